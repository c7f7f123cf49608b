#!/usr/bin/env python
"""GPU micro-benchmark of the specialised segment kernels on the 30-qubit ansatz: per-segment
device time and GB/s (B200Q_TILE_L, B200Q_IO_LANES to vary the geometry).  Not part of bench.py."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200 import segjit  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    ops_ = bench.hea_ops(n)
    sv = StateVector(n)
    S2 = 2 * 16 * (1 << n)
    out = {"n": n, "L": os.environ.get("B200Q_TILE_L"), "jit": sv.jit_enabled(1)}
    t0 = time.time()
    segs = sv.compile_fused(ops_, level=1)
    sv.prepare_segments(segs)
    out["prepare_s"] = time.time() - t0
    sv.reset()
    for s in segs:
        sv.run_segment(s)
    torch.cuda.synchronize()
    out["first_pass_s"] = time.time() - t0
    out["jit"] = segjit.stats()
    for rep in range(2):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in segs]
        sv.reset()
        for s, (a, b) in zip(segs, ev):
            a.record(); sv.run_segment(s); b.record()
        torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    out["segments"] = len(segs)
    out["total_ms"] = sum(ms)
    out["gates_per_s"] = len(ops_) / (sum(ms) * 1e-3)
    out["gbps"] = S2 * len(segs) / (sum(ms) * 1e-3) / 1e9
    per = []
    for m, s in zip(ms, segs):
        plan = getattr(s, "_sk_plan", None)
        if plan is None:
            per.append((round(m, 2),))
            continue
        kinds = {}
        for r in plan.ir:
            kinds[r[0]] = kinds.get(r[0], 0) + 1
        per.append((round(m, 2), len(plan.rounds), kinds.get("dk", 0), kinds.get("cx", 0)))
    out["per_segment(ms,rounds,dk,cx)"] = per
    out["norm2"] = float(sv.norm2())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
