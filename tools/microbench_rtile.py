#!/usr/bin/env python
"""GPU micro-benchmark of the fused segment kernels on the 30-qubit ansatz: per-segment device
time and GB/s for several contiguous-run lengths L (B200Q_TILE_L).  Not part of bench.py."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200.compiler import compile_ops, schedule_rounds  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    ops_ = bench.hea_ops(n)
    sv = StateVector(n)
    T, RB, _ = sv.rt_geometry(1)
    S2 = 2 * 16 * (1 << n)
    out = {}
    for L in (5, 4, 3, 6):
        for level in (1,):
            segs = compile_ops(ops_, n, level=level, T=T, L=L)
            sv.reset()
            for s in segs:
                sv.run_segment(s)          # warm-up + encode cache
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in segs]
            sv.reset()
            for s, (a, b) in zip(segs, ev):
                a.record(); sv.run_segment(s); b.record()
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in ev]
            rounds = [len(schedule_rounds(s.prims, s.tile_bits, RB)) for s in segs]
            out[f"L{L}_lvl{level}"] = {
                "segments": len(segs), "total_ms": sum(ms), "gates_per_s": len(ops_) / (sum(ms) * 1e-3),
                "gbps": S2 * len(segs) / (sum(ms) * 1e-3) / 1e9,
                "per_segment": [(round(m, 2), r, len(s.prims)) for m, r, s in zip(ms, rounds, segs)]}
    # empty program (one IO round, no gates): the copy ceiling of the kernel
    from pennylane_b200.compiler import Segment
    for L in (5, 4, 3):
        seg = Segment(list(range(L)) + list(range(n - (T - L), n)), [], 0)
        sv.run_segment(seg)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            sv.run_segment(seg)
        b.record(); torch.cuda.synchronize()
        out[f"empty_L{L}_gbps"] = S2 * 5 / (a.elapsed_time(b) * 1e-3) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
