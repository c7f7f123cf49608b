#!/usr/bin/env python
"""GPU micro-benchmark of the fused adjoint (execute_and_compute_derivatives of the ansatz):
seconds per call and per-segment device times of the reverse sweep.  Not part of bench.py."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200 import adjoint, segjit  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    tape = bench.hea_tape(n, 8)
    out = {"n": n, "geom": os.environ.get("B200Q_SK_ADJGEOM"), "L": os.environ.get("B200Q_TILE_L")}
    ts = []
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        jac = adjoint.adjoint_jacobian(tape, fusion=1)
        torch.cuda.synchronize()
        ts.append(time.time() - t0)
    out["seconds"] = ts
    out["jit"] = segjit.stats()
    j = np.array(jac, dtype=float)
    out["jac_norm"] = float(np.linalg.norm(j))
    out["jac_head"] = [float(x) for x in j[:4]]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
