#!/usr/bin/env python
"""Target for `ncu -k regex:sk_kernel`: runs chosen fused segments of the ansatz through the
specialised kernels (one warm-up + one profiled launch each)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
which = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 3]
sv = StateVector(n)
segs = sv.compile_fused(bench.hea_ops(n), level=1)
sv.prepare_segments(segs)
for i in which:
    for _ in range(2):
        sv.run_segment(segs[i])
torch.cuda.synchronize()
for i in which:
    p = segs[i]._sk_plan
    print("segment", i, "rounds", len(p.rounds), {k: sum(r[0] == k for r in p.ir) for k in ("dk", "cx")})
