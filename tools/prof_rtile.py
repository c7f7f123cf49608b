#!/usr/bin/env python
"""Target for `ncu -k regex:k_rtile`: runs one mid-circuit fused segment of the ansatz (and an
empty one) a few times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200.compiler import Segment, compile_ops  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
which = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 3]
sv = StateVector(n)
T, RB, _ = sv.rt_geometry(1)
segs = compile_ops(bench.hea_ops(n), n, level=1, T=T, L=5)
for i in which:
    for _ in range(2):
        sv.run_segment(segs[i])
empty = Segment(list(range(5)) + list(range(8, 8 + T - 5)), [], 0)
for _ in range(2):
    sv.run_segment(empty)
torch.cuda.synchronize()
print("segments:", [(i, len(segs[i].prims)) for i in which])
