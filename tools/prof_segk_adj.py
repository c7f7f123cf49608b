#!/usr/bin/env python
"""Target for `ncu --profile-from-start off -k regex:sk_kernel`: one warm call of the fused
adjoint of the ansatz, then a second call inside cudaProfilerStart/Stop (the reverse-sweep
launches are the NV = 2 instances of sk_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200 import adjoint  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
tape = bench.hea_tape(n, 8)
adjoint.adjoint_jacobian(tape, fusion=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
adjoint.adjoint_jacobian(tape, fusion=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
