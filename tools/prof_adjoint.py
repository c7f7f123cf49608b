#!/usr/bin/env python
"""Target for ncu: one adjoint Jacobian of the hardware-efficient ansatz (default 26 qubits)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import pennylane_b200 as qb  # noqa: E402
from pennylane_b200 import ops as q  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tape = qb.QuantumScript(bench.hea_ops(n, layers), [qb.expval(q.PauliZ(wires=0))])
dev = qb.B200Qubit(wires=n, fusion=1)
res, jac = dev.execute_and_compute_derivatives(tape)
print("expval", float(res), "grad norm", float(np.linalg.norm(np.asarray(jac, dtype=float))))
