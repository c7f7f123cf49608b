#!/usr/bin/env python
"""Target for ncu: the per-gate kernels and the reductions at 28 qubits (one launch each after a
warm-up): RY, RZ, CNOT (unfused), expval(Z0), expval(X0 X1), probs(all)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pennylane_b200 import ops as q  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
sv = StateVector(n)
gates = [q.RY(0.3, wires=n // 2), q.RZ(0.7, wires=3), q.CNOT(wires=[5, n - 2])]
obs = [q.PauliZ(wires=0), q.PauliX(wires=0) @ q.PauliX(wires=1)]
for rep in range(2):
    for g in gates:
        sv.apply_operation(g)
    for o in obs:
        sv.expval_pauli_sentence(o.pauli_rep)
    p = sv.probs_device()
    del p
torch.cuda.synchronize()
print("done")
