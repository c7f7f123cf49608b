#!/usr/bin/env python
"""Target for ncu: the two sweeps of a native MidMeasure at N qubits (default 28), complex128 —
the single-wire marginal (k_probs_marginal + split sum) and k_collapse — on a high, a middle and
the lowest state bit, then the reduced-density-matrix kernel (k_gram_block) on two and three
kept wires; one launch each after a warm-up."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pennylane_b200 import ops as q  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
sv = StateVector(n)
for w in range(0, n, 3):
    sv.apply_operation(q.Hadamard(w))
for rep in range(2):
    for wire in (0, n // 2, n - 1):
        p = sv.probs_device([wire])
        sv.collapse(wire, 0, False, 1.0)
        del p
    sv.reduced_dm([0, n - 1])          # k_gram_block<MI=2, same>
    sv.reduced_dm([3, 0, n - 1])       # + the off-diagonal block variant
torch.cuda.synchronize()
print("done")
