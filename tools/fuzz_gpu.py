#!/usr/bin/env python
"""Differential fuzz on the GPU: random circuits (every primitive kind of tests/test_compiler.py's
generator, plus broadcast rotations) through the fused path at several sizes, fusion levels and
both precisions, against the oracle.  Prints one summary line; exit code 1 on any mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pennylane_b200 as qb  # noqa: E402
from pennylane_b200 import StateVector, ops as q  # noqa: E402
from oracle.apply_operation import apply_operation as o_apply  # noqa: E402
from conftest import random_state  # noqa: E402
from test_compiler import _random_circuit, _trainable_circuit  # noqa: E402


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    bad, worst, runs = [], {np.dtype(np.complex128): 0.0, np.dtype(np.complex64): 0.0}, 0
    for seed in range(seeds):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(12, 18))
        dtype = np.complex128 if seed % 3 else np.complex64
        level = seed % 3
        L = int(rng.integers(3, 7))
        ops_ = _random_circuit(n, 120, seed=5000 + seed)
        if seed % 4 == 0:                                  # sprinkle broadcast rotations
            B = 2 + seed % 2
            for i in range(0, len(ops_), 7):
                w = int(rng.integers(n))
                ops_.insert(i, [q.RX, q.RY, q.RZ][i % 3](rng.uniform(0, 6, B), wires=w))
        state = random_state(n, seed=seed)
        ref = state
        for op in ops_:
            ref = o_apply(op, ref, is_state_batched=ref.ndim > n)
        sv = StateVector(n, dtype=dtype)
        T = sv.rt_geometry(1)[0]
        if n < T:
            continue
        sv.set_state(state.astype(dtype))
        sv.apply_operations_fused(ops_, level=level, T=T, L=min(L, T))
        got = sv.to_numpy()
        err = float(np.max(np.abs(got.reshape(ref.shape) - ref)))
        tol = 1e-12 if dtype == np.complex128 else 3e-5
        worst[np.dtype(dtype)] = max(worst[np.dtype(dtype)], err)
        runs += 1
        if not err < tol:
            bad.append((seed, n, np.dtype(dtype).name, level, L, err))
    # adjoint Jacobians of random trainable circuits through the device
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    jworst = 0.0
    for seed in range(max(4, seeds // 4)):
        n = 12 + seed % 4
        ops_ = _trainable_circuit(n, 50, seed=700 + seed)
        tape = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0) @ q.PauliX(wires=2)), qb.expval(q.PauliY(wires=1))])
        dev = qb.B200Qubit(wires=n, fusion=1)
        res, jac = dev.execute_and_compute_derivatives(tape)
        st, _ = o_sim.get_final_state(tape)
        rj = np.array(o_adj.adjoint_jacobian(tape, st), dtype=float)
        e = float(np.max(np.abs(np.array(jac, dtype=float) - rj)))
        jworst = max(jworst, e)
        runs += 1
        if not e < 1e-12:
            bad.append((seed, n, "adjoint", e))
    print({"runs": runs, "worst_c128": worst[np.dtype(np.complex128)], "worst_c64": worst[np.dtype(np.complex64)],
           "worst_jacobian": jworst, "failures": bad})
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
