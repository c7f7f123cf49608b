"""CPU: the oracle's apply_operation (reference numpy kernels, independent gate matrices) against
a brute-force Kronecker expansion of the package's operator matrices — the differential check of
tests/devices/qubit/test_apply_operation.py:1028-1083 in the reference.  Two independently
written sets of gate matrices (oracle/gates.py via expm of generators, pennylane_b200/ops.py
closed forms) must agree on every gate."""
import itertools

import numpy as np
import pytest

from conftest import random_state
from oracle.apply_operation import apply_operation
from oracle.gates import matrix_of
from pennylane_b200 import ops as q


def all_ops(n, rng):
    w = [int(x) for x in rng.permutation(n)]
    a, b = w[0], w[1 % n]
    out = [q.PauliX(wires=a), q.PauliY(wires=a), q.PauliZ(wires=a), q.Hadamard(wires=a),
           q.S(wires=a), q.T(wires=a), q.SX(wires=a), q.RX(0.432, wires=a), q.RY(-1.2, wires=a),
           q.RZ(2.1, wires=a), q.PhaseShift(0.77, wires=a), q.U1(0.3, wires=a),
           q.Rot(0.1, 0.2, 0.3, wires=a), q.U2(0.3, -0.4, wires=a), q.U3(0.5, 0.6, 0.7, wires=a),
           q.GlobalPhase(0.31, wires=a), q.adjoint(q.S(wires=a)), q.adjoint(q.RX(0.3, wires=a))]
    if n >= 2:
        out += [q.CNOT(wires=[a, b]), q.CZ(wires=[a, b]), q.CY(wires=[a, b]), q.CH(wires=[a, b]),
                q.SWAP(wires=[a, b]), q.ISWAP(wires=[a, b]), q.SISWAP(wires=[a, b]),
                q.ECR(wires=[a, b]), q.CRX(0.3, wires=[a, b]), q.CRY(0.4, wires=[a, b]),
                q.CRZ(0.5, wires=[a, b]), q.CRot(0.1, 0.2, 0.3, wires=[a, b]),
                q.ControlledPhaseShift(0.9, wires=[a, b]), q.IsingXX(0.3, wires=[a, b]),
                q.IsingYY(0.4, wires=[a, b]), q.IsingZZ(0.5, wires=[a, b]),
                q.IsingXY(0.6, wires=[a, b]), q.PSWAP(0.7, wires=[a, b]),
                q.SingleExcitation(0.8, wires=[a, b]), q.SingleExcitationPlus(0.8, wires=[a, b]),
                q.SingleExcitationMinus(0.8, wires=[a, b]), q.MultiRZ(0.45, wires=[a, b]),
                q.PauliRot(0.3, "XY", wires=[a, b]), q.ctrl(q.RY(0.2, wires=b), a, [0]),
                q.adjoint(q.ISWAP(wires=[a, b]))]
    if n >= 3:
        c = w[2]
        out += [q.Toffoli(wires=[a, b, c]), q.CSWAP(wires=[a, b, c]), q.CCZ(wires=[a, b, c]),
                q.MultiControlledX(wires=[a, b, c], control_values=[0, 1]),
                q.PauliRot(0.7, "YIX", wires=[a, b, c]), q.MultiRZ(0.2, wires=[a, b, c])]
    if n >= 4:
        out += [q.DoubleExcitation(0.5, wires=w[:4])]
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_oracle_matches_kronecker_expansion(n):
    rng = np.random.default_rng(n)
    state = random_state(n, seed=n)
    for op in all_ops(n, rng):
        ref = (q.matrix(op, wire_order=range(n)) @ state.reshape(-1)).reshape(state.shape)
        got = apply_operation(op, state)
        assert np.allclose(got, ref, atol=1e-13), op


def test_oracle_gate_matrices_match_package():
    rng = np.random.default_rng(0)
    for op in all_ops(5, rng):
        if op.name in ("GlobalPhase",):
            continue
        assert np.allclose(matrix_of(op), op.matrix(), atol=1e-14), op.name


def test_oracle_tensordot_branch_and_batching():
    n = 13      # state.ndim >= 13 -> tensordot (apply_operation.py:29-30)
    state = random_state(n, seed=1)
    for op in [q.RY(0.3, wires=12), q.CNOT(wires=[12, 0]), q.IsingXX(0.2, wires=[3, 9]),
               q.Hadamard(wires=5), q.Toffoli(wires=[1, 11, 6])]:
        ref = (q.matrix(op, wire_order=range(n)) @ state.reshape(-1)).reshape(state.shape)
        assert np.allclose(apply_operation(op, state), ref, atol=1e-13)
    th = np.array([0.1, 0.2, 0.3])
    st = random_state(4, seed=2)
    for op in [q.RX(th, wires=2), q.IsingZZ(th, wires=[0, 3]), q.PhaseShift(th, wires=1),
               q.GlobalPhase(th, wires=0)]:
        got = apply_operation(op, st)
        assert got.shape == (3,) + st.shape
        for b in range(3):
            single = op._with_params([th[b]])
            assert np.allclose(got[b], apply_operation(single, st), atol=1e-14)
    bst = random_state(4, seed=3, batch=3)
    got = apply_operation(q.RX(th, wires=1), bst, is_state_batched=True)
    for b in range(3):
        assert np.allclose(got[b], apply_operation(q.RX(th[b], wires=1), bst[b]), atol=1e-14)


def test_oracle_multicontrolledx_wide_path():
    n = 10
    state = random_state(n, seed=4)
    cv = [1, 0, 1, 1, 0, 1, 1, 0, 1]
    op = q.MultiControlledX(wires=[3, 0, 9, 1, 7, 4, 2, 8, 6, 5], control_values=cv)
    ref = (q.matrix(op, wire_order=range(n)) @ state.reshape(-1)).reshape(state.shape)
    assert np.allclose(apply_operation(op, state), ref, atol=1e-14)


# ---- GroverOperator: apply_operation.py:836-880 ----------------------------------------------------
def test_grover_known_answer_and_matrix_free_branch():
    """test_apply_operation.py:391-412 (two-qubit known answer) and :1168-1300 (the matrix-free
    kernel on >= 9 wires against the dense matrix; full and partial wire sets, batched)."""
    import numpy as np
    from oracle.apply_operation import apply_operation
    from pennylane_b200 import ops as q

    initial = np.array([[0.04624539 + 0.3895457j, 0.22399401 + 0.53870339j],
                        [-0.483054 + 0.2468498j, -0.02772249 - 0.45901669j]])
    for wire in (0, 1):
        op = q.GroverOperator(wires=[wire, 1 - wire])
        new = apply_operation(op, initial)
        expected = 2 * (np.ones_like(initial) / 2) * (initial.sum() / 2) - initial
        assert np.allclose(new, expected)
    rng = np.random.default_rng(3)
    for op_wires, n, batch in ((list(range(9)), 9, None), ([10, 0, 3, 4, 5, 6, 7, 8, 9], 11, None),
                               (list(range(9)), 10, 2)):
        shape = ([batch] if batch else []) + [2] * n
        state = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        op = q.GroverOperator(wires=op_wires)
        got = apply_operation(op, state, is_state_batched=bool(batch))
        k = len(op_wires)
        # dense reference: 2|s><s| - I on the operator's axes
        axes = [w + bool(batch) for w in op_wires]
        mean = state.sum(axis=tuple(axes), keepdims=True) / (1 << k)
        assert np.allclose(got, 2 * mean - state)
    # the decomposition the CUDA device uses for wide operators equals the matrix
    op = q.GroverOperator(wires=[2, 0, 1])
    state = rng.normal(size=[2] * 3) + 1j * rng.normal(size=[2] * 3)
    ref = apply_operation(op, state)
    dec = state
    for sub in op.decomposition():
        dec = apply_operation(sub, dec)
    assert np.allclose(dec, ref, atol=1e-14)


# ---- SparseHamiltonian: measure.py:74-118 (scipy branch) ------------------------------------------
def test_sparse_hamiltonian_expval_oracle_and_wire_expansion():
    """test_measure.py:125-190 measures the same Hamiltonian as LinearCombination, SparseHamiltonian
    and Hermitian and expects one value; here: the oracle's CSR branch against the dense
    contraction, with the observable on a permuted subset of the wires and on a batch."""
    import numpy as np
    import scipy.sparse as sp
    from types import SimpleNamespace
    from oracle import measure as o_meas
    from pennylane_b200 import ops as q
    from pennylane_b200.ops import expand_matrix

    rng = np.random.default_rng(12)
    n, wires = 6, [4, 1, 3]
    A = sp.random(8, 8, density=0.4, random_state=3) + 1j * sp.random(8, 8, density=0.4, random_state=4)
    H = (A + A.conj().T).tocsr()
    obs = q.SparseHamiltonian(H, wires=wires)
    dense_full = expand_matrix(H.toarray(), wires, list(range(n)))
    assert np.allclose(obs.sparse_matrix(wire_order=list(range(n))).toarray(), dense_full)
    for batch in (None, 3):
        shape = ([batch] if batch else []) + [2] * n
        st = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        st /= np.sqrt((np.abs(st) ** 2).sum(axis=tuple(range(int(bool(batch)), st.ndim)), keepdims=True))
        mp = SimpleNamespace(kind="expval", obs=obs, wires=obs.wires)
        got = o_meas.measure(mp, st, bool(batch))
        flat = st.reshape(batch or 1, -1)
        ref = np.real(np.einsum("bi,ij,bj->b", flat.conj(), dense_full, flat))
        assert np.allclose(got, ref if batch else ref[0], atol=1e-13)
