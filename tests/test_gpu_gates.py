"""GPU parity: every gate kernel against the oracle's restatement of apply_operation.py.

Mirrors tests/devices/qubit/test_apply_operation.py of the reference: each operation is applied
to a random state on every wire placement and compared with the dense reference
(:1028-1083 "broadcast vs Kronecker reference", :255-445 fixed-state cases).
Tolerance: 1e-12 (complex128), 1e-5 (complex64) on unit-norm states, as north_star states.
"""
import itertools

import numpy as np
import pytest

from conftest import TOL, random_state

pytestmark = pytest.mark.gpu


def _sv(state, dtype=np.complex128):
    from pennylane_b200 import StateVector

    n = state.ndim if state.shape[0] == 2 and all(s == 2 for s in state.shape) else state.ndim - 1
    sv = StateVector(n, dtype=dtype)
    sv.set_state(state)
    return sv


def _check(op, state, batched=False, dtype=np.complex128):
    from oracle.apply_operation import apply_operation

    from pennylane_b200 import StateVector

    n = state.ndim - (1 if batched else 0)
    sv = StateVector(n, dtype=dtype)
    sv.set_state(state.astype(dtype))
    sv.apply_operation(op)
    got = sv.to_numpy()
    ref = apply_operation(op, state.astype(np.complex128), is_state_batched=batched)
    assert got.shape == ref.shape, (op, got.shape, ref.shape)
    err = np.max(np.abs(got - ref))
    assert err < TOL[np.dtype(dtype)], f"{op}: max abs err {err:.3e}"


def _one_qubit_ops(q, w):
    return [q.PauliX(wires=w), q.PauliY(wires=w), q.PauliZ(wires=w), q.Hadamard(wires=w),
            q.S(wires=w), q.T(wires=w), q.SX(wires=w), q.RX(0.432, wires=w), q.RY(-1.2, wires=w),
            q.RZ(2.1, wires=w), q.PhaseShift(0.77, wires=w), q.Rot(0.1, 0.2, 0.3, wires=w),
            q.U2(0.3, -0.4, wires=w), q.U3(0.5, 0.6, 0.7, wires=w), q.Identity(wires=w),
            q.GlobalPhase(0.31, wires=w), q.adjoint(q.S(wires=w)), q.adjoint(q.T(wires=w)),
            q.adjoint(q.SX(wires=w))]


@pytest.mark.parametrize("n", [1, 3, 6, 9])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_single_qubit_gates_every_wire(n, dtype):
    from pennylane_b200 import ops as q

    state = random_state(n, seed=n)
    for w in range(n):
        for op in _one_qubit_ops(q, w):
            _check(op, state, dtype=dtype)


def _two_qubit_ops(q, a, b):
    return [q.CNOT(wires=[a, b]), q.CZ(wires=[a, b]), q.CY(wires=[a, b]), q.CH(wires=[a, b]),
            q.SWAP(wires=[a, b]), q.ISWAP(wires=[a, b]), q.SISWAP(wires=[a, b]),
            q.ECR(wires=[a, b]), q.CRX(0.3, wires=[a, b]), q.CRY(0.4, wires=[a, b]),
            q.CRZ(0.5, wires=[a, b]), q.CRot(0.1, 0.2, 0.3, wires=[a, b]),
            q.ControlledPhaseShift(0.9, wires=[a, b]), q.IsingXX(0.3, wires=[a, b]),
            q.IsingYY(0.4, wires=[a, b]), q.IsingZZ(0.5, wires=[a, b]),
            q.IsingXY(0.6, wires=[a, b]), q.PSWAP(0.7, wires=[a, b]),
            q.SingleExcitation(0.8, wires=[a, b]), q.SingleExcitationPlus(0.8, wires=[a, b]),
            q.SingleExcitationMinus(0.8, wires=[a, b]), q.MultiRZ(0.45, wires=[a, b]),
            q.PauliRot(0.3, "XY", wires=[a, b]), q.PauliRot(0.3, "ZZ", wires=[a, b]),
            q.PauliRot(0.3, "IY", wires=[a, b]), q.PauliRot(0.3, "ZX", wires=[a, b]),
            q.adjoint(q.ISWAP(wires=[a, b])), q.ctrl(q.RX(0.2, wires=b), a),
            q.ctrl(q.RY(0.2, wires=b), a, control_values=[0])]


@pytest.mark.parametrize("n", [2, 4, 7])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_two_qubit_gates_every_pair(n, dtype):
    from pennylane_b200 import ops as q

    state = random_state(n, seed=10 + n)
    for a, b in itertools.permutations(range(n), 2):
        for op in _two_qubit_ops(q, a, b):
            _check(op, state, dtype=dtype)


@pytest.mark.parametrize("n", [3, 5, 8])
def test_three_and_four_qubit_gates(n):
    from pennylane_b200 import ops as q

    rng = np.random.default_rng(n)
    state = random_state(n, seed=20 + n)
    for ws in itertools.islice(itertools.permutations(range(n), 3), 0, 40):
        ws = list(ws)
        U = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]
        for op in [q.Toffoli(wires=ws), q.CSWAP(wires=ws), q.CCZ(wires=ws),
                   q.MultiControlledX(wires=ws, control_values=[1, 0]),
                   q.MultiRZ(0.3, wires=ws), q.PauliRot(0.7, "XYZ", wires=ws),
                   q.PauliRot(0.7, "YIX", wires=ws), q.QubitUnitary(U, wires=ws),
                   q.ctrl(q.IsingXY(0.4, wires=ws[1:]), ws[0]),
                   q.ctrl(q.PhaseShift(0.4, wires=ws[2]), ws[:2], control_values=[0, 1]),
                   q.DiagonalQubitUnitary(np.exp(1j * rng.normal(size=8)), wires=ws)]:
            _check(op, state)
    if n >= 4:
        for ws in itertools.islice(itertools.permutations(range(n), 4), 0, 30):
            ws = list(ws)
            U = np.linalg.qr(rng.normal(size=(16, 16)) + 1j * rng.normal(size=(16, 16)))[0]
            for op in [q.DoubleExcitation(0.5, wires=ws), q.QubitUnitary(U, wires=ws),
                       q.MultiControlledX(wires=ws), q.PauliRot(0.2, "XZYX", wires=ws),
                       q.ctrl(q.SWAP(wires=ws[2:]), ws[:2])]:
                _check(op, state)


@pytest.mark.parametrize("k", [5, 6, 8, 10])
def test_large_dense_unitary(k):
    """apply_operation.py:202-255 tensordot path for wide QubitUnitary."""
    from pennylane_b200 import ops as q

    n = 12
    rng = np.random.default_rng(k)
    U = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))[0]
    ws = list(rng.permutation(n)[:k])
    _check(q.QubitUnitary(U, wires=[int(w) for w in ws]), random_state(n, seed=k))


def test_multicontrolledx_wide():
    """apply_operation.py:782-832: the >= 9-wire matrix-free MultiControlledX path."""
    from pennylane_b200 import ops as q

    n = 11
    state = random_state(n, seed=3)
    cv = [1, 0, 1, 1, 0, 1, 1, 0, 1, 1]
    _check(q.MultiControlledX(wires=[3, 0, 9, 1, 7, 4, 10, 2, 8, 6, 5], control_values=cv), state)
    _check(q.MultiControlledX(wires=list(range(n))), state)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_parameter_broadcasting(dtype):
    """apply_operation.py:686-697 / :191-197 — a batched parameter gives the state a batch axis;
    reference tests: test_apply_operation.py:1028-1083."""
    from pennylane_b200 import ops as q

    n, B = 5, 3
    th = np.array([0.1, -0.7, 2.3])
    single = random_state(n, seed=1)
    batched = random_state(n, seed=2, batch=B)
    for w in range(n):
        for op in [q.RX(th, wires=w), q.RY(th, wires=w), q.RZ(th, wires=w),
                   q.PhaseShift(th, wires=w), q.Rot(th, 0.2, th, wires=w),
                   q.GlobalPhase(th, wires=w)]:
            _check(op, single, dtype=dtype)
            _check(op, batched, batched=True, dtype=dtype)
    for a, b in [(0, 1), (3, 1), (4, 0)]:
        for op in [q.IsingXX(th, wires=[a, b]), q.IsingZZ(th, wires=[a, b]),
                   q.CRX(th, wires=[a, b]), q.ControlledPhaseShift(th, wires=[a, b]),
                   q.PauliRot(th, "XY", wires=[a, b]), q.MultiRZ(th, wires=[a, b]),
                   q.SingleExcitation(th, wires=[a, b])]:
            _check(op, single, dtype=dtype)
            _check(op, batched, batched=True, dtype=dtype)
    # unbatched gate on a batched state
    _check(q.Hadamard(wires=2), batched, batched=True, dtype=dtype)
    _check(q.CNOT(wires=[1, 3]), batched, batched=True, dtype=dtype)
    _check(q.Toffoli(wires=[1, 3, 0]), batched, batched=True, dtype=dtype)


@pytest.mark.parametrize("n", [14, 20])
def test_larger_states_high_and_low_wires(n):
    """n >= 13 takes the reference's tensordot branch (apply_operation.py:29-30,346-351)."""
    from pennylane_b200 import ops as q

    state = random_state(n, seed=n)
    for w in [0, 1, n // 2, n - 2, n - 1]:
        _check(q.RY(0.3, wires=w), state)
        _check(q.RZ(0.3, wires=w), state)
        _check(q.Hadamard(wires=w), state)
    for a, b in [(0, n - 1), (n - 1, 0), (n - 2, n - 1), (0, 1), (n // 2, 2)]:
        _check(q.CNOT(wires=[a, b]), state)
        _check(q.IsingXX(0.4, wires=[a, b]), state)
        _check(q.CZ(wires=[a, b]), state)
    _check(q.Toffoli(wires=[n - 1, 0, n // 2]), state)


def test_gate_sequence_matches_oracle_circuit():
    """simulate.py:214-235 gate loop on a 12-qubit layered circuit."""
    from oracle import simulate as o_sim

    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.simulate import get_final_state

    n = 12
    rng = np.random.default_rng(0)
    ops_ = []
    for layer in range(3):
        for i in range(n):
            ops_.append(q.Rot(*rng.uniform(0, 2 * np.pi, 3), wires=i))
        for i in range(n):
            ops_.append(q.CNOT(wires=[i, (i + layer + 1) % n]))
    tape = qb.QuantumScript(ops_, [qb.state()])
    sv, _ = get_final_state(tape)
    ref, _ = o_sim.get_final_state(tape)
    assert np.max(np.abs(sv.to_numpy() - ref)) < 1e-12


@pytest.mark.parametrize("fusion", [0, 1])
def test_grover_operator_matrix_and_decomposed_widths(fusion):
    """GroverOperator (apply_operation.py:836-880): below nine wires through its matrix, from nine
    wires on matrix-free — here as H^k (2|0><0| - I) H^k — against the oracle's restatement of the
    reference kernel."""
    import pennylane_b200 as qb
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q

    n = 13
    rng = np.random.default_rng(17)
    prep = [q.RY(rng.uniform(0, 6), wires=i) for i in range(n)] + \
           [q.CNOT(wires=[i, (i + 1) % n]) for i in range(n)] + [q.RZ(rng.uniform(0, 6), wires=i) for i in range(n)]
    for wires in ([0, 1], [4, 2, 9], list(range(9)), [12, 0, 3, 4, 5, 6, 7, 8, 9, 11], list(range(n))):
        tape = qb.QuantumScript(prep + [q.GroverOperator(wires=wires), q.RX(0.3, wires=wires[0])], [qb.state()])
        dev = qb.B200Qubit(wires=n, fusion=fusion)
        (ptape,), _ = dev.preprocess(tape)
        got = dev.execute(ptape)
        ref = np.asarray(o_sim.simulate(tape)).reshape(-1)
        assert np.max(np.abs(np.asarray(got).reshape(-1) - ref)) < 1e-12, wires
