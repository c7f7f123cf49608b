"""CPU: the operator / tape / measurement mirror of the PennyLane data model."""
import numpy as np
import pytest
from scipy.linalg import expm

import pennylane_b200 as qb
from pennylane_b200 import ops as q
from pennylane_b200.pauli import PauliSentence, PauliWord, word_masks


def _prod(ops_, wires):
    m = np.eye(2 ** len(wires), dtype=complex)
    for o in ops_:
        m = q.matrix(o, wire_order=wires) @ m
    return m


@pytest.mark.parametrize("op", [q.Rot(.1, .2, .3, wires=0), q.U2(.4, .5, wires=0),
                                q.U3(.1, .2, .3, wires=0), q.CRot(.1, .2, .3, wires=[0, 1])],
                         ids=lambda o: o.name)
def test_decompositions_reproduce_matrix(op):
    assert np.allclose(_prod(op.decomposition(), op.wires), op.matrix())


@pytest.mark.parametrize("op", [
    q.RX(.3, wires=0), q.RY(.3, wires=0), q.RZ(.3, wires=0), q.PhaseShift(.3, wires=0),
    q.IsingXX(.3, wires=[0, 1]), q.IsingYY(.3, wires=[0, 1]), q.IsingZZ(.3, wires=[0, 1]),
    q.IsingXY(.3, wires=[0, 1]), q.CRX(.3, wires=[0, 1]), q.CRY(.3, wires=[0, 1]),
    q.CRZ(.3, wires=[0, 1]), q.ControlledPhaseShift(.3, wires=[0, 1]),
    q.SingleExcitation(.3, wires=[0, 1]), q.SingleExcitationMinus(.3, wires=[0, 1]),
    q.SingleExcitationPlus(.3, wires=[0, 1]), q.DoubleExcitation(.3, wires=[0, 1, 2, 3]),
    q.MultiRZ(.3, wires=[0, 1, 2]), q.PauliRot(.3, "XYZ", wires=[0, 1, 2]),
    q.GlobalPhase(.3, wires=[0]), q.ctrl(q.RX(.3, wires=2), [0, 1], [1, 0])], ids=lambda o: o.name)
def test_generators_exponentiate_to_matrix(op):
    """pennylane/operation.py:40-60: U = exp(i theta G)."""
    assert op.has_generator
    assert np.allclose(expm(1j * .3 * q.generator_matrix(op)), op.matrix())
    assert np.allclose(q.operation_derivative(op), 1j * q.generator_matrix(op) @ op.matrix())


def test_adjoint_rules():
    for op in [q.RX(.3, wires=0), q.Rot(.1, .2, .3, wires=0), q.S(wires=0), q.CNOT(wires=[0, 1]),
               q.IsingXY(.4, wires=[0, 1]), q.QubitUnitary(q.SX(wires=0).matrix(), wires=0),
               q.ctrl(q.RY(.2, wires=1), 0), q.MultiControlledX(wires=[0, 1, 2], control_values=[0, 1])]:
        adj = q.adjoint(op)
        assert np.allclose(adj.matrix() @ op.matrix(), np.eye(2 ** len(op.wires))), op


def test_broadcast_matrices_and_batch_size():
    th = np.array([.1, .2, .3])
    for op in [q.RX(th, wires=0), q.IsingXX(th, wires=[0, 1]), q.PauliRot(th, "XY", wires=[0, 1]),
               q.Rot(th, .1, th, wires=0), q.CRZ(th, wires=[0, 1])]:
        assert op.batch_size == 3
        m = op.matrix()
        assert m.shape[0] == 3
        for b in range(3):
            single = op._with_params([p[b] if np.ndim(p) else p for p in op.data])
            assert np.allclose(m[b], single.matrix())
    assert q.RX(.1, wires=0).batch_size is None
    with pytest.raises(ValueError, match="wrong number of wires"):
        q.CNOT(wires=[0])
    with pytest.raises(ValueError, match="not allowed|wrong length"):
        q.PauliRot(.1, "XQ", wires=[0, 1])


def test_expand_matrix_and_observable_arithmetic():
    H = q.LinearCombination([0.5, -1.5], [q.PauliZ(wires=1) @ q.PauliZ(wires=0), q.PauliX(wires=1)])
    Z, X, I = np.diag([1., -1.]), np.array([[0, 1.], [1, 0]]), np.eye(2)
    assert np.allclose(H.matrix(wire_order=[0, 1]), 0.5 * np.kron(Z, Z) - 1.5 * np.kron(I, X))
    assert np.allclose(H.matrix(wire_order=[1, 0]), 0.5 * np.kron(Z, Z) - 1.5 * np.kron(X, I))
    ps = H.pauli_rep
    assert ps[PauliWord({0: "Z", 1: "Z"})] == 0.5 and ps[PauliWord({1: "X"})] == -1.5
    assert (q.PauliX(wires=0) @ q.PauliY(wires=0)).pauli_rep == PauliSentence({PauliWord({0: "Z"}): 1j})
    assert word_masks(PauliWord({0: "X", 2: "Y", 3: "Z"}), {0: 3, 2: 1, 3: 0}) == (0b1010, 0b0011, 1)
    assert q.Hermitian(np.eye(2), wires=0).pauli_rep is None
    with pytest.raises(ValueError, match="Hermitian"):
        q.Hermitian(np.array([[0, 1], [2, 0]]), wires=0)


def test_diagonalizing_gates_diagonalise():
    A = np.array([[1, 1 - 2j], [1 + 2j, -0.5]])
    for obs in [q.PauliX(wires=0), q.PauliY(wires=0), q.PauliZ(wires=0), q.Hadamard(wires=0),
                q.Hermitian(A, wires=0), q.PauliX(wires=0) @ q.PauliY(wires=1),
                q.Projector(np.array([1, 0]), wires=[0, 1]),
                q.Projector(np.array([1, 1j, 0, -1]) / np.sqrt(3), wires=[0, 1])]:
        U = _prod(obs.diagonalizing_gates(), obs.wires)
        D = U @ obs.matrix(wire_order=obs.wires) @ U.conj().T
        assert np.allclose(D, np.diag(np.asarray(obs.eigvals(), dtype=complex)), atol=1e-12), obs


def test_state_prep_vectors():
    st = q.StatePrep(np.array([0, 1, 0, 0]), wires=[1, 0])
    v = st.state_vector(wire_order=[0, 1, 2]).reshape(-1)
    assert v[0b100] == 1 and abs(v).sum() == 1
    b = q.BasisState(np.array([1, 0, 1]), wires=[2, 0, 1])
    assert b.state_vector(wire_order=[0, 1, 2]).reshape(-1)[0b011] == 1
    with pytest.raises(ValueError, match="norm"):
        q.StatePrep(np.array([1, 1]), wires=0)


def test_shots_and_quantum_script():
    s = qb.Shots([10, (5, 2)])
    assert s.total_shots == 20 and s.has_partitioned_shots and s.num_copies == 3
    assert list(s.bins()) == [(0, 10), (10, 15), (15, 20)] and list(s) == [10, 5, 5]
    assert not qb.Shots(None) and not qb.Shots(7).has_partitioned_shots
    with pytest.raises(ValueError):
        qb.Shots(0)
    t = qb.QuantumScript([q.RX(.1, wires="a"), q.CNOT(wires=["a", 3]), q.Rot(.1, .2, .3, wires=3)],
                         [qb.expval(q.PauliZ(wires="c")), qb.probs(wires=[3])])
    assert t.wires == ("a", 3, "c") and t.num_wires == 3 and t.trainable_params == [0, 1, 2, 3]
    m = t.map_to_standard_wires()
    assert [o.wires for o in m.operations] == [(0,), (0, 1), (1,)] and m.measurements[0].wires == (2,)
    std = qb.QuantumScript([q.RX(.1, wires=0)], [qb.expval(q.PauliZ(wires=1))])
    assert std.map_to_standard_wires() is std
    assert t.hash == t.copy().hash != std.hash
    t.trainable_params = [0, 2]
    assert t.get_parameters() == [.1, .2]
    bt = qb.QuantumScript([q.RX(np.array([.1, .2]), wires=0), q.RY(.3, wires=0)], [qb.state()])
    assert bt.batch_size == 2


def test_measurement_process_samples():
    samples = np.array([[0, 0], [0, 1], [1, 1], [1, 1]])
    assert qb.expval(q.PauliZ(wires=1)).process_samples(samples, [0, 1]) == -0.5
    assert np.allclose(qb.probs(wires=[0, 1]).process_samples(samples, [0, 1]), [.25, .25, 0, .5])
    assert qb.counts(wires=[1]).process_samples(samples, [0, 1]) == {"0": 1, "1": 3}
    assert qb.counts(wires=[0, 1], all_outcomes=True).process_samples(samples, [0, 1]) == {
        "00": 1, "01": 1, "10": 0, "11": 2}
    assert np.array_equal(qb.sample(wires=[1, 0]).process_samples(samples, [0, 1]), samples[:, ::-1])
    assert np.array_equal(qb.sample(q.PauliZ(wires=0)).process_samples(samples, [0, 1]), [1, 1, -1, -1])
    assert qb.var(q.PauliZ(wires=0)).process_samples(samples, [0, 1]) == 1.0
