"""GPU parity of ``b200q_gram_block`` / ``StateVector.reduced_dm`` and the measurements built on
it (density_matrix, purity, vn_entropy, mutual_info) against the oracle
(math/quantum.py:386-487) at 1e-12 (complex128) / 1e-5 (complex64), the reference's closed
forms, and — at sizes the reference's 4^n route cannot reach — trace / Hermiticity / purity
properties."""
import numpy as np
import pytest

from conftest import TOL, random_state

pytestmark = pytest.mark.gpu


def _sv(state, dtype=np.complex128, batched=False):
    from pennylane_b200 import StateVector

    n = state.ndim - (1 if batched else 0)
    sv = StateVector(n, dtype=dtype, batch=state.shape[0] if batched else 1)
    sv.set_state(state.astype(dtype))
    return sv


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n, wires", [
    (1, [0]), (2, [1]), (2, [1, 0]), (5, [0]), (5, [4]), (5, [2, 0]), (5, [4, 0, 2]),
    (9, [8, 0, 4, 3]), (9, [1, 2, 3, 4, 5]), (12, [11, 0]), (12, [3, 10, 7, 1, 5, 0]),
    (14, [13]), (14, [6, 7]), (6, [0, 1, 2, 3, 4, 5])])
def test_reduced_dm_matches_oracle(n, wires, dtype):
    from oracle.measure import reduce_statevector

    state = random_state(n, seed=n + len(wires), dtype=dtype)
    got = _sv(state, dtype).reduced_dm(wires)
    ref = reduce_statevector(state.reshape(-1).astype(np.complex128), wires)
    assert got.shape == ref.shape == (2 ** len(wires),) * 2
    assert np.max(np.abs(got - ref)) < TOL[np.dtype(dtype)]
    assert np.array_equal(got, got.conj().T)


def test_batched_state():
    from oracle.measure import reduce_statevector

    state = random_state(7, seed=3, batch=3)
    got = _sv(state, batched=True).reduced_dm([5, 1, 6])
    ref = reduce_statevector(state.reshape(3, -1), [5, 1, 6])
    assert got.shape == (3, 8, 8) and np.max(np.abs(got - ref)) < 1e-12


@pytest.mark.parametrize("param", np.linspace(0.05, 2 * np.pi - 0.05, 5))
def test_closed_forms_through_the_device(param):
    """tests/measurements/test_purity_measurement.py:25-31, test_vn_entropy.py:29-41,
    test_mutual_info.py:186-203."""
    import pennylane_b200 as pb
    from pennylane_b200 import QuantumScript, measurements as M, ops

    r = np.sqrt(1 - 4 * np.cos(param / 2) ** 2 * np.sin(param / 2) ** 2)
    eigs = np.array([e for e in ((1 + r) / 2, (1 - r) / 2) if e > 0])
    dev = pb.device("b200.qubit")
    tape = QuantumScript([ops.IsingXX(param, wires=[0, 1])],
                         [M.purity([0]), M.vn_entropy([1]), M.vn_entropy([0], log_base=2),
                          M.density_matrix([0]), M.purity([0, 1])])
    pur, ent, ent2, rho, pur_all = dev.execute(tape)
    assert np.isclose(pur, np.sum(eigs ** 2)) and np.isclose(ent, -np.sum(eigs * np.log(eigs)))
    assert np.isclose(ent2, -np.sum(eigs * np.log(eigs)) / np.log(2)) and np.isclose(pur_all, 1)
    assert np.allclose(rho, np.diag([np.cos(param / 2) ** 2, np.sin(param / 2) ** 2]))
    tape = QuantumScript([ops.RY(param, wires=0), ops.CNOT(wires=[0, 1])], [M.mutual_info([0], [1])])
    expected = (-2 * np.cos(param / 2) ** 2 * np.log(np.cos(param / 2) ** 2 + 1e-10)
                - 2 * np.sin(param / 2) ** 2 * np.log(np.sin(param / 2) ** 2 + 1e-10))
    assert np.allclose(dev.execute(tape), expected)


def test_measurements_match_oracle_on_a_circuit():
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200 import QuantumScript, measurements as M, ops
    from pennylane_b200.simulate import simulate

    n = 10
    rng = np.random.default_rng(2)
    gates = []
    for _ in range(3):
        gates += [ops.RY(rng.uniform(0, 6), wires=w) for w in range(n)]
        gates += [ops.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
    mps = [M.density_matrix([7, 2]), M.purity([0, 1, 2, 3, 4]), M.vn_entropy([9, 0, 5], log_base=2),
           M.mutual_info([0, 1], [8, 3]), M.expval(ops.PauliZ(0))]
    tape = QuantumScript(gates, mps)
    got, ref = simulate(tape, fusion=1), oracle_simulate(tape)
    for g, r in zip(got, ref):
        assert np.allclose(g, r, rtol=1e-10, atol=1e-12)
    with pytest.raises(Exception, match="finite shots"):
        import pennylane_b200 as pb
        pb.device("b200.qubit").preprocess(QuantumScript(gates, [M.purity([0])], shots=10))


def test_large_state_properties():
    """24 qubits (the reference's route would need a 4^24 matrix): a GHZ-like state has
    rho_A = diag(1/2, 0, ..., 0, 1/2) on any proper subset, purity 1/2, entropy log 2."""
    from pennylane_b200 import StateVector, measurements as M, ops
    from pennylane_b200.simulate import measure

    n = 24
    sv = StateVector(n)
    sv.apply_operation(ops.Hadamard(0))
    for w in range(n - 1):
        sv.apply_operation(ops.CNOT(wires=[w, w + 1]))
    rho = sv.reduced_dm([23, 0, 11, 5])
    expect = np.zeros((16, 16), dtype=complex)
    expect[0, 0] = expect[15, 15] = 0.5
    assert np.allclose(rho, expect, atol=1e-13) and np.isclose(np.trace(rho).real, 1.0)
    assert np.isclose(measure(M.purity([3, 20]), sv), 0.5)
    assert np.isclose(measure(M.vn_entropy([12]), sv), np.log(2))
    assert np.isclose(measure(M.mutual_info([0], [23]), sv), np.log(2))
