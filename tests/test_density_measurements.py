"""``density_matrix`` / ``purity`` / ``vn_entropy`` / ``mutual_info`` on the CPU side: the oracle
restatement (math/quantum.py:386-487, 563-663, 741-775) against the closed forms the reference's
tests assert (tests/measurements/test_purity_measurement.py:25-31, test_vn_entropy.py:29-41,
test_mutual_info.py:141-160, 186-203) and against a brute-force partial trace."""
import numpy as np
import pytest

from oracle.measure import measure as oracle_measure, reduce_statevector
from oracle.simulate import get_final_state
from pennylane_b200 import QuantumScript, measurements as M, ops

PARAMS = np.linspace(0.05, 2 * np.pi - 0.05, 7)


def _state(gates, n):
    st, _ = get_final_state(QuantumScript(gates, [M.probs(wires=list(range(n)))]))
    return st


def expected_purity_ising_xx(param):            # test_purity_measurement.py:25-31
    r = np.sqrt(1 - 4 * np.cos(param / 2) ** 2 * np.sin(param / 2) ** 2)
    return ((1 + r) / 2) ** 2 + ((1 - r) / 2) ** 2


def expected_entropy_ising_xx(param):           # test_vn_entropy.py:29-41
    r = np.sqrt(1 - 4 * np.cos(param / 2) ** 2 * np.sin(param / 2) ** 2)
    eigs = np.array([e for e in ((1 + r) / 2, (1 - r) / 2) if e > 0])
    return -np.sum(eigs * np.log(eigs))


@pytest.mark.parametrize("param", PARAMS)
def test_ising_xx_closed_forms(param):
    st = _state([ops.IsingXX(param, wires=[0, 1])], 2)
    for w in ([0], [1]):
        assert np.isclose(oracle_measure(M.purity(w), st), expected_purity_ising_xx(param))
        assert np.isclose(oracle_measure(M.vn_entropy(w), st), expected_entropy_ising_xx(param))
        assert np.isclose(oracle_measure(M.vn_entropy(w, log_base=2), st),
                          expected_entropy_ising_xx(param) / np.log(2))
    assert np.isclose(oracle_measure(M.purity([0, 1]), st), 1.0)


@pytest.mark.parametrize("state, expected", [                 # test_mutual_info.py:141-160
    ([1.0, 0.0, 0.0, 0.0], 0), ([np.sqrt(2) / 2, 0.0, np.sqrt(2) / 2, 0.0], 0),
    ([np.sqrt(2) / 2, 0.0, 0.0, np.sqrt(2) / 2], 2 * np.log(2)), (np.ones(4) * 0.5, 0.0)])
def test_mutual_info_known_answers(state, expected):
    full = np.kron(np.asarray(state, dtype=complex), np.array([1, 0, 0, 0])).reshape((2,) * 4)
    res = oracle_measure(M.mutual_info([0, 2], [1, 3]), full)
    assert np.allclose(res, expected, atol=1e-6)


@pytest.mark.parametrize("param", PARAMS)
def test_mutual_info_ry_cnot(param):                          # test_mutual_info.py:186-203
    st = _state([ops.RY(param, wires=0), ops.CNOT(wires=[0, 1])], 2)
    expected = (-2 * np.cos(param / 2) ** 2 * np.log(np.cos(param / 2) ** 2 + 1e-10)
                - 2 * np.sin(param / 2) ** 2 * np.log(np.sin(param / 2) ** 2 + 1e-10))
    assert np.allclose(oracle_measure(M.mutual_info([0], [1]), st), expected)


def test_reduce_statevector_against_brute_force():
    rng = np.random.default_rng(0)
    n = 6
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    full = np.outer(psi, psi.conj()).reshape([2] * (2 * n))
    for ind in ([0], [3, 1], [4, 0, 2], [1, 2, 3, 5], [2, 0, 4, 1, 3], list(range(n))):
        sub = list(range(n)) + [n + i if i in ind else i for i in range(n)]
        ref = np.einsum(full, sub, list(ind) + [n + i for i in ind]).reshape(2 ** len(ind), -1)
        assert np.allclose(reduce_statevector(psi, ind), ref)
    batch = np.stack([psi, np.roll(psi, 3)])
    assert np.allclose(reduce_statevector(batch, [2, 0])[1], reduce_statevector(batch[1], [2, 0]))


def test_measurement_classes():
    mp = M.mutual_info(["a"], ["b", "c"], log_base=2).map_wires({"a": 0, "b": 1, "c": 2})
    assert mp.wires == (0, 1, 2) and mp._wires == ((0,), (1, 2)) and mp.log_base == 2
    with pytest.raises(ValueError, match="must not overlap"):
        M.mutual_info([0, 1], [1])
    assert M.density_matrix([1, 0]).kind == "density_matrix" and M.purity(3).wires == (3,)
