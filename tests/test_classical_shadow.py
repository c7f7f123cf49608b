"""``classical_shadow`` (measurements/classical_shadow.py:142-257, sampling.py:338-374): the
oracle restatement on known states, and — on the GPU — bits and recipes identical to the
oracle's under the same seeds."""
import numpy as np
import pytest

from pennylane_b200 import QuantumScript, measurements as M, ops
from pennylane_b200.tape import Shots


def _circuit(n, seed):
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(2):
        gates += [ops.RY(rng.uniform(0, 6), wires=w) for w in range(n)]
        gates += [ops.RX(rng.uniform(0, 6), wires=w) for w in range(n)]
        gates += [ops.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
    return gates


def test_oracle_on_basis_and_plus_states():
    """tests/measurements/test_classical_shadow.py: Z-recipe bits of |0..0> are 0, X-recipe bits of
    |+..+> are 0, and the recipes depend on the measurement's seed only."""
    from oracle.sampling import classical_shadow_process_state_with_shots as shadow

    n, shots = 3, 200
    zero = np.zeros((2,) * n, dtype=complex)
    zero[(0,) * n] = 1
    bits, recipes = shadow(M.classical_shadow(range(n), seed=7), zero, shots, rng=1)
    assert bits.shape == recipes.shape == (shots, n) and bits.dtype == np.int8
    assert set(np.unique(recipes)) == {0, 1, 2} and not bits[recipes == 2].any()
    assert 0.3 < bits[recipes != 2].mean() < 0.7
    plus = np.full((2,) * n, 2 ** (-n / 2), dtype=complex)
    bits_p, recipes_p = shadow(M.classical_shadow(range(n), seed=7), plus, shots, rng=2)
    assert np.array_equal(recipes_p, recipes) and not bits_p[recipes_p == 0].any()


@pytest.mark.gpu
@pytest.mark.parametrize("n, wires", [(3, [0, 1, 2]), (5, [4, 0, 2]), (8, list(range(8)))])
def test_device_bits_equal_oracle(n, wires):
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200.simulate import simulate

    tape = QuantumScript(_circuit(n, n), [M.classical_shadow(wires, seed=11 + n),
                                          M.sample(wires=[0])], shots=60)
    got = simulate(tape, rng=np.random.default_rng(5))
    ref = oracle_simulate(tape, rng=np.random.default_rng(5))
    assert got[0].shape == (2, 60, len(wires)) and got[0].dtype == np.int8
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


@pytest.mark.gpu
def test_device_shot_vector_and_many_wires():
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200.simulate import simulate

    n = 6
    tape = QuantumScript(_circuit(n, 1), [M.classical_shadow(range(n), seed=3)], shots=[5, 9])
    got = simulate(tape, rng=np.random.default_rng(2))
    ref = oracle_simulate(tape, rng=np.random.default_rng(2))
    assert len(got) == 2 and got[0].shape == (2, 5, n) and got[1].shape == (2, 9, n)
    assert all(np.array_equal(g, r) for g, r in zip(got, ref))
    # 20 wires measured: more than the 16 control bits one sweep carries
    n = 20
    gates = [ops.Hadamard(0)] + [ops.CNOT(wires=[w, w + 1]) for w in range(n - 1)]
    bits, recipes = simulate(QuantumScript(gates, [M.classical_shadow(range(n), seed=1)], shots=20),
                             rng=np.random.default_rng(0))
    z = recipes == 2                      # GHZ: all Z-basis bits of a shot agree
    for t in range(20):
        if z[t].any():
            assert len(set(bits[t][z[t]])) == 1


def _bell_tape(shots, k=1):
    H = [ops.PauliZ(0) @ ops.PauliZ(1), ops.PauliX(0) @ ops.PauliX(1),
         0.5 * (ops.PauliY(0) @ ops.PauliY(1)) + 2.0 * ops.PauliZ(2)]
    gates = [ops.Hadamard(0), ops.CNOT(wires=[0, 1]), ops.RY(0.4, wires=2)]
    return QuantumScript(gates, [M.shadow_expval(H, k=k, seed=13)], shots=shots)


def test_shadow_expval_oracle_estimates_bell_state():
    """Bell pair: <ZZ> = <XX> = 1, <YY> = -1 (tests/measurements/test_classical_shadow.py,
    ``TestExpvalForward``: estimates within the shadow's statistical error)."""
    from oracle.simulate import simulate as oracle_simulate

    res = oracle_simulate(_bell_tape(4000), rng=np.random.default_rng(1))
    assert np.allclose(res, [1.0, 1.0, -0.5 + 2.0 * np.cos(0.4)], atol=0.2)
    res5 = oracle_simulate(_bell_tape(4000, k=5), rng=np.random.default_rng(1))
    assert np.allclose(res5, [1.0, 1.0, -0.5 + 2.0 * np.cos(0.4)], atol=0.3)


def test_host_estimator_equals_oracle_estimator():
    from oracle.sampling import _median_of_means, _pauli_expval
    from pennylane_b200.shadows import median_of_means, pauli_expval

    rng = np.random.default_rng(0)
    bits, recipes = rng.integers(0, 2, (50, 4)), rng.integers(0, 3, (50, 4))
    words = np.array([[0, -1, 2, 1], [-1, -1, -1, 2], [2, 2, 2, 2]])
    a, b = pauli_expval(bits, recipes, words), _pauli_expval(bits, recipes, words)
    assert np.array_equal(a, b) and a.shape == (50, 3)
    assert np.array_equal(median_of_means(a, 3), _median_of_means(b, 3))


@pytest.mark.gpu
def test_device_shadow_expval_equals_oracle():
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200.simulate import simulate

    for k in (1, 4):
        got = simulate(_bell_tape(500, k), rng=np.random.default_rng(3))
        ref = oracle_simulate(_bell_tape(500, k), rng=np.random.default_rng(3))
        assert np.array_equal(got, ref)
    assert np.allclose(got, [1.0, 1.0, -0.5 + 2.0 * np.cos(0.4)], atol=0.5)
