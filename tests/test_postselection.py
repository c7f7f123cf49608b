"""Projectors as operations = postselection (simulate.py:120-171, 226-232).  Known answers:
tests/devices/qubit/test_simulate.py:440-476."""
import numpy as np
import pytest

from pennylane_b200 import QuantumScript, measurements as M, ops


def _tapes():
    norm = QuantumScript([ops.PauliX(0), ops.RX(0.123, wires=1), ops.Projector([0], wires=1)],
                         [M.state()])
    nan = QuantumScript([ops.PauliX(0), ops.Projector([0], wires=0)], [M.state()])
    bcast = QuantumScript([ops.RX([0.1, 0.2], wires=0), ops.Projector([0], wires=0)], [M.state()])
    return norm, nan, bcast


def _bell_postselected(shots):
    return QuantumScript([ops.Hadamard(0), ops.CNOT(wires=[0, 1]), ops.RY(0.8, wires=0),
                          ops.Projector([1], wires=0), ops.CNOT(wires=[1, 2])],
                         [M.sample(wires=[0, 1, 2]), M.expval(ops.PauliZ(2))], shots=shots)


def test_oracle_known_answers():
    from oracle.simulate import simulate

    norm, nan, bcast = _tapes()
    assert np.isclose(np.linalg.norm(simulate(norm)), 1.0)             # :443-449
    assert np.all(np.isnan(simulate(nan)))                             # :467-476
    with pytest.raises(ValueError, match="Cannot postselect on circuits with broadcasting"):
        simulate(bcast)                                                # :451-463
    samples, ez = simulate(_bell_postselected(400), rng=np.random.default_rng(3))
    assert 100 < len(samples) < 300 and np.all(samples[:, 0] == 1)
    assert np.array_equal(samples[:, 1], samples[:, 2])


@pytest.mark.gpu
@pytest.mark.parametrize("fusion", [0, 1])
def test_device_matches_oracle(fusion):
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200.simulate import simulate

    norm, nan, bcast = _tapes()
    got = simulate(norm, fusion=fusion)
    assert np.isclose(np.linalg.norm(got), 1.0)
    assert np.max(np.abs(got - oracle_simulate(norm))) < 1e-12
    assert np.all(np.isnan(simulate(nan, fusion=fusion)))
    with pytest.raises(ValueError, match="Cannot postselect on circuits with broadcasting"):
        simulate(bcast, fusion=fusion)
    for shots in (400, [50, 70]):
        got = simulate(_bell_postselected(shots), rng=np.random.default_rng(3), fusion=fusion)
        ref = oracle_simulate(_bell_postselected(shots), rng=np.random.default_rng(3))
        if isinstance(shots, list):
            for g, r in zip(got, ref):
                assert np.array_equal(g[0], r[0]) and np.isclose(g[1], r[1])
        else:
            assert np.array_equal(got[0], ref[0]) and np.isclose(got[1], ref[1])
    # a projector that removes every shot: empty samples, no kernel launch with zero shots
    dead = QuantumScript([ops.Projector([1], wires=0)], [M.sample(wires=[0])], shots=5)
    assert simulate(dead, rng=np.random.default_rng(0), fusion=fusion).shape[0] == 0
