"""``_adjoint_jacobian_state`` (adjoint_jacobian.py:43-73): the Jacobian of the state itself.
Known answers: tests/devices/qubit/test_adjoint_jacobian.py:316-345."""
import numpy as np
import pytest

from pennylane_b200 import QuantumScript, measurements as M, ops

X, Y = 0.5, 0.6
C_X, S_X, C_Y, S_Y = np.cos(X / 2), np.sin(X / 2), np.cos(Y / 2), np.sin(Y / 2)
X_JAC = np.array([-0.5 * C_Y * S_X, -0.5 * S_Y * S_X, -0.5j * S_Y * C_X, -0.5j * C_X * C_Y])
Y_JAC = np.array([-0.5 * C_X * S_Y, 0.5 * C_X * C_Y, -0.5j * S_X * C_Y, 0.5j * S_X * S_Y])


def _tapes():
    one = QuantumScript([ops.RX(1.2, wires=0)], [M.state()])
    two = QuantumScript([ops.RX(X, wires=0), ops.RY(Y, wires=1), ops.CNOT(wires=[0, 1])], [M.state()])
    return one, two


def _layered(n=6, seed=0):
    rng = np.random.default_rng(seed)
    gates = [ops.StatePrep(np.eye(2 ** n)[3], wires=list(range(n)))]
    for _ in range(2):
        gates += [ops.RY(rng.uniform(0, 6), wires=w) for w in range(n)]
        gates += [ops.IsingXX(rng.uniform(0, 6), wires=[w, (w + 1) % n]) for w in range(0, n, 2)]
        gates += [ops.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
    tape = QuantumScript(gates, [M.state()])
    tape.trainable_params = [1, 2, 5, 7, 8, 13, 16]
    return tape


def test_oracle_known_answers():
    from oracle.adjoint_jacobian import adjoint_jacobian_state

    one, two = _tapes()
    (jac,) = adjoint_jacobian_state(one)
    assert np.allclose(jac, [-0.5 * np.sin(0.6), -0.5j * np.cos(0.6)])
    x_jac, y_jac = adjoint_jacobian_state(two)
    assert np.allclose(x_jac, X_JAC) and np.allclose(y_jac, Y_JAC)


@pytest.mark.gpu
def test_device_known_answers_and_oracle_parity():
    from oracle.adjoint_jacobian import adjoint_jacobian_state
    from pennylane_b200.adjoint import adjoint_jacobian

    one, two = _tapes()
    (jac,) = adjoint_jacobian(one)
    assert np.allclose(jac, [-0.5 * np.sin(0.6), -0.5j * np.cos(0.6)])
    x_jac, y_jac = adjoint_jacobian(two)
    assert np.allclose(x_jac, X_JAC) and np.allclose(y_jac, Y_JAC)
    tape = _layered()
    got, ref = adjoint_jacobian(tape), adjoint_jacobian_state(tape)
    assert len(got) == len(ref) == 7
    for g, r in zip(got, ref):
        assert g.shape == (64,) and np.max(np.abs(g - r)) < 1e-12
    got32 = adjoint_jacobian(tape, dtype=np.complex64)
    assert max(np.max(np.abs(g - r)) for g, r in zip(got32, ref)) < 1e-5


def test_oracle_state_vjp_known_answers():
    """test_adjoint_jacobian.py:328-331, 352-359."""
    from oracle.adjoint_jacobian import adjoint_vjp
    from oracle.simulate import get_final_state

    one, two = _tapes()
    dy = np.array([0.5, 2.0], dtype=np.complex128)
    (vjp,) = adjoint_vjp(one, dy, get_final_state(one)[0])
    assert np.allclose(vjp, dy[0] * -0.5 * np.sin(0.6) + dy[1] * -0.5j * np.cos(0.6))
    dy = np.array([0.5, 1.0, 2.0, 2.5], dtype=np.complex128)
    x_vjp, y_vjp = adjoint_vjp(two, dy, get_final_state(two)[0])
    assert np.allclose(x_vjp, np.dot(X_JAC, dy)) and np.allclose(y_vjp, np.dot(Y_JAC, dy))


@pytest.mark.gpu
def test_device_state_vjp():
    from oracle.adjoint_jacobian import adjoint_vjp as oracle_vjp
    from oracle.simulate import get_final_state
    from pennylane_b200.adjoint import adjoint_jacobian, adjoint_vjp

    _, two = _tapes()
    dy = np.array([0.5, 1.0, 2.0, 2.5], dtype=np.complex128)
    x_vjp, y_vjp = adjoint_vjp(two, dy)
    assert np.allclose(x_vjp, np.dot(X_JAC, dy)) and np.allclose(y_vjp, np.dot(Y_JAC, dy))
    tape = _layered()
    rng = np.random.default_rng(1)
    dy = rng.normal(size=64) + 1j * rng.normal(size=64)
    got = adjoint_vjp(tape, dy, fusion=1)
    ref = oracle_vjp(tape, dy, get_final_state(tape)[0])
    jac = adjoint_jacobian(tape)
    assert len(got) == 7 and np.max(np.abs(np.array(got) - np.array(ref))) < 1e-12
    assert np.allclose(got, [np.dot(j, dy) for j in jac])


@pytest.mark.gpu
def test_device_pipeline_accepts_state_tapes():
    """default_qubit.py:243-283: the adjoint pipeline lets a state-only tape through."""
    import pennylane_b200 as pb

    _, two = _tapes()
    dev = pb.device("b200.qubit")
    tapes, cfg = dev.preprocess(two, pb.ExecutionConfig(gradient_method="adjoint"))
    (res,), (jac,) = dev.execute_and_compute_derivatives(tapes, cfg)
    assert np.allclose(jac[0], X_JAC) and np.allclose(jac[1], Y_JAC) and res.shape == (4,)
    (vjp,) = dev.compute_vjp(tapes, (np.array([0.5, 1.0, 2.0, 2.5], dtype=complex),), cfg)
    assert np.allclose(vjp[0], np.dot(X_JAC, [0.5, 1.0, 2.0, 2.5]))


# ---- batched cotangents: tests/devices/qubit/test_adjoint_jacobian.py:566-700 ---------------------
_X, _Y = 0.654, 1.221
_OBS3 = lambda: [M.expval(ops.PauliZ(0)), M.expval(ops.PauliY(1)), M.expval(ops.PauliX(0))]
_JAC3 = np.array([[-np.sin(_X), 0], [0, -np.cos(_Y)], [np.cos(_X), 0]])


@pytest.mark.gpu
@pytest.mark.parametrize("cotangents", ((0, 1.23), (1.232, -2.098, 0.323, 1.112),
                                        (5.212, -0.354, -2.575), (0.0, 0.0, 0.0)))
@pytest.mark.parametrize("trainable", ([0], [0, 1]))
def test_batched_cotangents_single_obs(cotangents, trainable):
    from pennylane_b200.adjoint import adjoint_vjp

    qs = QuantumScript([ops.RY(_X, wires=0), ops.RX(_Y, wires=1)], [M.expval(ops.PauliZ(0))],
                       trainable_params=trainable)
    actual = adjoint_vjp(qs, cotangents)
    assert isinstance(actual, tuple) and len(actual) == len(trainable)
    jac = np.array([[-np.sin(_X), 0]])[:, :len(trainable)]
    assert np.allclose(actual, jac.T @ np.expand_dims(np.array(cotangents), 0), atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("cotangents", [
    (np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0])),
    (np.array([0.653, 0, 0]), np.array([0, 0.573, 0]), np.array([0, 0, 1.232])),
    (np.array([0.653, -1.456]), np.array([0.498, 0.573]), np.array([0, 1.232])),
    (np.array([0.653, 0, 0, -1.234]), np.array([-0.323, 0.573, -1.449, -0.573]),
     np.array([0, 1, 1.232, 1.232])),
    (np.array([0.0, 0, 0]), np.array([0.0, 0, 0]), np.array([0.0, 0, 0])),
    (np.array([0.498, 0.573]), np.array([0.653, -1.456]), 0.0),          # :702-730 inhomogeneous
    (np.array([0.498, 0.573, -1.456]), 0.0, 0.0)])
@pytest.mark.parametrize("trainable", ([0], [0, 1]))
def test_batched_cotangents_multi_obs(cotangents, trainable):
    from pennylane_b200.adjoint import adjoint_vjp

    qs = QuantumScript([ops.RY(_X, wires=0), ops.RX(_Y, wires=1)], _OBS3(),
                       trainable_params=trainable)
    actual = adjoint_vjp(qs, cotangents, fusion=1)
    assert isinstance(actual, tuple) and len(actual) == len(trainable)
    B = next(len(c) for c in cotangents if np.ndim(c))
    dense = np.array([np.broadcast_to(c, (B,)) for c in cotangents], dtype=float)
    assert np.allclose(actual, _JAC3[:, :len(trainable)].T @ dense, atol=1e-12)


@pytest.mark.gpu
def test_compute_vjp_reuses_the_forward_state(monkeypatch):
    """default_qubit.py:1021-1029: with ``use_device_jacobian_product`` the state of ``execute`` is
    kept per tape hash and ``compute_vjp`` starts its reverse sweep from it."""
    import pennylane_b200 as pb
    from pennylane_b200 import adjoint as adj

    rng = np.random.default_rng(0)
    n = 8
    gates = []
    for _ in range(2):
        gates += [ops.RY(rng.uniform(0, 6), wires=f"q{w}") for w in range(n)]
        gates += [ops.CNOT(wires=[f"q{w}", f"q{(w + 1) % n}"]) for w in range(n)]
    tape = QuantumScript(gates, [M.expval(ops.PauliZ("q0")), M.expval(ops.PauliX("q3"))])
    dev = pb.device("b200.qubit")
    cfg = pb.ExecutionConfig(gradient_method="adjoint", use_device_jacobian_product=True)
    tapes, cfg = dev.preprocess(tape, cfg)
    cots = ((0.3, -1.2),)
    plain = pb.device("b200.qubit").compute_vjp(tapes, cots, cfg)          # no cache: own forward
    dev.execute(tapes, cfg)
    assert len(dev._state_cache) == 1

    def boom(*a, **k):
        raise AssertionError("forward pass re-run although the state was cached")

    monkeypatch.setattr(adj, "get_final_state", boom)
    cached = dev.compute_vjp(tapes, cots, cfg)
    assert np.allclose(cached, plain, rtol=0, atol=1e-14)
