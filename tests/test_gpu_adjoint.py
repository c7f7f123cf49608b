"""GPU parity: adjoint Jacobian / JVP / VJP against the oracle's restatement of
adjoint_jacobian.py and against closed forms / finite differences, as
tests/devices/qubit/test_adjoint_jacobian.py does in the reference (:32-47 finite-diff helper,
:363-393 and :475-500 closed forms)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_jac(tape):
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim

    st, _ = o_sim.get_final_state(tape)
    return np.array(o_adj.adjoint_jacobian(tape, st), dtype=float)


@pytest.mark.parametrize("theta", [0.1, -2.3, np.pi / 3])
@pytest.mark.parametrize("G,ob,expected", [
    ("RX", "PauliZ", lambda t: -np.sin(t)),     # test_adjoint_jacobian.py:363-377
    ("RY", "PauliZ", lambda t: -np.sin(t)),
    ("RX", "PauliY", lambda t: -np.cos(t)),
    ("RY", "PauliX", lambda t: np.cos(t)),
])
def test_closed_form_single_rotation(theta, G, ob, expected):
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jacobian

    tape = qb.QuantumScript([getattr(q, G)(theta, wires=0)], [qb.expval(getattr(q, ob)(wires=0))])
    jac = adjoint_jacobian(tape)
    assert abs(float(jac) - expected(theta)) < 1e-13


def _layered_tape(n, layers, seed, obs):
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    rng = np.random.default_rng(seed)
    ops_ = []
    for layer in range(layers):
        for i in range(n):
            ops_ += [q.RZ(rng.uniform(0, 6), wires=i), q.RY(rng.uniform(0, 6), wires=i),
                     q.RZ(rng.uniform(0, 6), wires=i)]
        for i in range(n):
            ops_.append(q.CNOT(wires=[i, (i + layer + 1) % n]))
    return qb.QuantumScript(ops_, [qb.expval(o) for o in obs],
                            trainable_params=list(range(3 * n * layers)))


@pytest.mark.parametrize("n", [3, 4, 9, 13])
def test_jacobian_matches_oracle_layered(n):
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jacobian

    obs = [q.PauliZ(wires=0), q.PauliX(wires=n - 1) @ q.PauliY(wires=0),
           q.LinearCombination([0.5, -1.5], [q.PauliZ(wires=1) @ q.PauliZ(wires=0), q.PauliX(wires=1)])]
    tape = _layered_tape(n, 2, n, obs)
    jac = np.array(adjoint_jacobian(tape), dtype=float)
    ref = _oracle_jac(tape)
    assert jac.shape == ref.shape == (3, 6 * n)
    assert np.max(np.abs(jac - ref)) < 1e-12


def test_all_generator_gates_against_oracle():
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jacobian

    n = 5
    ops_ = [q.Hadamard(wires=i) for i in range(n)] + [
        q.RX(0.1, wires=0), q.PhaseShift(0.2, wires=1), q.IsingXX(0.3, wires=[0, 4]),
        q.IsingYY(0.4, wires=[3, 1]), q.IsingZZ(0.5, wires=[2, 0]), q.IsingXY(0.6, wires=[1, 2]),
        q.CRX(0.7, wires=[4, 2]), q.CRY(0.8, wires=[0, 3]), q.CRZ(0.9, wires=[2, 1]),
        q.ControlledPhaseShift(1.0, wires=[1, 4]), q.SingleExcitation(1.1, wires=[3, 4]),
        q.SingleExcitationMinus(1.2, wires=[0, 1]), q.SingleExcitationPlus(1.3, wires=[2, 3]),
        q.MultiRZ(1.4, wires=[0, 2, 4]), q.PauliRot(1.5, "XYZ", wires=[1, 3, 0]),
        q.DoubleExcitation(1.6, wires=[0, 1, 2, 3]), q.MultiRZ(1.7, wires=[4, 3, 2, 1]),
        q.PauliRot(1.8, "XXYZ", wires=[0, 2, 1, 4]), q.T(wires=2), q.SWAP(wires=[0, 3]),
        q.Toffoli(wires=[0, 1, 2]), q.RY(1.9, wires=3), q.GlobalPhase(0.4, wires=0),
        q.adjoint(q.RX(0.33, wires=1)), q.ctrl(q.RZ(0.5, wires=0), 4),
    ]
    obs = [q.PauliZ(wires=0) @ q.PauliZ(wires=3), q.PauliY(wires=2),
           q.Hermitian(np.array([[1, 1j], [-1j, -0.5]]), wires=4)]
    tape = qb.QuantumScript(ops_, [qb.expval(o) for o in obs])
    jac = np.array(adjoint_jacobian(tape), dtype=float)
    ref = _oracle_jac(tape)
    assert jac.shape == ref.shape
    assert np.max(np.abs(jac - ref)) < 1e-12


def test_trainable_subset_and_state_prep():
    """adjoint_jacobian.py:117-133 bookkeeping: non-trainable params are skipped, StatePrep is
    excluded from the sweep (num_preps)."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jacobian

    prep = np.array([1, 1j, -1, 0.5, 0, 0.3, 0.2j, -0.1], dtype=complex)
    prep /= np.linalg.norm(prep)
    ops_ = [q.StatePrep(prep, wires=[0, 1, 2]), q.RX(0.4, wires=0), q.CNOT(wires=[0, 1]),
            q.RY(0.5, wires=1), q.Rot(0.1, 0.2, 0.3, wires=2), q.RZ(0.6, wires=2)]
    tape = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=1)), qb.expval(q.PauliX(wires=2))],
                            trainable_params=[1, 2, 6])
    jac = np.array(adjoint_jacobian(tape), dtype=float)
    ref = _oracle_jac(tape)
    assert jac.shape == (2, 3)
    assert np.max(np.abs(jac - ref)) < 1e-13


def test_finite_difference_cross_check():
    """test_adjoint_jacobian.py:32-47."""
    import pennylane_b200 as qb
    from pennylane_b200.adjoint import adjoint_jacobian
    from pennylane_b200 import ops as q

    n = 6
    tape = _layered_tape(n, 1, 3, [q.PauliZ(wires=0) @ q.PauliX(wires=3)])
    jac = np.array(adjoint_jacobian(tape), dtype=float)
    dev = qb.B200Qubit(wires=n)
    eps = 1e-6
    for p in [0, 5, 11, 17]:
        def val(shift):
            ops_ = list(tape.operations)
            k = 0
            for i, op in enumerate(ops_):
                if len(op.data) == 1:
                    if k == p:
                        ops_[i] = op._with_params([op.data[0] + shift])
                    k += 1
            return dev.execute(qb.QuantumScript(ops_, tape.measurements))
        fd = (val(eps) - val(-eps)) / (2 * eps)
        assert abs(fd - jac[p]) < 1e-8


def test_jvp_and_vjp_match_oracle():
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim

    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jvp, adjoint_vjp

    n = 6
    obs = [q.PauliZ(wires=0), q.PauliX(wires=2) @ q.PauliY(wires=5), q.PauliY(wires=3)]
    tape = _layered_tape(n, 2, 21, obs)
    st, _ = o_sim.get_final_state(tape)
    rng = np.random.default_rng(0)
    tangents = rng.normal(size=len(tape.trainable_params))
    tangents[3] = 0.0
    got = np.array(adjoint_jvp(tape, tangents), dtype=float)
    ref = np.array(o_adj.adjoint_jvp(tape, tangents, st), dtype=float)
    assert np.max(np.abs(got - ref)) < 1e-12
    cots = (0.3, 0.0, -1.7)
    got = np.array(adjoint_vjp(tape, cots), dtype=float)
    ref = np.array(o_adj.adjoint_vjp(tape, cots, st), dtype=float)
    assert np.max(np.abs(got - ref)) < 1e-12
    assert adjoint_vjp(tape, (0.0, 0.0, 0.0)) == tuple(0.0 for _ in tape.trainable_params)


def test_single_precision_adjoint():
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jacobian

    n = 8
    tape = _layered_tape(n, 2, 5, [q.PauliZ(wires=0)])
    jac = np.array(adjoint_jacobian(tape, dtype=np.complex64), dtype=float)
    assert np.max(np.abs(jac - _oracle_jac(tape))) < 1e-5


# ---- fused reverse sweep (b200q_apply_rtile in adjoint mode) -------------------------------------
@pytest.mark.parametrize("dtype,n,tol", [(np.complex128, 12, 1e-12), (np.complex128, 14, 1e-12),
                                         (np.complex64, 13, 2e-5), (np.complex64, 14, 2e-5)])
@pytest.mark.parametrize("fusion", [1, 2])
def test_fused_reverse_sweep_layered(dtype, n, tol, fusion):
    """Three observables (three bras sharing one ket) through the fused sweep."""
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jacobian

    obs = [q.PauliZ(wires=0), q.PauliX(wires=n - 1) @ q.PauliY(wires=0),
           q.LinearCombination([0.5, -1.5], [q.PauliZ(wires=1) @ q.PauliZ(wires=0), q.PauliX(wires=1)])]
    tape = _layered_tape(n, 2, n, obs)
    jac = np.array(adjoint_jacobian(tape, dtype=dtype, fusion=fusion), dtype=float)
    assert jac.shape == (3, 6 * n)
    assert np.max(np.abs(jac - _oracle_jac(tape))) < tol


@pytest.mark.parametrize("n", [12, 13])
def test_fused_reverse_sweep_mixed_gates(n):
    """Every generator shape the register kernel takes (X/Y/Z words, two-term projector
    generators, three-wire Z strings) next to non-trainable gates of all kinds, with only a
    subset of the parameters trainable."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import _Sweep, _reverse_sweep_fused, adjoint_jacobian
    from test_compiler import _trainable_circuit

    ops_ = _trainable_circuit(n, 90, seed=40 + n)
    npar = sum(len(o.data) for o in ops_)
    trainable = [i for i in range(npar) if i % 3 != 1]
    tape = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0) @ q.PauliX(wires=2)),
                                   qb.expval(q.PauliY(wires=n - 1))], trainable_params=trainable)
    jac = np.array(adjoint_jacobian(tape, fusion=1), dtype=float)
    assert np.max(np.abs(jac - _oracle_jac(tape))) < 1e-12
    # the fused path really ran (it returns None when it has to fall back)
    sw = _Sweep(tape.map_to_standard_wires(), np.complex128, None, 2, fusion=1)
    assert _reverse_sweep_fused(tape.map_to_standard_wires(), sw, 1) is not None


def test_fused_vjp_and_jvp_match_oracle():
    from pennylane_b200 import ops as q
    from pennylane_b200.adjoint import adjoint_jvp, adjoint_vjp

    n = 12
    obs = [q.PauliZ(wires=0), q.PauliX(wires=3)]
    tape = _layered_tape(n, 1, 5, obs)
    ref = _oracle_jac(tape)
    rng = np.random.default_rng(0)
    cot = rng.normal(size=2)
    tan = rng.normal(size=ref.shape[1])
    vjp = np.array(adjoint_vjp(tape, cot, fusion=1), dtype=float)
    jvp = np.array(adjoint_jvp(tape, tan, fusion=1), dtype=float)
    assert np.max(np.abs(vjp - cot @ ref)) < 1e-12
    assert np.max(np.abs(jvp - ref @ tan)) < 1e-12
