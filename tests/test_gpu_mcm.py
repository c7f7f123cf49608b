"""GPU parity of native mid-circuit measurements (apply_operation.py:355-497,
simulate.py:354-381, 947-990): ``b200q_collapse`` against numpy, ``apply_mid_measure`` and the
one-shot loop against the oracle under the same seed (identical sampled bits, states within
1e-12 / 1e-5), and the reference's known answers
(tests/devices/qubit/test_apply_operation.py:1526-1640)."""
import numpy as np
import pytest

from conftest import TOL, random_state

pytestmark = pytest.mark.gpu


def _sv(state, dtype=np.complex128):
    from pennylane_b200 import StateVector

    sv = StateVector(state.ndim, dtype=dtype)
    sv.set_state(state.astype(dtype))
    return sv


def _get(sv):
    return sv.to_numpy().reshape((2,) * sv.n)


class _FixedBinomial:
    def __init__(self, value):
        self.value = value

    def binomial(self, *_):
        return self.value


# ---- the kernel -----------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [1, 2, 5, 11, 17])
def test_collapse_kernel_every_wire(n, dtype):
    state = random_state(n, seed=n, dtype=dtype)
    for wire in sorted({0, n // 2, n - 1}):
        for sample in (0, 1):
            for reset in (False, True):
                sv = _sv(state, dtype)
                scale = 1.37
                sv.collapse(wire, sample, reset, scale)
                kept = np.take(state, sample, axis=wire) * np.asarray(scale, dtype=np.float64)
                zero = np.zeros_like(kept)
                lands = 0 if (reset or sample == 0) else 1
                ref = np.stack([kept, zero] if lands == 0 else [zero, kept], axis=wire)
                got = _get(sv)
                assert np.allclose(got, ref, rtol=TOL[np.dtype(dtype)], atol=0), (wire, sample, reset)
                assert np.array_equal(got == 0, ref == 0)


# ---- tests/devices/qubit/test_apply_operation.py:1578-1604 -------------------------------------
@pytest.mark.parametrize("m_res", [(0, 0), (1, 1)])
def test_mid_measure_known_answer(m_res):
    from pennylane_b200.mcm import measure

    initial_state = np.array([[0.09068964 + 0.36775595j, 0.37578343 + 0.4786927j],
                              [0.3537292 + 0.27214766j, 0.01928256 + 0.53536021j]])
    mid_state, end_state = np.zeros((2, 2), dtype=complex), np.zeros((2, 2), dtype=complex)
    mid_state[m_res[0]] = initial_state[m_res[0]] / np.linalg.norm(initial_state[m_res[0]])
    end_state[m_res] = mid_state[m_res] / np.abs(mid_state[m_res])
    m0, m1 = measure(0).measurements[0], measure(1).measurements[0]
    mid_meas = {}
    rng = _FixedBinomial(m_res[0])
    sv = _sv(initial_state)
    sv.apply_operation(m0, mid_measurements=mid_meas, rng=rng)
    assert np.allclose(mid_state, _get(sv))
    sv.apply_operation(m1, mid_measurements=mid_meas, rng=rng)
    assert np.allclose(end_state, _get(sv))
    assert mid_meas == {m0: m_res[0], m1: m_res[1]}


# ---- tests/devices/qubit/test_apply_operation.py:1529-1576 -------------------------------------
@pytest.mark.parametrize("unitary_name", ("CRX", "CRZ"))
@pytest.mark.parametrize("wires", ([0, 1], [1, 0]))
def test_conditional_known_answer(wires, unitary_name):
    from pennylane_b200 import ops
    from pennylane_b200.mcm import Conditional, measure

    unitary = getattr(ops, unitary_name)
    initial_state = np.array([0.3541035 + 0.05231577j, 0.6912382 + 0.49474503j,
                              0.29276263 + 0.06231887j, 0.10736635 + 0.21947607j])
    rotated = (initial_state @ ops.matrix(unitary(-0.238, wires=wires), wire_order=[0, 1]).T
               ).reshape(2, 2)
    m0 = measure(0)
    op = Conditional(m0, unitary(0.238, wires=wires))
    sv = _sv(rotated)
    sv.apply_operation(op, mid_measurements={m0.measurements[0]: 0})
    assert np.allclose(rotated, _get(sv))
    sv.apply_operation(op, mid_measurements={m0.measurements[0]: 1})
    assert np.allclose(initial_state, _get(sv).reshape(4))


# ---- tests/devices/qubit/test_apply_operation.py:1606-1637 -------------------------------------
def test_floating_point_edge_case_and_errors():
    from pennylane_b200 import StateVector, ops
    from pennylane_b200.mcm import MidMeasure

    sv = StateVector(4)
    rng = np.random.default_rng(0)
    for op in [ops.RX(-5.754168297787336, wires=0), ops.Hadamard(1), MidMeasure(1), MidMeasure(2),
               MidMeasure(3)]:
        sv.apply_operation(op, mid_measurements={}, rng=rng)
    assert np.isclose(np.linalg.norm(sv.to_numpy()), 1.0)

    big = StateVector(1)
    big.set_state(np.array([1.0005, 0.0], dtype=complex))
    with pytest.raises(ValueError, match="probabilities greater than 1."):
        big.apply_operation(MidMeasure(0), mid_measurements={})

    batched = StateVector(1, batch=2)
    with pytest.raises(ValueError, match="MidMeasure cannot be applied to batched states."):
        batched.apply_operation(MidMeasure(0), mid_measurements={})


# ---- against the oracle, same seed ----------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [3, 10, 16])
def test_mid_measure_sequence_matches_oracle(n, dtype):
    from oracle.apply_operation import apply_operation as oracle_apply
    from pennylane_b200 import ops
    from pennylane_b200.mcm import MidMeasure

    state = random_state(n, seed=100 + n, dtype=dtype)
    seq = []
    for i, w in enumerate([0, n - 1, n // 2, 0, 1]):
        seq += [ops.RY(0.3 + i, wires=w), ops.CNOT(wires=[w, (w + 1) % n]),
                MidMeasure(w, reset=bool(i % 2))]
    sv = _sv(state, dtype)
    ref = state.copy()
    mm_dev, mm_ref = {}, {}
    r1, r2 = np.random.default_rng(9), np.random.default_rng(9)
    for op in seq:
        sv.apply_operation(op, mid_measurements=mm_dev, rng=r1)
        ref = oracle_apply(op, ref, mid_measurements=mm_ref, rng=r2)
    assert list(mm_dev.values()) == [int(v) for v in mm_ref.values()]
    assert np.max(np.abs(_get(sv) - ref)) < 20 * TOL[np.dtype(dtype)]
    assert r1.random() == r2.random()                       # same stream position


def _dynamic_tape(n, shots, seed):
    """Entangling prefix, three measurements (one with reset, one postselected), conditionals on
    single values and on an expression, terminal samples / expval / probs."""
    from pennylane_b200 import QuantumScript, measurements as M, ops
    from pennylane_b200.mcm import cond, measure

    rng = np.random.default_rng(seed)
    tape_ops = []
    for layer in range(2):
        for w in range(n):
            tape_ops.append(ops.RY(rng.uniform(0, 2 * np.pi), wires=w))
            tape_ops.append(ops.RZ(rng.uniform(0, 2 * np.pi), wires=w))
        for w in range(n):
            tape_ops.append(ops.CNOT(wires=[w, (w + 1) % n]))
    m0, m1, m2 = measure(0), measure(n - 1, reset=True), measure(n // 2, postselect=1)
    tape_ops += [m0.measurements[0], cond(m0, ops.RX(0.4, wires=1)), ops.CNOT(wires=[1, 2]),
                 m1.measurements[0], cond(m0 & m1, ops.Hadamard(n - 1)),
                 cond(m0 + m1 == 1, ops.IsingXX(1.1, wires=[0, 2])),
                 ops.RY(0.9, wires=n // 2), m2.measurements[0], cond(~m2, ops.PauliX(0))]
    mps = [M.sample(wires=list(range(n))), M.expval(ops.PauliZ(0) @ ops.PauliX(1)),
           M.probs(wires=[0, n - 1]), M.sample(m0), M.sample(m1), M.sample(m2)]
    return QuantumScript(tape_ops, mps, shots=[1] * shots)


@pytest.mark.parametrize("fusion", [0, 1])
@pytest.mark.parametrize("n", [4, 12])
def test_one_shot_simulate_matches_oracle(n, fusion):
    """Every shot's result tuple — terminal samples, one-shot expval / probs, and the three
    mid-circuit bits — equals the oracle's under the same seed."""
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200.simulate import simulate

    tape = _dynamic_tape(n, shots=25, seed=n)
    got = simulate(tape, rng=np.random.default_rng(77), fusion=fusion)
    ref = oracle_simulate(tape, rng=np.random.default_rng(77))
    assert len(got) == len(ref) == 25
    for g, r in zip(got, ref):
        assert len(g) == len(r) == 6
        assert np.array_equal(np.asarray(g[0]), np.asarray(r[0]))
        assert np.allclose(g[1], r[1]) and np.allclose(g[2], r[2])
        assert [int(x) for x in g[3:]] == [int(x) for x in r[3:]]


def test_device_one_shot_pipeline_teleportation():
    """``preprocess`` -> ``execute`` -> ``postprocessing`` on the device: teleport RY(0.7)|0> with
    measured corrections; the statistics follow the closed form and equal the oracle-executed
    pipeline bit for bit under the same seed."""
    import pennylane_b200 as pb
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200 import QuantumScript, measurements as M, ops
    from pennylane_b200.mcm import cond, measure

    m0, m1 = measure("a"), measure("b")
    tape = QuantumScript(
        [ops.RY(0.7, wires="a"), ops.Hadamard("b"), ops.CNOT(wires=["b", "c"]),
         ops.CNOT(wires=["a", "b"]), ops.Hadamard("a"), m0.measurements[0], m1.measurements[0],
         cond(m1, ops.PauliX("c")), cond(m0, ops.PauliZ("c"))],
        [M.expval(ops.PauliZ("c")), M.probs(op=m0), M.counts(m1)], shots=600)
    dev = pb.device("b200.qubit", seed=21)
    tapes, cfg = dev.preprocess(tape)
    assert tapes[0].shots.total_shots == 600 and len(tapes[0].measurements) == 3
    (ez, p0, c1), = tapes.postprocessing(dev.execute(tapes, cfg))
    assert abs(ez - np.cos(0.7)) < 0.1 and abs(p0[1] - 0.5) < 0.1 and sum(c1.values()) == 600
    ref_raw = oracle_simulate(tapes[0].map_to_standard_wires(), rng=np.random.default_rng(21))
    (ez_r, p0_r, c1_r), = tapes.postprocessing((ref_raw,))
    assert ez == ez_r and np.array_equal(p0, p0_r) and c1 == c1_r

    with pytest.raises(pb.DeviceError, match="finite shots"):
        dev.preprocess(QuantumScript([m0.measurements[0]], [M.expval(ops.PauliZ("a"))]))


def test_prefix_is_simulated_once(monkeypatch):
    """The gates in front of the first MidMeasure run once, not once per shot."""
    from pennylane_b200 import StateVector
    from pennylane_b200.simulate import simulate

    calls = {"n": 0}
    orig = StateVector.apply_operation

    def counting(self, op, **kw):
        calls["n"] += 1
        return orig(self, op, **kw)

    monkeypatch.setattr(StateVector, "apply_operation", counting)
    tape = _dynamic_tape(4, shots=10, seed=1)
    simulate(tape, rng=np.random.default_rng(0), fusion=0)
    n_prefix = next(i for i, op in enumerate(tape.operations) if op.name == "MidMeasureMP")
    # per shot at most the 6 unitary gates behind the first measurement (+ diagonalising gates)
    assert calls["n"] <= n_prefix + 10 * 12


# ---- Snapshot: tests/devices/qubit/test_apply_operation.py:771-873 on the device ---------------------
_SNAP_STATE = np.array([[0.04624539 + 0.3895457j, 0.22399401 + 0.53870339j],
                        [-0.483054 + 0.2468498j, -0.02772249 - 0.45901669j]])


def test_snapshot_known_answers():
    from oracle.measure import measure as oracle_measure
    from pennylane_b200 import StateVector, measurements as M, ops
    from pennylane_b200.device import Debugger
    from pennylane_b200.simulate import apply_gates

    sv = _sv(_SNAP_STATE)
    apply_gates(sv, [ops.Snapshot()])                                   # no debugger: nothing
    assert np.array_equal(_get(sv), _SNAP_STATE)
    dbg = Debugger()
    apply_gates(sv, [ops.Snapshot(), ops.Snapshot("abcd")], debugger=dbg)
    assert list(dbg.snapshots) == [0, "abcd"] and dbg.snapshots[0].shape == (4,)
    assert np.array_equal(dbg.snapshots[0], _SNAP_STATE.ravel())
    assert np.array_equal(dbg.snapshots["abcd"], _SNAP_STATE.ravel())
    for mp in (M.expval(ops.PauliX(0)), M.var(ops.PauliZ(1)), M.probs(wires=[0])):
        dbg = Debugger()
        apply_gates(sv, [ops.Snapshot(measurement=mp)], debugger=dbg)
        assert np.allclose(dbg.snapshots[0], oracle_measure(mp, _SNAP_STATE), rtol=1e-12, atol=1e-15)
    dbg = Debugger()
    one = StateVector(1)
    apply_gates(one, [ops.Snapshot("tag", M.sample(wires=0), shots=50)], debugger=dbg)
    assert dbg.snapshots["tag"].shape == (50, 1) and not dbg.snapshots["tag"].any()
    dbg = Debugger()
    two = StateVector(1, batch=2)
    two.set_state(np.array([[1.0, 0.0], [0.0, 0.1]], dtype=complex))
    apply_gates(two, [ops.Snapshot()], debugger=dbg)
    assert np.array_equal(dbg.snapshots[0], np.array([[1.0, 0.0], [0.0, 0.1]]))


@pytest.mark.parametrize("fusion", [0, 1])
def test_snapshots_through_the_device_match_oracle(fusion):
    """Snapshots inside a fused run split it; state / expval / shot snapshots and the final
    samples equal the oracle's under the same seed."""
    import pennylane_b200 as pb
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200 import QuantumScript, measurements as M, ops
    from pennylane_b200.device import Debugger

    n = 6
    rng = np.random.default_rng(4)
    gates = []
    for layer in range(3):
        gates += [ops.RY(rng.uniform(0, 6), wires=w) for w in range(n)]
        gates += [ops.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
        gates.append(ops.Snapshot(f"s{layer}"))
        gates.append(ops.Snapshot("e", M.expval(ops.PauliZ(0) @ ops.PauliZ(1)), shots=None))
        gates.append(ops.Snapshot("shots", M.sample(wires=[0, 2])))
    tape = QuantumScript(gates, [M.sample(wires=list(range(n)))], shots=20)
    dev = pb.device("b200.qubit", seed=8, fusion=fusion)
    with Debugger(dev) as dbg:
        res = dev.execute(tape)
    assert dev._debugger is None
    ref_dbg = Debugger()
    ref = oracle_simulate(tape, rng=np.random.default_rng(8), debugger=ref_dbg)
    assert np.array_equal(res, ref)
    assert set(dbg.snapshots) == set(ref_dbg.snapshots) == {"s0", "s1", "s2", "e", "shots"}
    for k in ("s0", "s1", "s2"):
        assert np.max(np.abs(dbg.snapshots[k] - ref_dbg.snapshots[k])) < 1e-12
    assert np.allclose(dbg.snapshots["e"], ref_dbg.snapshots["e"], atol=1e-12)
    assert len(dbg.snapshots["shots"]) == 3
    for a, b in zip(dbg.snapshots["shots"], ref_dbg.snapshots["shots"]):
        assert a.shape == (20, 2) and np.array_equal(a, b)


@pytest.mark.parametrize("fusion", [0, 1])
def test_branch_cache_is_transparent(fusion):
    """Cached branch states (default), no cache at all, and an exhausted cache (every new branch
    in the scratch buffer) give identical per-shot results — and the oracle's."""
    from oracle.simulate import simulate as oracle_simulate
    from pennylane_b200.simulate import _simulate_native_mcm

    tape = _dynamic_tape(8, shots=60, seed=5).map_to_standard_wires()
    runs = [_simulate_native_mcm(tape, np.random.default_rng(9), np.complex128, None, True, fusion,
                                 cache_fraction=f) for f in (0.5, 0.0, 1e-15)]
    ref = oracle_simulate(tape, rng=np.random.default_rng(9))
    for res in runs:
        assert len(res) == 60
        for g, r in zip(res, ref):
            assert np.array_equal(np.asarray(g[0]), np.asarray(r[0]))
            assert np.allclose(g[1], r[1]) and np.allclose(g[2], r[2])
            assert [int(x) for x in g[3:]] == [int(x) for x in r[3:]]
