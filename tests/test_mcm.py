"""Mid-circuit measurements on the CPU side: the oracle restatement of ``apply_mid_measure`` /
``apply_conditional`` (apply_operation.py:355-497) against the known answers the reference's own
tests hold (tests/devices/qubit/test_apply_operation.py:1526-1640), the host classes
(``MidMeasure``, ``MeasurementValue``, ``Conditional``) and the one-shot result layout
(simulate.py:354-381, sampling.py:235-267)."""
import numpy as np
import pytest

from oracle.apply_operation import apply_operation
from oracle.simulate import simulate as oracle_simulate
from pennylane_b200 import QuantumScript, measurements as M, ops
from pennylane_b200.mcm import Conditional, MeasurementValue, MidMeasure, cond, measure


class _FixedBinomial:
    def __init__(self, value):
        self.value = value

    def binomial(self, *_):
        return self.value


# ---- tests/devices/qubit/test_apply_operation.py:1578-1604 -------------------------------------
@pytest.mark.parametrize("m_res", [(0, 0), (1, 1)])
def test_mid_measure_known_answer(m_res):
    initial_state = np.array([[0.09068964 + 0.36775595j, 0.37578343 + 0.4786927j],
                              [0.3537292 + 0.27214766j, 0.01928256 + 0.53536021j]])
    mid_state, end_state = np.zeros((2, 2), dtype=complex), np.zeros((2, 2), dtype=complex)
    mid_state[m_res[0]] = initial_state[m_res[0]] / np.linalg.norm(initial_state[m_res[0]])
    end_state[m_res] = mid_state[m_res] / np.abs(mid_state[m_res])
    m0, m1 = measure(0).measurements[0], measure(1).measurements[0]
    mid_meas = {}
    rng = _FixedBinomial(m_res[0])
    res = apply_operation(m0, initial_state, mid_measurements=mid_meas, rng=rng)
    assert np.allclose(mid_state, res)
    res = apply_operation(m1, res, mid_measurements=mid_meas, rng=rng)
    assert np.allclose(end_state, res)
    assert mid_meas == {m0: m_res[0], m1: m_res[1]}


# ---- tests/devices/qubit/test_apply_operation.py:1529-1576 -------------------------------------
@pytest.mark.parametrize("batched", (False, True))
@pytest.mark.parametrize("unitary", (ops.CRX, ops.CRZ))
@pytest.mark.parametrize("wires", ([0, 1], [1, 0]))
def test_conditional_known_answer(wires, unitary, batched):
    n_states = int(batched) + 1
    initial_state = np.array([
        [0.3541035 + 0.05231577j, 0.6912382 + 0.49474503j,
         0.29276263 + 0.06231887j, 0.10736635 + 0.21947607j],
        [0.09803567 + 0.47557068j, 0.4427561 + 0.13810454j,
         0.26421703 + 0.5366283j, 0.03825933 + 0.4357423j]][:n_states])
    rotated = initial_state @ ops.matrix(unitary(-0.238, wires=wires), wire_order=[0, 1]).T
    rotated = np.squeeze(np.reshape(rotated, (n_states, 2, 2)))
    m0 = measure(0)
    op = Conditional(m0, unitary(0.238, wires=wires))
    mid_meas = {m0.measurements[0]: 0}
    old = apply_operation(op, rotated, batched, mid_measurements=mid_meas)
    assert np.allclose(rotated, old)
    mid_meas[m0.measurements[0]] = 1
    new = apply_operation(op, rotated, batched, mid_measurements=mid_meas)
    assert np.allclose(np.squeeze(initial_state), np.reshape(new, (n_states, 4)))


# ---- tests/devices/qubit/test_apply_operation.py:1606-1637 -------------------------------------
def test_floating_point_mcm_edge_case():
    tape_ops = [ops.RX(-5.754168297787336, wires=0), ops.Hadamard(1), MidMeasure(1),
                MidMeasure(2), MidMeasure(3)]
    state = np.zeros((2, 2, 2, 2), dtype=complex)
    state[0, 0, 0, 0] = 1
    rng = np.random.default_rng(0)
    with np.errstate(all="raise"):
        for op in tape_ops:
            state = apply_operation(op, state, mid_measurements={}, rng=rng)
    assert np.isclose(np.linalg.norm(state), 1.0)


def test_norm_greater_than_one():
    state = np.zeros((2,), dtype=complex)
    state[0] = 1.0005
    with pytest.raises(ValueError, match="probabilities greater than 1."):
        apply_operation(MidMeasure(0), state, mid_measurements={})


def test_batched_state_rejected():
    with pytest.raises(ValueError, match="MidMeasure cannot be applied to batched states."):
        apply_operation(measure(0).measurements[0], np.array([[1, 0], [1, 0]], dtype=complex),
                        is_state_batched=True)


# ---- host classes ---------------------------------------------------------------------------------
def test_measurement_value_arithmetic_and_concretize():
    a, b = measure(0), measure("b")
    ma, mb = a.measurements[0], b.measurements[0]
    assert ma.name == "MidMeasureMP" and ma.wires == (0,) and not ma.reset and ma.postselect is None
    assert a.concretize({ma: 1}) == 1
    both = a & b
    assert [m.wires for m in both.measurements] == [(0,), ("b",)]
    assert both.concretize({ma: 1, mb: 1}) and not both.concretize({ma: 1, mb: 0})
    assert (a + 2 * b).concretize({ma: 1, mb: 1}) == 3
    assert (~a).concretize({ma: 0}) and (a == 0).concretize({ma: 0})
    with pytest.raises(ValueError, match="truth value"):
        bool(a)


def test_map_wires_keeps_measurement_identity():
    mv = measure("q", reset=True, postselect=1)
    m = mv.measurements[0]
    tape = QuantumScript([ops.Hadamard("q"), m, cond(mv, ops.PauliX("r"))],
                         [M.sample(wires=["r"]), M.sample(mv)], shots=3)
    std = tape.map_to_standard_wires()
    m_std = std.operations[1]
    assert m_std.wires == (0,) and m_std.reset and m_std.postselect == 1
    assert m_std == m and hash(m_std) == hash(m)
    assert std.operations[2].meas_val.concretize({m_std: 1}) == 1
    assert std.measurements[1].mv.measurements[0] == m_std and std.measurements[1].obs is None


# ---- one-shot result layout (simulate.py:354-381) ---------------------------------------------------
def test_one_shot_layout_and_correlation():
    """H on wire 0, measure it, X on wire 1 iff the outcome was 1: every shot's terminal sample
    of wire 1 equals the mid-circuit sample; one result tuple per shot, MCM values last."""
    mv = measure(0)
    tape = QuantumScript([ops.Hadamard(0), mv.measurements[0], cond(mv, ops.PauliX(1))],
                         [M.sample(wires=[1]), M.sample(mv)], shots=40)
    res = oracle_simulate(tape, rng=np.random.default_rng(3))
    assert len(res) == 40 and all(len(r) == 2 for r in res)
    term = np.array([int(np.squeeze(r[0])) for r in res])
    mid = np.array([int(r[1]) for r in res])
    assert np.array_equal(term, mid) and 0 < mid.sum() < 40


def test_reset_and_analytic_rejection():
    mv = measure(0, reset=True)
    tape = QuantumScript([ops.PauliX(0), mv.measurements[0]], [M.sample(wires=[0]), M.sample(mv)],
                         shots=5)
    res = oracle_simulate(tape, rng=np.random.default_rng(1))
    assert all(int(np.squeeze(r[0])) == 0 and int(r[1]) == 1 for r in res)
    with pytest.raises(TypeError, match="only supported with finite shots"):
        from oracle.simulate import get_final_state, measure_final_state
        t0 = QuantumScript([mv.measurements[0]], [M.expval(ops.PauliZ(0))])
        mm = {}
        st, b = get_final_state(t0, mid_measurements=mm, rng=np.random.default_rng(0))
        measure_final_state(t0, st, b, mid_measurements=mm)


def test_measurement_value_is_exported():
    assert isinstance(measure(1), MeasurementValue)


# ---- one-shot post-processing (transforms/dynamic_one_shot.py) with the oracle as executor ----------
def _run_one_shot(tape, seed):
    from pennylane_b200.one_shot import dynamic_one_shot

    aux, post = dynamic_one_shot(tape)
    assert aux.shots.total_shots == tape.shots.total_shots
    assert all(sc.shots == 1 for sc in aux.shots.shot_vector)
    return post(oracle_simulate(aux, rng=np.random.default_rng(seed)))


def test_one_shot_teleportation_is_deterministic():
    """Teleport RY(0.7)|0> from wire 0 to wire 2 with measured corrections: <Z_2> sampled over the
    shots must follow cos(0.7) and the two mid-circuit outcomes must be uniform."""
    m0, m1 = measure(0), measure(1)
    tape = QuantumScript(
        [ops.RY(0.7, wires=0), ops.Hadamard(1), ops.CNOT(wires=[1, 2]), ops.CNOT(wires=[0, 1]),
         ops.Hadamard(0), m0.measurements[0], m1.measurements[0],
         cond(m1, ops.PauliX(2)), cond(m0, ops.PauliZ(2))],
        [M.expval(ops.PauliZ(2)), M.probs(op=m0), M.expval(m1), M.counts(m0), M.var(ops.PauliZ(2))],
        shots=4000)
    ez, p0, e1, c0, vz = _run_one_shot(tape, 11)
    assert abs(ez - np.cos(0.7)) < 0.05
    assert p0.shape == (2,) and abs(p0[1] - 0.5) < 0.05 and np.isclose(p0.sum(), 1)
    assert abs(e1 - 0.5) < 0.05
    assert set(c0) == {0.0, 1.0} and sum(c0.values()) == 4000
    assert abs(vz - np.sin(0.7) ** 2) < 0.05


def test_one_shot_postselection_drops_invalid_shots():
    mv = measure(0, postselect=1)
    tape = QuantumScript([ops.Hadamard(0), mv.measurements[0], ops.CNOT(wires=[0, 1])],
                         [M.sample(wires=[1]), M.sample(mv), M.expval(ops.PauliZ(1))], shots=200)
    s1, smv, ez = _run_one_shot(tape, 5)
    assert 60 < len(s1) < 140 and len(s1) == len(smv)
    assert np.all(s1 == 1) and np.all(smv == 1) and ez == -1.0


def test_one_shot_shot_vector_and_passthrough():
    from pennylane_b200.one_shot import dynamic_one_shot

    mv = measure(0)
    tape = QuantumScript([ops.PauliX(0), mv.measurements[0]], [M.expval(mv)], shots=[3, 5])
    res = _run_one_shot(tape, 0)
    assert res == (1.0, 1.0)
    plain = QuantumScript([ops.PauliX(0)], [M.expval(ops.PauliZ(0))], shots=3)
    same, post = dynamic_one_shot(plain)
    assert same is plain and post("x") == "x"
    with pytest.raises(ValueError, match="finite shots"):
        dynamic_one_shot(QuantumScript([mv.measurements[0]], [M.expval(mv)]))


# ---- Snapshot: tests/devices/qubit/test_apply_operation.py:771-873 -----------------------------------
_SNAP_STATE = np.array([[0.04624539 + 0.3895457j, 0.22399401 + 0.53870339j],
                        [-0.483054 + 0.2468498j, -0.02772249 - 0.45901669j]])


def test_snapshot_known_answers_oracle():
    from oracle.measure import measure as oracle_measure
    from pennylane_b200.device import Debugger

    assert apply_operation(ops.Snapshot(), _SNAP_STATE) is _SNAP_STATE          # :774-786
    dbg = Debugger()
    new = apply_operation(ops.Snapshot(), _SNAP_STATE, debugger=dbg)            # :788-806
    assert np.allclose(new, _SNAP_STATE) and list(dbg.snapshots) == [0]
    assert dbg.snapshots[0].shape == (4,) and np.allclose(dbg.snapshots[0], _SNAP_STATE.ravel())
    dbg = Debugger()
    apply_operation(ops.Snapshot("abcd"), _SNAP_STATE, debugger=dbg)            # :808-827
    assert list(dbg.snapshots) == ["abcd"] and dbg.snapshots["abcd"].shape == (4,)
    for mp in (M.expval(ops.PauliX(0)), M.var(ops.PauliZ(1)), M.probs(wires=[0])):   # :829-850
        dbg = Debugger()
        apply_operation(ops.Snapshot(measurement=mp), _SNAP_STATE, debugger=dbg)
        assert np.array_equal(dbg.snapshots[0], oracle_measure(mp, _SNAP_STATE))
    dbg = Debugger()                                                            # :852-861
    apply_operation(ops.Snapshot("tag", M.sample(wires=0), shots=50),
                    np.array([1.0, 0.0], dtype=complex), debugger=dbg)
    assert dbg.snapshots["tag"].shape == (50, 1)
    dbg = Debugger()                                                            # :863-873
    batched = np.array([[1.0, 0.0], [0.0, 0.1]], dtype=complex)
    apply_operation(ops.Snapshot(), batched, is_state_batched=True, debugger=dbg)
    assert set(dbg.snapshots) == {0} and np.array_equal(dbg.snapshots[0], batched)
    # repeated tags collect into a list (apply_operation.py:908-915)
    dbg = Debugger()
    for _ in range(3):
        apply_operation(ops.Snapshot("t", M.probs(wires=[1])), _SNAP_STATE, debugger=dbg)
    assert isinstance(dbg.snapshots["t"], list) and len(dbg.snapshots["t"]) == 3


def test_snapshot_operator_rules():
    snap = ops.Snapshot()
    assert not snap.hyperparameters["shots"] and snap.hyperparameters["measurement"].kind == "state"
    assert ops.Snapshot(measurement=M.expval(ops.PauliZ("a"))).hyperparameters["shots"] == "workflow"
    assert ops.Snapshot("t", M.probs(wires=["a"])).map_wires({"a": 0}).wires == (0,)
    with pytest.raises(ValueError, match="tags can only be"):
        ops.Snapshot(tag=1.5)
    with pytest.raises(ValueError, match="not supported"):
        ops.Snapshot(measurement=ops.PauliZ(0))
