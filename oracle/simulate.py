"""ORACLE (test infrastructure only — never imported by ``pennylane_b200``).

CPU restatement of pennylane/devices/qubit/simulate.py (``get_final_state`` :174-242,
``measure_final_state`` :246-304, ``simulate`` :308-393) and initialize_state.py:24-59.
"""
import numpy as np

from .apply_operation import apply_operation
from .measure import measure
from .sampling import measure_with_samples


def create_initial_state(num_wires, prep_operation=None):   # initialize_state.py:24-59
    if prep_operation is None:
        state = np.zeros((2,) * num_wires, dtype=complex)
        state[(0,) * num_wires] = 1
        return state
    sv = prep_operation.state_vector(wire_order=list(range(num_wires)))
    dtype = "complex64" if str(sv.dtype) in ("float32", "complex64") else "complex128"
    return np.asarray(sv).astype(dtype)


def _is_prep(op):
    return hasattr(op, "state_vector")


def _op_batch(op):
    bs = getattr(op, "batch_size", None)
    return bs


class _FlexShots:                                           # simulate.py:104-117
    def __init__(self, shots):
        self.shot_vector = [(int(s), 1) for s in shots]
        self.total_shots = sum(int(s) for s in shots)
        self.has_partitioned_shots = len(self.shot_vector) > 1

    def __bool__(self):
        return True

    def __iter__(self):
        return iter(s for s, _ in self.shot_vector)


def _postselection_postprocess(state, is_state_batched, shots, rng=None, postselect_mode=None):
    """simulate.py:120-171."""
    if is_state_batched:
        raise ValueError("Cannot postselect on circuits with broadcasting.")
    norm = np.linalg.norm(state)
    if np.allclose(norm, 0.0):
        if postselect_mode == "fill-shots" and shots:
            raise RuntimeError("The probability of the postselected mid-circuit measurement "
                               "outcome is 0.")
        norm = 0.0
    if shots:
        binomial_fn = np.random.binomial if rng is None else rng.binomial
        postselected = (list(shots) if postselect_mode == "fill-shots"
                        else [int(binomial_fn(s, float(norm ** 2))) for s in shots])
        shots = _FlexShots(postselected)
    with np.errstate(divide="ignore", invalid="ignore"):
        state = state / norm
    return state, shots


class _ShotsView:
    def __init__(self, circuit, shots):
        self._c, self.shots = circuit, shots

    def __getattr__(self, name):
        return getattr(self._c, name)


def get_final_state(circuit, mid_measurements=None, rng=None, debugger=None):   # simulate.py:174-242
    ops = list(circuit.operations)
    prep = ops[0] if ops and _is_prep(ops[0]) else None
    op_wires = sorted({w for op in ops for w in op.wires})
    state = create_initial_state(len(op_wires), prep)
    is_state_batched = bool(prep is not None and _op_batch(prep) is not None)
    for op in ops[bool(prep):]:
        state = apply_operation(op, state, is_state_batched=is_state_batched,
                                mid_measurements=mid_measurements, rng=rng, debugger=debugger,
                                tape_shots=circuit.shots)
        if op.name == "Projector":                          # simulate.py:226-232
            state, new_shots = _postselection_postprocess(state, is_state_batched, circuit.shots,
                                                          rng=rng)
            circuit._postselected_shots = new_shots
        is_state_batched = is_state_batched or (_op_batch(op) is not None)
    for _ in range(circuit.num_wires - len(op_wires)):
        state = np.stack([state, np.zeros_like(state)], axis=-1)
    return state, is_state_batched


def measure_final_state(circuit, state, is_state_batched, rng=None, mid_measurements=None):
    """simulate.py:246-304."""
    if not circuit.shots:
        if mid_measurements is not None:
            raise TypeError("Native mid-circuit measurements are only supported with finite shots.")
        if len(circuit.measurements) == 1:
            return measure(circuit.measurements[0], state, is_state_batched)
        return tuple(measure(mp, state, is_state_batched) for mp in circuit.measurements)
    rng = np.random.default_rng(rng)
    results = measure_with_samples(circuit.measurements, state, circuit.shots,
                                   is_state_batched=is_state_batched, rng=rng,
                                   mid_measurements=mid_measurements)
    if len(circuit.measurements) == 1:
        if circuit.shots.has_partitioned_shots:
            return tuple(res[0] for res in results)
        return results[0]
    return results


def simulate_one_shot_native_mcm(circuit, rng=None, debugger=None):   # simulate.py:947-990
    mid_measurements = {}
    state, is_state_batched = get_final_state(circuit, mid_measurements=mid_measurements, rng=rng,
                                              debugger=debugger)
    return measure_final_state(circuit, state, is_state_batched, rng=rng,
                               mid_measurements=mid_measurements)


def simulate(circuit, rng=None, debugger=None):             # simulate.py:308-393
    has_mcm = any(op.name == "MidMeasureMP" for op in circuit.operations)
    if has_mcm:                                             # :354-381, one-shot method
        # the device hands ONE Generator to every shot (default_qubit.py:798)
        rng = np.random.default_rng(rng)
        aux_circ = circuit.copy(shots=[1])
        return tuple(simulate_one_shot_native_mcm(aux_circ, rng=rng, debugger=debugger)
                     for _ in range(circuit.shots.total_shots))
    if circuit.shots and any(op.name == "Projector" for op in circuit.operations):
        rng = np.random.default_rng(rng)
    circuit = _ShotsView(circuit, circuit.shots)
    state, is_state_batched = get_final_state(circuit, rng=rng, debugger=debugger)
    if getattr(circuit, "_postselected_shots", None) is not None:
        circuit.shots = circuit._postselected_shots         # circuit._shots = new_shots, :232
    return measure_final_state(circuit, state, is_state_batched, rng=rng)
