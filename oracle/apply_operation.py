"""ORACLE (test infrastructure only — never imported by ``pennylane_b200``).

CPU restatement of pennylane/devices/qubit/apply_operation.py: the same numpy primitives
(``einsum``, ``tensordot``+``transpose``, ``roll``, ``stack``) in the same order and with the
same dispatch thresholds, on a state of shape ``[2]*n`` (``[B] + [2]*n`` when batched).
"""
import numpy as np

from .gates import matrix_of

EINSUM_OP_WIRECOUNT_PERF_THRESHOLD = 3      # apply_operation.py:29
EINSUM_STATE_WIRECOUNT_PERF_THRESHOLD = 13  # apply_operation.py:30
alphabet = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"

_HADAMARD = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)   # :32-41


def _get_slice(index, axis, num_axes):       # apply_operation.py:44-70
    idx = [slice(None)] * num_axes
    idx[axis] = index
    return tuple(idx)


def _batch_size(mat, dim):
    return mat.shape[0] if mat.ndim == 3 else None


def apply_operation_einsum(op, state, is_state_batched=False):   # apply_operation.py:151-199
    mat = matrix_of(op) + 0j
    total_indices = state.ndim - is_state_batched
    num_indices = len(op.wires)
    state_indices = alphabet[:total_indices]
    affected_indices = "".join(alphabet[i] for i in op.wires)
    new_indices = alphabet[total_indices: total_indices + num_indices]
    new_state_indices = state_indices
    for old, new in zip(affected_indices, new_indices):
        new_state_indices = new_state_indices.replace(old, new)
    einsum_indices = f"...{new_indices}{affected_indices},...{state_indices}->...{new_state_indices}"
    new_mat_shape = [2] * (num_indices * 2)
    bs = _batch_size(mat, 2**num_indices)
    if bs is not None:
        new_mat_shape = [bs] + new_mat_shape
    return np.einsum(einsum_indices, mat.reshape(new_mat_shape), state)


def apply_operation_tensordot(op, state, is_state_batched=False):   # apply_operation.py:202-255
    mat = matrix_of(op) + 0j
    total_indices = state.ndim - is_state_batched
    num_indices = len(op.wires)
    new_mat_shape = [2] * (num_indices * 2)
    bs = _batch_size(mat, 2**num_indices)
    is_mat_batched = bs is not None
    if is_mat_batched:
        new_mat_shape = [bs] + new_mat_shape
    reshaped_mat = mat.reshape(new_mat_shape)
    mat_axes = list(range(-num_indices, 0))
    state_axes = [i + is_state_batched for i in op.wires]
    tdot = np.tensordot(reshaped_mat, state, axes=(mat_axes, state_axes))
    unused_idxs = [i for i in range(total_indices) if i not in op.wires]
    perm = list(op.wires) + unused_idxs
    if is_mat_batched:
        perm = [0] + [i + 1 for i in perm]
    if is_state_batched:
        perm.insert(num_indices, -1)
    inv_perm = np.argsort(perm)
    return np.transpose(tdot, inv_perm)


def _op_batch_size(op):
    m = matrix_of(op) if op.name not in ("Identity",) else np.eye(2)
    return m.shape[0] if m.ndim == 3 else None


def _apply_operation_default(op, state, is_state_batched):   # apply_operation.py:341-351
    if (len(op.wires) < EINSUM_OP_WIRECOUNT_PERF_THRESHOLD
            and state.ndim < EINSUM_STATE_WIRECOUNT_PERF_THRESHOLD) or (
            _op_batch_size(op) and is_state_batched):
        return apply_operation_einsum(op, state, is_state_batched)
    return apply_operation_tensordot(op, state, is_state_batched)


def _apply_single_qubit_np(mat, state, axis):   # apply_operation.py:73-88
    return np.moveaxis(np.tensordot(mat, state, axes=[[1], [axis]]), 0, axis)


def _prepare_batched_params(params, state0, state1, axis, n_dim, is_state_batched):
    """apply_operation.py:116-148."""
    if is_state_batched:
        params = np.reshape(params, (-1,) + (1,) * (n_dim - 2))
    else:
        axis = axis + 1
        params = np.reshape(params, (-1,) + (1,) * (n_dim - 1))
        state0 = np.expand_dims(state0, 0) + np.zeros_like(params)
        state1 = np.expand_dims(state1, 0)
    return params, state0, state1, axis


def _rz_coeffs(params):   # :726-729
    e = np.exp(-0.5j * params)
    z = np.zeros_like(e)
    return e, z, z, np.conj(e)


def _rx_coeffs(params):   # :732-735
    c = np.cos(params / 2)
    js = -1j * np.sin(params / 2)
    return c, js, js, c


def _ry_coeffs(params):   # :738-741
    c = np.cos(params / 2)
    s = np.sin(params / 2)
    return c, -s, s, c


def _rotation_1q(op, state, is_state_batched, compute_coeffs):   # apply_operation.py:645-723
    n_dim = state.ndim
    axis = op.wires[0] + is_state_batched
    state0 = state[_get_slice(0, axis, n_dim)]
    state1 = state[_get_slice(1, axis, n_dim)]
    params = op.data[0]
    if np.ndim(params) == 1:                                       # op.batch_size is not None
        params = np.asarray(params).astype(complex)
        params, state0, state1, axis = _prepare_batched_params(
            params, state0, state1, axis, n_dim, is_state_batched)
        a, b, c, d = compute_coeffs(params)
        state0 = state0.astype(complex)
        state1 = state1.astype(complex)
        new0 = np.multiply(state0, a) + np.multiply(state1, b)
        new1 = np.multiply(state0, c) + np.multiply(state1, d)
        return np.stack([new0, new1], axis=axis)
    p = np.asarray(params, dtype=state.dtype)
    a, b, c, d = compute_coeffs(p)
    if n_dim < EINSUM_STATE_WIRECOUNT_PERF_THRESHOLD:
        return np.stack([np.multiply(state0, a) + np.multiply(state1, b),
                         np.multiply(state0, c) + np.multiply(state1, d)], axis=axis)
    mat = np.array([[a, b], [c, d]], dtype=state.dtype)
    return _apply_single_qubit_np(mat, state, axis)


def _apply_grover_without_matrix(state, op_wires, is_state_batched):
    """apply_operation.py:849-880: G = 2P - I with P the projector on the all-plus state of the
    operator's wires — sum over those axes, refill with the (unnormalised) all-plus state."""
    num_wires = len(op_wires)
    prefactor = 2 ** (1 - num_wires)
    sum_axes = [w + is_state_batched for w in op_wires]
    collapsed = np.sum(state, axis=tuple(sum_axes))
    if num_wires == state.ndim - is_state_batched:
        new_shape = (-1,) + (1,) * num_wires if is_state_batched else (1,) * num_wires
        return prefactor * np.reshape(collapsed, new_shape) - state
    all_plus = np.full([2] * num_wires, prefactor).astype(state.dtype)
    source = list(range(np.ndim(collapsed), np.ndim(state)))
    return np.moveaxis(np.tensordot(collapsed, all_plus, axes=0), source, sum_axes) - state


def _qubit_unitary(matrix, wires):
    from types import SimpleNamespace
    return SimpleNamespace(name="QubitUnitary", wires=tuple(wires), data=(matrix,),
                           hyperparameters={}, batch_size=None)


def apply_conditional(op, state, is_state_batched=False, mid_measurements=None, rng=None):
    """apply_operation.py:355-411 (numpy branch)."""
    if op.meas_val.concretize(mid_measurements):
        return apply_operation(op.base, state, is_state_batched=is_state_batched,
                               mid_measurements=mid_measurements, rng=rng)
    return state


def apply_mid_measure(op, state, is_state_batched=False, mid_measurements=None, rng=None):
    """apply_operation.py:415-497 (numpy branch): P(0) from the norm of the bit-0 slice,
    ``binomial(1, 1 - P0)``, projector and reset as 2x2 ``QubitUnitary`` sweeps."""
    if is_state_batched:
        raise ValueError("MidMeasure cannot be applied to batched states.")          # :441-442
    wire = list(op.wires)
    axis = wire[0]
    slices = [slice(None)] * np.ndim(state)
    slices[axis] = 0
    prob0 = np.real(np.linalg.norm(state[tuple(slices)])) ** 2                       # :450
    norm = np.sum(prob0, axis=-1)                                                     # :453
    eps = 10 * np.finfo(state.dtype).eps
    if (norm - 1) > eps:
        raise ValueError(f"probabilities greater than 1. Got norm {norm}.")
    if norm > 1:
        prob0 = prob0 / norm
    binomial_fn = np.random.binomial if rng is None else rng.binomial                # :468
    sample = binomial_fn(1, 1 - prob0)
    assert mid_measurements is not None
    mid_measurements[op] = sample                                                     # :473
    matrix = np.array([[(sample + 1) % 2, 0.0], [0.0, (sample) % 2]])                # :477
    state = apply_operation(_qubit_unitary(matrix, wire), state, is_state_batched)
    state = state / np.linalg.norm(state)                                             # :484
    element = op.reset and sample == 1                                                # :488
    matrix = np.array([[(element + 1) % 2, (element) % 2],
                       [(element) % 2, (element + 1) % 2]], dtype=float)
    return apply_operation(_qubit_unitary(matrix, wire), state, is_state_batched)


def apply_snapshot(op, state, is_state_batched=False, debugger=None, tape_shots=None, rng=None):
    """apply_operation.py:883-917."""
    if debugger is None or not debugger.active:
        return state
    from .measure import measure
    from .sampling import measure_with_samples

    measurement = op.hyperparameters["measurement"]
    shots = op.hyperparameters["shots"]
    if isinstance(shots, str) and shots == "workflow":
        shots = tape_shots
    if shots:
        snapshot = measure_with_samples([measurement], state, shots, is_state_batched, rng)[0]
    else:
        snapshot = measure(measurement, state, is_state_batched)
    tag = op.hyperparameters["tag"]
    if tag is None:
        debugger.snapshots[len(debugger.snapshots)] = snapshot
    elif tag not in debugger.snapshots:
        debugger.snapshots[tag] = snapshot
    elif isinstance(debugger.snapshots[tag], list):
        debugger.snapshots[tag].append(snapshot)
    else:
        debugger.snapshots[tag] = [debugger.snapshots[tag], snapshot]
    return state


def apply_operation(op, state, is_state_batched=False, mid_measurements=None, rng=None,
                    debugger=None, tape_shots=None):
    """Dispatcher — apply_operation.py:258-324 (singledispatch) and the registered kernels."""
    name = op.name
    n_dim = state.ndim
    if name == "Snapshot":
        return apply_snapshot(op, state, is_state_batched, debugger, tape_shots, rng)
    if name == "MidMeasureMP":
        return apply_mid_measure(op, state, is_state_batched, mid_measurements, rng)
    if name.startswith("Conditional") and hasattr(op, "meas_val"):
        return apply_conditional(op, state, is_state_batched, mid_measurements, rng)
    if name in ("Identity", "Barrier"):                                   # :501
        return state
    if name == "GroverOperator" and len(op.wires) >= 9:                   # :836-846
        return _apply_grover_without_matrix(state, list(op.wires), is_state_batched)
    if name == "GlobalPhase":                                             # :507-517
        phase = np.exp(-1j * np.asarray(op.data[0], dtype=complex))
        if phase.ndim > 0:
            if not is_state_batched:
                state = state.reshape((1,) + state.shape)
            phase = phase.reshape((-1,) + (1,) * (state.ndim - 1))
        return phase * state
    if name == "PauliX":                                                  # :521-524
        return np.roll(state, 1, op.wires[0] + is_state_batched)
    if name in ("PauliZ", "T", "S") or (name == "PhaseShift" and np.ndim(op.data[0]) == 0):
        # :528-609 — scale the |1> slice, stack
        axis = op.wires[0] + is_state_batched
        factor = {"PauliZ": -1, "T": np.exp(0.25j * np.pi), "S": 1j}.get(name)
        if factor is None:
            factor = np.exp(1j * np.asarray(op.data[0], dtype=complex))
        sl_0 = _get_slice(0, axis, n_dim)
        sl_1 = _get_slice(1, axis, n_dim)
        return np.stack([state[sl_0], state[sl_1] * factor], axis=axis)
    if name == "Hadamard":                                                # :613-642
        return _apply_single_qubit_np(_HADAMARD.astype(state.dtype), state,
                                      op.wires[0] + is_state_batched)
    if name in ("RX", "RY", "RZ"):                                        # :744-759
        fn = {"RX": _rx_coeffs, "RY": _ry_coeffs, "RZ": _rz_coeffs}[name]
        return _rotation_1q(op, state, is_state_batched, fn)
    if name == "CNOT":                                                    # :763-778
        target_axes = (op.wires[1] - 1 if op.wires[1] > op.wires[0] else op.wires[1]) + is_state_batched
        control_axes = op.wires[0] + is_state_batched
        sl_0 = _get_slice(0, control_axes, n_dim)
        sl_1 = _get_slice(1, control_axes, n_dim)
        state_x = np.roll(state[sl_1], 1, target_axes)
        return np.stack([state[sl_0], state_x], axis=control_axes)
    if name == "MultiControlledX" and len(op.wires) >= 9:                 # :782-832
        return _apply_mcx_big(op, state, is_state_batched)
    return _apply_operation_default(op, state, is_state_batched)


def _apply_mcx_big(op, state, is_state_batched):   # apply_operation.py:792-832
    wires = list(op.wires)
    ctrl_wires = [w + is_state_batched for w in wires[:-1]]
    cvals = op.hyperparameters.get("control_values") or [True] * len(ctrl_wires)
    roll_axes = [w for val, w in zip(cvals, ctrl_wires) if not val]
    for ax in roll_axes:
        state = np.roll(state, 1, ax)
    orig_shape = state.shape
    transpose_axes = (
        np.array([w - is_state_batched for w in range(len(orig_shape))
                  if w - is_state_batched not in wires] + [wires[-1]] + wires[:-1])
        + is_state_batched
    )
    state = np.transpose(state, transpose_axes)
    state = state.reshape((-1, 2, 2 ** (len(wires) - 1)))
    state_x = np.roll(state[:, :, -1], 1, 1)[:, :, np.newaxis]
    state = np.concatenate([state[:, :, :-1], state_x], axis=2)
    state = np.transpose(state.reshape(orig_shape), np.argsort(transpose_axes))
    for ax in roll_axes:
        state = np.roll(state, 1, ax)
    return state
