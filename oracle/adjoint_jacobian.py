"""ORACLE (test infrastructure only — never imported by ``pennylane_b200``).

CPU restatement of pennylane/devices/qubit/adjoint_jacobian.py: ``adjoint_jacobian`` (:77-149),
``adjoint_jvp`` (:153-223), ``adjoint_vjp`` (:327-419) and ``operation_derivative``
(pennylane/operation.py:40-60), on duck-typed tapes (``operations``, ``observables``,
``measurements``, ``trainable_params``, ``num_preps``, ``num_wires``).
"""
from types import SimpleNamespace

import numpy as np

from .apply_operation import apply_operation
from .gates import generator_matrix, matrix_of
from .measure import _apply_observable, pauli_sentence_dot


def _dot_product_real(bra, ket, num_wires):               # adjoint_jacobian.py:36-40
    sum_axes = tuple(range(1, num_wires + 1))
    return np.real(np.sum(np.conj(bra) * ket, axis=sum_axes))


def operation_derivative(op):                              # operation.py:40-60
    return 1j * generator_matrix(op) @ matrix_of(op)


def _unitary(mat, wires):
    return SimpleNamespace(name="QubitUnitary", wires=tuple(wires), data=(mat,),
                           hyperparameters={}, ndim_params=(2,))


def _adjoint_op(op):                                       # qml.adjoint(op): conj-transpose
    return _unitary(np.conj(matrix_of(op)).T, op.wires)


def _op_param_layout(tape):
    """Index of every operation parameter in the tape's flat parameter list."""
    idx = 0
    layout = []
    for op in tape.operations:
        layout.append(list(range(idx, idx + len(op.data))))
        idx += len(op.data)
    return layout, idx


def adjoint_jacobian_state(tape):                          # adjoint_jacobian.py:43-73
    """Forward-mode Jacobian of the state itself: one derivative state per trainable parameter,
    carried through the rest of the circuit."""
    from .simulate import create_initial_state

    ops = list(tape.operations)
    has_prep = bool(ops) and hasattr(ops[0], "state_vector")
    state = create_initial_state(tape.num_wires, ops[0] if has_prep else None)
    jacobian = []
    param_idx = int(has_prep)
    for op in ops[has_prep:]:
        jacobian = [apply_operation(op, jac) for jac in jacobian]
        if len(op.data) == 1:
            if param_idx in tape.trainable_params:
                d_op_matrix = operation_derivative(op)
                jacobian.append(apply_operation(_unitary(d_op_matrix, op.wires), state))
            param_idx += 1
        state = apply_operation(op, state)
    return tuple(jac.flatten() for jac in jacobian)


def adjoint_jacobian(tape, state):                         # adjoint_jacobian.py:77-149
    n = tape.num_wires
    ket = state
    obs = list(tape.observables)
    n_obs = len(obs)
    bras = np.empty([n_obs] + [2] * n, dtype=np.complex128)
    for kk, o in enumerate(obs):
        bras[kk, ...] = 2 * _apply_observable(o, ket)
    trainable = list(tape.trainable_params)
    jac = np.zeros((n_obs, len(trainable)))
    _, n_op_params = _op_param_layout(tape)
    param_number = n_op_params - 1
    trainable_param_number = len(trainable) - 1
    # trainable observable parameters come last in the flat list and get zero columns
    while trainable_param_number >= 0 and trainable[trainable_param_number] > param_number:
        trainable_param_number -= 1
    for op in reversed(tape.operations[tape.num_preps:]):
        if op.name == "Snapshot":
            continue
        adj_op = _adjoint_op(op)
        ket = apply_operation(adj_op, ket)
        if len(op.data) == 1:
            if param_number in trainable:
                d_op_matrix = operation_derivative(op)
                ket_temp = apply_operation(_unitary(d_op_matrix, op.wires), ket)
                jac[:, trainable_param_number] = _dot_product_real(bras, ket_temp, n)
                trainable_param_number -= 1
            param_number -= 1
        else:
            param_number -= len(op.data)
        for kk in range(n_obs):
            bras[kk, ...] = apply_operation(adj_op, bras[kk, ...])
    jac = np.squeeze(jac)
    if jac.ndim == 0:
        return np.array(jac)
    if jac.ndim == 1:
        return tuple(np.array(j) for j in jac)
    return tuple(tuple(np.array(j_) for j_ in j) for j in jac)


def adjoint_jvp(tape, tangents, state):                    # adjoint_jacobian.py:153-223
    n = tape.num_wires
    ket = state
    obs = list(tape.observables)
    n_obs = len(obs)
    bras = np.empty([n_obs] + [2] * n, dtype=np.complex128)
    for i, o in enumerate(obs):
        bras[i] = _apply_observable(o, ket)
    trainable = list(tape.trainable_params)
    _, n_op_params = _op_param_layout(tape)
    param_number = n_op_params - 1
    trainable_param_number = len(trainable) - 1
    while trainable_param_number >= 0 and trainable[trainable_param_number] > param_number:
        trainable_param_number -= 1
    tangents_out = np.zeros(n_obs)
    for op in reversed(tape.operations[tape.num_preps:]):
        adj_op = _adjoint_op(op)
        ket = apply_operation(adj_op, ket)
        if len(op.data) == 1:
            if param_number in trainable:
                if not np.allclose(tangents[trainable_param_number], 0):
                    ket_temp = apply_operation(_unitary(operation_derivative(op), op.wires), ket)
                    tangents_out += (2 * _dot_product_real(bras, ket_temp, n)
                                     * tangents[trainable_param_number])
                trainable_param_number -= 1
            param_number -= 1
        else:
            param_number -= len(op.data)
        for i in range(n_obs):
            bras[i] = apply_operation(adj_op, bras[i])
    if n_obs == 1:
        return np.array(tangents_out[0])
    return tuple(np.array(t) for t in tangents_out)


def adjoint_vjp_state(tape, cotangents, state):            # adjoint_jacobian.py:240-245, 378-419
    """VJP of a tape that returns the state: bra = conj(cotangent), complex results."""
    ket = state
    bras = np.conj(np.asarray(cotangents).reshape(-1, *ket.shape))
    bras = np.squeeze(bras, axis=0)
    trainable = list(tape.trainable_params)
    _, n_op_params = _op_param_layout(tape)
    param_number = n_op_params - 1
    trainable_param_number = len(trainable) - 1
    out = np.empty(len(trainable), dtype=complex)
    for op in reversed(tape.operations[tape.num_preps:]):
        adj_op = _adjoint_op(op)
        ket = apply_operation(adj_op, ket)
        if len(op.data) == 1:
            if param_number in trainable:
                ket_temp = apply_operation(_unitary(operation_derivative(op), op.wires), ket)
                out[trainable_param_number] = np.sum(np.conj(bras) * ket_temp)
                trainable_param_number -= 1
            param_number -= 1
        else:
            param_number -= len(op.data)
        bras = apply_operation(adj_op, bras)
    return tuple(out)


def adjoint_vjp(tape, cotangents, state):                  # adjoint_jacobian.py:327-419 (unbatched)
    if tape.measurements[0].kind == "state":
        return adjoint_vjp_state(tape, cotangents, state)
    n = tape.num_wires
    ket = state
    obs = list(tape.observables)
    cots = np.atleast_1d(np.asarray(cotangents, dtype=float))
    trainable = list(tape.trainable_params)
    if np.allclose(cots, 0.0):
        return tuple(0.0 for _ in trainable)
    # bra = 2 * (sum_k cot_k O_k)|ket>        (:300-317)
    bra = np.zeros_like(ket)
    for c, o in zip(cots, obs):
        if np.allclose(c, 0.0):
            continue
        ps = getattr(o, "pauli_rep", None)
        if ps is not None:
            bra = bra + c * pauli_sentence_dot(ps, ket.reshape(-1), list(range(n))).reshape(ket.shape)
        else:
            bra = bra + c * _apply_observable(o, ket)
    bras = 2 * bra
    _, n_op_params = _op_param_layout(tape)
    param_number = n_op_params - 1
    trainable_param_number = len(trainable) - 1
    while trainable_param_number >= 0 and trainable[trainable_param_number] > param_number:
        trainable_param_number -= 1
    out = np.zeros(len(trainable))
    for op in reversed(tape.operations[tape.num_preps:]):
        adj_op = _adjoint_op(op)
        ket = apply_operation(adj_op, ket)
        if len(op.data) == 1:
            if param_number in trainable:
                ket_temp = apply_operation(_unitary(operation_derivative(op), op.wires), ket)
                out[trainable_param_number] = np.real(np.sum(np.conj(bras) * ket_temp))
                trainable_param_number -= 1
            param_number -= 1
        else:
            param_number -= len(op.data)
        bras = apply_operation(adj_op, bras)
    return tuple(out)
