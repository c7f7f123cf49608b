"""ORACLE (test infrastructure only — never imported by ``pennylane_b200``).

CPU restatement of pennylane/devices/qubit/measure.py and of the ``process_state`` methods it
calls (pennylane/measurements/probs.py:101-135, expval.py:81-118, var.py:81-115) plus
``PauliSentence.dot`` (pennylane/pauli/pauli_arithmetic.py:924-964).

Measurement processes are duck-typed: ``mp.kind`` in {"expval","var","probs","state","sample",
"counts"}, ``mp.obs`` (or None) and ``mp.wires``.
"""
import numpy as np

from .apply_operation import apply_operation
from .gates import matrix_of, observable_matrix


def flatten_state(state, num_wires):          # measure.py:35-49
    dim = 2**num_wires
    batched = state.size != dim
    return state.reshape((-1, dim) if batched else (dim,))


def probs_process_state(flat_state, wires, num_device_wires):   # probs.py:101-135
    prob = np.real(flat_state) ** 2 + np.imag(flat_state) ** 2
    wires = list(wires)
    if not wires:
        return prob
    inactive = [w for w in range(num_device_wires) if w not in wires]
    shape = [2] * num_device_wires
    desired_axes = np.argsort(np.argsort(wires))
    flat_shape = (-1,)
    batched = prob.ndim == 2
    if batched:
        shape.insert(0, prob.shape[0])
        inactive = [i + 1 for i in inactive]
        desired_axes = np.insert(desired_axes + 1, 0, 0)
        flat_shape = (prob.shape[0], -1)
    prob = prob.reshape(shape)
    prob = np.sum(prob, axis=tuple(inactive))
    prob = np.transpose(prob, desired_axes)
    return prob.reshape(flat_shape)


def diagonalizing_gates(obs):
    """The gates that rotate ``obs`` into the computational basis (each observable's
    ``compute_diagonalizing_gates`` in the reference, non_parametric_ops.py:152,419,628,860;
    observables.py:170-190; op_math/prod.py composite rule)."""
    return list(obs.diagonalizing_gates())


def eigvals(obs):
    name = obs.name
    if name in ("PauliX", "PauliY", "PauliZ", "Hadamard"):
        return np.array([1.0, -1.0])
    if name == "Identity":
        return np.ones(2 ** max(1, len(obs.wires)))
    if name == "Prod":
        ws = [w for o in obs.operands for w in o.wires]
        if len(ws) == len(set(ws)):
            ev = np.array([1.0])
            for o in obs.operands:
                ev = np.kron(ev, eigvals(o))
            return ev
    if name == "SProd":
        return obs.scalar * eigvals(obs.base)
    if name == "Hermitian":
        return np.linalg.eigh(np.asarray(obs.data[0], dtype=complex))[0]
    if name == "Projector":
        return np.asarray(obs.eigvals())
    return np.linalg.eigvalsh(observable_matrix(obs, list(obs.wires)))


def reduce_statevector(state, indices):
    """pennylane/math/quantum.py:386-487: ``einsum`` of the state with its conjugate over the
    traced wires, then the permutation to the requested wire order (``_permute_dense_matrix``).
    ``state``: ``(2^n,)`` or ``(B, 2^n)``."""
    from string import ascii_letters

    state = np.asarray(state, dtype="complex128")
    batched = state.ndim == 2
    dim = state.shape[-1]
    num_wires = int(np.log2(dim))
    consecutive = list(range(num_wires))
    st = np.reshape(state, [state.shape[0] if batched else 1] + [2] * num_wires)
    indices1 = ascii_letters[1: num_wires + 1]
    indices2 = "".join(ascii_letters[num_wires + i + 1] if i in indices else ascii_letters[i + 1]
                       for i in consecutive)
    target = "".join([ascii_letters[i + 1] for i in sorted(indices)]
                     + [ascii_letters[num_wires + i + 1] for i in sorted(indices)])
    dm = np.einsum(f"a{indices1},a{indices2}->a{target}", st, np.conj(st), optimize="greedy")
    k = len(indices)
    dm = np.reshape(dm, (-1, 2 ** k, 2 ** k))
    # _permute_dense_matrix: from sorted(indices) to the requested order
    srt = sorted(indices)
    if list(indices) != srt:
        perm = [srt.index(w) for w in indices]
        t = np.reshape(dm, [dm.shape[0]] + [2] * (2 * k))
        axes = [0] + [1 + p for p in perm] + [1 + k + p for p in perm]
        dm = np.reshape(np.transpose(t, axes), (-1, 2 ** k, 2 ** k))
    return dm if batched else dm[0]


def _compute_vn_entropy(density_matrix, base=None):   # math/quantum.py:632-663
    div_base = np.log(base) if base else 1
    evs = np.linalg.eigvalsh(density_matrix)
    evs = np.where(evs > 0, evs, 1.0)
    return np.sum(-evs * np.log(evs), axis=-1) / div_base


def density_process_state(mp, flat):
    """measurements/purity.py:50-54, vn_entropy.py:65-67, mutual_info.py:92-100 and
    ``DensityMatrixMP.process_state``.  The reference first forms the full density matrix
    (``dm_from_state_vector``) and reduces it with ``reduce_dm``; tracing |psi><psi| over the
    complement is the same contraction ``reduce_statevector`` does, which keeps this oracle
    usable beyond a dozen wires."""
    if mp.kind == "mutual_info":
        w0, w1 = (list(w) for w in mp._wires)
        base = getattr(mp, "log_base", None)
        return (_compute_vn_entropy(reduce_statevector(flat, w0), base)
                + _compute_vn_entropy(reduce_statevector(flat, w1), base)
                - _compute_vn_entropy(reduce_statevector(flat, sorted(w0 + w1)), base))
    rho = reduce_statevector(flat, list(mp.wires))
    if mp.kind == "density_matrix":
        return rho
    if mp.kind == "purity":                                # math/quantum.py:563-589
        if rho.ndim > 2:
            return np.real(np.einsum("abc,acb->a", rho, rho))
        return np.real(np.einsum("ab,ba", rho, rho))
    return _compute_vn_entropy(rho, getattr(mp, "log_base", None))


def state_diagonalizing_gates(mp, state, is_state_batched=False):   # measure.py:52-71
    if mp.obs is not None:
        for op in diagonalizing_gates(mp.obs):
            state = apply_operation(op, state, is_state_batched=is_state_batched)
    total = state.ndim - is_state_batched
    flat = flatten_state(state, total)
    wires = list(mp.wires)
    if mp.kind == "probs":
        return probs_process_state(flat, wires, total)
    if mp.kind == "state":
        return flat
    if mp.kind in ("density_matrix", "purity", "vn_entropy", "mutual_info"):
        return density_process_state(mp, flat)
    prob = probs_process_state(flat, wires, total)
    ev = np.asarray(eigvals(mp.obs), dtype="float64")
    if mp.kind == "expval":                                  # expval.py:81-118
        return np.dot(prob, ev)
    if mp.kind == "var":                                     # var.py:81-115
        return np.dot(prob, ev**2) - np.dot(prob, ev) ** 2
    raise NotImplementedError(mp.kind)


# ---- PauliSentence.dot -------------------------------------------------------------------------
_SPARSE = {  # _cached_sparse_data (pauli_arithmetic.py:63-98): (data, col index) of a 2x2 Pauli
    "I": (np.array([1.0, 1.0], dtype=complex), np.array([0, 1])),
    "X": (np.array([1.0, 1.0], dtype=complex), np.array([1, 0])),
    "Y": (np.array([-1.0j, 1.0j], dtype=complex), np.array([1, 0])),
    "Z": (np.array([1.0, -1.0], dtype=complex), np.array([0, 1])),
}


def _csr_data(word, wire_order, coeff):        # pauli_arithmetic.py:464-487
    full_word = [word.get(w, "I") for w in wire_order]
    matrix_size = 2 ** len(wire_order)
    if len(word) == 0:
        return np.full(matrix_size, coeff, dtype=np.complex128)
    data = np.empty(matrix_size, dtype=np.complex128)
    current_size = 2
    data[:current_size] = _SPARSE[full_word[-1]][0]
    data[:current_size] *= coeff
    for s in full_word[-2::-1]:
        if s in ("I", "X"):
            data[current_size: 2 * current_size] = data[:current_size]
        elif s == "Y":
            data[current_size: 2 * current_size] = 1j * data[:current_size]
            data[:current_size] *= -1j
        elif s == "Z":
            data[current_size: 2 * current_size] = -data[:current_size]
        current_size *= 2
    return data


def _csr_indices(word, wire_order):            # pauli_arithmetic.py:498-520
    full_word = [word.get(w, "I") for w in wire_order]
    matrix_size = 2 ** len(wire_order)
    if len(word) == 0:
        return np.arange(matrix_size)
    indices = np.empty(matrix_size, dtype=np.int64)
    current_size = 2
    indices[:current_size] = _SPARSE[full_word[-1]][1]
    for s in full_word[-2::-1]:
        if s in ("I", "Z"):
            indices[current_size: 2 * current_size] = indices[:current_size] + current_size
        else:
            indices[current_size: 2 * current_size] = indices[:current_size]
            indices[:current_size] += current_size
        current_size *= 2
    return indices


def pauli_sentence_dot(ps, vector, wire_order):   # pauli_arithmetic.py:924-950
    words = list(ps)
    structure = [tuple(1 if w.get(x, "I") in "XY" else 0 for x in wire_order) for w in words]
    uniq = sorted(set(structure))                  # np.unique(axis=0) sorts rows
    if vector.ndim == 1:
        vector = vector.reshape(1, -1)
    mv = np.zeros_like(vector, dtype=np.complex128)
    for st in uniq:
        group = [w for w, s in zip(words, structure) if s == st]
        entries = _csr_indices(group[0], wire_order)
        data = np.zeros(vector.shape[1], dtype=np.complex128)
        for w in group:
            data = data + _csr_data(w, wire_order, complex(ps[w]))
        mv += vector[:, entries] * data.reshape(1, -1)
    return mv


def csr_dot_products(mp, state, is_state_batched=False):   # measure.py:74-118 (Pauli branch)
    total_wires = state.ndim - is_state_batched
    st = state.reshape(state.shape[0], -1) if is_state_batched else state.reshape(1, -1)
    bra = np.conj(st)
    ps = mp.obs.pauli_rep
    new_ket = pauli_sentence_dot(ps, st, list(range(total_wires)))
    res = (bra * new_ket).sum(axis=1)
    return np.real(np.squeeze(res))


def full_dot_products(mp, state, is_state_batched=False):   # measure.py:121-139
    ket = _apply_observable(mp.obs, state, is_state_batched)
    dot = np.sum(np.conj(state) * ket, axis=tuple(range(int(is_state_batched), state.ndim)))
    return np.real(dot)


class _MatOp:
    """QubitUnitary-like carrier so an observable matrix goes through apply_operation's
    default dense path, as ``apply_operation(obs, state)`` does in the reference."""
    name = "QubitUnitary"
    hyperparameters = {}
    ndim_params = (2,)

    def __init__(self, mat, wires):
        self.data = (mat,)
        self.wires = tuple(wires)


def _apply_observable(obs, state, is_state_batched=False):
    wires = list(obs.wires)
    if not wires:
        return observable_matrix(obs, [0])[0, 0] * state
    return apply_operation(_MatOp(observable_matrix(obs, wires), wires), state, is_state_batched)


def sum_of_terms_method(mp, state, is_state_batched=False):   # measure.py:142-161
    from types import SimpleNamespace

    cs, os_ = mp.obs.terms()
    return sum(c * measure(SimpleNamespace(kind="expval", obs=o, wires=o.wires), state,
                           is_state_batched) for c, o in zip(cs, os_))


def csr_dot_products_sparse(mp, state, is_state_batched=False):   # measure.py:100-118 (scipy branch)
    from scipy.sparse import csr_matrix

    total_wires = state.ndim - is_state_batched
    Hmat = mp.obs.sparse_matrix(wire_order=list(range(total_wires)))
    if is_state_batched:
        st = state.reshape(state.shape[0], -1)
        bra = csr_matrix(np.conj(st))
        ket = csr_matrix(st)
        new_bra = bra.dot(Hmat)
        res = np.asarray(new_bra.multiply(ket).sum(axis=1)).reshape(-1)
    else:
        st = state.flatten()
        bra = csr_matrix(np.conj(st))
        ket = csr_matrix(st[..., None])
        new_ket = csr_matrix.dot(Hmat, ket)
        res = csr_matrix.dot(bra, new_ket).toarray()[0]
    return np.real(np.squeeze(res))


def get_measurement_function(mp, state):          # measure.py:165-221
    if mp.kind in ("expval",) and mp.obs is not None:
        name = mp.obs.name
        if name == "SparseHamiltonian":                                  # :198-199
            return csr_dot_products_sparse
        if name == "Hermitian":
            return full_dot_products
        if name in ("LinearCombination", "Hamiltonian"):
            if mp.obs.pauli_rep is None:
                return sum_of_terms_method
            return csr_dot_products
        if name == "Sum":
            if mp.obs.pauli_rep is None:
                return sum_of_terms_method
            ws = [w for o in mp.obs.operands for w in o.wires]
            if len(ws) != len(set(ws)) and len(mp.obs.wires) > 7:
                return csr_dot_products
            if not _has_diag_gates(mp.obs):
                return csr_dot_products
    if mp.kind in ("expval", "var", "probs", "state", "density_matrix", "purity", "vn_entropy",
                   "mutual_info"):
        if mp.obs is None or _has_diag_gates(mp.obs):
            return state_diagonalizing_gates
        if mp.kind == "expval":
            return full_dot_products
    raise NotImplementedError(f"no analytic measurement for {mp.kind} of {mp.obs}")


def _has_diag_gates(obs):
    return bool(getattr(obs, "has_diagonalizing_gates", False))


def measure(mp, state, is_state_batched=False):   # measure.py:224-239
    return get_measurement_function(mp, state)(mp, state, is_state_batched)
