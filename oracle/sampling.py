"""ORACLE (test infrastructure only — never imported by ``pennylane_b200``).

CPU restatement of pennylane/devices/qubit/sampling.py: ``sample_state`` (:439-476),
``_sample_probs_numpy`` (:500-531) — which IS ``numpy.random.Generator.choice`` — the grouping of
measurements (:46-98) and ``measure_with_samples`` (:205-335, :377-436).
"""
import numpy as np

from .apply_operation import apply_operation
from .measure import diagonalizing_gates, eigvals, flatten_state, probs_process_state

_PROB_NORMALISATION_TOLERANCE = 1e-6    # sampling.py:33


def _sample_probs_numpy(probs, shots, num_wires, is_state_batched, rng):   # sampling.py:500-531
    rng = np.random.default_rng(rng)
    norm = np.sum(probs, axis=-1)
    norm_err = np.abs(norm - 1.0)
    norm_err = norm_err if is_state_batched else norm_err[..., np.newaxis]
    if np.any(norm_err > _PROB_NORMALISATION_TOLERANCE):
        raise ValueError("probabilities do not sum to 1")
    basis_states = np.arange(2**num_wires)
    if is_state_batched:
        probs = probs / norm[:, np.newaxis] if norm.shape else probs / norm
        samples = np.stack([rng.choice(basis_states, shots, p=p) for p in probs])
    else:
        probs = probs / norm
        samples = rng.choice(basis_states, shots, p=probs)
    powers_of_two = 1 << np.arange(num_wires, dtype=np.int64)[::-1]
    states_sampled_base_ten = samples[..., None] & powers_of_two
    return (states_sampled_base_ten > 0).astype(np.int64)


def choice_restated(probs, shots, rng):
    """What ``Generator.choice(arange(N), shots, p=probs)`` computes (numpy/random/_generator.pyx,
    ``choice`` with replacement and ``p``): used to pin the CUDA sampler's algorithm."""
    cdf = probs.cumsum()
    cdf /= cdf[-1]
    u = rng.random(shots)
    return cdf.searchsorted(u, side="right")


def sample_state(state, shots, is_state_batched=False, wires=None, rng=None):   # :439-476
    total_indices = state.ndim - is_state_batched
    wires_to_sample = list(wires) if wires else list(range(total_indices))
    flat_state = flatten_state(state, total_indices)
    probs = probs_process_state(flat_state, wires_to_sample
                                if wires_to_sample != list(range(total_indices)) else [],
                                total_indices)
    return _sample_probs_numpy(probs, shots, len(wires_to_sample), is_state_batched, rng)


def process_samples(mp, samples, wire_order):
    """measurements/{sample,expval,var,probs,counts}.py ``process_samples`` for one shot bin.
    ``samples``: (shots, n_wires) ints (or with a leading batch axis)."""
    wire_order = list(wire_order)
    wires = list(mp.wires) if len(mp.wires) else wire_order
    cols = [wire_order.index(w) for w in wires]
    sub = samples[..., cols]
    if mp.obs is None:
        if mp.kind == "sample":
            return sub
        powers = 2 ** np.arange(len(wires))[::-1]
        idx = sub @ powers
        if mp.kind == "probs":
            dim = 2 ** len(wires)
            if idx.ndim == 1:
                return np.bincount(idx, minlength=dim) / idx.shape[0]
            return np.stack([np.bincount(i, minlength=dim) / i.shape[0] for i in idx])
        if mp.kind == "counts":
            out = {}
            for i in idx:
                key = format(int(i), f"0{len(wires)}b")
                out[key] = out.get(key, 0) + 1
            return out
        raise NotImplementedError(mp.kind)
    ev = np.asarray(eigvals(mp.obs))
    powers = 2 ** np.arange(len(wires))[::-1]
    idx = sub @ powers
    vals = ev[idx]
    if mp.kind == "sample":
        return vals
    if mp.kind == "expval":
        return np.squeeze(np.mean(vals, axis=-1))
    if mp.kind == "var":
        return np.squeeze(np.var(vals, axis=-1))
    if mp.kind == "counts":
        out = {}
        for v in vals:
            out[float(v)] = out.get(float(v), 0) + 1
        return out
    raise NotImplementedError(mp.kind)


def _pauli_word_of(obs):
    ps = getattr(obs, "pauli_rep", None)
    if ps is None or len(ps) != 1:
        return None
    (w, c), = ps.items()
    return w


def _qwc(w1, w2):
    return all(w1[k] == w2[k] for k in w1 if k in w2)


def compute_partition_indices(words):
    """pauli/grouping/group_observables.py:389-432 with grouping_type 'qwc' and method 'lf', on
    Pauli words given as {wire: 'X'|'Y'|'Z'}.  The adjacency matrix of the complement graph
    (:340-386: entry (i, j) set when some wire carries two different non-identity factors), its
    largest-first greedy colouring (rustworkx ``graph_greedy_color``: nodes by descending degree,
    stable; smallest colour absent from the neighbours), and the grouping of :238-243 (colours in
    the order of their lowest index, indices ascending)."""
    m = len(words)
    if all(len(w) == 0 for w in words):
        return (tuple(range(m)),)
    wires = sorted({k for w in words for k in w}, key=str)
    code = {"X": 1, "Y": 2, "Z": 3}
    P = np.array([[code.get(w.get(k), 0) for k in wires] for w in words], dtype=np.int8)
    Pb = P[:, None]
    adj = np.logical_or.reduce((P * Pb) * (P - Pb), axis=2)
    degree = adj.sum(axis=1)
    order = np.argsort(-degree, kind="stable")
    colours = -np.ones(m, dtype=int)
    for i in order:
        taken = set(colours[np.nonzero(adj[i])[0]].tolist())
        c = 0
        while c in taken:
            c += 1
        colours[i] = c
    parts = {}
    for i in range(m):
        parts.setdefault(int(colours[i]), []).append(i)
    return tuple(tuple(v) for v in parts.values())


def _group_measurements(mps):                  # sampling.py:46-98
    if len(mps) == 1:
        return [mps], [[0]]
    pauli, other, other_idx, no_obs, no_obs_idx = [], [], [], [], []
    for i, mp in enumerate(mps):
        if mp.kind in ("shadow", "shadow_expval"):                        # :68-70
            other.append([mp]); other_idx.append([i])
        elif mp.obs is None:
            no_obs.append(mp); no_obs_idx.append(i)
        elif _pauli_word_of(mp.obs) is not None and mp.obs.name not in ("LinearCombination", "Sum", "Hamiltonian"):
            pauli.append((i, mp))
        else:
            other.append([mp]); other_idx.append([i])
    groups, gidx = [], []
    if pauli:
        for part in compute_partition_indices([_pauli_word_of(mp.obs) for _, mp in pauli]):
            groups.append([pauli[k][1] for k in part])
            gidx.append([pauli[k][0] for k in part])
    if no_obs:
        groups.append(no_obs); gidx.append(no_obs_idx)
    return groups + other, gidx + other_idx


def _apply_diagonalizing_gates(mps, state, is_state_batched=False):   # sampling.py:188-202
    from types import SimpleNamespace

    if len(mps) == 1:
        gates = diagonalizing_gates(mps[0].obs) if mps[0].obs is not None else []
    elif all(mp.obs is not None for mp in mps):
        # pauli/utils.py:1059-1081 (diagonalize_qwc_pauli_words): one rotation per wire
        full = {}
        for mp in mps:
            for wire, ch in _pauli_word_of(mp.obs).items():
                full.setdefault(wire, ch)
        gates = []
        for w, ch in full.items():
            if ch == "X":
                gates.append(SimpleNamespace(name="RY", wires=(w,), data=(-np.pi / 2,),
                                             hyperparameters={}))
            elif ch == "Y":
                gates.append(SimpleNamespace(name="RX", wires=(w,), data=(np.pi / 2,),
                                             hyperparameters={}))
    else:
        gates = []
    for op in gates:
        state = apply_operation(op, state, is_state_batched=is_state_batched)
    return state


def shot_bins(shot_vector):
    lower = 0
    for s, copies in shot_vector:
        for _ in range(copies):
            yield lower, lower + s
            lower += s


def measure_with_samples(mps, state, shots, is_state_batched=False, rng=None,
                         mid_measurements=None):                                  # :205-273
    """``shots``: object with ``total_shots``, ``shot_vector`` [(shots, copies)...],
    ``has_partitioned_shots``."""
    mps = list(mps)
    if mid_measurements:                                                         # :235-236
        mps = mps[0: len(mps) - len(mid_measurements)]
    groups, indices = _group_measurements(mps)
    all_res = []
    for group in groups:
        mp0 = group[0]
        if mp0.kind == "expval" and mp0.obs is not None and mp0.obs.name in (
                "LinearCombination", "Hamiltonian", "Sum"):
            all_res.extend(_measure_sum_with_samples(group, state, shots, is_state_batched, rng))
        elif mp0.kind in ("shadow", "shadow_expval"):
            all_res.extend(_measure_classical_shadow(group, state, shots, rng))
        else:
            all_res.extend(_measure_with_samples_diagonalizing_gates(
                group, state, shots, is_state_batched, rng))
    flat_indices = [i for idx in indices for i in idx]
    sorted_res = tuple(res for _, res in sorted(enumerate(all_res), key=lambda r: flat_indices[r[0]]))
    if mid_measurements:                                                         # :266-267
        sorted_res += tuple(mid_measurements.values())
    if shots.has_partitioned_shots:
        sorted_res = tuple(zip(*sorted_res))
    return sorted_res


def classical_shadow_process_state_with_shots(mp, state, shots, rng=None):
    """measurements/classical_shadow.py:142-257 (``ClassicalShadowMP.process_state_with_shots``):
    the same stacked-state einsum walk over the measured wires."""
    from string import ascii_letters

    mapped_wires = list(mp.wires)
    n_qubits = len(mapped_wires)
    num_dev_qubits = len(state.shape)
    recipe_rng = np.random.RandomState(mp.seed)
    recipes = recipe_rng.randint(0, 3, size=(shots, n_qubits))
    bit_rng = np.random.default_rng(rng)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Y = np.array([[0, -1j], [1j, 0]])
    Z = np.array([[1, 0], [0, -1]], dtype=complex)
    H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
    rz = np.array([[np.exp(-0.5j * (-np.pi / 2)), 0], [0, np.exp(0.5j * (-np.pi / 2))]])
    obs_list = np.stack([X, Y, Z])
    diag_list = np.stack([H, H @ rz, np.eye(2, dtype=complex)])
    obs = obs_list[recipes]
    diagonalizers = diag_list[recipes]
    unmeasured = [i for i in range(num_dev_qubits) if i not in mapped_wires]
    transposed_state = np.transpose(state, axes=mapped_wires + unmeasured)
    outcomes = np.zeros((shots, n_qubits))
    stacked_state = np.repeat(transposed_state[np.newaxis, ...], shots, axis=0)
    for active_qubit in range(n_qubits):
        num_remaining = num_dev_qubits - active_qubit
        conj_first = ascii_letters[num_remaining]
        stacked_dim = ascii_letters[num_remaining + 1]
        state_str = f"{stacked_dim}{ascii_letters[:num_remaining]}"
        conj_state_str = f"{stacked_dim}{conj_first}{ascii_letters[1:num_remaining]}"
        target_str = f"{stacked_dim}a{conj_first}"
        first_qubit_state = np.einsum(f"{state_str},{conj_state_str}->{target_str}",
                                      stacked_state, np.conj(stacked_state))
        probs = (np.einsum("abc,acb->a", first_qubit_state, obs[:, active_qubit]) + 1) / 2
        samples = bit_rng.random(size=probs.shape) > np.real(probs)
        outcomes[:, active_qubit] = samples
        rotated_state = np.einsum("ab...,acb->ac...", stacked_state, diagonalizers[:, active_qubit])
        stacked_state = rotated_state[np.arange(shots), samples.astype(np.int8)]
        sum_indices = tuple(range(1, num_remaining))
        state_squared = np.abs(stacked_state) ** 2
        norms = np.sqrt(np.sum(state_squared, sum_indices, keepdims=True))
        stacked_state /= norms
    return np.stack([outcomes, recipes]).astype(np.int8)


def _median_of_means(arr, num_batches, axis=0):            # shadows/classical_shadow.py:466-486
    batch_size = int(np.ceil(arr.shape[0] / num_batches))
    means = [np.mean(arr[i * batch_size: (i + 1) * batch_size], 0) for i in range(num_batches)]
    return np.median(means, axis=axis)


def _pauli_expval(bits, recipes, word):                    # shadows/classical_shadow.py:489-547
    T, n = recipes.shape
    b = word.shape[0]
    bits, recipes = bits.astype(np.int64), recipes.astype(np.int64)
    id_mask = word == -1
    indices = np.equal(np.reshape(recipes, (T, 1, n)), np.reshape(word, (1, b, n)))
    indices = np.logical_or(indices, np.tile(np.reshape(id_mask, (1, b, n)), (T, 1, 1)))
    indices = np.all(indices, axis=2)
    bits = np.where(id_mask, 0, np.tile(np.expand_dims(bits, 1), (1, b, 1)))
    bits = np.sum(bits, axis=2) % 2
    expvals = np.where(indices, 1 - 2 * bits, 0) * 3 ** np.count_nonzero(
        np.logical_not(id_mask), axis=1)
    return expvals.astype(np.float64)


def shadow_expval_process_state_with_shots(mp, state, shots, rng=None):
    """measurements/classical_shadow.py:490-514 with ``ClassicalShadow.expval``
    (shadows/classical_shadow.py:238-250, 283-344)."""
    from types import SimpleNamespace

    wire_map = list(mp.wires)
    bits, recipes = classical_shadow_process_state_with_shots(
        SimpleNamespace(wires=mp.wires, seed=mp.seed), state, shots, rng=rng)
    Hs = list(mp.H) if isinstance(mp.H, (list, tuple)) else [mp.H]
    to_recipe = {"X": 0, "Y": 1, "Z": 2, "I": -1}
    coeffs_and_words = []
    for h in Hs:
        cw = []
        for pw, c in h.pauli_rep.items():
            word = [-1] * bits.shape[1]
            for i, ch in pw.items():
                word[wire_map.index(i)] = to_recipe[ch]
            cw.append((c, word))
        coeffs_and_words.append(cw)
    expvals = _pauli_expval(bits, recipes,
                            np.array([word for cw in coeffs_and_words for _, word in cw]))
    expvals = _median_of_means(expvals, mp.k, axis=0)
    expvals = expvals * np.array([np.real(c) for cw in coeffs_and_words for c, _ in cw])
    start, results = 0, []
    for cw in coeffs_and_words:
        results.append(np.sum(expvals[start: start + len(cw)]))
        start += len(cw)
    return np.squeeze(results)


def _measure_classical_shadow(mps, state, shots, rng):                       # sampling.py:338-374
    mp = mps[0]
    if mp.kind == "shadow_expval":
        if shots.has_partitioned_shots:
            return [tuple(shadow_expval_process_state_with_shots(mp, state, s, rng)
                          for s, copies in shots.shot_vector for _ in range(copies))]
        return [shadow_expval_process_state_with_shots(mp, state, shots.total_shots, rng)]
    if shots.has_partitioned_shots:
        return [tuple(classical_shadow_process_state_with_shots(mp, state, s, rng)
                      for s, copies in shots.shot_vector for _ in range(copies))]
    return [classical_shadow_process_state_with_shots(mp, state, shots.total_shots, rng)]


def _measure_with_samples_diagonalizing_gates(mps, state, shots, is_state_batched, rng):  # :276-335
    state = _apply_diagonalizing_gates(mps, state, is_state_batched)
    total_indices = state.ndim - is_state_batched
    wires = list(range(total_indices))
    try:
        samples = sample_state(state, shots.total_shots, is_state_batched, wires=wires, rng=rng)
    except ValueError as e:
        if "probabilities contain nan" not in str(e).lower():
            raise
        samples = np.zeros((shots.total_shots, len(wires)), dtype=np.int64)
    processed = []
    for lower, upper in shot_bins(shots.shot_vector):
        processed.append(tuple(process_samples(mp, samples[..., lower:upper, :], wires) for mp in mps))
    if shots.has_partitioned_shots:
        return tuple(zip(*processed))
    return processed[0]


class _OneShots:
    has_partitioned_shots = False

    def __init__(self, s):
        self.total_shots = s
        self.shot_vector = [(s, 1)]


def _measure_sum_with_samples(mps, state, shots, is_state_batched, rng):   # :377-436
    from types import SimpleNamespace

    mp = mps[0]
    cs, os_ = mp.obs.terms()

    def one(s):
        res = measure_with_samples(
            [SimpleNamespace(kind="expval", obs=o, wires=o.wires) for o in os_], state,
            _OneShots(s), is_state_batched, rng)
        return sum(c * r for c, r in zip(cs, res))

    unsq = tuple(one(s) for s, copies in shots.shot_vector for _ in range(copies))
    return [unsq] if shots.has_partitioned_shots else [unsq[0]]
