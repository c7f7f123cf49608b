"""Classical-shadow estimator — the host post-processing behind ``shadow_expval``
(measurements/classical_shadow.py:490-514 → pennylane/shadows/classical_shadow.py:283-344,
``pauli_expval`` :489-547, ``median_of_means`` :466-486).  Works on the ``(T, n)`` bit and
recipe arrays only; the amplitudes are handled by ``simulate._classical_shadow``."""
from __future__ import annotations

import numpy as np

_RECIPE = {"X": 0, "Y": 1, "Z": 2, "I": -1}


def median_of_means(arr, num_batches, axis=0):
    batch_size = int(np.ceil(arr.shape[0] / num_batches))
    means = [np.mean(arr[i * batch_size:(i + 1) * batch_size], 0) for i in range(num_batches)]
    return np.median(means, axis=axis)


def pauli_expval(bits, recipes, word):
    """Per-snapshot value of each Pauli word (rows of ``word``; -1 = identity): 3^|word| times
    the parity sign when every non-identity factor was measured in its own basis, else 0."""
    T, n = recipes.shape
    b = word.shape[0]
    bits = bits.astype(np.int64)
    recipes = recipes.astype(np.int64)
    id_mask = word == -1
    hit = np.equal(recipes.reshape(T, 1, n), word.reshape(1, b, n))
    hit = np.all(np.logical_or(hit, np.tile(id_mask.reshape(1, b, n), (T, 1, 1))), axis=2)
    masked = np.where(id_mask, 0, np.tile(np.expand_dims(bits, 1), (1, b, 1)))
    parity = np.sum(masked, axis=2) % 2
    vals = np.where(hit, 1 - 2 * parity, 0) * 3 ** np.count_nonzero(np.logical_not(id_mask), axis=1)
    return vals.astype(np.float64)


def _coeffs_and_words(obs, wire_map):
    pr = obs.pauli_rep
    if pr is None:
        raise ValueError(f"Observable must have a valid pauli representation. Received {obs}")
    out = []
    for pw, c in pr.items():
        word = [-1] * len(wire_map)
        for w, ch in pw.items():
            word[wire_map.index(w)] = _RECIPE[ch]
        out.append((c, word))
    return out


def shadow_expval(bits, recipes, H, k=1, wire_map=None):
    """``ClassicalShadow(bits, recipes, wire_map).expval(H, k)`` (shadows/classical_shadow.py:283-344)."""
    wire_map = list(range(bits.shape[1])) if wire_map is None else list(wire_map)
    Hs = list(H) if isinstance(H, (list, tuple)) else [H]
    cw = [_coeffs_and_words(h, wire_map) for h in Hs]
    words = np.array([word for terms in cw for _, word in terms])
    vals = median_of_means(pauli_expval(bits, recipes, words), k, axis=0)
    vals = vals * np.array([np.real(c) for terms in cw for c, _ in terms])
    res, start = [], 0
    for terms in cw:
        res.append(np.sum(vals[start:start + len(terms)]))
        start += len(terms)
    return np.squeeze(res)
