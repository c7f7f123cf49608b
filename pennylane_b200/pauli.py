"""Minimal Pauli-word / Pauli-sentence algebra used to hand observables to the kernels.

Mirrors the data model of ``pennylane/pauli/pauli_arithmetic.py`` (``PauliWord`` :165 is a
mapping wire -> "X"/"Y"/"Z", ``PauliSentence`` :539 a mapping word -> coefficient) closely
enough that a real PennyLane ``op.pauli_rep`` can be consumed by the same code (duck typing:
``for word, coeff in ps.items(): for wire, char in word.items()``).
"""
from __future__ import annotations

from typing import Iterable

_MUL = {  # (a, b) -> (phase, c) with a*b = phase * c
    ("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"),
    ("Y", "Z"): (1j, "X"), ("Z", "Y"): (-1j, "X"),
    ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y"),
}


class PauliWord(dict):
    """Immutable mapping wire -> 'X' | 'Y' | 'Z' (identity factors are not stored)."""

    def __init__(self, mapping=()):
        super().__init__({w: c for w, c in dict(mapping).items() if c != "I"})
        self._hash = hash(frozenset(self.items()))

    def __hash__(self):
        return self._hash

    def __setitem__(self, k, v):  # pragma: no cover - defensive
        raise TypeError("PauliWord is immutable")

    @property
    def wires(self):
        return tuple(self.keys())

    def mul(self, other: "PauliWord"):
        """Return (phase, word) with self * other = phase * word."""
        out = dict(self)
        phase = 1.0 + 0j
        for w, c in other.items():
            if w not in out:
                out[w] = c
            elif out[w] == c:
                del out[w]
            else:
                ph, r = _MUL[(out[w], c)]
                phase *= ph
                out[w] = r
        return phase, PauliWord(out)


class PauliSentence(dict):
    """Mapping PauliWord -> complex coefficient."""

    def add_term(self, word: PauliWord, coeff):
        self[word] = self.get(word, 0.0) + coeff

    def __add__(self, other):
        out = PauliSentence(self)
        for w, c in other.items():
            out.add_term(w, c)
        return out

    def scale(self, s):
        return PauliSentence({w: c * s for w, c in self.items()})

    def __matmul__(self, other):
        out = PauliSentence()
        for w1, c1 in self.items():
            for w2, c2 in other.items():
                ph, w = w1.mul(w2)
                out.add_term(w, ph * c1 * c2)
        return out

    @property
    def wires(self):
        seen = []
        for w in self:
            for x in w:
                if x not in seen:
                    seen.append(x)
        return tuple(seen)


def word_masks(word, wire_to_bit) -> tuple[int, int, int]:
    """(xmask, zmask, ny) of a Pauli word for the kernels: xmask covers X and Y factors,
    zmask covers Z and Y factors, ny counts Y factors
    (P|j> = i^ny (-1)^popcount(j & zmask) |j ^ xmask>)."""
    xm = zm = ny = 0
    for wire, ch in word.items():
        b = 1 << wire_to_bit[wire]
        if ch == "X":
            xm |= b
        elif ch == "Z":
            zm |= b
        elif ch == "Y":
            xm |= b
            zm |= b
            ny += 1
        elif ch != "I":  # pragma: no cover
            raise ValueError(f"bad Pauli character {ch!r}")
    return xm, zm, ny


def sentence_terms(ps, wire_to_bit, tol: float = 0.0):
    """Flatten a Pauli sentence into kernel term lists (xmasks, zmasks, nys, coeffs)."""
    xs, zs, ys, cs = [], [], [], []
    for word, coeff in ps.items():
        if tol and abs(coeff) <= tol:
            continue
        xm, zm, ny = word_masks(word, wire_to_bit)
        xs.append(xm); zs.append(zm); ys.append(ny); cs.append(coeff)
    return xs, zs, ys, cs


def is_pauli_word_like(ps) -> bool:
    return ps is not None and len(ps) == 1


def pauli_word_from_string(s: str, wires: Iterable) -> PauliWord:
    return PauliWord({w: c for w, c in zip(wires, s) if c != "I"})
