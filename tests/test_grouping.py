"""Qubit-wise-commuting partition of sampled Pauli observables (sampling.py:46-98 ->
pauli/grouping/group_observables.py:389-432, largest-first colouring).  The ORDER of the groups is
the order in which the shot budget's uniforms are consumed, so it is part of the same-seed sample
parity with default.qubit.  rustworkx is not available here: the expected partitions below are
the reference's docstring example and cases coloured by hand with its rules (nodes by descending
degree of the non-commutation graph, ties in index order; smallest free colour; groups in the
order of their lowest index)."""
import numpy as np

from oracle.sampling import compute_partition_indices
from pennylane_b200.simulate import _qwc_partition

CASES = [
    # group_observables.py:412-416 (docstring): [X0 @ Z1, Z0, X1] -> ((0,), (1, 2))
    ([{0: "X", 1: "Z"}, {0: "Z"}, {1: "X"}], [[0], [1, 2]]),
    # greedy-in-order would give [[0, 1, 3], [2]]: degrees (0, 1, 2, 1) -> order 2, 1, 3, 0 ->
    # colours 2:0, 1:1, 3:1, 0:0
    ([{}, {1: "X", 2: "Y"}, {0: "Y", 1: "Y", 2: "Z"}, {0: "X", 1: "X"}], [[0, 2], [1, 3]]),
    # all identities: one group (:389-394)
    ([{}, {}], [[0, 1]]),
    # X0, Y0, Z0, X0: a triangle plus a twin of node 0; degrees (2, 3, 3, 2) -> order 1, 2, 0, 3
    ([{0: "X"}, {0: "Y"}, {0: "Z"}, {0: "X"}], [[0, 3], [1], [2]]),
]


def test_partitions_follow_largest_first_colouring():
    for words, want in CASES:
        assert _qwc_partition(words) == want
        assert [list(t) for t in compute_partition_indices(words)] == want


def test_product_and_oracle_agree_on_random_words():
    rng = np.random.default_rng(0)
    for _ in range(300):
        words = []
        for _ in range(int(rng.integers(2, 7))):
            words.append({q: "XYZ"[int(c) - 1] for q, c in enumerate(rng.integers(0, 4, size=4)) if c})
        a = _qwc_partition(words)
        assert a == [list(t) for t in compute_partition_indices(words)]
        flat = sorted(i for g in a for i in g)
        assert flat == list(range(len(words)))
        for g in a:                                   # every group commutes qubit-wise
            for i in g:
                for j in g:
                    assert all(words[i][k] == words[j][k] for k in words[i] if k in words[j])
