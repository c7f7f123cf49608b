"""CPU: pin the oracle against every known-answer vector the reference's own tests hold for the
hot path (tests/golden/reference_known_answers.json; provenance per case in `source`)."""
import numpy as np
import pytest

from golden_utils import build_mp, build_op, build_tape, cx, load_cases
from oracle import adjoint_jacobian as o_adj
from oracle import simulate as o_sim
from oracle.apply_operation import apply_operation
from oracle.measure import measure
from oracle.sampling import sample_state

CASES = load_cases()


def _ids(kind):
    return [c for c in CASES if c["type"] == kind]


@pytest.mark.parametrize("case", _ids("apply"), ids=lambda c: c["id"])
def test_apply(case):
    got = apply_operation(build_op(case["op"]), cx(case["state"]))
    assert np.allclose(got, cx(case["expected"]), atol=case["atol"], rtol=0)


@pytest.mark.parametrize("case", _ids("measure"), ids=lambda c: c["id"])
def test_measure(case):
    got = measure(build_mp(case["measurement"]), cx(case["state"]))
    assert np.allclose(got, cx(case["expected"]).real if np.isrealobj(got) else cx(case["expected"]),
                       atol=case["atol"], rtol=0)


@pytest.mark.parametrize("case", _ids("simulate"), ids=lambda c: c["id"])
def test_simulate(case):
    tape = build_tape(case)
    res = o_sim.simulate(tape)
    res = res if isinstance(res, tuple) else (res,)
    for r, e in zip(res, case["expected"]):
        e = cx(e)
        assert np.allclose(r, e.real if np.isrealobj(np.asarray(r)) else e, atol=case["atol"], rtol=0)


@pytest.mark.parametrize("case", _ids("sample"), ids=lambda c: c["id"])
def test_sample_bit_exact(case):
    got = sample_state(cx(case["state"]), case["shots"], rng=np.random.default_rng(case["seed"]))
    assert got.tolist() == case["expected_samples"]


def test_adjoint_known_answers():
    for case in CASES:
        if case["type"] not in ("jacobian", "jvp", "vjp"):
            continue
        tape = build_tape(case)
        st, _ = o_sim.get_final_state(tape)
        if case["type"] == "jacobian":
            got = np.atleast_2d(np.array(o_adj.adjoint_jacobian(tape, st), dtype=float))
        elif case["type"] == "jvp":
            got = np.array(o_adj.adjoint_jvp(tape, case["tangents"], st), dtype=float)
        else:
            got = np.array(o_adj.adjoint_vjp(tape, case["cotangents"], st), dtype=float)
        assert np.allclose(got, np.array(case["expected"]), atol=case["atol"], rtol=0), case["id"]


def test_choice_restatement_is_numpy_choice():
    """The algorithm the CUDA sampler implements (cumsum, /cdf[-1], searchsorted right) is
    bit-for-bit what numpy's Generator.choice does (sampling.py:527)."""
    from oracle.sampling import choice_restated

    for n, seed in [(3, 0), (64, 1), (1000, 2), (4096, 3)]:
        p = np.random.default_rng(seed).random(n)
        p /= p.sum()
        a = np.random.default_rng(seed + 10).choice(np.arange(n), 5000, p=p)
        b = choice_restated(p.copy(), 5000, np.random.default_rng(seed + 10))
        assert np.array_equal(a, b)


def test_numpy_pairwise_sum_restatement():
    """b200q's k_np_leaf_sums / k_np_tree reproduce numpy's pairwise summation order
    (128-element leaves with 8 accumulators, then a perfect binary tree for power-of-two
    lengths).  Checked here against np.sum bit for bit, in the same arithmetic."""
    def leaf(a):
        r = a[:8].copy()
        for i in range(8, len(a) - len(a) % 8, 8):
            r += a[i:i + 8]
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for x in a[len(a) - len(a) % 8:]:
            res += x
        return res

    def np_sum_restated(a):
        n = len(a)
        if n < 8:
            res = 0.0
            for x in a:
                res += x
            return res
        if n <= 128:
            return leaf(a)
        leaves = np.array([leaf(a[i:i + 128]) for i in range(0, n, 128)])
        while len(leaves) > 1:
            leaves = leaves[0::2] + leaves[1::2]
        return leaves[0]

    rng = np.random.default_rng(0)
    for m in [0, 1, 2, 3, 5, 7, 8, 11, 14, 17]:
        a = rng.random(2**m) ** 3
        assert np_sum_restated(a) == np.sum(a), m
