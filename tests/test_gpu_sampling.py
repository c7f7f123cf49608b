"""GPU parity: the shot sampler against numpy's Generator.choice (the reference's sampler,
sampling.py:500-531) — bit-exact under a fixed seed in exact mode.  Mirrors
tests/devices/qubit/test_sampling.py (seed-pinned vector :137-142, shapes, shot vectors)."""
import numpy as np
import pytest

from conftest import random_state

pytestmark = pytest.mark.gpu


def _sv(state, batched=False):
    from pennylane_b200 import StateVector

    n = state.ndim - (1 if batched else 0)
    sv = StateVector(n)
    sv.set_state(state)
    return sv


def test_reference_golden_vector():
    """tests/devices/qubit/test_sampling.py:137-142: default_rng(12345), two-qubit state
    [[0, 1j], [-1, 0]]/sqrt(2), 4 shots -> [[0,1],[0,1],[1,0],[1,0]]."""
    state = np.array([[0, 1j], [-1, 0]], dtype=np.complex128) / np.sqrt(2)
    sv = _sv(state)
    got = sv.sample(4, np.random.default_rng(12345))
    assert got.dtype == np.int64
    assert np.array_equal(got, np.array([[0, 1], [0, 1], [1, 0], [1, 0]]))


@pytest.mark.parametrize("n", [1, 2, 3, 6, 7, 8, 11, 12, 16, 20])
def test_bit_exact_against_numpy_choice(n):
    from oracle.sampling import sample_state

    state = random_state(n, seed=n)
    sv = _sv(state)
    shots = 5000
    got = sv.sample(shots, np.random.default_rng(42 + n))
    ref = sample_state(state, shots, rng=np.random.default_rng(42 + n))
    assert got.shape == (shots, n)
    assert np.array_equal(got, ref)


def test_rng_stream_advances_like_reference():
    """default_qubit.py:798 threads ONE Generator through all executions: two consecutive calls
    must consume the stream exactly like two numpy choice calls."""
    from oracle.sampling import sample_state

    n = 9
    state = random_state(n, seed=1)
    sv = _sv(state)
    r1, r2 = np.random.default_rng(7), np.random.default_rng(7)
    a1, a2 = sv.sample(100, r1), sv.sample(333, r1)
    b1, b2 = sample_state(state, 100, rng=r2), sample_state(state, 333, rng=r2)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


def test_marginal_wires_and_batched_state():
    from oracle.sampling import sample_state

    n, B = 6, 3
    state = random_state(n, seed=3, batch=B)
    sv = _sv(state, batched=True)
    got = sv.sample(200, np.random.default_rng(5))
    ref = sample_state(state, 200, is_state_batched=True, rng=np.random.default_rng(5))
    assert got.shape == (B, 200, n)
    assert np.array_equal(got, ref)
    single = random_state(n, seed=8)
    sv1 = _sv(single)
    got = sv1.sample(300, np.random.default_rng(6), wires=[4, 1])
    ref = sample_state(single, 300, wires=[4, 1], rng=np.random.default_rng(6))
    assert np.array_equal(got, ref)


def test_fast_mode_statistics_and_sortedness():
    """Fast (parallel scan) CDF: same distribution; indices from sorted uniforms are sorted."""
    n = 10
    state = random_state(n, seed=11)
    sv = _sv(state)
    shots = 200000
    got = sv.sample(shots, np.random.default_rng(3), exact=False)
    idx = got @ (1 << np.arange(n)[::-1])
    freq = np.bincount(idx, minlength=2**n) / shots
    p = np.abs(state.reshape(-1)) ** 2
    assert np.max(np.abs(freq - p)) < 5 * np.sqrt(p.max() / shots) + 1e-3

    class SortedRng:
        def random(self, k):
            return np.sort(np.random.default_rng(0).random(k))

    s = sv.sample(1000, SortedRng(), exact=False)
    idx = s @ (1 << np.arange(n)[::-1])
    assert np.all(np.diff(idx) >= 0)
    # exact and fast agree except (possibly) on a vanishing fraction of boundary shots
    a = sv.sample(50000, np.random.default_rng(9), exact=True)
    b = sv.sample(50000, np.random.default_rng(9), exact=False)
    assert np.mean(np.any(a != b, axis=1)) < 1e-3


def test_basis_state_and_unnormalised_errors():
    from pennylane_b200 import StateVector

    sv = StateVector(5)
    sv.reset(0b10110)
    got = sv.sample(50, np.random.default_rng(0))
    assert np.array_equal(got, np.tile([1, 0, 1, 1, 0], (50, 1)))
    sv.apply_phase(1.1)          # norm^2 = 1.21 -> sampling.py:514-519
    with pytest.raises(ValueError, match="probabilities do not sum to 1"):
        sv.sample(10, np.random.default_rng(0))
    sv.apply_phase(float("nan"))  # sampling.py:322-325 -> zeros, no exception
    assert np.array_equal(sv.sample(7, np.random.default_rng(0)), np.zeros((7, 5), dtype=np.int64))


def test_measure_with_samples_matches_oracle():
    """sampling.py:205-335 through the tape-level API, incl. shot vectors and obs samples."""
    from oracle import simulate as o_sim

    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    n = 5
    rng = np.random.default_rng(0)
    ops_ = [q.Rot(*rng.uniform(0, 6, 3), wires=i) for i in range(n)] + \
           [q.CNOT(wires=[i, i + 1]) for i in range(n - 1)]
    cases = [
        ([qb.sample(wires=range(n))], 100),
        ([qb.sample(wires=[3, 0])], 100),
        ([qb.expval(q.PauliZ(wires=1))], 1000),
        ([qb.expval(q.PauliX(wires=1)), qb.var(q.PauliY(wires=2))], 500),
        ([qb.probs(wires=[0, 2]), qb.expval(q.PauliZ(wires=4))], 400),
        ([qb.expval(q.PauliX(wires=0) @ q.PauliY(wires=3))], [100, (50, 2)]),
        ([qb.sample(q.PauliX(wires=2))], 64),
        ([qb.counts(wires=[1, 2])], 300),
        ([qb.expval(q.LinearCombination([0.3, -1.2], [q.PauliZ(wires=0), q.PauliX(wires=1) @ q.PauliX(wires=2)]))], 250),
    ]
    for mps, shots in cases:
        tape = qb.QuantumScript(ops_, mps, shots=shots)
        got = qb.B200Qubit(wires=n, seed=99).execute(tape)
        ref = o_sim.simulate(tape, rng=np.random.default_rng(99))
        _assert_equal_nested(got, ref)


def _assert_equal_nested(a, b):
    if isinstance(a, dict):
        assert a == b
    elif isinstance(a, (tuple, list)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _assert_equal_nested(x, y)
    else:
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape
        if a.dtype.kind in "iu":
            assert np.array_equal(a, b)
        else:
            assert np.allclose(a, b, atol=1e-12, rtol=0)


def _cumsum_dev(p, mode, carry=None):
    """b200q_cumsum on a device copy of ``p`` (length a power of two)."""
    import ctypes as C

    import torch

    from pennylane_b200 import StateVector
    from pennylane_b200._lib import check

    m = int(np.log2(p.size))
    assert 1 << m == p.size
    sv = StateVector(12)
    w, wb = sv.workspace(((1 << m) // 2048 + 128) * 48 + (4 << 20))
    d = torch.from_numpy(np.ascontiguousarray(p, dtype=np.float64)).to(sv.device)
    c = None if carry is None else torch.tensor([carry], dtype=torch.float64, device=sv.device)
    check(sv.lib.b200q_cumsum(C.c_void_p(d.data_ptr()), m, mode,
                              None if c is None else C.c_void_p(c.data_ptr()), w, wb, sv.stream))
    return d.cpu().numpy()


@pytest.mark.parametrize("m", [14, 17, 20])
def test_parallel_exact_cumsum_is_numpy_bit_for_bit(m):
    """The parallel exact scan (integer significand arithmetic per binade, sample.cuh) reproduces
    numpy's sequential float64 additions (sampling.py:527 -> Generator.choice -> p.cumsum()) bit
    for bit, also on inputs built to hit ties, zeros, subnormals, huge dynamic range and many
    binade crossings; the serial kernel is the second witness."""
    rng = np.random.default_rng(m)
    N = 1 << m
    cases = {}
    x = rng.random(N); cases["uniform"] = x / x.sum()
    x = rng.random(N) ** 9; cases["skewed"] = x / x.sum()
    cases["ties_power_of_two"] = np.full(N, 2.0 ** -(m + 3))
    x = np.where(rng.random(N) < 0.7, 0.0, rng.random(N)); cases["mostly_zero"] = x / x.sum()
    x = np.exp(rng.normal(0, 14, N)); cases["wide_range"] = x / x.sum()
    x = np.concatenate([np.full(N // 2, 5e-324), rng.random(N // 4) * 1e-300, rng.random(N // 4)])
    cases["subnormal_start"] = x
    cases["grid_ties"] = rng.integers(0, 8, N) * 2.0 ** -(m + 6)
    x = np.abs(np.fft.fft(rng.normal(size=N))) ** 2; cases["abs2_of_amplitudes"] = x / x.sum()
    x = np.zeros(N); x[N // 3] = 0.25; x[N // 2] = 0.75; cases["two_spikes"] = x
    for name, p in cases.items():
        for carry in (None, 0.3125, 1e-9):
            ref = np.cumsum(p) if carry is None else np.cumsum(np.concatenate([[carry], p]))[1:]
            par = _cumsum_dev(p, 0, carry)
            ser = _cumsum_dev(p, 2, carry)
            assert np.array_equal(ser.view(np.uint64), ref.view(np.uint64)), (name, carry, "serial")
            assert np.array_equal(par.view(np.uint64), ref.view(np.uint64)), (name, carry, "parallel")
