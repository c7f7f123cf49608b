"""GPU parity: measurement kernels against the oracle's restatement of measure.py / probs.py /
PauliSentence.dot.  Mirrors tests/devices/qubit/test_measure.py of the reference (:125-190
state/probs/Hamiltonian known answers, 8-wire Sum)."""
import itertools

import numpy as np
import pytest

from conftest import TOL, random_state

pytestmark = pytest.mark.gpu


def _sv(state, batched=False, dtype=np.complex128):
    from pennylane_b200 import StateVector

    n = state.ndim - (1 if batched else 0)
    sv = StateVector(n, dtype=dtype)
    sv.set_state(state.astype(dtype))
    return sv


@pytest.mark.parametrize("n", [1, 2, 4, 5, 6, 9, 13])
def test_probs_all_subsets_and_orders(n):
    from oracle.measure import flatten_state, probs_process_state

    state = random_state(n, seed=n)
    sv = _sv(state)
    flat = flatten_state(state, n)
    rng = np.random.default_rng(n)
    subsets = [list(range(n)), list(range(n))[::-1]]
    for m in range(1, n + 1):
        for _ in range(3):
            subsets.append([int(x) for x in rng.permutation(n)[:m]])
    for wires in subsets:
        ref = probs_process_state(flat, wires, n)
        got = sv.probs(wires)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) < 1e-13, wires


def test_probs_batched_and_single_precision():
    from oracle.measure import flatten_state, probs_process_state

    n, B = 7, 4
    state = random_state(n, seed=5, batch=B)
    flat = flatten_state(state, n)
    for dtype, tol in [(np.complex128, 1e-13), (np.complex64, 1e-6)]:
        sv = _sv(state, batched=True, dtype=dtype)
        for wires in ([0], [6], [2, 5], [6, 0, 3], list(range(n))):
            ref = probs_process_state(flat, wires, n)
            got = sv.probs(wires)
            assert got.shape == ref.shape
            assert np.max(np.abs(got - ref)) < tol


def _rand_pauli_sentence(q, n, nterms, rng):
    cs, os_ = [], []
    for _ in range(nterms):
        k = int(rng.integers(1, min(n, 4) + 1))
        ws = [int(x) for x in rng.permutation(n)[:k]]
        word = "".join(rng.choice(list("XYZ"), size=k))
        os_.append(q.pauli_word_op(word, ws))
        cs.append(float(rng.normal()))
    return q.LinearCombination(cs, os_)


@pytest.mark.parametrize("n", [2, 5, 10, 14])
def test_expval_pauli_sums(n):
    """measure.py:74-99 csr_dot_products (Pauli branch) -> PauliSentence.dot."""
    from types import SimpleNamespace

    from oracle.measure import csr_dot_products

    from pennylane_b200 import ops as q
    from pennylane_b200.measurements import expval
    from pennylane_b200.simulate import measure

    rng = np.random.default_rng(n)
    state = random_state(n, seed=n)
    sv = _sv(state)
    for nterms in (1, 3, 17):
        H = _rand_pauli_sentence(q, n, nterms, rng)
        ref = csr_dot_products(SimpleNamespace(kind="expval", obs=H, wires=H.wires), state)
        got = measure(expval(H), sv)
        assert abs(got - ref) < 1e-12 * max(1.0, abs(ref)), (nterms, got, ref)


def test_expval_batched_state():
    from types import SimpleNamespace

    from oracle.measure import csr_dot_products

    from pennylane_b200 import ops as q
    from pennylane_b200.measurements import expval
    from pennylane_b200.simulate import measure

    n, B = 6, 3
    rng = np.random.default_rng(1)
    state = random_state(n, seed=4, batch=B)
    sv = _sv(state, batched=True)
    H = _rand_pauli_sentence(q, n, 9, rng)
    ref = csr_dot_products(SimpleNamespace(kind="expval", obs=H, wires=H.wires), state, True)
    got = measure(expval(H), sv, True)
    assert got.shape == (B,)
    assert np.max(np.abs(got - ref)) < 1e-12


@pytest.mark.parametrize("n", [3, 8])
def test_single_observables_expval_var(n):
    """measure.py:52-71 state_diagonalizing_gates for X/Y/Z/H/Hermitian/Prod, expval and var."""
    from types import SimpleNamespace

    from oracle.measure import measure as o_measure

    from pennylane_b200 import ops as q
    from pennylane_b200.measurements import expval, var
    from pennylane_b200.simulate import measure

    rng = np.random.default_rng(n)
    state = random_state(n, seed=30 + n)
    sv = _sv(state)
    A = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    A = A + A.conj().T
    obs_list = []
    for w in range(n):
        obs_list += [q.PauliX(wires=w), q.PauliY(wires=w), q.PauliZ(wires=w), q.Hadamard(wires=w)]
    obs_list += [q.PauliX(wires=0) @ q.PauliY(wires=n - 1), q.PauliZ(wires=1) @ q.Hadamard(wires=2),
                 q.Hermitian(A, wires=[n - 1, 0]), q.Hermitian(A, wires=[1, 2]),
                 2.5 * q.PauliY(wires=1), q.Projector(np.array([1, 0]), wires=[0, 2])]
    for obs in obs_list:
        for kind, fn in (("expval", expval), ("var", var)):
            ref = o_measure(SimpleNamespace(kind=kind, obs=obs, wires=obs.wires), state)
            got = measure(fn(obs), sv)
            assert abs(got - ref) < 1e-12, (obs, kind, got, ref)


def test_probs_of_observable_and_state_measurement():
    from types import SimpleNamespace

    from oracle.measure import measure as o_measure

    from pennylane_b200 import ops as q
    from pennylane_b200.measurements import probs, state as state_mp
    from pennylane_b200.simulate import measure

    n = 5
    st = random_state(n, seed=9)
    sv = _sv(st)
    obs = q.PauliX(wires=1) @ q.PauliY(wires=3)
    ref = o_measure(SimpleNamespace(kind="probs", obs=obs, wires=obs.wires), st)
    got = measure(probs(op=obs), sv)
    assert np.max(np.abs(got - ref)) < 1e-13
    got_state = measure(state_mp(), sv)
    assert np.max(np.abs(got_state - st.reshape(-1))) == 0.0


def test_inner_and_norm():
    n = 11
    a, b = random_state(n, seed=1), random_state(n, seed=2)
    sa, sb = _sv(a), _sv(b)
    assert abs(sa.inner(sb) - np.vdot(a, b)) < 1e-13
    assert abs(sa.norm2() - 1.0) < 1e-13
    # deterministic reductions: bit-identical on repetition
    assert sa.inner(sb) == sa.inner(sb)


def test_heisenberg_and_maxcut_style_hamiltonians():
    """Shapes of BASELINE configs 2 and 4: a Z-type cost Hamiltonian (single mask group) and
    XX+YY+ZZ chains (one group per edge + one diagonal group)."""
    from types import SimpleNamespace

    from oracle.measure import csr_dot_products

    from pennylane_b200 import ops as q
    from pennylane_b200.measurements import expval
    from pennylane_b200.simulate import measure

    n = 12
    state = random_state(n, seed=77)
    sv = _sv(state)
    edges = [(i, (i + 1) % n) for i in range(n)] + [(0, 5), (2, 9), (3, 7)]
    cost = q.LinearCombination(
        [0.5] * len(edges) + [-0.5] * len(edges),
        [q.PauliZ(wires=a) @ q.PauliZ(wires=b) for a, b in edges] + [q.Identity(wires=a) for a, _ in edges])
    heis = q.LinearCombination(
        [1.0] * (3 * (n - 1)),
        [P(wires=i) @ P(wires=i + 1) for i in range(n - 1) for P in (q.PauliX, q.PauliY, q.PauliZ)])
    for H in (cost, heis):
        ref = csr_dot_products(SimpleNamespace(kind="expval", obs=H, wires=H.wires), state)
        got = measure(expval(H), sv)
        assert abs(got - ref) < 1e-12 * max(1, abs(ref))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_sparse_hamiltonian_expval(dtype):
    """expval(SparseHamiltonian) (measure.py:74-118, scipy CSR branch): the CSR stays 2^k x 2^k on
    the device (b200q_expval_csr), observable on a permuted subset of wires, on all wires, and on
    a broadcast state, against the oracle's restatement."""
    import scipy.sparse as sp

    import pennylane_b200 as qb
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q

    n = 12
    rng = np.random.default_rng(5)
    prep = [q.RY(rng.uniform(0, 6), wires=i) for i in range(n)] + \
           [q.CNOT(wires=[i, (i + 1) % n]) for i in range(n)] + [q.RX(rng.uniform(0, 6), wires=i) for i in range(n)]
    tol = TOL[np.dtype(dtype)] * 30
    for wires in ([7, 2, 9], [0], list(range(n))):
        k = len(wires)
        A = sp.random(1 << k, 1 << k, density=min(0.5, 8.0 / (1 << k)), random_state=k) \
            + 1j * sp.random(1 << k, 1 << k, density=min(0.5, 8.0 / (1 << k)), random_state=k + 1)
        H = (A + A.conj().T + sp.identity(1 << k)).tocsr()
        obs = q.SparseHamiltonian(H, wires=wires)
        tape = qb.QuantumScript(prep, [qb.expval(obs)])
        got = qb.B200Qubit(wires=n, c_dtype=dtype, fusion=1).execute(tape)
        ref = o_sim.simulate(tape)
        assert abs(got - ref) < tol * max(1.0, abs(ref)), (wires, got, ref)
    B = 3
    bprep = prep + [q.RZ(rng.uniform(0, 6, B), wires=3), q.RY(rng.uniform(0, 6, B), wires=8)]
    tape = qb.QuantumScript(bprep, [qb.expval(q.SparseHamiltonian(H, wires=list(range(n))))])
    got = qb.B200Qubit(wires=n, c_dtype=dtype).execute(tape)
    ref = o_sim.simulate(tape)
    assert got.shape == (B,) and np.max(np.abs(got - ref)) < tol * max(1.0, np.max(np.abs(ref)))
