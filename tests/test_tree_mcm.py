"""Tree-traversal mid-circuit measurements (simulate.py:396-611).

CPU: the oracle restatement (oracle/tree_mcm.py) pinned the way the reference's own tests pin the
method (tests/devices/default_qubit/test_default_qubit_native_mcm.py): analytic mode against an
exact enumeration of the outcome branches, finite shots against those exact values statistically,
reset / postselection / conditionals included.  GPU (``-m gpu``): the CUDA implementation
(pennylane_b200/tree_mcm.py) against the oracle — exact in analytic mode, bit-identical samples
and counts under the same seed with shots."""
import itertools

import numpy as np
import pytest

import pennylane_b200 as qb
from pennylane_b200 import ops as q
from pennylane_b200.mcm import cond, measure


def _circuit(kind, shots=None, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.3, 2.8, size=12)
    ops_ = [q.RY(a[0], wires=0), q.RX(a[1], wires=1), q.CNOT(wires=[0, 2]), q.RY(a[2], wires=2)]
    # post-selection sits on the FIRST measurement: the reference weights the two children of a
    # node by the node's own outcome probabilities and does not re-weight ancestors by the
    # probability that a LATER post-selection succeeds (simulate.py:543-553 zeroes only the
    # counts / probs of the post-selected node), so nested post-selection agrees with the exact
    # conditional distribution only at the root; "postselect2" keeps one on the second
    # measurement for the CUDA-vs-oracle parity tests, where the reference's behaviour is the bar
    m0 = measure(0, reset=(kind == "reset"), postselect=(1 if kind == "postselect" else None))
    ops_ += list(m0.measurements)
    ops_ += [cond(m0 == 1, q.RY(a[3], wires=1)), q.CNOT(wires=[1, 2]), q.RX(a[4], wires=0)]
    m1 = measure(1, postselect=(1 if kind == "postselect2" else None))
    ops_ += list(m1.measurements)
    ops_ += [cond(m0 + m1 == 1, q.RZ(a[5], wires=2)), q.RY(a[6], wires=2), q.CNOT(wires=[2, 0])]
    m2 = measure(2)
    ops_ += list(m2.measurements)
    ops_ += [cond(m2 == 1, q.PauliX(wires=1)), q.RY(a[7], wires=0), q.RX(a[8], wires=1)]
    if kind == "mv":
        meas = [qb.expval(q.PauliZ(wires=0)), qb.probs(wires=[1, 2]), qb.expval(m1), qb.probs(op=m0)]
    elif shots:
        meas = [qb.expval(q.PauliZ(wires=0) @ q.PauliX(wires=1)), qb.probs(wires=[1, 2]),
                qb.var(q.PauliY(wires=2)), qb.counts(wires=[0, 1]), qb.sample(wires=[2])]
    else:
        meas = [qb.expval(q.PauliZ(wires=0) @ q.PauliX(wires=1)), qb.probs(wires=[1, 2]),
                qb.var(q.PauliY(wires=2))]
    return qb.QuantumScript(ops_, meas, shots=shots), (m0, m1, m2)


def _exact(tape):
    """Exact values of the NON-MCM measurements by enumerating all outcome branches with the
    oracle's gate kernels (projector, renormalise, reset; postselected branches dropped and the
    rest renormalised)."""
    from oracle.apply_operation import apply_operation
    from oracle.measure import measure as o_measure

    n = tape.num_wires
    mcms = [op for op in tape.operations if op.name == "MidMeasureMP"]
    acc = None
    total = 0.0
    for outcome in itertools.product([0, 1], repeat=len(mcms)):
        state = np.zeros((2,) * n, dtype=complex)
        state[(0,) * n] = 1
        mid, p = {}, 1.0
        dead = False
        for op in tape.operations:
            if op.name == "MidMeasureMP":
                b = outcome[len(mid)]
                w = int(op.wires[0])
                sl = [slice(None)] * n
                sl[w] = 1 - b
                state = state.copy()
                state[tuple(sl)] = 0
                nb = np.linalg.norm(state) ** 2
                if nb < 1e-300 or (op.postselect is not None and b != op.postselect):
                    dead = True
                    break
                p *= nb
                state = state / np.sqrt(nb)
                if op.reset and b == 1:
                    state = apply_operation(q.PauliX(wires=[w]), state)
                mid[op] = b
            elif op.name.startswith("Conditional") or hasattr(op, "meas_val"):
                if op.meas_val.concretize(mid):
                    state = apply_operation(op.base, state)
            else:
                state = apply_operation(op, state)
        if dead:
            continue
        vals = []
        for m in tape.measurements:
            if getattr(m, "mv", None) is not None or m.kind in ("counts", "sample"):
                continue
            if m.kind == "var":
                e2 = o_measure(qb.expval(m.obs @ m.obs), state)
                e1 = o_measure(qb.expval(m.obs), state)
                vals.append(np.array([e2, e1]))
            else:
                vals.append(np.asarray(o_measure(m, state)))
        acc = [p * v for v in vals] if acc is None else [x + p * v for x, v in zip(acc, vals)]
        total += p
    out = []
    i = 0
    for m in tape.measurements:
        if getattr(m, "mv", None) is not None or m.kind in ("counts", "sample"):
            continue
        v = acc[i] / total
        out.append(v[0] - v[1] ** 2 if m.kind == "var" else v)
        i += 1
    return out


@pytest.mark.parametrize("kind", ["plain", "reset", "postselect"])
def test_oracle_tree_analytic_matches_branch_enumeration(kind):
    from oracle.tree_mcm import simulate_tree_mcm

    tape, _ = _circuit(kind)
    res = simulate_tree_mcm(tape, rng=np.random.default_rng(1), model=qb)
    ref = _exact(tape)
    assert len(res) == len(ref)
    for a, b in zip(res, ref):
        assert np.allclose(a, b, atol=1e-12)


@pytest.mark.parametrize("kind", ["plain", "reset", "postselect"])
def test_oracle_tree_with_shots_matches_exact_values_statistically(kind):
    from oracle.tree_mcm import simulate_tree_mcm

    shots = 40000
    tape, _ = _circuit(kind, shots=shots)
    res = simulate_tree_mcm(tape, rng=np.random.default_rng(2), model=qb)
    ref = _exact(tape)
    assert abs(res[0] - ref[0]) < 0.03
    assert np.max(np.abs(res[1] - ref[1])) < 0.02
    assert abs(res[2] - ref[2]) < 0.03
    assert isinstance(res[3], dict) and res[4].ndim >= 1
    kept = sum(res[3].values())
    assert kept == len(np.atleast_1d(res[4]))
    if kind != "postselect":
        assert kept == shots


def test_oracle_tree_mcm_valued_measurements():
    from oracle.tree_mcm import simulate_tree_mcm

    tape_a, _ = _circuit("mv")
    res_a = simulate_tree_mcm(tape_a, rng=np.random.default_rng(3), model=qb)
    tape_s, _ = _circuit("mv", shots=50000)
    res_s = simulate_tree_mcm(tape_s, rng=np.random.default_rng(3), model=qb)
    for a, s in zip(res_a, res_s):
        assert np.max(np.abs(np.asarray(a, dtype=float) - np.asarray(s, dtype=float))) < 0.02
    assert abs(np.sum(res_a[3]) - 1) < 1e-12


# ---- CUDA implementation ------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["plain", "reset", "postselect", "postselect2", "mv"])
@pytest.mark.parametrize("fusion", [0, 1])
def test_gpu_tree_analytic_matches_oracle(kind, fusion):
    from oracle.tree_mcm import simulate_tree_mcm as o_tree
    from pennylane_b200.tree_mcm import simulate_tree_mcm

    tape, _ = _circuit(kind)
    ref = o_tree(tape, rng=np.random.default_rng(1), model=qb)
    res = simulate_tree_mcm(tape, rng=np.random.default_rng(1), fusion=fusion)
    assert len(res) == len(ref)
    for a, b in zip(res, ref):
        assert np.allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["plain", "reset", "postselect", "postselect2", "mv"])
def test_gpu_tree_with_shots_is_bit_identical_to_oracle(kind):
    """Same Generator, same depth-first order: counts, samples, estimates are identical."""
    from oracle.tree_mcm import simulate_tree_mcm as o_tree
    from pennylane_b200.tree_mcm import simulate_tree_mcm

    tape, _ = _circuit(kind, shots=3000)
    ref = o_tree(tape, rng=np.random.default_rng(7), model=qb)
    res = simulate_tree_mcm(tape, rng=np.random.default_rng(7))
    assert len(res) == len(ref)
    for a, b in zip(res, ref):
        if isinstance(b, dict):
            assert a == b
        else:
            a, b = np.asarray(a), np.asarray(b)
            assert a.shape == b.shape
            if np.issubdtype(b.dtype, np.integer):
                assert np.array_equal(a, b)
            else:
                assert np.allclose(a, b, atol=1e-12, equal_nan=True)


@pytest.mark.gpu
def test_gpu_device_accepts_tree_traversal_and_shot_vectors():
    from oracle.tree_mcm import simulate_tree_mcm as o_tree

    tape, _ = _circuit("plain", shots=[500, 700])
    dev = qb.B200Qubit(wires=3, seed=11)
    cfg = qb.ExecutionConfig(mcm_config=qb.MCMConfig(mcm_method="tree-traversal"))
    batch, cfg2 = dev.preprocess(tape, cfg)
    assert cfg2.mcm_config.mcm_method == "tree-traversal"
    res = dev.execute(batch, cfg2)[0]
    ref = o_tree(tape.map_to_standard_wires(), rng=np.random.default_rng(11), model=qb)
    assert len(res) == 2
    for ra, rb in zip(res, ref):
        for a, b in zip(ra, rb):
            if isinstance(b, dict):
                assert a == b
            else:
                assert np.allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), atol=1e-12)
