"""Tree-traversal simulation of dynamic circuits on the CUDA engine (``mcm_method =
"tree-traversal"``), pennylane/devices/qubit/simulate.py:396-611 (``simulate_tree_mcm``) with its
helpers ``split_circuit_at_mcms`` :623-667, ``branch_state`` :727-753, ``update_mcm_samples``
:792-812, ``variance_transform`` :815-855, ``combine_measurements`` :865-1000.

The reference walks the outcome tree with an explicit stack and keeps one host state per depth.
Here the walk is a depth-first recursion over device-resident states:

* an edge = the unitary stretch between two mid-circuit measurements, run through the fused
  segment path (``apply_gates``; ``Conditional`` gates are resolved from the branch's outcomes);
* a node = ONE marginal sweep of the measured wire (``b200q_probs``); with shots the wire is
  sampled with the node's shot budget exactly as the reference samples its ``sample(wires=mcm)``
  segment measurement (same Generator calls in the same depth-first order, so results are
  bit-identical to the oracle under the same seed), the two children get the counts as budgets and
  a child with no shots is never simulated;
* a child state = ``b200q_collapse`` (projector + 1/norm + optional reset in one pass: S/2 read +
  S written, where the reference copies, zeroes, normalises and flips in four passes).  The
  0-branch works on a clone, the 1-branch re-uses its parent's buffer (the parent is dead once
  its 0-subtree is done): ``n_mcm + 1`` state vectors at most, like the reference;
* terminal measurements are taken once per leaf with the leaf's whole shot budget and combined
  upwards, count- (or probability-) weighted.
"""
from __future__ import annotations

from collections import Counter

import numpy as np

from .mcm import is_mcm
from .one_shot import _gather_mcm


def _post_processed_mcms(circuit):
    """simulate.py:52-67: measurements that are post-selected or read by a terminal measurement."""
    post = {op for op in circuit.operations if is_mcm(op) and op.postselect is not None}
    for m in circuit.measurements:
        mv = getattr(m, "mv", None)
        if isinstance(mv, (list, tuple)):
            for v in mv:
                post |= set(v.measurements)
        elif mv is not None:
            post |= set(mv.measurements)
    return post


def _variance_transform(circuit):
    """simulate.py:815-855: var(O) -> (expval(O @ O), expval(O)); the global variance is formed
    after the branches are combined."""
    from .measurements import expval
    from .tape import QuantumScript

    orig = list(circuit.measurements)
    if not any(m.kind == "var" for m in orig):
        return circuit, None
    new, extra = [], []
    for m in orig:
        if m.kind == "var":
            mv = getattr(m, "mv", None)
            new.append(expval(mv * mv if mv is not None else m.obs @ m.obs))
            extra.append(expval(mv if mv is not None else m.obs))
        else:
            new.append(m)

    def post(results):
        res = list(results)
        offset = len(orig)
        for i, m in enumerate(orig):
            if m.kind == "var":
                e = res.pop(offset)
                res[i] = res[i] - e ** 2
        return res[0] if len(res) == 1 else tuple(res)

    return QuantumScript(circuit.operations, new + extra, shots=circuit.shots), post


def _no_shots(m):
    return np.nan * np.ones(2 ** len(m.wires)) if m.kind == "probs" else np.nan


def _empty(x):
    return isinstance(x, tuple) and len(x) == 0


def _combine_one(m, by_branch):
    """simulate.py:940-1000 (``combine_measurements_core``): ``by_branch`` = {branch: (weight,
    result)} with weight = the branch's probability (or count fraction)."""
    if m.kind == "counts":
        acc = Counter()
        for w, r in by_branch.values():
            if w:
                acc.update(r)
        return dict(sorted(acc.items()))
    if m.kind in ("expval", "probs"):
        cum, tot = 0, 0
        for w, r in by_branch.values():
            if not w or _empty(r):
                continue
            cum = cum + np.multiply(w, r)
            tot = tot + w
        return cum / tot
    if m.kind == "sample":
        return np.concatenate(tuple(np.atleast_1d(r) for w, r in by_branch.values() if w and not _empty(r)))
    raise TypeError(f"Native mid-circuit measurement mode does not support {m.kind} measurements.")


def _combine(terminal, per_measurement, mcm_samples):
    """simulate.py:865-937 (``combine_measurements``)."""
    need = (not all(v is None for v in mcm_samples.values())
            and any(getattr(m, "mv", None) is not None for m in terminal))
    empty = need and len(next(iter(mcm_samples.values()))) == 0
    out = []
    results = list(per_measurement)
    for m in terminal:
        has_mv = getattr(m, "mv", None) is not None
        if need and has_mv and empty:
            out.append(_no_shots(m))
        elif need and has_mv:
            cols = {k: v.reshape((-1, 1)) for k, v in mcm_samples.items()}
            valid = np.ones(next(iter(cols.values())).shape[0], dtype=bool)
            out.append(_gather_mcm(m, cols, valid))
        elif not results or not results[0]:
            if results:
                results.pop(0)
            out.append(_no_shots(m))
        else:
            out.append(_combine_one(m, results.pop(0)))
    return out[0] if len(out) == 1 else tuple(out)


def simulate_tree_mcm(circuit, rng=None, dtype=np.complex128, device=None, exact_sampling: bool = True,
                      fusion: int = 0, debugger=None):
    """simulate.py:396-611."""
    from .measurements import probs as probs_mp
    from .measurements import sample as sample_mp
    from .simulate import apply_gates, measure_final_state
    from .statevector import StateVector
    from .tape import QuantumScript, Shots

    if circuit.shots and circuit.shots.has_partitioned_shots:               # :430-437
        return tuple(simulate_tree_mcm(circuit.copy(shots=Shots(s)), rng, dtype, device,
                                       exact_sampling, fusion, debugger) for s in circuit.shots)
    rng = np.random.default_rng(rng)
    circuit, var_post = _variance_transform(circuit)
    finite = bool(circuit.shots)
    n = circuit.num_wires
    ops_ = list(circuit.operations)
    mcm_pos = [i for i, op in enumerate(ops_) if is_mcm(op)]
    mcms = [ops_[i] for i in mcm_pos]
    n_mcms = len(mcms)
    # circuit segments between the measurements (split_circuit_at_mcms :623-667)
    bounds = [-1] + mcm_pos + [len(ops_)]
    segments = [ops_[bounds[d] + 1: bounds[d + 1]] for d in range(n_mcms + 1)]
    leaf_meas = [m for m in circuit.measurements if getattr(m, "mv", None) is None]
    terminal = leaf_meas if finite else list(circuit.measurements)
    measured = _post_processed_mcms(circuit)
    total = circuit.shots.total_shots if finite else 0
    # register of correlated mid-circuit samples, one row per measured MCM (depth = index + 1)
    mcm_samples = {d + 1: (np.empty((total,), dtype=int) if finite else None)
                   for d, mcm in enumerate(mcms) if mcm in measured}
    cumcounts = [0] * (n_mcms + 1)

    def record(depth, samples):                                             # update_mcm_samples :792-812
        if depth not in mcm_samples or mcm_samples[depth] is None:
            return
        c1 = int(np.sum(samples))
        c0 = samples.size - c1
        reg, at = mcm_samples[depth], cumcounts[depth]
        reg[at: at + c0] = 0
        reg[at + c0: at + c0 + c1] = 1
        cumcounts[depth] = at + c0 + c1

    def leaf(sv, shots, mid):
        if not terminal:
            return tuple()
        tape = QuantumScript([], leaf_meas, shots=Shots(shots) if finite else None)
        res = measure_final_state(tape, sv, False, rng=rng, exact_sampling=exact_sampling) if leaf_meas else ()
        if finite or all(getattr(m, "mv", None) is None for m in circuit.measurements):
            return res
        # analytic mode: terminal measurements of MCM values are the branch's own values (:688-702)
        vals = list(res) if isinstance(res, tuple) else ([res] if leaf_meas else [])
        mid1 = {k: np.array([[v]]) for k, v in mid.items()}
        out = []
        for m in circuit.measurements:
            if getattr(m, "mv", None) is None:
                out.append(vals.pop(0))
            else:
                out.append(_gather_mcm(m, mid1, np.array([[True]])))
        return out[0] if len(out) == 1 else tuple(out)

    def visit(depth, sv, shots, mid):
        """Results of the subtree whose edge ``depth`` starts from ``sv`` (consumed)."""
        apply_gates(sv, segments[depth], fusion, mid, rng, debugger, circuit.shots, exact_sampling)
        if depth == n_mcms:
            return leaf(sv, shots, mid)
        mcm = mcms[depth]
        wire = mcm.wires[0]
        p_true = np.asarray(sv.probs([wire]), dtype=float).reshape(-1)      # one marginal sweep
        if finite:
            tape = QuantumScript([], [sample_mp(wires=[wire])], shots=Shots(shots))
            samples = np.atleast_1d(measure_final_state(tape, sv, False, rng=rng,
                                                        exact_sampling=exact_sampling))
            c1 = int(np.count_nonzero(samples))
            counts = {0: samples.size - c1, 1: c1}
            tot = counts[0] + counts[1]
            weights = {0: counts[0] / tot, 1: counts[1] / tot}
            record(depth + 1, samples)
        else:
            counts = None
            weights = {0: p_true[0], 1: p_true[1]}
        results = {}
        for branch in (0, 1):
            skip = (counts[branch] == 0) if finite else (float(weights[branch]) <= 0.0)
            invalid = mcm.postselect is not None and branch != mcm.postselect
            if skip or invalid:
                if invalid:
                    if finite:
                        for d in range(depth + 2, n_mcms + 1):
                            cumcounts[d] += counts[branch]
                        counts[branch] = 0
                    else:
                        weights[branch] = 0
                results[branch] = tuple()
                continue
            # the 0-branch works on a copy; the 1-branch re-uses the parent's buffer
            child = sv.clone() if (branch == 0) else sv
            child.collapse(wire, branch, bool(mcm.reset), 1.0 / np.sqrt(p_true[branch]))
            mid[mcm] = branch
            results[branch] = visit(depth + 1, child, counts[branch] if finite else None, mid)
            if branch == 0:
                del child
        mid.pop(mcm, None)
        per_meas = [{} for _ in terminal]
        for branch in (0, 1):
            r = results[branch]
            r = [r] if len(terminal) == 1 else r
            for i, x in enumerate(r):
                per_meas[i][branch] = (weights[branch], x)
        return _combine(terminal, per_meas, mcm_samples)

    if n_mcms == 0:
        raise ValueError("tree-traversal needs at least one mid-circuit measurement")
    prep = ops_[0] if ops_ and hasattr(ops_[0], "state_vector") else None
    root = StateVector(n, dtype=dtype, device=device)
    if prep is not None:
        root.set_state(np.asarray(prep.state_vector(wire_order=list(range(n)))))
        segments[0] = segments[0][1:]
    # inner combinations never read the register (their terminal list has no MCM-valued
    # measurement with shots; in analytic mode the register is all None): only the root does
    res_root = _visit_root(visit, root, total if finite else None, {}, circuit, mcms, mcm_samples,
                           finite, terminal)
    return var_post(res_root if isinstance(res_root, tuple) else (res_root,)) if var_post else res_root


def _visit_root(visit, root, shots, mid, circuit, mcms, mcm_samples, finite, terminal):
    """The top of the tree (simulate.py:598-611): run the walk, then combine with the register of
    mid-circuit samples pruned by post-selection (``prune_mcm_samples`` :773-789) so that
    MCM-valued terminal measurements are gathered over the valid shots."""
    res = visit(0, root, shots, mid)
    if not finite:
        return res
    if not any(getattr(m, "mv", None) is not None for m in circuit.measurements):
        return res
    # re-assemble: non-MCM results come from the walk (in order), MCM-valued ones from the register
    by_mcm = {mcms[d - 1]: v for d, v in mcm_samples.items()}
    if by_mcm and not all(v is None for v in by_mcm.values()):
        mask = np.ones(next(iter(by_mcm.values())).shape, dtype=bool)
        for mcm, s in by_mcm.items():
            if mcm.postselect is not None:
                mask = np.logical_and(mask, s == mcm.postselect)
        by_mcm = {k: v[mask] for k, v in by_mcm.items()}
    walk = list(res) if isinstance(res, tuple) else ([res] if terminal else [])
    out = []
    empty = len(next(iter(by_mcm.values()))) == 0 if by_mcm else True
    for m in circuit.measurements:
        if getattr(m, "mv", None) is None:
            out.append(walk.pop(0))
        elif empty:
            out.append(_no_shots(m))
        else:
            cols = {k: v.reshape((-1, 1)) for k, v in by_mcm.items()}
            valid = np.ones(next(iter(cols.values())).shape[0], dtype=bool)
            out.append(_gather_mcm(m, cols, valid))
    return out[0] if len(out) == 1 else tuple(out)
