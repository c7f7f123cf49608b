"""ORACLE (test infrastructure only — never imported by ``pennylane_b200``).

CPU restatement of the tree-traversal simulation of dynamic circuits,
pennylane/devices/qubit/simulate.py: ``simulate_tree_mcm`` :396-611, ``split_circuit_at_mcms``
:623-667, ``insert_mcms`` :688-702, ``get_measurement_dicts`` :705-724, ``branch_state``
:727-753, ``samples_to_counts`` / ``counts_to_probs`` :756-770, ``prune_mcm_samples`` :773-789,
``update_mcm_samples`` :792-812, ``variance_transform`` :815-855, ``combine_measurements``
:865-937, ``TreeTraversalStack`` :70-101, ``_find_post_processed_mcms`` :52-67.

The control flow (the explicit stack, the order in which edges are simulated and therefore the
order in which the Generator is consumed) follows the reference statement by statement; states
are plain numpy arrays and the per-segment simulation uses the oracle's own ``get_final_state``
/ ``measure_final_state``.

Parity note: the reference's tests for this path (tests/devices/default_qubit/
test_default_qubit_native_mcm.py) compare tree-traversal against ``defer_measurements`` and
against analytic values; ``tests/test_tree_mcm.py`` pins this file the same way (exact branch
enumeration in analytic mode, deferred-measurement statistics with shots).
"""
from collections import Counter

import numpy as np

from .apply_operation import apply_operation
from .simulate import get_final_state, measure_final_state


class _Model:
    """The data-model classes the algorithm has to CONSTRUCT (tapes of circuit segments, the
    ``sample`` / ``probs`` / ``expval`` measurements on them, ``StatePrep``, ``PauliX``,
    ``Shots``).  The oracle does not import the product: the caller (a test) passes the package
    that provides them — ``pennylane_b200`` here, ``pennylane`` where it can be imported."""

    def __init__(self, pkg):
        self.QuantumScript = pkg.QuantumScript
        self.sample, self.probs, self.expval = pkg.sample, pkg.probs, pkg.expval
        ops = getattr(pkg, "ops", pkg)
        self.StatePrep, self.PauliX = ops.StatePrep, ops.PauliX
        self.Shots = pkg.Shots


def _branches(mv):                                           # measurements/mid_measure.py MeasurementValue.branches
    import itertools
    out = {}
    ms = list(mv.measurements)
    for bits in itertools.product([0, 1], repeat=len(ms)):
        out[bits] = mv.processing_fn(*bits)
    return out


def _gather_non_mcm(m, samples, is_valid):                   # dynamic_one_shot.py:401-507
    is_valid = np.asarray(is_valid)
    if m.kind == "counts":
        tmp = Counter()
        for i, d in enumerate(samples):
            tmp.update({k if isinstance(k, str) else float(k): v * bool(np.ravel(is_valid)[i])
                        for k, v in d.items()})
        tmp = Counter({k: v for k, v in tmp.items() if v > 0})
        return dict(sorted(tmp.items()))
    if m.kind == "sample":
        arr = np.concatenate(samples) if isinstance(samples, (list, tuple)) else np.asarray(samples)
        return arr[np.ravel(is_valid)] if arr.ndim == 1 else arr[np.ravel(is_valid)]
    arr = np.squeeze(np.stack([np.asarray(x) for x in samples]))
    v = np.ravel(is_valid)
    if m.kind == "expval":
        return np.sum(arr * v) / np.sum(v)
    if m.kind == "var":
        e = np.sum(arr * v) / np.sum(v)
        return np.sum((arr - e) ** 2 * v) / np.sum(v)
    raise TypeError(m.kind)


def _gather_mcm(m, samples, is_valid):                       # dynamic_one_shot.py:512-575 (single value)
    vals = np.asarray(m.mv.concretize(samples))
    if m.kind == "probs":
        vals = np.squeeze(vals)
        cnt = np.array([np.count_nonzero(np.logical_and(vals == v, np.squeeze(is_valid)))
                        for v in _branches(m.mv).values()])
        return cnt / np.sum(cnt)
    data = vals
    if m.kind == "counts":
        data = [{float(np.asarray(s).item()): 1} for s in vals]
    res = _gather_non_mcm(m, data, is_valid)
    return np.squeeze(res) if m.kind == "sample" else res


def _is_mcm(op):
    return op.name == "MidMeasureMP"


def _find_post_processed_mcms(circuit):                      # :52-67
    post = {op for op in circuit.operations if _is_mcm(op) and op.postselect is not None}
    for m in circuit.measurements:
        mv = getattr(m, "mv", None)
        if isinstance(mv, list):
            for v in mv:
                post |= set(v.measurements)
        elif mv is not None:
            post |= set(mv.measurements)
    return post


class TreeTraversalStack:                                    # :70-101
    def __init__(self, max_depth):
        self.counts = [None] * max_depth
        self.probs = [None] * max_depth
        self.results_0 = [None] * max_depth
        self.results_1 = [None] * max_depth
        self.states = [None] * max_depth

    def any_is_empty(self, depth):
        return self.results_0[depth] is None or self.results_1[depth] is None

    def is_full(self, depth):
        return self.results_0[depth] is not None and self.results_1[depth] is not None

    def prune(self, depth):
        self.counts[depth] = self.probs[depth] = None
        self.results_0[depth] = self.results_1[depth] = self.states[depth] = None


def split_circuit_at_mcms(circuit, qb):                      # :623-667
    circuits, first = [], 0
    ops_ = list(circuit.operations)
    for last, op in enumerate(ops_):
        if not _is_mcm(op):
            continue
        meas = [qb.sample(wires=op.wires)] if circuit.shots else [qb.probs(wires=op.wires)]
        circuits.append(qb.QuantumScript(ops_[first:last], meas, shots=circuit.shots))
        first = last + 1
    final = [m for m in circuit.measurements if getattr(m, "mv", None) is None]
    circuits.append(qb.QuantumScript(ops_[first:], final, shots=circuit.shots))
    return circuits


def branch_state(state, branch, mcm, qb):                    # :727-753
    state = state.copy()
    slices = [slice(None)] * state.ndim
    slices[int(mcm.wires[0])] = int(not branch)
    state[tuple(slices)] = 0.0
    state /= np.linalg.norm(state)
    if mcm.reset and branch == 1:
        state = apply_operation(qb.PauliX(wires=mcm.wires), state)
    return state


def samples_to_counts(samples):                              # :756-763
    c1 = int(np.count_nonzero(samples))
    return {0: samples.size - c1, 1: c1}


def counts_to_probs(counts):                                 # :766-770
    p = np.array(list(counts.values()))
    p = p / np.sum(p)
    return dict(zip(counts.keys(), p))


def prune_mcm_samples(mcm_samples):                          # :773-789
    if not mcm_samples or all(v is None for v in mcm_samples.values()):
        return mcm_samples
    mask = np.ones(list(mcm_samples.values())[0].shape, dtype=bool)
    for mcm, s in mcm_samples.items():
        if mcm.postselect is None:
            continue
        mask = np.logical_and(mask, s == mcm.postselect)
    return {k: v[mask] for k, v in mcm_samples.items()}


def update_mcm_samples(samples, mcm_samples, depth, cumcounts):   # :792-812
    if depth not in mcm_samples or mcm_samples[depth] is None:
        return mcm_samples, cumcounts
    count1 = int(np.sum(samples))
    count0 = samples.size - count1
    mcm_samples[depth][cumcounts[depth]: cumcounts[depth] + count0] = 0
    cumcounts[depth] += count0
    mcm_samples[depth][cumcounts[depth]: cumcounts[depth] + count1] = 1
    cumcounts[depth] += count1
    return mcm_samples, cumcounts


def variance_transform(circuit, qb):                         # :815-855
    if not any(m.kind == "var" for m in circuit.measurements):
        return circuit, lambda x: x
    new, extra = [], []
    for m in circuit.measurements:
        if m.kind == "var":
            mv = getattr(m, "mv", None)
            obs2 = mv * mv if mv is not None else m.obs @ m.obs
            new.append(qb.expval(obs2))
            extra.append(qb.expval(mv if mv is not None else m.obs))
        else:
            new.append(m)
    orig = list(circuit.measurements)

    def post(results):
        res = list(results)
        offset = len(orig)
        for i, m in enumerate(orig):
            if m.kind == "var":
                e = res.pop(offset)
                res[i] = res[i] - e ** 2
        return res[0] if len(res) == 1 else tuple(res)

    return qb.QuantumScript(circuit.operations, new + extra, shots=circuit.shots), post


def measurement_with_no_shots(m):                            # :858-862
    if m.kind == "probs":
        return np.nan * np.ones(2 ** len(m.wires))
    return np.nan


def insert_mcms(circuit, results, mid_measurements):         # :688-702
    if circuit.shots or all(getattr(m, "mv", None) is None for m in circuit.measurements):
        return results
    results = list(results) if isinstance(results, (list, tuple)) else [results]
    new = []
    mid = {k: np.array([[v]]) for k, v in mid_measurements.items()}
    for m in circuit.measurements:
        if getattr(m, "mv", None) is None:
            new.append(results.pop(0))
        else:
            new.append(_gather_mcm(m, mid, np.array([[True]])))
    return new


def get_measurement_dicts(measurements, stack, depth):       # :705-724
    probs, r0, r1 = stack.probs[depth], stack.results_0[depth], stack.results_1[depth]
    dicts = [{} for _ in measurements]
    single = len(measurements) == 1
    for branch, prob in probs.items():
        meas = r0 if branch == 0 else r1
        if single:
            meas = [meas]
        for i, m in enumerate(meas):
            dicts[i][branch] = (prob, m)
    return dicts


def _is_empty(x):
    return isinstance(x, tuple) and len(x) == 0


def combine_measurements_core(m, measures):                  # :940-1000
    if m.kind == "counts":
        c = Counter()
        for k in list(measures.keys()):
            if not measures[k][0]:
                continue
            c.update(measures[k][1])
        return dict(sorted(c.items()))
    if m.kind in ("expval", "probs"):
        cum, tot = 0, 0
        for v in measures.values():
            if not v[0] or _is_empty(v[1]):
                continue
            cum = cum + np.multiply(v[0], v[1])
            tot = tot + v[0]
        return cum / tot
    if m.kind == "sample":
        parts = tuple(np.atleast_1d(v[1]) for v in measures.values() if v[0] and not _is_empty(v[1]))
        return np.concatenate(parts)
    raise TypeError(f"Native mid-circuit measurement mode does not support {m.kind}")


def combine_measurements(terminal_measurements, results, mcm_samples):   # :865-937
    empty = False
    need = not all(v is None for v in mcm_samples.values())
    need = need and any(getattr(m, "mv", None) is not None for m in terminal_measurements)
    if need:
        empty = len(next(iter(mcm_samples.values()))) == 0
    final = []
    for m in terminal_measurements:
        has_mv = getattr(m, "mv", None) is not None
        if need and has_mv and empty:
            comb = measurement_with_no_shots(m)
        elif need and has_mv:
            mcm_samples = {k: v.reshape((-1, 1)) for k, v in mcm_samples.items()}
            is_valid = np.ones(list(mcm_samples.values())[0].shape[0], dtype=bool)
            comb = _gather_mcm(m, mcm_samples, is_valid)
        elif not results or not results[0]:
            if len(results) > 0:
                results.pop(0)
            comb = measurement_with_no_shots(m)
        else:
            comb = combine_measurements_core(m, results.pop(0))
        final.append(comb)
    return final[0] if len(final) == 1 else tuple(final)


def simulate_tree_mcm(circuit, rng=None, model=None):        # :396-611
    """``model``: the package providing the data-model classes (see :class:`_Model`)."""
    PROBS_TOL = 0.0
    qb = model if isinstance(model, _Model) else _Model(model)
    if circuit.shots and circuit.shots.has_partitioned_shots:        # :430-437
        return tuple(simulate_tree_mcm(circuit.copy(shots=s), rng=rng, model=qb) for s in circuit.shots)
    circuit, variance_post = variance_transform(circuit, qb)
    finite = bool(circuit.shots)
    n = circuit.num_wires
    mcms = tuple([None] + [op for op in circuit.operations if _is_mcm(op)])
    n_mcms = len(mcms) - 1
    measured = _find_post_processed_mcms(circuit)
    measured_idx = [i for i, mcm in enumerate(mcms[1:]) if mcm in measured]
    mcm_samples = {k + 1: (np.empty((circuit.shots.total_shots,), dtype=int) if finite else None)
                   for k in measured_idx}
    mcm_current = np.zeros(n_mcms + 1, dtype=int)
    mid_measurements = dict(zip(mcms[1:], mcm_current[1:].tolist()))
    circuits = split_circuit_at_mcms(circuit, qb)
    init = np.zeros((2,) * n, dtype=complex)
    init[(0,) * n] = 1.0
    terminal = circuits[-1].measurements if finite else circuit.measurements
    cumcounts = [0] * (n_mcms + 1)
    stack = TreeTraversalStack(n_mcms + 1)
    stack.states[0] = init
    depth = 0
    if n_mcms == 0:
        raise ValueError("simulate_tree_mcm needs at least one mid-circuit measurement")

    def run_segment(d, state, shots):
        seg = circuits[d]
        prep = qb.StatePrep(state.reshape(-1), wires=list(range(n)), validate_norm=False)
        tape = qb.QuantumScript([prep] + list(seg.operations), seg.measurements,
                                shots=qb.Shots(shots) if shots is not None else None)
        st, batched = get_final_state(tape, mid_measurements=mid_measurements, rng=rng)
        return st, measure_final_state(tape, st, batched, rng=rng)

    while stack.any_is_empty(1):
        if stack.is_full(depth):                                     # :493-512
            dicts = get_measurement_dicts(terminal, stack, depth)
            measurements = combine_measurements(terminal, dicts, mcm_samples)
            mcm_current[depth:] = 0
            stack.prune(depth)
            depth -= 1
            if mcm_current[depth] == 1:
                stack.results_1[depth] = measurements
                mcm_current[depth] = 0
            else:
                stack.results_0[depth] = measurements
                mcm_current[depth] = 1
            mid_measurements.update((k, v) for k, v in zip(mcms[depth:], mcm_current[depth:].tolist()))
            continue
        if finite:                                                   # :519-530
            if stack.counts[depth]:
                shots = stack.counts[depth][mcm_current[depth]]
            else:
                shots = circuits[depth].shots.total_shots
            skip = not bool(shots)
        else:
            shots = None
            skip = (stack.probs[depth] is not None
                    and float(stack.probs[depth][mcm_current[depth]]) <= PROBS_TOL)
        invalid = (depth > 0 and mcms[depth].postselect is not None
                   and mcm_current[depth] != mcms[depth].postselect)
        if skip or invalid:                                          # :543-553
            if invalid:
                if finite:
                    for d in range(depth + 1, n_mcms + 1):
                        cumcounts[d] += stack.counts[depth][mcm_current[depth]]
                    stack.counts[depth][mcm_current[depth]] = 0
                else:
                    stack.probs[depth][mcm_current[depth]] = 0
            measurements = tuple()
        else:                                                        # :555-569
            if depth == 0:
                initial_state = stack.states[0]
            else:
                initial_state = branch_state(stack.states[depth], mcm_current[depth], mcms[depth], qb)
            state, measurements = run_segment(depth, initial_state, shots)
        if depth < n_mcms and (not skip and not invalid):            # :575-589
            depth += 1
            if finite:
                samples = np.atleast_1d(measurements)
                stack.counts[depth] = samples_to_counts(samples)
                stack.probs[depth] = counts_to_probs(stack.counts[depth])
            else:
                stack.probs[depth] = dict(zip([False, True], measurements))
                samples = None
            stack.states[depth] = state
            mcm_samples, cumcounts = update_mcm_samples(samples, mcm_samples, depth, cumcounts)
            continue
        if not skip and not invalid:                                 # :595-596
            measurements = insert_mcms(circuit, measurements, mid_measurements)
        if mcm_current[depth] == 0:                                  # :598-605
            stack.results_0[depth] = measurements
            mcm_current[depth] = True
            mid_measurements[mcms[depth]] = True
            continue
        stack.results_1[depth] = measurements
    dicts = get_measurement_dicts(terminal, stack, depth)            # :611-618
    if finite:
        terminal = circuit.measurements
    samples_by_mcm = {mcms[i]: v for i, v in mcm_samples.items()}
    samples_by_mcm = prune_mcm_samples(samples_by_mcm)
    results = combine_measurements(terminal, dicts, samples_by_mcm)
    return variance_post(results)
