"""Circuit execution on the CUDA engine — the device-side mirror of
pennylane/devices/qubit/simulate.py (``get_final_state`` :174-242, ``measure_final_state``
:246-304, ``simulate`` :308-393), measure.py (``measure`` :224-239 and its strategy choice
:165-221) and sampling.py (``measure_with_samples`` :205-335, ``_group_measurements`` :46-98,
``sample_state`` :439-476).

Every amplitude-sized operation is a kernel call on a :class:`StateVector`; the host only
handles scalars, 2^m-sized marginals (m = observable wires) and the shot bookkeeping.
"""
from __future__ import annotations

import numpy as np

from .mcm import is_conditional, is_mcm
from .pauli import PauliSentence, PauliWord
from .statevector import StateVector


def _is_prep(op) -> bool:
    return hasattr(op, "state_vector")


def get_final_state(circuit, dtype=np.complex128, device=None, buffer=None, fusion: int = 0,
                    mid_measurements=None, rng=None, debugger=None, exact_sampling: bool = True):
    """Run the gate loop (simulate.py:214-235).  ``circuit`` must be in standard wire order.
    ``mid_measurements`` (a dict, filled in place) and ``rng`` are needed only when the tape
    holds ``MidMeasure`` operations.

    Returns ``(StateVector, is_state_batched)``.  Measurement-only wires are simply extra
    low-order qubits left in |0> (the reference pads them afterwards, simulate.py:237-240).
    """
    ops_ = list(circuit.operations)
    n = circuit.num_wires
    prep = ops_[0] if ops_ and _is_prep(ops_[0]) else None
    # A broadcast tape gets its (B, 2^n) state up front: growing the state at the first batched
    # gate (simulate.py:235) would hold the old and the new buffer at once — at 32 qubits,
    # B = 2, complex128 that is 64 + 128 GiB on a 180 GB device.  Gates before the first
    # batched one then act on B identical copies, which costs time but not memory.
    B = getattr(circuit, "batch_size", None)
    if buffer is None and B is not None and B > 1 and prep is None:
        sv = StateVector(n, dtype=dtype, device=device, batch=int(B))
    else:
        sv = StateVector(n, dtype=dtype, device=device, buffer=buffer)
    if buffer is not None:
        sv.reset()
    if prep is not None:
        vec = np.asarray(prep.state_vector(wire_order=list(range(n))))
        # single-precision StatePrep gives a complex64 simulation (initialize_state.py:47-51)
        sv.set_state(vec)
    apply_gates(sv, ops_[bool(prep):], fusion, mid_measurements, rng, debugger,
                getattr(circuit, "shots", None), exact_sampling)
    return sv, sv.batch > 1


#: set by ``pl_plugin`` to wrap PennyLane measurement processes met inside ``Snapshot`` operators
MEASUREMENT_ADAPTER = None


def apply_snapshot(op, sv: StateVector, debugger, tape_shots=None, rng=None, exact: bool = True):
    """apply_operation.py:883-917: measure the current state into ``debugger.snapshots``."""
    if debugger is None or not debugger.active:
        return
    measurement = op.hyperparameters["measurement"]
    if not hasattr(measurement, "kind") and MEASUREMENT_ADAPTER is not None:
        measurement = MEASUREMENT_ADAPTER(measurement)       # a genuine PennyLane measurement
    shots = op.hyperparameters["shots"]
    if isinstance(shots, str) and shots == "workflow":
        shots = tape_shots
    batched = sv.batch > 1
    if shots:
        snapshot = measure_with_samples([measurement], sv, shots, np.random.default_rng(rng),
                                        exact)[0]
    else:
        snapshot = measure(measurement, sv, batched)
    tag = op.hyperparameters["tag"]
    snaps = debugger.snapshots
    if tag is None:
        snaps[len(snaps)] = snapshot
    elif tag not in snaps:
        snaps[tag] = snapshot
    elif isinstance(snaps[tag], list):
        snaps[tag].append(snapshot)
    else:
        snaps[tag] = [snaps[tag], snapshot]


def _postselection_postprocess(sv: StateVector, shots, rng=None, postselect_mode=None):
    """simulate.py:120-171: after a projector, renormalise and thin the shot budget out with
    ``binomial(shots, |P psi|^2)`` (``hw-like``; ``fill-shots`` keeps it).  A zero norm gives a
    NaN state, as the reference's division does."""
    from .tape import FlexShots

    if sv.batch > 1:
        raise ValueError(
            "Cannot postselect on circuits with broadcasting. Use the "
            "qp.transforms.broadcast_expand transform to split a broadcasted "
            "tape into multiple non-broadcasted tapes before executing if "
            "postselection is used.")
    norm = float(np.sqrt(sv.norm2()))
    if np.allclose(norm, 0.0):
        if postselect_mode == "fill-shots" and shots:
            raise RuntimeError(
                "The probability of the postselected mid-circuit measurement outcome is 0. "
                "This leads to invalid results when using postselect_mode='fill-shots'.")
        norm = 0.0
    if shots:
        binomial = np.random.binomial if rng is None else rng.binomial
        shots = FlexShots(list(shots) if postselect_mode == "fill-shots"
                          else [int(binomial(s, float(norm ** 2))) for s in shots])
    with np.errstate(divide="ignore", invalid="ignore"):
        sv.scale(np.complex128(np.float64(1.0) / np.float64(norm)) if norm else np.nan)
    return shots


def apply_gates(sv: StateVector, gates, fusion: int = 0, mid_measurements=None, rng=None,
                debugger=None, tape_shots=None, exact: bool = True, postselect_mode=None):
    """The gate loop of simulate.py:213-235.  Mid-circuit measurements split the gate list:
    the unitary runs between them go through the fused path (or gate by gate), a ``MidMeasure``
    is one probs sweep + one collapse sweep, and a ``Conditional`` is decided on the host from
    the sampled values (apply_operation.py:355-411) — its base gate simply joins the current
    run when the condition holds."""
    run = []

    def flush():
        if not run:
            return
        if fusion:
            # host fusion pass + tile kernel: one state sweep per segment (compiler.py)
            sv.apply_operations_fused(list(run), level=fusion)
        else:
            for g in run:
                sv.apply_operation(g)
        run.clear()

    for op in gates:
        if is_mcm(op):
            flush()
            sv.apply_mid_measure(op, mid_measurements, rng)
        elif is_conditional(op):
            if op.meas_val.concretize(mid_measurements):
                run.append(op.base)
        elif op.name == "Snapshot":
            if debugger is not None and debugger.active:
                flush()
                apply_snapshot(op, sv, debugger, tape_shots, rng, exact)
        elif op.name == "Projector":
            # postselection on a mid-circuit measurement, simulate.py:226-232
            if getattr(op, "batch_size", None) is not None:
                raise ValueError("Cannot postselect on circuits with broadcasting.")
            run.append(op)
            flush()
            tape_shots = _postselection_postprocess(sv, tape_shots, rng, postselect_mode)
            sv.postselected_shots = tape_shots
        else:
            run.append(op)
    flush()


# ---------------------------------------------------------------------------------------------
# analytic measurements
# ---------------------------------------------------------------------------------------------
def _rotated(sv: StateVector, gates):
    """State after the diagonalizing gates; a copy only when there is something to apply."""
    if not gates:
        return sv
    rot = sv.clone()
    for g in gates:
        rot.apply_operation(g)
    return rot


def _squeeze(res, batched):
    res = np.asarray(res)
    return res if batched else (res[0] if res.ndim and res.shape[0] == 1 else res)


def _pauli_rep(obs):
    try:
        return obs.pauli_rep
    except Exception:  # pragma: no cover - foreign operator classes
        return None


def _expval_pauli(sv, ps):
    return sv.expval_pauli_sentence(ps)


def measure(mp, sv: StateVector, is_state_batched: bool = False, return_torch: bool = False):
    """One analytic measurement (measure.py:224-239).  ``return_torch`` (device option, SURVEY
    section 8 f2): ``state`` and ``probs`` — the results whose size grows with the register — stay
    on the device and come back as torch CUDA tensors; no device-to-host copy is made.

    Strategy (the kernel-side replacement of ``get_measurement_function``, measure.py:165-221):
    observables with a Pauli representation go through the fused Pauli-sum reduction (no copy of
    the state, no diagonalizing gates); other observables rotate a copy with their diagonalizing
    gates and reduce marginal probabilities against the eigenvalues.
    """
    kind = mp.kind
    obs = mp.obs
    if kind == "state":
        if return_torch:
            flat = sv.data.clone()
            return flat if is_state_batched else flat[0]
        flat = sv.data.cpu().numpy()
        return flat if is_state_batched else flat[0]
    if kind == "probs":
        rot = _rotated(sv, mp.diagonalizing_gates()) if obs is not None else sv
        wires = list(mp.wires) if len(mp.wires) else list(range(sv.n))
        if return_torch:
            p = rot.probs_device(wires)
            return p if sv.batch > 1 else p.reshape(-1)
        p = rot.probs(wires)
        return p
    if kind in ("density_matrix", "purity", "vn_entropy", "mutual_info"):
        return _measure_density(mp, sv, is_state_batched)
    if kind == "expval":
        if getattr(obs, "name", "") == "SparseHamiltonian":          # measure.py:198-199
            r = sv.expval_csr(obs.sparse_matrix(), list(obs.wires))
            return r if is_state_batched else np.float64(r[0])
        ps = _pauli_rep(obs)
        if ps is not None:
            return np.float64(_expval_pauli(sv, ps)) if not is_state_batched else _expval_pauli(sv, ps)
        if getattr(obs, "has_diagonalizing_gates", False):
            rot = _rotated(sv, mp.diagonalizing_gates())
            p = rot.probs(list(mp.wires))
            return np.dot(p, np.real(np.asarray(mp.eigvals())))
        if hasattr(obs, "terms"):                       # sum_of_terms_method, measure.py:142-161
            cs, os_ = obs.terms()
            from .measurements import ExpectationMP
            return sum(c * measure(ExpectationMP(o), sv, is_state_batched) for c, o in zip(cs, os_))
        # full_dot_products (measure.py:121-139): <psi| O |psi> with O applied to a copy
        tmp = sv.clone()
        tmp.apply_matrix(np.asarray(obs.matrix()), list(obs.wires))
        return np.real(sv.inner(tmp))
    if kind == "var":
        if getattr(obs, "has_diagonalizing_gates", False):
            rot = _rotated(sv, mp.diagonalizing_gates())
            p = rot.probs(list(mp.wires))
            ev = np.real(np.asarray(mp.eigvals(), dtype=complex)).astype("float64")
            return np.dot(p, ev**2) - np.dot(p, ev) ** 2      # var.py:107-115
        ps = _pauli_rep(obs)
        if ps is not None:                              # <H^2> - <H>^2 without diagonalising
            ps2 = ps @ ps
            return _expval_pauli(sv, _real_sentence(ps2)) - _expval_pauli(sv, ps) ** 2
        raise NotImplementedError(f"variance of {obs} is not supported")
    raise NotImplementedError(f"analytic measurement {kind} is not supported")


def _entropy(rho, base):
    """math/quantum.py:632-663 (``_compute_vn_entropy``): eigvalsh, non-positive eigenvalues
    dropped, ``entr`` = -x log x summed, divided by log(base)."""
    evs = np.linalg.eigvalsh(rho)
    evs = np.where(evs > 0, evs, 1.0)
    div = np.log(base) if base else 1
    return np.sum(-evs * np.log(evs), axis=-1) / div


def _measure_density(mp, sv: StateVector, is_state_batched: bool):
    """``qml.density_matrix`` / ``purity`` / ``vn_entropy`` / ``mutual_info``
    (measurements/purity.py:50-54, vn_entropy.py:65-67, mutual_info.py:92-100).  The reference
    expands the state into the full 4^n density matrix and traces it down; here
    ``StateVector.reduced_dm`` reads the statevector and only 2^m x 2^m matrices reach the host."""
    kind = mp.kind
    if kind == "mutual_info":
        w0, w1 = (list(w) for w in mp._wires)
        base = getattr(mp, "log_base", None)
        s0 = _entropy(sv.reduced_dm(w0), base)
        s1 = _entropy(sv.reduced_dm(w1), base)
        s01 = _entropy(sv.reduced_dm(sorted(w0 + w1)), base)
        return s0 + s1 - s01
    rho = sv.reduced_dm(list(mp.wires))
    if kind == "density_matrix":
        return rho
    if kind == "purity":                       # math/quantum.py:563-589: Re tr(rho rho)
        return np.real(np.einsum("...ab,...ba->...", rho, rho))
    return _entropy(rho, getattr(mp, "log_base", None))


def _real_sentence(ps):
    out = PauliSentence()
    for w, c in ps.items():
        if abs(c) > 0:
            out[w] = np.real(c)
    return out


# ---------------------------------------------------------------------------------------------
# finite shots
# ---------------------------------------------------------------------------------------------
def _pauli_word_of(obs):
    if obs is None or obs.name in ("LinearCombination", "Hamiltonian", "Sum"):
        return None
    ps = _pauli_rep(obs)
    if ps is None or len(ps) != 1:
        return None
    (w, _c), = ps.items()
    return w


def _qwc(w1, w2) -> bool:
    return all(w1[k] == w2[k] for k in w1 if k in w2)


def _qwc_partition(words):
    """``compute_partition_indices(observables, "qwc", "lf")`` (pauli/grouping/
    group_observables.py:389-432) without rustworkx: greedy colouring of the graph whose edges
    join the words that do NOT commute qubit-wise (:340-386, :168-193), nodes taken by descending
    degree — ``graph_greedy_color``'s largest-first order, a stable sort, so ties stay in index
    order — each getting the smallest colour none of its neighbours has; the groups come out in
    the order of the lowest index of each colour, indices ascending inside a group (:238-243).
    The order matters: it is the order in which the shot budget's uniforms are drawn."""
    m = len(words)
    if all(len(w) == 0 for w in words):                     # :389-394: nothing acts on a wire
        return [list(range(m))]
    adj = [[j for j in range(m) if j != i and not _qwc(words[i], words[j])] for i in range(m)]
    colour: dict = {}
    for i in sorted(range(m), key=lambda k: -len(adj[k])):
        used = {colour[j] for j in adj[i] if j in colour}
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    groups: dict = {}
    for i in range(m):
        groups.setdefault(colour[i], []).append(i)
    return list(groups.values())


def _group_measurements(mps):
    """sampling.py:46-98.  Pauli-word observables are partitioned into qubit-wise commuting
    groups by :func:`_qwc_partition` (the reference's largest-first colouring)."""
    if len(mps) == 1:
        return [list(mps)], [[0]]
    pauli, other, other_idx, no_obs, no_obs_idx = [], [], [], [], []
    for i, mp in enumerate(mps):
        if mp.kind in ("shadow", "shadow_expval"):     # sampling.py:68-70: a group of its own
            other.append([mp]); other_idx.append([i])
        elif mp.obs is None:
            no_obs.append(mp); no_obs_idx.append(i)
        elif _pauli_word_of(mp.obs) is not None:
            pauli.append((i, mp))
        else:
            other.append([mp]); other_idx.append([i])
    groups, gidx = [], []
    if pauli:
        for part in _qwc_partition([_pauli_word_of(mp.obs) for _, mp in pauli]):
            groups.append([pauli[k][1] for k in part])
            gidx.append([pauli[k][0] for k in part])
    if no_obs:
        groups.append(no_obs); gidx.append(no_obs_idx)
    return groups + other, gidx + other_idx


def _group_diagonalizing_gates(mps):
    """sampling.py:188-202 (and pauli/utils.py:1059-1081 for a QWC group)."""
    from . import ops as _ops

    if len(mps) == 1:
        return mps[0].diagonalizing_gates()
    if all(mp.obs is not None for mp in mps):
        full = {}
        for mp in mps:
            for wire, ch in _pauli_word_of(mp.obs).items():
                full.setdefault(wire, ch)
        gates = []
        for w, ch in full.items():
            if ch == "X":
                gates.append(_ops.RY(-np.pi / 2, wires=w))
            elif ch == "Y":
                gates.append(_ops.RX(np.pi / 2, wires=w))
        return gates
    return []


def sample_state(sv: StateVector, shots: int, rng, wires=None, exact: bool = True):
    """sampling.py:439-476 + :500-531 on the device (uniforms from the host Generator)."""
    return sv.sample(shots, rng, wires=wires, exact=exact)


def _measure_group(mps, sv, shots, rng, exact):
    """sampling.py:276-335."""
    rot = _rotated(sv, _group_diagonalizing_gates(mps))
    wires = list(range(sv.n))
    samples = sample_state(rot, shots.total_shots, rng, wires=wires, exact=exact)
    processed = []
    for lower, upper in shots.bins():
        processed.append(tuple(mp.process_samples(samples[..., lower:upper, :], wires) for mp in mps))
    if shots.has_partitioned_shots:
        return tuple(zip(*processed))
    return processed[0]


def _measure_sum(mps, sv, shots, rng, exact):
    """sampling.py:377-436: term-by-term sampling of LinearCombination / Sum expectation values."""
    from .measurements import ExpectationMP
    from .tape import Shots

    mp = mps[0]
    cs, os_ = mp.obs.terms()

    def one(s):
        res = measure_with_samples([ExpectationMP(o) for o in os_], sv, Shots(s), rng, exact)
        return sum(c * r for c, r in zip(cs, res))

    unsq = tuple(one(s) for s in shots)
    return [unsq] if shots.has_partitioned_shots else [unsq[0]]


_SHADOW_OBS = np.array([[[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]], dtype=complex)
_H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
# classical_shadow.py:185-191: H, H RZ(-pi/2), I
_SHADOW_DIAG = np.stack([_H, _H @ np.diag([np.exp(0.25j * np.pi), np.exp(-0.25j * np.pi)]),
                         np.eye(2, dtype=complex)])


def _classical_shadow(mp, sv: StateVector, shots: int, rng):
    """``ClassicalShadowMP.process_state_with_shots`` (measurements/classical_shadow.py:142-257).

    The reference stacks ``shots`` copies of the state (``shots * S`` bytes) and walks the wires:
    single-wire density matrix, expectation of the recipe's Pauli, ``bit_rng.random(shots) >
    probs``, collapse, renormalise.  Here one device-resident copy is walked per shot: the
    single-wire density matrix is ``b200q_gram_block`` restricted to the part of the state where
    the wires measured so far are 0, and "rotate into the recipe's basis, keep the sampled row,
    renormalise" is one controlled 2x2 sweep that parks the survivor in the wire's |0> half — so
    every step touches half of what the step before did, like the reference's shrinking stack.
    The uniforms are drawn in the reference's order (all shots of wire 0, then wire 1, ...)."""
    wires = list(mp.wires)
    nq = len(wires)
    recipes = np.random.RandomState(mp.seed).randint(0, 3, size=(shots, nq))   # :171-172
    bit_rng = np.random.default_rng(rng)
    uniforms = np.stack([bit_rng.random(size=shots) for _ in range(nq)], axis=1)
    outcomes = np.zeros((shots, nq))
    work = sv.clone()
    for t in range(shots):
        if t:
            work.data.copy_(sv.data)
        done = []
        for q, w in enumerate(wires):
            r = recipes[t, q]
            rho = work.reduced_dm([w], fixed_zero=done)
            prob = (np.einsum("bc,cb->", rho, _SHADOW_OBS[r]) + 1) / 2            # :239
            sample = int(uniforms[t, q] > prob.real)
            outcomes[t, q] = sample
            p_s = prob.real if sample == 0 else 1 - prob.real
            row = _SHADOW_DIAG[r][sample] / np.sqrt(p_s)
            ctrl = done[-16:]
            work.apply_matrix(np.array([row, [0, 0]]), [w], ctrl, [0] * len(ctrl))
            done.append(w)
    return np.stack([outcomes, recipes]).astype(np.int8)


def _shadow_expval(mp, sv: StateVector, shots: int, rng):
    """``ShadowExpvalMP.process_state_with_shots`` (measurements/classical_shadow.py:490-514):
    a shadow on the observable's wires, then the host estimator."""
    from .shadows import shadow_expval

    bits, recipes = _classical_shadow(mp, sv, shots, rng)
    return shadow_expval(bits, recipes, mp.H, mp.k, wire_map=list(mp.wires))


def _measure_classical_shadow(mps, sv, shots, rng):
    """sampling.py:338-374."""
    mp = mps[0]
    fn = _shadow_expval if mp.kind == "shadow_expval" else _classical_shadow
    if shots.has_partitioned_shots:
        return [tuple(fn(mp, sv, s, rng) for s in shots)]
    return [fn(mp, sv, shots.total_shots, rng)]


def measure_with_samples(mps, sv: StateVector, shots, rng, exact: bool = True,
                         mid_measurements=None):
    """sampling.py:205-273."""
    mps = list(mps)
    if mid_measurements:
        # the last N measurements are the sampled MCMs of the one-shot tape (sampling.py:235-236)
        mps = mps[: len(mps) - len(mid_measurements)]
    groups, indices = _group_measurements(mps)
    all_res = []
    for group in groups:
        mp0 = group[0]
        if mp0.kind == "expval" and mp0.obs is not None and mp0.obs.name in (
                "LinearCombination", "Hamiltonian", "Sum"):
            all_res.extend(_measure_sum(group, sv, shots, rng, exact))
        elif mp0.kind in ("shadow", "shadow_expval"):
            all_res.extend(_measure_classical_shadow(group, sv, shots, rng))
        else:
            all_res.extend(_measure_group(group, sv, shots, rng, exact))
    flat_indices = [i for idx in indices for i in idx]
    sorted_res = tuple(r for _, r in sorted(enumerate(all_res), key=lambda t: flat_indices[t[0]]))
    if mid_measurements:
        sorted_res += tuple(mid_measurements.values())          # sampling.py:266-267
    if shots.has_partitioned_shots:
        sorted_res = tuple(zip(*sorted_res))
    return sorted_res


def measure_final_state(circuit, sv: StateVector, is_state_batched: bool, rng=None,
                        exact_sampling: bool = True, mid_measurements=None, return_torch: bool = False):
    """simulate.py:246-304."""
    if not circuit.shots:
        if mid_measurements is not None:
            raise TypeError("Native mid-circuit measurements are only supported with finite shots.")
        if len(circuit.measurements) == 1:
            return measure(circuit.measurements[0], sv, is_state_batched, return_torch)
        return tuple(measure(mp, sv, is_state_batched, return_torch) for mp in circuit.measurements)
    rng = np.random.default_rng(rng)
    results = measure_with_samples(circuit.measurements, sv, circuit.shots, rng, exact_sampling,
                                   mid_measurements=mid_measurements)
    if len(circuit.measurements) == 1:
        if circuit.shots.has_partitioned_shots:
            return tuple(res[0] for res in results)
        return results[0]
    return results


def simulate_one_shot_native_mcm(circuit, sv: StateVector, gates, rng, exact_sampling: bool = True,
                                 fusion: int = 0, debugger=None):
    """simulate.py:947-990: one shot of a tape with native mid-circuit measurements.  ``sv``
    already holds the state in front of ``gates`` (the part of the tape from its first
    ``MidMeasure`` on)."""
    mid_measurements = {}
    apply_gates(sv, gates, fusion, mid_measurements, rng, debugger, circuit.shots, exact_sampling)
    return measure_final_state(circuit, sv, False, rng=rng, exact_sampling=exact_sampling,
                               mid_measurements=mid_measurements)


class _Branch:
    """A node of the outcome tree: the state reached by one sequence of mid-circuit outcomes,
    parked in front of the next ``MidMeasure`` (or final), and that measurement's marginal."""
    __slots__ = ("state", "p", "children")

    def __init__(self, state, p=None):
        self.state, self.p, self.children = state, p, {}


def _simulate_native_mcm(circuit, rng, dtype, device, exact_sampling, fusion, debugger=None,
                         cache_fraction: float = 0.5):
    """The one-shot loop of simulate.py:356-381 (``mcm_method="one-shot"``): every shot re-runs
    the tape with ``shots=[1]`` and returns its own result tuple; ``dynamic_one_shot``'s
    post-processing (above the device boundary) combines them.

    The reference re-simulates the whole tape per shot.  Here
    * everything in front of the first ``MidMeasure`` is shot-independent and simulated ONCE;
    * the state a shot reaches depends only on its outcomes so far, so collapsed branch states
      stay resident in HBM in a tree keyed by outcome (up to ``cache_fraction`` of the free
      memory; beyond that a shot continues in a scratch buffer).  A shot whose branch is cached
      costs its ``binomial`` draws on cached marginals plus a search in the leaf's cached CDF.
    The host Generator is consumed exactly as in the reference's loop — one ``binomial`` per
    measurement in tape order, then the terminal draws — so per-shot results are bit-identical
    to the oracle's under the same seed (this is the one-shot method's stream, not
    tree-traversal's, which the reference samples differently)."""
    if not circuit.shots:
        raise TypeError("Native mid-circuit measurements are only supported with finite shots.")
    from .tape import Shots

    rng = np.random.default_rng(rng)
    ops_ = list(circuit.operations)
    snap = debugger is not None and debugger.active       # snapshots are recorded per shot
    first = next(i for i, op in enumerate(ops_)
                 if is_mcm(op) or (snap and op.name == "Snapshot"))
    base, batched = _prefix_state(ops_[:first], circuit.num_wires, dtype, device, fusion)
    if batched:
        raise ValueError("MidMeasure cannot be applied to batched states.")
    rest = ops_[first:]
    aux = _OneShotView(circuit, Shots([1]))
    n_shots = circuit.shots.total_shots
    results = []
    if snap or any(op.name == "Projector" for op in rest) or cache_fraction <= 0:
        # side effects per shot (snapshots, shot thinning): plain loop from the prefix state
        work = base.clone()
        for i in range(n_shots):
            if i:
                work.data.copy_(base.data)
            results.append(simulate_one_shot_native_mcm(aux, work, rest, rng, exact_sampling,
                                                        fusion, debugger))
        return tuple(results)

    mcms, segs = [], []
    for op in rest:                                        # rest starts with a MidMeasure
        if is_mcm(op):
            mcms.append(op)
            segs.append([])
        else:
            segs[-1].append(op)
    state_bytes = base.data.numel() * base.data.element_size()
    budget = max(0, int(_free_bytes(base.device) * cache_fraction) // state_bytes - 1)
    root = _Branch(base, base.probs([mcms[0].wires[0]]))
    work = None
    for _ in range(n_shots):
        mm, node, scratch = {}, root, None
        for j, mcm in enumerate(mcms):
            if scratch is not None:                        # off the cached tree
                scratch.apply_mid_measure(mcm, mm, rng)
                apply_gates(scratch, segs[j], fusion, mm, rng)
                continue
            sample, scale = StateVector.mid_measure_draw(node.p, node.state.np_dtype, rng)
            mm[mcm] = sample
            child = node.children.get(sample)
            if child is None:
                leaf = j + 1 == len(mcms)
                cost = 2 if leaf else 1                    # a leaf also keeps its sampling CDF
                if budget >= cost:
                    budget -= cost
                    target = node.state.clone()
                else:
                    if work is None:
                        work = node.state.clone()
                    else:
                        work.data.copy_(node.state.data)
                    target = scratch = work
                target.collapse(mcm.wires[0], sample, bool(getattr(mcm, "reset", False)), scale)
                apply_gates(target, segs[j], fusion, mm, rng)
                if scratch is None:
                    nxt = None if leaf else target.probs([mcms[j + 1].wires[0]])
                    target.frozen = leaf                   # terminal draws reuse its CDF
                    child = node.children[sample] = _Branch(target, nxt)
            if child is not None:
                node = child
        final = scratch if scratch is not None else node.state
        results.append(measure_final_state(aux, final, False, rng=rng,
                                           exact_sampling=exact_sampling, mid_measurements=mm))
    return tuple(results)


def _free_bytes(device):
    import torch

    return torch.cuda.mem_get_info(device)[0]


class _OneShotView:
    """``circuit.copy(shots=[1])`` (simulate.py:359) without copying the tape."""

    def __init__(self, circuit, shots):
        self._c = circuit
        self.shots = shots

    def __getattr__(self, name):
        return getattr(self._c, name)


def _prefix_state(ops_, n, dtype, device, fusion):
    prep = ops_[0] if ops_ and _is_prep(ops_[0]) else None
    sv = StateVector(n, dtype=dtype, device=device)
    if prep is not None:
        sv.set_state(np.asarray(prep.state_vector(wire_order=list(range(n)))))
    apply_gates(sv, ops_[bool(prep):], fusion)
    return sv, sv.batch > 1


def simulate(circuit, rng=None, dtype=np.complex128, device=None, exact_sampling: bool = True,
             state_cache=None, fusion: int = 0, debugger=None, mcm_method=None,
             return_torch: bool = False):
    """simulate.py:308-393.  Tapes with ``MidMeasure`` operations take the tree-traversal path
    when ``mcm_method == "tree-traversal"`` (:350-352, tree_mcm.py) and the native one-shot path
    otherwise (:356-381)."""
    circuit = circuit.map_to_standard_wires()
    if any(is_mcm(op) for op in circuit.operations):
        if mcm_method == "tree-traversal":
            from .tree_mcm import simulate_tree_mcm
            return simulate_tree_mcm(circuit, rng, dtype, device, exact_sampling, fusion, debugger)
        return _simulate_native_mcm(circuit, rng, dtype, device, exact_sampling, fusion, debugger)
    if debugger is not None and debugger.active and circuit.shots:
        rng = np.random.default_rng(rng)        # snapshots and final sampling share one stream
    if circuit.shots and any(op.name == "Projector" for op in circuit.operations):
        rng = np.random.default_rng(rng)        # the shot thinning and the sampling share a stream
    sv, batched = get_final_state(circuit, dtype=dtype, device=device, fusion=fusion, rng=rng,
                                  debugger=debugger, exact_sampling=exact_sampling)
    if getattr(sv, "postselected_shots", None) is not None:
        circuit = _OneShotView(circuit, sv.postselected_shots)     # circuit._shots = new_shots, :232
    if state_cache is not None:
        state_cache[circuit.hash] = sv
    return measure_final_state(circuit, sv, batched, rng=rng, exact_sampling=exact_sampling,
                               return_torch=return_torch)


__all__ = ["get_final_state", "apply_gates", "measure", "measure_with_samples",
           "measure_final_state", "simulate", "simulate_one_shot_native_mcm", "sample_state",
           "PauliWord"]
