"""Statevector sharded over ``G = 2**g`` ranks (one process per GPU) — SURVEY.md section 8(e).

north_star: "the statevector is partitioned across the GPUs by the high-order qubit bits, and
gates on global qubits trigger qubit-remapping swaps via NCCL send/recv over NVLink".  The
reference has no analogue (default.qubit holds one numpy array; its only multi-worker mode is a
process pool over independent tapes, default_qubit.py:810-829).

Layout.  The flat index of default.qubit's state (initialize_state.py:43-44) has one bit per
wire: *logical* bit ``b = n-1-wire``.  A host-side permutation ``phys[b]`` sends logical bits to
*physical* bits; physical bits ``0..nl-1`` (``nl = n-g``) index the rank's shard, physical bits
``nl..n-1`` are the rank id.  Every rank holds ``2**nl`` amplitudes per batch element.

Gates.  A gate needs no communication when every wire it acts on *non-diagonally* is local:
controls and diagonal factors on rank bits are resolved on the host (``specialise``): the rank
either skips the gate, drops the control, or flips the sign of the angle — the shard then sees an
ordinary gate on ``nl`` wires and the single-GPU engine (fused segments included) runs it
unchanged.  When a non-diagonal target sits on a rank bit the planner (``plan``) inserts a
*remap*: ``k`` rank bits are exchanged with the top ``k`` local bits.  With that choice the data
a rank sends to each of its ``2**k - 1`` partners is one contiguous slice of its shard and what
it receives lands in the very same slice, so the exchange is in place through bounded staging
buffers; nothing the size of a second shard is allocated.  Victims are chosen Belady-style
(the local bits whose next non-diagonal use is furthest away) and moved to the top local
positions with ordinary SWAP gates that ride in the preceding fused segment.

The exchange (K9).  On CUDA the slab pieces are PUSHED into the partners' staging buffers by
copy-engine copies whose destination is peer memory (``torch.distributed._symmetric_memory``
supplies the mapping), *landed* / *consumed* flags are stream memory operations, and the staging
buffer is unpacked into place by a TMA kernel that shares the SMs with the sweeps
(``csrc/remap.cu``); three staging buffers rotate on two communication streams.  An exchange is
cut into pieces along index bits that the segments before and after it leave alone, and those
segments run piece by piece (partial launches) so that transfer and sweeps overlap
(``_schedule`` / ``_run_window``).  Grouped ``isend``/``irecv`` (``ncclSend``/``ncclRecv`` under
``torch.distributed``) stay as the fallback when the symmetric allocation is refused, and are
what the ``gloo`` tests exercise.

Reductions.  Pauli-sum expectation values: every rank evaluates the terms whose X/Y factors are
local (Z factors on rank bits are signs), partial sums are all-gathered and added in rank order
(bit-identical run to run); terms that flip a rank bit are measured after one more remap.
Marginal probabilities: local marginals placed by rank bits, same ordered sum.  Sampling: the map
is first restored to the identity, so rank ``r`` holds the logical indices
``[r 2**nl, (r+1) 2**nl)``; numpy's pairwise ``sum`` is the same tree over ranks, the sequential
``cumsum`` carries its running value from rank to rank, and ``searchsorted`` is the sum over
ranks of local counts — samples stay bit-identical to sampling.py:500-531 under a fixed seed.

The numerical work is done by a *local engine* (``CudaEngine`` below: a ``StateVector`` and the
C ABI).  The class is written against the small engine interface so that the host logic (plan,
maps, exchanges, reductions) is testable under ``gloo`` on CPU with a test-only engine; the
product constructs ``CudaEngine`` only and never falls back.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import ops as _ops

# ---------------------------------------------------------------------------------------------
# which wires of an operator must be local?
# ---------------------------------------------------------------------------------------------
_DIAG_NAMES = {"PauliZ", "S", "T", "CZ", "CCZ", "PhaseShift", "U1", "ControlledPhaseShift", "RZ",
               "MultiRZ", "IsingZZ", "CRZ", "GlobalPhase", "Identity", "Barrier", "Snapshot",
               "WireCut", "DiagonalQubitUnitary"}
_FIXED_PHASE = {"PauliZ": -1.0 + 0j, "S": 1j, "T": np.exp(0.25j * np.pi), "CZ": -1.0 + 0j,
                "CCZ": -1.0 + 0j}
_CTRL_FIXED = {"CNOT": (1, "PauliX"), "Toffoli": (2, "PauliX"), "CY": (1, "PauliY"),
               "CH": (1, "Hadamard"), "CSWAP": (1, "SWAP")}
_CTRL_PARAM = {"CRX": "RX", "CRY": "RY", "CRot": "Rot"}


def _is_generic_controlled(op) -> bool:
    return (getattr(op, "base", None) is not None and len(getattr(op, "control_wires", ()) or ())
            and (op.name.startswith("C(") or op.name == "ControlledQubitUnitary"))


def op_locality(op):
    """(must_local, may_be_global): wires the operator acts on non-diagonally / only as a
    control or a diagonal factor."""
    name = op.name
    wires = list(op.wires)
    if name in _DIAG_NAMES:
        return [], wires
    if name == "PauliRot":
        word = op.hyperparameters["pauli_word"]
        must = [w for w, c in zip(wires, word) if c in "XY"]
        return must, [w for w in wires if w not in must]
    if name in _CTRL_FIXED:
        nc = _CTRL_FIXED[name][0]
        return wires[nc:], wires[:nc]
    if name in _CTRL_PARAM:
        return wires[1:], wires[:1]
    if name == "MultiControlledX":
        return wires[-1:], wires[:-1]
    if _is_generic_controlled(op):
        cw = list(op.control_wires)
        return [w for w in wires if w not in cw], cw
    return wires, []


def _phase_on_ones(phase, local_wires, spare_wire):
    """Multiply the subspace where all ``local_wires`` are 1 by ``phase`` (scalar or (B,))."""
    phase = np.asarray(phase, dtype=complex)
    one = np.ones_like(phase)
    if not local_wires:
        return _ops.DiagonalQubitUnitary(np.stack([phase, phase], axis=-1), wires=[spare_wire])
    d = _ops.DiagonalQubitUnitary(np.stack([one, phase], axis=-1), wires=[local_wires[-1]])
    if len(local_wires) == 1:
        return d
    return _ops.Controlled(d, list(local_wires[:-1]))


def specialise(op, gvals: dict, wmap: dict, spare_wire=0):
    """The operator a rank applies to its shard: ``gvals`` maps the operator's wires that sit on
    rank bits to this rank's bit value, ``wmap`` maps the others to local wire labels.  Returns
    ``None`` when the gate is the identity on this rank."""
    name = op.name
    wires = list(op.wires)
    if name in ("Identity", "Barrier", "Snapshot", "WireCut"):
        return None
    if not any(w in gvals for w in wires):
        return op.map_wires(wmap)
    loc = [wmap[w] for w in wires if w not in gvals]
    if name == "GlobalPhase":
        new = op.map_wires(wmap)
        new.wires = tuple(loc)
        return new
    if name in _FIXED_PHASE or name in ("PhaseShift", "U1", "ControlledPhaseShift"):
        if any(gvals[w] == 0 for w in wires if w in gvals):
            return None
        ph = _FIXED_PHASE[name] if name in _FIXED_PHASE else np.exp(1j * np.asarray(op.data[0], dtype=float))
        return _phase_on_ones(ph, loc, spare_wire)
    if name in ("RZ", "MultiRZ", "IsingZZ"):
        sign = -1.0 if sum(gvals[w] for w in wires if w in gvals) & 1 else 1.0
        th = sign * np.asarray(op.data[0], dtype=float)
        if not loc:
            return _phase_on_ones(np.exp(-0.5j * th), [], spare_wire)
        return _ops.RZ(th, wires=loc[0]) if len(loc) == 1 else _ops.MultiRZ(th, wires=loc)
    if name == "CRZ":
        c, t = wires
        th = np.asarray(op.data[0], dtype=float)
        if c in gvals:
            if gvals[c] == 0:
                return None
            return specialise(_ops.RZ(th, wires=t), gvals, wmap, spare_wire)
        sign = -1.0 if gvals[t] else 1.0
        return _phase_on_ones(np.exp(-0.5j * sign * th), [wmap[c]], spare_wire)
    if name == "PauliRot":
        word = op.hyperparameters["pauli_word"]
        sign, lw, lword = 1.0, [], ""
        for w, ch in zip(wires, word):
            if w in gvals:
                if ch in "XY":
                    raise ValueError("PauliRot with an X/Y factor on a rank bit needs a remap")
                if ch == "Z" and gvals[w]:
                    sign = -sign
            else:
                lw.append(wmap[w]); lword += ch
        th = sign * np.asarray(op.data[0], dtype=float)
        if not lw or all(ch == "I" for ch in lword):
            return _phase_on_ones(np.exp(-0.5j * th), [], spare_wire)
        return _ops.PauliRot(th, lword, wires=lw)
    if name == "DiagonalQubitUnitary":
        D = np.asarray(op.data[0], dtype=complex)
        k = len(wires)
        batched = D.ndim == 2
        sub = D.reshape((D.shape[0],) * batched + (2,) * k)
        # index the rank-bit axes, highest axis first so the remaining positions stay valid
        for ax in reversed(range(k)):
            if wires[ax] in gvals:
                sub = np.take(sub, gvals[wires[ax]], axis=ax + batched)
        if not loc:
            ph = sub.reshape(-1) if batched else sub.reshape(())
            return _phase_on_ones(ph, [], spare_wire)
        return _ops.DiagonalQubitUnitary(sub.reshape((D.shape[0], -1) if batched else (-1,)), wires=loc)
    if name in _CTRL_FIXED or name in _CTRL_PARAM:
        nc = _CTRL_FIXED[name][0] if name in _CTRL_FIXED else 1
        cw, tw = wires[:nc], wires[nc:]
        if any(gvals.get(w, 1) == 0 for w in cw):
            return None
        rem = [wmap[w] for w in cw if w not in gvals]
        base_cls = getattr(_ops, _CTRL_FIXED[name][1] if name in _CTRL_FIXED else _CTRL_PARAM[name])
        base = base_cls(*op.data, wires=[wmap[w] for w in tw])
        if not rem:
            return base
        if name == "Toffoli" and len(rem) == 1:
            return _ops.CNOT(wires=rem + [wmap[tw[0]]])
        return _ops.Controlled(base, rem)
    if name == "MultiControlledX":
        cw, cv = wires[:-1], [int(bool(v)) for v in op.control_values]
        if any(w in gvals and gvals[w] != v for w, v in zip(cw, cv)):
            return None
        rem = [(wmap[w], v) for w, v in zip(cw, cv) if w not in gvals]
        t = wmap[wires[-1]]
        if not rem:
            return _ops.PauliX(wires=t)
        return _ops.MultiControlledX(wires=[w for w, _ in rem] + [t], control_values=[v for _, v in rem])
    if _is_generic_controlled(op):
        cw = list(op.control_wires)
        cv = [int(bool(v)) for v in op.control_values]
        if any(w in gvals and gvals[w] != v for w, v in zip(cw, cv)):
            return None
        base = op.base.map_wires(wmap)
        rem = [(wmap[w], v) for w, v in zip(cw, cv) if w not in gvals]
        if not rem:
            return base
        return _ops.Controlled(base, [w for w, _ in rem], [v for _, v in rem])
    raise ValueError(f"{name} acts non-diagonally on a rank bit: the planner must remap first")


# ---------------------------------------------------------------------------------------------
# planner
# ---------------------------------------------------------------------------------------------
@dataclass
class RunStep:
    ops: list                      # original operators (wire labels of the circuit)
    phys: list                     # logical bit -> physical bit while these run
    swaps: list = field(default_factory=list)   # (pa, pb): local physical-bit swaps AFTER the ops


@dataclass
class ExchangeStep:
    rank_bits: list                # rank-bit indices (physical bit - nl), ascending
    k: int                         # exchanged with local physical bits nl-k .. nl-1, in order


def _next_use(ops_, start, n, bit_of):
    """For every logical bit: index of the first operator at or after ``start`` that needs it
    local (len(ops_) + something if never)."""
    nxt = [len(ops_) + 1 + b for b in range(n)]         # stable tie-break
    seen = set()
    for i in range(start, len(ops_)):
        must, _ = op_locality(ops_[i])
        for w in must:
            b = bit_of(w)
            if b not in seen:
                seen.add(b)
                nxt[b] = i
        if len(seen) == n:
            break
    return nxt


def plan_remap(phys, nl, want_local, nxt):
    """Choose the bits to exchange: returns (local swaps, ExchangeStep, new phys).

    ``want_local``: logical bits that must become local now.  Every other rank-resident bit
    whose next use precedes that of the best remaining victim comes along (one exchange of k
    bits moves ``1 - 2**-k`` of the shard; k exchanges of one bit move ``k/2``)."""
    n = len(phys)
    phys = list(phys)
    glob = [b for b in range(n) if phys[b] >= nl]
    local = [b for b in range(n) if phys[b] < nl]
    need = [b for b in want_local if phys[b] >= nl]
    cand = sorted((b for b in local if b not in want_local), key=lambda b: -nxt[b])
    incoming = list(need)
    victims = cand[: len(incoming)]
    if len(victims) < len(incoming):
        raise ValueError("not enough local qubits to remap (gate wider than the shard?)")
    rest_g = sorted((b for b in glob if b not in incoming), key=lambda b: nxt[b])
    ci = len(victims)
    for b in rest_g:
        if ci < len(cand) and nxt[cand[ci]] > nxt[b]:
            incoming.append(b); victims.append(cand[ci]); ci += 1
    k = len(incoming)
    # move the victims to the top k local positions
    swaps = []
    at = {phys[b]: b for b in range(n)}
    top = list(range(nl - k, nl))
    vict_set = set(victims)
    free_top = [p for p in top if at[p] not in vict_set]
    for v in victims:
        if phys[v] >= nl - k:
            continue
        p, t = phys[v], free_top.pop()
        other = at[t]
        swaps.append((p, t))
        phys[v], phys[other] = t, p
        at[t], at[p] = v, other
    incoming.sort(key=lambda b: phys[b])
    rank_bits = [phys[b] - nl for b in incoming]
    for i, b in enumerate(incoming):
        slot = nl - k + i
        v = at[slot]
        phys[v], phys[b] = phys[b], slot
    return swaps, ExchangeStep(rank_bits, k), phys


def plan(ops_, n, g, phys=None, bit_of=None, defer=None):
    """Cut an operator list into run steps separated by exchanges.  Returns (steps, final phys).

    List scheduling over the circuit's dependency order: a run step takes every operator that
    is executable under the current map and not behind a blocked operator on any of its wires
    (operators on disjoint wires commute), so one exchange is amortised over as many gates as
    the dependencies allow — a layered ansatz costs about one exchange per layer, not one per
    gate on a rank bit."""
    nl = n - g
    if bit_of is None:
        bit_of = lambda w: n - 1 - int(w)            # noqa: E731
    phys = list(range(n)) if phys is None else list(phys)
    if defer is None:
        defer = os.environ.get("B200Q_SHARD_DEFER", "1") != "0"
    deferred_once: set = set()
    steps = []
    remaining = list(ops_)
    while remaining:
        run, keep, blocked = [], [], set()
        for op in remaining:
            must, _ = op_locality(op)
            allb = {bit_of(w) for w in op.wires}
            if len(must) > nl:
                raise ValueError(f"{op.name} acts on more qubits than one shard holds")
            if (blocked & allb) or any(phys[bit_of(w)] >= nl for w in must):
                blocked |= allb
                keep.append(op)
            else:
                run.append(op)
        if not keep:
            steps.append(RunStep(run, list(phys)))
            break
        first_must = [bit_of(w) for w in op_locality(keep[0])[0]]
        if not any(phys[b] >= nl for b in first_must):      # pragma: no cover - defensive
            raise RuntimeError("sharded planner made no progress")
        nxt = _next_use(keep, 0, n, bit_of)
        swaps, ex, new_phys = plan_remap(phys, nl, first_must, nxt)
        if defer:
            leaving = {b for b in range(n) if new_phys[b] >= nl and phys[b] < nl}
            run, keep = _defer_trailing_1q(run, keep, remaining, leaving, bit_of, deferred_once)
        steps.append(RunStep(run, list(phys), swaps))
        steps.append(ex)
        phys, remaining = new_phys, keep
    return steps, phys


def _free_bit_window(busy, top, pb_max, min_bit):
    """Highest run of ``pb`` consecutive index bits in [min_bit, top) that ``busy`` leaves free,
    for the largest ``pb <= pb_max`` that has one: (lo, pb) or None."""
    for pb in range(pb_max, 0, -1):
        m = (1 << pb) - 1
        for lo in range(top - pb, min_bit - 1, -1):
            if not (busy >> lo) & m:
                return lo, pb
    return None


def _defer_trailing_1q(run, keep, remaining, leaving, bit_of, deferred_once):
    """Move the trailing single-qubit gates of a run step into the next one when the first gate
    waiting on their wire is a CNOT targeting it.

    A greedy run step of a layered ansatz is [rest of the CNOT ring of layer l][rotations of layer
    l+1]: the rotations are runnable, the CNOTs after them are not.  In that order the fusion
    pass cannot fold a CNOT into the rotation block on its target (the next CNOT of the ring
    reads that wire as its control in between), and every CNOT stays a record of its own.  Run
    one step later, the rotations precede their CNOT and the pair is ONE controlled-select record
    (compiler.merge_blocks).  Wires about to become rank bits are left alone (their gates would
    be blocked in the next step), and a gate is deferred at most once (termination)."""
    first_keep = {}
    for op in keep:
        for w in op.wires:
            first_keep.setdefault(bit_of(w), op)
    moved, closed = set(), set()
    for op in reversed(run):
        bits = {bit_of(w) for w in op.wires}
        if len(bits) == 1:
            (b,) = bits
            nxt_op = first_keep.get(b)
            if (b not in closed and b not in leaving and id(op) not in deferred_once
                    and nxt_op is not None and nxt_op.name == "CNOT"
                    and bit_of(nxt_op.wires[1]) == b):
                moved.add(id(op))
                continue
        closed |= bits
    if not moved:
        return run, keep
    deferred_once |= moved
    kept = {id(op) for op in keep} | moved
    return [op for op in run if id(op) not in moved], [op for op in remaining if id(op) in kept]


# ---------------------------------------------------------------------------------------------
# the CUDA local engine
# ---------------------------------------------------------------------------------------------
class CudaEngine:
    """Local engine over a :class:`~pennylane_b200.statevector.StateVector` (the product path)."""

    def __init__(self, nl, dtype=np.complex128, batch=1, device=None, fusion=1):
        from .statevector import StateVector

        self.sv = StateVector(nl, dtype=dtype, batch=batch, device=device)
        self.fusion = fusion

    # -- state ---------------------------------------------------------------------------
    n = property(lambda self: self.sv.n)
    batch = property(lambda self: self.sv.batch)
    data = property(lambda self: self.sv.data)
    device = property(lambda self: self.sv.device)

    def reset(self, index=None):
        """|0..0> restricted to this shard: ``index`` = local index of the single 1 or None."""
        if index is not None:
            self.sv.reset(int(index))
        else:
            self.sv.data.zero_()           # cudaMemset: a shard that holds no basis-state 1

    def set_local_state(self, arr):
        self.sv.set_state(arr)

    def resize_batch(self, batch):
        self.sv._resize_batch(batch)

    # -- gates ---------------------------------------------------------------------------
    def compile(self, local_ops):
        if not self.fusion:
            return ("ops", list(local_ops))
        segs = self.sv.compile_fused(local_ops, level=self.fusion)
        self.sv.prepare_segments(segs)
        return ("segs", segs)

    def run(self, handle):
        kind, items = handle
        if kind == "ops":
            for op in items:
                self.sv.apply_operation(op)
            return len(items)
        for seg in items:
            self.sv.run_segment(seg)
        return len(items)

    def units(self, handle):
        """The separately launchable parts of a compiled run step: ``[(unit, busy)]`` with
        ``busy`` = mask of the local index bits the unit acts on when it can also run on a
        SUBSET of the state (partial launches of the specialised segment kernels: any index bit
        outside ``busy`` may be held fixed), else ``None``.  ``None`` for the whole handle: not
        divisible (per-gate mode, broadcast states)."""
        kind, items = handle
        if kind != "segs" or self.sv.batch != 1:
            return None
        out = []
        for seg in items:
            busy = None
            if self.sv.segment_partial_ok(seg):
                busy = 0
                for b in seg.tile_bits:
                    busy |= 1 << int(b)
            out.append((seg, busy))
        return out

    def run_unit(self, seg, fix_mask: int = 0, fix_val: int = 0):
        self.sv.run_segment(seg, 0, fix_mask, fix_val)
        return 1

    def remap_copy(self, dst_ptr, dst_pitch, src_ptr, src_pitch, run_bytes, count, stream):
        """``b200q_remap_copy``: the pack / unpack copies of an exchange piece, on the copy engines."""
        from ._lib import check

        check(self.sv.lib.b200q_remap_copy(C.c_void_p(dst_ptr), dst_pitch, C.c_void_p(src_ptr), src_pitch,
                                           run_bytes, count, C.c_void_p(stream)))

    def remap_unpack(self, dst_ptr, dst_pitch, src_ptr, src_pitch, run_bytes, count, stream, mode=0):
        """``b200q_remap_unpack``: staging buffer -> state as a kernel that shares the SMs with
        the segment kernels (mode 0: TMA bulk copies, 1: through registers)."""
        from ._lib import check

        check(self.sv.lib.b200q_remap_unpack(C.c_void_p(dst_ptr), dst_pitch, C.c_void_p(src_ptr), src_pitch,
                                             run_bytes, count, int(mode), 0, C.c_void_p(stream)))

    def unit_rounds(self, seg) -> int:
        """Rounds (= 1 + shared-memory transpositions per tile) of a compiled unit."""
        plan = getattr(seg, "_sk_plan", None)
        return len(plan.rounds) if plan is not None else 0

    def stream_write32(self, addr, value, stream):
        from ._lib import check

        check(self.sv.lib.b200q_stream_write32(C.c_void_p(addr), int(value) & 0xFFFFFFFF, C.c_void_p(stream)))

    def stream_wait_geq32(self, addr, value, stream):
        from ._lib import check

        check(self.sv.lib.b200q_stream_wait_geq32(C.c_void_p(addr), int(value) & 0xFFFFFFFF, C.c_void_p(stream)))

    # -- reductions ------------------------------------------------------------------------
    def expval_terms(self, xs, zs, ys, cs):
        """sum_t cs[t] <psi_local| P_t |psi_local> per batch element -> (B,) float64."""
        from ._lib import check, f64_array, int_array, u64_array

        sv = self.sv
        w, wb = sv.workspace()
        check(sv.lib.b200q_expval_pauli_sum(
            sv.ptr, sv.n, sv.dtype_code, sv.batch, u64_array(xs), u64_array(zs),
            int_array(ys) if ys else None, f64_array(cs), len(cs),
            C.c_void_p(sv._scal.data_ptr()), w, wb, sv.stream))
        return sv._scal[: sv.batch].cpu().numpy().copy()

    def probs(self, local_wires):
        """(B, 2**m) float64 marginal over local wires (host array)."""
        return self.sv.probs_device(local_wires).cpu().numpy()

    def probs_device(self, local_wires):
        return self.sv.probs_device(local_wires)

    def reduced_dm(self, local_wires):
        return self.sv.reduced_dm(local_wires)

    def probs_inplace_device(self, chunk_bits: int = 26):
        """|psi|^2 over all local wires written over the state itself (the state is consumed):
        at 33 local qubits the complex128 shard is 128 GiB and a separate 64 GiB probability
        vector does not fit next to it.  Chunk c of 2^chunk_bits amplitudes is reduced into a
        staging buffer by the ordinary probs kernel and then copied to doubles [c*K, (c+1)*K)
        of the state buffer — bytes whose amplitudes were consumed by this or an earlier chunk
        (8(c+1)K <= 16cK for c >= 1; chunk 0 is already in the staging buffer)."""
        import torch
        from ._lib import check, int_array

        sv = self.sv
        if sv.batch != 1:
            raise NotImplementedError("in-place probabilities of a broadcast state")
        kb = min(int(chunk_bits), sv.n)
        K = 1 << kb
        stage = torch.empty(K, dtype=torch.float64, device=sv.device)
        front = sv.data.view(torch.float64).reshape(-1)[: 1 << sv.n]
        elem = 16 if sv.dtype_code else 8
        bits = int_array(list(range(kb - 1, -1, -1)))
        w, wb = sv.workspace()
        for c in range(1 << (sv.n - kb)):
            check(sv.lib.b200q_probs(C.c_void_p(sv.data.data_ptr() + c * K * elem), kb, sv.dtype_code, 1,
                                     bits, kb, C.c_void_p(stage.data_ptr()), w, wb, sv.stream))
            front[c * K:(c + 1) * K].copy_(stage)
        return front

    # -- sampler building blocks (all on 2**m float64 device vectors) ---------------------------
    def _scalar(self, value):
        import torch

        return torch.tensor([float(value)], dtype=torch.float64, device=self.sv.device)

    def has_nan(self, p, m):
        import torch
        from ._lib import check

        flag = torch.zeros(1, dtype=torch.int32, device=self.sv.device)
        check(self.sv.lib.b200q_has_nan(C.c_void_p(p.data_ptr()), m, C.c_void_p(flag.data_ptr()),
                                        self.sv.stream))
        return bool(flag.item())

    def np_sum(self, p, m):
        from ._lib import check

        need = ((1 << m) // 128 + (1 << m) // (128 * 2047) + 128) * 8 + (4 << 20)
        w, wb = self.sv.workspace(need)
        out = self._scalar(0.0)
        check(self.sv.lib.b200q_np_sum(C.c_void_p(p.data_ptr()), m, C.c_void_p(out.data_ptr()), w, wb,
                                       self.sv.stream))
        return float(out.item())

    def div_by(self, p, m, value):
        from ._lib import check

        d = self._scalar(value)
        check(self.sv.lib.b200q_div_by(C.c_void_p(p.data_ptr()), m, C.c_void_p(d.data_ptr()),
                                       self.sv.stream))

    def cumsum(self, p, m, exact, carry):
        """In-place inclusive cumsum starting from ``carry`` (None: start of the vector);
        returns the last entry."""
        from ._lib import check

        need = ((1 << m) // 2048 + 128) * 48 + (4 << 20)      # exact mode: 48 bytes per 2048-term chunk
        w, wb = self.sv.workspace(need)
        c = None if carry is None else self._scalar(carry)
        check(self.sv.lib.b200q_cumsum(C.c_void_p(p.data_ptr()), m, 0 if exact else 1,
                                       None if c is None else C.c_void_p(c.data_ptr()), w, wb,
                                       self.sv.stream))
        return float(p[-1].item())

    def search(self, cdf, m, u):
        """Number of cdf entries <= u, per uniform (device int64 tensor)."""
        import torch
        from ._lib import check

        ud = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(self.sv.device)
        out = torch.empty(len(u), dtype=torch.int64, device=self.sv.device)
        check(self.sv.lib.b200q_search(C.c_void_p(cdf.data_ptr()), m, C.c_void_p(ud.data_ptr()),
                                       len(u), C.c_void_p(out.data_ptr()), self.sv.stream))
        return out

    def unpack_bits(self, idx, m):
        import torch
        from ._lib import check

        bits = torch.empty((idx.numel(), m), dtype=torch.int64, device=idx.device)
        check(self.sv.lib.b200q_unpack_bits(C.c_void_p(idx.data_ptr()), idx.numel(), m,
                                            C.c_void_p(bits.data_ptr()), self.sv.stream))
        return bits.cpu().numpy()

    def sample_replicated(self, probs_host, shots, rng, exact):
        """Single-GPU sampler on a (small) probability vector every rank holds."""
        import torch
        from ._lib import check

        sv = self.sv
        m = int(np.log2(probs_host.size))
        p = torch.from_numpy(np.ascontiguousarray(probs_host, dtype=np.float64)).to(sv.device)
        u = torch.from_numpy(rng.random(shots)).to(sv.device)
        bits = torch.empty((shots, m), dtype=torch.int64, device=sv.device)
        flags = torch.zeros(1, dtype=torch.int32, device=sv.device)
        need = ((1 << m) // 128 + (1 << m) // (128 * 2047) + 128) * 8 + (4 << 20)
        w, wb = sv.workspace(need)
        check(sv.lib.b200q_sample(C.c_void_p(p.data_ptr()), m, C.c_void_p(u.data_ptr()), shots,
                                  0 if exact else 1, None, C.c_void_p(bits.data_ptr()),
                                  C.c_void_p(sv._scal.data_ptr()), C.c_void_p(flags.data_ptr()),
                                  w, wb, sv.stream))
        norm = float(sv._scal[0].item())
        if int(flags.item()):
            return np.zeros((shots, m), dtype=np.int64)
        if abs(norm - 1.0) > 1e-6:
            raise ValueError("probabilities do not sum to 1")
        return bits.cpu().numpy()

    def synchronize(self):
        import torch

        torch.cuda.synchronize(self.sv.device)


# ---------------------------------------------------------------------------------------------
# the sharded statevector
#: process-wide symmetric staging buffers: (id(group), dtype) -> (buffer, handle, capacity) | False
_SYMM_CACHE: dict = {}


# ---------------------------------------------------------------------------------------------
class ShardedStateVector:
    """``2**n`` amplitudes over ``dist.get_world_size()`` ranks (a power of two)."""

    def __init__(self, num_wires, dist, engine=None, dtype=np.complex128, batch=1, device=None,
                 fusion=1, stage_bytes=None, group=None):
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        g = self.world.bit_length() - 1
        if (1 << g) != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.n, self.g, self.nl = int(num_wires), g, int(num_wires) - g
        if self.nl < 1:
            raise ValueError("fewer than one local qubit per rank")
        self.engine = engine if engine is not None else CudaEngine(self.nl, dtype, batch, device,
                                                                   fusion)
        # real dtype of the shards (the engine's buffer when it has one)
        data = getattr(self.engine, "data", None)
        if data is None:
            data = getattr(getattr(self.engine, "sv", None), "data", None)
        self.np_dtype = (np.dtype(np.complex64) if data is not None and "complex64" in str(data.dtype)
                         else np.dtype(dtype) if data is None else np.dtype(np.complex128))
        self.phys = list(range(self.n))
        # bytes moved per exchange step and staging buffer (three buffers are allocated)
        # (1 GiB; 512 MiB beside shards of 128 GiB and more, where HBM is nearly full)
        shard_bytes = (16 if self.np_dtype == np.dtype(np.complex128) else 8) << self.nl
        self.stage_bytes = int(stage_bytes if stage_bytes is not None
                               else os.environ.get("B200Q_STAGE_BYTES", (1 << 29) if shard_bytes >= (1 << 37) else (1 << 30)))
        self._stage = None
        self._symm = None          # (buffer, handle, capacity, state) of the symmetric staging buffers
        self.stats = {"exchanges": 0, "exchange_bytes": 0, "run_steps": 0, "sweeps": 0}
        self.timer = None          # optional: callable(kind, fn) -> fn() (bench.py times steps)
        self.comm_records = []     # (start, end) events of the exchange pieces while a timer is set
        self.trace = None          # list: per-piece CUDA events of windows (tools/trace_window.py)
        self.reset()

    # -- bookkeeping ---------------------------------------------------------------------------
    def bit_of(self, wire):
        return self.n - 1 - int(wire)

    @property
    def batch(self):
        return self.engine.batch

    def rank_bit(self, j, rank=None):
        return ((self.rank if rank is None else rank) >> j) & 1

    def _maps(self, phys, wires):
        """(gvals, wmap) of ``wires`` under ``phys`` for this rank."""
        gvals, wmap = {}, {}
        for w in wires:
            p = phys[self.bit_of(w)]
            if p >= self.nl:
                gvals[w] = self.rank_bit(p - self.nl)
            else:
                wmap[w] = self.nl - 1 - p
        return gvals, wmap

    def reset(self):
        """|0...0>."""
        self.phys = list(range(self.n))
        self.engine.reset(0 if self.rank == 0 else None)

    def set_state(self, full_state):
        """Every rank passes the same full host vector ((2**n,), (2,)*n or batched)."""
        self.phys = list(range(self.n))
        arr = np.asarray(full_state)
        dim = 1 << self.n
        flat = arr.reshape(-1, dim)
        ln = 1 << self.nl
        self.engine.set_local_state(flat[:, self.rank * ln:(self.rank + 1) * ln])

    # -- gates -------------------------------------------------------------------------------
    def localise(self, step: RunStep):
        """Run step -> operators on local wire labels for THIS rank."""
        out = []
        for op in step.ops:
            gvals, wmap = self._maps(step.phys, op.wires)
            loc = specialise(op, gvals, wmap, spare_wire=0)
            if loc is not None:
                out.append(loc)
        for pa, pb in step.swaps:
            out.append(_ops.SWAP(wires=[self.nl - 1 - pa, self.nl - 1 - pb]))
        return out

    def compile(self, ops_):
        """Plan + per-rank lowering.  Returns a program to pass to :meth:`run`; the current
        map is consumed (a program is valid for the map it was compiled from)."""
        steps, final = plan(ops_, self.n, self.g, self.phys, self.bit_of)
        prog = []
        for st in steps:
            if isinstance(st, RunStep):
                lops = self.localise(st)
                prog.append(("run", self.engine.compile(lops) if lops else None))
            else:
                prog.append(("exchange", st))
        return {"steps": prog, "schedule": self._schedule(prog), "start": list(self.phys), "final": final,
                "n_exchanges": sum(1 for k, _ in prog if k == "exchange")}

    # -- overlap of exchanges with the sweeps around them ----------------------------------------
    def _schedule(self, steps):
        """Program steps -> execution schedule.  An exchange moves whole slabs of the shard (the
        values of the top ``k`` local bits); cut along ``pb`` further index bits it becomes
        ``2**pb`` independent PIECES.  A fused segment whose tile does not contain those bits
        can run piece by piece too (a partial launch), so the segments just before and just
        after an exchange are pipelined against it: piece p is swept, its transfer starts on
        the copy engines / NVLink while piece p+1 is swept, and the first segment after the
        exchange starts on piece 0 as soon as that piece has arrived.  Entries:
          ("run", handle) | ("units", [unit]) | ("exchange", ex)
          ("window", [units before], ex, [units after], lo, pb)   # piece bits lo .. lo+pb-1
        Knobs: B200Q_EXCHANGE_PIECE_BITS (default 3, 0 = no overlap), B200Q_EXCHANGE_WINDOW
        (segments per side, default 2)."""
        pb_max = int(os.environ.get("B200Q_EXCHANGE_PIECE_BITS", "3"))
        wmax = int(os.environ.get("B200Q_EXCHANGE_WINDOW", "2"))
        min_bit = int(os.environ.get("B200Q_EXCHANGE_MIN_BIT", "5"))
        units_of = getattr(self.engine, "units", None)
        items = []                                   # ["run", handle] | ["unit", unit, busy] | ["exchange", ex]
        for kind, item in steps:
            if kind == "run":
                us = units_of(item) if (item is not None and units_of is not None and pb_max > 0 and wmax > 0) else None
                if us is None:
                    items.append(["run", item])
                else:
                    items.extend(["unit", u, busy] for u, busy in us)
            else:
                items.append(["exchange", item])
        claimed: dict = {}                           # item index -> exchange index
        windows: dict = {}
        overlap = units_of is not None and pb_max > 0 and wmax > 0      # the same on every rank
        for i, it in enumerate(items):
            if it[0] != "exchange" or not overlap:
                continue
            top = self.nl - it[1].k
            before, j = [], i - 1
            while j >= 0 and len(before) < wmax and items[j][0] == "unit" and items[j][2] is not None \
                    and j not in claimed:
                before.insert(0, j)
                j -= 1
            after, j = [], i + 1
            while j < len(items) and len(after) < wmax and items[j][0] == "unit" and items[j][2] is not None:
                after.append(j)
                j += 1
            # busy masks of the (da, db) candidate windows; the choice of the piece bits must be
            # the SAME on every rank (the pieces of an exchange have to match), while the segments
            # differ from rank to rank (gates controlled by a rank bit are skipped on half of the
            # ranks): OR the masks over the ranks, then every rank scores the same table
            tab = []
            for da in range(wmax + 1):
                for db in range(wmax + 1):
                    busy = 0
                    for x in before[max(0, len(before) - da):] + after[:db]:
                        busy |= items[x][2]
                    tab.append(busy)
            avail = self._gather_masks(tab + [len(before), len(after)])
            tab = [0] * len(tab)
            for row in avail:
                for t in range(len(tab)):
                    tab[t] |= row[t]
            max_a, max_b = max(r[-2] for r in avail), max(r[-1] for r in avail)
            best = None
            for da in range(min(wmax, max_a), -1, -1):
                for db in range(min(wmax, max_b), -1, -1):
                    if da + db == 0:
                        continue
                    got = _free_bit_window(tab[da * (wmax + 1) + db], top, pb_max, min_bit)
                    if got is None:
                        continue
                    lo, pb = got
                    # long runs first (2**lo amplitudes per copy row: the unpack copy of 128 KiB
                    # rows runs at less than half the rate of 16 MiB rows), then window size
                    score = (min(pb, 2), min(lo, 16), da + db, min(da, db), pb, lo)
                    if best is None or score > best[0]:
                        best = (score, da, db, lo, pb)
            if best is not None:
                _, da, db, lo, pb = best
                a_idx, b_idx = before[max(0, len(before) - da):], after[:db]
                windows[i] = (a_idx, b_idx, lo, pb)
                for x in a_idx + b_idx:
                    claimed[x] = i
        sched, pending = [], []

        def flush():
            if pending:
                sched.append(("units", list(pending)))
                pending.clear()

        for i, it in enumerate(items):
            if i in claimed:
                continue
            if it[0] == "unit":
                pending.append(it[1])
                continue
            flush()
            if it[0] == "run":
                sched.append(("run", it[1]))
            elif i in windows:
                a_idx, b_idx, lo, pb = windows[i]
                sched.append(("window", [items[x][1] for x in a_idx], it[1], [items[x][1] for x in b_idx], lo, pb))
            else:
                sched.append(("exchange", it[1]))
        flush()
        return sched

    def run(self, program):
        if program["start"] != self.phys:
            raise ValueError("program compiled for another qubit map")
        timed = self.timer if self.timer is not None else (lambda kind, fn: fn())
        for entry in program["schedule"]:
            kind = entry[0]
            if kind == "run":
                if entry[1] is not None:
                    self.stats["sweeps"] += timed("run", lambda: self.engine.run(entry[1]))
                self.stats["run_steps"] += 1
            elif kind == "units":
                self.stats["sweeps"] += timed("run", lambda: sum(self.engine.run_unit(u) for u in entry[1]))
            elif kind == "exchange":
                timed("exchange", lambda: self.exchange(entry[1]))
            else:
                _, a_units, ex, b_units, lo, pb = entry
                self.stats["sweeps"] += timed("window", lambda: self._run_window(a_units, ex, b_units, lo, pb))
                self.stats["windows"] = self.stats.get("windows", 0) + 1
        self.phys = list(program["final"])

    def _gather_masks(self, values):
        """All ranks' copies of a short list of non-negative integers (< 2**62): [[...] per rank]."""
        if self.world == 1:
            return [list(values)]
        import torch

        data = getattr(self.engine, "data", None)
        dev = data.device if data is not None else "cpu"
        mine = torch.tensor([int(v) for v in values], dtype=torch.int64, device=dev)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(out, mine, group=self.group)
        return [[int(v) for v in t.cpu().tolist()] for t in out]

    def _run_window(self, a_units, ex, b_units, lo, pb):
        """[segments] -> exchange -> [segments], piece by piece (see :meth:`_schedule`)."""
        P = 1 << pb
        mask = (P - 1) << lo
        ctx = self._exchange_begin(ex, lo, pb)
        # unpack kernel: TMA bulk copies unless a segment of the window has many shared-memory
        # transpositions per tile (they starve the bulk copies: csrc/remap.cu), then registers
        mode = os.environ.get("B200Q_UNPACK_MODE", "tma")      # "auto" / "regs": measured slower end to end
        rounds_of = getattr(self.engine, "unit_rounds", None)
        heavy = rounds_of is not None and any(rounds_of(u) >= 4 for u in list(a_units) + list(b_units))
        ctx["unpack_mode"] = 1 if (mode == "regs" or (mode == "auto" and heavy)) else 0
        tr = self.trace

        def mark(tag, p):
            if tr is not None and "compute" in ctx:
                import torch

                e = torch.cuda.Event(enable_timing=True)
                e.record(ctx["compute"])
                tr.append((tag, p, [e]))

        for p in range(P):
            mark("a0", p)
            for u in a_units:
                self.engine.run_unit(u, mask, p << lo)
            mark("a1", p)
            self._exchange_piece(ctx, p)
        for p in range(P):
            self._exchange_wait(ctx, p)
            mark("b0", p)
            for u in b_units:
                self.engine.run_unit(u, mask, p << lo)
            mark("b1", p)
        self._exchange_end(ctx)
        return len(a_units) + len(b_units)

    def apply_mid_measure(self, op, mid_measurements: dict, rng=None):
        """``apply_mid_measure`` (apply_operation.py:415-497) on the sharded state.  The marginal
        of the measured wire is the ordered all-reduce of the ranks' partial marginals
        (:meth:`probs`); every rank makes the reference's ``binomial`` draw on its own copy of the
        same Generator; projector, ``1 / ||P psi||`` and reset are ONE 2x2 operator handed to the
        planner — diagonal (no reset), so a global wire needs no communication; with a reset on a
        global wire the planner swaps it local like for any other non-diagonal gate."""
        from .statevector import StateVector

        if self.batch > 1:
            raise ValueError("MidMeasure cannot be applied to batched states.")
        if mid_measurements is None:
            raise AssertionError("mid_measurements dictionary is required for MidMeasure")
        wire = op.wires[0]
        p = self.probs([wire])
        # the state's own precision decides the renormalisation tolerance (finfo(state.dtype) in
        # apply_operation.py:451-457): a complex64 state drifts in norm by ~1e-7
        sample, scale = StateVector.mid_measure_draw(p, self.np_dtype, rng)
        mid_measurements[op] = sample
        if sample == 1 and getattr(op, "reset", False):
            mat = np.zeros((2, 2), dtype=complex)
            mat[0, 1] = scale
            self.apply_operations([_ops.QubitUnitary(mat, wires=[wire])])
        else:
            diag = np.zeros(2, dtype=complex)
            diag[sample] = scale
            self.apply_operations([_ops.DiagonalQubitUnitary(diag, wires=[wire])])
        return sample

    def apply_gates(self, gates, mid_measurements=None, rng=None):
        """The gate loop with mid-circuit measurements and conditionals (simulate.py:213-235,
        apply_operation.py:355-411): unitary stretches go to the planner as before."""
        run = []
        for op in gates:
            name = op.name
            if name == "MidMeasureMP":
                if run:
                    self.apply_operations(run)
                    run = []
                self.apply_mid_measure(op, mid_measurements, rng)
            elif name.startswith("Conditional") and hasattr(op, "meas_val"):
                if op.meas_val.concretize(mid_measurements):
                    run.append(op.base)
            else:
                run.append(op)
        if run:
            self.apply_operations(run)

    def apply_operations(self, ops_):
        bs = [getattr(o, "batch_size", None) for o in ops_]
        bs = [b for b in bs if b is not None]
        if bs and self.engine.batch == 1 and bs[0] != 1:
            self.engine.resize_batch(bs[0])
        self.run(self.compile(list(ops_)))

    # -- the exchange --------------------------------------------------------------------------
    def _symm_stage(self, numel, dtype, device):
        """Three staging buffers in SYMMETRIC memory (torch.distributed._symmetric_memory: every
        rank can address every other rank's copy over NVLink).  Allocated and rendezvous'ed ONCE
        per process and (group, dtype) and shared by every ShardedStateVector; the decision to use
        them is COLLECTIVE (a rank whose allocation failed would otherwise take the NCCL path
        while its partners wait in the device-side barrier).  None -> NCCL send/recv."""
        if self._symm is False:
            return None
        if self._symm is not None and self._symm[2] >= numel and self._symm[0].dtype == dtype:
            return self._symm
        key = (id(self.group), str(dtype))
        cached = _SYMM_CACHE.get(key)
        if cached is False:
            self._symm = False
            return None
        if cached is not None and cached[2] >= numel:
            self._symm = cached
            return cached
        import torch

        entry, err = None, ""
        try:
            import torch.distributed._symmetric_memory as symm

            if device.type != "cuda" or os.environ.get("B200Q_EXCHANGE", "symm") != "symm":
                raise RuntimeError("symmetric exchange disabled")
            # + 512 bytes of flags (uint32 landed[64], consumed[64]) behind the three buffers
            buf = symm.empty(3 * numel + 512 // torch.empty(0, dtype=dtype).element_size(), dtype=dtype,
                             device=device)
            buf[3 * numel:].zero_()
            ok = torch.ones(1, dtype=torch.int32, device=device)
        except Exception as e:                       # noqa: BLE001 - any failure -> NCCL path
            err = repr(e)
            buf = None
            ok = torch.zeros(1, dtype=torch.int32, device=device) if device.type == "cuda" else None
        if ok is None:
            _SYMM_CACHE[key] = self._symm = False
            return None
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            try:
                hdl = symm.rendezvous(buf, self.group if self.group is not None else self.dist.group.WORLD)
                entry = (buf, hdl, numel, {"t": 0, "flag_off": 3 * numel * buf.element_size()})
            except Exception as e:                   # noqa: BLE001
                err = repr(e)
        ok2 = torch.tensor([1 if entry is not None else 0], dtype=torch.int32, device=device)
        self.dist.all_reduce(ok2, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok2.item()) != 1:
            entry = None
        _SYMM_CACHE[key] = entry if entry is not None else False
        self._symm = _SYMM_CACHE[key]
        self._symm_error = err
        return entry

    def exchange(self, ex: ExchangeStep):
        """Swap rank bits ``ex.rank_bits`` with local physical bits ``nl-k .. nl-1`` (one piece:
        the whole slabs)."""
        ctx = self._exchange_begin(ex, self.nl - ex.k, 0)
        self._exchange_piece(ctx, 0)
        self._exchange_wait(ctx, 0)
        self._exchange_end(ctx)

    def _exchange_begin(self, ex: ExchangeStep, lo: int, pb: int):
        """Geometry of an exchange cut into ``2**pb`` pieces along index bits ``lo .. lo+pb-1``.
        The data a rank gives to the partner that holds value ``j`` of the exchanged bits is slab
        ``j`` of its shard (``chunk`` contiguous amplitudes) and what it receives lands in the same
        slab; piece ``p`` of a slab is ``chunk >> (lo+pb)`` runs of ``2**lo`` amplitudes."""
        k = ex.k
        data = self.engine.data                      # (B, 2**nl)
        chunk = 1 << (self.nl - k)
        q = 0
        for i, rb in enumerate(ex.rank_bits):
            q |= self.rank_bit(rb) << i
        partners = []
        for j in range(1 << k):
            if j == q:
                continue
            r = self.rank
            for i, rb in enumerate(ex.rank_bits):
                r = (r & ~(1 << rb)) | (((j >> i) & 1) << rb)
            partners.append((j, r))
        itemsize = data.element_size()
        per_partner = max(1, self.stage_bytes // (itemsize * len(partners)))
        cap = min(chunk >> pb, 1 << (per_partner.bit_length() - 1))      # amplitudes per partner and step
        ctx = {"ex": ex, "data": data, "chunk": chunk, "q": q, "partners": partners, "lo": lo, "pb": pb,
               "cap": cap, "itemsize": itemsize, "events": {}, "symm": None}
        if data.is_cuda and self._symm_stage(max(cap * len(partners), self.stage_bytes // itemsize), data.dtype,
                                             data.device) is not None:
            import torch

            ctx["symm"] = self._symm
            if getattr(self, "_comm_stream", None) is None:
                # high priority: the barrier kernels and copies of an exchange step must not queue
                # behind the next (multi-millisecond, every-SM) segment launch of the compute stream
                self._comm_stream = torch.cuda.Stream(device=data.device, priority=-1)    # NVLink pushes + barriers
                self._unpack_stream = torch.cuda.Stream(device=data.device, priority=-1)  # staging -> state
            ctx["compute"] = torch.cuda.current_stream(data.device)
        elif self._stage is None or self._stage.numel() < cap * len(partners) or self._stage.dtype != data.dtype:
            import torch

            self._stage = torch.empty(cap * len(partners), dtype=data.dtype, device=data.device)
        return ctx

    @staticmethod
    def _piece_steps(chunk, lo, pb, p, cap):
        """Piece ``p`` of one slab in steps of at most ``cap`` amplitudes:
        (offset in the slab, run, pitch, count) — ``count`` runs of ``run`` amplitudes ``pitch`` apart."""
        run, pitch = 1 << lo, 1 << (lo + pb)
        count = max(1, chunk // pitch)
        first = p << lo
        if run > cap:
            for o in range(count):
                for sub in range(0, run, cap):
                    yield first + o * pitch + sub, cap, cap, 1
        else:
            rows = max(1, cap // run)
            for o in range(0, count, rows):
                yield first + o * pitch, run, pitch, min(rows, count - o)

    def _exchange_piece(self, ctx, p):
        """Start the transfer of piece ``p`` (every slab).  Symmetric-memory path (K9 of SURVEY.md
        section 2c): asynchronous, on two communication streams behind everything issued so far
        on the compute stream, and made of copy-engine copies and stream memory operations ONLY —
        no kernel: a segment launch keeps every SM busy for milliseconds (at 34 qubits a piece
        of a segment runs 6.5 ms) and a kernel of the communication streams, the device-side
        barrier of torch's symmetric memory included, was measured to wait for the next launch
        boundary (tools/trace_window.py, profiles/r2_exchange_trace.txt).  Per step (at most
        ``cap`` amplitudes per partner)
          NVLink stream : [wait until the partner has consumed step t-3]  PUSH my slab piece
                          straight into the partner's staging buffer (copy-engine copy with a
                          peer destination: 690-740 GB/s beside the sweeps, where pulls drop to
                          440 because remote reads wait on the partner's saturated HBM), then
                          write ``landed[me] = t+1`` into the partner's flags
                          (cuStreamWriteValue32 over NVLink);
          unpack stream : wait for ``landed[partner] >= t+1`` in MY flags (cuStreamWaitValue32),
                          copy what the partners left in my staging buffer into place, then
                          write ``consumed[me] = t+1`` into every rank's flags.
        ``t`` counts the steps of all exchanges of the process (the same sequence on every
        rank); three staging buffers rotate, so the push of step t+1 overlaps the unpack of
        step t.  Other backends: blocking send / recv."""
        data, chunk, partners, q = ctx["data"], ctx["chunk"], ctx["partners"], ctx["q"]
        lo, pb, cap, isz = ctx["lo"], ctx["pb"], ctx["cap"], ctx["itemsize"]
        B = data.shape[0]
        nelem = B * (chunk >> pb) * len(partners)
        self.stats["exchange_bytes"] += nelem * isz
        if ctx["symm"] is None:
            self._exchange_piece_p2p(ctx, p)
            return
        import torch

        buf, hdl, third, st = ctx["symm"]
        link, unpack = self._comm_stream, self._unpack_stream
        ev = torch.cuda.Event()
        ev.record(ctx["compute"])
        link.wait_event(ev)
        unpack.wait_event(ev)
        timing = self.timer is not None or self.trace is not None
        marks = []

        def mark(stream):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            marks.append(e)

        if timing:
            mark(link)
        eng = self.engine
        base, sbase = data.data_ptr(), buf.data_ptr()
        peers = hdl.buffer_ptrs
        row_stride = data.stride(0) * isz
        ls, us = link.cuda_stream, unpack.cuda_stream
        me, foff = self.rank, st["flag_off"]
        landed = lambda r, src: peers[r] + foff + 4 * src              # noqa: E731  landed[src] on rank r
        consumed = lambda r, src: peers[r] + foff + 256 + 4 * src      # noqa: E731
        done = None
        for b in range(B):
            for off, run, pitch, count in self._piece_steps(chunk, lo, pb, p, cap):
                t = st["t"]
                st["t"] += 1
                slot = (t % 3) * third
                for s_, (j, r) in enumerate(partners):
                    if t >= 3:
                        eng.stream_wait_geq32(consumed(me, r), t - 2, ls)
                    # partner r holds value j of the exchanged bits; in ITS partner list (all
                    # values but j, ascending) my value q sits at index q - (q > j)
                    s_there = q - (1 if q > j else 0)
                    eng.remap_copy(peers[r] + (slot + s_there * cap) * isz, run * isz,
                                   base + b * row_stride + (j * chunk + off) * isz, pitch * isz,
                                   run * isz, count, ls)
                    eng.stream_write32(landed(r, me), t + 1, ls)
                # my unpack overwrites the slab piece my own push reads: it also waits for that push
                pushed = torch.cuda.Event(enable_timing=self.trace is not None)
                pushed.record(link)
                if self.trace is not None:
                    marks.append(pushed)
                unpack.wait_event(pushed)
                for s_, (j, r) in enumerate(partners):
                    eng.stream_wait_geq32(landed(me, r), t + 1, us)
                if self.trace is not None:
                    mark(unpack)
                for s_, (j, r) in enumerate(partners):
                    eng.remap_unpack(base + b * row_stride + (j * chunk + off) * isz, pitch * isz,
                                     sbase + (slot + s_ * cap) * isz, run * isz, run * isz, count, us,
                                     ctx.get("unpack_mode", 0))
                for r in range(self.world):
                    if r != me:
                        eng.stream_write32(consumed(r, me), t + 1, us)
                if self.trace is not None:
                    mark(unpack)
        done = torch.cuda.Event(enable_timing=timing)
        done.record(unpack)
        if timing:
            mark(link)
            self.comm_records.append((marks[0], done))
        if self.trace is not None:
            self.trace.append(("comm", p, marks + [done]))
        ctx["events"][p] = done

    def _exchange_piece_p2p(self, ctx, p):
        import torch

        dist = self.dist
        data, chunk, partners = ctx["data"], ctx["chunk"], ctx["partners"]
        lo, pb, cap = ctx["lo"], ctx["pb"], ctx["cap"]
        for b in range(data.shape[0]):
            row = data[b]
            for off, run, pitch, count in self._piece_steps(chunk, lo, pb, p, cap):
                views, sends, p2p = [], [], []
                for s_, (j, r) in enumerate(partners):
                    v = torch.as_strided(row, (count, run), (pitch, 1), row.storage_offset() + j * chunk + off)
                    snd = v.contiguous()
                    stage = self._stage[s_ * cap: s_ * cap + count * run]
                    gr = r if self.group is None else dist.get_global_rank(self.group, r)
                    p2p.append(dist.P2POp(dist.isend, snd, gr, self.group))
                    p2p.append(dist.P2POp(dist.irecv, stage, gr, self.group))
                    views.append((v, stage))
                    sends.append(snd)
                for req in dist.batch_isend_irecv(p2p):
                    req.wait()
                for v, stage in views:
                    v.copy_(stage.view(count, run))

    def _exchange_wait(self, ctx, p):
        """The compute stream waits for piece ``p`` to be in place."""
        ev = ctx["events"].pop(p, None)
        if ev is not None:
            ctx["compute"].wait_event(ev)

    def _exchange_end(self, ctx):
        for p in list(ctx["events"]):
            self._exchange_wait(ctx, p)
        self.stats["exchanges"] += 1

    def remap(self, want_local_bits, nxt=None):
        """Make the given logical bits local (one exchange at most)."""
        if all(self.phys[b] < self.nl for b in want_local_bits):
            return
        if nxt is None:
            nxt = [2 * self.n - b for b in range(self.n)]
            for b in want_local_bits:
                nxt[b] = 0
        swaps, ex, new_phys = plan_remap(self.phys, self.nl, list(want_local_bits), nxt)
        self._run_swaps(swaps)
        self.exchange(ex)
        self.phys = new_phys

    def _run_swaps(self, swaps):
        if not swaps:
            return
        lops = [_ops.SWAP(wires=[self.nl - 1 - a, self.nl - 1 - b]) for a, b in swaps]
        self.stats["sweeps"] += self.engine.run(self.engine.compile(lops))

    def restore_identity_map(self):
        """Bring every logical bit back to its own physical position (needed before sampling:
        the CDF must run over logical indices).  At most two exchanges + one local permutation."""
        n, nl = self.n, self.nl
        for _ in range(3):
            at = {self.phys[b]: b for b in range(n)}
            wrong = [p for p in range(nl, n) if at[p] != p]
            if not wrong:
                break
            if any(self.phys[p] >= nl for p in wrong):
                # a rightful owner sits on another rank position: park every misplaced
                # occupant on local bits first, the next pass brings the owners in
                self._exchange_exact([(at[p], None) for p in wrong])
            else:
                self._exchange_exact([(at[p], p) for p in wrong])
        # local permutation by transpositions
        swaps = []
        phys = list(self.phys)
        at = {phys[b]: b for b in range(n)}
        for p in range(nl):
            if at[p] != p:
                src = phys[p]                       # where logical p currently lives
                other = at[p]
                swaps.append((src, p))
                phys[p], phys[other] = p, src
                at[p], at[src] = p, other
        self._run_swaps(swaps)
        self.phys = phys
        assert self.phys == list(range(n)), self.phys

    def _exchange_exact(self, pairs):
        """Exchange the rank-resident logical bits ``b`` of ``pairs`` = [(b, want)] with local
        bits: ``want`` = logical bit that must take b's rank position (None: any victim)."""
        n, nl = self.n, self.nl
        k = len(pairs)
        phys = list(self.phys)
        at = {phys[b]: b for b in range(n)}
        pairs = sorted(pairs, key=lambda t: phys[t[0]])
        fixed = {w for _, w in pairs if w is not None}
        spare = [b for b in range(n) if phys[b] < nl and b not in fixed and b < nl]
        spare.sort(key=lambda b: -phys[b])
        victims = [w if w is not None else spare.pop(0) for _, w in pairs]
        # victims[i] must sit at local position nl-k+i
        swaps = []
        for i, v in enumerate(victims):
            t = nl - k + i
            if phys[v] != t:
                other = at[t]
                swaps.append((phys[v], t))
                pv = phys[v]
                phys[v], phys[other] = t, pv
                at[t], at[pv] = v, other
        self._run_swaps(swaps)
        ex = ExchangeStep([phys[b] - nl for b, _ in pairs], k)
        self.exchange(ex)
        for i, (b, _) in enumerate(pairs):
            v = victims[i]
            phys[v], phys[b] = phys[b], nl - k + i
        self.phys = phys

    # -- reductions ----------------------------------------------------------------------------
    def _ordered_sum(self, local):
        """Sum of per-rank float64 arrays in rank order (deterministic)."""
        import torch

        t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64).reshape(-1))
        dev = self.engine.data.device
        t = t.to(dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        acc = out[0].cpu().numpy().copy()
        for o in out[1:]:
            acc = acc + o.cpu().numpy()
        return acc.reshape(np.shape(local))

    def expval_pauli_sentence(self, ps):
        """<psi| sum_t c_t P_t |psi> (float, or (B,) when batched)."""
        terms = []
        for word, coeff in ps.items():
            xb = [self.bit_of(w) for w, ch in word.items() if ch in "XY"]
            zb = [self.bit_of(w) for w, ch in word.items() if ch in "ZY"]
            ny = sum(1 for ch in word.values() if ch == "Y")
            terms.append((xb, zb, ny, float(np.real(coeff))))
        total = np.zeros(self.batch)
        pending = terms
        guard = 0
        while pending:
            now = [t for t in pending if all(self.phys[b] < self.nl for b in t[0])]
            later = [t for t in pending if not all(self.phys[b] < self.nl for b in t[0])]
            if now:
                xs, zs, ys, cs = [], [], [], []
                for xb, zb, ny, c in now:
                    xm = zm = 0
                    sign = 1.0
                    for b in xb:
                        xm |= 1 << self.phys[b]
                    for b in zb:
                        p = self.phys[b]
                        if p < self.nl:
                            zm |= 1 << p
                        elif self.rank_bit(p - self.nl):
                            sign = -sign
                    xs.append(xm); zs.append(zm); ys.append(ny); cs.append(sign * c)
                total = total + self._ordered_sum(self.engine.expval_terms(xs, zs, ys, cs))
            if later:
                guard += 1
                if guard > 4 * self.n:
                    raise RuntimeError("expval remapping made no progress")   # pragma: no cover
                first = later[0]
                flips = {}
                for xb, _, _, _ in later:
                    for b in xb:
                        flips[b] = flips.get(b, 0) + 1
                # victims: local bits that few pending terms flip, never one the first pending
                # term flips (so that term is measured next pass: guaranteed progress); other
                # rank-resident flip bits come along when a cheaper victim exists
                nxt = [-flips.get(b, 0) for b in range(self.n)]
                for b in first[0]:
                    nxt[b] = -10**6
                self.remap([b for b in first[0] if self.phys[b] >= self.nl], nxt)
            pending = later
        return total if self.batch > 1 else float(total[0])

    def probs(self, wires=None):
        """Marginal probabilities over ``wires`` in the given order (every rank gets the full
        2**m vector; m is assumed small — use :meth:`sample` for all-wire sampling)."""
        wires = list(range(self.n)) if wires is None else list(wires)
        m = len(wires)
        loc = [(i, w) for i, w in enumerate(wires) if self.phys[self.bit_of(w)] < self.nl]
        glo = [(i, w) for i, w in enumerate(wires) if self.phys[self.bit_of(w)] >= self.nl]
        lw = [self.nl - 1 - self.phys[self.bit_of(w)] for _, w in loc]
        pl = self.engine.probs(lw) if lw else None
        if pl is None:
            pl = self.engine.probs([0]).sum(axis=-1, keepdims=True)
        B = pl.shape[0]
        full = np.zeros((B,) + (2,) * m)
        idx = [slice(None)] * (m + 1)
        for i, w in glo:
            idx[i + 1] = self.rank_bit(self.phys[self.bit_of(w)] - self.nl)
        sub = pl.reshape((B,) + (2,) * len(loc))
        # axes of `sub` follow `loc` order, which is the order of the remaining axes of `full`
        full[tuple(idx)] = sub
        out = self._ordered_sum(full.reshape(B, -1))
        return out if B > 1 else out[0]

    def reduced_dm(self, wires):
        """Reduced density matrix over ``wires`` (``reduce_statevector``, math/quantum.py:386-487)
        of the sharded state: the kept wires are made local (one exchange at most), every rank
        forms the Gram blocks of its shard (``b200q_gram_block`` — a rank's bits are part of the
        traced index) and the 4^m partial matrices are added in rank order."""
        wires = list(wires)
        if len(wires) > self.nl:
            raise ValueError(f"reduced_dm over {len(wires)} wires needs them local; shards hold "
                             f"{self.nl} qubits")
        self.remap([self.bit_of(w) for w in wires])
        lw = [self.nl - 1 - self.phys[self.bit_of(w)] for w in wires]
        part = np.asarray(self.engine.reduced_dm(lw), dtype=np.complex128)
        tot = self._ordered_sum(np.stack([part.real, part.imag]))
        return tot[0] + 1j * tot[1]

    def norm2(self):
        xs, zs, ys, cs = [0], [0], [0], [1.0]
        r = self._ordered_sum(self.engine.expval_terms(xs, zs, ys, cs))
        return r if self.batch > 1 else float(r[0])

    # -- sampling ------------------------------------------------------------------------------
    def sample(self, shots, rng, wires=None, exact=True, consume=None):
        """(shots, m) int64 samples, bit-identical on every rank and to sampling.py:500-531 under
        the same Generator state (every rank must hold an identically seeded ``rng``).
        ``consume``: build the probabilities over the state itself (the state is destroyed);
        default: only when a separate probability vector would not fit in free device memory."""
        import torch

        if self.batch != 1:
            raise NotImplementedError("sharded sampling of broadcast states")
        wires_all = list(range(self.n))
        if wires is not None and list(wires) != wires_all:
            p = self.probs(wires)
            return self.engine.sample_replicated(p, shots, rng, exact)
        self.restore_identity_map()
        eng, dist, nl = self.engine, self.dist, self.nl
        if consume is None and hasattr(eng, "probs_inplace_device") and torch.cuda.is_available():
            free, _ = torch.cuda.mem_get_info()
            consume = free < (8 << nl) + (2 << 30)
        if consume and hasattr(eng, "probs_inplace_device"):
            p = eng.probs_inplace_device()
        else:
            p = eng.probs_device(list(range(nl)))[0]
        u = rng.random(shots)
        nan = np.array([1.0 if eng.has_nan(p, nl) else 0.0])
        if self._ordered_sum(nan)[0] > 0:
            return np.zeros((shots, self.n), dtype=np.int64)       # sampling.py:322-325
        # norm = probs.sum(): numpy's pairwise tree = same tree over the per-rank sums
        sums = self._gather_scalar(eng.np_sum(p, nl))
        while len(sums) > 1:
            sums = [sums[i] + sums[i + 1] for i in range(0, len(sums), 2)]
        norm = float(sums[0])
        if abs(norm - 1.0) > 1e-6:                                  # sampling.py:514-519
            raise ValueError("probabilities do not sum to 1")
        eng.div_by(p, nl, norm)
        # cdf = probs.cumsum(): the running value travels from rank to rank
        dev = p.device
        carry = None
        if self.rank > 0:
            t = torch.zeros(1, dtype=torch.float64, device=dev)
            dist.recv(t, self._grank(self.rank - 1), group=self.group)
            carry = float(t.item())
        last = eng.cumsum(p, nl, bool(exact), carry)
        if self.rank + 1 < self.world:
            dist.send(torch.tensor([last], dtype=torch.float64, device=dev),
                      self._grank(self.rank + 1), group=self.group)
        total = self._gather_scalar(last)[-1]
        eng.div_by(p, nl, total)                                    # cdf /= cdf[-1]
        cnt = eng.search(p, nl, u)
        dist.all_reduce(cnt, group=self.group)                      # int64 sum: exact
        return eng.unpack_bits(cnt, self.n)

    def _grank(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def _gather_scalar(self, x):
        import torch

        dev = self.engine.data.device
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return [float(o.item()) for o in out]

    # -- host copy (tests / small states) ---------------------------------------------------------
    def to_numpy(self):
        """Full state in LOGICAL order on every rank ((B, 2**n) or (2**n,)).  Small states only."""
        import torch

        data = self.engine.data
        out = [torch.empty_like(data) for _ in range(self.world)]
        self.dist.all_gather(out, data.contiguous(), group=self.group)
        B = data.shape[0]
        phys_state = torch.stack(out, dim=1).cpu().numpy().reshape(B, 1 << self.n)
        # axis order of the physical index, MSB first: physical bit n-1 .. 0
        arr = phys_state.reshape((B,) + (2,) * self.n)
        # logical bit b lives on physical bit phys[b] -> numpy axis (n-1-phys[b]) + 1
        axes = [0] + [1 + (self.n - 1 - self.phys[b]) for b in range(self.n - 1, -1, -1)]
        logical = arr.transpose(axes).reshape(B, 1 << self.n)
        return logical if B > 1 else logical[0]


# ---------------------------------------------------------------------------------------------
# circuit level
# ---------------------------------------------------------------------------------------------
def _simulate_sharded_one_shot(circuit, dist, rng, dtype, engine, fusion, exact_sampling, device):
    """The one-shot loop of simulate.py:354-381 on the sharded state: per shot the tape is run
    from |0..0> (as the reference does), mid-circuit values are drawn identically on every rank,
    the terminal all-wire sample is one shot of the distributed sampler, and the sampled MCM
    values close the result tuple (sampling.py:235-267)."""
    if not circuit.shots:
        raise TypeError("Native mid-circuit measurements are only supported with finite shots.")
    rng = np.random.default_rng(rng)
    n = circuit.num_wires
    ops_ = list(circuit.operations)
    wires = list(range(n))
    sv = ShardedStateVector(n, dist, engine=engine, dtype=dtype, device=device, fusion=fusion)
    prep = ops_[0] if ops_ and hasattr(ops_[0], "state_vector") else None
    results = []
    for _ in range(circuit.shots.total_shots):
        if prep is not None:
            sv.set_state(np.asarray(prep.state_vector(wire_order=wires)))
        else:
            sv.reset()
        mm = {}
        sv.apply_gates(ops_[bool(prep):], mm, rng)
        mps = list(circuit.measurements)[: len(circuit.measurements) - len(mm)]
        if any(mp.obs is not None for mp in mps):
            raise NotImplementedError("sharded finite-shot measurement of an observable")
        res = []
        if mps:
            samples = sv.sample(1, rng, None, exact_sampling)
            res = [mp.process_samples(samples, wires) for mp in mps]
        res += list(mm.values())
        results.append(res[0] if len(circuit.measurements) == 1 else tuple(res))
    return tuple(results)


def simulate_sharded(circuit, dist, rng=None, dtype=np.complex128, engine=None, fusion=1,
                     exact_sampling=True, device=None, return_state=False):
    """The sharded mirror of simulate.py:308-393 for a tape in standard wire order: gate loop,
    then expval (Pauli observables) / probs analytically, or sample() with shots."""
    circuit = circuit.map_to_standard_wires()
    n = circuit.num_wires
    ops_ = list(circuit.operations)
    if any(op.name == "MidMeasureMP" for op in ops_):
        return _simulate_sharded_one_shot(circuit, dist, rng, dtype, engine, fusion,
                                          exact_sampling, device)
    sv = ShardedStateVector(n, dist, engine=engine, dtype=dtype, device=device, fusion=fusion)
    if ops_ and hasattr(ops_[0], "state_vector"):
        sv.set_state(np.asarray(ops_[0].state_vector(wire_order=list(range(n)))))
        ops_ = ops_[1:]
    if ops_:
        sv.apply_operations(ops_)
    results = []
    if circuit.shots:
        rng = np.random.default_rng(rng)
        # measure_with_samples (sampling.py:276-335): one all-wire sampling, then per-measurement
        # post-processing of the sample matrix
        if any(mp.obs is not None for mp in circuit.measurements):
            raise NotImplementedError("sharded finite-shot measurement of an observable")
        wires = list(range(n))
        samples = sv.sample(circuit.shots.total_shots, rng, None, exact_sampling)
        for mp in circuit.measurements:
            results.append(mp.process_samples(samples, wires))
    else:
        for mp in circuit.measurements:
            if mp.kind == "expval" and getattr(mp.obs, "pauli_rep", None) is not None:
                r = sv.expval_pauli_sentence(mp.obs.pauli_rep)
                results.append(np.float64(r) if np.ndim(r) == 0 else r)
            elif mp.kind == "probs" and mp.obs is None:
                results.append(sv.probs(list(mp.wires) if len(mp.wires) else None))
            elif mp.kind in ("density_matrix", "purity", "vn_entropy", "mutual_info"):
                from .simulate import _measure_density

                results.append(_measure_density(mp, sv, False))
            elif mp.kind == "state":
                results.append(sv.to_numpy())
            else:
                raise NotImplementedError(f"sharded measurement {mp.kind} of {mp.obs}")
    res = results[0] if len(results) == 1 else tuple(results)
    return (res, sv) if return_state else res
