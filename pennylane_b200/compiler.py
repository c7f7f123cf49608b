"""Host-side gate-fusion pass: operators -> primitives -> dense blocks -> tile segments.

north_star: "a host-side gate-fusion pass that packs runs of gates into dense k-qubit blocks".
The reference has no analogue: default.qubit sweeps the whole state once per gate
(pennylane/devices/qubit/simulate.py:214-235).  Here

  1. every operator is lowered to a *primitive* the tile kernel understands (dense 2x2 / 4x4
     with controls, controlled-X, swap, parity phase, small diagonal table); anything else
     (wide dense unitaries, broadcast parameters) stays a *generic* op for the per-gate kernels;
  2. runs of single-qubit gates on one wire are multiplied into one 2x2 block on the host, and
     (level 2) absorbed together with neighbouring two-qubit gates into 4x4 blocks;
  3. primitives are packed greedily — respecting the circuit's dependency order but hopping
     over gates on unrelated wires — into *segments*: sets of gates whose targets fit the T
     index bits of one shared-memory tile (L low bits + T-L free high bits);
  4. each segment is ONE launch of ``b200q_apply_tile``: one read and one write of the state.

Fusion changes floating-point rounding (products of matrices are formed on the host), so it is
a device option; the unfused path stays the bit-for-bit-stable default for parity tests.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

DENSE1, DENSE2, CX, PARITY, DIAG, SWAP = 0, 1, 2, 3, 4, 5
GENERIC = -1
GEN = 7          # adjoint generator term (RT_GEN record); defined here for merge_blocks


class TileOp(C.Structure):
    """Mirror of ``struct b200q_tile_op`` (include/b200q.h)."""
    _fields_ = [
        ("kind", C.c_int32), ("t0", C.c_int32), ("t1", C.c_int32), ("mat_off", C.c_int32),
        ("ctrl_mask_l", C.c_uint32), ("ctrl_val_l", C.c_uint32), ("par_mask_l", C.c_uint32),
        ("ndiag", C.c_int32),
        ("ctrl_mask_e", C.c_uint64), ("ctrl_val_e", C.c_uint64),
        ("par_mask_e", C.c_uint64), ("dbits", C.c_int8 * 8),
    ]


assert C.sizeof(TileOp) == 64


@dataclass
class Prim:
    kind: int
    targets: list = field(default_factory=list)      # global bit positions that must be in the tile
    ctrl: dict = field(default_factory=dict)         # bit -> required value (may be outside)
    other: list = field(default_factory=list)        # parity / diagonal-table bits (may be outside)
    mat: np.ndarray | None = None                    # dense matrix / (p0, p1) / table
    mat0: np.ndarray | None = None                   # DENSE1: matrix where the control FAILS
                                                     # (None: identity, i.e. a plain controlled gate)
    ngates: int = 1
    op: object = None                                # GENERIC: the original operator
    slot: int = 0                                    # GEN: output slot inside its segment
    param: int = -1                                  # GEN: trainable-parameter index
    ny: int = 0                                      # GEN: number of Y factors
    coef: float = 0.0                                # GEN: real coefficient of the Pauli term
    zbits: list = field(default_factory=list)        # GEN: bits carrying Z or Y
    seq: int = 0                                     # merge_blocks: stream position of the first
                                                     # constituent of an accumulated block
    pure_cx: bool = False                            # merge_blocks: block is exactly one CX so far
    deferred: list = field(default_factory=list)     # merge_blocks: (GEN, 2x2 Pauli) evaluated
                                                     # right after this block
    src: list | None = None                          # provenance for parameter rebinding: indices
                                                     # of the source operators, in order of
                                                     # application (None: unknown)
    src_j: int = 0                                   # position among the primitives of src[0]
    src_exact: bool = False                          # mat == product of the operators' matrices

    @property
    def bits(self):
        return set(self.targets) | set(self.ctrl) | set(self.other) | set(self.zbits)


_X = np.array([[0, 1], [1, 0]], dtype=complex)
_PHASES = {"PauliZ": -1.0 + 0j, "S": 1j, "T": np.exp(0.25j * np.pi)}


def _is_diag(m):
    return not np.any(m[~np.eye(m.shape[0], dtype=bool)])


def lower(op, bit_of, batched_ok: bool = False) -> list[Prim]:
    """Operator -> primitives (apply_operation.py:258-351 dispatch, re-done for the tile kernel).

    ``batched_ok``: single-qubit gates with broadcast parameters (``op.batch_size``,
    apply_operation.py:186-197) become dense blocks whose matrix carries a leading batch axis —
    the register kernel reads one matrix table per batch element; without it they stay generic
    per-gate launches."""
    name = op.name
    wires = list(op.wires)
    bits = [bit_of(w) for w in wires]
    if name in ("Identity", "Barrier", "Snapshot", "WireCut"):
        return []
    if hasattr(op, "state_vector"):
        return [Prim(GENERIC, targets=bits, op=op)]
    if getattr(op, "batch_size", None) is not None:
        if batched_ok and len(bits) == 1 and getattr(op, "has_matrix", True):
            m = np.asarray(op.matrix(), dtype=complex)
            if m.ndim == 3 and m.shape[1:] == (2, 2):
                return [Prim(DENSE1, targets=bits, mat=m)]
        return [Prim(GENERIC, targets=bits, op=op)]
    data = op.data
    if name == "GlobalPhase":
        ph = np.exp(-1j * float(data[0]))
        return [Prim(PARITY, mat=np.array([ph, ph]))]
    if name in _PHASES:
        return [Prim(PARITY, ctrl={bits[0]: 1}, mat=np.array([_PHASES[name]] * 2))]
    if name in ("PhaseShift", "U1", "ControlledPhaseShift"):
        ph = np.exp(1j * float(data[0]))
        return [Prim(PARITY, ctrl={b: 1 for b in bits}, mat=np.array([ph, ph]))]
    if name in ("CZ", "CCZ"):
        return [Prim(PARITY, ctrl={b: 1 for b in bits}, mat=np.array([-1.0 + 0j, -1.0 + 0j]))]
    if name in ("RZ", "IsingZZ", "MultiRZ"):
        th = float(data[0])
        return [Prim(PARITY, other=bits, mat=np.array([np.exp(-0.5j * th), np.exp(0.5j * th)]))]
    if name == "PauliRot":
        word = op.hyperparameters["pauli_word"]
        act = [(c, b) for c, b in zip(word, bits) if c != "I"]
        th = float(data[0])
        if not act:
            ph = np.exp(-0.5j * th)
            return [Prim(PARITY, mat=np.array([ph, ph]))]
        if all(c == "Z" for c, _ in act):
            return [Prim(PARITY, other=[b for _, b in act],
                         mat=np.array([np.exp(-0.5j * th), np.exp(0.5j * th)]))]
        if len(act) <= 2:
            from .ops import PauliRot as _PR
            sub = _PR.compute_matrix(th, pauli_word="".join(c for c, _ in act))
            return [_dense([b for _, b in act], {}, np.asarray(sub))]
        return [Prim(GENERIC, targets=bits, op=op)]
    if name in ("CNOT", "Toffoli"):
        return [Prim(CX, targets=[bits[-1]], ctrl={b: 1 for b in bits[:-1]})]
    if name == "MultiControlledX":
        cv = op.hyperparameters.get("control_values") or [True] * (len(bits) - 1)
        return [Prim(CX, targets=[bits[-1]], ctrl={b: int(bool(v)) for b, v in zip(bits[:-1], cv)})]
    if name == "PauliX":
        return [Prim(CX, targets=bits)]
    if name == "SWAP":
        return [Prim(SWAP, targets=bits)]
    if name == "CSWAP":
        return [Prim(SWAP, targets=bits[1:], ctrl={bits[0]: 1})]
    # controlled wrappers: base matrix on the base wires, controls as masks
    base = getattr(op, "base", None)
    cw = list(getattr(op, "control_wires", ()) or ())
    if cw and (name.startswith("C(") or name in ("ControlledQubitUnitary", "CRX", "CRY", "CRZ",
                                                   "CRot", "CY", "CH")):
        cvals = getattr(op, "control_values", None) or [True] * len(cw)
        tw = [w for w in wires if w not in cw]
        if base is not None and getattr(base, "has_matrix", True):
            m = np.asarray(base.matrix())
            tb = [bit_of(w) for w in base.wires]
        else:
            full = np.asarray(op.matrix())
            d = 1 << len(tw)
            m = full[-d:, -d:]
            tb = [bit_of(w) for w in tw]
        ctrl = {bit_of(w): int(bool(v)) for w, v in zip(cw, cvals)}
        if len(tb) <= 2 and m.ndim == 2:
            return [_dense(tb, ctrl, m)]
        return [Prim(GENERIC, targets=bits, op=op)]
    if not getattr(op, "has_matrix", True) or len(bits) > 2:
        if getattr(op, "has_matrix", True) and len(bits) <= 4:
            m = np.asarray(op.matrix())
            if m.ndim == 2 and _is_diag(m):
                return [Prim(DIAG, other=bits, mat=np.diag(m).copy())]
        return [Prim(GENERIC, targets=bits, op=op)]
    m = np.asarray(op.matrix())
    if m.ndim != 2:
        return [Prim(GENERIC, targets=bits, op=op)]
    return [_dense(bits, {}, m)]


def _dense(bits, ctrl, m) -> Prim:
    m = np.asarray(m, dtype=complex)
    if _is_diag(m) and not ctrl:
        if len(bits) == 1 and m[0, 0] == 1:
            return Prim(PARITY, ctrl={bits[0]: 1}, mat=np.array([m[1, 1], m[1, 1]]))
        return Prim(DIAG, other=list(bits), mat=np.diag(m).copy())
    if len(bits) == 1:
        if np.array_equal(m, _X):
            return Prim(CX, targets=list(bits), ctrl=dict(ctrl))
        return Prim(DENSE1, targets=list(bits), ctrl=dict(ctrl), mat=m)
    return Prim(DENSE2, targets=list(bits), ctrl=dict(ctrl), mat=m)


# ---------------------------------------------------------------------------------------------
# pass A: dense-block merging
# ---------------------------------------------------------------------------------------------
def _as_1q_matrix(p: Prim):
    """2x2 matrix of a primitive acting on exactly one bit with no other dependence, else None."""
    if p.kind == DENSE1 and not p.ctrl:
        return p.targets[0], p.mat
    if p.kind == CX and not p.ctrl:
        return p.targets[0], _X
    if p.kind == PARITY and not p.ctrl and len(p.other) == 1:
        return p.other[0], np.diag(p.mat)
    if p.kind == PARITY and len(p.ctrl) == 1 and not p.other and list(p.ctrl.values()) == [1]:
        return next(iter(p.ctrl)), np.diag([1.0, p.mat[0]])
    if p.kind == DIAG and len(p.other) == 1:
        return p.other[0], np.diag(p.mat)
    return None


def _embed(m1, pos):
    """2x2 on position pos (0 = MSB) of a two-bit block."""
    return np.kron(m1, np.eye(2)) if pos == 0 else np.kron(np.eye(2), m1)


_PAULI_2x2 = {
    "X": np.array([[0, 1], [1, 0]], dtype=complex),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "Z": np.array([[1, 0], [0, -1]], dtype=complex),
}


def _commute(a, b) -> bool:
    return bool(np.allclose(a @ b, b @ a, rtol=0, atol=1e-14))


def _gen_single_bit(g: Prim):
    """(bit, letter) when the generator term is a Pauli on exactly one bit, else None."""
    bits = g.bits
    if len(bits) != 1:
        return None
    (b,) = bits
    x, z = b in g.targets, b in g.zbits
    return b, ("Y" if x and z else "X" if x else "Z")


def _place_generator(g: Prim, pending: dict, out: list) -> bool:
    """Try to place a one-bit generator term without breaking up the pending block on its bit.

    ``<bra|G|ket>`` is unchanged when the same unitary acts on bra and ket and G is conjugated
    along, so the term may be evaluated (a) *before* a pending pure CNOT on its bit, conjugated
    through it (Z_t -> Z_c Z_t, Y_t -> Z_c Y_t, X_t -> X_t), or (b) *after* the pending block,
    as long as every gate merged into the block from now on commutes with it (RY with Y, any
    diagonal with Z ...).  With both moves the reverse sweep of [RY, RZ, CNOT] on a wire stays
    ONE controlled-select record with its two generator terms on either side, instead of three
    records separated by inner products.  Returns False when neither move is valid."""
    one = _gen_single_bit(g)
    if one is None:
        return False
    b, letter = one
    q = pending.get(b)
    readers = [t for t in pending if b in pending[t].ctrl]     # pending blocks controlled by b
    if q is None:
        if readers and letter != "Z":
            return False
        out.append(g)                 # Z on a control commutes with the controlled blocks
        return True
    if q.pure_cx and not q.deferred:
        (c, v), = q.ctrl.items()
        hoisted = Prim(GEN, targets=list(g.targets), zbits=list(g.zbits), ny=g.ny, coef=g.coef,
                       ngates=0, param=g.param, slot=g.slot)
        if letter in "ZY":
            hoisted.zbits = hoisted.zbits + [c]
            if not v:
                hoisted.coef = -hoisted.coef          # control on |0>: X_c CX X_c
        # the hoisted term is emitted ahead of every pending block; blocks that started before
        # the CNOT in the stream are overtaken and must commute with it
        for t, q2 in pending.items():
            if q2 is q or q2.seq > q.seq:
                continue
            for bit in hoisted.bits & q2.bits:
                if bit in q2.targets or bit in hoisted.targets:
                    return False                      # only Z on a control of q2 commutes
        out.append(hoisted)
        return True
    if readers:
        return False                                  # cannot happen (see merge_blocks), be safe
    q.deferred.append((g, _PAULI_2x2[letter]))
    return True


def merge_blocks(prims: list[Prim], level: int = 1, fold_cx: bool = True) -> list[Prim]:
    """``fold_cx=False`` (the specialised kernels of segjit.py): CNOTs stay separate records — a
    register renaming or a per-thread select there — and single-qubit runs stay uncontrolled, so
    that they can be applied in their normalised form.

    level 0: nothing; 1: products of single-qubit runs, with singly-controlled X gates on
    the same target folded in as *controlled-select* blocks (``mat`` where the control holds,
    ``mat0`` where it does not: CNOT next to a single-qubit block costs no extra pass and no
    data movement); 2: also absorb single-qubit blocks into adjacent dense two-qubit gates and
    merge consecutive dense gates on the same pair."""
    if level <= 0:
        return list(prims)
    out: list[Prim] = []
    pending: dict[int, Prim] = {}            # target bit -> accumulated block (not yet emitted)

    def emit(t):
        q = pending.pop(t)
        out.append(q)
        out.extend(g for g, _ in q.deferred)
        q.deferred = []

    def flush(bits):
        """Emit every pending block that shares a bit (target or control) with ``bits``."""
        bits = set(bits)
        for b in sorted(pending):
            if pending[b].bits & bits:
                emit(b)

    def flush_controlled_by(b):
        for t in sorted(pending):
            if b in pending[t].ctrl:
                emit(t)

    for seq, p in enumerate(prims):
        if p.kind == GEN:
            # Generator terms (adjoint reverse sweep) are inner products, not gates: they only
            # have to be evaluated at a point of the stream where the value is the same.
            if not p.bits:
                out.append(p)                     # identity term: Im<bra|ket> is invariant
                continue
            if _place_generator(p, pending, out):
                continue
            flush(p.bits)
            out.append(p)
            continue
        one = _as_1q_matrix(p) if p.kind != GENERIC else None
        if one is not None:
            b, m = one
            m = np.asarray(m, dtype=complex)
            flush_controlled_by(b)               # blocks that read b as a control come first
            if b in pending and any(not _commute(P, m) for _, P in pending[b].deferred):
                emit(b)                          # a deferred generator must stay before m
            if b in pending:
                q = pending[b]
                q.mat = m @ q.mat
                if q.mat0 is not None:
                    q.mat0 = m @ q.mat0
                q.ngates += p.ngates
                q.pure_cx = False
                q.src = (q.src + p.src) if (q.src is not None and p.src is not None) else None
                q.src_exact = q.src_exact and p.src_exact
            else:
                pending[b] = Prim(DENSE1, targets=[b], mat=m.copy(), ngates=p.ngates, seq=seq,
                                  src=None if p.src is None else list(p.src), src_exact=p.src_exact)
            continue
        if fold_cx and p.kind == CX and len(p.ctrl) == 1:
            t = p.targets[0]
            (c, v), = p.ctrl.items()
            flush_controlled_by(t)               # they read t before it is flipped
            if c in pending:
                emit(c)                          # the control's own block acts first
            q = pending.get(t)
            if q is not None and ((q.ctrl and q.ctrl != {c: v}) or q.deferred):
                emit(t)
                q = None
            if q is None:
                pending[t] = Prim(DENSE1, targets=[t], ctrl={c: v}, mat=_X.copy(),
                                  mat0=np.eye(2, dtype=complex), ngates=p.ngates, seq=seq,
                                  pure_cx=True)
            else:
                if not q.ctrl:
                    q.ctrl = {c: v}
                    q.mat0 = q.mat
                q.mat = _X @ q.mat
                q.ngates += p.ngates
                q.pure_cx = False
                q.src, q.src_exact = None, False
            continue
        if level >= 2 and p.kind == DENSE2 and not p.ctrl:
            # absorb pending (uncontrolled) single-qubit blocks that precede this gate
            for b in p.targets:
                flush_controlled_by(b)
            m = p.mat
            for pos, b in enumerate(p.targets):
                if b in pending and not pending[b].ctrl and not pending[b].deferred \
                        and np.ndim(pending[b].mat) == 2:
                    q = pending.pop(b)
                    m = m @ _embed(q.mat, pos)
                    p.ngates += q.ngates
            p.mat = m
            # merge with an immediately preceding dense gate on the same ordered pair
            if (out and out[-1].kind == DENSE2 and not out[-1].ctrl and out[-1].targets == p.targets
                    and not any(q.bits & p.bits for q in pending.values())):
                out[-1].mat = p.mat @ out[-1].mat
                out[-1].ngates += p.ngates
            else:
                flush(p.bits)
                out.append(p)
            continue
        flush(p.bits)
        if not p.bits:
            flush(list(pending))                 # global phases keep their place trivially
        out.append(p)
    flush([b for q in pending.values() for b in q.bits])
    # normalise accumulated blocks (a product may have become diagonal / X / identity)
    norm = []
    for p in out:
        if p.kind == DENSE1 and np.ndim(p.mat) == 3:
            norm.append(p)                       # broadcast block: one matrix per batch element
        elif p.kind == DENSE1 and not p.ctrl:
            q = _dense(p.targets, {}, p.mat)
            q.ngates = p.ngates
            if q.kind == DIAG and np.allclose(q.mat, 1.0):
                q = Prim(PARITY, mat=np.array([1.0 + 0j, 1.0 + 0j]), ngates=p.ngates)
            q.src, q.src_exact = p.src, p.src_exact and q.kind == DENSE1
            norm.append(q)
        elif p.kind == DENSE1 and p.mat0 is not None and np.array_equal(p.mat0, np.eye(2)):
            p.mat0 = None                        # plain controlled gate: skip where control fails
            if np.array_equal(p.mat, _X):
                p = Prim(CX, targets=p.targets, ctrl=p.ctrl, ngates=p.ngates)
            norm.append(p)
        else:
            norm.append(p)
    return norm


# ---------------------------------------------------------------------------------------------
# pass B: segment packing
# ---------------------------------------------------------------------------------------------
@dataclass
class Segment:
    tile_bits: list | None                    # None -> generic op
    prims: list
    ngates: int = 0
    rounds: list | None = None                # round schedule fixed by the packer (round budget)


def pack_segments(prims: list[Prim], n: int, T: int = 12, L: int = 5, max_ops: int = 96,
                  max_mat: int = 1024, round_budget: int | None = None, RB: int = 4,
                  sww: int = 3, trim_rounds: bool = False) -> list[Segment]:
    """``round_budget``: the primitives of a segment that do not fit that many rounds of the
    register kernel (``RB`` register bits per round) are handed back and packed later — an extra
    shared-memory transposition per tile costs about a quarter of a sweep, so single-qubit blocks
    that a later segment can take for free should not force one.

    ``trim_rounds``: the same hand-back, applied only where it is free.  Measured per-segment time
    at 30 qubits: ``max(5.75 ms, 2.0 + 1.0 (rounds - 1) + 0.2 blocks)`` — a segment of 4+ rounds is
    bound by its transpositions while its neighbours of 2 rounds idle at the memory time.  The
    greedy packer lets the segment that closes a layer's CNOT chain also take the NEXT layer's
    single-qubit blocks on its bits, which costs it a round or two; if dropping a few trailing
    uncontrolled single-qubit blocks saves a round, they are handed back (the lowest ``L`` bits are
    in every tile, the following two-round segments have a spare register slot for them)."""
    T = min(T, n)
    L = min(L, T)
    free = T - L
    remaining = list(prims)
    segments: list[Segment] = []
    while remaining:
        hi: set[int] = set()
        seg: list[Prim] = []
        keep: list[Prim] = []
        blocked: set[int] = set()
        nmat = 0
        for idx, p in enumerate(remaining):
            pb = p.bits
            if blocked & pb or len(seg) >= max_ops:
                blocked |= pb
                if not pb:
                    blocked |= set(range(n))    # global phase: keep order trivially
                keep.append(p)
                continue
            if p.kind == GENERIC:
                if not seg and not keep:
                    segments.append(Segment(None, [p], p.ngates))
                    keep.extend(remaining[idx + 1:])
                    seg = None
                    break
                blocked |= pb
                keep.append(p)
                continue
            if p.kind == DIAG and len(p.other) > 4:
                # wide diagonal table: not a tile op
                if not seg and not keep:
                    segments.append(Segment(None, [p], p.ngates))
                    keep.extend(remaining[idx + 1:])
                    seg = None
                    break
                blocked |= pb
                keep.append(p)
                continue
            need = {b for b in p.targets if b >= L and b not in hi}
            msize = (0 if p.mat is None else int(np.size(p.mat))) + \
                (0 if p.mat0 is None else int(np.size(p.mat0)))
            if len(hi) + len(need) <= free and nmat + msize <= max_mat:
                hi |= need
                nmat += msize
                seg.append(p)
            else:
                blocked |= pb
                keep.append(p)
        if seg is None:
            remaining = keep
            continue
        if seg:
            fill = [b for b in range(L, n) if b not in hi]
            bits = list(range(L)) + sorted(list(hi) + fill[: free - len(hi)])
            if round_budget is not None and len(bits) == T and len(bits) > RB \
                    and not any(p.kind == SWAP for p in seg):
                rounds, left = schedule_rounds_budget(seg, bits, RB, sww, round_budget)
                if left and len(left) < len(seg):
                    drop = {id(p) for p in left}
                    seg = [p for p in seg if id(p) not in drop]
                    order = {id(p): i for i, p in enumerate(remaining)}
                    keep = sorted(keep + left, key=lambda p: order[id(p)])
                    left = []
                if not left:
                    segments.append(Segment(bits, seg, sum(p.ngates for p in seg), rounds))
                    remaining = keep
                    continue
            elif trim_rounds and len(bits) == T and len(bits) > RB and not any(p.kind == SWAP for p in seg):
                full = schedule_rounds(seg, bits, RB, sww)
                best = None
                for budget in (len(full) - 1, len(full) - 2):
                    if len(full) < 4 or budget < 2:
                        break
                    rounds, left = schedule_rounds_budget(seg, bits, RB, sww, budget)
                    if left and len(left) < len(seg) and len(left) <= 6 and all(
                            q.kind == DENSE1 and not q.ctrl and q.mat0 is None for q in left):
                        best = (rounds, left)
                if best is not None:
                    rounds, left = best
                    drop = {id(p) for p in left}
                    seg = [p for p in seg if id(p) not in drop]
                    order = {id(p): i for i, p in enumerate(remaining)}
                    keep = sorted(keep + left, key=lambda p: order[id(p)])
                    segments.append(Segment(bits, seg, sum(p.ngates for p in seg), rounds))
                    remaining = keep
                    continue
            segments.append(Segment(bits, seg, sum(p.ngates for p in seg)))
        remaining = keep
    return segments


def lower_all(ops_, bit_of, batched_ok: bool = False) -> list[Prim]:
    """:func:`lower` over a list of operators, with provenance (``Prim.src``) for rebinding."""
    prims: list[Prim] = []
    for i, op in enumerate(ops_):
        low = lower(op, bit_of, batched_ok)
        exact = len(low) == 1 and len(op.wires) == 1 and getattr(op, "batch_size", None) is None \
            and low[0].kind != GENERIC
        for j, p in enumerate(low):
            p.src, p.src_j, p.src_exact = [i], j, exact
        prims.extend(low)
    return prims


def compile_ops(ops_, n: int, bit_of=None, level: int = 1, T: int = 12, L: int = 5,
                batched_ok: bool = False, fold_cx: bool = True, round_budget: int | None = None,
                RB: int = 4, sww: int = 3, trim_rounds: bool | None = None):
    """Operators -> list of :class:`Segment`.  ``trim_rounds`` (default: on for the specialised
    kernels, i.e. ``fold_cx`` off; B200Q_TRIM_ROUNDS=0 switches it off): see pack_segments."""
    if trim_rounds is None:
        trim_rounds = (not fold_cx) and os.environ.get("B200Q_TRIM_ROUNDS", "1") != "0"
    if bit_of is None:
        bit_of = lambda w: n - 1 - int(w)          # noqa: E731
    prims = lower_all(ops_, bit_of, batched_ok)
    prims = merge_blocks(prims, level, fold_cx)
    return pack_segments(prims, n, T=T, L=L, round_budget=round_budget, RB=RB, sww=sww,
                         trim_rounds=bool(trim_rounds))


def _expand_select(prims):
    """Controlled-select block -> two controlled records (the shared-memory kernel has no
    select): ``mat`` on control == v, ``mat0`` on control == not v."""
    out = []
    for p in prims:
        if p.kind == DENSE1 and p.mat0 is not None:
            (c, v), = p.ctrl.items()
            out.append(Prim(DENSE1, targets=p.targets, ctrl={c: v}, mat=p.mat, ngates=p.ngates))
            out.append(Prim(DENSE1, targets=p.targets, ctrl={c: 1 - v}, mat=p.mat0, ngates=0))
        else:
            out.append(p)
    return out


def encode_segment(seg: Segment):
    """Segment -> (ctypes TileOp array, complex128 matrix table) for ``b200q_apply_tile``."""
    pos = {b: i for i, b in enumerate(seg.tile_bits)}
    prims = _expand_select(seg.prims)
    ops_arr = (TileOp * len(prims))()
    mats: list[complex] = []
    for i, p in enumerate(prims):
        o = ops_arr[i]
        o.kind = p.kind
        cml = cvl = cme = cve = 0
        for b, v in p.ctrl.items():
            if b in pos:
                cml |= 1 << pos[b]
                cvl |= (1 << pos[b]) if v else 0
            else:
                cme |= 1 << b
                cve |= (1 << b) if v else 0
        o.ctrl_mask_l, o.ctrl_val_l, o.ctrl_mask_e, o.ctrl_val_e = cml, cvl, cme, cve
        o.mat_off = len(mats)
        if p.kind in (DENSE1, DENSE2):
            o.t0 = pos[p.targets[0]]
            o.t1 = pos[p.targets[1]] if p.kind == DENSE2 else 0
            mats.extend(np.asarray(p.mat, dtype=complex).reshape(-1))
        elif p.kind == CX:
            o.t0 = pos[p.targets[0]]
        elif p.kind == SWAP:
            o.t0, o.t1 = pos[p.targets[0]], pos[p.targets[1]]
        elif p.kind == PARITY:
            pml = pme = 0
            for b in p.other:
                if b in pos:
                    pml |= 1 << pos[b]
                else:
                    pme |= 1 << b
            o.par_mask_l, o.par_mask_e = pml, pme
            mats.extend(np.asarray(p.mat, dtype=complex).reshape(-1)[:2])
        elif p.kind == DIAG:
            o.ndiag = len(p.other)
            for j, b in enumerate(p.other):
                o.dbits[j] = pos[b] if b in pos else -(b + 1)
            mats.extend(np.asarray(p.mat, dtype=complex).reshape(-1))
        if len(mats) % 2:
            mats.append(0j)
    table = np.ascontiguousarray(np.array(mats if mats else [0j, 0j], dtype=np.complex128))
    return ops_arr, table


# ---------------------------------------------------------------------------------------------
# register-tiled kernel (rtile.cuh): round scheduling + record encoding
# ---------------------------------------------------------------------------------------------
RT_ROUND, RT_GEN = 6, 7
GEN = 7            # Prim kind: generator Pauli term of a trainable gate (adjoint sweeps)
import os as _os

# tile positions 0.._IO_LANES-1 stay on the lanes in the last (store) round: every warp store
# instruction then covers whole 2^_IO_LANES-amplitude runs (3 = single 128-byte lines of
# complex128 measured the same gates/s as 5 on B200; 5 keeps 512-byte runs)
_IO_LANES = int(_os.environ.get("B200Q_IO_LANES", 5))


class _RtGate(C.Structure):
    _fields_ = [("ctrl_r", C.c_uint32), ("cval_r", C.c_uint32), ("ctrl_t", C.c_uint32),
                ("cval_t", C.c_uint32), ("ctrl_e", C.c_uint64), ("cval_e", C.c_uint64),
                ("par_r", C.c_uint32), ("par_t", C.c_uint32), ("par_e", C.c_uint64)]


class _RtDiag(C.Structure):
    _fields_ = [("src", C.c_int8 * 16), ("pad", C.c_int32 * 8)]


class _RtRound(C.Structure):
    _fields_ = [("rbits", C.c_int8 * 8), ("tbits", C.c_int8 * 16), ("pad", C.c_int32 * 6)]


class _RtGen(C.Structure):
    _fields_ = [("xr", C.c_uint32), ("zr", C.c_uint32), ("zt", C.c_uint32), ("pad", C.c_uint32),
                ("ze", C.c_uint64), ("coef", C.c_double), ("pad2", C.c_int32 * 4)]


class _RtU(C.Union):
    _fields_ = [("g", _RtGate), ("d", _RtDiag), ("r", _RtRound), ("p", _RtGen)]


class RtOp(C.Structure):
    """Mirror of ``struct RtOp`` (pennylane_b200/csrc/rtile.cuh)."""
    _fields_ = [("kind", C.c_int32), ("q0", C.c_int32), ("q1", C.c_int32), ("mat_off", C.c_int32),
                ("u", _RtU)]


assert C.sizeof(RtOp) == 64


def _first_min_pos(sww):
    """Lowest tile position that may be a register bit in the FIRST round.  The first round reads
    the landing buffer in natural (unswizzled) order: a register bit below position ``sww`` makes
    that one read 2- to 8-way bank-conflicting (tuning knob B200Q_FIRST_MIN_POS; default: sww)."""
    v = _os.environ.get("B200Q_FIRST_MIN_POS")
    return sww if v is None else int(v)


@dataclass
class Round:
    rpos: list                 # tile positions held by register bits 0..RB-1
    tpos: list                 # tile positions held by thread bits 0..TB-1
    prims: list = field(default_factory=list)


def _expand_swaps(prims):
    """SWAP(a, b) = CX(a<-b) CX(b<-a) CX(a<-b): the register kernel has no swap record."""
    out = []
    for p in prims:
        if p.kind == SWAP:
            a, b = p.targets
            for t, c in ((a, b), (b, a), (a, b)):
                ctrl = dict(p.ctrl)
                ctrl[c] = 1
                out.append(Prim(CX, targets=[t], ctrl=ctrl, ngates=0))
            out[-1].ngates = p.ngates
        else:
            out.append(p)
    return out


def _thread_positions(free, sww, io):
    """Order the non-register tile positions over the thread bits.  IO rounds keep positions
    0..4 on the lanes (coalesced global access); inner rounds put positions that are distinct
    modulo the swizzle width on the lowest lane bits (conflict-free shared-memory phases)."""
    free = sorted(free)
    if io:
        return free
    chosen, seen = [], set()
    for p in free:
        if p % sww not in seen:
            chosen.append(p)
            seen.add(p % sww)
        if len(chosen) == sww:
            break
    rest = [p for p in free if p not in chosen]
    return chosen + rest


def schedule_rounds(prims, tile_bits, RB: int, sww: int = 3):
    """Assign the primitives of one segment to rounds.  Returns a list of :class:`Round`.

    A primitive can run in a round when all its targets are register bits of that round and no
    earlier unscheduled primitive shares a bit with it.  The first and last rounds are "IO
    rounds": their register bits avoid tile positions 0..4, which stay on the lanes.

    Two list schedulers are tried and the one with fewer rounds wins: program order (the first
    fit) and critical path first (the ready primitive with the longest chain of dependants picks
    the next register bit) — on a CNOT ring the former spends register bits on single-qubit
    blocks whose chain comes much later and revisits them (4 rounds for 12 targets), the latter
    follows the chain (3)."""
    a = _schedule_rounds_order(prims, tile_bits, RB, sww)
    if len(a) <= 2:
        return a
    b = _schedule_rounds_critical(prims, tile_bits, RB, sww)
    return b if len(b) < len(a) else a


def low_run(tile_bits) -> int:
    """Number of leading tile bits that are contiguous from bit 0 (the kernel's L)."""
    run = 0
    for i, b in enumerate(tile_bits):
        if b != i:
            break
        run += 1
    return run


def io_lanes(tile_bits, RB) -> int:
    """Tile positions that stay on the lanes in the last (store) round: the contiguous low run of
    the tile, at most _IO_LANES — a position above the run is not contiguous in memory, keeping it on
    the lanes buys no coalescing and costs the round its register slot (with L = 4 tiles it forced a
    third round on every 8-target segment)."""
    return min(_IO_LANES, low_run(tile_bits), len(tile_bits) - RB)


def _round_helpers(tile_bits, RB, sww):
    T = len(tile_bits)
    lanes = io_lanes(tile_bits, RB)
    io_allowed = set(range(lanes, T))

    def finish(R, io):
        pool = [p for p in (sorted(io_allowed, reverse=True) if io else range(T - 1, -1, -1))
                if p not in R]
        R = list(R) + pool[: RB - len(R)]
        free = [p for p in range(T) if p not in R]
        return R, _thread_positions(free, sww, io)

    return T, lanes, io_allowed, finish


def _close_rounds(rounds, lanes, io_allowed, finish):
    last = rounds[-1]
    if any(r not in io_allowed for r in last.rpos) or last.tpos[:lanes] != list(range(lanes)):
        R, tpos = finish([], True)
        rounds.append(Round(R, tpos, []))
    return rounds


def schedule_rounds_budget(prims, tile_bits, RB: int, sww: int, max_rounds: int):
    """At most ``max_rounds`` rounds (the last one IO-compatible): returns (rounds, leftover) where
    ``leftover`` are the primitives, in their original order, that did not fit — a set closed
    under "depends on", so the packer can hand them to a later segment."""
    return _schedule_rounds_critical(prims, tile_bits, RB, sww, max_rounds)


def _schedule_rounds_critical(prims, tile_bits, RB: int, sww: int = 3, max_rounds: int | None = None):
    T, lanes, io_allowed, finish = _round_helpers(tile_bits, RB, sww)
    pos = {b: i for i, b in enumerate(tile_bits)}
    ps = _expand_swaps(prims)
    n = len(ps)
    bits = [p.bits for p in ps]
    tpos_of = [[pos[b] for b in p.targets] for p in ps]
    # immediate predecessors / successors through shared bits (global phases order with everything)
    preds = [set() for _ in range(n)]
    succs = [set() for _ in range(n)]
    last_on: dict = {}
    last_global = -1
    for i in range(n):
        if not bits[i]:
            for j in range(i):
                preds[i].add(j)
            last_global = i
        else:
            for b in bits[i]:
                if b in last_on:
                    preds[i].add(last_on[b])
                last_on[b] = i
            if last_global >= 0:
                preds[i].add(last_global)
    for i in range(n):
        for j in preds[i]:
            succs[j].add(i)
    height = [1] * n
    for i in range(n - 1, -1, -1):
        for j in succs[i]:
            height[i] = max(height[i], 1 + height[j])
    done = [False] * n
    npred = [len(preds[i]) for i in range(n)]
    left = n
    rounds: list[Round] = []
    first = True

    def run_round(allowed, commit):
        nonlocal left
        d = list(done)
        c = list(npred)
        R, run = [], []
        ready = sorted((i for i in range(n) if not d[i] and c[i] == 0), key=lambda i: (-height[i], i))
        while True:
            progressed = False
            # everything that fits the current register set, in program order
            for i in sorted(ready):
                if all(t in R for t in tpos_of[i]):
                    run.append(i)
                    d[i] = True
                    ready.remove(i)
                    for j in succs[i]:
                        c[j] -= 1
                        if c[j] == 0:
                            ready.append(j)
                    progressed = True
                    break
            if progressed:
                continue
            # the ready primitive with the longest chain behind it picks the next register bits
            for i in sorted(ready, key=lambda i: (-height[i], i)):
                need = [t for t in tpos_of[i] if t not in R]
                if all(t in allowed for t in tpos_of[i]) and len(R) + len(set(need)) <= RB:
                    for t in need:
                        if t not in R:
                            R.append(t)
                    progressed = True
                    break
            if not progressed:
                break
        if commit:
            for i in run:
                done[i] = True
            for i in range(n):
                npred[i] = c[i]
            left -= len(run)
        return R, run, sum(1 for x in d if not x)

    while left or first:
        closing = max_rounds is not None and len(rounds) == max_rounds - 1
        if first:
            allowed = set(range(min(_first_min_pos(sww), T - RB), T))
            if closing:
                allowed &= io_allowed
            R, run, _ = run_round(allowed, True)
            io = True
        else:
            R, run, rest = run_round(io_allowed, False)
            if rest == 0 or closing:
                R, run, _ = run_round(io_allowed, True)
                io = True
            else:
                R, run, _ = run_round(set(range(T)), True)
                io = all(r in io_allowed for r in R)
        if not run and not first and not closing:
            raise RuntimeError("round scheduling made no progress")   # pragma: no cover
        Rf, tp = finish(R, io)
        rounds.append(Round(Rf, tp, [ps[i] for i in run]))
        first = False
        if closing:
            break
    rounds = _close_rounds(rounds, lanes, io_allowed, finish)
    if max_rounds is None:
        return rounds
    return rounds, [ps[i] for i in range(n) if not done[i]]


def _schedule_rounds_order(prims, tile_bits, RB: int, sww: int = 3):
    T, lanes, io_allowed, finish = _round_helpers(tile_bits, RB, sww)
    pos = {b: i for i, b in enumerate(tile_bits)}
    remaining = _expand_swaps(prims)
    rounds: list[Round] = []

    def greedy(allowed, rem):
        R, run, keep, blocked = [], [], [], set()
        for p in rem:
            pb = p.bits
            tp = [pos[b] for b in p.targets]
            if blocked & pb:
                blocked |= pb
                keep.append(p)
                continue
            need = [t for t in tp if t not in R]
            if all(t in allowed for t in tp) and len(R) + len(need) <= RB:
                R += need
                run.append(p)
            else:
                blocked |= pb
                keep.append(p)
        return R, run, keep

    first = True
    while remaining or first:
        if first:
            # The first round is read out of the landing buffer the bulk copies filled (natural,
            # unswizzled order), not from global memory: any position >= sww may be a register
            # bit; positions 0..sww-1 stay on the lowest lane bits (conflict-free reads).
            R, run, keep = greedy(set(range(min(_first_min_pos(sww), T - RB), T)), remaining)
            io = True
        else:
            # prefer a round that is IO-compatible when it finishes the segment
            R, run, keep = greedy(io_allowed, remaining)
            io = True
            if keep:
                R, run, keep = greedy(set(range(T)), remaining)
                io = all(r in io_allowed for r in R)
        if not run and not first:
            raise RuntimeError("round scheduling made no progress")   # pragma: no cover
        R, tpos = finish(R, io)
        rounds.append(Round(R, tpos, run))
        remaining = keep
        first = False
    return _close_rounds(rounds, lanes, io_allowed, finish)


def _swap_2q(m):
    """4x4 matrix with the roles of its two index bits exchanged."""
    perm = [0, 2, 1, 3]
    return np.asarray(m)[np.ix_(perm, perm)]


def encode_rt_segment(seg: Segment, RB: int, sww: int = 3, rounds=None):
    """Segment -> (ctypes RtOp array, complex128 table, n_records).  ``rounds`` can be passed
    when the caller already scheduled them."""
    tile_bits = seg.tile_bits
    pos = {b: i for i, b in enumerate(tile_bits)}
    if rounds is None:
        rounds = schedule_rounds(seg.prims, tile_bits, RB, sww)
    nrec = sum(1 + len(r.prims) for r in rounds)
    ops_arr = (RtOp * nrec)()
    cols: list[np.ndarray] = []      # (k,) or (B, k) pieces of the matrix table, in order
    nmat = 0

    def push(arr, batched=False):
        nonlocal nmat
        arr = np.asarray(arr, dtype=complex)
        arr = arr.reshape(arr.shape[0], -1) if batched else arr.reshape(-1)
        cols.append(arr)
        nmat += arr.shape[-1]

    i = 0
    for rnd in rounds:
        o = ops_arr[i]; i += 1
        o.kind = RT_ROUND
        for b, p in enumerate(rnd.rpos):
            o.u.r.rbits[b] = p
        for b, p in enumerate(rnd.tpos):
            o.u.r.tbits[b] = p
        rbit = {p: b for b, p in enumerate(rnd.rpos)}
        tbit = {p: b for b, p in enumerate(rnd.tpos)}

        def split_mask(bits_vals):
            """{global bit: value} -> (mask_r, val_r, mask_t, val_t, mask_e, val_e)."""
            mr = vr = mt = vt = me = ve = 0
            for b, v in bits_vals.items():
                if b in pos:
                    p = pos[b]
                    if p in rbit:
                        mr |= 1 << rbit[p]; vr |= (1 << rbit[p]) if v else 0
                    else:
                        mt |= 1 << tbit[p]; vt |= (1 << tbit[p]) if v else 0
                else:
                    me |= 1 << b; ve |= (1 << b) if v else 0
            return mr, vr, mt, vt, me, ve

        for p in rnd.prims:
            o = ops_arr[i]; i += 1
            o.mat_off = nmat
            if p.kind == GEN:
                o.kind = RT_GEN
                o.q0 = int(p.slot)
                o.q1 = int(p.ny) & 3
                xr = 0
                for b in p.targets:
                    xr |= 1 << rbit[pos[b]]
                zr, _, zt, _, ze, _ = split_mask({b: 1 for b in p.zbits})
                o.u.p.xr, o.u.p.zr, o.u.p.zt, o.u.p.ze = xr, zr, zt, ze
                o.u.p.coef = float(p.coef)
                continue
            if p.kind == DIAG:
                o.kind = DIAG
                o.q0 = len(p.other)
                for j, b in enumerate(p.other):
                    if b in pos:
                        pp = pos[b]
                        o.u.d.src[j] = rbit[pp] if pp in rbit else 32 + tbit[pp]
                    else:
                        o.u.d.src[j] = 64 + b
                push(p.mat)
            else:
                o.kind = p.kind
                g = o.u.g
                g.ctrl_r, g.cval_r, g.ctrl_t, g.cval_t, g.ctrl_e, g.cval_e = split_mask(p.ctrl)
                if p.kind == DENSE1:
                    o.q0 = rbit[pos[p.targets[0]]]
                    push(p.mat, np.ndim(p.mat) == 3)
                    if p.mat0 is not None:
                        o.kind = DENSE1 | 0x100
                        push(p.mat0, np.ndim(p.mat0) == 3)
                elif p.kind == DENSE2:
                    q0, q1 = rbit[pos[p.targets[0]]], rbit[pos[p.targets[1]]]
                    m = np.asarray(p.mat, dtype=complex)
                    if q0 < q1:
                        q0, q1, m = q1, q0, _swap_2q(m)
                    o.q0, o.q1 = q0, q1
                    push(m)
                elif p.kind == CX:
                    o.q0 = rbit[pos[p.targets[0]]]
                elif p.kind == PARITY:
                    g.par_r, _, g.par_t, _, g.par_e, _ = split_mask({b: 1 for b in p.other})
                    push(np.asarray(p.mat, dtype=complex).reshape(-1)[:2])
                else:  # pragma: no cover
                    raise ValueError(f"primitive kind {p.kind} has no register-kernel record")
            if nmat % 2:
                push([0j])
    if not cols:
        push([0j, 0j])
    B = max((c.shape[0] for c in cols if c.ndim == 2), default=0)
    if B:        # broadcast parameters: one table row per batch element
        table = np.concatenate([c if c.ndim == 2 else np.broadcast_to(c, (B, c.shape[0])) for c in cols],
                               axis=1)
    else:
        table = np.concatenate(cols)
    table = np.ascontiguousarray(table, dtype=np.complex128)
    return ops_arr, table, nrec
