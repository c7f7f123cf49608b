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
from dataclasses import dataclass, field

import numpy as np

DENSE1, DENSE2, CX, PARITY, DIAG, SWAP = 0, 1, 2, 3, 4, 5
GENERIC = -1


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
    ngates: int = 1
    op: object = None                                # GENERIC: the original operator

    @property
    def bits(self):
        return set(self.targets) | set(self.ctrl) | set(self.other)


_X = np.array([[0, 1], [1, 0]], dtype=complex)
_PHASES = {"PauliZ": -1.0 + 0j, "S": 1j, "T": np.exp(0.25j * np.pi)}


def _is_diag(m):
    return not np.any(m[~np.eye(m.shape[0], dtype=bool)])


def lower(op, bit_of) -> list[Prim]:
    """Operator -> primitives (apply_operation.py:258-351 dispatch, re-done for the tile kernel)."""
    name = op.name
    wires = list(op.wires)
    bits = [bit_of(w) for w in wires]
    if name in ("Identity", "Barrier", "Snapshot", "WireCut"):
        return []
    if getattr(op, "batch_size", None) is not None or hasattr(op, "state_vector"):
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


def merge_blocks(prims: list[Prim], level: int = 1) -> list[Prim]:
    """level 0: nothing; 1: products of single-qubit runs; 2: also absorb single-qubit blocks
    into adjacent dense two-qubit gates and merge consecutive dense gates on the same pair."""
    if level <= 0:
        return list(prims)
    out: list[Prim] = []
    pending: dict[int, Prim] = {}            # bit -> accumulated 1q block (not yet emitted)

    def flush(bits):
        for b in sorted(bits):
            if b in pending:
                out.append(pending.pop(b))

    for p in prims:
        one = _as_1q_matrix(p) if p.kind != GENERIC else None
        if one is not None:
            b, m = one
            if b in pending:
                q = pending[b]
                q.mat = np.asarray(m) @ q.mat
                q.ngates += p.ngates
            else:
                pending[b] = Prim(DENSE1, targets=[b], mat=np.asarray(m, dtype=complex),
                                  ngates=p.ngates)
            continue
        if level >= 2 and p.kind == DENSE2 and not p.ctrl:
            # absorb pending single-qubit blocks that precede this gate on its two bits
            m = p.mat
            for pos, b in enumerate(p.targets):
                if b in pending:
                    q = pending.pop(b)
                    m = m @ _embed(q.mat, pos)
                    p.ngates += q.ngates
            p.mat = m
            # merge with an immediately preceding dense gate on the same ordered pair
            if out and out[-1].kind == DENSE2 and not out[-1].ctrl and out[-1].targets == p.targets:
                out[-1].mat = p.mat @ out[-1].mat
                out[-1].ngates += p.ngates
            else:
                flush(p.bits)
                out.append(p)
            continue
        flush(p.bits)
        out.append(p)
    flush(list(pending))
    # normalise accumulated 1q blocks (a product may have become diagonal / X / identity)
    norm = []
    for p in out:
        if p.kind == DENSE1 and not p.ctrl:
            q = _dense(p.targets, {}, p.mat)
            q.ngates = p.ngates
            if q.kind == DIAG and np.allclose(q.mat, 1.0):
                q = Prim(PARITY, mat=np.array([1.0 + 0j, 1.0 + 0j]), ngates=p.ngates)
            norm.append(q)
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


def pack_segments(prims: list[Prim], n: int, T: int = 12, L: int = 5, max_ops: int = 96,
                  max_mat: int = 1024) -> list[Segment]:
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
            msize = 0 if p.mat is None else int(np.size(p.mat))
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
            segments.append(Segment(bits, seg, sum(p.ngates for p in seg)))
        remaining = keep
    return segments


def compile_ops(ops_, n: int, bit_of=None, level: int = 1, T: int = 12, L: int = 5):
    """Operators -> list of :class:`Segment`."""
    if bit_of is None:
        bit_of = lambda w: n - 1 - int(w)          # noqa: E731
    prims: list[Prim] = []
    for op in ops_:
        prims.extend(lower(op, bit_of))
    prims = merge_blocks(prims, level)
    return pack_segments(prims, n, T=T, L=L)


def encode_segment(seg: Segment):
    """Segment -> (ctypes TileOp array, complex128 matrix table)."""
    pos = {b: i for i, b in enumerate(seg.tile_bits)}
    ops_arr = (TileOp * len(seg.prims))()
    mats: list[complex] = []
    for i, p in enumerate(seg.prims):
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
