"""Structure-specialised segment kernels: plan -> CUDA source -> NVRTC -> launch.

The register-tiled interpreter (csrc/rtile.cuh) decodes 64-byte records per tile; ncu showed more
than half of its issued instructions were decode / address / predicate work.  Here the host turns
a :class:`~pennylane_b200.compiler.Segment` into

  * a *structure*: round layouts, record kinds, register bits, control locations, Pauli masks —
    emitted as two small headers (``sk_config.inc``, ``sk_body.inc``) for ``csrc/segk.cuh`` and
    compiled ONCE per structure by NVRTC (``b200q_jit_compile``; cached in memory and on disk by
    the hash of the two headers + the kernel source), and
  * the *values*: a flat coefficient table in the order the body consumes it, recomputed per
    call (parameter rebinding never recompiles).

Single-qubit blocks are applied in the normalised form ``U = s * diag(1,l) K(t) diag(1,r)``
(:func:`canon_1q`); the scalars ``s`` of all records of a segment are multiplied together on the
host and applied once by the kernel at the end of the segment.

The plan is kept as a small IR (``SegPlan.ir``) from which both the CUDA body and the test-only
numpy emulator (``tests/sk_emulator.py``) are derived.

Reference analogue: none — default.qubit applies one gate per sweep
(pennylane/devices/qubit/simulate.py:214-235, adjoint_jacobian.py:121-137).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import threading
from dataclasses import dataclass, field

import numpy as np

from . import compiler as cc
from ._lib import B200QError, check, int_array, load

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")

#: |phase - 1| below this is treated as exactly 1 (the record then skips the multiplication);
#: the relative error this leaves is far below the 1e-12 parity bar
PHASE_TOL = 1e-15
#: coefficient tables up to this size travel as a kernel parameter (constant bank operands)
COEF_PARAM_MAX_BYTES = 3072
#: a 2x2 block is normalised only when it is unitary to this tolerance (else the plain product)
UNITARY_TOL = 1e-12


# ---------------------------------------------------------------------------------------------
# normalised single-qubit blocks
# ---------------------------------------------------------------------------------------------
@dataclass
class Canon:
    kern: int          # 0: real kernel [[1,-t],[t,1]],  1: imaginary kernel [[1,-it],[-it,1]]
    sinp: bool         # sin pivot: [[t,-1],[1,t]] / [[t,-i],[-i,t]]
    t: float
    r: complex         # right phase diag(1, r)
    l: complex         # left phase diag(1, l)
    s: complex         # scalar left out: U = s * diag(1,l) K diag(1,r)

    @property
    def dr(self) -> bool:
        return abs(self.r - 1.0) > PHASE_TOL

    @property
    def dl(self) -> bool:
        return abs(self.l - 1.0) > PHASE_TOL

    def matrix(self) -> np.ndarray:
        t = self.t
        if self.kern == 0:
            k = np.array([[t, -1.0], [1.0, t]]) if self.sinp else np.array([[1.0, -t], [t, 1.0]])
        else:
            k = np.array([[t, -1j], [-1j, t]]) if self.sinp else np.array([[1.0, -1j * t], [-1j * t, 1.0]])
        return self.s * (np.diag([1.0, self.l]) @ k @ np.diag([1.0, self.r]))


def _canon_candidates(u: np.ndarray):
    u00, u01, u10, u11 = u[0, 0], u[0, 1], u[1, 0], u[1, 1]
    sinp = abs(u00) < abs(u10)
    for kern in (0, 1):
        for sigma in (1.0, -1.0):
            if not sinp:
                s = u00
                t = sigma * abs(u10 / u00)
                if t == 0.0:
                    if sigma < 0:
                        continue
                    r, l = 1.0 + 0j, u11 / u00
                elif kern == 0:
                    l, r = (u10 / s) / t, -(u01 / s) / t
                else:
                    l, r = 1j * (u10 / s) / t, 1j * (u01 / s) / t
            else:
                if u00 == 0.0:
                    if sigma < 0:
                        continue
                    t = 0.0
                    s = u10 if kern == 0 else 1j * u10         # l = 1
                else:
                    t = sigma * abs(u00) / abs(u10)
                    s = u00 / t
                if kern == 0:
                    l, r = u10 / s, -u01 / s
                else:
                    l, r = 1j * u10 / s, 1j * u01 / s
            yield Canon(kern, sinp, float(t), complex(r), complex(l), complex(s))


def canon_1q(u, hint=None):
    """Normalised form of a 2x2 unitary, or None when ``u`` is not unitary (to UNITARY_TOL) —
    the caller then applies the plain product.  ``hint = (kern, dl, dr)``: prefer a form that
    fits a previously compiled structure (same kernel, no phase the structure does not have)."""
    u = np.asarray(u, dtype=np.complex128)
    if u.shape != (2, 2) or not np.all(np.isfinite(u)):
        return None
    if np.max(np.abs(u.conj().T @ u - np.eye(2))) > UNITARY_TOL:
        return None
    best, best_cost = None, None
    for c in _canon_candidates(u):
        if np.max(np.abs(c.matrix() - u)) > 1e-13:
            continue                                   # pragma: no cover - numerical safety net
        cost = int(c.dl) + int(c.dr)
        if hint is not None:
            fits = c.kern == hint[0] and (hint[1] or not c.dl) and (hint[2] or not c.dr)
            cost = cost if fits else cost + 10
        elif c.kern == 1:
            cost += 0.5 if cost else 0.25              # ties go to the real kernel
        if best is None or cost < best_cost:
            best, best_cost = c, cost
    return best


# ---------------------------------------------------------------------------------------------
# plan
# ---------------------------------------------------------------------------------------------
@dataclass
class Geometry:
    """Kernel shape: complex128 forward = 256 threads x 16 amplitudes, 2 CTAs/SM."""
    dtype_code: int            # 1 = complex128
    RB: int
    TB: int
    NV: int = 1
    MINB: int = 2

    @property
    def T(self) -> int:
        return self.RB + self.TB

    @property
    def sww(self) -> int:
        return 3 if self.dtype_code else 4


def default_low_bits(dtype_code: int, nv: int = 1) -> int:
    """Contiguous low tile bits (L) the packer reserves for the specialised kernels.  complex128
    forward: 4 (256-byte runs: the same copy ceiling as 512-byte ones, and 8 free tile bits instead
    of 7 — the 30-qubit ansatz packs into 27 segments / 67 rounds instead of 29 / 77: 187.8 ->
    174.5 ms per step, tools/microbench_segk.py); everything else: 5."""
    env = os.environ.get("B200Q_TILE_L")
    if env:
        return int(env)
    return 4 if (dtype_code and nv == 1) else 5


def default_geometry(dtype_code: int, nv: int = 1) -> Geometry:
    """Forward: T = 12 (c128) / 13 (c64), 256 threads, 2 CTAs per SM.  Adjoint (two vectors per
    thread): one register bit fewer, 512 threads, 1 CTA per SM (B200Q_SK_ADJ=1: 256 x 2)."""
    env = os.environ.get("B200Q_SK_FWD" if nv == 1 else "B200Q_SK_ADJGEOM")     # tuning knob: "RB,TB,MINB"
    if env:
        rb, tb, minb = (int(x) for x in env.split(","))
        return Geometry(dtype_code, rb + (0 if dtype_code else 1), tb, nv, minb)
    if nv == 1:
        return Geometry(dtype_code, 4 if dtype_code else 5, 8, 1, 2)
    if int(os.environ.get("B200Q_SK_ADJ", "0")):
        return Geometry(dtype_code, 3 if dtype_code else 4, 8, 2, 2)
    return Geometry(dtype_code, 3 if dtype_code else 4, 9, 2, 1)


@dataclass
class SegPlan:
    geom: Geometry
    L: int
    tile_bits: list
    rounds: list                      # [(rpos, tpos)]
    ir: list                          # records, see _Builder
    ext_pos: list                     # global bit positions of the external predicate bits
    ncoef: int                        # reals in the coefficient table (even)
    nslots: int
    fill: list                        # value producers, see coefficients()
    forms: dict                       # prim index -> (kern, dl, dr) of the normalised records
    config: str = ""
    body: str = ""
    key: str = ""
    slot_params: list = field(default_factory=list)   # slot -> trainable parameter index
    coef_param: bool = False          # the coefficient table is a kernel parameter (constant bank)


class FormMismatch(Exception):
    """New parameter values do not fit the normalised forms of a cached structure."""


class _Builder:
    def __init__(self, geom: Geometry, tile_bits, L):
        self.g = geom
        self.tile_bits = list(tile_bits)
        self.L = L
        self.pos = {b: i for i, b in enumerate(tile_bits)}
        self.ir = []
        self.ext = []                 # external bits in slot order
        self.ncoef = 0
        self.fill = []
        self.forms = {}
        self.norm_idx = []            # prim indices whose scalar goes into the segment scalar

    # -- coefficient table ---------------------------------------------------------------------
    def alloc(self, nreal, *producer):
        """Reserve ``nreal`` reals; ``producer`` = (kind, prim index, extra) says how to fill them."""
        off = self.ncoef
        self.ncoef += nreal + (nreal & 1)
        self.fill.append((off, nreal) + tuple(producer))
        return off

    # -- locations -------------------------------------------------------------------------------
    def ext_slot(self, b):
        if b not in self.ext:
            if len(self.ext) >= 16:
                raise ValueError("more than 16 external predicate bits in one segment")
            self.ext.append(b)
        return self.ext.index(b)

    def split(self, bits_vals, rbit, tbit):
        """{global bit: value} -> (mask_r, val_r, mask_t, val_t, mask_e, val_e) with the external
        part over the segment's external slots."""
        mr = vr = mt = vt = me = ve = 0
        for b, v in bits_vals.items():
            if b in self.pos:
                p = self.pos[b]
                if p in rbit:
                    mr |= 1 << rbit[p]; vr |= (1 << rbit[p]) if v else 0
                else:
                    mt |= 1 << tbit[p]; vt |= (1 << tbit[p]) if v else 0
            else:
                e = self.ext_slot(b)
                me |= 1 << e; ve |= (1 << e) if v else 0
        return mr, vr, mt, vt, me, ve


def _complex_pairs(arr):
    arr = np.asarray(arr, dtype=np.complex128).reshape(-1)
    out = np.empty(2 * arr.size)
    out[0::2] = arr.real
    out[1::2] = arr.imag
    return out


def plan_segment(seg, geom: Geometry, L: int, forms_hint=None, final_scale: bool = True) -> SegPlan:
    """Segment -> :class:`SegPlan`.  ``forms_hint``: ``SegPlan.forms`` of a cached plan of the
    same circuit structure (the new plan then has the same key whenever the values allow)."""
    tile_bits = list(seg.tile_bits)
    assert len(tile_bits) == geom.T
    RB, TB = geom.RB, geom.TB
    rounds = getattr(seg, "rounds", None)
    if rounds and len(rounds[0].rpos) != RB:         # scheduled by the packer for another geometry
        rounds = None
    rounds = rounds or cc.schedule_rounds(seg.prims, tile_bits, RB, geom.sww)
    bld = _Builder(geom, tile_bits, L)
    index_of = {id(p): i for i, p in enumerate(seg.prims)}
    nrounds = len(rounds)
    slot_params: list = []
    # scalars of normalised records, in program order, for the GEN corrections: gen_pre[i] = list
    # of prim indices normalised before GEN record i
    normalised_so_far: list = []
    batched = any(p.kind == cc.DENSE1 and np.ndim(p.mat) == 3 for p in seg.prims)

    for ri, rnd in enumerate(rounds):
        rbit = {p: b for b, p in enumerate(rnd.rpos)}
        tbit = {p: b for b, p in enumerate(rnd.tpos)}
        if ri == 0:
            bld.ir.append(("load", 0))
            if nrounds == 1:
                bld.ir.append(("fetch",))
        else:
            bld.ir.append(("xpose", ri - 1, ri))
            if ri == nrounds - 1:
                bld.ir.append(("fetch",))
        for p in rnd.prims:
            pi = index_of.get(id(p), -1)
            if p.kind == cc.GEN:
                xr = 0
                for b in p.targets:
                    xr |= 1 << rbit[bld.pos[b]]
                zr, _, zt, _, ze, _ = bld.split({b: 1 for b in p.zbits}, rbit, tbit)
                if p.param not in slot_params:
                    slot_params.append(p.param)
                slot = slot_params.index(p.param)
                pre = list(normalised_so_far)
                off = bld.alloc(2, "gen", pi, pre)
                bld.ir.append(("gen", xr, zr, int(p.ny) & 1, slot, off, (zt, ze, (int(p.ny) >> 1) & 1)))
                continue
            if p.kind == cc.DIAG:
                nd = len(p.other)
                rc = [0] * 5
                it = []                                  # (kind, index, shift)
                for j, b in enumerate(p.other):
                    w = nd - 1 - j
                    if b in bld.pos:
                        pp = bld.pos[b]
                        if pp in rbit:
                            rc[rbit[pp]] |= 1 << w
                        else:
                            it.append(("t", tbit[pp], w))
                    else:
                        it.append(("e", bld.ext_slot(b), w))
                off = bld.alloc(2 << nd, "mat", pi, "mat")
                bld.ir.append(("diag", tuple(rc), off, tuple(it), nd))
                continue
            mr, vr, mt, vt, me, ve = bld.split(p.ctrl, rbit, tbit)
            pred = (mt, vt, me, ve)
            if p.kind == cc.DENSE1:
                q = rbit[bld.pos[p.targets[0]]]
                form = None
                if not p.ctrl and p.mat0 is None:
                    form = _choose_form(p.mat, None if forms_hint is None else forms_hint.get(pi))
                if form is not None:
                    kern, dl, dr = form
                    bld.forms[pi] = form
                    off = bld.alloc(2 + 2 * dl + 2 * dr, "dk", pi, form)
                    bld.ir.append(("dk", q, kern, int(dl), int(dr), off))
                    normalised_so_far.append(pi)
                    bld.norm_idx.append(pi)
                elif p.mat0 is not None:
                    off = bld.alloc(16, "mat", pi, "mat+mat0")
                    bld.ir.append(("f16", q, mr, vr, 1, off, pred))
                else:
                    off = bld.alloc(8, "mat", pi, "mat")
                    bld.ir.append(("f16", q, mr, vr, 0, off, pred))
            elif p.kind == cc.DENSE2:
                q0, q1 = rbit[bld.pos[p.targets[0]]], rbit[bld.pos[p.targets[1]]]
                swap = q0 < q1
                if swap:
                    q0, q1 = q1, q0
                off = bld.alloc(32, "mat", pi, "mat2s" if swap else "mat")
                bld.ir.append(("d2", q0, q1, mr, vr, off, pred))
            elif p.kind == cc.CX:
                q = rbit[bld.pos[p.targets[0]]]
                bld.ir.append(("cx", q, mr, vr, pred if (mt or me) else None))
            elif p.kind == cc.PARITY:
                pr, _, pt, _, pe, _ = bld.split({b: 1 for b in p.other}, rbit, tbit)
                norm = (not p.ctrl) and bool(p.other) and not batched and _parity_normalisable(p.mat)
                if norm:
                    off = bld.alloc(2, "par", pi, True)
                    normalised_so_far.append(pi)
                    bld.norm_idx.append(pi)
                else:
                    off = bld.alloc(4, "par", pi, False)
                bld.ir.append(("par", mr, vr, pr, (pt, pe), int(norm), off, pred))
            else:  # pragma: no cover
                raise ValueError(f"primitive kind {p.kind} has no specialised record")
    if bld.norm_idx and final_scale:
        off = bld.alloc(2, "scale", -1, list(bld.norm_idx))
        bld.ir.append(("scale", off))
    last = rounds[-1]
    lanes = cc.io_lanes(tile_bits, RB)
    assert all(r >= lanes for r in last.rpos) and last.tpos[:lanes] == list(range(lanes))
    plan = SegPlan(geom, L, tile_bits, [(list(r.rpos), list(r.tpos)) for r in rounds], bld.ir,
                   list(bld.ext), max(2, bld.ncoef), len(slot_params), bld.fill, bld.forms,
                   slot_params=slot_params)
    real_bytes = 8 if geom.dtype_code else 4
    plan.coef_param = (not batched) and plan.ncoef * real_bytes <= COEF_PARAM_MAX_BYTES
    plan.config = emit_config(plan)
    plan.body = emit_body(plan)
    plan.key = hashlib.sha256((plan.config + "\n//--\n" + plan.body).encode()).hexdigest()
    return plan




# ---------------------------------------------------------------------------------------------
# values: coefficient table of a plan for given primitives
# ---------------------------------------------------------------------------------------------
def _parity_normalisable(mat):
    m = np.asarray(mat, dtype=complex).reshape(-1)
    return abs(abs(m[0]) - 1.0) < 1e-12


def _choose_form(mat, hint):
    """(kern, dl, dr) for an uncontrolled block, None -> plain product.  Broadcast blocks
    (leading batch axis) take the union over the batch."""
    m = np.asarray(mat, dtype=np.complex128)
    forms = []
    for u in (m if m.ndim == 3 else m[None]):
        c = canon_1q(u, hint)
        if c is None:
            return None
        forms.append((c.kern, c.dl, c.dr))
    kern = forms[0][0]
    if any(f[0] != kern for f in forms):
        return (0, True, True)
    dl = any(f[1] for f in forms)
    dr = any(f[2] for f in forms)
    if hint is not None and hint[0] == kern:
        dl, dr = dl or hint[1], dr or hint[2]
    return (kern, dl, dr)


def _canon_for(u, form):
    c = canon_1q(u, form)
    if c is None or c.kern != form[0] or (c.dl and not form[1]) or (c.dr and not form[2]):
        raise FormMismatch()
    return c


def _complex_pairs(arr):
    arr = np.asarray(arr, dtype=np.complex128).reshape(-1)
    out = np.empty(2 * arr.size)
    out[0::2] = arr.real
    out[1::2] = arr.imag
    return out


def coefficients(plan: SegPlan, prims) -> np.ndarray:
    """The coefficient table of ``plan`` for the values in ``prims`` (same structure as the
    primitives the plan was made from): float64 ``[ncoef]``, or ``[B, ncoef]`` when some block
    carries broadcast matrices.  Raises :class:`FormMismatch` when a block no longer fits the
    normalised form compiled into the plan."""
    B = 0
    for p in prims:
        if p.kind == cc.DENSE1 and np.ndim(p.mat) == 3:
            B = max(B, np.shape(p.mat)[0])

    def pick(m, b):
        m = np.asarray(m, dtype=np.complex128)
        return m[b] if (m.ndim == 3 and b is not None) else m

    rows = []
    for b in (range(B) if B else [None]):
        tab = np.zeros(plan.ncoef)
        scalars, canon = {}, {}
        for off, nreal, kind, pi, extra in plan.fill:        # scalars of the normalised records
            if kind == "dk":
                canon[pi] = c = _canon_for(pick(prims[pi].mat, b), extra)
                scalars[pi] = c.s
            elif kind == "par" and extra:
                m = np.asarray(prims[pi].mat, dtype=np.complex128).reshape(-1)
                if abs(abs(m[0]) - 1.0) >= 1e-12:
                    raise FormMismatch()
                scalars[pi] = complex(m[0])
        for off, nreal, kind, pi, extra in plan.fill:
            if kind == "dk":
                c = canon[pi]
                vals = [c.t, 1.0 if c.sinp else 0.0]
                if extra[2]:
                    vals += [c.r.real, c.r.imag]
                if extra[1]:
                    vals += [c.l.real, c.l.imag]
            elif kind == "mat":
                p = prims[pi]
                if extra == "mat":
                    vals = _complex_pairs(pick(p.mat, b))
                elif extra == "mat2s":
                    vals = _complex_pairs(cc._swap_2q(pick(p.mat, b)))
                else:
                    vals = np.concatenate([_complex_pairs(pick(p.mat, b)), _complex_pairs(pick(p.mat0, b))])
            elif kind == "par":
                m = np.asarray(prims[pi].mat, dtype=np.complex128).reshape(-1)[:2]
                if extra:
                    q = m[1] / m[0]
                    vals = [q.real, q.imag]
                else:
                    vals = [m[0].real, m[0].imag, m[1].real, m[1].imag]
            elif kind == "gen":
                s2 = 1.0
                for j in extra:
                    s2 *= abs(scalars[j]) ** 2
                vals = [float(prims[pi].coef) * s2, 0.0]
            elif kind == "scale":
                s = 1.0 + 0j
                for j in extra:
                    s *= scalars[j]
                vals = [s.real, s.imag]
            else:  # pragma: no cover
                raise AssertionError(kind)
            if len(vals) != nreal:
                raise FormMismatch()
            tab[off: off + nreal] = np.asarray(vals, dtype=float)
        rows.append(tab)
    return np.ascontiguousarray(rows[0] if not B else np.stack(rows))


# ---------------------------------------------------------------------------------------------
# structure -> CUDA headers
# ---------------------------------------------------------------------------------------------
def _arr2(rows):
    return "{" + ",".join("{" + ",".join(str(int(v)) for v in r) + "}" for r in rows) + "}"


def emit_config(plan: SegPlan) -> str:
    g = plan.geom
    lines = [
        f"#define SK_REAL {'double' if g.dtype_code else 'float'}",
        f"#define SK_IS_DOUBLE {1 if g.dtype_code else 0}",
        f"#define SK_RB {g.RB}", f"#define SK_TB {g.TB}", f"#define SK_NV {g.NV}",
        f"#define SK_L {plan.L}", f"#define SK_MINB {g.MINB}",
        f"#define SK_NROUNDS {len(plan.rounds)}", f"#define SK_NCOEF {plan.ncoef}",
        f"#define SK_NSLOTS {plan.nslots}", f"#define SK_NEXT {len(plan.ext_pos)}",
        f"#define SK_SWW {g.sww}",
        f"#define SK_COEF_PARAM {1 if plan.coef_param else 0}",
        # the cap leaves room for the exchange's unpack kernel beside FORWARD segments; the reverse
        # sweep (NV = 2, one 512-thread CTA per SM) never runs inside an exchange window
        f"#define SK_MAXREG {_max_registers() if g.NV == 1 else 128}",
        f"#define SK_RPOS {_arr2([r for r, _ in plan.rounds])}",
        f"#define SK_TPOS {_arr2([t for _, t in plan.rounds])}",
    ]
    return "\n".join(lines) + "\n"


def _pred_expr(pred) -> str:
    mt, vt, me, ve = pred
    parts = []
    if mt:
        parts.append(f"((tid & {mt}u) == {vt}u)")
    if me:
        parts.append(f"((ext & {me}u) == {ve}u)")
    return " && ".join(parts) if parts else "true"


def _par_expr(mt, me, const=0) -> str:
    parts = []
    if mt:
        parts.append(f"__popc(tid & {mt}u)")
    if me:
        parts.append(f"__popc(ext & {me}u)")
    if const:
        parts.append(f"{int(const)}")
    return "((" + " + ".join(parts) + ") & 1u)" if parts else "0u"


#: Generator terms whose warp sums share one butterfly.  0 (default): ADJACENT terms only (no gate
#: between them), so no value stays live across gate code; 2 / 4: groups across gates — measured
#: slower (0.875 s vs 0.785 s per 30-qubit Jacobian: the pending values spill at the register cap);
#: 1: one reduction per term (0.794 s).
_GEN_GROUP = int(os.environ.get("B200Q_SK_GEN_GROUP", "0"))


def emit_body(plan: SegPlan) -> str:
    out = []
    pending = []                                    # [(variable, slot)] generator terms awaiting their warp sum

    def flush():
        # one butterfly per group of 4 / 2 / 1 values with DISTINCT slots (sk_gen_flush*)
        while pending:
            grp = pending[:4] if len(pending) >= 4 else pending[:len(pending)]
            del pending[:len(grp)]
            names = ", ".join(v for v, _ in grp)
            slots = ", ".join(str(sl) for _, sl in grp)
            out.append(f"sk_gen_flush{len(grp)}<{slots}>({names}, accs, tid);")

    ngen = 0
    for rec in plan.ir:
        k = rec[0]
        if k in ("xpose", "scale") or (_GEN_GROUP == 0 and k != "gen"):
            flush()                                 # do not carry pending values across a transposition
                                                    # (group 0: nor across any gate — adjacent terms only)
        if k == "load":
            out.append(f"SK_LOAD({rec[1]})")
        elif k == "fetch":
            out.append("SK_FETCH_NEXT()")
        elif k == "xpose":
            out.append(f"SK_XPOSE({rec[1]}, {rec[2]})")
        elif k == "dk":
            _, q, kern, dl, dr, off = rec
            out.append(f"sk_dk<{q}, {kern}, {'true' if dl else 'false'}, {'true' if dr else 'false'}>"
                       f"(A, SK_COEF({off}));")
        elif k == "f16":
            _, q, mr, vr, has0, off, pred = rec
            out.append(f"sk_f16<{q}, {mr}u, {vr}u, {'true' if has0 else 'false'}>(A, SK_COEF({off}), "
                       f"{_pred_expr(pred)});")
        elif k == "d2":
            _, q0, q1, mr, vr, off, pred = rec
            out.append(f"sk_d2<{q0}, {q1}, {mr}u, {vr}u>(A, SK_COEF({off}), {_pred_expr(pred)});")
        elif k == "cx":
            _, q, mr, vr, pred = rec
            if pred is None:
                out.append(f"sk_cx<{q}, {mr}u, {vr}u, false>(A, true);")
            else:
                out.append(f"sk_cx<{q}, {mr}u, {vr}u, true>(A, {_pred_expr(pred)});")
        elif k == "par":
            _, mr, vr, pr, (pt, pe), norm, off, pred = rec
            rt = bool(pt or pe)
            out.append(f"sk_par<{mr}u, {vr}u, {pr}u, {'true' if rt else 'false'}, "
                       f"{'true' if norm else 'false'}>(A, SK_COEF({off}), {_pred_expr(pred)}, "
                       f"{_par_expr(pt, pe)});")
        elif k == "diag":
            _, rc, off, items, nd = rec
            parts = []
            for kind, idx, w in items:
                parts.append(f"(SK_TBIT({idx}) << {w})" if kind == "t" else f"(SK_EBIT({idx}) << {w})")
            i0 = " | ".join(parts) if parts else "0u"
            out.append(f"sk_diag<{rc[0]}u, {rc[1]}u, {rc[2]}u, {rc[3]}u, {rc[4]}u>(A, SK_COEF({off}), {i0});")
        elif k == "gen":
            _, xr, zr, odd, slot, off, (zt, ze, c) = rec
            expr = (f"sk_gen_val<{xr}u, {zr}u, {'true' if odd else 'false'}>(A, SK_COEF({off}), "
                    f"{_par_expr(zt, ze, c)})")
            same = [v for v, sl in pending if sl == slot]
            if same:                                # same parameter: add in registers, one slot per butterfly
                out.append(f"{same[0]} += {expr};")
            else:
                out.append(f"double g{ngen} = {expr};")
                pending.append((f"g{ngen}", slot))
                ngen += 1
                if len(pending) == (_GEN_GROUP or 4):
                    flush()
        elif k == "scale":
            out.append(f"sk_scale(A, SK_COEF({rec[1]}));")
        else:  # pragma: no cover
            raise AssertionError(k)
    flush()
    return "\n".join("    " + line for line in out) + "\n"


# ---------------------------------------------------------------------------------------------
# kernel cache: structure key -> loaded kernel
# ---------------------------------------------------------------------------------------------
_LOCK = threading.Lock()
_KERNELS: dict = {}              # key -> ctypes handle (void*)
_CUBINS: dict = {}               # key -> bytes (compiled in this process, not yet loaded)
_STATS = {"compiled": 0, "disk_hits": 0, "mem_hits": 0}


def _source():
    with open(os.path.join(_CSRC, "segk.cuh")) as f:
        src = f.read()
    with open(os.path.join(_CSRC, "segk_args.h")) as f:
        args = f.read()
    return src, args


def _cache_dirs():
    """In-tree cache first (cubins of the benchmark structures can be built ahead of time by
    ``build()``: NVRTC needs no GPU), then the user's cache directory."""
    dirs = [os.path.join(_CSRC, "jit_cache")]
    user = os.environ.get("B200Q_JIT_CACHE") or os.path.join(
        os.environ.get("XDG_CACHE_HOME", os.path.join(os.path.expanduser("~"), ".cache")), "b200q")
    dirs.append(user)
    return dirs


_SRC_HASH = None


def _max_registers() -> int:
    """Register cap of the specialised kernels (``--maxrregcount``).  120: two resident CTAs of
    256 threads (or one of 512) leave 4096 registers of an SM free, which is what lets the
    one-warp CTAs of the exchange's unpack kernel (csrc/remap.cu) run BESIDE a segment launch;
    at 126-128 registers (the five-round segments) they waited for the launch boundary
    (tools/micro_corun.py).  B200Q_SK_MAXREG overrides (0 = no cap)."""
    return int(os.environ.get("B200Q_SK_MAXREG", "120"))


def _full_key(plan_key: str) -> str:
    global _SRC_HASH
    if _SRC_HASH is None:
        src, args = _source()
        _SRC_HASH = hashlib.sha256((src + args).encode()).hexdigest()
    return hashlib.sha256((plan_key + _SRC_HASH).encode()).hexdigest()[:40]


def compile_plan(plan: SegPlan, save_dir: str | None = None) -> bytes:
    """cubin of the plan's structure (memory -> disk -> NVRTC).  Needs no GPU."""
    key = _full_key(plan.key)
    with _LOCK:
        if key in _CUBINS:
            return _CUBINS[key]
    for d in _cache_dirs():
        path = os.path.join(d, key + ".cubin")
        if os.path.exists(path):
            with open(path, "rb") as f:
                data = f.read()
            with _LOCK:
                _CUBINS[key] = data
                _STATS["disk_hits"] += 1
            return data
    lib = load()
    src, args = _source()
    names = [b"sk_config.inc", b"sk_body.inc", b"segk_args.h"]
    texts = [plan.config.encode(), plan.body.encode(), args.encode()]
    n_arr = (C.c_char_p * 3)(*names)
    t_arr = (C.c_char_p * 3)(*texts)
    out, size = C.c_void_p(), C.c_size_t()
    check(lib.b200q_jit_compile(src.encode(), n_arr, t_arr, 3, 1, 0, C.byref(out), C.byref(size)))
    data = C.string_at(out, size.value)
    lib.b200q_jit_free(out)
    with _LOCK:
        _CUBINS[key] = data
        _STATS["compiled"] += 1
    for d in ([save_dir] if save_dir else _cache_dirs()[1:]):
        try:
            os.makedirs(d, exist_ok=True)
            tmp = os.path.join(d, f".{key}.{os.getpid()}.tmp")
            with open(tmp, "wb") as f:
                f.write(data)
            os.replace(tmp, os.path.join(d, key + ".cubin"))
            break
        except OSError:
            continue
    return data


def ensure_compiled(plans) -> None:
    """Compile the structures of ``plans`` that are in no cache yet, in parallel (NVRTC programs
    are independent and the ctypes call releases the GIL): the 13 kernels of the 30-qubit ansatz
    take about 3 s instead of 20 s on the first call of a process with a cold disk cache."""
    todo, seen = [], set()
    for p in plans:
        k = _full_key(p.key)
        if k in seen or k in _CUBINS or k in _KERNELS:
            continue
        seen.add(k)
        todo.append(p)
    if len(todo) <= 1:
        for p in todo:
            compile_plan(p)
        return
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4, 16)) as pool:
        list(pool.map(compile_plan, todo))


def kernel_for(plan: SegPlan):
    """Loaded kernel handle for the plan's structure (per process and CUDA context)."""
    key = _full_key(plan.key)
    with _LOCK:
        h = _KERNELS.get(key)
        if h is not None:
            _STATS["mem_hits"] += 1
            return h
    data = compile_plan(plan)
    lib = load()
    handle = C.c_void_p()
    buf = C.create_string_buffer(data, len(data))
    check(lib.b200q_seg_load(buf, len(data), C.byref(handle)))
    with _LOCK:
        _KERNELS[key] = handle
    return handle


def stats() -> dict:
    return dict(_STATS, kernels=len(_KERNELS))


def launch(plan: SegPlan, coefs: np.ndarray, vec0_ptr, vec1_ptr, n: int, batch: int, work, work_bytes,
           stream, base_hi: int = 0, write0: int = 1, scale: float = 1.0, out_ptr=None,
           fix_mask: int = 0, fix_val: int = 0):
    """One segment launch (``b200q_seg_launch``).  ``fix_mask`` / ``fix_val``: partial launch over
    the tiles whose (non-tile) index bits ``fix_mask`` equal ``fix_val``."""
    lib = load()
    g = plan.geom
    h = kernel_for(plan)
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    batched = coefs.ndim == 2
    if batched and coefs.shape[0] != batch:
        raise ValueError(f"coefficient tables for batch {coefs.shape[0]} on a state of batch {batch}")
    check(lib.b200q_seg_launch(
        h, vec0_ptr, vec1_ptr, n, g.dtype_code, batch, int_array(plan.tile_bits), g.T, plan.L, g.RB,
        g.MINB, int_array(plan.ext_pos) if plan.ext_pos else None, len(plan.ext_pos),
        coefs.ctypes.data_as(C.POINTER(C.c_double)), int(coefs.shape[-1]),
        (2 if plan.coef_param else 1 if batched else 0),
        plan.nslots, write0, int(base_hi), int(fix_mask), int(fix_val), float(scale), out_ptr, work,
        work_bytes, stream))
