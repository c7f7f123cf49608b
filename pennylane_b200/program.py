"""Structure-keyed cache of fused programs with parameter rebinding.

The reference re-reads every operator of a tape on every execution (simulate.py:214-235) and
has no compiled form; its cache key for a circuit is ``QuantumScript.hash``
(pennylane/core/qscript.py:193), which includes the parameter values.  Here the host fusion
pass (lowering, block merging, segment packing, round scheduling, kernel generation) depends on
the circuit's STRUCTURE only — operator names, wires, hyper-parameters — so its result is cached
under a key built from exactly those, and a call with new parameter values only

  1. builds the 2x2 matrices of the single-qubit operators, vectorised per gate type,
  2. multiplies them along the merged blocks (batched matmul per chain length),
  3. brings every block to its normalised form (vectorised :func:`canon_batch`), and
  4. scatters the coefficients into the per-segment tables,

about 2 ms for the 720-gate ansatz instead of ~130 ms for the full pass.  Primitives whose values
do not come from a product of single-qubit operator matrices (controlled blocks, 4x4 blocks,
parity phases of multi-qubit gates, diagonal tables) are re-lowered individually.

The cache is validated, not trusted: if the new values no longer fit the compiled structure (a
block that happened to be diagonal at the first call and is not any more, a normalised form that
needs a phase the kernel does not have) the entry is dropped and the circuit recompiled.
"""
from __future__ import annotations

import collections

import numpy as np

from . import compiler as cc
from . import segjit

_MAX_ENTRIES = 64
_CACHE: "collections.OrderedDict" = collections.OrderedDict()
_SEEN: "collections.OrderedDict" = collections.OrderedDict()
STATS = {"hits": 0, "misses": 0, "rebind_failures": 0, "compile_calls": 0}


def _hashable(x):
    if isinstance(x, dict):
        return tuple(sorted((k, _hashable(v)) for k, v in x.items()))
    if isinstance(x, (list, tuple)):
        return tuple(_hashable(v) for v in x)
    if isinstance(x, np.ndarray):
        return (x.shape, x.dtype.str, x.tobytes())
    try:
        hash(x)
        return x
    except TypeError:
        return repr(x)


_SCALARS = (float, int, np.float64, np.float32, np.int64)


class _Key:
    """A structure key with its hash computed once (a tuple of a few hundred tuples is re-hashed by
    every dictionary operation: 0.1 ms each at 320 operators, several per execution)."""
    __slots__ = ("t", "h")

    def __init__(self, t):
        self.t = t
        self.h = hash(t)

    def __hash__(self):
        return self.h

    def __eq__(self, other):
        return isinstance(other, _Key) and self.h == other.h and self.t == other.t

    def __repr__(self):
        return f"_Key({self.t!r})"


def _shapes(data):
    return tuple(() if type(d) in _SCALARS else np.shape(d) for d in data)


def structure_key(ops_, extra=()) -> tuple:
    """Names, wires, hyper-parameters and parameter SHAPES of the operators — everything of
    ``QuantumScript.hash`` (qscript.py:193) except the parameter values."""
    return _structure_key(ops_, extra)[0]


def _structure_key(ops_, extra=()):
    """(key, all parameters are scalars).  One pass over the operators: this runs on every
    execution of a cached circuit (0.6 ms for 320 operators), so the common case — an operator
    without hyper-parameters, control values or a base — takes the short route."""
    items = []
    add = items.append
    scalar = True
    for op in ops_:
        data = op.data
        shapes = _shapes(data)
        if scalar and any(shapes):
            scalar = False
        hyper = getattr(op, "hyperparameters", None)
        cv = getattr(op, "control_values", None)
        base = getattr(op, "base", None)
        if not hyper and cv is None and base is None:
            add((op.name, tuple(op.wires), shapes))
        else:
            add((op.name, tuple(op.wires), _hashable(hyper) if hyper else (), shapes,
                 _hashable(cv) if cv is not None else None,
                 structure_key([base]).t if base is not None else None,
                 # operators whose matrix is their datum (QubitUnitary ...) keep the structure
                 # of that matrix out of the key: it is a value
                 ))
    return _Key((tuple(items),) + tuple(extra)), scalar


# ---------------------------------------------------------------------------------------------
# vectorised single-qubit matrices
# ---------------------------------------------------------------------------------------------
def _rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    m = np.zeros(t.shape + (2, 2), dtype=complex)
    m[..., 0, 0] = c; m[..., 1, 1] = c
    m[..., 0, 1] = -1j * s; m[..., 1, 0] = -1j * s
    return m


def _ry(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    m = np.zeros(t.shape + (2, 2), dtype=complex)
    m[..., 0, 0] = c; m[..., 1, 1] = c
    m[..., 0, 1] = -s; m[..., 1, 0] = s
    return m


def _rz(t):
    m = np.zeros(t.shape + (2, 2), dtype=complex)
    m[..., 0, 0] = np.exp(-0.5j * t); m[..., 1, 1] = np.exp(0.5j * t)
    return m


def _phase(t):
    m = np.zeros(t.shape + (2, 2), dtype=complex)
    m[..., 0, 0] = 1.0; m[..., 1, 1] = np.exp(1j * t)
    return m


_VECTOR_GATES = {"RX": _rx, "RY": _ry, "RZ": _rz, "PhaseShift": _phase, "U1": _phase}


def op_matrices(ops_, idx) -> np.ndarray:
    """(len(idx), 2, 2) matrices of the single-qubit operators ``ops_[i] for i in idx``: one
    vectorised evaluation per gate type, ``op.matrix()`` for the rest.  For an adjoint sweep the
    list holds the adjoint operators already."""
    out = np.empty((len(idx), 2, 2), dtype=complex)
    groups: dict = {}
    for k, i in enumerate(idx):
        op = ops_[i]
        f = _VECTOR_GATES.get(op.name)
        if f is not None and len(op.data) == 1 and np.ndim(op.data[0]) == 0:
            groups.setdefault(op.name, ([], []))
            groups[op.name][0].append(k)
            groups[op.name][1].append(float(op.data[0]))
        else:
            out[k] = np.asarray(op.matrix(), dtype=complex)
    for name, (ks, ts) in groups.items():
        out[np.asarray(ks)] = _VECTOR_GATES[name](np.asarray(ts))
    return out


# ---------------------------------------------------------------------------------------------
# vectorised normalised forms
# ---------------------------------------------------------------------------------------------
def canon_batch(u: np.ndarray, kern: np.ndarray, dl: np.ndarray, dr: np.ndarray):
    """Normalised forms of the unitaries ``u[k]`` under the per-block structure (kern, dl, dr)
    compiled into the kernels: returns (t, sinp, r, l, s, ok).  ``ok[k]`` is False when block k
    does not fit its structure (the caller recompiles).  Same conventions as
    :func:`segjit.canon_1q` (``U = s * diag(1,l) K(t) diag(1,r)``)."""
    u00, u01, u10, u11 = u[:, 0, 0], u[:, 0, 1], u[:, 1, 0], u[:, 1, 1]
    k = u.shape[0]
    sinp = np.abs(u00) < np.abs(u10)
    ph = np.where(kern == 0, 1.0 + 0j, 1j)              # K's off-diagonal: -+t (real) / -it
    sg = np.where(kern == 0, -1.0 + 0j, 1j)             # r relation: r t = sg * u01 / s
    with np.errstate(divide="ignore", invalid="ignore"):
        # ---- cos pivot: s = u00,  l t = ph u10 / s,  r t = sg u01 / s
        A = ph * u10 / u00
        Bc = sg * u01 / u00
        t_c = np.where(dl & dr, np.abs(A), np.where(dl, Bc.real, A.real))
        tz = t_c == 0
        safe_t = np.where(tz, 1.0, t_c)
        l_c = np.where(dl, np.where(tz, u11 / u00, A / safe_t), 1.0 + 0j)
        r_c = np.where(dr & ~tz, Bc / safe_t, 1.0 + 0j)
        s_c = u00
        # ---- sin pivot: s t = u00,  l = ph u10 / s,  r = sg u01 / s
        both = dl & dr
        s_free = u00 / np.where(u00 == 0, 1.0, np.abs(u00) / np.abs(u10))     # sigma = +1
        s_l1 = ph * u10                                                         # l = 1
        s_r1 = sg * u01                                                         # r = 1
        zero = u00 == 0
        s_s = np.where(both & ~zero, s_free, np.where(dl, s_r1, s_l1))
        t_s = np.where(zero, 0.0, (u00 / s_s).real)
        l_s = np.where(dl, ph * u10 / s_s, 1.0 + 0j)
        r_s = np.where(dr, sg * u01 / s_s, 1.0 + 0j)
    t = np.where(sinp, t_s, t_c)
    l = np.where(sinp, l_s, l_c)
    r = np.where(sinp, r_s, r_c)
    s = np.where(sinp, s_s, s_c)
    # reconstruction check
    K = np.empty((k, 2, 2), dtype=complex)
    off = np.where(kern == 0, 1.0 + 0j, 1j)
    K[:, 0, 0] = np.where(sinp, t, 1.0); K[:, 1, 1] = K[:, 0, 0]
    K[:, 0, 1] = -off * np.where(sinp, 1.0, t)
    K[:, 1, 0] = np.where(kern == 0, 1.0, -1j) * np.where(sinp, 1.0, t)
    rec = np.empty_like(K)
    rec[:, 0, 0] = s * K[:, 0, 0]
    rec[:, 0, 1] = s * K[:, 0, 1] * r
    rec[:, 1, 0] = s * l * K[:, 1, 0]
    rec[:, 1, 1] = s * l * K[:, 1, 1] * r
    err = np.max(np.abs(rec - u).reshape(k, -1), axis=1)
    ok = np.isfinite(err) & (err < 1e-13) & (np.abs(t) <= 1.0 + 1e-12)
    return t, sinp, r, l, s, ok


# ---------------------------------------------------------------------------------------------
# cached program
# ---------------------------------------------------------------------------------------------
class RebindError(Exception):
    """The new values do not fit the cached structure."""


class FusedProgram:
    """Segments + plans of one circuit structure, and the recipe that turns the parameter values
    of a structurally identical operator list into the segments' coefficient tables."""

    def __init__(self, segs, plans, n_ops):
        self.segs = segs
        self.plans = plans            # plans[i] is None for generic (per-gate) segments
        self.n_ops = n_ops
        self.tables = None            # coefficient tables of the last binding
        self._build_recipe()

    def _build_recipe(self):
        blocks = []                   # (segment, fill entry index) of vectorisable DK entries
        chains = []
        self.slow = []                # (segment, prim index) that must be re-lowered
        self.slow_dk = []             # DK entries without exact provenance
        for si, (seg, plan) in enumerate(zip(self.segs, self.plans)):
            if plan is None:
                for pi, p in enumerate(seg.prims):
                    self.slow.append((si, pi))
                continue
            valued = set()
            for off, nreal, kind, pi, extra in plan.fill:
                if kind == "dk":
                    p = seg.prims[pi]
                    if p.src is not None and p.src_exact and np.ndim(p.mat) == 2:
                        blocks.append((si, pi, off, extra))
                        chains.append(list(p.src))
                    else:
                        valued.add(pi)
                elif kind in ("mat", "par"):
                    valued.add(pi)
            for pi in sorted(valued):
                self.slow.append((si, pi))
        self.blocks = blocks
        self.chain_groups = {}
        for b, ch in enumerate(chains):
            self.chain_groups.setdefault(len(ch), ([], []))
            self.chain_groups[len(ch)][0].append(b)
            self.chain_groups[len(ch)][1].append(ch)
        self.used_ops = sorted({i for ch in chains for i in ch})
        self.op_pos = {i: k for k, i in enumerate(self.used_ops)}
        self.chain_groups = {ln: (np.asarray(bs), np.asarray([[self.op_pos[i] for i in ch] for ch in chs]))
                             for ln, (bs, chs) in self.chain_groups.items()}
        nb = len(blocks)
        self.b_kern = np.array([b[3][0] for b in blocks], dtype=int).reshape(nb)
        self.b_dl = np.array([b[3][1] for b in blocks], dtype=bool).reshape(nb)
        self.b_dr = np.array([b[3][2] for b in blocks], dtype=bool).reshape(nb)
        # every slow primitive must know which operator(s) to re-lower
        for si, pi in self.slow:
            p = self.segs[si].prims[pi]
            if p.src is None:
                raise RebindError("primitive without provenance")

    def _used_matrices(self, ops_):
        """:func:`op_matrices` of ``self.used_ops`` with the grouping by gate type done ONCE per
        program: names and parameter shapes are part of the structure key, so the groups of the first
        binding hold for every later one and a rebinding only gathers the angles."""
        groups = getattr(self, "_mat_groups", None)
        if groups is None:
            by_name: dict = {}
            slow = []
            for k, i in enumerate(self.used_ops):
                op = ops_[i]
                if op.name in _VECTOR_GATES and len(op.data) == 1 and np.ndim(op.data[0]) == 0:
                    by_name.setdefault(op.name, ([], []))
                    by_name[op.name][0].append(k)
                    by_name[op.name][1].append(i)
                else:
                    slow.append((k, i))
            groups = self._mat_groups = ([(name, np.asarray(ks), idx) for name, (ks, idx) in by_name.items()], slow)
        out = np.empty((len(self.used_ops), 2, 2), dtype=complex)
        for name, ks, idx in groups[0]:
            ts = np.fromiter((ops_[i].data[0] for i in idx), dtype=float, count=len(idx))
            out[ks] = _VECTOR_GATES[name](ts)
        for k, i in groups[1]:
            out[k] = np.asarray(ops_[i].matrix(), dtype=complex)
        return out

    def bind(self, ops_, bit_of, batched_ok, adjoint: bool = False):
        """Coefficient tables (and refreshed primitive values) for ``ops_``.  ``adjoint``: the
        program applies the ADJOINT of every operator of ``ops_`` (a reverse sweep; ``ops_`` in
        sweep order)."""
        if adjoint:
            from .adjoint import _op_adjoint
            lower_op = lambda o: cc.lower(_op_adjoint(o), bit_of, batched_ok)     # noqa: E731
        else:
            lower_op = lambda o: cc.lower(o, bit_of, batched_ok)                   # noqa: E731
        if len(ops_) != self.n_ops:
            raise RebindError("operator count changed")
        segs, plans = self.segs, self.plans
        # 1-2. block matrices
        nb = len(self.blocks)
        if nb:
            M = self._used_matrices(ops_)
            if adjoint:
                M = np.conj(np.swapaxes(M, 1, 2))
            U = np.empty((nb, 2, 2), dtype=complex)
            for ln, (bs, ch) in self.chain_groups.items():
                acc = M[ch[:, 0]]
                for j in range(1, ln):
                    acc = M[ch[:, j]] @ acc
                U[bs] = acc
            t, sinp, r, l, s, ok = canon_batch(U, self.b_kern, self.b_dl, self.b_dr)
            if not np.all(ok):
                raise RebindError("a block does not fit its normalised form")
            for b, (si, pi, off, form) in enumerate(self.blocks):
                segs[si].prims[pi].mat = U[b]
        # slow primitives: re-lower their operator(s)
        for si, pi in self.slow:
            p = segs[si].prims[pi]
            if len(p.src) == 1:
                low = lower_op(ops_[p.src[0]])
                q = low[p.src_j] if p.src_j < len(low) else None
                if q is not None and q.kind == p.kind and q.targets == p.targets and q.ctrl == p.ctrl \
                        and q.other == p.other and np.shape(q.mat) == np.shape(p.mat):
                    p.mat, p.mat0, p.op = q.mat, q.mat0, q.op
                    continue
            # a run of single-qubit primitives multiplied (and re-classified) by merge_blocks
            m = None
            for i in p.src:
                low = lower_op(ops_[i])
                one = cc._as_1q_matrix(low[0]) if len(low) == 1 else None
                if one is None:
                    raise RebindError("lowering changed")
                mm = np.asarray(one[1], dtype=complex)
                m = mm if m is None else mm @ m
            bit = (p.targets or p.other or list(p.ctrl) or [None])[0]
            if bit is None:
                if np.allclose(m, np.eye(2)):
                    continue                              # identity block (kept as a unit phase)
                raise RebindError("a merged block changed kind")
            q = cc._dense([bit], {}, m)
            if q.kind == p.kind and q.ctrl == p.ctrl and list(q.other) == list(p.other) \
                    and np.shape(q.mat) == np.shape(p.mat):
                p.mat = q.mat
            elif p.kind == cc.DENSE1 and not p.ctrl:
                p.mat = m
            else:
                raise RebindError("a merged block changed kind")
        # 3-4. tables
        tables = []
        bi = 0
        block_of = {}
        for b, (si, pi, off, form) in enumerate(self.blocks):
            block_of[(si, pi)] = b
        for si, (seg, plan) in enumerate(zip(segs, plans)):
            if plan is None:
                tables.append(None)
                continue
            try:
                if not any((si, e[3]) in block_of for e in plan.fill if e[2] == "dk"):
                    tables.append(segjit.coefficients(plan, seg.prims))
                    continue
                tab = np.zeros(plan.ncoef)
                scalars = {}
                for off, nreal, kind, pi, extra in plan.fill:
                    if kind == "dk":
                        b = block_of.get((si, pi))
                        if b is None:
                            c = segjit._canon_for(np.asarray(seg.prims[pi].mat, dtype=complex), extra)
                            vals = [c.t, 1.0 if c.sinp else 0.0]
                            if extra[2]:
                                vals += [c.r.real, c.r.imag]
                            if extra[1]:
                                vals += [c.l.real, c.l.imag]
                            tab[off: off + nreal] = vals
                            scalars[pi] = c.s
                        else:
                            tab[off] = t[b]
                            tab[off + 1] = 1.0 if sinp[b] else 0.0
                            o = off + 2
                            if extra[2]:
                                tab[o] = r[b].real; tab[o + 1] = r[b].imag; o += 2
                            if extra[1]:
                                tab[o] = l[b].real; tab[o + 1] = l[b].imag
                            scalars[pi] = s[b]
                    elif kind == "par" and extra:
                        m = np.asarray(seg.prims[pi].mat, dtype=complex).reshape(-1)
                        if abs(abs(m[0]) - 1.0) >= 1e-12:
                            raise segjit.FormMismatch()
                        scalars[pi] = complex(m[0])
                for off, nreal, kind, pi, extra in plan.fill:
                    if kind == "mat":
                        p = seg.prims[pi]
                        if extra == "mat":
                            vals = segjit._complex_pairs(p.mat)
                        elif extra == "mat2s":
                            vals = segjit._complex_pairs(cc._swap_2q(p.mat))
                        else:
                            vals = np.concatenate([segjit._complex_pairs(p.mat), segjit._complex_pairs(p.mat0)])
                    elif kind == "par":
                        m = np.asarray(seg.prims[pi].mat, dtype=complex).reshape(-1)[:2]
                        if extra:
                            qv = m[1] / m[0]
                            vals = [qv.real, qv.imag]
                        else:
                            vals = [m[0].real, m[0].imag, m[1].real, m[1].imag]
                    elif kind == "gen":
                        s2 = 1.0
                        for j in extra:
                            s2 *= abs(scalars[j]) ** 2
                        vals = [float(seg.prims[pi].coef) * s2, 0.0]
                    elif kind == "scale":
                        sc = 1.0 + 0j
                        for j in extra:
                            sc *= scalars[j]
                        vals = [sc.real, sc.imag]
                    else:
                        continue
                    if len(vals) != nreal:
                        raise segjit.FormMismatch()
                    tab[off: off + nreal] = np.asarray(vals, dtype=float)
                tables.append(tab)
            except segjit.FormMismatch as e:
                raise RebindError("normalised form changed") from e
        self.tables = tables
        return tables


def get_program(sv, ops_, level, T=None, L=None, bit_of=None):
    """Cached :class:`FusedProgram` for ``ops_`` on the state ``sv`` with its tables bound to the
    current parameter values.  Returns (program, hit)."""
    import os

    n = sv.n
    dT, dL = sv.default_tile()
    T = T or dT
    if L is None:                                    # programs always run on the specialised kernels
        L = min(segjit.default_low_bits(sv.dtype_code, 1), T)
    if bit_of is None:
        bit_of_f = lambda w: n - 1 - int(w)          # noqa: E731
        bkey = None
    else:
        bit_of_f = bit_of
        bkey = tuple(bit_of(w) for op in ops_ for w in op.wires)
    key, scalar = _structure_key(ops_, (n, sv.dtype_code, int(level), T, L, bkey,
                                        os.environ.get("B200Q_ROUND_BUDGET"), os.environ.get("B200Q_IO_LANES"),
                                        os.environ.get("B200Q_SK_FWD")))
    if not scalar and any(getattr(op, "batch_size", None) is not None for op in ops_):
        return None, False                           # broadcast parameters: uncached path
    hot = sv.jit_enabled(1)
    if not hot and not sv.jit_possible(1):
        return None, False                           # the interpreter path keeps its own encodings
    if not hot:
        # below the size threshold a structure is compiled the SECOND time it is seen
        if key not in _CACHE and key not in _SEEN:
            _SEEN[key] = True
            while len(_SEEN) > 4 * _MAX_ENTRIES:
                _SEEN.popitem(last=False)
            return None, False
    with sv.hot():
        return _get_program_hot(sv, ops_, level, T, L, bit_of, bit_of_f, key)


def _get_program_hot(sv, ops_, level, T, L, bit_of, bit_of_f, key):
    n = sv.n
    rtT = sv.rt_geometry(1)[0]
    if T != rtT:
        return None, False
    batched_ok = T == rtT and n >= rtT
    prog = _CACHE.get(key)
    if prog is not None:
        try:
            prog.bind(ops_, bit_of_f, batched_ok)
            _CACHE.move_to_end(key)
            STATS["hits"] += 1
            return prog, True
        except RebindError:
            STATS["rebind_failures"] += 1
            del _CACHE[key]
    STATS["misses"] += 1
    STATS["compile_calls"] += 1
    segs = sv.compile_fused(ops_, level, T, L, bit_of)
    sv.prepare_segments(segs)
    plans = [getattr(s, "_sk_plan", None) for s in segs]
    try:
        prog = FusedProgram(segs, plans, len(ops_))
    except RebindError:
        return None, False
    prog.tables = [getattr(s, "_sk_coefs", None) for s in segs]
    _CACHE[key] = prog
    while len(_CACHE) > _MAX_ENTRIES:
        _CACHE.popitem(last=False)
    return prog, False


def seen_before(key) -> bool:
    """True from the second call with ``key`` on (promotion of repeated structures)."""
    if key in _SEEN:
        _SEEN.move_to_end(key)
        return True
    _SEEN[key] = True
    while len(_SEEN) > 4 * _MAX_ENTRIES:
        _SEEN.popitem(last=False)
    return False


def cached(key) -> bool:
    return key in _CACHE


def lookup(key):
    prog = _CACHE.get(key)
    if prog is not None:
        _CACHE.move_to_end(key)
        STATS["hits"] += 1
    return prog


def store(key, prog):
    STATS["misses"] += 1
    _CACHE[key] = prog
    while len(_CACHE) > _MAX_ENTRIES:
        _CACHE.popitem(last=False)


def drop(key):
    STATS["rebind_failures"] += 1
    _CACHE.pop(key, None)


def clear():
    _CACHE.clear()
    _SEEN.clear()
