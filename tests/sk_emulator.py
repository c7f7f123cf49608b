"""Test-only numpy interpreter of the specialised segment kernel's plan (pennylane_b200/segjit.py
``SegPlan.ir`` + coefficient table).  It resolves every record's register / thread / external
locations back to global bit positions through the current round layout — the same mapping
csrc/segk.cuh compiles in — and applies the record to a flat numpy state, so the host side
(round scheduling, normalised forms, segment scalars, generator corrections, coefficient order)
is checked on CPU against the oracle.  Not product code."""
import numpy as np


def _bit(idx, g):
    return ((idx >> np.uint64(g)) & np.uint64(1)).astype(np.int64)


class _Layout:
    def __init__(self, plan, ri):
        rpos, tpos = plan.rounds[ri]
        self.rglob = [plan.tile_bits[p] for p in rpos]
        self.tglob = [plan.tile_bits[p] for p in tpos]
        self.ext = plan.ext_pos

    def reg_mask_ok(self, idx, mr, vr):
        ok = np.ones(idx.shape, dtype=bool)
        for b, g in enumerate(self.rglob):
            if (mr >> b) & 1:
                ok &= _bit(idx, g) == ((vr >> b) & 1)
        return ok

    def pred(self, idx, pred):
        if pred is None:
            return np.ones(idx.shape, dtype=bool)
        mt, vt, me, ve = pred
        ok = np.ones(idx.shape, dtype=bool)
        for b, g in enumerate(self.tglob):
            if (mt >> b) & 1:
                ok &= _bit(idx, g) == ((vt >> b) & 1)
        for e, g in enumerate(self.ext):
            if (me >> e) & 1:
                ok &= _bit(idx, g) == ((ve >> e) & 1)
        return ok

    def parity(self, idx, pr, pt, pe, const=0):
        p = np.full(idx.shape, const & 1, dtype=np.int64)
        for b, g in enumerate(self.rglob):
            if (pr >> b) & 1:
                p ^= _bit(idx, g)
        for b, g in enumerate(self.tglob):
            if (pt >> b) & 1:
                p ^= _bit(idx, g)
        for e, g in enumerate(self.ext):
            if (pe >> e) & 1:
                p ^= _bit(idx, g)
        return p


def _cplx(tab, off):
    return tab[off] + 1j * tab[off + 1]


def run_plan(plan, coefs, state, n, bra=None, base_hi=0):
    """Apply the plan to ``state`` (flat, 2^n).  With ``bra``: returns (ket, bra, slot sums)."""
    tab = np.asarray(coefs, dtype=float)
    vecs = [np.array(state, dtype=complex).reshape(-1)]
    if bra is not None:
        vecs.append(np.array(bra, dtype=complex).reshape(-1))
    idx = np.arange(1 << n, dtype=np.uint64) | np.uint64(base_hi)
    low = (idx & np.uint64((1 << n) - 1)).astype(np.int64)
    sums = np.zeros(plan.nslots)
    lay = None
    nfetch = 0
    for rec in plan.ir:
        k = rec[0]
        if k == "load":
            lay = _Layout(plan, rec[1])
        elif k == "fetch":
            nfetch += 1
        elif k == "xpose":
            lay = _Layout(plan, rec[2])
        elif k == "dk":
            _, q, kern, dl, dr, off = rec
            t, sinp = tab[off], tab[off + 1] != 0.0
            o = off + 2
            r = l = 1.0
            if dr:
                r = _cplx(tab, o); o += 2
            if dl:
                l = _cplx(tab, o)
            if kern == 0:
                K = np.array([[t, -1.0], [1.0, t]]) if sinp else np.array([[1.0, -t], [t, 1.0]])
            else:
                K = np.array([[t, -1j], [-1j, t]]) if sinp else np.array([[1.0, -1j * t], [-1j * t, 1.0]])
            m = np.diag([1.0, l]) @ K @ np.diag([1.0, r])
            g = lay.rglob[q]
            i0 = low[_bit(idx, g) == 0]
            i1 = i0 | (1 << g)
            for v in vecs:
                x0, x1 = v[i0].copy(), v[i1].copy()
                v[i0] = m[0, 0] * x0 + m[0, 1] * x1
                v[i1] = m[1, 0] * x0 + m[1, 1] * x1
        elif k == "f16":
            _, q, mr, vr, has0, off, pred = rec
            g = lay.rglob[q]
            ok = lay.reg_mask_ok(idx, mr, vr) & lay.pred(idx, pred)
            m1 = (tab[off: off + 8: 2] + 1j * tab[off + 1: off + 8: 2]).reshape(2, 2)
            variants = [(ok, m1)]
            if has0:
                m0 = (tab[off + 8: off + 16: 2] + 1j * tab[off + 9: off + 16: 2]).reshape(2, 2)
                variants.append((~ok, m0))
            for okv, m in variants:
                i0 = low[okv & (_bit(idx, g) == 0)]
                i1 = i0 | (1 << g)
                for v in vecs:
                    x0, x1 = v[i0].copy(), v[i1].copy()
                    v[i0] = m[0, 0] * x0 + m[0, 1] * x1
                    v[i1] = m[1, 0] * x0 + m[1, 1] * x1
        elif k == "d2":
            _, q0, q1, mr, vr, off, pred = rec
            assert q0 > q1
            g0, g1 = lay.rglob[q0], lay.rglob[q1]
            ok = lay.reg_mask_ok(idx, mr, vr) & lay.pred(idx, pred)
            m = (tab[off: off + 32: 2] + 1j * tab[off + 1: off + 32: 2]).reshape(4, 4)
            i0 = low[ok & (_bit(idx, g0) == 0) & (_bit(idx, g1) == 0)]
            ii = [i0, i0 | (1 << g1), i0 | (1 << g0), i0 | (1 << g0) | (1 << g1)]
            for v in vecs:
                x = [v[i].copy() for i in ii]
                for r_ in range(4):
                    v[ii[r_]] = sum(m[r_, c] * x[c] for c in range(4))
        elif k == "cx":
            _, q, mr, vr, pred = rec
            g = lay.rglob[q]
            ok = lay.reg_mask_ok(idx, mr, vr) & lay.pred(idx, pred)
            i0 = low[ok & (_bit(idx, g) == 0)]
            i1 = i0 | (1 << g)
            for v in vecs:
                x0 = v[i0].copy()
                v[i0] = v[i1]
                v[i1] = x0
        elif k == "par":
            _, mr, vr, pr, (pt, pe), norm, off, pred = rec
            ok = lay.reg_mask_ok(idx, mr, vr) & lay.pred(idx, pred)
            par = lay.parity(idx, pr, pt, pe)
            if norm:
                m0, m1 = 1.0, _cplx(tab, off)
            else:
                m0, m1 = _cplx(tab, off), _cplx(tab, off + 2)
            ph = np.where(par == 1, m1, m0)
            for v in vecs:
                v[ok] *= ph[ok]
        elif k == "diag":
            _, rc, off, items, nd = rec
            ti = np.zeros(idx.shape, dtype=np.int64)
            for b, g in enumerate(lay.rglob):
                ti |= _bit(idx, g) * rc[b]
            for kind, i, w in items:
                g = lay.tglob[i] if kind == "t" else lay.ext[i]
                ti |= _bit(idx, g) << w
            d = tab[off + 2 * ti] + 1j * tab[off + 2 * ti + 1]
            for v in vecs:
                v *= d
        elif k == "gen":
            _, xr, zr, odd, slot, off, (zt, ze, c) = rec
            xm = 0
            for b, g in enumerate(lay.rglob):
                if (xr >> b) & 1:
                    xm |= 1 << g
            ket, b_ = vecs[0], vecs[1]
            j = low ^ xm
            sign_par = _Layout.parity(lay, idx ^ np.uint64(xm), zr, zt, ze, c)
            # note: thread / external bits are not flipped by xm (x part lives on register bits)
            sign = np.where(sign_par == 1, -1.0, 1.0)
            prod = np.conj(b_) * ket[j]
            val = prod.real if odd else prod.imag
            sums[slot] += tab[off] * np.sum(sign * val)
        elif k == "scale":
            s = _cplx(tab, rec[1])
            for v in vecs:
                v *= s
        else:
            raise AssertionError(k)
    assert nfetch == 1, "the body must prefetch the next tile exactly once"
    if bra is not None:
        return vecs[0], vecs[1], sums
    return vecs[0]
