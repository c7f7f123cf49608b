"""Test-only numpy interpreter of the register-tiled kernel's record stream
(pennylane_b200/csrc/rtile.cuh ``RtOp``).  It maps every record's register / thread / external
masks back to global bit positions through the current ROUND layout and applies the record to a
flat numpy state, so the host-side scheduler and encoder (compiler.schedule_rounds /
encode_rt_segment) can be checked on CPU against the oracle without a GPU.  Not product code."""
import numpy as np

from pennylane_b200 import compiler as cc


def _masks(mr, mt, me, rglob, tglob):
    m = int(me)
    for b, g in enumerate(rglob):
        if (mr >> b) & 1:
            m |= 1 << g
    for b, g in enumerate(tglob):
        if (mt >> b) & 1:
            m |= 1 << g
    return m


def run_records(state, n, tile_bits, ops_arr, table, RB, bra=None, nslots=0, base_hi=0):
    """Apply the records to ``state`` (flat, length 2^n; modified copy returned).  With ``bra``
    also returns the per-slot sums of coef * Im<bra|P|ket>.  ``base_hi`` is OR-ed into every
    index for external predicates (rank bits of a sharded state)."""
    T = len(tile_bits)
    TB = T - RB
    vecs = [np.array(state, dtype=complex).reshape(-1)]
    if bra is not None:
        vecs.append(np.array(bra, dtype=complex).reshape(-1))
    idx = np.arange(1 << n, dtype=np.uint64) | np.uint64(base_hi)
    sums = np.zeros(nslots)
    rglob = tglob = None
    assert ops_arr[0].kind == cc.RT_ROUND

    def par(mask):
        x = idx & np.uint64(mask)
        p = np.zeros(idx.shape, dtype=np.uint64)
        while x.any():
            p ^= x & np.uint64(1)
            x >>= np.uint64(1)
        return p.astype(bool)

    for o in ops_arr:
        k = o.kind & 0xff
        has0 = (o.kind >> 8) & 1
        if k == cc.RT_ROUND:
            rpos = [o.u.r.rbits[b] for b in range(RB)]
            tpos = [o.u.r.tbits[b] for b in range(TB)]
            assert sorted(rpos + tpos) == list(range(T)), "round is not a permutation"
            rglob = [tile_bits[p] for p in rpos]
            tglob = [tile_bits[p] for p in tpos]
            continue
        if k == cc.RT_GEN:
            p = o.u.p
            assert p.zt >> TB == 0 and p.xr >> RB == 0
            xm = _masks(p.xr, 0, 0, rglob, tglob)
            zm = _masks(p.zr, p.zt, p.ze, rglob, tglob)
            ket, b = vecs[0], vecs[1]
            j = (idx ^ np.uint64(xm)) & np.uint64((1 << n) - 1)
            # (P ket)_i = i^ny (-1)^{popc((i^x) & z)} ket_{i^x}
            sign = np.where(_par_of(idx ^ np.uint64(xm), zm), -1.0, 1.0)
            pk = (1j ** o.q1) * sign * ket[j.astype(np.int64)]
            sums[o.q0] += p.coef * np.imag(np.vdot(b, pk))
            continue
        if k == cc.DIAG:
            nd = o.q0
            tabi = np.zeros(idx.shape, dtype=np.int64)
            for b in range(nd):
                s = o.u.d.src[b]
                g = rglob[s] if s < 32 else (tglob[s - 32] if s < 64 else s - 64)
                tabi |= (((idx >> np.uint64(g)) & np.uint64(1)).astype(np.int64)) << (nd - 1 - b)
            d = table[o.mat_off: o.mat_off + (1 << nd)][tabi]
            for v in vecs:
                v *= d
            continue
        g = o.u.g
        cm = _masks(g.ctrl_r, g.ctrl_t, g.ctrl_e, rglob, tglob)
        cv = _masks(g.cval_r, g.cval_t, g.cval_e, rglob, tglob)
        ok = (idx & np.uint64(cm)) == np.uint64(cv)
        if k == cc.PARITY:
            pm = _masks(g.par_r, g.par_t, g.par_e, rglob, tglob)
            ph = np.where(par(pm), table[o.mat_off + 1], table[o.mat_off])
            for v in vecs:
                v[ok] *= ph[ok]
            continue
        low = idx & np.uint64((1 << n) - 1)
        if k in (cc.DENSE1, cc.CX):
            t = rglob[o.q0]
            variants = [(ok, 0)] + ([(~ok, 4)] if has0 else [])
            for okv, moff in variants:
                sel = okv & (((idx >> np.uint64(t)) & np.uint64(1)) == 0)
                i0 = low[sel].astype(np.int64)
                i1 = i0 | (1 << t)
                m = (np.array([[0, 1], [1, 0]], dtype=complex) if k == cc.CX
                     else table[o.mat_off + moff: o.mat_off + moff + 4].reshape(2, 2))
                for v in vecs:
                    x0, x1 = v[i0].copy(), v[i1].copy()
                    v[i0] = m[0, 0] * x0 + m[0, 1] * x1
                    v[i1] = m[1, 0] * x0 + m[1, 1] * x1
            continue
        if k == cc.DENSE2:
            assert o.q0 > o.q1
            t0, t1 = rglob[o.q0], rglob[o.q1]
            sel = ok & (((idx >> np.uint64(t0)) & np.uint64(1)) == 0) & \
                (((idx >> np.uint64(t1)) & np.uint64(1)) == 0)
            i0 = low[sel].astype(np.int64)
            ii = [i0, i0 | (1 << t1), i0 | (1 << t0), i0 | (1 << t0) | (1 << t1)]
            m = table[o.mat_off: o.mat_off + 16].reshape(4, 4)
            for v in vecs:
                x = [v[i].copy() for i in ii]
                for r in range(4):
                    v[ii[r]] = sum(m[r, c] * x[c] for c in range(4))
            continue
        raise AssertionError(f"unknown record kind {k}")
    if bra is not None:
        return vecs[0], vecs[1], sums
    return vecs[0]


def _par_of(x, mask):
    x = x & np.uint64(mask)
    p = np.zeros(x.shape, dtype=np.uint64)
    while x.any():
        p ^= x & np.uint64(1)
        x = x >> np.uint64(1)
    return p.astype(bool)
