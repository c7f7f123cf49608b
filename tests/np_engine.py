"""TEST INFRASTRUCTURE: a numpy local engine for ``pennylane_b200.sharded.ShardedStateVector``.

It implements the engine interface of ``sharded.CudaEngine`` with the oracle's numpy kernels so
that the sharded HOST logic (planner, qubit maps, in-place exchanges, ordered reductions, the
distributed sampler) can run under ``gloo`` on CPU ranks.  Never imported by the product.
"""
import numpy as np
import torch

from oracle.apply_operation import apply_operation


class NumpyEngine:
    def __init__(self, nl, batch=1, dtype=np.complex128):
        self.n = nl
        self.np_dtype = np.dtype(dtype)
        self.data = torch.zeros((batch, 1 << nl),
                                dtype=torch.complex64 if self.np_dtype == np.complex64 else torch.complex128)
        self.applied = 0

    @property
    def batch(self):
        return self.data.shape[0]

    @property
    def device(self):
        return self.data.device

    def reset(self, index=None):
        self.data.zero_()
        if index is not None:
            self.data[:, index] = 1.0

    def set_local_state(self, arr):
        arr = np.asarray(arr, dtype=self.np_dtype).reshape(-1, 1 << self.n)
        self.data = torch.from_numpy(np.ascontiguousarray(arr).copy())

    def resize_batch(self, batch):
        if batch != self.batch:
            self.data = self.data.expand(batch, -1).contiguous()

    def compile(self, local_ops):
        return list(local_ops)

    def run(self, handle):
        for op in handle:
            bs = getattr(op, "batch_size", None)
            if bs is not None and self.batch == 1 and bs != 1:
                self.resize_batch(bs)
            batched = self.batch > 1
            st = self.data.numpy().reshape(((self.batch,) if batched else ()) + (2,) * self.n)
            out = apply_operation(op, st, is_state_batched=batched)
            # results are stored in the engine's precision (a complex64 shard drifts in norm by 1e-7)
            self.data = torch.from_numpy(np.ascontiguousarray(out, dtype=self.np_dtype).reshape(-1, 1 << self.n).copy())
            self.applied += 1
        return len(handle)

    # partial runs (the window schedule of sharded.ShardedStateVector._schedule): every operator
    # is a unit; it can run on the part of the state where index bits outside its wires are fixed
    def units(self, handle):
        if self.batch != 1 or not getattr(self, "divisible", True):
            return None
        out = []
        for op in handle:
            busy = 0
            for w in op.wires:
                busy |= 1 << (self.n - 1 - int(w))
            out.append((op, busy))
        return out

    def run_unit(self, op, fix_mask=0, fix_val=0):
        if not fix_mask:
            self.run([op])
            return 1
        n = self.n
        st = self.data.numpy().reshape((2,) * n)
        index, wmap, nxt = [], {}, 0
        for axis in range(n):
            b = n - 1 - axis
            if (fix_mask >> b) & 1:
                index.append((fix_val >> b) & 1)
            else:
                index.append(slice(None))
                wmap[axis] = nxt
                nxt += 1
        assert all(int(w) in wmap for w in op.wires), "a fixed bit is one of the operator's wires"
        sub = st[tuple(index)]
        sub[...] = apply_operation(op.map_wires(wmap), np.ascontiguousarray(sub))
        self.applied += 1
        self.partial_runs = getattr(self, "partial_runs", 0) + 1
        return 1

    # reductions
    def expval_terms(self, xs, zs, ys, cs):
        psi = self.data.numpy()
        idx = np.arange(1 << self.n, dtype=np.int64)
        out = np.zeros(self.batch)
        for xm, zm, ny, c in zip(xs, zs, ys, cs):
            par = np.zeros_like(idx)
            z = zm
            b = 0
            while z:
                if z & 1:
                    par ^= (idx >> b) & 1
                z >>= 1
                b += 1
            sign = 1.0 - 2.0 * par
            # P|j> = i^ny (-1)^popc(j & zm) |j ^ xm>   =>   <psi|P|psi> = sum_j conj(psi[j^xm]) ...
            val = (np.conj(psi[:, idx ^ xm]) * psi * sign).sum(axis=1) * (1j ** ny)
            out += c * np.real(val)
        return out

    def probs(self, local_wires):
        from oracle.measure import probs_process_state

        flat = self.data.numpy()
        full = list(local_wires) == list(range(self.n))
        p = probs_process_state(flat if self.batch > 1 else flat[0], [] if full else list(local_wires), self.n)
        return p.reshape(self.batch, -1)

    def reduced_dm(self, local_wires):
        from oracle.measure import reduce_statevector

        flat = self.data.numpy()
        return reduce_statevector(flat if self.batch > 1 else flat[0], list(local_wires))

    def probs_device(self, local_wires):
        return torch.from_numpy(np.ascontiguousarray(self.probs(local_wires)).copy())

    # sampler building blocks
    def has_nan(self, p, m):
        return bool(torch.isnan(p).any())

    def np_sum(self, p, m):
        return float(np.sum(p.numpy()))

    def div_by(self, p, m, value):
        a = p.numpy()
        a /= value

    def cumsum(self, p, m, exact, carry):
        a = p.numpy()
        if carry is not None:
            a[0] = carry + a[0]
        np.cumsum(a, out=a)
        return float(a[-1])

    def search(self, cdf, m, u):
        return torch.from_numpy(np.searchsorted(cdf.numpy(), u, side="right").astype(np.int64))

    def unpack_bits(self, idx, m):
        v = idx.numpy()
        return ((v[:, None] >> np.arange(m - 1, -1, -1)) & 1).astype(np.int64)

    def sample_replicated(self, probs_host, shots, rng, exact):
        from oracle.sampling import _sample_probs_numpy

        m = int(np.log2(probs_host.size))
        return _sample_probs_numpy(np.asarray(probs_host).reshape(-1), shots, m, False, rng)

    def synchronize(self):
        pass
