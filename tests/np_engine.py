"""TEST INFRASTRUCTURE: a numpy local engine for ``pennylane_b200.sharded.ShardedStateVector``.

It implements the engine interface of ``sharded.CudaEngine`` with the oracle's numpy kernels so
that the sharded HOST logic (planner, qubit maps, in-place exchanges, ordered reductions, the
distributed sampler) can run under ``gloo`` on CPU ranks.  Never imported by the product.
"""
import numpy as np
import torch

from oracle.apply_operation import apply_operation


class NumpyEngine:
    def __init__(self, nl, batch=1):
        self.n = nl
        self.data = torch.zeros((batch, 1 << nl), dtype=torch.complex128)
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
        arr = np.asarray(arr, dtype=np.complex128).reshape(-1, 1 << self.n)
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
            self.data = torch.from_numpy(np.ascontiguousarray(out).reshape(-1, 1 << self.n).copy())
            self.applied += 1
        return len(handle)

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
