// b200q — shot sampler kernels (K7 of SURVEY.md section 2c).
//
// Replaces pennylane/devices/qubit/sampling.py:500-531, i.e. numpy's
//     norm = probs.sum(); probs /= norm                      (sampling.py:510-526)
//     cdf = probs.cumsum(); cdf /= cdf[-1]                   (Generator.choice, :527)
//     idx = cdf.searchsorted(rng.random(shots), side="right")
//     bits = (idx[:, None] & (1 << arange(m)[::-1])) > 0     (:529-531)
// The uniforms come from the HOST numpy Generator (same PCG64 stream as the reference);
// everything else runs here.  Two CDF modes:
//   exact  : float64 additions performed in numpy's order (pairwise `sum`, sequential
//            `cumsum`) -> bit-identical CDF for identical probabilities -> bit-identical shots;
//   fast   : blocked parallel scan (different rounding, ~1e-16 relative CDF differences).
#pragma once
#include "common.cuh"

namespace b200q {

// ---- numpy pairwise sum, leaf level: 128 consecutive doubles with 8 running accumulators ------
// (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum, PW_BLOCKSIZE = 128).
// One thread per leaf; leaves[i] = leaf sum.  count must be a multiple of 128 here (power-of-two
// probability vectors); shorter vectors are handled by k_np_sum_small.
__global__ void __launch_bounds__(128)
k_np_leaf_sums(const double* __restrict__ p, double* __restrict__ leaves, const uint64_t nleaves) {
  const uint64_t leaf = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= nleaves) return;
  const double* a = p + leaf * 128;
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  for (int i = 8; i < 128; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  }
  leaves[leaf] = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
}

// Perfect binary tree over adjacent pairs (the recursion of pairwise_sum on a power-of-two
// count).  Each CTA collapses up to 2048 consecutive values to one; launched repeatedly.
__global__ void __launch_bounds__(1024)
k_np_tree(const double* __restrict__ in, double* __restrict__ out, const uint64_t count) {
  __shared__ double s[2048];
  const uint64_t base = (uint64_t)blockIdx.x * 2048;
  const uint64_t rem = count - base;
  const int len = rem < 2048 ? (int)rem : 2048;            // power of two
  for (int i = threadIdx.x; i < len; i += blockDim.x) s[i] = in[base + i];
  __syncthreads();
  for (int w = len >> 1; w >= 1; w >>= 1) {               // w <= 1024 = blockDim.x
    const bool act = (int)threadIdx.x < w;
    double v = 0.0;
    if (act) v = __dadd_rn(s[2 * threadIdx.x], s[2 * threadIdx.x + 1]);
    __syncthreads();
    if (act) s[threadIdx.x] = v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// count <= 128 (single leaf or the n < 8 sequential case)
__global__ void k_np_sum_small(const double* __restrict__ a, double* __restrict__ out, const int n) {
  if (threadIdx.x || blockIdx.x) return;
  double res;
  if (n < 8) {
    res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
  } else {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                    __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  }
  out[0] = res;
}

// p[i] /= *norm  (IEEE division, like numpy)
__global__ void __launch_bounds__(256)
k_div_by(double* __restrict__ p, const double* __restrict__ norm, const uint64_t count) {
  const double d = *norm;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    p[i] = __ddiv_rn(p[i], d);
}
// cdf[i] /= cdf[count-1] (the last element must be read before anything is overwritten: the
// divisor is passed through a one-element device buffer filled by k_copy_last)
__global__ void k_copy_last(const double* __restrict__ cdf, double* __restrict__ dst,
                            const uint64_t count) {
  if (threadIdx.x == 0 && blockIdx.x == 0) dst[0] = cdf[count - 1];
}

// ---- exact mode: sequential cumsum in numpy's order ------------------------------------------
// Two warps.  Warp 1 stages CHUNK values into shared memory with coalesced loads and writes
// finished chunks back; lane 0 of warp 0 runs the dependent chain of rounded additions
// (numpy's DOUBLE_add.accumulate: out[0] = p[0]; out[i] = out[i-1] + p[i]).  Triple buffered so
// global traffic overlaps the chain.  In place allowed (cdf may alias p).
// carry_in (nullable): running sum of everything that precedes p[0] in the global vector (the
// previous rank's last CDF entry when the probability vector is sharded): out[0] = carry + p[0].
template <int CHUNK>
__global__ void __launch_bounds__(64)
k_cumsum_serial(const double* p, double* cdf, const uint64_t count,
                const double* __restrict__ carry_in = nullptr) {
  __shared__ double buf[3][CHUNK];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double run = 0.0;
  const uint64_t nchunks = (count + CHUNK - 1) / CHUNK;
  if (wid == 1) {
    for (int i = lane; i < CHUNK; i += 32) buf[0][i] = (uint64_t)i < count ? p[i] : 0.0;
  }
  __syncthreads();
  for (uint64_t c = 0; c < nchunks; ++c) {
    const int cur = (int)(c % 3), nxt = (int)((c + 1) % 3), prv = (int)((c + 2) % 3);
    if (wid == 1) {
      if (c >= 1) {                                         // write back chunk c-1
        for (int i = lane; i < CHUNK; i += 32) {
          const uint64_t g = (c - 1) * CHUNK + i;
          if (g < count) cdf[g] = buf[prv][i];
        }
      }
      if (c + 1 < nchunks) {                                // prefetch chunk c+1
        // NOTE in-place use: chunk c+1 of p is read before chunk c+1 of cdf is written
        // (that write happens at iteration c+2), and buf[nxt] == buffer of chunk c-2,
        // already written back at iteration c-1.
        for (int i = lane; i < CHUNK; i += 32) {
          const uint64_t g = (c + 1) * CHUNK + i;
          buf[nxt][i] = g < count ? p[g] : 0.0;
        }
      }
    } else if (lane == 0) {
      double* b = buf[cur];
      int i0 = 0;
      if (c == 0) { run = carry_in ? __dadd_rn(*carry_in, b[0]) : b[0]; b[0] = run; i0 = 1; }
#pragma unroll 8
      for (int i = i0; i < CHUNK; ++i) { run = __dadd_rn(run, b[i]); b[i] = run; }
    }
    __syncthreads();
  }
  if (wid == 1) {
    const uint64_t c = nchunks - 1;
    const int cur = (int)(c % 3);
    for (int i = lane; i < CHUNK; i += 32) {
      const uint64_t g = c * CHUNK + i;
      if (g < count) cdf[g] = buf[cur][i];
    }
  }
}

// ---- fast mode: blocked parallel inclusive scan (deterministic, not numpy-ordered) ------------
// phase 1: per-block totals; phase 2: exclusive scan of the totals by a single CTA;
// phase 3: per-block scan + offset.  2048 elements per block.
__global__ void __launch_bounds__(256)
k_scan_block_totals(const double* __restrict__ p, double* __restrict__ totals, const uint64_t count) {
  __shared__ double sh[32];
  const uint64_t base = (uint64_t)blockIdx.x * 2048;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t g = base + (uint64_t)threadIdx.x * 8 + k;
    if (g < count) acc += p[g];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) totals[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(1024)
k_scan_totals(double* __restrict__ totals, const uint64_t nblocks,
              const double* __restrict__ carry_in = nullptr) {
  // exclusive scan in place, single CTA, processes 1024 entries per round
  __shared__ double s[1024];
  __shared__ double carry;
  if (threadIdx.x == 0) carry = carry_in ? *carry_in : 0.0;
  __syncthreads();
  for (uint64_t base = 0; base < nblocks; base += 1024) {
    const uint64_t g = base + threadIdx.x;
    const double v = g < nblocks ? totals[g] : 0.0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {            // Hillis-Steele inclusive
      double t = threadIdx.x >= off ? s[threadIdx.x - off] : 0.0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    const double incl = s[threadIdx.x];
    const double c = carry;
    if (g < nblocks) totals[g] = c + (incl - v);           // exclusive prefix
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_scan_apply(const double* __restrict__ p, const double* __restrict__ totals,
             double* __restrict__ cdf, const uint64_t count) {
  __shared__ double wsum[8];
  const uint64_t base = (uint64_t)blockIdx.x * 2048;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double v[8];
  double tsum = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t g = base + (uint64_t)threadIdx.x * 8 + k;
    v[k] = g < count ? p[g] : 0.0;
    tsum += v[k];
    v[k] = tsum;                                           // inclusive within thread
  }
  // warp inclusive scan of thread totals
  double incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  double woff = 0.0;
  for (int w = 0; w < wid; ++w) woff += wsum[w];
  const double off = totals[blockIdx.x] + woff + (incl - tsum);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t g = base + (uint64_t)threadIdx.x * 8 + k;
    if (g < count) cdf[g] = off + v[k];
  }
}

// ---- searchsorted(side="right") + bit unpack -----------------------------------------------------
// idx = number of cdf entries <= u.  Writes the basis-state index and, if bits != nullptr, the
// (shots, m) int64 sample matrix with column 0 = most significant bit (= first wire).
__global__ void __launch_bounds__(256)
k_search(const double* __restrict__ cdf, const uint64_t count, const double* __restrict__ u,
         const uint64_t shots, long long* __restrict__ idx_out, long long* __restrict__ bits,
         const int m) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= shots) return;
  const double x = u[s];
  uint64_t lo = 0, hi = count;                              // first i in [lo,hi) with cdf[i] > x
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(cdf + mid) <= x) lo = mid + 1; else hi = mid;
  }
  if (idx_out) idx_out[s] = (long long)lo;
  if (bits) {
    long long* row = bits + s * (uint64_t)m;
    for (int j = 0; j < m; ++j) row[j] = (long long)((lo >> (m - 1 - j)) & 1ull);
  }
}

// idx -> (shots, m) int64 bit rows, column 0 = most significant bit (sampling.py:529-531)
__global__ void __launch_bounds__(256)
k_unpack_bits(const long long* __restrict__ idx, const uint64_t shots, const int m,
              long long* __restrict__ bits) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= shots) return;
  const unsigned long long v = (unsigned long long)idx[s];
  long long* row = bits + s * (uint64_t)m;
  for (int j = 0; j < m; ++j) row[j] = (long long)((v >> (m - 1 - j)) & 1ull);
}

// NaN scan (sampling.py:322-325: NaN probabilities -> all-zero samples, no exception)
__global__ void __launch_bounds__(256)
k_has_nan(const double* __restrict__ p, const uint64_t count, int* __restrict__ flag) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  int f = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    f |= isnan(p[i]) ? 1 : 0;
  if (f) atomicOr(flag, 1);
}

}  // namespace b200q
