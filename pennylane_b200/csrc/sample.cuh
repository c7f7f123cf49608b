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

// ---- exact mode, parallel --------------------------------------------------------------------
// numpy's cumsum is the dependent chain s_i = fl(s_{i-1} + p_i).  With non-negative terms the
// chain can be cut into independent pieces WITHOUT changing a single bit: while the running sum
// stays inside one binade [2^e, 2^(e+1)) it is an integer multiple M of u = ulp = 2^(e-52), and
//     fl(M u + p) = (M + k) u,   k = p / u rounded to the nearest integer,
// where k does not depend on M — except when p / u lies exactly half-way (ties go to the even
// M + k), and a tie only needs the PARITY of M.  So a run of terms acts on the state as
//     M -> M + (M even ? a0 : a1),
// and such maps compose associatively: a chunk is summarised by the pair (a0, a1), chunks are
// combined in order by one warp, and every chunk is then replayed from its exact start state.
// A chunk is only summarised for ONE binade, predicted from an ordinary (approximate) parallel
// prefix of the chunk totals; where the prediction is wrong or the sum crosses into the next
// binade inside the chunk (a few dozen chunks out of half a million at 30 qubits) the ordered
// pass falls back to the plain rounded additions for that chunk.  Zero, subnormal and huge
// terms are covered by the same rules (unit exponent 1 for zero / subnormal sums; a term whose
// own unit is larger than the sum's forces the fallback).  Checked bit for bit against the
// serial kernel and numpy on adversarial inputs (tests/test_gpu_sampling.py).
constexpr int EXC_LANE = 64;                    // consecutive terms per lane
constexpr int EXC_CHUNK = 32 * EXC_LANE;        // terms per warp-chunk (= block of the fast scan)

struct ExcSum { unsigned long long a0, a1; int bad; };

// unit exponent (biased; 1 for zero / subnormal) and integer significand of a non-negative double
__device__ __forceinline__ void exc_split(const double s, int& ue, unsigned long long& m) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(s);
  const int eb = (int)((b >> 52) & 0x7ffull);
  const unsigned long long man = b & ((1ull << 52) - 1ull);
  ue = eb ? eb : 1;
  m = eb ? (man | (1ull << 52)) : man;
}
__device__ __forceinline__ double exc_join(const int ue, const unsigned long long m) {
  const unsigned long long b = m < (1ull << 52) ? m : (((unsigned long long)ue << 52) | (m - (1ull << 52)));
  return __longlong_as_double((long long)b);
}
// p / u = k (+ 1/2 if tie); bad: the term's unit is larger than the sum's (certain binade change)
__device__ __forceinline__ void exc_term(const double p, const int ue, unsigned long long& k, bool& tie,
                                         bool& bad) {
  int ep;
  unsigned long long mp;
  exc_split(p, ep, mp);
  k = 0; tie = false; bad = false;
  if (mp == 0) return;
  const int d = ue - ep;
  if (d < 0) { bad = true; return; }
  if (d == 0) { k = mp; return; }
  if (d >= 54) return;
  const unsigned long long r = mp & ((1ull << d) - 1ull), half = 1ull << (d - 1);
  k = (mp >> d) + (r > half ? 1ull : 0ull);
  tie = r == half;
}
// f then g
__device__ __forceinline__ ExcSum exc_compose(const ExcSum f, const ExcSum g) {
  ExcSum r;
  r.a0 = f.a0 + ((f.a0 & 1ull) ? g.a1 : g.a0);
  r.a1 = f.a1 + ((f.a1 & 1ull) ? g.a0 : g.a1);       // start odd: parity after f = 1 ^ (a1 & 1)
  r.bad = f.bad | g.bad;
  return r;
}
// summary of the lane's EXC_LANE consecutive terms for unit exponent `ue`
__device__ __forceinline__ ExcSum exc_lane_summary(const double* __restrict__ p, const uint64_t first,
                                                   const uint64_t count, const int ue) {
  ExcSum s; s.a0 = 0; s.a1 = 0; s.bad = 0;
  for (int i = 0; i < EXC_LANE; ++i) {
    const uint64_t g = first + i;
    if (g >= count) break;
    unsigned long long k; bool tie, bad;
    exc_term(p[g], ue, k, tie, bad);
    s.bad |= bad ? 1 : 0;
    const unsigned long long t0 = s.a0 + k, t1 = s.a1 + k;
    s.a0 = t0 + ((tie && (t0 & 1ull)) ? 1ull : 0ull);           // start even: parity of M + a0 + k
    s.a1 = t1 + ((tie && !(t1 & 1ull)) ? 1ull : 0ull);          // start odd
  }
  return s;
}
__device__ __forceinline__ ExcSum exc_shfl_up(const ExcSum v, const int o) {
  ExcSum r;
  r.a0 = __shfl_up_sync(0xffffffffu, v.a0, o);
  r.a1 = __shfl_up_sync(0xffffffffu, v.a1, o);
  r.bad = __shfl_up_sync(0xffffffffu, v.bad, o);
  return r;
}

// K1: one warp per chunk; approx[c] = approximate value of the running sum before chunk c
__global__ void __launch_bounds__(128)
k_excum_summaries(const double* __restrict__ p, const uint64_t count, const double* __restrict__ approx,
                  const uint64_t nchunks, unsigned long long* __restrict__ a0,
                  unsigned long long* __restrict__ a1, int* __restrict__ flags) {
  const uint64_t c = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= nchunks) return;
  const int lane = threadIdx.x & 31;
  int ue; unsigned long long m0;
  exc_split(approx[c], ue, m0);
  ExcSum s = exc_lane_summary(p, c * EXC_CHUNK + (uint64_t)lane * EXC_LANE, count, ue);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const ExcSum e = exc_shfl_up(s, o);
    if (lane >= o) s = exc_compose(e, s);
  }
  if (lane == 31) { a0[c] = s.a0; a1[c] = s.a1; flags[c] = s.bad; }
}

// K2: ONE warp walks the chunks in order with the exact state (replicated in every lane).
// start[c] = exact running sum before chunk c; flags[c] |= 2 where the chunk was done here.
__global__ void __launch_bounds__(32)
k_excum_ordered(const double* p, double* cdf, const uint64_t count, const double* __restrict__ approx,
                const uint64_t nchunks, const unsigned long long* __restrict__ a0,
                const unsigned long long* __restrict__ a1, int* __restrict__ flags,
                double* __restrict__ start, const double* __restrict__ carry_in) {
  const int lane = threadIdx.x;
  double s = carry_in ? *carry_in : 0.0;
  bool first = carry_in == nullptr;              // numpy: out[0] = p[0] (no addition)
  for (uint64_t base = 0; base < nchunks; base += 32) {
    const uint64_t mine = base + lane;
    unsigned long long la0 = 0, la1 = 0; int lbad = 1, lue = 0;
    if (mine < nchunks) {
      la0 = a0[mine]; la1 = a1[mine]; lbad = flags[mine];
      unsigned long long mm; exc_split(approx[mine], lue, mm);
    }
    double my_start = 0.0; int my_flag = 0;
    const int lim = (int)min((uint64_t)32, nchunks - base);
    for (int j = 0; j < lim; ++j) {
      const unsigned long long ca0 = __shfl_sync(0xffffffffu, la0, j), ca1 = __shfl_sync(0xffffffffu, la1, j);
      const int cbad = __shfl_sync(0xffffffffu, lbad, j), cue = __shfl_sync(0xffffffffu, lue, j);
      int ue; unsigned long long m;
      exc_split(s, ue, m);
      const unsigned long long add = (m & 1ull) ? ca1 : ca0;
      const bool ok = !first && !cbad && cue == ue && add < (1ull << 53) && m + add < (1ull << 53);
      if (lane == j) { my_start = s; my_flag = ok ? 0 : 2; }
      if (ok) {
        s = exc_join(ue, m + add);
      } else {
        // plain rounded additions for this chunk (lane 0 owns the chain; loads are independent)
        const uint64_t b0 = (base + j) * EXC_CHUNK;
        const uint64_t e0 = min(count, b0 + (uint64_t)EXC_CHUNK);
        if (lane == 0) {
          double run = s;
          uint64_t i = b0;
          if (first) { run = p[i]; cdf[i] = run; ++i; }
#pragma unroll 4
          for (; i < e0; ++i) { run = __dadd_rn(run, p[i]); cdf[i] = run; }
          s = run;
        }
        s = __shfl_sync(0xffffffffu, s, 0);
        first = false;
      }
    }
    if (mine < nchunks) { start[mine] = my_start; flags[mine] |= my_flag; }
  }
}

// K3: replay every chunk that was not done by the ordered pass from its exact start state
__global__ void __launch_bounds__(128)
k_excum_replay(const double* p, double* cdf, const uint64_t count, const uint64_t nchunks,
               const int* __restrict__ flags, const double* __restrict__ start) {
  const uint64_t c = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= nchunks || (flags[c] & 2)) return;
  const int lane = threadIdx.x & 31;
  int ue; unsigned long long m;
  exc_split(start[c], ue, m);
  const uint64_t first = c * EXC_CHUNK + (uint64_t)lane * EXC_LANE;
  const ExcSum mine = exc_lane_summary(p, first, count, ue);
  ExcSum s = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const ExcSum e = exc_shfl_up(s, o);
    if (lane >= o) s = exc_compose(e, s);
  }
  // exclusive prefix of the lanes before this one
  ExcSum ex = exc_shfl_up(s, 1);
  if (lane == 0) { ex.a0 = 0; ex.a1 = 0; }
  m += (m & 1ull) ? ex.a1 : ex.a0;
  for (int i = 0; i < EXC_LANE; ++i) {
    const uint64_t g = first + i;
    if (g >= count) break;
    unsigned long long k; bool tie, bad;
    exc_term(p[g], ue, k, tie, bad);
    if (tie && ((m + k) & 1ull)) ++k;
    m += k;
    cdf[g] = exc_join(ue, m);
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
