// b200q — measurement kernels (K6/K10 of SURVEY.md section 2c).
//
// Replaces pennylane/measurements/probs.py:101-135 (|psi|^2 + marginalisation),
// pennylane/devices/qubit/measure.py:74-118 + pennylane/pauli/pauli_arithmetic.py:924-950
// (Pauli-sentence expectation; the reference materialises an int64 index array and a complex
// data array of the full state size per flip-mask group), measure.py:121-139 (inner
// products), simulate.py:120-170 (norms / renormalisation).
//
// All reductions accumulate in float64 in a FIXED order (xor-shuffle tree -> per-CTA partial
// -> one final CTA walking the partials in index order): bit-identical run to run.
#pragma once
#include "common.cuh"

namespace b200q {

// |psi_i|^2 rounded exactly like numpy's `real**2 + imag**2` (two rounded products, one
// rounded add, NO fma contraction) so the sampler sees the same float64 probabilities as the
// reference does for the same amplitudes (measurements/probs.py:102).
__device__ __forceinline__ double abs2_exact(double2 a) {
  return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y));
}
__device__ __forceinline__ double abs2_exact(float2 a) {
  // complex64 states: numpy would produce float32 probabilities; we widen first (strictly
  // more accurate, compared against the complex128 oracle at 1e-5).
  const double x = (double)a.x, y = (double)a.y;
  return __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
}

// ---- full probability vector ------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_probs_full(const cx<T>* __restrict__ state, double* __restrict__ out, const uint64_t total) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;               // independent loads in flight per thread
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < total; i += U * stride) {
    cx<T> a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = state[i + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) out[i + u * stride] = abs2_exact(a[u]);
  }
  for (; i < total; i += stride) out[i] = abs2_exact(state[i]);
}

// ---- marginal probabilities ---------------------------------------------------------------
// Output bin index: bit (m-1-j) of the bin <- state bit tbits[j]  (the requested wire order).
// Work decomposition (deterministic, coalesced for ANY choice of bits):
//   * a warp always reads 32 consecutive amplitudes: lane <-> state bits 0..4;
//   * "outer" target bits (>= 5) are fixed per warp-task, "inner" target bits (< 5) are a
//     function of the lane;
//   * the remaining (summed-over) bits >= 5 are split into `nsplit` contiguous ranges per
//     task; each (task, split) produces a partial; a second kernel adds splits in order.
struct MargArgs {
  int n, m;
  int8_t tbits[B200Q_MAX_BITS];        // requested order, MSB of the bin first
  int n_outer;                         // target bits >= 5
  int8_t outer_pos[B200Q_MAX_BITS];    // ascending state-bit positions of outer targets
  int n_sum_hi;                        // summed-over bits >= 5
  int8_t sum_pos[B200Q_MAX_BITS];      // ascending positions of summed-over bits >= 5
  uint64_t sum_mask;                   // OR of (1 << sum_pos[j])
  unsigned lane_sum_mask;              // lane bits (0..4) that are summed over
  unsigned lane_valid;                 // number of valid lanes: min(32, 2^n)
  int lg_nsplit;                       // log2(number of splits of the summed range)
};

template <typename T>
__global__ void __launch_bounds__(256)
k_probs_marginal(const cx<T>* __restrict__ state, double* __restrict__ partials, const MargArgs a) {
  // partials layout: [batch][nsplit][2^m]
  const unsigned lane = threadIdx.x & 31;
  const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps_total = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t n_tasks = 1ull << (a.n_outer + a.lg_nsplit);
  const int lg_per_split = a.n_sum_hi - a.lg_nsplit;           // summed hi bits per split
  const uint64_t per_split = 1ull << lg_per_split;
  const cx<T>* st = state + ((uint64_t)blockIdx.y << a.n);
  double* outb = partials + ((uint64_t)blockIdx.y << (a.m + a.lg_nsplit));
  const bool lane_ok = lane < a.lane_valid;                     // n < 5: only 2^n lanes carry data
  for (uint64_t task = warp_global; task < n_tasks; task += nwarps_total) {
    const uint64_t o = task >> a.lg_nsplit;                    // outer target assignment
    const uint64_t split = task & ((1ull << a.lg_nsplit) - 1);
    uint64_t base = 0;
    for (int j = 0; j < a.n_outer; ++j) base |= ((o >> j) & 1ull) << a.outer_pos[j];
    // first summed-bit pattern of this split, then enumerate sub-masks in increasing order
    const uint64_t r0 = split * per_split;
    uint64_t sub = 0;
    for (int j = 0; j < a.n_sum_hi; ++j) sub |= ((r0 >> j) & 1ull) << a.sum_pos[j];
    // two accumulators, eight loads in flight per lane; (r even) + (r odd) in a fixed order
    double acc = 0.0, acc1 = 0.0;
#pragma unroll 4
    for (uint64_t r = 0; r + 1 < per_split; r += 2) {
      const uint64_t idx0 = base | sub | (uint64_t)lane;
      sub = ((sub | ~a.sum_mask) + 1ull) & a.sum_mask;
      const uint64_t idx1 = base | sub | (uint64_t)lane;
      sub = ((sub | ~a.sum_mask) + 1ull) & a.sum_mask;
      if (lane_ok) { acc += abs2_exact(st[idx0]); acc1 += abs2_exact(st[idx1]); }
    }
    if (per_split & 1ull) {
      const uint64_t idx = base | sub | (uint64_t)lane;
      if (lane_ok) acc += abs2_exact(st[idx]);
    }
    acc += acc1;
    // reduce over summed lane bits (fixed xor order: bit 4 down to bit 0)
#pragma unroll
    for (int b = 4; b >= 0; --b)
      if (a.lane_sum_mask & (1u << b)) acc += __shfl_xor_sync(0xffffffffu, acc, 1 << b);
    if (lane_ok && (lane & a.lane_sum_mask) == 0) {
      const uint64_t idx = base | (uint64_t)lane;
      uint64_t bin = 0;
      for (int j = 0; j < a.m; ++j) bin = (bin << 1) | ((idx >> a.tbits[j]) & 1ull);
      outb[(split << a.m) + bin] = acc;
    }
  }
}

__global__ void __launch_bounds__(256)
k_sum_splits(const double* __restrict__ partials, double* __restrict__ out, const int m,
             const int lg_nsplit) {
  // out[batch][2^m] = sum over splits in index order
  const uint64_t bins = 1ull << m;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bins) return;
  const double* p = partials + ((uint64_t)blockIdx.y << (m + lg_nsplit));
  double acc = 0.0;
  for (uint64_t s = 0; s < (1ull << lg_nsplit); ++s) acc += p[(s << m) + i];
  out[((uint64_t)blockIdx.y << m) + i] = acc;
}

// ---- final reduction of per-CTA partials ----------------------------------------------------
// partials: [nlaunch][rows][ncta] doubles; out[row] = scale * sum over (launch, cta) in index
// order (one CTA per row: strided accumulation then block_sum => fixed order).
// Same sum for few bins and many splits: one CTA per bin, strided accumulation then block_sum
// (fixed order), instead of one thread walking 2^lg_nsplit partials.
__global__ void __launch_bounds__(256)
k_sum_splits_cta(const double* __restrict__ partials, double* __restrict__ out, const int m,
                 const int lg_nsplit) {
  __shared__ double sh[32];
  const double* p = partials + ((uint64_t)blockIdx.y << (m + lg_nsplit));
  double acc = 0.0;
  for (uint64_t s = threadIdx.x; s < (1ull << lg_nsplit); s += blockDim.x)
    acc += p[(s << m) + blockIdx.x];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[((uint64_t)blockIdx.y << m) + blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256)
k_final_reduce(const double* __restrict__ partials, double* __restrict__ out, const int ncta,
               const int nlaunch, const int rows, const double scale) {
  __shared__ double sh[32];
  double acc = 0.0;
  const int total = nlaunch * ncta;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int l = i / ncta, c = i - l * ncta;
    acc += partials[((size_t)l * rows + blockIdx.x) * ncta + c];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = acc * scale;
}

// ---- Pauli-sentence expectation ---------------------------------------------------------------
// A term is (xmask, zmask, ny, coeff): P|j> = i^{ny} (-1)^{popc(j & zmask)} |j ^ xmask>,
// zmask covering Z and Y factors.  Terms come grouped by xmask (the reference's "sparse
// structure", pauli_arithmetic.py:933-937).
//
// Diagonal group (xmask = 0): sum_j |psi_j|^2 * sum_t c_t (-1)^{popc(j & z_t)}  — every
// diagonal term of the Hamiltonian in ONE read of the state.
struct PauliTerm { uint64_t zmask; double coeff; };       // coeff already includes i^{ny} sign handling

template <typename T>
__global__ void __launch_bounds__(256)
k_expval_diag(const cx<T>* __restrict__ state, const int n, const PauliTerm* __restrict__ terms,
              const int nterms, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PauliTerm* tt = reinterpret_cast<PauliTerm*>(smem_raw);
  __shared__ double sh[32];
  for (int i = threadIdx.x; i < nterms; i += blockDim.x) tt[i] = terms[i];
  __syncthreads();
  const cx<T>* st = state + ((uint64_t)blockIdx.y << n);
  const uint64_t N = 1ull << n;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  // U independent 16-byte loads in flight per thread: one load per iteration left the kernel at
  // 0.55 of the HBM peak (2.4 MB outstanding over the whole GPU against ~13 MB of bandwidth x
  // latency).  The summation order stays a fixed function of (grid, n): deterministic.
  constexpr int U = 8;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < N; i += U * stride) {
    cx<T> a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = st[i + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t idx = i + u * stride;
      const double p = (double)a[u].x * (double)a[u].x + (double)a[u].y * (double)a[u].y;
      double w = 0.0;
      for (int t = 0; t < nterms; ++t) {
        const double c = tt[t].coeff;
        w += (__popcll(idx & tt[t].zmask) & 1) ? -c : c;
      }
      acc = fma(p, w, acc);
    }
  }
  for (; i < N; i += stride) {
    const cx<T> a = st[i];
    const double p = (double)a.x * (double)a.x + (double)a.y * (double)a.y;
    double w = 0.0;
    for (int t = 0; t < nterms; ++t) {
      const double c = tt[t].coeff;
      w += (__popcll(i & tt[t].zmask) & 1) ? -c : c;
    }
    acc = fma(p, w, acc);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// Off-diagonal group (shared xmask != 0): pairs (i, j = i ^ xmask) visited once (pivot bit of
// i cleared).  For a Hermitian Pauli word the two pair contributions are complex conjugates,
// so   <psi|P|psi> += 2 Re( conj(psi_j) f(i) psi_i ),  f(i) = i^{ny} (-1)^{popc(i & z)}.
// All words sharing the xmask are accumulated from the same two loads.
struct PauliTermXY { uint64_t zmask; double coeff; int ny; int pad; };

template <typename T>
__global__ void __launch_bounds__(256)
k_expval_offdiag(const cx<T>* __restrict__ state, const int n, const uint64_t xmask, const int pivot,
                 const PauliTermXY* __restrict__ terms, const int nterms,
                 double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PauliTermXY* tt = reinterpret_cast<PauliTermXY*>(smem_raw);
  __shared__ double sh[32];
  for (int i = threadIdx.x; i < nterms; i += blockDim.x) tt[i] = terms[i];
  __syncthreads();
  const cx<T>* st = state + ((uint64_t)blockIdx.y << n);
  const uint64_t half = 1ull << (n - 1);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const int8_t piv = (int8_t)pivot;
  double acc = 0.0;
  constexpr int U = 4;               // U pairs = 2U independent loads in flight per thread
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; g + (U - 1) * stride < half; g += U * stride) {
    cx<T> av[U], bv[U];
    uint64_t iv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      iv[u] = insert_zero_bits(g + u * stride, &piv, 1);
      av[u] = st[iv[u]];
      bv[u] = st[iv[u] ^ xmask];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = iv[u];
      const cx<T> a = av[u], b = bv[u];
      const double zr = (double)b.x * (double)a.x + (double)b.y * (double)a.y;
      const double zi = (double)b.x * (double)a.y - (double)b.y * (double)a.x;
      double w_re = 0.0, w_im = 0.0;
      for (int t = 0; t < nterms; ++t) {
        double c = tt[t].coeff;
        if (__popcll(i & tt[t].zmask) & 1) c = -c;
        switch (tt[t].ny & 3) {
          case 0: w_re += c; break;
          case 1: w_im += c; break;
          case 2: w_re -= c; break;
          default: w_im -= c; break;
        }
      }
      acc += 2.0 * (w_re * zr - w_im * zi);
    }
  }
  for (; g < half; g += stride) {
    const uint64_t i = insert_zero_bits(g, &piv, 1);
    const uint64_t j = i ^ xmask;
    const cx<T> a = st[i], b = st[j];
    // z = conj(b) * a
    const double zr = (double)b.x * (double)a.x + (double)b.y * (double)a.y;
    const double zi = (double)b.x * (double)a.y - (double)b.y * (double)a.x;
    double w_re = 0.0, w_im = 0.0;     // sum_t c_t f_t(i)
    for (int t = 0; t < nterms; ++t) {
      double c = tt[t].coeff;
      if (__popcll(i & tt[t].zmask) & 1) c = -c;
      switch (tt[t].ny & 3) {
        case 0: w_re += c; break;
        case 1: w_im += c; break;
        case 2: w_re -= c; break;
        default: w_im -= c; break;
      }
    }
    // Re( (w_re + i w_im) * (zr + i zi) ) * 2
    acc += 2.0 * (w_re * zr - w_im * zi);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// ---- <psi| H |psi> for a CSR matrix on k of the n wires (SparseHamiltonian) --------------------
// measure.py:74-118 (csr_dot_products, scipy branch) builds the 2^n x 2^n matrix H (x) I and two
// sparse products; here H stays 2^k x 2^k: amplitude i contributes
//   conj(psi_i) * sum_e H[r(i), c_e] * psi_{i with its target bits replaced by c_e},
// r(i) = the k target bits of i (MSB = first wire of the observable).  Fixed-order reduction.
struct CsrArgs {
  int n, k;
  int8_t tbits[B200Q_MAX_BITS];     // state bit of matrix-index bit (k-1-j)  <- tbits[j]
  uint64_t tmask;
};

template <typename T>
__global__ void __launch_bounds__(256)
k_expval_csr(const cx<T>* __restrict__ state, const CsrArgs a, const long long* __restrict__ indptr,
             const long long* __restrict__ indices, const double2* __restrict__ data,
             double* __restrict__ partials) {
  __shared__ double sh[32];
  const cx<T>* st = state + ((uint64_t)blockIdx.y << a.n);
  const uint64_t N = 1ull << a.n;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    uint64_t r = 0;
    for (int j = 0; j < a.k; ++j) r |= ((i >> a.tbits[j]) & 1ull) << (a.k - 1 - j);
    const cx<T> b = st[i];
    const uint64_t rest = i & ~a.tmask;
    double sr = 0.0, si = 0.0;                       // (H psi)_i
    for (long long e = indptr[r]; e < indptr[r + 1]; ++e) {
      const uint64_t c = (uint64_t)indices[e];
      uint64_t j2 = rest;
      for (int j = 0; j < a.k; ++j) j2 |= ((c >> (a.k - 1 - j)) & 1ull) << a.tbits[j];
      const cx<T> x = st[j2];
      const double2 h = data[e];
      sr += h.x * (double)x.x - h.y * (double)x.y;
      si += h.x * (double)x.y + h.y * (double)x.x;
    }
    acc += (double)b.x * sr + (double)b.y * si;      // Re(conj(b) * (H psi)_i)
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// ---- <a|b> (complex) and ||a||^2 ------------------------------------------------------------------
// partials layout: [2][batch][ncta]  (real plane, imaginary plane)
template <typename T>
__global__ void __launch_bounds__(256)
k_inner(const cx<T>* __restrict__ a, const cx<T>* __restrict__ b, const int n,
        double* __restrict__ partials) {
  __shared__ double sh[32];
  const cx<T>* pa = a + ((uint64_t)blockIdx.y << n);
  const cx<T>* pb = b + ((uint64_t)blockIdx.y << n);
  const uint64_t N = 1ull << n;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double re = 0.0, im = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const cx<T> x = pa[i], y = pb[i];
    re += (double)x.x * (double)y.x + (double)x.y * (double)y.y;
    im += (double)x.x * (double)y.y - (double)x.y * (double)y.x;
  }
  re = block_sum(re, sh);
  im = block_sum(im, sh);
  if (threadIdx.x == 0) {
    const size_t plane = (size_t)gridDim.y * gridDim.x;
    partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = re;
    partials[plane + (size_t)blockIdx.y * gridDim.x + blockIdx.x] = im;
  }
}

// ---- z = <bra| P |ket> for one Pauli word (complex), used for Pauli-generator derivatives of
// gates wider than the fused adjoint kernel handles.  partials: [2][1][ncta].
template <typename T>
__global__ void __launch_bounds__(256)
k_pauli_braket(const cx<T>* __restrict__ bra, const cx<T>* __restrict__ ket, const int n,
               const uint64_t xmask, const uint64_t zmask, const int ny,
               double* __restrict__ partials) {
  __shared__ double sh[32];
  const uint64_t N = 1ull << n;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double re = 0.0, im = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const uint64_t j = i ^ xmask;
    const cx<T> y = bra[i], x = ket[j];
    // conj(y) * x
    double pr = (double)y.x * (double)x.x + (double)y.y * (double)x.y;
    double pi = (double)y.x * (double)x.y - (double)y.y * (double)x.x;
    if (__popcll(j & zmask) & 1) { pr = -pr; pi = -pi; }
    switch (ny & 3) {                       // multiply by i^{ny}
      case 0: re += pr; im += pi; break;
      case 1: re -= pi; im += pr; break;
      case 2: re -= pr; im -= pi; break;
      default: re += pi; im -= pr; break;
    }
  }
  re = block_sum(re, sh);
  im = block_sum(im, sh);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = re;
    partials[gridDim.x + blockIdx.x] = im;
  }
}

// ---- out = sum_t c_t P_t |in>  (observable applied to a ket; builds adjoint bras) ------------------
// One term per launch-loop iteration on the host side would cost T sweeps; instead each thread
// gathers all T partner amplitudes of its own output index (reads hit L2 for nearby masks).
struct PauliTermFull { uint64_t xmask; uint64_t zmask; double cre; double cim; int ny; int pad; };

template <typename T>
__global__ void __launch_bounds__(256)
k_pauli_sum_apply(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, const int n,
                  const PauliTermFull* __restrict__ terms, const int nterms, const double scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PauliTermFull* tt = reinterpret_cast<PauliTermFull*>(smem_raw);
  for (int i = threadIdx.x; i < nterms; i += blockDim.x) tt[i] = terms[i];
  __syncthreads();
  const cx<T>* pin = in + ((uint64_t)blockIdx.y << n);
  cx<T>* pout = out + ((uint64_t)blockIdx.y << n);
  const uint64_t N = 1ull << n;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    double re = 0.0, im = 0.0;
    for (int t = 0; t < nterms; ++t) {
      const uint64_t j = i ^ tt[t].xmask;
      const cx<T> v = pin[j];
      // (P psi)_i = f(j) psi_j,  f(j) = i^{ny} (-1)^{popc(j & z)}
      double fr, fi;
      switch (tt[t].ny & 3) {
        case 0: fr = 1; fi = 0; break;
        case 1: fr = 0; fi = 1; break;
        case 2: fr = -1; fi = 0; break;
        default: fr = 0; fi = -1; break;
      }
      if (__popcll(j & tt[t].zmask) & 1) { fr = -fr; fi = -fi; }
      // coefficient (cre + i cim) * f
      const double kr = tt[t].cre * fr - tt[t].cim * fi;
      const double ki = tt[t].cre * fi + tt[t].cim * fr;
      re += kr * (double)v.x - ki * (double)v.y;
      im += kr * (double)v.y + ki * (double)v.x;
    }
    pout[i] = make_cx<T>((T)(re * scale), (T)(im * scale));
  }
}

}  // namespace b200q
