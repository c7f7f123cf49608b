// b200q — gate-application kernels (K1-K5, K10 of SURVEY.md section 2c).
//
// Replaces the numpy einsum/tensordot/roll/stack paths of
// pennylane/devices/qubit/apply_operation.py:151-255 (dense), :521-609 (X / diagonal),
// :645-759 (RX/RY/RZ), :763-832 (CNOT / MultiControlledX), :507-517 (GlobalPhase).
// Everything is in place: one read and one write of each touched amplitude per gate
// (2*f*S algorithmic bytes, f = touched fraction), versus >= 3 full copies in the reference.
#pragma once
#include "common.cuh"

namespace b200q {

// ---------------------------------------------------------------------------------------
// Argument blocks passed BY VALUE (kernel parameter space -> constant bank, no extra
// H2D copies, CUDA-graph friendly).
// ---------------------------------------------------------------------------------------
struct GroupArgs {
  int n;                               // qubits in this (local) state
  int nins;                            // number of zero-insert positions (targets + controls)
  int8_t ins[B200Q_MAX_CTRL + 12];     // ascending bit positions
  uint64_t ctrl_or;                    // control bits that must be 1
  uint64_t ngroups;                    // 2^(n - nins)
};

template <int K> struct DenseOff { uint64_t off[1 << K]; };
template <int K> struct MatVal { double2 m[(1 << K) * (1 << K)]; };   // row-major, complex128

// ---------------------------------------------------------------------------------------
// K1/K4/K5: dense 2^K x 2^K matrix on arbitrary target bits with arbitrary controls,
// K <= 3, one thread per group of 2^K amplitudes, matrix staged in shared memory
// (warp-broadcast LDS).  gridDim.y = batch.
// ---------------------------------------------------------------------------------------
template <typename T, int K, bool BYVAL>
__global__ void __launch_bounds__(256)
k_dense(cx<T>* __restrict__ state, const GroupArgs a, const DenseOff<K> o, const MatVal<K> mv,
        const cx<T>* __restrict__ mat_dev, const long long mat_bstride) {
  constexpr int D = 1 << K;
  __shared__ cx<T> sm[D * D];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    if (BYVAL) sm[i] = make_cx<T>((T)mv.m[i].x, (T)mv.m[i].y);
    else sm[i] = mat_dev[(long long)blockIdx.y * mat_bstride + i];
  }
  __syncthreads();
  cx<T>* st = state + ((uint64_t)blockIdx.y << a.n);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // U groups per thread and iteration: all their loads are issued before the first FMA, so
  // 8 (K = 1, 2) amplitudes are in flight per thread instead of 2^K (one group per iteration
  // left the single-qubit sweep at 0.80 of the HBM peak).
  constexpr int U = K == 1 ? 4 : K == 2 ? 2 : 1;
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; g + (U - 1) * stride < a.ngroups; g += U * stride) {
    uint64_t base[U];
    cx<T> x[U][D];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      base[u] = insert_zero_bits(g + u * stride, a.ins, a.nins) | a.ctrl_or;
#pragma unroll
      for (int r = 0; r < D; ++r) x[u][r] = st[base[u] | o.off[r]];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int r = 0; r < D; ++r) {
        cx<T> y = make_cx<T>(0, 0);
#pragma unroll
        for (int c = 0; c < D; ++c) cmac(y, sm[r * D + c], x[u][c]);
        st[base[u] | o.off[r]] = y;
      }
    }
  }
  for (; g < a.ngroups; g += stride) {
    const uint64_t base = insert_zero_bits(g, a.ins, a.nins) | a.ctrl_or;
    cx<T> x[D];
#pragma unroll
    for (int r = 0; r < D; ++r) x[r] = st[base | o.off[r]];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      cx<T> y = make_cx<T>(0, 0);
#pragma unroll
      for (int c = 0; c < D; ++c) cmac(y, sm[r * D + c], x[c]);
      st[base | o.off[r]] = y;
    }
  }
}

// ---------------------------------------------------------------------------------------
// K5 (large blocks): dense 2^K x 2^K, 4 <= K <= 10, through shared memory.  A CTA stages
// G groups (G * 2^K = TILE amplitudes), each thread produces TILE / blockDim outputs.
// The matrix is read through the read-only path (L1/L2 resident: <= 16 MiB at K = 10).
// ---------------------------------------------------------------------------------------
struct BigArgs {
  int n, k, nins;
  int8_t ins[B200Q_MAX_CTRL + 12];
  int8_t tbits[12];                    // target bit of matrix-index bit (k-1-j)  <- tbits[j]
  uint64_t ctrl_or;
  uint64_t ngroups;
};

template <typename T, int TILE>
__global__ void __launch_bounds__(256)
k_dense_big(cx<T>* __restrict__ state, const BigArgs a, const cx<T>* __restrict__ mat_dev,
            const long long mat_bstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* xs = reinterpret_cast<cx<T>*>(smem_raw);                 // [D][G]  (c-major)
  uint64_t* offs = reinterpret_cast<uint64_t*>(xs + TILE);        // [D]
  const int D = 1 << a.k;
  const int G = TILE / D;                                          // groups per tile
  const int lgG = 31 - __clz(G);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    uint64_t off = 0;
    for (int j = 0; j < a.k; ++j)
      if ((c >> (a.k - 1 - j)) & 1) off |= 1ull << a.tbits[j];
    offs[c] = off;
  }
  __syncthreads();
  cx<T>* st = state + ((uint64_t)blockIdx.y << a.n);
  const cx<T>* mat = mat_dev + (long long)blockIdx.y * mat_bstride;
  const uint64_t ntiles = (a.ngroups + G - 1) / G;
  constexpr int PER = TILE / 256;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // stage
    for (int e = threadIdx.x; e < TILE; e += 256) {
      const int gi = e & (G - 1), c = e >> lgG;
      const uint64_t g = t * G + gi;
      if (g < a.ngroups) {
        const uint64_t base = insert_zero_bits(g, a.ins, a.nins) | a.ctrl_or;
        xs[c * G + gi] = st[base | offs[c]];
      }
    }
    __syncthreads();
    // Register blocking: the PER outputs of a thread (e = tid + 256 p) share either their group
    // (G <= 256: the column entry xs[c][gi] is loaded once per c and meets PER matrix rows) or
    // their row (G > 256: one matrix entry meets PER groups).  The first version looped over c
    // per output: two loads for every complex multiply-add (4 DFMA), i.e. bound by load issue;
    // now PER complex multiply-adds share 1 + PER loads, and the contraction runs at the FP64
    // pipe rate from K = 5 on (its roofline: 4 * 2^K DFMA per amplitude, see
    // bench.py `per_gate_kernels.dense_k*`).
    cx<T> y[PER];
#pragma unroll
    for (int p = 0; p < PER; ++p) y[p] = make_cx<T>(0, 0);
    if (G <= 256) {
      const int gi = threadIdx.x & (G - 1);
      const int r0 = threadIdx.x >> lgG;                 // rows r0 + p * (256 / G)
      const int rstep = 256 >> lgG;
      const cx<T>* mrow = mat + (size_t)r0 * D;
      const size_t mstep = (size_t)rstep * D;
#pragma unroll 2
      for (int c = 0; c < D; ++c) {
        const cx<T> x = xs[c * G + gi];
#pragma unroll
        for (int p = 0; p < PER; ++p) cmac(y[p], __ldg(mrow + p * mstep + c), x);
      }
    } else {
      // G is a multiple of 256: all PER outputs of a thread are in rows r = p * 256 / G ... with
      // group gi = (tid + 256 p) & (G - 1)
#pragma unroll
      for (int p = 0; p < PER; ++p) {
        const int e = threadIdx.x + p * 256;
        const int gi = e & (G - 1), r = e >> lgG;
        const cx<T>* mrow = mat + (size_t)r * D;
        cx<T> acc = make_cx<T>(0, 0);
        for (int c = 0; c < D; ++c) cmac(acc, __ldg(mrow + c), xs[c * G + gi]);
        y[p] = acc;
      }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int e = threadIdx.x + p * 256;
      const int gi = e & (G - 1), r = e >> lgG;
      const uint64_t g = t * G + gi;
      if (g < a.ngroups) {
        const uint64_t base = insert_zero_bits(g, a.ins, a.nins) | a.ctrl_or;
        st[base | offs[r]] = y[p];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K2: general diagonal gate on k bits (table of 2^k complex entries, matrix-index MSB first).
// Pure streaming: every thread owns whole amplitudes, perfectly coalesced for any bits.
// Table in shared memory when it fits (k <= 11), else read-only global.
// ---------------------------------------------------------------------------------------
struct DiagArgs {
  int n, k;
  int8_t bits[24];                     // bits[j] <- matrix-index bit (k-1-j)
};

// MODE 0: table by value (k <= 6), 1: device table staged in shared memory (k <= 11),
// 2: device table read through the read-only path.
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_diag(cx<T>* __restrict__ state, const DiagArgs a, const MatVal<3> dv,
       const cx<T>* __restrict__ diag_dev, const long long diag_bstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* tab = reinterpret_cast<cx<T>*>(smem_raw);
  const cx<T>* src = diag_dev + (long long)blockIdx.y * diag_bstride;
  if (MODE == 0) {
    for (int i = threadIdx.x; i < (1 << a.k); i += blockDim.x)
      tab[i] = make_cx<T>((T)dv.m[i].x, (T)dv.m[i].y);
    __syncthreads();
  } else if (MODE == 1) {
    for (int i = threadIdx.x; i < (1 << a.k); i += blockDim.x) tab[i] = src[i];
    __syncthreads();
  }
  cx<T>* st = state + ((uint64_t)blockIdx.y << a.n);
  const uint64_t N = 1ull << a.n;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    unsigned idx = 0;
    for (int j = 0; j < a.k; ++j) idx = (idx << 1) | (unsigned)((i >> a.bits[j]) & 1ull);
    const cx<T> d = (MODE == 2) ? __ldg(src + idx) : tab[idx];
    st[i] = cmul(d, st[i]);
  }
}

// ---------------------------------------------------------------------------------------
// K2/K10: multiply the subspace where all control bits are 1 by one complex scalar.
// Covers PauliZ, S, T, PhaseShift, CZ, CCZ, ControlledPhaseShift (touch 2^-C of the state),
// GlobalPhase / scale / renormalise (C = 0).  Scalar by value, or per-batch from device.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_phase(cx<T>* __restrict__ state, const GroupArgs a, const double2 phase,
        const cx<T>* __restrict__ phase_dev) {
  cx<T> ph = make_cx<T>((T)phase.x, (T)phase.y);
  if (phase_dev) ph = phase_dev[blockIdx.y];
  cx<T>* st = state + ((uint64_t)blockIdx.y << a.n);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < a.ngroups; g += stride) {
    const uint64_t i = insert_zero_bits(g, a.ins, a.nins) | a.ctrl_or;
    st[i] = cmul(ph, st[i]);
  }
}

// ---------------------------------------------------------------------------------------
// K2: parity phase: amp *= (popcount(i & mask) odd ? p1 : p0).  RZ, IsingZZ, MultiRZ and
// PauliRot over Z strings of ANY width are this one streaming kernel.
// phases: by value (p0,p1) or per-batch pairs from device memory.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_parity_phase(cx<T>* __restrict__ state, const int n, const uint64_t mask, const double2 p0v,
               const double2 p1v, const cx<T>* __restrict__ phases_dev) {
  cx<T> p0 = make_cx<T>((T)p0v.x, (T)p0v.y), p1 = make_cx<T>((T)p1v.x, (T)p1v.y);
  if (phases_dev) { p0 = phases_dev[2 * blockIdx.y]; p1 = phases_dev[2 * blockIdx.y + 1]; }
  cx<T>* st = state + ((uint64_t)blockIdx.y << n);
  const uint64_t N = 1ull << n;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const cx<T> p = (__popcll(i & mask) & 1) ? p1 : p0;
    st[i] = cmul(p, st[i]);
  }
}

// ---------------------------------------------------------------------------------------
// Pauli-string rotation exp(-i theta/2 P) for a string with at least one X/Y:
//   psi'_i = c psi_i - i s f(i^x) psi_{i^x},   f(j) = i^{nY} (-1)^{popc(j & zmask)}
// One sweep for ANY string width (reference builds the dense 2^k matrix:
// ops/qubit/parametric_ops_multi_qubit.py:380-436).  Pairs (i, i^xmask) enumerated by
// clearing the pivot bit (highest X/Y bit).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_pauli_rot(cx<T>* __restrict__ state, const int n, const int pivot, const uint64_t xmask,
            const uint64_t zmask, const int ny, const double2 csv,
            const cx<T>* __restrict__ cs_dev) {
  T c = (T)csv.x, s = (T)csv.y;
  if (cs_dev) { c = cs_dev[blockIdx.y].x; s = cs_dev[blockIdx.y].y; }
  cx<T>* st = state + ((uint64_t)blockIdx.y << n);
  // -i * i^{ny}: table over ny mod 4 -> (re, im)
  const int q = ny & 3;
  const T fr = (q == 1) ? (T)1 : (q == 3) ? (T)-1 : (T)0;    // real part of -i * i^ny
  const T fi = (q == 0) ? (T)-1 : (q == 2) ? (T)1 : (T)0;    // imag part
  const uint64_t half = 1ull << (n - 1);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const int8_t piv = (int8_t)pivot;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < half; g += stride) {
    const uint64_t i = insert_zero_bits(g, &piv, 1);
    const uint64_t j = i ^ xmask;
    const cx<T> a = st[i], b = st[j];
    const T sj = (__popcll(j & zmask) & 1) ? -s : s;          // s * (-1)^{popc(j&z)}
    const T si = (__popcll(i & zmask) & 1) ? -s : s;
    cx<T> ni, nj;
    // ni = c a + (fr + i fi) * sj * b
    ni.x = c * a.x + sj * (fr * b.x - fi * b.y);
    ni.y = c * a.y + sj * (fr * b.y + fi * b.x);
    nj.x = c * b.x + si * (fr * a.x - fi * a.y);
    nj.y = c * b.y + si * (fr * a.y + fi * a.x);
    st[i] = ni; st[j] = nj;
  }
}

// ---------------------------------------------------------------------------------------
// Mid-circuit measurement collapse (apply_operation.py:478-495): project bit `q` on `sample`,
// rescale the surviving half by 1/||P psi||, and (reset && sample == 1) move it to the
// bit = 0 half — the reference's projector sweep, `state / norm` sweep and reset sweep in one
// pass that reads only the surviving half: S/2 read + S written.
//   src  = half that survives          (bit == sample)
//   keep = where it lands              (bit == sample, or 0 when reset)
//   zero = the other half
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_collapse(cx<T>* __restrict__ state, const int n, const int q, const int sample,
           const int reset, const double scale) {
  const uint64_t half = 1ull << (n - 1);
  const uint64_t bitq = 1ull << q;
  const uint64_t src_or = sample ? bitq : 0ull;
  const uint64_t keep_or = (sample && !reset) ? bitq : 0ull;
  const uint64_t zero_or = keep_or ^ bitq;
  const T sc = (T)scale;
  const int8_t pos = (int8_t)q;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < half; g += stride) {
    const uint64_t i = insert_zero_bits(g, &pos, 1);
    cx<T> v = state[i | src_or];
    v.x *= sc; v.y *= sc;
    state[i | keep_or] = v;
    state[i | zero_or] = make_cx<T>((T)0, (T)0);
  }
}

// ---------------------------------------------------------------------------------------
// State initialisation: |index> for every batch element (initialize_state.py:43-44).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void k_set_one(cx<T>* __restrict__ state, const int n, const uint64_t index) {
  state[((uint64_t)blockIdx.x << n) + index] = make_cx<T>(1, 0);
}

}  // namespace b200q
