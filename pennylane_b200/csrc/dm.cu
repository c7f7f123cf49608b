// b200q — reduced density matrices straight from the statevector.
//
// Reference: pennylane/math/quantum.py:386-487 (`reduce_statevector`: einsum of the state with
// its conjugate over the traced wires) as used by `qml.density_matrix` / `purity` /
// `vn_entropy` / `mutual_info` (pennylane/measurements/purity.py:50-54, vn_entropy.py:65-67,
// mutual_info.py:92-100).  purity / vn_entropy / mutual_info go through `dm_from_state_vector`
// first — a 4^n matrix — so the reference stops near 14 qubits; here the statevector is read
// and only the 2^m x 2^m result ever exists.
//
// rho[a, b] = sum_r psi[a, r] conj(psi[b, r]),   a, b over the m kept bits, r over the rest.
//
// The kept bits are split into `mi` <= 2 INNER bits, handled inside a thread, and `mo` OUTER
// bits, fixed per launch: one launch computes the 2^mi x 2^mi block G[ai, bi] of rho between the
// outer assignments A (rows) and B (columns).  A thread owns whole columns r: it loads the 2^mi
// amplitudes of (A, r) — and of (B, r) when B != A — with 16-byte loads (the k_dense access
// pattern: adjacent threads touch adjacent addresses for any choice of bits) and accumulates
// the block in 2 * 4^mi FP64 registers.  Deterministic: grid-stride partial per thread, fixed
// xor-tree per CTA, CTAs added in index order by a second kernel.
// Algorithmic bytes per launch: S / 2^mo when A == B, 2 S / 2^mo otherwise; m <= 2 is one
// launch over S.  Diagonal blocks (A == B) accumulate their upper triangle only (real diagonal);
// the host fills every lower triangle by Hermiticity.
#include "common.cuh"
#include "../../include/b200q.h"

namespace b200q {

struct GramArgs {
  int n, mi, nfix;
  int8_t inner[2];          // state-bit positions of the inner kept bits, matrix MSB first
  int8_t fixpos[B200Q_MAX_BITS];   // ascending positions of ALL kept bits (zero-bit insertion)
  uint64_t a_or, b_or;      // outer assignment of the row / column block
  uint64_t ncols;           // 2^(n - m)
};

template <typename T, int MI, bool SAME>
__global__ void __launch_bounds__(256)
k_gram_block(const cx<T>* __restrict__ st, double* __restrict__ partials, const GramArgs a) {
  constexpr int D = 1 << MI;
  __shared__ double sh[32];
  double gr[D][D], gi[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) { gr[i][j] = 0.0; gi[i][j] = 0.0; }
  uint64_t off[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    uint64_t o = 0;
#pragma unroll
    for (int b = 0; b < MI; ++b) o |= (uint64_t)((i >> (MI - 1 - b)) & 1) << a.inner[b];
    off[i] = o;
  }
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.ncols; r += stride) {
    const uint64_t base = insert_zero_bits(r, a.fixpos, a.nfix);
    double ar[D], ai[D], br[D], bi[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const cx<T> v = st[base | a.a_or | off[i]];
      ar[i] = (double)v.x; ai[i] = (double)v.y;
    }
    if (SAME) {
#pragma unroll
      for (int i = 0; i < D; ++i) { br[i] = ar[i]; bi[i] = ai[i]; }
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const cx<T> v = st[base | a.b_or | off[i]];
        br[i] = (double)v.x; bi[i] = (double)v.y;
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        if (SAME && j < i) continue;      // a diagonal block is Hermitian: upper triangle only
        // a_i * conj(b_j)
        gr[i][j] = fma(ar[i], br[j], gr[i][j]); gr[i][j] = fma(ai[i], bi[j], gr[i][j]);
        if (SAME && j == i) continue;     // real diagonal
        gi[i][j] = fma(ai[i], br[j], gi[i][j]); gi[i][j] = fma(-ar[i], bi[j], gi[i][j]);
      }
  }
  // partials layout [row = 2 * (i * D + j) + {re, im}][cta]
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      // (SAME, j < i) entries were never accumulated: they stay 0 and the host mirrors the
      // upper triangle
      const double sr = (SAME && j < i) ? 0.0 : block_sum(gr[i][j], sh);
      const double si = (SAME && j <= i) ? 0.0 : block_sum(gi[i][j], sh);
      if (threadIdx.x == 0) {
        partials[(size_t)(2 * (i * D + j)) * gridDim.x + blockIdx.x] = sr;
        partials[(size_t)(2 * (i * D + j) + 1) * gridDim.x + blockIdx.x] = si;
      }
    }
}

// out[row] = sum over CTAs in index order (one CTA per row, strided then block_sum)
__global__ void __launch_bounds__(256)
k_gram_reduce(const double* __restrict__ partials, double* __restrict__ out, const int ncta) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int c = threadIdx.x; c < ncta; c += blockDim.x) acc += partials[(size_t)blockIdx.x * ncta + c];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

template <typename T>
static int gram_block_t(const void* state, const GramArgs& a, bool same, double* out, void* work,
                        size_t work_bytes, cudaStream_t s) {
  const int D = 1 << a.mi, rows = 2 * D * D;
  unsigned ncta = grid_for(a.ncols, 256, 4);
  B200Q_REQUIRE(work && work_bytes >= (size_t)rows * ncta * sizeof(double),
                "gram_block: workspace too small");
  double* partials = (double*)work;
  const cx<T>* st = (const cx<T>*)state;
#define GRAM_LAUNCH(MI, SAME) \
  k_gram_block<T, MI, SAME><<<ncta, 256, 0, s>>>(st, partials, a)
  if (a.mi == 0) { if (same) GRAM_LAUNCH(0, true); else GRAM_LAUNCH(0, false); }
  else if (a.mi == 1) { if (same) GRAM_LAUNCH(1, true); else GRAM_LAUNCH(1, false); }
  else { if (same) GRAM_LAUNCH(2, true); else GRAM_LAUNCH(2, false); }
#undef GRAM_LAUNCH
  B200Q_LAUNCH_CHECK();
  k_gram_reduce<<<rows, 256, 0, s>>>(partials, out, (int)ncta);
  B200Q_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_gram_block(const void* state, int n, int dtype, const int* inner_bits, int mi,
                                const int* outer_bits, int mo, uint64_t row_assign,
                                uint64_t col_assign, double* out_dev, void* work,
                                size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(n >= 0 && n <= B200Q_MAX_BITS, "gram_block: bad n=%d", n);
  B200Q_REQUIRE(mi >= 0 && mi <= 2 && mo >= 0 && mi + mo <= n, "gram_block: bad mi=%d mo=%d", mi, mo);
  B200Q_REQUIRE(row_assign < (1ull << mo) && col_assign < (1ull << mo),
                "gram_block: outer assignment out of range");
  GramArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.mi = mi;
  uint64_t used = 0;
  for (int j = 0; j < mi; ++j) {
    const int q = inner_bits[j];
    B200Q_REQUIRE(q >= 0 && q < n && !((used >> q) & 1), "gram_block: bad inner bit %d", q);
    used |= 1ull << q;
    a.inner[j] = (int8_t)q;
  }
  for (int j = 0; j < mo; ++j) {
    const int q = outer_bits[j];
    B200Q_REQUIRE(q >= 0 && q < n && !((used >> q) & 1), "gram_block: bad outer bit %d", q);
    used |= 1ull << q;
    // outer_bits[0] is the most significant bit of the outer assignment
    if ((row_assign >> (mo - 1 - j)) & 1ull) a.a_or |= 1ull << q;
    if ((col_assign >> (mo - 1 - j)) & 1ull) a.b_or |= 1ull << q;
  }
  for (int q = 0; q < n; ++q)
    if ((used >> q) & 1ull) a.fixpos[a.nfix++] = (int8_t)q;
  a.ncols = 1ull << (n - mi - mo);
  const bool same = row_assign == col_assign;
  if (dtype == B200Q_DTYPE_C128) return gram_block_t<double>(state, a, same, out_dev, work, work_bytes, s);
  if (dtype == B200Q_DTYPE_C64) return gram_block_t<float>(state, a, same, out_dev, work, work_bytes, s);
  set_error("unknown dtype %d", dtype);
  return 2;
}
