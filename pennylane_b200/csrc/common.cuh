// b200q — shared device/host helpers for the sm_100a statevector kernels.
//
// Conventions used by every kernel in this directory:
//   * the state is a flat array of `batch * 2^n` complex amplitudes (float2 / double2);
//     batch element b starts at b * 2^n.  This is default.qubit's `(B, 2, ..., 2)` array
//     flattened (reference: pennylane/devices/qubit/initialize_state.py:43-44,
//     apply_operation.py:301), so PennyLane wire w is BIT POSITION q = n-1-w of the index.
//   * all C-ABI entry points talk in bit positions (q = 0 has stride 1); the Python host
//     converts wires to bits (and, when sharded, logical to physical bits).
//   * indices are 64-bit everywhere (n can exceed 31).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define B200Q_C64 0
#define B200Q_C128 1

#define B200Q_MAX_DENSE_K 3     // dense gates applied from registers (by-value matrix)
#define B200Q_MAX_BIG_K 10      // dense gates applied through shared memory
#define B200Q_MAX_CTRL 16
#define B200Q_MAX_BITS 48

namespace b200q {

void set_error(const char* fmt, ...);
int sm_count();

#define B200Q_CHECK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      b200q::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                       cudaGetErrorString(_e));                                        \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

#define B200Q_LAUNCH_CHECK()                                                           \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      b200q::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,               \
                       cudaGetErrorString(_e));                                        \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

#define B200Q_REQUIRE(cond, ...)                                                       \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      b200q::set_error(__VA_ARGS__);                                                   \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };
template <typename T> using cx = typename vec2<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> make_cx(T x, T y) {
  cx<T> r; r.x = x; r.y = y; return r;
}
// r += a * b   (complex multiply-accumulate, 4 FMAs)
template <typename C> __device__ __forceinline__ void cmac(C& r, const C a, const C b) {
  r.x = fma(a.x, b.x, r.x); r.x = fma(-a.y, b.y, r.x);
  r.y = fma(a.x, b.y, r.y); r.y = fma(a.y, b.x, r.y);
}
template <typename C> __device__ __forceinline__ C cmul(const C a, const C b) {
  C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// conj(a) * b
template <typename C> __device__ __forceinline__ C cmulc(const C a, const C b) {
  C r; r.x = a.x * b.x + a.y * b.y; r.y = a.x * b.y - a.y * b.x; return r;
}

// Insert a zero bit at each of the (ascending) positions pos[0..m): the standard
// "free index -> amplitude index" expansion.
__host__ __device__ __forceinline__ uint64_t insert_zero_bits(uint64_t g, const int8_t* pos, int m) {
#pragma unroll 4
  for (int i = 0; i < m; ++i) {
    const int p = pos[i];
    const uint64_t low = g & ((1ull << p) - 1ull);
    g = ((g >> p) << (p + 1)) | low;
  }
  return g;
}

// streaming (evict-first-ish) access helpers are deliberately NOT used for the state:
// at n <= 22 the whole state is L2 resident and default caching wins; at n >= 28 nothing fits
// anyway.

// ---- deterministic block reduction (double) --------------------------------------------
// Every thread contributes v; thread 0 of the block gets the block sum.  Fixed order
// (xor-shuffle tree inside a warp, then a sequential pass over warps by warp 0) so that
// repeated runs give bit-identical results.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse between consecutive calls
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = (lane < nw) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// Launch geometry for streaming kernels: enough CTAs to fill 148 SMs several times over,
// but bounded so grid-stride loops amortise index setup.
inline unsigned grid_for(uint64_t work_items, int block, int max_ctas_per_sm = 8) {
  uint64_t need = (work_items + block - 1) / block;
  uint64_t cap = (uint64_t)sm_count() * max_ctas_per_sm;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

}  // namespace b200q
