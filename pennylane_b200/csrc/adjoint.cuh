// b200q — adjoint-differentiation reverse-sweep kernel (K8 of SURVEY.md section 2c).
//
// Replaces one iteration of the loop at pennylane/devices/qubit/adjoint_jacobian.py:121-137:
//     ket      = U^dagger ket                                   (:125)
//     ket_temp = (i G U) ket            [operation.py:40-60]    (:129-130)
//     jac[:,p] = Re sum conj(bras) * ket_temp                   (:131)
//     bras[k]  = U^dagger bras[k]                               (:136-137)
// Because (i G U) U^dagger = i G, the derivative term only needs the generator matrix G applied
// to the ket BEFORE U^dagger:  jac = Re <bra| i G |ket_after> = -Im <bra|G|ket_after>.
// The kernel therefore loads each group of 2^K amplitudes of the ket and of one bra ONCE,
// accumulates z = <bra|G|ket>, applies A = U^dagger to both and writes them back:
// 2*S*(1 + n_bras) bytes per op (+ S per extra bra for re-reading the ket) instead of the
// reference's extra ket_temp materialisation and separate (n_obs + 1) * S inner-product read.
//
// Buffer layout: `vecs` = [1 + n_bras][2^n], row 0 = ket, rows 1.. = bras (so that
// non-trainable ops are one batched k_dense launch over all rows).
#pragma once
#include "common.cuh"
#include "gates.cuh"

namespace b200q {

// gridDim.y = bras handled by this launch; bra index b = bra_first + blockIdx.y (row b+1).
// The ket is written back only when write_ket != 0, which the host sets on a launch that covers
// exactly ONE bra and runs AFTER the launches for all other bras (they read the old ket).
// partials: [2][n_bras_total][gridDim.x] (real plane, imag plane) of z.
template <typename T, int K>
__global__ void __launch_bounds__(256)
k_adjoint_step(cx<T>* __restrict__ vecs, const GroupArgs a, const DenseOff<K> o,
               const MatVal<K> adj, const MatVal<K> gen, double* __restrict__ partials,
               const int bra_first, const int n_bras_total, const int write_ket_flag) {
  constexpr int D = 1 << K;
  __shared__ cx<T> sa[D * D];
  __shared__ cx<T> sg[D * D];
  __shared__ double sh[32];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    sa[i] = make_cx<T>((T)adj.m[i].x, (T)adj.m[i].y);
    sg[i] = make_cx<T>((T)gen.m[i].x, (T)gen.m[i].y);
  }
  __syncthreads();
  cx<T>* ket = vecs;
  const int b = bra_first + (int)blockIdx.y;
  cx<T>* bra = vecs + ((uint64_t)(b + 1) << a.n);
  const bool write_ket = (write_ket_flag != 0);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double zr = 0.0, zi = 0.0;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < a.ngroups; g += stride) {
    const uint64_t base = insert_zero_bits(g, a.ins, a.nins) | a.ctrl_or;
    cx<T> x[D], y[D];
#pragma unroll
    for (int r = 0; r < D; ++r) { x[r] = ket[base | o.off[r]]; y[r] = bra[base | o.off[r]]; }
    // z += sum_r conj(y_r) (G x)_r
#pragma unroll
    for (int r = 0; r < D; ++r) {
      cx<T> gx = make_cx<T>(0, 0);
#pragma unroll
      for (int c = 0; c < D; ++c) cmac(gx, sg[r * D + c], x[c]);
      zr += (double)y[r].x * (double)gx.x + (double)y[r].y * (double)gx.y;
      zi += (double)y[r].x * (double)gx.y - (double)y[r].y * (double)gx.x;
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      cx<T> nb = make_cx<T>(0, 0);
#pragma unroll
      for (int c = 0; c < D; ++c) cmac(nb, sa[r * D + c], y[c]);
      bra[base | o.off[r]] = nb;
    }
    if (write_ket) {
#pragma unroll
      for (int r = 0; r < D; ++r) {
        cx<T> nk = make_cx<T>(0, 0);
#pragma unroll
        for (int c = 0; c < D; ++c) cmac(nk, sa[r * D + c], x[c]);
        ket[base | o.off[r]] = nk;
      }
    }
  }
  zr = block_sum(zr, sh);
  zi = block_sum(zi, sh);
  if (threadIdx.x == 0) {
    const size_t plane = (size_t)n_bras_total * gridDim.x;
    partials[(size_t)b * gridDim.x + blockIdx.x] = zr;
    partials[plane + (size_t)b * gridDim.x + blockIdx.x] = zi;
  }
}

}  // namespace b200q
