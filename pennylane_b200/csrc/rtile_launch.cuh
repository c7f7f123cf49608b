// b200q — launch template of the register-tiled kernel.  Every kernel variant is explicitly
// instantiated in its own translation unit (rtile_k_*.cu, one line each) so that the large
// interpreter kernels compile in parallel; rtile.cu sees them as extern templates.
#pragma once
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "rtile.cuh"
#include "rtile_host.h"

namespace b200q {

// out[row] = scale * sum over CTAs, fixed order (same contract as k_final_reduce, measure.cuh)
static __global__ void __launch_bounds__(256)
k_rt_final_reduce(const double* __restrict__ partials, double* __restrict__ out, const int ncta,
                  const double scale) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < ncta; i += blockDim.x) acc += partials[(size_t)blockIdx.x * ncta + i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = acc * scale;
}

template <typename T, int RB, int NV, int THREADS, int MINB, bool WS>
int rtile_launch(void* v0, void* v1, const RtArgs& a, int64_t batch, const RtOp* ops_dev,
                        const double2* mats_dev, long long mat_bstride, int nslots, double scale, double* out_dev,
                        double* partials, size_t partial_cap, cudaStream_t s) {
  const size_t smem = ((size_t)NV * sizeof(cx<T>) << a.T) * (WS ? 2 : 1) + sizeof(cx<T>) * ((a.nmat + 1) & ~1) +
                      (size_t)a.nops * sizeof(RtOp) +
                      (size_t)((2 << RB) + 2 * THREADS + 2) * sizeof(unsigned long long) +
                      (size_t)nslots * (THREADS / 32) * sizeof(double) +
                      ((size_t)a.nrounds * (THREADS + 8) + THREADS + (1 << RB)) * sizeof(unsigned short) +
                      ((size_t)a.nops + 1 + (1 << RB)) * sizeof(unsigned) +
                      (size_t)a.nd1 * THREADS * sizeof(unsigned short) + 32;
  B200Q_REQUIRE(smem <= 227 * 1024, "rtile: %zu bytes of shared memory needed (%d ops, %d slots)",
                smem, a.nops, nslots);
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CHECK(cudaFuncSetAttribute(k_rtile<T, RB, NV, THREADS, MINB, WS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  uint64_t per_sm = std::max<uint64_t>(1, std::min<uint64_t>(MINB, (227 * 1024) / smem));
  static const int ctas_knob = getenv("B200Q_RT_CTAS") ? atoi(getenv("B200Q_RT_CTAS")) : 0;   // tuning knob
  if (ctas_knob > 0) per_sm = ctas_knob;
  const uint64_t cap = (uint64_t)sm_count() * per_sm;
  dim3 grid((unsigned)std::min<uint64_t>(a.ntiles, cap), (unsigned)batch);
  if (nslots > 0)
    B200Q_REQUIRE((size_t)batch * nslots * grid.x <= partial_cap, "rtile: workspace too small for %d slots",
                  nslots);
  k_rtile<T, RB, NV, THREADS, MINB, WS><<<grid, THREADS + (WS ? 32 : 0), smem, s>>>(
      a, (cx<T>*)v0, (cx<T>*)v1, ops_dev, mats_dev, mat_bstride, partials);
  B200Q_LAUNCH_CHECK();
  if (nslots > 0) {
    k_rt_final_reduce<<<(unsigned)(batch * nslots), 256, 0, s>>>(partials, out_dev, (int)grid.x, scale);
    B200Q_LAUNCH_CHECK();
  }
  return 0;
}

#define RT_LAUNCH_ARGS                                                                          \
  void *v0, void *v1, const RtArgs &a, int64_t batch, const RtOp *ops_dev,                      \
      const double2 *mats_dev, long long mat_bstride, int nslots, double scale,                \
      double *out_dev, double *partials,                                                        \
      size_t partial_cap, cudaStream_t s

// the variants that exist (kept in sync with rtile_k_*.cu and rtile_run in rtile.cu)
#define RT_FOR_EACH_VARIANT(X)                                                                  \
  X(double, 4, 1, 256, 2, false) X(double, 4, 1, 128, 3, true) X(double, 3, 2, 512, 1, false)   \
  X(double, 3, 2, 256, 2, false)                                                                \
  X(float, 5, 1, 256, 2, false) X(float, 5, 1, 128, 3, true) X(float, 4, 2, 512, 1, false)      \
  X(float, 4, 2, 256, 2, false)

}  // namespace b200q
