// FP64 issue-rate micro-benchmark of the record loop of k_rtile: the single-qubit block handlers
// (d1_uniform / d1_regctl) on 16 amplitudes held in registers, no global traffic inside the loop.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pennylane_b200/csrc tools/micro/fp64_micro.cu -o /tmp/fp64_micro && /tmp/fp64_micro
#include <cstdio>
#include <cuda_runtime.h>
#include "rtile.cuh"

using namespace b200q;
using K = RtKernel<double, 4, 1, 256, false>;

template <int VAR>
__global__ void __launch_bounds__(256, 2) k_micro(double2* out, const double2* mats, int iters) {
  __shared__ double2 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = mats[threadIdx.x];
  __syncthreads();
  const unsigned ms = (unsigned)__cvta_generic_to_shared(sm);
  double2 A[1][16];
#pragma unroll
  for (int k = 0; k < 16; ++k) A[0][k] = make_double2(1.0 + threadIdx.x * 1e-3 + k, 0.5 - k * 1e-2);
  for (int it = 0; it < iters; ++it) {
    if (VAR == 0) {            // four records, no register control
      K::d1_uniform<0>(A, ms); K::d1_uniform<1>(A, ms + 64); K::d1_uniform<2>(A, ms + 128); K::d1_uniform<3>(A, ms + 192);
    } else if (VAR == 1) {     // four records with a control on another register bit (two halves)
      K::d1_regctl<0, 1>(A, ms, 0, 2); K::d1_regctl<1, 2>(A, ms, 4, 6); K::d1_regctl<2, 3>(A, ms, 8, 10); K::d1_regctl<3, 0>(A, ms, 12, 14);
    } else {                   // textbook y = M x on pairs (not in place), for comparison
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 m0 = sm[q * 4], m1 = sm[q * 4 + 1], m2 = sm[q * 4 + 2], m3 = sm[q * 4 + 3];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if ((k >> q) & 1) continue;
          const double2 x0 = A[0][k], x1 = A[0][k | (1 << q)];
          double2 y0, y1;
          y0.x = m0.x * x0.x - m0.y * x0.y + m1.x * x1.x - m1.y * x1.y;
          y0.y = m0.x * x0.y + m0.y * x0.x + m1.x * x1.y + m1.y * x1.x;
          y1.x = m2.x * x0.x - m2.y * x0.y + m3.x * x1.x - m3.y * x1.y;
          y1.y = m2.x * x0.y + m2.y * x0.x + m3.x * x1.y + m3.y * x1.x;
          A[0][k] = y0; A[0][k | (1 << q)] = y1;
        }
      }
    }
  }
  double2 acc = make_double2(0, 0);
#pragma unroll
  for (int k = 0; k < 16; ++k) { acc.x += A[0][k].x; acc.y += A[0][k].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int VAR> void run(const char* name, double2* out, double2* mats, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 2;
  k_micro<VAR><<<grid, 256>>>(out, mats, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_micro<VAR><<<grid, 256>>>(out, mats, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // 4 records x 8 pairs x 16 FP64 instructions per thread and iteration
  const double fp64_thread_instr = (double)iters * 4 * 8 * 16;
  const double flops = fp64_thread_instr * grid * 256 * 2;     // count every DMUL/DFMA as 2 flop (upper bound: 2 of 16 are DMUL)
  printf("%-28s %8.3f ms  %7.2f TFLOP/s (DFMA-equivalent)  err=%s\n", name, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double2 *out, *mats, h[64];
  for (int i = 0; i < 64; ++i) h[i] = make_double2(0.6 + 0.001 * i, 0.3 - 0.002 * i);   // contractive enough
  for (int i = 0; i < 64; ++i) { h[i].x *= 0.7; h[i].y *= 0.7; }
  cudaMalloc(&out, sizeof(double2) * 148 * 2 * 256); cudaMalloc(&mats, sizeof(h));
  cudaMemcpy(mats, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("d1_uniform x4", out, mats, 4000);
  run<1>("d1_regctl x4 (two halves)", out, mats, 4000);
  run<2>("textbook y = M x", out, mats, 4000);
  return 0;
}
