// FP64 micro-benchmarks that decide the shape of the specialised segment kernel (round 2).
//   (1) k_dfma<ILP>: pure DFMA chains -> the FP64 pipe ceiling at several occupancies / ILP.
//   (2) k_forms<FORM>: one "record" = a 2x2 block on a register bit of 16 register-resident
//       complex128 amplitudes, straight-line (what a structure-specialised kernel would run),
//       for the normalised forms of the block:
//         F16  out = M x, all four entries complex                              (16 FP64 / pair)
//         F12  out0 = x0 + a x1 ; out1 = b x0 + g x1                            (12)
//         F8   out0 = x0 + r x1 ; out1 = p (s x0 + x1), r s real, p complex     ( 8)  RZ.RY
//         F4   out0 = x0 + r x1 ; out1 = s x0 + x1                              ( 4)  RY
//         F12S F12 followed by a per-thread conditional swap of the outputs (folded CNOT whose
//              control sits on a lane bit), as SELs
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/micro/fp64_forms.cu -o /tmp/fp64_forms && /tmp/fp64_forms
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP> void run_dfma(double* out, int threads, int bps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * bps, iters = 20000 / ILP + 1;
  k_dfma<ILP><<<grid, threads>>>(out, 10, 0.999, 1e-3);
  cudaEventRecord(e0);
  k_dfma<ILP><<<grid, threads>>>(out, iters, 0.999, 1e-3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)iters * 8 * ILP * grid * threads;
  printf("dfma ILP=%2d threads=%4d blocks/SM=%d warps/SM=%2d : %7.2f TFLOP/s  %6.1f DFMA/clk/SM @1.9GHz\n", ILP,
         threads, bps, threads * bps / 32, 2 * n / ms / 1e9, n / (ms * 1e-3) / 148 / 1.9e9);
}

struct C2 { double x, y; };

template <int FORM, int Q>
__device__ __forceinline__ void rec(C2 (&A)[16], const double* __restrict__ m, const bool swp) {
  // volatile shared-memory loads: the table is loop invariant and ptxas would otherwise hoist the
  // coefficients of every record out of the tile loop (64 registers)
  double c[8];
  const unsigned ma = (unsigned)__cvta_generic_to_shared(m);
#pragma unroll
  for (int i = 0; i < (FORM == 4 ? 2 : FORM == 8 ? 4 : FORM == 16 ? 8 : 6); i += 2)
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c[i]), "=d"(c[i + 1]) : "r"(ma + 8u * i));
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if ((k >> Q) & 1) continue;
    C2& x0 = A[k];
    C2& x1 = A[k | (1 << Q)];
    if (FORM == 16) {
      double tx0 = -c[1] * x0.y, ty0 = c[1] * x0.x, tx1 = -c[7] * x1.y, ty1 = c[7] * x1.x;
      tx0 = fma(c[2], x1.x, tx0); tx0 = fma(-c[3], x1.y, tx0);
      ty0 = fma(c[3], x1.x, ty0); ty0 = fma(c[2], x1.y, ty0);
      tx1 = fma(c[4], x0.x, tx1); tx1 = fma(-c[5], x0.y, tx1);
      ty1 = fma(c[5], x0.x, ty1); ty1 = fma(c[4], x0.y, ty1);
      x0.x = fma(c[0], x0.x, tx0); x0.y = fma(c[0], x0.y, ty0);
      x1.x = fma(c[6], x1.x, tx1); x1.y = fma(c[6], x1.y, ty1);
    } else if (FORM == 12 || FORM == 13) {
      // out1 first (needs the old x0), then out0
      double tx1 = -c[5] * x1.y, ty1 = c[5] * x1.x;
      tx1 = fma(c[2], x0.x, tx1); tx1 = fma(-c[3], x0.y, tx1);
      ty1 = fma(c[3], x0.x, ty1); ty1 = fma(c[2], x0.y, ty1);
      double ox = fma(c[0], x1.x, x0.x); ox = fma(-c[1], x1.y, ox);
      double oy = fma(c[1], x1.x, x0.y); oy = fma(c[0], x1.y, oy);
      double px = fma(c[4], x1.x, tx1), py = fma(c[4], x1.y, ty1);
      if (FORM == 13) {
        x0.x = swp ? px : ox; x0.y = swp ? py : oy;
        x1.x = swp ? ox : px; x1.y = swp ? oy : py;
      } else {
        x0.x = ox; x0.y = oy; x1.x = px; x1.y = py;
      }
    } else if (FORM == 8) {
      const double wx = fma(c[1], x0.x, x1.x), wy = fma(c[1], x0.y, x1.y);
      x0.x = fma(c[0], x1.x, x0.x); x0.y = fma(c[0], x1.y, x0.y);
      const double t = -c[3] * wy;
      const double u = c[3] * wx;
      x1.x = fma(c[2], wx, t); x1.y = fma(c[2], wy, u);
    } else {
      const double wx = fma(c[1], x0.x, x1.x), wy = fma(c[1], x0.y, x1.y);
      x0.x = fma(c[0], x1.x, x0.x); x0.y = fma(c[0], x1.y, x0.y);
      x1.x = wx; x1.y = wy;
    }
  }
}

template <int FORM, int MINB>
__global__ void __launch_bounds__(256, MINB) k_forms(C2* out, const double* mats, int iters) {
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = mats[threadIdx.x];
  __syncthreads();
  C2 A[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { A[k].x = 1.0 + threadIdx.x * 1e-3 + k; A[k].y = 0.5 - k * 1e-2; }
  const bool swp = (threadIdx.x >> 2) & 1;
  for (int it = 0; it < iters; ++it) {
    rec<FORM, 0>(A, sm, swp); rec<FORM, 1>(A, sm + 8, swp); rec<FORM, 2>(A, sm + 16, swp); rec<FORM, 3>(A, sm + 24, swp);
  }
  C2 acc = {0, 0};
#pragma unroll
  for (int k = 0; k < 16; ++k) { acc.x += A[k].x; acc.y += A[k].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int FORM, int MINB> void run_forms(const char* name, C2* out, double* mats) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * MINB, iters = 4000;
  k_forms<FORM, MINB><<<grid, 256>>>(out, mats, 10);
  cudaEventRecord(e0);
  k_forms<FORM, MINB><<<grid, 256>>>(out, mats, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // records applied to 2^30 amplitudes per ms:  amplitudes touched per launch = grid*256*16 per record
  const double amp_rec = (double)iters * 4 * grid * 256 * 16;
  const int ops = FORM == 13 ? 12 : FORM;
  printf("%-22s CTAs/SM=%d : %8.3f ms  -> %6.3f ms per record over 2^30 amplitudes, %6.2f TFLOP/s (2 flop per FP64 instr)  %s\n", name,
         MINB, ms, ms / (amp_rec / 1073741824.0), amp_rec / 2 * ops * 2 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* outd; cudaMalloc(&outd, sizeof(double) * 148 * 4 * 1024);
  run_dfma<1>(outd, 256, 2); run_dfma<2>(outd, 256, 2); run_dfma<4>(outd, 256, 2); run_dfma<8>(outd, 256, 2);
  run_dfma<16>(outd, 256, 2);
  run_dfma<8>(outd, 128, 1); run_dfma<8>(outd, 256, 1); run_dfma<8>(outd, 256, 3); run_dfma<8>(outd, 256, 4);
  run_dfma<4>(outd, 1024, 1); run_dfma<4>(outd, 1024, 2); run_dfma<2>(outd, 1024, 2);
  C2* out; double *mats, h[64];
  for (int i = 0; i < 64; ++i) h[i] = (i % 8 == 0 || i % 8 == 6) ? 0.62 : 0.05 * ((i % 5) - 2);   // contractive
  cudaMalloc(&out, sizeof(C2) * 148 * 4 * 256); cudaMalloc(&mats, sizeof(h));
  cudaMemcpy(mats, h, sizeof(h), cudaMemcpyHostToDevice);
  run_forms<16, 2>("F16 generic", out, mats); run_forms<16, 3>("F16 generic", out, mats);
  run_forms<12, 2>("F12 unit pivot", out, mats); run_forms<12, 3>("F12 unit pivot", out, mats);
  run_forms<13, 2>("F12 + lane select", out, mats); run_forms<13, 3>("F12 + lane select", out, mats);
  run_forms<8, 2>("F8 RZ.RY", out, mats); run_forms<8, 3>("F8 RZ.RY", out, mats);
  run_forms<4, 2>("F4 RY", out, mats); run_forms<4, 3>("F4 RY", out, mats);
  return 0;
}
