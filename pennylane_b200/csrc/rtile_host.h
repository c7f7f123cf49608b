// b200q — host-side interface of the register-tiled fused segment kernel (rtile.cu is its own
// translation unit so that the large interpreter kernels compile in parallel with api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200q {

struct RtOp;   // 64-byte record, rtile.cuh

static const size_t kWorkBytes = 32ull << 20;      // 32 MiB scratch
static const size_t kTermRegion = 4ull << 20;      // first 4 MiB: uploaded term tables

// (dtype, nvec) -> tile bits T, register bits RB, threads per CTA
void rtile_geom(int dtype, int nvec, int& T, int& RB, int& threads);

int rtile_dispatch(void* v0, void* v1, int n, int dtype, int64_t batch, const int* tile_bits,
                   int Tn, int L, const RtOp* ops_host, int nops, const double2* mats_host,
                   int nmat, int nslots, int write0, uint64_t base_hi, double scale,
                   double* out_dev, void* work, size_t work_bytes, cudaStream_t s,
                   int mat_batched = 0);

}  // namespace b200q
