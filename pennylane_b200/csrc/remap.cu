// b200q — K9: data movement of the qubit-remapping exchange.
//
//   b200q_remap_copy    copy-engine copies (cudaMemcpy[2D]Async): used for the NVLink PUSH of a
//                       slab piece into a partner's staging buffer (peer destination).
//   b200q_remap_unpack  staging buffer -> state (or state -> staging) as a TMA kernel.  Why a
//                       kernel: beside a fused segment launch (which saturates HBM from every SM)
//                       a copy-engine device-to-device copy was measured at 0.4 TB/s for
//                       contiguous runs, and a PITCHED 2D copy did not start before the next
//                       launch boundary (tools/micro_corun.py, profiles/r2_corun.txt); memory
//                       requests issued from the SMs get the same arbitration as the segment
//                       kernel's.  Why TMA: the kernel has to fit into what two resident segment
//                       CTAs leave of an SM — one warp, < 32 registers per thread, 64 KiB of shared
//                       memory: ONE thread per CTA keeps four 16 KiB bulk copies in flight
//                       (cp.async.bulk global -> shared with mbarrier completion, then shared ->
//                       global as a bulk group).
//   b200q_stream_write32 / _wait_geq32   flags of the exchange protocol as stream memory
//                       operations: no kernel, nothing waits for an SM.
// The reference has no analogue (default.qubit holds one array on one device).
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/b200q.h"
#include "common.cuh"

namespace b200q {

constexpr int kRemapChunk = 16384;     // bytes per bulk copy
constexpr int kRemapStages = 4;
constexpr int kRegsThreads = 96;

__device__ __forceinline__ unsigned rm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32, 1)
k_remap_tma(char* __restrict__ dst, const unsigned long long dst_pitch, const char* __restrict__ src,
            const unsigned long long src_pitch, const unsigned long long chunks_per_run,
            const unsigned long long total_chunks) {
  extern __shared__ __align__(128) unsigned char rm_smem[];          // kRemapStages * kRemapChunk
  __shared__ __align__(8) unsigned long long full[kRemapStages];
  if (threadIdx.x != 0) return;                                      // one thread drives the CTA
  for (int s = 0; s < kRemapStages; ++s)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rm_smem_u32(&full[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const unsigned long long first = blockIdx.x, step = gridDim.x;
  if (first >= total_chunks) return;
  const unsigned long long mine = (total_chunks - first + step - 1) / step;

  auto issue_load = [&](unsigned long long i) {
    const unsigned long long c = first + i * step;
    const unsigned long long row = c / chunks_per_run, col = c - row * chunks_per_run;
    const char* g = src + row * src_pitch + col * kRemapChunk;
    const int s = (int)(i % kRemapStages);
    const unsigned bar = rm_smem_u32(&full[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kRemapChunk) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rm_smem_u32(rm_smem + s * kRemapChunk)), "l"(g), "r"(kRemapChunk), "r"(bar) : "memory");
  };

  const unsigned long long pro = mine < (unsigned long long)kRemapStages ? mine : (unsigned long long)kRemapStages;
  for (unsigned long long i = 0; i < pro; ++i) issue_load(i);
  for (unsigned long long i = 0; i < mine; ++i) {
    const int s = (int)(i % kRemapStages);
    const unsigned parity = (unsigned)((i / kRemapStages) & 1ull);
    const unsigned bar = rm_smem_u32(&full[s]);
    asm volatile(
        "{\n.reg .pred p;\nRM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RM_DONE_%=;\nbra RM_WAIT_%=;\nRM_DONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
    const unsigned long long c = first + i * step;
    const unsigned long long row = c / chunks_per_run, col = c - row * chunks_per_run;
    char* g = dst + row * dst_pitch + col * kRemapChunk;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(g), "r"(rm_smem_u32(rm_smem + s * kRemapChunk)), "r"(kRemapChunk) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // the stage of the PREVIOUS store may be refilled once that store has read its data
    if (i >= 1 && i - 1 + kRemapStages < mine) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      issue_load(i - 1 + kRemapStages);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The same copy through REGISTERS (96 threads, four 16-byte loads in flight per thread, two CTAs
// per SM).  Beside a segment whose rounds keep the shared-memory port busy (five rounds = four
// transpositions per tile) the bulk copies of k_remap_tma were measured not to progress at all until
// the segment's CTAs left (7.6 ms per GiB, tools/micro_corun.py); this kernel touches no shared
// memory: 1.8-2.0 ms per GiB beside any segment (the TMA kernel: 1.0 ms beside a two-round one).
__global__ void __launch_bounds__(kRegsThreads, 2)
k_remap_regs(char* __restrict__ dst, const unsigned long long dst_pitch, const char* __restrict__ src,
             const unsigned long long src_pitch, const unsigned long long run_bytes, const unsigned long long count) {
  const unsigned long long per_run = run_bytes / 64, total = per_run * count;
  for (unsigned long long i = (unsigned long long)blockIdx.x * kRegsThreads + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * kRegsThreads) {
    const unsigned long long row = i / per_run, col = i - row * per_run;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + row * src_pitch + col * 64);
    uint4* d4 = reinterpret_cast<uint4*>(dst + row * dst_pitch + col * 64);
    const uint4 a = __ldcs(s4), b = __ldcs(s4 + 1), c = __ldcs(s4 + 2), d = __ldcs(s4 + 3);
    __stcs(d4, a); __stcs(d4 + 1, b); __stcs(d4 + 2, c); __stcs(d4 + 3, d);
  }
}

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamValue32Fn stream_value_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    p = nullptr;
  return (StreamValue32Fn)p;
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_remap_copy(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t run_bytes,
                     size_t count, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(dst && src && run_bytes > 0 && count > 0, "remap_copy: null / empty argument");
  B200Q_REQUIRE(dst_pitch >= run_bytes && src_pitch >= run_bytes, "remap_copy: pitch smaller than a run");
  if (count == 1 || (dst_pitch == run_bytes && src_pitch == run_bytes)) {
    B200Q_CHECK(cudaMemcpyAsync(dst, src, run_bytes * count, cudaMemcpyDefault, s));
    return 0;
  }
  static const size_t max_pitch = [] {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return (size_t)0x7fffffff;
    return (size_t)p.memPitch;
  }();
  if (dst_pitch <= max_pitch && src_pitch <= max_pitch) {
    B200Q_CHECK(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, run_bytes, count, cudaMemcpyDefault, s));
    return 0;
  }
  B200Q_REQUIRE(count <= 4096, "remap_copy: %zu runs with a pitch above the 2D-copy limit", count);
  for (size_t i = 0; i < count; ++i)
    B200Q_CHECK(cudaMemcpyAsync((char*)dst + i * dst_pitch, (const char*)src + i * src_pitch, run_bytes,
                                cudaMemcpyDefault, s));
  return 0;
}

int b200q_remap_unpack(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t run_bytes,
                       size_t count, int mode, int ctas, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(dst && src && run_bytes > 0 && count > 0, "remap_unpack: null / empty argument");
  B200Q_REQUIRE(dst_pitch >= run_bytes && src_pitch >= run_bytes, "remap_unpack: pitch smaller than a run");
  B200Q_REQUIRE(mode == B200Q_UNPACK_TMA || mode == B200Q_UNPACK_REGS, "remap_unpack: unknown mode %d", mode);
  const size_t unit = mode == B200Q_UNPACK_TMA ? (size_t)kRemapChunk : 64;
  if (run_bytes % unit != 0 || (((uintptr_t)dst | (uintptr_t)src | dst_pitch | src_pitch) & 15) != 0)
    return b200q_remap_copy(dst, dst_pitch, src, src_pitch, run_bytes, count, stream);
  static bool configured = false;
  if (!configured) {
    // the same shared-memory carve-out as the segment kernels: CTAs of both must be able to
    // share an SM
    B200Q_CHECK(cudaFuncSetAttribute(k_remap_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kRemapStages * kRemapChunk));
    B200Q_CHECK(cudaFuncSetAttribute(k_remap_tma, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    B200Q_CHECK(cudaFuncSetAttribute(k_remap_regs, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  if (mode == B200Q_UNPACK_REGS) {
    const unsigned grid = (unsigned)(ctas > 0 ? ctas : 2 * sm_count());
    k_remap_regs<<<grid, kRegsThreads, 0, s>>>((char*)dst, (unsigned long long)dst_pitch, (const char*)src,
                                               (unsigned long long)src_pitch, (unsigned long long)run_bytes,
                                               (unsigned long long)count);
  } else {
    const unsigned long long per_run = run_bytes / kRemapChunk, total = per_run * count;
    const unsigned long long cap = (unsigned long long)(ctas > 0 ? ctas : sm_count());
    const unsigned grid = (unsigned)(total < cap ? total : cap);
    k_remap_tma<<<grid, 32, kRemapStages * kRemapChunk, s>>>((char*)dst, (unsigned long long)dst_pitch,
                                                             (const char*)src, (unsigned long long)src_pitch,
                                                             per_run, total);
  }
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_stream_write32(void* addr, uint32_t value, void* stream) {
  static StreamValue32Fn fn = stream_value_fn("cuStreamWriteValue32");
  B200Q_REQUIRE(fn, "stream_write32: cuStreamWriteValue32 is not available in this driver");
  B200Q_REQUIRE(addr && ((uintptr_t)addr & 3) == 0, "stream_write32: unaligned / null address");
  CUresult rc = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
  B200Q_REQUIRE(rc == CUDA_SUCCESS, "stream_write32: driver error %d", (int)rc);
  return 0;
}

int b200q_stream_wait_geq32(void* addr, uint32_t value, void* stream) {
  static StreamValue32Fn fn = stream_value_fn("cuStreamWaitValue32");
  B200Q_REQUIRE(fn, "stream_wait_geq32: cuStreamWaitValue32 is not available in this driver");
  B200Q_REQUIRE(addr && ((uintptr_t)addr & 3) == 0, "stream_wait_geq32: unaligned / null address");
  CUresult rc = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  B200Q_REQUIRE(rc == CUDA_SUCCESS, "stream_wait_geq32: driver error %d", (int)rc);
  return 0;
}

}  // extern "C"
