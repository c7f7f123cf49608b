// b200q — host side of the structure-specialised segment kernel (segk.cuh):
//   b200q_jit_compile : CUDA C++ source + in-memory headers -> sm_100a cubin through NVRTC
//                       (libnvrtc is dlopen'ed: no link-time dependency; works without a GPU);
//   b200q_seg_load    : cubin -> kernel handle (cudaLibraryLoadData, runtime API only);
//   b200q_seg_launch  : SkArgs + tensor maps + coefficient upload + launch (+ fixed-order
//                       reduction of the generator partial sums in adjoint mode).
// The hot path of the fused forward / reverse sweeps goes through here; rtile.cu (the record
// interpreter) stays for structures that are not worth a compilation.
#include <cuda.h>
#include <dlfcn.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200q.h"
#include "common.cuh"
#include "rtile_host.h"
#include "segk_args.h"

namespace b200q {

// ---- NVRTC through dlopen ----------------------------------------------------------------------
typedef int nvrtcResult_;
typedef struct _nvrtcProgram* nvrtcProgram_;
struct Nvrtc {
  void* h = nullptr;
  nvrtcResult_ (*CreateProgram)(nvrtcProgram_*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult_ (*CompileProgram)(nvrtcProgram_, int, const char* const*);
  nvrtcResult_ (*GetCUBINSize)(nvrtcProgram_, size_t*);
  nvrtcResult_ (*GetCUBIN)(nvrtcProgram_, char*);
  nvrtcResult_ (*GetProgramLogSize)(nvrtcProgram_, size_t*);
  nvrtcResult_ (*GetProgramLog)(nvrtcProgram_, char*);
  nvrtcResult_ (*DestroyProgram)(nvrtcProgram_*);
  const char* (*GetErrorString)(nvrtcResult_);
};

static Nvrtc* nvrtc() {
  static Nvrtc lib;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("B200Q_NVRTC");
    const char* names[] = {env ? env : "libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
      lib.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (lib.h) break;
    }
    if (!lib.h) return;
#define NVRTC_SYM(f)                                             \
  *(void**)(&lib.f) = dlsym(lib.h, "nvrtc" #f);                  \
  if (!lib.f) { dlclose(lib.h); lib.h = nullptr; return; }
    NVRTC_SYM(CreateProgram) NVRTC_SYM(CompileProgram) NVRTC_SYM(GetCUBINSize) NVRTC_SYM(GetCUBIN)
    NVRTC_SYM(GetProgramLogSize) NVRTC_SYM(GetProgramLog) NVRTC_SYM(DestroyProgram) NVRTC_SYM(GetErrorString)
#undef NVRTC_SYM
  });
  return lib.h ? &lib : nullptr;
}

// ---- tensor maps (same construction as rtile.cu: the tile of a segment as ONE box) -------------
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
static TmaEncodeFn sk_tma_encode_fn() {
  static TmaEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (TmaEncodeFn)p;
  }();
  return fn;
}

static void sk_build_tile_maps(SkArgs& a, int n, int dtype, uint64_t inmask, int L, void* v0, void* v1,
                               CUtensorMap* tm) {
  a.tma_rank = 0;
  TmaEncodeFn enc = sk_tma_encode_fn();
  if (!enc) return;
  const uint64_t ampB = dtype == B200Q_C128 ? 16 : 8;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  dims[0] = (ampB / 8) << L;
  box[0] = (cuuint32_t)dims[0];
  if (dims[0] > 256 || dims[0] * 8 < 16) return;
  int rank = 1;
  signed char lo[5] = {0, 0, 0, 0, 0}, len[5] = {0, 0, 0, 0, 0};
  for (int b = L; b < n;) {
    const bool tile = (inmask >> b) & 1;
    int e = b;
    while (e < n && (((inmask >> e) & 1) != 0) == tile) ++e;
    if (rank >= 5 || (tile && e - b > 8) || e - b > 31) return;
    dims[rank] = 1ull << (e - b);
    strides[rank - 1] = (1ull << b) * ampB;
    box[rank] = tile ? (cuuint32_t)dims[rank] : 1u;
    lo[rank] = (signed char)b;
    len[rank] = tile ? 0 : (signed char)(e - b);
    ++rank;
    b = e;
  }
  if (rank < 2) return;
  for (int v = 0; v < (v1 ? 2 : 1); ++v) {
    CUresult rc = enc(tm + v, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, v ? v1 : v0, dims, strides,
                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return;
  }
  a.tma_rank = rank;
  for (int r = 0; r < 5; ++r) { a.tma_lo[r] = lo[r]; a.tma_len[r] = len[r]; }
}

struct SegKernel {
  cudaLibrary_t lib;
  cudaKernel_t kern;
};

// out[row] = scale * sum over CTAs, fixed order
static __global__ void __launch_bounds__(256)
k_sk_final_reduce(const double* __restrict__ partials, double* __restrict__ out, const int ncta,
                  const double scale) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < ncta; i += blockDim.x) acc += partials[(size_t)blockIdx.x * ncta + i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = acc * scale;
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_jit_available(void) { return nvrtc() ? 1 : 0; }

int b200q_jit_compile(const char* source, const char* const* header_names, const char* const* header_sources,
                      int n_headers, int lineinfo, int maxreg, void** cubin_out, size_t* size_out) {
  B200Q_REQUIRE(source && cubin_out && size_out && n_headers >= 0, "jit_compile: null argument");
  Nvrtc* rt = nvrtc();
  B200Q_REQUIRE(rt, "jit_compile: libnvrtc.so.12 not found (set B200Q_NVRTC to its path)");
  nvrtcProgram_ prog = nullptr;
  nvrtcResult_ rc = rt->CreateProgram(&prog, source, "segk.cu", n_headers, header_sources, header_names);
  B200Q_REQUIRE(rc == 0, "jit_compile: nvrtcCreateProgram -> %s", rt->GetErrorString(rc));
  std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device"};
  if (lineinfo) opts.push_back("-lineinfo");
  // register cap: two resident 256-thread CTAs at <= 120 registers leave 4096 registers of an SM
  // free — room for the one-warp CTAs of the exchange's unpack kernel (remap.cu)
  std::string maxreg_opt;
  if (maxreg > 0) {
    maxreg_opt = "--maxrregcount=" + std::to_string(maxreg);
    opts.push_back(maxreg_opt.c_str());
  }
  rc = rt->CompileProgram(prog, (int)opts.size(), opts.data());
  if (rc != 0) {
    size_t ls = 0;
    rt->GetProgramLogSize(prog, &ls);
    std::string log(ls + 1, '\0');
    if (ls) rt->GetProgramLog(prog, &log[0]);
    if (log.size() > 3500) log.resize(3500);
    set_error("jit_compile: %s\n%s", rt->GetErrorString(rc), log.c_str());
    rt->DestroyProgram(&prog);
    return 3;
  }
  size_t sz = 0;
  rc = rt->GetCUBINSize(prog, &sz);
  if (rc != 0 || sz == 0) {
    set_error("jit_compile: no cubin (%s)", rt->GetErrorString(rc));
    rt->DestroyProgram(&prog);
    return 3;
  }
  char* buf = (char*)malloc(sz);
  rc = rt->GetCUBIN(prog, buf);
  rt->DestroyProgram(&prog);
  if (rc != 0) {
    free(buf);
    set_error("jit_compile: nvrtcGetCUBIN -> %s", rt->GetErrorString(rc));
    return 3;
  }
  *cubin_out = buf;
  *size_out = sz;
  return 0;
}

void b200q_jit_free(void* cubin) { free(cubin); }

int b200q_seg_load(const void* cubin, size_t size, void** handle_out) {
  B200Q_REQUIRE(cubin && size > 0 && handle_out, "seg_load: null argument");
  SegKernel* k = new SegKernel();
  cudaError_t e = cudaLibraryLoadData(&k->lib, cubin, nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->kern, k->lib, "sk_kernel");
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute((const void*)k->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  // fixed carve-out (all of it shared): the exchange's unpack kernel (remap.cu) asks for the same,
  // so its CTAs can join the two resident segment CTAs of an SM
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute((const void*)k->kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) {
    set_error("seg_load: %s", cudaGetErrorString(e));
    delete k;
    return 1;
  }
  *handle_out = k;
  return 0;
}

int b200q_seg_unload(void* handle) {
  if (!handle) return 0;
  SegKernel* k = (SegKernel*)handle;
  cudaLibraryUnload(k->lib);
  delete k;
  return 0;
}

int b200q_seg_launch(void* handle, void* vec0, void* vec1, int n, int dtype, int64_t batch,
                     const int* tile_bits, int T, int L, int RB, int minb, const int* ext_pos, int n_ext,
                     const double* coef_host, int n_coef, int coef_mode, int nslots, int write0,
                     uint64_t base_hi, uint64_t fix_mask, uint64_t fix_val, double scale, double* out_dev,
                     void* work, size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(handle && vec0 && tile_bits && coef_host, "seg_launch: null argument");
  B200Q_REQUIRE(dtype == B200Q_C64 || dtype == B200Q_C128, "seg_launch: unknown dtype %d", dtype);
  B200Q_REQUIRE(T >= 1 && T <= n && n <= B200Q_MAX_BITS && L >= 1 && L <= T && RB >= 1 && RB < T && batch >= 1,
                "seg_launch: bad geometry n=%d T=%d L=%d RB=%d", n, T, L, RB);
  B200Q_REQUIRE(n_ext >= 0 && n_ext <= 16 && n_coef >= 2 && nslots >= 0, "seg_launch: bad sizes");
  B200Q_REQUIRE(nslots == 0 || (vec1 && out_dev), "seg_launch: generator slots need a bra and an output");
  const size_t elem = dtype == B200Q_C128 ? 16 : 8;
  B200Q_REQUIRE(((elem << L) >= 16) && (((uintptr_t)vec0 | (uintptr_t)vec1) % 16 == 0),
                "seg_launch: bulk copies need 16-byte aligned runs");
  SkArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.write0 = write0; a.base_hi = base_hi; a.base_fix = fix_val;
  B200Q_REQUIRE((fix_val & ~fix_mask) == 0 && (n >= 64 || (fix_mask >> n) == 0),
                "seg_launch: fixed bits outside the mask / the state");
  uint64_t inmask = 0;
  for (int i = 0; i < T; ++i) {
    const int b = tile_bits[i];
    B200Q_REQUIRE(b >= 0 && b < n && !((inmask >> b) & 1), "seg_launch: bad tile bit %d", b);
    B200Q_REQUIRE(i < L ? b == i : (i == 0 || b > tile_bits[i - 1]),
                  "seg_launch: bits must be ascending with the first L equal to 0..L-1");
    inmask |= 1ull << b;
    if (i >= L) a.hi_bits[i - L] = (signed char)b;
  }
  B200Q_REQUIRE((inmask & fix_mask) == 0, "seg_launch: a tile bit cannot be held fixed");
  int no = 0, nfix = 0;
  for (int b = 0; b < n; ++b) {
    if ((inmask >> b) & 1) continue;
    if ((fix_mask >> b) & 1) { ++nfix; continue; }
    if (a.nruns > 0 && a.run_g[a.nruns - 1] + a.run_len[a.nruns - 1] == b) {
      a.run_len[a.nruns - 1]++;
    } else {
      B200Q_REQUIRE(a.nruns < 16, "seg_launch: too many runs of non-tile bits");
      a.run_s[a.nruns] = (signed char)no; a.run_len[a.nruns] = 1; a.run_g[a.nruns] = (signed char)b;
      a.nruns++;
    }
    ++no;
  }
  for (int e = 0; e < n_ext; ++e) {
    B200Q_REQUIRE(ext_pos[e] >= 0 && ext_pos[e] < 64, "seg_launch: bad external bit %d", ext_pos[e]);
    a.ext_pos[e] = (signed char)ext_pos[e];
  }
  a.ntiles = 1ull << (n - T - nfix);
  const int NV = vec1 ? 2 : 1;
  alignas(64) CUtensorMap tm[2];        // passed BY VALUE as a kernel parameter (SkMaps)
  memset(tm, 0, sizeof(tm));
  static const int tma_knob = getenv("B200Q_RT_TMA") ? atoi(getenv("B200Q_RT_TMA")) : 1;     // tuning knob
  if (tma_knob && batch == 1) sk_build_tile_maps(a, n, dtype, inmask, L, vec0, vec1, tm);
  // workspace: [coefficients | ... | partial sums]
  // coef_mode 0: one table, copied to shared memory by the kernel; 1: one table per batch
  // element; 2: the table is a kernel parameter (the kernel was compiled with SK_COEF_PARAM)
  const bool coef_batched = coef_mode == 1, coef_param = coef_mode == 2;
  B200Q_REQUIRE(coef_mode >= 0 && coef_mode <= 2, "seg_launch: bad coef_mode %d", coef_mode);
  const size_t coef_bytes = coef_param ? 0 : (size_t)n_coef * sizeof(double) * (coef_batched ? (size_t)batch : 1);
  B200Q_REQUIRE(work && coef_bytes + 1024 <= kTermRegion && work_bytes >= kWorkBytes,
                "seg_launch: coefficient table too large for the workspace");
  B200Q_REQUIRE(!coef_param || (size_t)n_coef * (elem / 2) <= 4000, "seg_launch: parameter table too large");
  char* w = (char*)work;
  if (!coef_param) B200Q_CHECK(cudaMemcpyAsync(w, coef_host, coef_bytes, cudaMemcpyHostToDevice, s));
  double* partials = (double*)(w + kTermRegion);
  const size_t pcap = (work_bytes - kTermRegion) / sizeof(double);
  const int threads = 1 << (T - RB);
  const size_t real_b = elem / 2;
  const size_t smem = ((size_t)NV * elem << T) + (coef_param ? 0 : (size_t)n_coef * real_b) + ((size_t)(1 << RB) + 1) * 8 +
                      (size_t)nslots * (threads / 32) * sizeof(double) + 64;
  B200Q_REQUIRE(smem <= 227 * 1024, "seg_launch: %zu bytes of shared memory needed", smem);
  uint64_t per_sm = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(minb, 1), (227 * 1024) / smem));
  static const int ctas_knob = getenv("B200Q_RT_CTAS") ? atoi(getenv("B200Q_RT_CTAS")) : 0;   // tuning knob
  if (ctas_knob > 0) per_sm = ctas_knob;
  const uint64_t cap = (uint64_t)sm_count() * per_sm;
  dim3 grid((unsigned)std::min<uint64_t>(a.ntiles, cap), (unsigned)batch);
  if (nslots > 0)
    B200Q_REQUIRE((size_t)batch * nslots * grid.x <= pcap, "seg_launch: workspace too small for %d slots", nslots);
  SegKernel* k = (SegKernel*)handle;
  const double* coef_dev = (const double*)w;
  long long bstride = coef_batched ? n_coef : 0;
  // parameter table in the kernel's precision (ignored by kernels compiled without SK_COEF_PARAM)
  std::vector<double> cf64;
  std::vector<float> cf32;
  void* cf_arg = (void*)coef_host;
  if (coef_param && dtype == B200Q_C64) {
    cf32.assign(coef_host, coef_host + n_coef);
    cf_arg = cf32.data();
  }
  void* args[] = {&a, &vec0, &vec1, &coef_dev, &bstride, tm, &partials, cf_arg};
  B200Q_CHECK(cudaLaunchKernel((const void*)k->kern, grid, dim3(threads), args, smem, s));
  if (nslots > 0) {
    k_sk_final_reduce<<<(unsigned)(batch * nslots), 256, 0, s>>>(partials, out_dev, (int)grid.x, scale);
    B200Q_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
