// b200q — host dispatch of the register-tiled fused segment kernel (rtile.cuh).  Separate
// translation unit: the interpreter kernels are the bulk of the library's compile time.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "rtile_launch.cuh"

namespace b200q {

#define RT_EXTERN(T, RB, NV, TH, MINB, WS) \
  extern template int rtile_launch<T, RB, NV, TH, MINB, WS>(RT_LAUNCH_ARGS);
RT_FOR_EACH_VARIANT(RT_EXTERN)
#undef RT_EXTERN

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encode_fn() {
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

// The tile of a segment as ONE box of a tensor map over the state: dim 0 = the contiguous run of
// 2^L amplitudes (as 8-byte elements), then one dim per group of consecutive tile bits (box =
// whole dim) or non-tile bits (box = 1, coordinate = those bits of the tile base).  Fills
// a.tma_rank / tma_lo / tma_len and both maps; leaves tma_rank = 0 when the shape does not fit
// (more than 5 dims, a tile group wider than 8 bits, runs shorter than 16 bytes).
static void build_tile_maps(RtArgs& a, int n, int dtype, uint64_t inmask, int L, void* v0, void* v1,
                            CUtensorMap& tm0, CUtensorMap& tm1) {
  a.tma_rank = 0;
  TmaEncodeFn enc = tma_encode_fn();
  if (!enc) return;
  const uint64_t ampB = dtype == B200Q_C128 ? 16 : 8;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  dims[0] = (ampB / 8) << L;
  box[0] = (cuuint32_t)dims[0];
  if (dims[0] > 256 || dims[0] * 8 < 16) return;
  int rank = 1;
  int8_t lo[5] = {0, 0, 0, 0, 0}, len[5] = {0, 0, 0, 0, 0};
  for (int b = L; b < n;) {
    const bool tile = (inmask >> b) & 1;
    int e = b;
    while (e < n && (((inmask >> e) & 1) != 0) == tile) ++e;
    if (rank >= 5 || (tile && e - b > 8) || e - b > 31) return;
    dims[rank] = 1ull << (e - b);
    strides[rank - 1] = (1ull << b) * ampB;
    box[rank] = tile ? (cuuint32_t)dims[rank] : 1u;
    lo[rank] = (int8_t)b;
    len[rank] = tile ? 0 : (int8_t)(e - b);
    ++rank;
    b = e;
  }
  if (rank < 2) return;
  for (int v = 0; v < (v1 ? 2 : 1); ++v) {
    CUresult rc = enc(v ? &tm1 : &tm0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, v ? v1 : v0, dims,
                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return;
  }
  a.tma_rank = rank;
  for (int r = 0; r < 5; ++r) { a.tma_lo[r] = lo[r]; a.tma_len[r] = len[r]; }
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// tuning knob, read once.  Forward kernel: 0 = warp-specialised (128 consumer threads + one
// producer warp, 3 CTAs/SM, landing buffer filled by bulk async copies), 1 = 256 threads x
// 2 CTAs/SM with the next tile fetched into the idle transpose buffer.
static int rtile_variant() {
  static const int v = env_int("B200Q_RT_VARIANT", 1);
  return v;
}

// adjoint kernel (two vectors): 0 = 512 threads x 1 CTA/SM, 1 = 256 threads x 2 CTAs/SM
static int rtile_adj_variant() {
  static const int v = env_int("B200Q_RT_ADJ_VARIANT", 0);
  return v;
}

// geometry of the register-tiled kernel: (dtype, nvec) -> (T, RB, THREADS = consumer threads)
void rtile_geom(int dtype, int nvec, int& T, int& RB, int& threads) {
  if (nvec <= 1) {
    threads = rtile_variant() == 0 ? 128 : 256; RB = dtype == B200Q_C128 ? 4 : 5;
  } else { threads = rtile_adj_variant() == 0 ? 512 : 256; RB = dtype == B200Q_C128 ? 3 : 4; }
  T = (threads == 128 ? 7 : threads == 256 ? 8 : 9) + RB;
}

int rtile_dispatch(void* v0, void* v1, int n, int dtype, int64_t batch, const int* tile_bits,
                   int Tn, int L, const RtOp* ops_host, int nops, const double2* mats_host,
                   int nmat, int nslots, int write0, uint64_t base_hi, double scale,
                   double* out_dev, void* work, size_t work_bytes, cudaStream_t s, int mat_batched) {
  int gT, gRB, gTh;
  rtile_geom(dtype, v1 ? 2 : 1, gT, gRB, gTh);
  B200Q_REQUIRE(Tn == gT && Tn <= n && L >= 0 && L <= Tn, "rtile: T=%d (need %d) L=%d n=%d", Tn, gT, L, n);
  B200Q_REQUIRE(nops >= 1 && nops <= 2048 && nmat >= 0 && nslots >= 0, "rtile: bad nops=%d nmat=%d", nops, nmat);
  B200Q_REQUIRE(ops_host[0].kind == RT_ROUND, "rtile: the first record must be a round");
  B200Q_REQUIRE(nslots == 0 || (v1 && out_dev), "rtile: generator slots need a bra and an output");
  RtArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.T = Tn; a.L = L; a.nops = nops; a.nmat = nmat; a.nslots = nslots; a.write0 = write0;
  a.base_hi = base_hi;
  for (int i = 0; i < nops; ++i) {
    if ((ops_host[i].kind & 0xff) == RT_DENSE1 && a.nd1 < 32) a.nd1++;
    if (ops_host[i].kind == RT_ROUND) {
      a.last_round = i;
      a.nrounds++;
      uint32_t seen = 0;
      const int tb = Tn - gRB;
      for (int b = 0; b < gRB; ++b) seen |= 1u << ops_host[i].u.r.rbits[b];
      for (int b = 0; b < tb; ++b) seen |= 1u << ops_host[i].u.r.tbits[b];
      B200Q_REQUIRE(seen == (1u << Tn) - 1u, "rtile: round %d is not a permutation of the tile bits", i);
    }
  }
  uint64_t inmask = 0;
  for (int i = 0; i < Tn; ++i) {
    const int b = tile_bits[i];
    B200Q_REQUIRE(b >= 0 && b < n && !((inmask >> b) & 1), "rtile: bad tile bit %d", b);
    B200Q_REQUIRE(i < L ? b == i : (i == 0 || b > tile_bits[i - 1]),
                  "rtile: bits must be ascending with the first L equal to 0..L-1");
    inmask |= 1ull << b;
    if (i >= L) a.hi_bits[i - L] = (int8_t)b;
  }
  // runs of consecutive non-tile bits: tile-number bits [s, s+len) -> global bits [g, g+len)
  int no = 0;
  for (int b = 0; b < n; ++b) {
    if ((inmask >> b) & 1) continue;
    if (a.nruns > 0 && a.run_g[a.nruns - 1] + a.run_len[a.nruns - 1] == b) {
      a.run_len[a.nruns - 1]++;
    } else {
      B200Q_REQUIRE(a.nruns < 16, "rtile: too many runs of non-tile bits");
      a.run_s[a.nruns] = (int8_t)no; a.run_len[a.nruns] = 1; a.run_g[a.nruns] = (int8_t)b;
      a.nruns++;
    }
    ++no;
  }
  a.ntiles = 1ull << (n - Tn);
  // bulk-copy prefetch of the next tile (default on); needs >= 16-byte contiguous runs
  static const int pf_knob = env_int("B200Q_RT_PREFETCH", 1);                  // tuning knob
  const size_t elem = dtype == B200Q_C128 ? 16 : 8;
  a.prefetch = ((pf_knob || (rtile_variant() == 0 && !v1)) && ((elem << L) >= 16) && (((uintptr_t)v0 | (uintptr_t)v1) % 16 == 0)) ? 1 : 0;
  // one TMA instruction per tile when the tile is a box of a <= 5-dim view of the state
  // (B200Q_RT_TMA=0: always one bulk copy per run); broadcast states keep the per-run copies
  CUtensorMap tm0, tm1;
  memset(&tm0, 0, sizeof(tm0));
  memset(&tm1, 0, sizeof(tm1));
  static const int tma_knob = env_int("B200Q_RT_TMA", 1);                      // tuning knob
  if (tma_knob && a.prefetch && batch == 1 && !v1 && rtile_variant() != 0)
    build_tile_maps(a, n, dtype, inmask, L, v0, v1, tm0, tm1);
  const size_t ops_bytes = (size_t)nops * sizeof(RtOp);
  const size_t mat_bytes = (size_t)nmat * sizeof(double2) * (mat_batched ? (size_t)batch : 1);
  const long long mat_bstride = mat_batched ? nmat : 0;
  B200Q_REQUIRE(work && ops_bytes + mat_bytes + 512 <= kTermRegion && work_bytes >= kWorkBytes,
                "rtile: segment tables too large for the workspace");
  char* w = (char*)work;
  B200Q_CHECK(cudaMemcpyAsync(w, ops_host, ops_bytes, cudaMemcpyHostToDevice, s));
  const size_t moff = (ops_bytes + 255) & ~(size_t)255;
  if (nmat) B200Q_CHECK(cudaMemcpyAsync(w + moff, mats_host, mat_bytes, cudaMemcpyHostToDevice, s));
  if (a.tma_rank > 0) {
    static_assert(kRtTensorMapOffset + 2 * sizeof(CUtensorMap) <= kTermRegion, "tensor maps outside the table region");
    CUtensorMap both[2] = {tm0, tm1};
    B200Q_CHECK(cudaMemcpyAsync(w + kRtTensorMapOffset, both, sizeof(both), cudaMemcpyHostToDevice, s));
  }
  double* partials = (double*)(w + kTermRegion);
  const size_t pcap = (work_bytes - kTermRegion) / sizeof(double);
  const RtOp* od = (const RtOp*)w;
  const double2* md = (const double2*)(w + moff);
  const bool ws = rtile_variant() == 0;
  if (ws && !v1)
    B200Q_REQUIRE(a.prefetch, "rtile: the warp-specialised kernel needs 16-byte aligned runs (L=%d)", L);
#define RT_GO(T, RB, NV, TH, MINB, WS) \
  return rtile_launch<T, RB, NV, TH, MINB, WS>(v0, v1, a, batch, od, md, mat_bstride, nslots, scale, out_dev, partials, pcap, s)
  const bool adj2 = rtile_adj_variant() != 0;
  if (dtype == B200Q_C128) {
    if (!v1 && ws) RT_GO(double, 4, 1, 128, 3, true);
    if (!v1) RT_GO(double, 4, 1, 256, 2, false);
    if (adj2) RT_GO(double, 3, 2, 256, 2, false);
    RT_GO(double, 3, 2, 512, 1, false);
  }
  if (dtype == B200Q_C64) {
    if (!v1 && ws) RT_GO(float, 5, 1, 128, 3, true);
    if (!v1) RT_GO(float, 5, 1, 256, 2, false);
    if (adj2) RT_GO(float, 4, 2, 256, 2, false);
    RT_GO(float, 4, 2, 512, 1, false);
  }
#undef RT_GO
  set_error("unknown dtype %d", dtype);
  return 2;
}

}  // namespace b200q
