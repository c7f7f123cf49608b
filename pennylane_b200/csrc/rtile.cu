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
  const size_t ops_bytes = (size_t)nops * sizeof(RtOp);
  const size_t mat_bytes = (size_t)nmat * sizeof(double2) * (mat_batched ? (size_t)batch : 1);
  const long long mat_bstride = mat_batched ? nmat : 0;
  B200Q_REQUIRE(work && ops_bytes + mat_bytes + 512 <= kTermRegion && work_bytes >= kWorkBytes,
                "rtile: segment tables too large for the workspace");
  char* w = (char*)work;
  B200Q_CHECK(cudaMemcpyAsync(w, ops_host, ops_bytes, cudaMemcpyHostToDevice, s));
  const size_t moff = (ops_bytes + 255) & ~(size_t)255;
  if (nmat) B200Q_CHECK(cudaMemcpyAsync(w + moff, mats_host, mat_bytes, cudaMemcpyHostToDevice, s));
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
