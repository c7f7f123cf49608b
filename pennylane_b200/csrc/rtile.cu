// b200q — host dispatch of the register-tiled fused segment kernel (rtile.cuh).  Separate
// translation unit: the interpreter kernels are the bulk of the library's compile time.
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

// geometry of the register-tiled kernel: (dtype, nvec) -> (T, RB, THREADS)
void rtile_geom(int dtype, int nvec, int& T, int& RB, int& threads) {
  if (nvec <= 1) {
    threads = 256; RB = dtype == B200Q_C128 ? 4 : 5;
  } else { threads = 512; RB = dtype == B200Q_C128 ? 3 : 4; }
  T = (threads == 256 ? 8 : 9) + RB;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T, int RB, int NV, int THREADS, int MINB>
static int rtile_launch(void* v0, void* v1, const RtArgs& a, int64_t batch, const RtOp* ops_dev,
                        const double2* mats_dev, int nslots, double scale, double* out_dev,
                        double* partials, size_t partial_cap, cudaStream_t s) {
  const size_t smem = ((size_t)NV * sizeof(cx<T>) << a.T) + sizeof(cx<T>) * ((a.nmat + 1) & ~1) +
                      (size_t)a.nops * sizeof(RtOp) +
                      (size_t)((2 << RB) + 2 * THREADS + 2) * sizeof(unsigned long long) +
                      (size_t)nslots * (THREADS / 32) * sizeof(double) +
                      ((size_t)a.nrounds * (THREADS + 8) + THREADS + (1 << RB)) * sizeof(unsigned short);
  B200Q_REQUIRE(smem <= 227 * 1024, "rtile: %zu bytes of shared memory needed (%d ops, %d slots)",
                smem, a.nops, nslots);
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CHECK(cudaFuncSetAttribute(k_rtile<T, RB, NV, THREADS, MINB>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  uint64_t per_sm = std::max<uint64_t>(1, std::min<uint64_t>(MINB, (227 * 1024) / smem));
  static const int ctas_knob = env_int("B200Q_RT_CTAS", 0);                    // tuning knob
  if (ctas_knob > 0) per_sm = ctas_knob;
  const uint64_t cap = (uint64_t)sm_count() * per_sm;
  dim3 grid((unsigned)std::min<uint64_t>(a.ntiles, cap), (unsigned)batch);
  if (nslots > 0)
    B200Q_REQUIRE((size_t)batch * nslots * grid.x <= partial_cap, "rtile: workspace too small for %d slots",
                  nslots);
  k_rtile<T, RB, NV, THREADS, MINB><<<grid, THREADS, smem, s>>>(
      a, (cx<T>*)v0, (cx<T>*)v1, ops_dev, mats_dev, 0, partials);
  B200Q_LAUNCH_CHECK();
  if (nslots > 0) {
    k_rt_final_reduce<<<(unsigned)(batch * nslots), 256, 0, s>>>(partials, out_dev, (int)grid.x, scale);
    B200Q_LAUNCH_CHECK();
  }
  return 0;
}

int rtile_dispatch(void* v0, void* v1, int n, int dtype, int64_t batch, const int* tile_bits,
                   int Tn, int L, const RtOp* ops_host, int nops, const double2* mats_host,
                   int nmat, int nslots, int write0, uint64_t base_hi, double scale,
                   double* out_dev, void* work, size_t work_bytes, cudaStream_t s) {
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
  a.prefetch = (pf_knob && ((elem << L) >= 16) && (((uintptr_t)v0 | (uintptr_t)v1) % 16 == 0)) ? 1 : 0;
  const size_t ops_bytes = (size_t)nops * sizeof(RtOp);
  const size_t mat_bytes = (size_t)nmat * sizeof(double2);
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
  if (dtype == B200Q_C128) {
    if (!v1) return rtile_launch<double, 4, 1, 256, 2>(v0, v1, a, batch, od, md, nslots, scale, out_dev, partials, pcap, s);
    return rtile_launch<double, 3, 2, 512, 1>(v0, v1, a, batch, od, md, nslots, scale, out_dev, partials, pcap, s);
  }
  if (dtype == B200Q_C64) {
    if (!v1) return rtile_launch<float, 5, 1, 256, 2>(v0, v1, a, batch, od, md, nslots, scale, out_dev, partials, pcap, s);
    return rtile_launch<float, 4, 2, 512, 1>(v0, v1, a, batch, od, md, nslots, scale, out_dev, partials, pcap, s);
  }
  set_error("unknown dtype %d", dtype);
  return 2;
}

}  // namespace b200q
