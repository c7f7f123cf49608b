// b200q — extern "C" entry points (see include/b200q.h for the contract of each function).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../../include/b200q.h"
#include "adjoint.cuh"
#include "common.cuh"
#include "gates.cuh"
#include "measure.cuh"
#include "sample.cuh"
#include "tile.cuh"
#include "rtile_host.h"

namespace b200q {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      cached = v;
    else
      cached = 148;
  }
  return cached;
}

static const int kReduceCtasPerSm = 4;

// every reduction kernel runs with exactly this many CTAs so partial rows have one stride
static inline unsigned reduce_ncta() { return (unsigned)(sm_count() * kReduceCtasPerSm); }

// Build the group enumeration (zero-insert positions + control OR mask).
static int build_group(GroupArgs& a, int n, const int* tgt, int k, const int* ctrl,
                       const int* cvals, int nc) {
  B200Q_REQUIRE(n >= 1 && n <= B200Q_MAX_BITS, "n=%d out of range", n);
  B200Q_REQUIRE(k >= 0 && nc >= 0 && nc <= B200Q_MAX_CTRL && k + nc <= n &&
                    k + nc <= (int)sizeof(a.ins),
                "bad gate arity k=%d nc=%d n=%d", k, nc, n);
  uint64_t seen = 0;
  std::vector<int> pos;
  for (int j = 0; j < k; ++j) pos.push_back(tgt[j]);
  for (int j = 0; j < nc; ++j) pos.push_back(ctrl[j]);
  for (int p : pos) {
    B200Q_REQUIRE(p >= 0 && p < n, "bit %d out of range for n=%d", p, n);
    B200Q_REQUIRE(!((seen >> p) & 1), "bit %d used twice", p);
    seen |= 1ull << p;
  }
  std::sort(pos.begin(), pos.end());
  memset(&a, 0, sizeof(a));
  a.n = n;
  a.nins = (int)pos.size();
  for (size_t i = 0; i < pos.size(); ++i) a.ins[i] = (int8_t)pos[i];
  a.ctrl_or = 0;
  for (int j = 0; j < nc; ++j)
    if (!cvals || cvals[j]) a.ctrl_or |= 1ull << ctrl[j];
  a.ngroups = 1ull << (n - a.nins);
  return 0;
}

template <int K> static void build_off(DenseOff<K>& o, const int* tgt) {
  for (int r = 0; r < (1 << K); ++r) {
    uint64_t off = 0;
    for (int j = 0; j < K; ++j)
      if ((r >> (K - 1 - j)) & 1) off |= 1ull << tgt[j];
    o.off[r] = off;
  }
}

template <int K> static void load_mat(MatVal<K>& m, const void* host) {
  if (host) memcpy(m.m, host, sizeof(m.m));
  else memset(m.m, 0, sizeof(m.m));
}

template <typename T, int K>
static int launch_dense(void* state, const GroupArgs& a, const int* tgt, int64_t batch,
                        const void* mat_host, const void* mat_dev, int64_t bstride,
                        cudaStream_t s) {
  DenseOff<K> o;
  build_off<K>(o, tgt);
  MatVal<K> mv;
  load_mat<K>(mv, mat_host);
  dim3 grid(grid_for(a.ngroups, 256, 8), (unsigned)batch);
  if (mat_host)
    k_dense<T, K, true><<<grid, 256, 0, s>>>((cx<T>*)state, a, o, mv, nullptr, 0);
  else
    k_dense<T, K, false><<<grid, 256, 0, s>>>((cx<T>*)state, a, o, mv, (const cx<T>*)mat_dev,
                                               (long long)bstride);
  B200Q_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_dense_big(void* state, int n, int64_t batch, const int* tgt, int k,
                            const int* ctrl, const int* cvals, int nc, const void* mat_dev,
                            int64_t bstride, cudaStream_t s) {
  GroupArgs g;
  if (int rc = build_group(g, n, tgt, k, ctrl, cvals, nc)) return rc;
  BigArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.k = k; a.nins = g.nins;
  memcpy(a.ins, g.ins, sizeof(a.ins));
  for (int j = 0; j < k; ++j) a.tbits[j] = (int8_t)tgt[j];
  a.ctrl_or = g.ctrl_or;
  a.ngroups = g.ngroups;
  constexpr int TILE = 2048;
  const int D = 1 << k;
  const int G = TILE / D;
  const uint64_t ntiles = (a.ngroups + G - 1) / G;
  const size_t smem = TILE * sizeof(cx<T>) + D * sizeof(uint64_t);
  static bool attr_set[2] = {false, false};
  const int ti = sizeof(T) == 8 ? 1 : 0;
  if (!attr_set[ti]) {
    B200Q_CHECK(cudaFuncSetAttribute(k_dense_big<T, TILE>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set[ti] = true;
  }
  uint64_t cap = (uint64_t)sm_count() * 4;
  dim3 grid((unsigned)std::min<uint64_t>(ntiles, cap), (unsigned)batch);
  k_dense_big<T, TILE><<<grid, 256, smem, s>>>((cx<T>*)state, a, (const cx<T>*)mat_dev,
                                                 (long long)bstride);
  B200Q_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int apply_matrix_t(void* state, int n, int64_t batch, const int* tgt, int k,
                          const int* ctrl, const int* cvals, int nc, const void* mat_host,
                          const void* mat_dev, int64_t bstride, cudaStream_t s) {
  B200Q_REQUIRE(k >= 1 && k <= B200Q_MAX_BIG_K, "apply_matrix: k=%d unsupported (1..%d)", k,
                B200Q_MAX_BIG_K);
  B200Q_REQUIRE(mat_host || mat_dev, "apply_matrix: no matrix given");
  B200Q_REQUIRE(!(bstride != 0 && !mat_dev), "apply_matrix: batched matrices must be on device");
  if (k > B200Q_MAX_DENSE_K) {
    B200Q_REQUIRE(mat_dev, "apply_matrix: k=%d needs mat_dev", k);
    return launch_dense_big<T>(state, n, batch, tgt, k, ctrl, cvals, nc, mat_dev, bstride, s);
  }
  GroupArgs a;
  if (int rc = build_group(a, n, tgt, k, ctrl, cvals, nc)) return rc;
  const void* mh = mat_dev ? nullptr : mat_host;
  switch (k) {
    case 1: return launch_dense<T, 1>(state, a, tgt, batch, mh, mat_dev, bstride, s);
    case 2: return launch_dense<T, 2>(state, a, tgt, batch, mh, mat_dev, bstride, s);
    default: return launch_dense<T, 3>(state, a, tgt, batch, mh, mat_dev, bstride, s);
  }
}

template <typename T>
static int apply_diag_t(void* state, int n, int64_t batch, const int* bits, int k,
                        const void* diag_host, const void* diag_dev, int64_t bstride,
                        cudaStream_t s) {
  B200Q_REQUIRE(k >= 1 && k <= 20 && k <= n, "apply_diag: k=%d unsupported", k);
  B200Q_REQUIRE(diag_host || diag_dev, "apply_diag: no table given");
  DiagArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.k = k;
  uint64_t seen = 0;
  for (int j = 0; j < k; ++j) {
    B200Q_REQUIRE(bits[j] >= 0 && bits[j] < n && !((seen >> bits[j]) & 1), "apply_diag: bad bit %d",
                  bits[j]);
    seen |= 1ull << bits[j];
    a.bits[j] = (int8_t)bits[j];
  }
  MatVal<3> dv;
  memset(&dv, 0, sizeof(dv));
  dim3 grid(grid_for(1ull << n, 256, 8), (unsigned)batch);
  const size_t tab_bytes = sizeof(cx<T>) << k;
  if (!diag_dev) {
    B200Q_REQUIRE(k <= 6, "apply_diag: k=%d host tables are limited to k <= 6", k);
    memcpy(dv.m, diag_host, sizeof(double2) << k);
    k_diag<T, 0><<<grid, 256, tab_bytes, s>>>((cx<T>*)state, a, dv, nullptr, 0);
  } else if (k <= 11) {
    k_diag<T, 1><<<grid, 256, tab_bytes, s>>>((cx<T>*)state, a, dv, (const cx<T>*)diag_dev,
                                               (long long)bstride);
  } else {
    k_diag<T, 2><<<grid, 256, 0, s>>>((cx<T>*)state, a, dv, (const cx<T>*)diag_dev,
                                       (long long)bstride);
  }
  B200Q_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int probs_t(const void* state, int n, int64_t batch, const int* bits, int m, double* out,
                   void* work, size_t work_bytes, cudaStream_t s) {
  B200Q_REQUIRE(m >= 0 && m <= n, "probs: m=%d out of range", m);
  bool full = (m == n);
  for (int j = 0; full && j < m; ++j) full = (bits[j] == n - 1 - j);
  if (full) {
    const uint64_t total = (uint64_t)batch << n;
    k_probs_full<T><<<grid_for(total, 256, 8), 256, 0, s>>>((const cx<T>*)state, out, total);
    B200Q_LAUNCH_CHECK();
    return 0;
  }
  MargArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.m = m;
  uint64_t tmask = 0;
  for (int j = 0; j < m; ++j) {
    B200Q_REQUIRE(bits[j] >= 0 && bits[j] < n && !((tmask >> bits[j]) & 1), "probs: bad bit %d",
                  bits[j]);
    tmask |= 1ull << bits[j];
    a.tbits[j] = (int8_t)bits[j];
  }
  a.lane_valid = n >= 5 ? 32u : (1u << n);
  for (int q = 0; q < n; ++q) {
    const bool is_t = (tmask >> q) & 1;
    if (q < 5) {
      if (!is_t) a.lane_sum_mask |= 1u << q;
    } else if (is_t) {
      a.outer_pos[a.n_outer++] = (int8_t)q;
    } else {
      a.sum_pos[a.n_sum_hi++] = (int8_t)q;
      a.sum_mask |= 1ull << q;
    }
  }
  // ~2^17 warp tasks: an order of magnitude more than resident warps (148 SMs x 64), so the
  // grid-stride tail costs a few percent instead of up to one task in two
  int lg = std::max(0, std::min(17 - a.n_outer, a.n_sum_hi - 4));
  const size_t avail = (work && work_bytes > kTermRegion) ? work_bytes - kTermRegion : 0;
  while (lg > 0 && ((uint64_t)batch << (m + lg)) * sizeof(double) > avail) --lg;
  a.lg_nsplit = lg;
  double* partials = (lg == 0) ? out : (double*)((char*)work + kTermRegion);
  const uint64_t ntasks = 1ull << (a.n_outer + lg);
  dim3 grid(grid_for(ntasks * 32, 256, 8), (unsigned)batch);
  k_probs_marginal<T><<<grid, 256, 0, s>>>((const cx<T>*)state, partials, a);
  B200Q_LAUNCH_CHECK();
  if (lg > 0) {
    if (m <= 10 && lg >= 8) {
      dim3 g2(1u << m, (unsigned)batch);
      k_sum_splits_cta<<<g2, 256, 0, s>>>(partials, out, m, lg);
    } else {
      dim3 g2((unsigned)(((1ull << m) + 255) / 256), (unsigned)batch);
      k_sum_splits<<<g2, 256, 0, s>>>(partials, out, m, lg);
    }
    B200Q_LAUNCH_CHECK();
  }
  return 0;
}

template <typename T>
static int expval_t(const void* state, int n, int64_t batch, const uint64_t* xm,
                    const uint64_t* zm, const int* nys, const double* coeffs, int nterms,
                    double* out, void* work, size_t work_bytes, cudaStream_t s) {
  B200Q_REQUIRE(work && work_bytes >= kWorkBytes, "expval: workspace too small");
  B200Q_REQUIRE(nterms >= 0, "expval: nterms < 0");
  // group by xmask, preserving first-appearance order (pauli_arithmetic.py:933-937)
  std::vector<uint64_t> masks;
  std::vector<std::vector<int>> groups;
  for (int t = 0; t < nterms; ++t) {
    size_t gi = 0;
    for (; gi < masks.size(); ++gi) if (masks[gi] == xm[t]) break;
    if (gi == masks.size()) { masks.push_back(xm[t]); groups.emplace_back(); }
    groups[gi].push_back(t);
  }
  const unsigned ncta = reduce_ncta();
  char* term_base = (char*)work;
  double* partials = (double*)((char*)work + kTermRegion);
  const size_t max_rows = (work_bytes - kTermRegion) / (sizeof(double) * ncta);
  size_t term_off = 0;
  int launches = 0;
  const size_t chunk = 2048;                                 // terms per launch (<= 48 KiB smem)
  for (size_t gi = 0; gi < masks.size(); ++gi) {
    const std::vector<int>& idx = groups[gi];
    for (size_t c0 = 0; c0 < idx.size(); c0 += chunk) {
      const size_t cn = std::min(chunk, idx.size() - c0);
      B200Q_REQUIRE((size_t)(launches + 1) * batch <= max_rows, "expval: too many term groups");
      double* prow = partials + (size_t)launches * batch * ncta;
      dim3 grid(ncta, (unsigned)batch);
      if (masks[gi] == 0) {
        std::vector<PauliTerm> tt(cn);
        for (size_t i = 0; i < cn; ++i) {
          const int t = idx[c0 + i];
          B200Q_REQUIRE(nys[t] == 0, "expval: diagonal term with Y count %d", nys[t]);
          tt[i].zmask = zm[t]; tt[i].coeff = coeffs[t];
        }
        const size_t bytes = cn * sizeof(PauliTerm);
        B200Q_REQUIRE(term_off + bytes <= kTermRegion, "expval: term table overflow");
        B200Q_CHECK(cudaMemcpyAsync(term_base + term_off, tt.data(), bytes,
                                    cudaMemcpyHostToDevice, s));
        k_expval_diag<T><<<grid, 256, bytes, s>>>(
            (const cx<T>*)state, n, (const PauliTerm*)(term_base + term_off), (int)cn, prow);
        B200Q_LAUNCH_CHECK();
        term_off += (bytes + 255) & ~(size_t)255;
      } else {
        std::vector<PauliTermXY> tt(cn);
        for (size_t i = 0; i < cn; ++i) {
          const int t = idx[c0 + i];
          tt[i].zmask = zm[t]; tt[i].coeff = coeffs[t]; tt[i].ny = nys[t]; tt[i].pad = 0;
        }
        const size_t bytes = cn * sizeof(PauliTermXY);
        B200Q_REQUIRE(term_off + bytes <= kTermRegion, "expval: term table overflow");
        B200Q_CHECK(cudaMemcpyAsync(term_base + term_off, tt.data(), bytes,
                                    cudaMemcpyHostToDevice, s));
        const int pivot = 63 - __builtin_clzll(masks[gi]);
        B200Q_REQUIRE(pivot < n, "expval: xmask outside the state");
        k_expval_offdiag<T><<<grid, 256, bytes, s>>>(
            (const cx<T>*)state, n, masks[gi], pivot,
            (const PauliTermXY*)(term_base + term_off), (int)cn, prow);
        B200Q_LAUNCH_CHECK();
        term_off += (bytes + 255) & ~(size_t)255;
      }
      ++launches;
    }
  }
  if (launches == 0) {
    B200Q_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * batch, s));
    return 0;
  }
  k_final_reduce<<<(unsigned)batch, 256, 0, s>>>(partials, out, (int)ncta, launches, (int)batch,
                                                  1.0);
  B200Q_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int inner_t(const void* a, const void* b, int n, int64_t batch, double* out, void* work,
                   size_t work_bytes, cudaStream_t s) {
  const unsigned ncta = reduce_ncta();
  B200Q_REQUIRE(work && work_bytes >= kTermRegion + 2 * (size_t)batch * ncta * sizeof(double),
                "inner: workspace too small");
  double* partials = (double*)((char*)work + kTermRegion);
  dim3 grid(ncta, (unsigned)batch);
  k_inner<T><<<grid, 256, 0, s>>>((const cx<T>*)a, (const cx<T>*)b, n, partials);
  B200Q_LAUNCH_CHECK();
  // planes [2][batch][ncta] -> out[2*batch]
  k_final_reduce<<<(unsigned)(2 * batch), 256, 0, s>>>(partials, out, (int)ncta, 1,
                                                        (int)(2 * batch), 1.0);
  B200Q_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int pauli_sum_apply_t(const void* in, void* out, int n, int64_t batch, const uint64_t* xm,
                             const uint64_t* zm, const int* nys, const double* cre,
                             const double* cim, int nterms, double scale, void* work,
                             size_t work_bytes, cudaStream_t s) {
  B200Q_REQUIRE(in != out, "pauli_sum_apply: must be out of place");
  B200Q_REQUIRE(nterms >= 1 && nterms <= 1024, "pauli_sum_apply: nterms=%d (1..1024)", nterms);
  B200Q_REQUIRE(work && work_bytes >= kTermRegion, "pauli_sum_apply: workspace too small");
  std::vector<PauliTermFull> tt(nterms);
  for (int t = 0; t < nterms; ++t) {
    tt[t].xmask = xm[t]; tt[t].zmask = zm[t]; tt[t].cre = cre[t]; tt[t].cim = cim ? cim[t] : 0.0;
    tt[t].ny = nys[t]; tt[t].pad = 0;
    B200Q_REQUIRE((xm[t] >> n) == 0 && (zm[t] >> n) == 0, "pauli_sum_apply: mask outside state");
  }
  const size_t bytes = nterms * sizeof(PauliTermFull);
  B200Q_CHECK(cudaMemcpyAsync(work, tt.data(), bytes, cudaMemcpyHostToDevice, s));
  dim3 grid(grid_for(1ull << n, 256, 8), (unsigned)batch);
  k_pauli_sum_apply<T><<<grid, 256, bytes, s>>>((const cx<T>*)in, (cx<T>*)out, n,
                                                 (const PauliTermFull*)work, nterms, scale);
  B200Q_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int pauli_braket_t(const void* bra, const void* ket, int n, uint64_t xmask, uint64_t zmask,
                          int ny, double* out, void* work, size_t work_bytes, cudaStream_t s) {
  const unsigned ncta = reduce_ncta();
  B200Q_REQUIRE(work && work_bytes >= kTermRegion + 2 * (size_t)ncta * sizeof(double),
                "pauli_braket: workspace too small");
  double* partials = (double*)((char*)work + kTermRegion);
  k_pauli_braket<T><<<ncta, 256, 0, s>>>((const cx<T>*)bra, (const cx<T>*)ket, n, xmask, zmask, ny,
                                          partials);
  B200Q_LAUNCH_CHECK();
  k_final_reduce<<<2, 256, 0, s>>>(partials, out, (int)ncta, 1, 2, 1.0);
  B200Q_LAUNCH_CHECK();
  return 0;
}

template <typename T, int K>
static int adjoint_step_k(void* vecs, const GroupArgs& a, const int* tgt, int n_bras,
                          const void* adj_host, const void* gen_host, double* out, void* work,
                          cudaStream_t s) {
  DenseOff<K> o;
  build_off<K>(o, tgt);
  MatVal<K> adj, gen;
  load_mat<K>(adj, adj_host);
  load_mat<K>(gen, gen_host);
  const unsigned ncta = reduce_ncta();
  double* partials = (double*)((char*)work + kTermRegion);
  if (n_bras > 1) {
    dim3 grid(ncta, (unsigned)(n_bras - 1));
    k_adjoint_step<T, K><<<grid, 256, 0, s>>>((cx<T>*)vecs, a, o, adj, gen, partials, 1, n_bras, 0);
    B200Q_LAUNCH_CHECK();
  }
  dim3 grid(ncta, 1);
  k_adjoint_step<T, K><<<grid, 256, 0, s>>>((cx<T>*)vecs, a, o, adj, gen, partials, 0, n_bras, 1);
  B200Q_LAUNCH_CHECK();
  // out[b] = -Im z_b : reduce the imaginary plane with scale -1
  k_final_reduce<<<(unsigned)n_bras, 256, 0, s>>>(partials + (size_t)n_bras * ncta, out, (int)ncta,
                                                   1, n_bras, -1.0);
  B200Q_LAUNCH_CHECK();
  return 0;
}

static int np_sum(const double* p, uint64_t count, double* out, double* scratch,
                  cudaStream_t s) {
  // numpy pairwise sum of a power-of-two-length vector -> out[0]
  if (count <= 128) {
    k_np_sum_small<<<1, 32, 0, s>>>(p, out, (int)count);
    B200Q_LAUNCH_CHECK();
    return 0;
  }
  uint64_t cur = count / 128;
  k_np_leaf_sums<<<(unsigned)((cur + 127) / 128), 128, 0, s>>>(p, scratch, cur);
  B200Q_LAUNCH_CHECK();
  double* src = scratch;
  double* dst = scratch + cur;
  while (true) {
    const uint64_t nb = (cur + 2047) / 2048;
    double* target = (nb == 1) ? out : dst;
    k_np_tree<<<(unsigned)nb, 1024, 0, s>>>(src, target, cur);
    B200Q_LAUNCH_CHECK();
    if (nb == 1) break;
    src = dst; dst = dst + nb; cur = nb;
  }
  return 0;
}

template <typename T>
static int tile_t(void* state, int n, int64_t batch, const int* tile_bits, int Tn, int L,
                  const TileOp* ops_host, int nops, const double2* mats_host, int nmat,
                  void* work, size_t work_bytes, cudaStream_t s) {
  B200Q_REQUIRE(Tn >= 1 && Tn <= n && Tn <= 14 && L >= 0 && L <= Tn, "tile: bad T=%d L=%d n=%d",
                Tn, L, n);
  B200Q_REQUIRE(nops >= 1 && nops <= 4096 && nmat >= 0, "tile: bad nops=%d nmat=%d", nops, nmat);
  TileArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.T = Tn; a.L = L; a.nops = nops; a.nmat = nmat;
  uint64_t inmask = 0;
  for (int i = 0; i < Tn; ++i) {
    const int b = tile_bits[i];
    B200Q_REQUIRE(b >= 0 && b < n && !((inmask >> b) & 1), "tile: bad tile bit %d", b);
    B200Q_REQUIRE(i < L ? b == i : (i == 0 || b > tile_bits[i - 1]),
                  "tile: bits must be ascending with the first L equal to 0..L-1");
    inmask |= 1ull << b;
    if (i >= L) a.hi_bits[i - L] = (int8_t)b;
  }
  int no = 0;
  for (int b = 0; b < n; ++b)
    if (!((inmask >> b) & 1)) a.out_bits[no++] = (int8_t)b;
  a.ntiles = 1ull << (n - Tn);
  const size_t ops_bytes = (size_t)nops * sizeof(TileOp);
  const size_t mat_bytes = (size_t)nmat * sizeof(double2);
  B200Q_REQUIRE(work && ops_bytes + mat_bytes + 512 <= kTermRegion && work_bytes >= kTermRegion,
                "tile: segment tables too large for the workspace");
  char* w = (char*)work;
  B200Q_CHECK(cudaMemcpyAsync(w, ops_host, ops_bytes, cudaMemcpyHostToDevice, s));
  const size_t moff = (ops_bytes + 255) & ~(size_t)255;
  if (nmat) B200Q_CHECK(cudaMemcpyAsync(w + moff, mats_host, mat_bytes, cudaMemcpyHostToDevice, s));
  const size_t smem = (sizeof(cx<T>) << Tn) + sizeof(cx<T>) * ((nmat + 1) & ~1) + ops_bytes;
  B200Q_REQUIRE(smem <= 227 * 1024, "tile: %zu bytes of shared memory needed (T=%d, %d ops)", smem,
                Tn, nops);
  constexpr int THREADS = 256;
  static size_t attr_set[2] = {0, 0};
  const int ti = sizeof(T) == 8 ? 1 : 0;
  if (attr_set[ti] < smem) {
    B200Q_CHECK(cudaFuncSetAttribute(k_tile<T, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    attr_set[ti] = 227 * 1024;
  }
  const uint64_t per_sm = std::max<uint64_t>(1, (227 * 1024) / smem);
  const uint64_t cap = (uint64_t)sm_count() * std::min<uint64_t>(per_sm, 8);
  dim3 grid((unsigned)std::min<uint64_t>(a.ntiles, cap), (unsigned)batch);
  k_tile<T, THREADS><<<grid, THREADS, smem, s>>>((cx<T>*)state, a, (const TileOp*)w,
                                                  (const double2*)(w + moff));
  B200Q_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200q

using namespace b200q;

#define DISPATCH(dtype, CALL_F, CALL_D)                          \
  do {                                                           \
    if ((dtype) == B200Q_DTYPE_C64) return CALL_F;               \
    if ((dtype) == B200Q_DTYPE_C128) return CALL_D;              \
    set_error("unknown dtype %d", (dtype));                      \
    return 2;                                                    \
  } while (0)

namespace b200q {

// In-place numpy-ordered cumsum (bit-exact), parallel: see sample.cuh "exact mode, parallel".
// Small vectors keep the serial chain (it is a few microseconds there).
static int cumsum_exact(double* p_dev, uint64_t count, const double* carry_in_dev, void* work,
                        size_t work_bytes, cudaStream_t s) {
  if (count < (uint64_t)4 * EXC_CHUNK) {
    k_cumsum_serial<1024><<<1, 64, 0, s>>>(p_dev, p_dev, count, carry_in_dev);
    B200Q_LAUNCH_CHECK();
    return 0;
  }
  const uint64_t nchunks = (count + EXC_CHUNK - 1) / EXC_CHUNK;
  const size_t need = (nchunks + 64) * 48;
  B200Q_REQUIRE(work && work_bytes >= kWorkBytes && need <= work_bytes - kTermRegion,
                "cumsum: workspace too small for 2^%d values (exact mode needs %zu bytes)",
                (int)(63 - __builtin_clzll(count)), need + kTermRegion);
  char* base = (char*)work + kTermRegion;
  double* approx = (double*)base + 8;                                   // nchunks (+ padding)
  unsigned long long* a0 = (unsigned long long*)(approx + nchunks + 8);
  unsigned long long* a1 = a0 + nchunks;
  double* start = (double*)(a1 + nchunks);
  int* flags = (int*)(start + nchunks);
  k_scan_block_totals<<<(unsigned)nchunks, 256, 0, s>>>(p_dev, approx, count);
  B200Q_LAUNCH_CHECK();
  k_scan_totals<<<1, 1024, 0, s>>>(approx, nchunks, carry_in_dev);
  B200Q_LAUNCH_CHECK();
  const unsigned grid = (unsigned)((nchunks + 3) / 4);
  k_excum_summaries<<<grid, 128, 0, s>>>(p_dev, count, approx, nchunks, a0, a1, flags);
  B200Q_LAUNCH_CHECK();
  k_excum_ordered<<<1, 32, 0, s>>>(p_dev, p_dev, count, approx, nchunks, a0, a1, flags, start,
                                   carry_in_dev);
  B200Q_LAUNCH_CHECK();
  k_excum_replay<<<grid, 128, 0, s>>>(p_dev, p_dev, count, nchunks, flags, start);
  B200Q_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200q

extern "C" {

const char* b200q_last_error(void) { return g_err; }
int b200q_version(void) { return B200Q_ABI_VERSION; }
int b200q_sm_count(void) { return sm_count(); }
size_t b200q_workspace_bytes(void) { return kWorkBytes; }

int b200q_set_basis_state(void* state, int n, int dtype, int64_t batch, uint64_t index,
                          void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(n >= 0 && n <= B200Q_MAX_BITS && batch >= 1, "set_basis_state: bad n/batch");
  B200Q_REQUIRE(index < (1ull << n), "set_basis_state: index out of range");
  const size_t esz = dtype == B200Q_DTYPE_C128 ? 16 : 8;
  B200Q_CHECK(cudaMemsetAsync(state, 0, esz * ((size_t)batch << n), s));
  if (dtype == B200Q_DTYPE_C128) k_set_one<double><<<(unsigned)batch, 1, 0, s>>>((double2*)state, n, index);
  else k_set_one<float><<<(unsigned)batch, 1, 0, s>>>((float2*)state, n, index);
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_apply_matrix(void* state, int n, int dtype, int64_t batch, const int* tgt_bits, int k,
                       const int* ctrl_bits, const int* ctrl_vals, int nc, const void* mat_host,
                       const void* mat_dev, int64_t mat_bstride, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype,
           apply_matrix_t<float>(state, n, batch, tgt_bits, k, ctrl_bits, ctrl_vals, nc, mat_host,
                                 mat_dev, mat_bstride, s),
           apply_matrix_t<double>(state, n, batch, tgt_bits, k, ctrl_bits, ctrl_vals, nc, mat_host,
                                  mat_dev, mat_bstride, s));
}

int b200q_apply_diag(void* state, int n, int dtype, int64_t batch, const int* bits, int k,
                     const void* diag_host, const void* diag_dev, int64_t diag_bstride,
                     void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype,
           apply_diag_t<float>(state, n, batch, bits, k, diag_host, diag_dev, diag_bstride, s),
           apply_diag_t<double>(state, n, batch, bits, k, diag_host, diag_dev, diag_bstride, s));
}

int b200q_apply_phase(void* state, int n, int dtype, int64_t batch, const int* ctrl_bits,
                      const int* ctrl_vals, int nc, double phase_re, double phase_im,
                      const void* phase_dev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GroupArgs a;
  if (int rc = build_group(a, n, nullptr, 0, ctrl_bits, ctrl_vals, nc)) return rc;
  dim3 grid(grid_for(a.ngroups, 256, 8), (unsigned)batch);
  const double2 ph = make_double2(phase_re, phase_im);
  if (dtype == B200Q_DTYPE_C128)
    k_phase<double><<<grid, 256, 0, s>>>((double2*)state, a, ph, (const double2*)phase_dev);
  else if (dtype == B200Q_DTYPE_C64)
    k_phase<float><<<grid, 256, 0, s>>>((float2*)state, a, ph, (const float2*)phase_dev);
  else { set_error("unknown dtype %d", dtype); return 2; }
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_collapse(void* state, int n, int dtype, int bit, int sample, int reset, double scale,
                   void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(n >= 1 && bit >= 0 && bit < n, "collapse: bit %d outside a %d-qubit state", bit, n);
  B200Q_REQUIRE(sample == 0 || sample == 1, "collapse: sample must be 0 or 1, got %d", sample);
  const unsigned grid = grid_for(1ull << (n - 1), 256, 8);
  if (dtype == B200Q_DTYPE_C128)
    k_collapse<double><<<grid, 256, 0, s>>>((double2*)state, n, bit, sample, reset != 0, scale);
  else if (dtype == B200Q_DTYPE_C64)
    k_collapse<float><<<grid, 256, 0, s>>>((float2*)state, n, bit, sample, reset != 0, scale);
  else { set_error("unknown dtype %d", dtype); return 2; }
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_apply_parity_phase(void* state, int n, int dtype, int64_t batch, uint64_t mask,
                             double p0_re, double p0_im, double p1_re, double p1_im,
                             const void* phases_dev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(n >= 1 && n <= B200Q_MAX_BITS && (mask >> n) == 0, "parity_phase: bad mask");
  dim3 grid(grid_for(1ull << n, 256, 8), (unsigned)batch);
  const double2 p0 = make_double2(p0_re, p0_im), p1 = make_double2(p1_re, p1_im);
  if (dtype == B200Q_DTYPE_C128)
    k_parity_phase<double><<<grid, 256, 0, s>>>((double2*)state, n, mask, p0, p1,
                                                 (const double2*)phases_dev);
  else if (dtype == B200Q_DTYPE_C64)
    k_parity_phase<float><<<grid, 256, 0, s>>>((float2*)state, n, mask, p0, p1,
                                                (const float2*)phases_dev);
  else { set_error("unknown dtype %d", dtype); return 2; }
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_apply_pauli_rot(void* state, int n, int dtype, int64_t batch, uint64_t xmask,
                          uint64_t zmask, int ny, double c, double sn, const void* cs_dev,
                          void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(n >= 1 && n <= B200Q_MAX_BITS && xmask != 0 && (xmask >> n) == 0 &&
                    (zmask >> n) == 0,
                "pauli_rot: bad masks");
  const int pivot = 63 - __builtin_clzll(xmask);
  dim3 grid(grid_for(1ull << (n - 1), 256, 8), (unsigned)batch);
  const double2 cs = make_double2(c, sn);
  if (dtype == B200Q_DTYPE_C128)
    k_pauli_rot<double><<<grid, 256, 0, s>>>((double2*)state, n, pivot, xmask, zmask, ny, cs,
                                              (const double2*)cs_dev);
  else if (dtype == B200Q_DTYPE_C64)
    k_pauli_rot<float><<<grid, 256, 0, s>>>((float2*)state, n, pivot, xmask, zmask, ny, cs,
                                             (const float2*)cs_dev);
  else { set_error("unknown dtype %d", dtype); return 2; }
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_probs(const void* state, int n, int dtype, int64_t batch, const int* bits, int m,
                double* out_dev, void* work, size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype, probs_t<float>(state, n, batch, bits, m, out_dev, work, work_bytes, s),
           probs_t<double>(state, n, batch, bits, m, out_dev, work, work_bytes, s));
}

int b200q_expval_pauli_sum(const void* state, int n, int dtype, int64_t batch,
                           const uint64_t* xmasks, const uint64_t* zmasks, const int* nys,
                           const double* coeffs, int nterms, double* out_dev, void* work,
                           size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype,
           expval_t<float>(state, n, batch, xmasks, zmasks, nys, coeffs, nterms, out_dev, work,
                           work_bytes, s),
           expval_t<double>(state, n, batch, xmasks, zmasks, nys, coeffs, nterms, out_dev, work,
                            work_bytes, s));
}

int b200q_inner(const void* a, const void* b, int n, int dtype, int64_t batch, double* out_dev,
                void* work, size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype, inner_t<float>(a, b, n, batch, out_dev, work, work_bytes, s),
           inner_t<double>(a, b, n, batch, out_dev, work, work_bytes, s));
}

int b200q_pauli_sum_apply(const void* in, void* out, int n, int dtype, int64_t batch,
                          const uint64_t* xmasks, const uint64_t* zmasks, const int* nys,
                          const double* cre, const double* cim, int nterms, double scale,
                          void* work, size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype,
           pauli_sum_apply_t<float>(in, out, n, batch, xmasks, zmasks, nys, cre, cim, nterms,
                                    scale, work, work_bytes, s),
           pauli_sum_apply_t<double>(in, out, n, batch, xmasks, zmasks, nys, cre, cim, nterms,
                                     scale, work, work_bytes, s));
}

int b200q_pauli_braket(const void* bra, const void* ket, int n, int dtype, uint64_t xmask,
                       uint64_t zmask, int ny, double* out_dev, void* work, size_t work_bytes,
                       void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype,
           pauli_braket_t<float>(bra, ket, n, xmask, zmask, ny, out_dev, work, work_bytes, s),
           pauli_braket_t<double>(bra, ket, n, xmask, zmask, ny, out_dev, work, work_bytes, s));
}

int b200q_sample(double* probs_dev, int m, const double* uniforms_dev, int64_t shots, int mode,
                 int64_t* idx_out_dev, int64_t* bits_out_dev, double* norm_out_dev,
                 int* flags_dev, void* work, size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(m >= 0 && m <= 40, "sample: m=%d out of range", m);
  B200Q_REQUIRE(work && work_bytes >= kWorkBytes, "sample: workspace too small");
  const uint64_t count = 1ull << m;
  double* scratch = (double*)((char*)work + kTermRegion);
  // scratch needs count/128 * (1 + 1/2048 + ...) doubles for the pairwise tree, or count/2048
  // block totals for the fast scan
  B200Q_REQUIRE((count / 128 + count / (128 * 2047) + 64) * sizeof(double) <=
                    work_bytes - kTermRegion,
                "sample: workspace too small for 2^%d probabilities", m);
  double* one = scratch;            // [0]: norm, [1]: cdf tail
  double* tree = scratch + 8;
  B200Q_CHECK(cudaMemsetAsync(flags_dev, 0, sizeof(int), s));
  k_has_nan<<<grid_for(count, 256, 8), 256, 0, s>>>(probs_dev, count, flags_dev);
  B200Q_LAUNCH_CHECK();
  // norm = probs.sum() in numpy's pairwise order; probs /= norm   (sampling.py:510, :526)
  if (int rc = np_sum(probs_dev, count, one, tree, s)) return rc;
  B200Q_CHECK(cudaMemcpyAsync(norm_out_dev, one, sizeof(double), cudaMemcpyDeviceToDevice, s));
  k_div_by<<<grid_for(count, 256, 8), 256, 0, s>>>(probs_dev, one, count);
  B200Q_LAUNCH_CHECK();
  // cdf = probs.cumsum()
  if (mode == B200Q_CDF_EXACT) {
    if (int rc = cumsum_exact(probs_dev, count, nullptr, work, work_bytes, s)) return rc;
  } else if (mode == B200Q_CDF_EXACT_SERIAL) {
    k_cumsum_serial<1024><<<1, 64, 0, s>>>(probs_dev, probs_dev, count);
    B200Q_LAUNCH_CHECK();
  } else if (mode == B200Q_CDF_FAST) {
    const uint64_t nblocks = (count + 2047) / 2048;
    k_scan_block_totals<<<(unsigned)nblocks, 256, 0, s>>>(probs_dev, tree, count);
    B200Q_LAUNCH_CHECK();
    k_scan_totals<<<1, 1024, 0, s>>>(tree, nblocks);
    B200Q_LAUNCH_CHECK();
    k_scan_apply<<<(unsigned)nblocks, 256, 0, s>>>(probs_dev, tree, probs_dev, count);
    B200Q_LAUNCH_CHECK();
  } else {
    set_error("sample: unknown mode %d", mode);
    return 2;
  }
  // cdf /= cdf[-1]
  k_copy_last<<<1, 1, 0, s>>>(probs_dev, one + 1, count);
  B200Q_LAUNCH_CHECK();
  k_div_by<<<grid_for(count, 256, 8), 256, 0, s>>>(probs_dev, one + 1, count);
  B200Q_LAUNCH_CHECK();
  if (shots > 0) {
    k_search<<<(unsigned)((shots + 255) / 256), 256, 0, s>>>(
        probs_dev, count, uniforms_dev, (uint64_t)shots, (long long*)idx_out_dev,
        (long long*)bits_out_dev, m);
    B200Q_LAUNCH_CHECK();
  }
  return 0;
}

// ---- sampler building blocks (the sharded sampler composes them around collectives) -------
int b200q_has_nan(const double* p_dev, int m, int* flag_dev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(m >= 0 && m <= 40 && p_dev && flag_dev, "has_nan: bad arguments");
  const uint64_t count = 1ull << m;
  B200Q_CHECK(cudaMemsetAsync(flag_dev, 0, sizeof(int), s));
  k_has_nan<<<grid_for(count, 256, 8), 256, 0, s>>>(p_dev, count, flag_dev);
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_np_sum(const double* p_dev, int m, double* out_dev, void* work, size_t work_bytes,
                 void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(m >= 0 && m <= 40 && p_dev && out_dev, "np_sum: bad arguments");
  const uint64_t count = 1ull << m;
  B200Q_REQUIRE(work && work_bytes >= kWorkBytes &&
                    (count / 128 + count / (128 * 2047) + 64) * sizeof(double) <=
                        work_bytes - kTermRegion,
                "np_sum: workspace too small for 2^%d values", m);
  double* scratch = (double*)((char*)work + kTermRegion);
  return np_sum(p_dev, count, out_dev, scratch + 8, s);
}

int b200q_div_by(double* p_dev, int m, const double* divisor_dev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(m >= 0 && m <= 40 && p_dev && divisor_dev, "div_by: bad arguments");
  const uint64_t count = 1ull << m;
  k_div_by<<<grid_for(count, 256, 8), 256, 0, s>>>(p_dev, divisor_dev, count);
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_cumsum(double* p_dev, int m, int mode, const double* carry_in_dev, void* work,
                 size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(m >= 0 && m <= 40 && p_dev, "cumsum: bad arguments");
  const uint64_t count = 1ull << m;
  if (mode == B200Q_CDF_EXACT) return cumsum_exact(p_dev, count, carry_in_dev, work, work_bytes, s);
  if (mode == B200Q_CDF_EXACT_SERIAL) {
    k_cumsum_serial<1024><<<1, 64, 0, s>>>(p_dev, p_dev, count, carry_in_dev);
    B200Q_LAUNCH_CHECK();
    return 0;
  }
  B200Q_REQUIRE(mode == B200Q_CDF_FAST, "cumsum: unknown mode %d", mode);
  const uint64_t nblocks = (count + 2047) / 2048;
  B200Q_REQUIRE(work && work_bytes >= kWorkBytes &&
                    (nblocks + 64) * sizeof(double) <= work_bytes - kTermRegion,
                "cumsum: workspace too small for 2^%d values", m);
  double* tree = (double*)((char*)work + kTermRegion) + 8;
  k_scan_block_totals<<<(unsigned)nblocks, 256, 0, s>>>(p_dev, tree, count);
  B200Q_LAUNCH_CHECK();
  k_scan_totals<<<1, 1024, 0, s>>>(tree, nblocks, carry_in_dev);
  B200Q_LAUNCH_CHECK();
  k_scan_apply<<<(unsigned)nblocks, 256, 0, s>>>(p_dev, tree, p_dev, count);
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_search(const double* cdf_dev, int m, const double* uniforms_dev, int64_t shots,
                 int64_t* idx_out_dev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(m >= 0 && m <= 40 && cdf_dev && uniforms_dev && idx_out_dev && shots >= 0,
                "search: bad arguments");
  if (shots == 0) return 0;
  k_search<<<(unsigned)((shots + 255) / 256), 256, 0, s>>>(
      cdf_dev, 1ull << m, uniforms_dev, (uint64_t)shots, (long long*)idx_out_dev, nullptr, m);
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_unpack_bits(const int64_t* idx_dev, int64_t shots, int m, int64_t* bits_out_dev,
                      void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(m >= 0 && m <= 62 && idx_dev && bits_out_dev && shots >= 0,
                "unpack_bits: bad arguments");
  if (shots == 0 || m == 0) return 0;
  k_unpack_bits<<<(unsigned)((shots + 255) / 256), 256, 0, s>>>(
      (const long long*)idx_dev, (uint64_t)shots, m, (long long*)bits_out_dev);
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_apply_tile(void* state, int n, int dtype, int64_t batch, const int* tile_bits, int T,
                     int L, const void* ops_host, int nops, const void* mats_host, int nmat,
                     void* work, size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH(dtype,
           tile_t<float>(state, n, batch, tile_bits, T, L, (const TileOp*)ops_host, nops,
                         (const double2*)mats_host, nmat, work, work_bytes, s),
           tile_t<double>(state, n, batch, tile_bits, T, L, (const TileOp*)ops_host, nops,
                          (const double2*)mats_host, nmat, work, work_bytes, s));
}

int b200q_rtile_geometry(int dtype, int nvec, int* T_out, int* RB_out, int* threads_out) {
  B200Q_REQUIRE(dtype == B200Q_C64 || dtype == B200Q_C128, "rtile_geometry: unknown dtype %d", dtype);
  int T, RB, th;
  rtile_geom(dtype, nvec, T, RB, th);
  if (T_out) *T_out = T;
  if (RB_out) *RB_out = RB;
  if (threads_out) *threads_out = th;
  return 0;
}

int b200q_apply_rtile(void* vec0, void* vec1, int n, int dtype, int64_t batch, const int* tile_bits,
                      int T, int L, const void* ops_host, int nops, const void* mats_host, int nmat,
                      int nslots, int write0, uint64_t base_hi, double scale, double* out_dev,
                      void* work, size_t work_bytes, void* stream) {
  return rtile_dispatch(vec0, vec1, n, dtype, batch, tile_bits, T, L, (const RtOp*)ops_host, nops,
                        (const double2*)mats_host, nmat, nslots, write0, base_hi, scale, out_dev,
                        work, work_bytes, (cudaStream_t)stream);
}

int b200q_expval_csr(const void* state, int n, int dtype, int64_t batch, const int* tbits, int k,
                     const int64_t* indptr_dev, const int64_t* indices_dev, const void* data_dev,
                     double* out_dev, void* work, size_t work_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(state && tbits && indptr_dev && indices_dev && data_dev && out_dev,
                "expval_csr: null argument");
  B200Q_REQUIRE(n >= 1 && n <= B200Q_MAX_BITS && k >= 1 && k <= n && batch >= 1,
                "expval_csr: bad sizes n=%d k=%d", n, k);
  const unsigned ncta = reduce_ncta();
  B200Q_REQUIRE(work && work_bytes >= kTermRegion + (size_t)batch * ncta * sizeof(double),
                "expval_csr: workspace too small");
  CsrArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.k = k;
  for (int j = 0; j < k; ++j) {
    B200Q_REQUIRE(tbits[j] >= 0 && tbits[j] < n && !((a.tmask >> tbits[j]) & 1), "expval_csr: bad bit %d", tbits[j]);
    a.tbits[j] = (int8_t)tbits[j];
    a.tmask |= 1ull << tbits[j];
  }
  double* partials = (double*)((char*)work + kTermRegion);
  dim3 grid(ncta, (unsigned)batch);
  if (dtype == B200Q_DTYPE_C128)
    k_expval_csr<double><<<grid, 256, 0, s>>>((const cx<double>*)state, a, (const long long*)indptr_dev,
                                              (const long long*)indices_dev, (const double2*)data_dev, partials);
  else if (dtype == B200Q_DTYPE_C64)
    k_expval_csr<float><<<grid, 256, 0, s>>>((const cx<float>*)state, a, (const long long*)indptr_dev,
                                             (const long long*)indices_dev, (const double2*)data_dev, partials);
  else { set_error("unknown dtype %d", dtype); return 2; }
  B200Q_LAUNCH_CHECK();
  k_final_reduce<<<(unsigned)batch, 256, 0, s>>>(partials, out_dev, (int)ncta, 1, (int)batch, 1.0);
  B200Q_LAUNCH_CHECK();
  return 0;
}

int b200q_apply_rtile_bcast(void* vec0, void* vec1, int n, int dtype, int64_t batch,
                            const int* tile_bits, int T, int L, const void* ops_host, int nops,
                            const void* mats_host, int nmat, int nslots, int write0,
                            uint64_t base_hi, double scale, double* out_dev, void* work,
                            size_t work_bytes, void* stream) {
  return rtile_dispatch(vec0, vec1, n, dtype, batch, tile_bits, T, L, (const RtOp*)ops_host, nops,
                        (const double2*)mats_host, nmat, nslots, write0, base_hi, scale, out_dev,
                        work, work_bytes, (cudaStream_t)stream, 1);
}

int b200q_adjoint_step(void* vecs, int n, int dtype, int n_bras, const int* tgt_bits, int k,
                       const int* ctrl_bits, const int* ctrl_vals, int nc, const void* adj_host,
                       const void* gen_host, double* out_dev, void* work, size_t work_bytes,
                       void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  B200Q_REQUIRE(k >= 1 && k <= B200Q_MAX_DENSE_K, "adjoint_step: k=%d unsupported (1..3)", k);
  B200Q_REQUIRE(n_bras >= 1 && adj_host, "adjoint_step: bad arguments");
  if (!gen_host) {
    // non-trainable op: one batched dense launch over ket + all bras
    return b200q_apply_matrix(vecs, n, dtype, 1 + n_bras, tgt_bits, k, ctrl_bits, ctrl_vals, nc,
                              adj_host, nullptr, 0, stream);
  }
  const unsigned ncta = reduce_ncta();
  B200Q_REQUIRE(work && work_bytes >= kTermRegion + 2 * (size_t)n_bras * ncta * sizeof(double),
                "adjoint_step: workspace too small");
  GroupArgs a;
  if (int rc = build_group(a, n, tgt_bits, k, ctrl_bits, ctrl_vals, nc)) return rc;
#define ADJ_CASE(T, K) \
  return adjoint_step_k<T, K>(vecs, a, tgt_bits, n_bras, adj_host, gen_host, out_dev, work, s)
  if (dtype == B200Q_DTYPE_C128) {
    switch (k) { case 1: ADJ_CASE(double, 1); case 2: ADJ_CASE(double, 2); default: ADJ_CASE(double, 3); }
  } else if (dtype == B200Q_DTYPE_C64) {
    switch (k) { case 1: ADJ_CASE(float, 1); case 2: ADJ_CASE(float, 2); default: ADJ_CASE(float, 3); }
  }
#undef ADJ_CASE
  set_error("unknown dtype %d", dtype);
  return 2;
}

}  // extern "C"
