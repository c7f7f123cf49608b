// b200q — register-tiled fused segment kernel (K5 + K8 of SURVEY.md section 2c).
//
// One launch = ONE read and ONE write of the state (forward), or of the ket and one bra
// (adjoint reverse sweep), applying a whole *segment* of gates.  A CTA owns a tile of 2^T
// amplitudes (T = log2(THREADS) + RB tile positions: the L lowest index bits plus T-L
// arbitrary higher bits chosen by the host).  Unlike k_tile (tile.cuh), which makes one
// shared-memory pass per gate and is therefore bound by the 128 B/clk shared-memory crossbar
// (measured: 18 % of HBM peak on the 30-qubit ansatz), here every thread keeps 2^RB
// amplitudes in REGISTERS.  The segment is a sequence of *rounds*; a round names the RB tile
// positions that are held in registers ("register bits"); every gate whose targets are
// register bits is applied with no memory traffic at all (controls / parity bits may sit on
// thread bits or outside the tile — they are predicates, not data).  Between rounds the tile
// is transposed through an XOR-swizzled shared-memory buffer (one conflict-free store + load
// of the tile), so the shared-memory traffic per tile is 2 * (rounds-1) tile copies instead
// of 2 per gate.  The CTA's NEXT tile is brought into the transpose buffer by the TMA engine
// (one tensor box, or one bulk copy per contiguous run; completion on an mbarrier) as soon as
// the last transposition of the current tile has left the buffer idle, so the first round reads
// shared memory; the last round stores straight back to global memory, and the host keeps tile
// positions 0..4 on its lanes so every warp store covers whole 512-byte runs.  Single-qubit
// block records are predecoded once per CTA (per-thread matrix choice tabulated, controls
// outside the tile resolved by one ballot per tile) and run with their matrix pinned in
// registers.  Template parameters: T_ (float / double), RB register bits, NV vectors (2 =
// adjoint), THREADS, WS (warp-specialised variant with a producer warp and a landing buffer).
//
// Adjoint mode (NV = 2): vector 0 is the ket, vector 1 a bra; gates are applied to both, and
// RT_GEN records accumulate coef * Im <bra| P |ket> for a Pauli term P of the generator of a
// trainable gate (adjoint_jacobian.py:121-137 fused: no ket_temp, no separate inner-product
// sweep).  Sums are kept per (slot, warp) in shared memory and reduced in a fixed order.
//
// Reference analogue: none (default.qubit sweeps the state once per gate,
// simulate.py:214-235).  Algorithmic bytes per launch: 2*S*NV.
#pragma once
#include <cuda.h>          // CUtensorMap (the descriptor is encoded on the host, rtile.cu)

#include "common.cuh"

namespace b200q {

enum RtKind : int {
  RT_DENSE1 = 0,   // 2x2 on register bit q0
  RT_DENSE2 = 1,   // 4x4 on register bits q0 (matrix MSB) > q1 (matrix LSB)
  RT_CX = 2,       // X on register bit q0
  RT_PARITY = 3,   // amp *= parity ? mat[1] : mat[0]
  RT_DIAG = 4,     // amp *= tab[idx]
  RT_ROUND = 6,    // new register/thread bit assignment (first op of every program)
  RT_GEN = 7,      // adjoint: acc[slot] += coef * Im <bra| P |ket>
};

struct RtGate {     // DENSE1 / DENSE2 / CX / PARITY
  unsigned ctrl_r, cval_r;                 // controls on register bits
  unsigned ctrl_t, cval_t;                 // controls on thread bits
  unsigned long long ctrl_e, cval_e;       // controls outside the tile (global bit positions)
  unsigned par_r, par_t;                   // PARITY: register / thread bits in the parity
  unsigned long long par_e;                // PARITY: bits outside the tile
};
struct RtDiag {     // table index bit b (MSB first): src < 32 register bit, < 64 thread bit
  signed char src[16];                     // (src-32), else external global bit (src-64)
  int pad[8];
};
struct RtRound {
  signed char rbits[8];                    // tile position held by register bit b
  signed char tbits[16];                   // tile position held by thread bit b
  int pad[6];
};
struct RtGen {      // Pauli term: x part on register bits only
  unsigned xr, zr, zt, pad;
  unsigned long long ze;
  double coef;
  int pad2[4];
};
struct __align__(16) RtOp {
  int kind;
  int q0, q1;       // DIAG: q0 = number of index bits.  GEN: q0 = slot, q1 = number of Y factors
  int mat_off;
  union {
    RtGate g;
    RtDiag d;
    RtRound r;
    RtGen p;
  } u;
};
static_assert(sizeof(RtOp) == 64, "RtOp must be 64 bytes (mirrored by ctypes in compiler.py)");

// byte offset of the two CUtensorMap descriptors from the start of the uploaded record table
// (the last 512 bytes of the 4 MiB table region of the workspace, see rtile_host.h)
static constexpr size_t kRtTensorMapOffset = (4ull << 20) - 512;

struct RtArgs {
  int n, T, L, nops, nmat, nslots;
  int write0;                    // write vector 0 back (adjoint passes over several bras)
  int last_round;                // index of the last RT_ROUND record
  int nrounds;                   // number of RT_ROUND records
  int prefetch;                  // 1: the CTA's next tile is brought into the (idle) transpose
                                 // buffer with bulk async copies while the last round computes
  int nruns;                     // runs of consecutive non-tile bits (tile number -> base)
  int nd1;                       // number of predecoded single-qubit block records (<= 32)
  int tma_rank;                  // > 0: the tile is ONE box of a rank-`tma_rank` tensor map over the
                                 // state (dim 0 = the contiguous run, then one dim per group of
                                 // consecutive tile / non-tile bits): one TMA instruction per tile
                                 // instead of one bulk copy per run
  int8_t tma_lo[5], tma_len[5];  // dims 1..rank-1: non-tile group -> coordinate = bits
                                 // [lo, lo+len) of the tile base; tile group -> len = 0
  int8_t hi_bits[16];            // global positions of tile positions L..T-1 (ascending)
  int8_t run_s[16], run_len[16], run_g[16];   // tile-number bits [s, s+len) -> global bits [g, g+len)
  unsigned long long ntiles;     // 2^(n-T)
  unsigned long long base_hi;    // OR-ed into the tile base for "external" predicates
                                 // (the rank's global-qubit bits when the state is sharded)
};

template <int W> __device__ __forceinline__ unsigned rt_sw(unsigned j) {
  unsigned s = 0;
#pragma unroll
  for (int sh = W; sh < 16; sh += W) s ^= (j >> sh);
  return s & ((1u << W) - 1u);
}

__device__ __forceinline__ unsigned long long rt_gscatter(unsigned j, const RtArgs& a) {
  unsigned long long off = j & ((1u << a.L) - 1u);
  const unsigned hi = j >> a.L;
  for (int b = 0; b < a.T - a.L; ++b) off |= (unsigned long long)((hi >> b) & 1u) << a.hi_bits[b];
  return off;
}

// ---- bulk async copy (TMA engine, 1-D) + mbarrier helpers --------------------------------
__device__ __forceinline__ unsigned rt_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void rt_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rt_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rt_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void rt_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RT_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RT_DONE_%=;\n"
      "bra RT_WAIT_%=;\n"
      "RT_DONE_%=:\n"
      "}\n" ::"r"(rt_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void rt_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          rt_smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(rt_smem_u32(bar))
      : "memory");
}
// one box of the tensor map `tm` -> shared memory, completion counted on `bar`
__device__ __forceinline__ void rt_tma_load(void* dst_smem, const CUtensorMap* tm, const int rank,
                                            const int (&c)[5], unsigned long long* bar) {
  const unsigned d = rt_smem_u32(dst_smem), b = rt_smem_u32(bar);
  const unsigned long long t = reinterpret_cast<unsigned long long>(tm);
  switch (rank) {
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(d), "l"(t), "r"(c[0]), "r"(c[1]), "r"(b) : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(d), "l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(b) : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(d), "l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(b) : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                   ::"r"(d), "l"(t), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(b) : "memory");
      break;
  }
}
__device__ __forceinline__ void rt_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// WS (warp-specialised): the CTA has THREADS consumer threads plus ONE producer warp.  The
// producer streams the CTA's next tile into a separate landing buffer with bulk async copies
// (full/empty mbarrier pair) while the consumers work on the current one in registers and in the
// transpose buffer, so the HBM read of tile i+1 overlaps ALL of tile i's compute.  Small tiles
// (128 consumer threads) keep 3 such CTAs resident per SM.
template <typename T_, int RB, int NV, int THREADS, bool WS = false>
struct RtKernel {
  static constexpr int NTH = THREADS + (WS ? 32 : 0);      // threads per CTA
  static constexpr int NA = 1 << RB;
  static constexpr int TB = (THREADS == 128) ? 7 : (THREADS == 256) ? 8 : (THREADS == 512) ? 9 : 10;
  static constexpr int SWW = (sizeof(T_) == 8) ? 3 : 4;    // swizzle width: 16-byte / 8-byte elements
  static constexpr int NW = THREADS / 32;
  using C = cx<T_>;

  // XOR of the per-register-bit constants selected by k (k is a compile-time constant at
  // every call site after unrolling)
  template <typename V> static __device__ __forceinline__ V sel_xor(const V (&c)[RB], int k) {
    V r = 0;
#pragma unroll
    for (int b = 0; b < RB; ++b)
      if ((k >> b) & 1) r ^= c[b];
    return r;
  }

  // In-place complex mat-vec on D amplitudes of one vector.  Every output component is formed
  // as (partial sum over the OTHER inputs) and then ONE final FMA that reads the input it
  // overwrites:  x_r.re <- m_rr.re * x_r.re + t.  The result is born in the register it lives
  // in, so no register moves are needed at the control-flow merges of the record interpreter
  // (the naive "y = M x; x = y" form costs one MOV per FMA, measured with ncu: 47 % of all
  // issued instructions).
  template <int D>
  static __device__ __forceinline__ void matvec_inplace(C (&A)[NA], const int (&ix)[D],
                                                        const C* __restrict__ m) {
    T_ tx[D], ty[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      const C mrr = m[r * D + r];
      tx[r] = -mrr.y * A[ix[r]].y;
      ty[r] = mrr.y * A[ix[r]].x;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        if (c == r) continue;
        const C mrc = m[r * D + c];
        tx[r] = fma(mrc.x, A[ix[c]].x, tx[r]);
        tx[r] = fma(-mrc.y, A[ix[c]].y, tx[r]);
        ty[r] = fma(mrc.y, A[ix[c]].x, ty[r]);
        ty[r] = fma(mrc.x, A[ix[c]].y, ty[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      const T_ d = m[r * D + r].x;
      A[ix[r]].x = fma(d, A[ix[r]].x, tx[r]);
      A[ix[r]].y = fma(d, A[ix[r]].y, ty[r]);
    }
  }

  // Controlled-select dense gates.  For every group of amplitudes the matrix is chosen by the
  // control predicate: `m` (controls satisfied) or `m + D*D` (controls not satisfied; only when
  // HAS0, otherwise the group is left untouched).  A CNOT next to a single-qubit block on its
  // target is folded by the host into ONE such record (M1 = X*U or U*X, M0 = U), so the ansatz'
  // CNOT ring costs no data movement at all.  `tsel` is the thread/external part of the
  // predicate; (cr, cv) the register-bit part, uniform over the warp.  The matrix is re-read
  // from shared memory (broadcast) per group: 64 data registers leave no room to pin it.
  template <int Q, bool HAS0>
  static __device__ __forceinline__ void dense1(C (&A)[NV][NA], const C* __restrict__ m,
                                                const bool tsel, const unsigned cr,
                                                const unsigned cv) {
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if ((k >> Q) & 1) continue;
      const bool sel = tsel && ((k & cr) == cv);
      if (!HAS0 && !sel) continue;
      const C* mm = (HAS0 && !sel) ? m + 4 : m;
      const int ix[2] = {k, k | (1 << Q)};
#pragma unroll
      for (int v = 0; v < NV; ++v) matvec_inplace<2>(A[v], ix, mm);
    }
  }

  template <int Q0, int Q1, bool HAS0>   // Q0 > Q1; Q0 is the matrix MSB
  static __device__ __forceinline__ void dense2(C (&A)[NV][NA], const C* __restrict__ m,
                                                const bool tsel, const unsigned cr,
                                                const unsigned cv) {
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if (((k >> Q0) & 1) || ((k >> Q1) & 1)) continue;
      const bool sel = tsel && ((k & cr) == cv);
      if (!HAS0 && !sel) continue;
      const C* mm = (HAS0 && !sel) ? m + 16 : m;
      const int ix[4] = {k, k | (1 << Q1), k | (1 << Q0), k | (1 << Q0) | (1 << Q1)};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        asm volatile("" ::: "memory");     // do not hoist the 16 matrix entries out of the quads
        matvec_inplace<4>(A[v], ix, mm);
      }
    }
  }

  // Conditional in-place exchange of two reals: (a, b) <- m ? (b, a) : (a, b), m = 0 or ~0, as
  // three XOR/AND operations per 32-bit word.  Branch-free and born in place: a register swap
  // written as moves under a (divergent or uniform) branch made ptxas re-shuffle the whole
  // amplitude file at the interpreter's control-flow merges (ncu on the adjoint kernel: 52 % of
  // all issued instructions were MOVs).
  static __device__ __forceinline__ void cswap(T_& a, T_& b, const unsigned m) {
    if constexpr (sizeof(T_) == 8) {
      unsigned al = (unsigned)__double2loint(a), ah = (unsigned)__double2hiint(a);
      unsigned bl = (unsigned)__double2loint(b), bh = (unsigned)__double2hiint(b);
      const unsigned tl = (al ^ bl) & m, th = (ah ^ bh) & m;
      al ^= tl; bl ^= tl; ah ^= th; bh ^= th;
      a = __hiloint2double((int)ah, (int)al);
      b = __hiloint2double((int)bh, (int)bl);
    } else {
      unsigned ua = __float_as_uint(a), ub = __float_as_uint(b);
      const unsigned t = (ua ^ ub) & m;
      a = __uint_as_float(ua ^ t);
      b = __uint_as_float(ub ^ t);
    }
  }

  template <int Q>
  static __device__ __forceinline__ void cx_gate(C (&A)[NV][NA], const bool tsel, unsigned cr,
                                                 unsigned cv) {
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if ((k >> Q) & 1) continue;
      const unsigned m = (tsel && ((k & cr) == cv)) ? 0xffffffffu : 0u;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        cswap(A[v][k].x, A[v][k | (1 << Q)].x, m);
        cswap(A[v][k].y, A[v][k | (1 << Q)].y, m);
      }
    }
  }

  // a <- p * a, both components formed by a final FMA on the overwritten input
  static __device__ __forceinline__ void cmul_inplace(C& a, const C p) {
    const T_ tx = -p.y * a.y, ty = p.y * a.x;
    a.x = fma(p.x, a.x, tx);
    a.y = fma(p.x, a.y, ty);
  }

  // sum_k sign_k * {Re, Im}(conj(bra_k) * ket_{k ^ PM}),  sign_k = (-1)^(popc((k ^ PM) & zr) + tpar).
  // PM (the X part of the generator term on register bits) is a compile-time constant: the
  // partner amplitude is named directly, nothing moves.  The sign is applied by flipping the
  // sign bit of the product.
  template <int PM>
  static __device__ __forceinline__ double gen_term(const C (&A)[NV][NA], const unsigned zr,
                                                    const unsigned tpar, const bool odd) {
    unsigned zb[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) zb[b] = (zr >> b) & 1u;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      const C b = A[NV - 1][k], x = A[0][k ^ PM];
      const unsigned par = sel_xor(zb, k ^ PM) ^ tpar;
      double val;
      if (odd) val = fma((double)b.y, (double)x.y, (double)b.x * (double)x.x);    // Re
      else val = fma(-(double)b.y, (double)x.x, (double)b.x * (double)x.y);       // Im
      val = __hiloint2double(__double2hiint(val) ^ (int)(par << 31), __double2loint(val));
      acc += val;
    }
    return acc;
  }

  static __device__ __forceinline__ double gen_dispatch(const C (&A)[NV][NA], const unsigned xr,
                                                        const unsigned zr, const unsigned tpar,
                                                        const bool odd) {
    switch (xr) {
#define RT_GEN_CASE(PM) case PM: if constexpr (PM < NA) return gen_term<PM>(A, zr, tpar, odd); break;
      RT_GEN_CASE(0) RT_GEN_CASE(1) RT_GEN_CASE(2) RT_GEN_CASE(3) RT_GEN_CASE(4) RT_GEN_CASE(5)
      RT_GEN_CASE(6) RT_GEN_CASE(7) RT_GEN_CASE(8) RT_GEN_CASE(9) RT_GEN_CASE(10) RT_GEN_CASE(11)
      RT_GEN_CASE(12) RT_GEN_CASE(13) RT_GEN_CASE(14) RT_GEN_CASE(15)
#undef RT_GEN_CASE
      default: break;
    }
    return 0.0;
  }

  // ---- single-qubit dense blocks with the matrix PINNED in registers -------------------------
  // The interpreter's first version re-read the 2x2 matrix from shared memory for every
  // amplitude pair (4 LDS.128 per 16 FP64 instructions; ncu: LDS 12 % of issued instructions,
  // short-scoreboard the second stall reason).  Here the matrix is read once per record (or
  // once per control half) into 8 registers and the 2^(RB-1) pairs run back to back as
  // independent FMA chains.
  //   d1_uniform: the same matrix for every pair of this thread (no control on a register bit);
  //   d1_regctl : one control on register bit CB — pairs with that bit clear use m_b0, pairs
  //               with it set use m_b1 (nullptr: leave those pairs untouched).

  // Scheduling fences: the amplitudes are "touched" by an empty volatile asm, so the front end
  // cannot hoist or common work across this point.  Without them the 2^(RB-1) independent
  // pairs of a record are software-pipelined so deep that the kernel wants 194 registers and,
  // at the 128-register cap of 2 CTAs/SM, spills amplitudes and re-shuffles the whole register
  // file at every control-flow merge of the interpreter (checked with -Xptxas -v / nvdisasm).
  static __device__ __forceinline__ void sfence(C& a, C& b, C& c, C& d) {
    if constexpr (sizeof(T_) == 8)
      asm volatile("" : "+d"(a.x), "+d"(a.y), "+d"(b.x), "+d"(b.y), "+d"(c.x), "+d"(c.y), "+d"(d.x), "+d"(d.y));
    else
      asm volatile("" : "+f"(a.x), "+f"(a.y), "+f"(b.x), "+f"(b.y), "+f"(c.x), "+f"(c.y), "+f"(d.x), "+f"(d.y));
  }
  static __device__ __forceinline__ void sfence_all(C (&A)[NV][NA]) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < NA; k += 4) sfence(A[v][k], A[v][k + 1], A[v][k + 2], A[v][k + 3]);
  }

  // 16-byte (complex128) / 8-byte (complex64) shared-memory load by 32-bit address
  static __device__ __forceinline__ C lds_c(const unsigned addr) {
    C r;
    if constexpr (sizeof(T_) == 8) asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
    else asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
    return r;
  }

  template <int Q>
  static __device__ __forceinline__ void d1_uniform(C (&A)[NV][NA], const unsigned m) {
    const C M[4] = {lds_c(m), lds_c(m + sizeof(C)), lds_c(m + 2 * sizeof(C)), lds_c(m + 3 * sizeof(C))};
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if ((k >> Q) & 1) continue;
      const int ix[2] = {k, k | (1 << Q)};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if ((cnt++ % 2) == 0) sfence_all(A);
        matvec_inplace<2>(A[v], ix, M);
      }
    }
  }

  template <int Q, int CB, int BV>
  static __device__ __forceinline__ void d1_half(C (&A)[NV][NA], const unsigned m) {
    const C M[4] = {lds_c(m), lds_c(m + sizeof(C)), lds_c(m + 2 * sizeof(C)), lds_c(m + 3 * sizeof(C))};
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if (((k >> Q) & 1) || (((k >> CB) & 1) != BV)) continue;
      const int ix[2] = {k, k | (1 << Q)};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if ((cnt++ % 2) == 0) sfence_all(A);
        matvec_inplace<2>(A[v], ix, M);
      }
    }
  }

  // matrices are named by their index in units of two table entries (255 = leave untouched)
  static constexpr unsigned D1_SKIP = 255u;
  template <int Q, int CB>
  static __device__ __forceinline__ void d1_regctl(C (&A)[NV][NA], const unsigned mats_s,
                                                   const unsigned b0, const unsigned b1) {
    if constexpr (Q < RB && CB < RB) {
      if constexpr (Q == CB) {
        if (b0 != D1_SKIP) d1_uniform<Q>(A, mats_s + b0 * 2u * (unsigned)sizeof(C));
      } else {
        if (b0 != D1_SKIP) d1_half<Q, CB, 0>(A, mats_s + b0 * 2u * (unsigned)sizeof(C));
        if (b1 != D1_SKIP) d1_half<Q, CB, 1>(A, mats_s + b1 * 2u * (unsigned)sizeof(C));
      }
    }
  }

  // tile number -> global offset of the tile (scatter over the non-tile bits, run by run)
  static __device__ __forceinline__ unsigned long long tile_base(const RtArgs& a,
                                                                 const unsigned long long t) {
    unsigned long long base = 0;
    for (int r = 0; r < a.nruns; ++r)
      base |= ((t >> a.run_s[r]) & ((1ull << a.run_len[r]) - 1ull)) << a.run_g[r];
    return base;
  }

  // bring tile `t` of every vector into the (currently idle) transpose buffer: one bulk async
  // copy per contiguous run of 2^L amplitudes, completion counted on `bar`
  static __device__ __forceinline__ void csync() {         // barrier over the consumer threads
    if constexpr (WS) asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    else __syncthreads();
  }

  static __device__ __forceinline__ void issue_tile_copies(const RtArgs& a, C* vec0, C* vec1,
                                                           C* tile, const unsigned long long t,
                                                           unsigned long long* bar,
                                                           const unsigned first = threadIdx.x,
                                                           const unsigned stride = THREADS) {
    const unsigned tsize = 1u << a.T;
    const unsigned run = 1u << a.L;
    const unsigned long long base = tile_base(a, t);
    for (unsigned j = first; j < (tsize >> a.L) * NV; j += stride) {
      const unsigned v = j >> (a.T - a.L), r = j & ((tsize >> a.L) - 1u);
      rt_bulk_g2s(tile + (size_t)v * tsize + ((size_t)r << a.L),
                  (v ? vec1 : vec0) + base + rt_gscatter(r << a.L, a), run * (unsigned)sizeof(C),
                  bar);
    }
  }

  // Fetch tile `t` of every vector into `tile` (non-WS kernels): every thread calls this once it
  // no longer needs the buffer.  With a tensor map it is one TMA instruction per vector, issued
  // by thread 0; otherwise one bulk copy per contiguous run, spread over the threads.
  static __device__ __forceinline__ void fetch_tile(const RtArgs& a, C* vec0, C* vec1, C* tile,
                                                    const unsigned long long t,
                                                    unsigned long long* bar, const CUtensorMap* tm0,
                                                    const CUtensorMap* tm1) {
    const unsigned bytes = (unsigned)(NV * sizeof(C)) << a.T;
    rt_fence_proxy_async();
    // (forward kernels only: in the two-vector adjoint kernel the extra path cost registers —
    // 416 -> 532 bytes of spills, 1.71 -> 1.82 s per Jacobian — for no measurable gain)
    if (NV == 1 && a.tma_rank > 0) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned long long base = tile_base(a, t);
        int c[5] = {0, 0, 0, 0, 0};
        for (int r = 1; r < a.tma_rank; ++r)
          if (a.tma_len[r]) c[r] = (int)((base >> a.tma_lo[r]) & ((1ull << a.tma_len[r]) - 1ull));
        rt_mbar_expect_tx(bar, bytes);
        rt_tma_load(tile, tm0, a.tma_rank, c, bar);
      }
    } else {
      if (threadIdx.x == 0) rt_mbar_expect_tx(bar, bytes);
      __syncthreads();
      issue_tile_copies(a, vec0, vec1, tile, t, bar);
    }
  }

  static __device__ __forceinline__ void run(const RtArgs& __restrict__ a, C* __restrict__ v0,
                                             C* __restrict__ v1, const RtOp* __restrict__ ops_g,
                                             const double2* __restrict__ mats_g,
                                             const long long mat_bstride,
                                             double* __restrict__ partials) {
    // the two tensor maps (vector 0, vector 1) sit at a fixed offset behind the uploaded records
    const CUtensorMap* tm0 = reinterpret_cast<const CUtensorMap*>(
        reinterpret_cast<const char*>(ops_g) + kRtTensorMapOffset);
    const CUtensorMap* tm1 = nullptr;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned tsize = 1u << a.T;
    C* tile = reinterpret_cast<C*>(smem_raw);                                  // NV * 2^T
    C* land = WS ? tile + (size_t)NV * tsize : tile;                            // WS: NV * 2^T more
    C* mats = land + (size_t)NV * tsize;                                        // nmat (padded even)
    RtOp* ops = reinterpret_cast<RtOp*>(mats + ((a.nmat + 1) & ~1));            // nops
    unsigned long long* koff = reinterpret_cast<unsigned long long*>(ops + a.nops);  // 2 * NA
    unsigned long long* toff = koff + 2 * NA;                                   // 2 * THREADS
    unsigned long long* bar = toff + 2 * THREADS;                               // full, empty
    double* accs = reinterpret_cast<double*>(bar + 2);                          // nslots * NW
    unsigned short* tsl = reinterpret_cast<unsigned short*>(accs + a.nslots * NW);   // nrounds * THREADS
    unsigned short* rsl = tsl + (size_t)a.nrounds * THREADS;                    // nrounds * 8
    unsigned short* ldsl = rsl + (size_t)a.nrounds * 8;                         // THREADS + NA: round-0
                                                                                // layout, unswizzled
    unsigned* rd = reinterpret_cast<unsigned*>(ldsl + THREADS + NA);            // nops + 1 descriptors
    unsigned* kst32 = rd + a.nops + 1;                                          // NA: store offsets
    unsigned short* ent = reinterpret_cast<unsigned short*>(kst32 + NA);        // nd1 * THREADS
    unsigned char* d1rec = reinterpret_cast<unsigned char*>(ent + (size_t)a.nd1 * THREADS);   // 32
    const unsigned tid = threadIdx.x;

    const double2* mg = mats_g + (long long)blockIdx.y * mat_bstride;
    for (int i = tid; i < a.nmat; i += NTH) mats[i] = make_cx<T_>((T_)mg[i].x, (T_)mg[i].y);
    for (int i = tid; i < a.nops; i += NTH) ops[i] = ops_g[i];
    for (int i = tid; i < a.nslots * NW; i += NTH) accs[i] = 0.0;
    if (tid == 0) {
      rt_mbar_init(bar, 1);
      rt_mbar_init(bar + 1, NW);        // empty: one arrival per consumer warp
    }
    __syncthreads();
    // Record descriptors (one uniform 32-bit load per record in the tile loop):
    //   byte 0: 0xF0 | kind for records decoded from their 64-byte form, or, for single-qubit
    //           blocks with at most one control on a register bit, the handler Q * 8 + CB
    //           (CB = Q: no register control);  byte 1: predecode slot;  bytes 2-3: the
    //           (b0, b1) matrix pair used when the record's controls OUTSIDE the tile fail.
    // The (b0, b1) pair of every (slot, thread) for the case that they hold is tabulated once
    // per CTA in `ent`: in the tile loop such a record costs one table load and a jump instead
    // of ~60 instructions of mask tests and pointer selects.
    if (tid == 0) {
      int r = 0, s1 = 0;
      for (int i = 0; i < a.nops; ++i) {
        const int k = ops[i].kind & 0xff;
        unsigned d = 0xF0u | (unsigned)k;
        if (k == RT_ROUND) ops[i].q0 = r++;
        else if (k == RT_DENSE1) {
          const unsigned cr = ops[i].u.g.ctrl_r;
          const bool has0 = (ops[i].kind >> 8) & 1;
          const unsigned idx0 = (unsigned)(ops[i].mat_off >> 1) + 2u;
          if ((cr & (cr - 1u)) == 0 && s1 < a.nd1 && s1 < 32 && !(ops[i].mat_off & 1) && idx0 < D1_SKIP) {
            const unsigned cb = cr ? (unsigned)(__ffs((int)cr) - 1) : (unsigned)ops[i].q0;
            const unsigned e0 = has0 ? idx0 : D1_SKIP;
            d = ((unsigned)ops[i].q0 * 8u + cb) | ((unsigned)s1 << 8) | (e0 << 16) | (e0 << 24);
            d1rec[s1++] = (unsigned char)i;
          }
        }
        rd[i] = d;
      }
      rd[a.nops] = 0xFFu | ((unsigned)s1 << 8);
    }
    __syncthreads();
    const unsigned nd1u = rd[a.nops] >> 8;           // predecoded records
    const unsigned mats_s = rt_smem_u32(mats);

    C* const vec[2] = {v0 + ((unsigned long long)blockIdx.y << a.n),
                       NV > 1 ? v1 + ((unsigned long long)blockIdx.y << a.n) : nullptr};

    // Per-CTA tables.  Global offsets of the first (load) and last (store) round layouts are
    // tile independent: per-thread parts in toff, per-register-index parts in koff.  The
    // swizzled shared-memory slot of every (round, thread) and (round, register bit) is
    // computed once here instead of once per tile and round.
    if (tid < THREADS) {
      const RtOp& f = ops[0];
      const RtOp& l = ops[a.last_round];
      unsigned tj = 0, tl = 0;
      for (int b = 0; b < TB; ++b) {
        tj |= ((tid >> b) & 1u) << f.u.r.tbits[b];
        tl |= ((tid >> b) & 1u) << l.u.r.tbits[b];
      }
      toff[tid] = rt_gscatter(tj, a);
      toff[THREADS + tid] = rt_gscatter(tl, a);
      ldsl[tid] = (unsigned short)tj;
      if (tid < 2 * NA) {
        const RtOp& r = tid < NA ? f : l;
        const unsigned k = tid & (NA - 1);
        unsigned j = 0;
        for (int b = 0; b < RB; ++b) j |= ((k >> b) & 1u) << r.u.r.rbits[b];
        koff[tid] = rt_gscatter(j, a);
        if (tid < NA) ldsl[THREADS + tid] = (unsigned short)j;
        else kst32[k] = (unsigned)koff[tid];          // store offsets, 32-bit (used when n <= 32)
      }
      for (unsigned s1 = 0; s1 < nd1u; ++s1) {
        const RtOp& r = ops[d1rec[s1]];
        const RtGate& g = r.u.g;
        const bool has0 = (r.kind >> 8) & 1;
        const unsigned idx1 = (unsigned)(r.mat_off >> 1), idx0 = idx1 + 2u;
        const unsigned mo = has0 ? idx0 : D1_SKIP;                          // control fails
        const unsigned ms = ((tid & g.ctrl_t) == g.cval_t) ? idx1 : mo;     // control holds
        unsigned b0 = ms, b1 = ms;
        if (g.ctrl_r) { b1 = g.cval_r ? ms : mo; b0 = g.cval_r ? mo : ms; }
        ent[s1 * THREADS + tid] = (unsigned short)(b0 | (b1 << 8));
      }
      for (int i = 0; i < a.nops; ++i) {
        const RtOp& r = ops[i];
        if ((r.kind & 0xff) != RT_ROUND) continue;
        unsigned t2 = 0;
        for (int b = 0; b < TB; ++b) t2 |= ((tid >> b) & 1u) << r.u.r.tbits[b];
        tsl[r.q0 * THREADS + tid] = (unsigned short)(t2 ^ rt_sw<SWW>(t2));
        if (tid < RB) {
          const unsigned rj = 1u << r.u.r.rbits[tid];
          rsl[r.q0 * 8 + tid] = (unsigned short)(rj ^ rt_sw<SWW>(rj));
        }
      }
    }
    __syncthreads();

    if constexpr (WS) {
      if (tid >= THREADS) {
        // ---- producer warp ------------------------------------------------------------------
        const unsigned lane = tid & 31u;
        unsigned i = 0;
        for (unsigned long long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++i) {
          if (i > 0) rt_mbar_wait(bar + 1, (i - 1u) & 1u);      // consumers drained tile i-1
          if (lane == 0) rt_mbar_expect_tx(bar, (unsigned)(NV * tsize * sizeof(C)));
          __syncwarp();
          issue_tile_copies(a, vec[0], vec[1], land, t, bar, lane, 32u);
        }
        return;
      }
    }
    const bool pf = !WS && a.prefetch != 0;
    if (pf) fetch_tile(a, vec[0], vec[1], tile, blockIdx.x, bar, tm0, tm1);   // the CTA's first tile
    unsigned phase = 0;

    C A[NV][NA];

    for (unsigned long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
      const unsigned long long base = tile_base(a, t);
      const unsigned long long baseE = base | a.base_hi;
      const bool more = t + gridDim.x < a.ntiles;
      bool issued = false;             // next tile's copies already in flight

      // ---- load (round 0 layout) ------------------------------------------------------------
      if constexpr (WS) {
        rt_mbar_wait(bar, phase);
        phase ^= 1u;
        const unsigned tj = ldsl[tid];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int k = 0; k < NA; ++k) A[v][k] = land[v * tsize + (tj | ldsl[THREADS + k])];
        __syncwarp();
        if ((tid & 31u) == 0)          // this warp has drained the landing buffer
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rt_smem_u32(bar + 1)) : "memory");
      } else if (pf) {
        rt_mbar_wait(bar, phase);
        phase ^= 1u;
        const unsigned tj = ldsl[tid];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int k = 0; k < NA; ++k) A[v][k] = tile[v * tsize + (tj | ldsl[THREADS + k])];
        if (a.last_round == 0 && more) {
          // single-round segment: the buffer is idle for the whole compute phase
          fetch_tile(a, vec[0], vec[1], tile, t + gridDim.x, bar, tm0, tm1);
          issued = true;
        }
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const C* src = vec[v] + base + toff[tid];
#pragma unroll
          for (int k = 0; k < NA; ++k) A[v][k] = src[koff[k]];
        }
      }
      int cur = 0;                     // round index (not record index) of the current layout

      // controls outside the tile of the predecoded records: one ballot per tile
      unsigned emask;
      {
        const unsigned lane = tid & 31u;
        bool p = false;
        if (lane < nd1u) {
          const RtGate& g = ops[d1rec[lane]].u.g;
          p = (baseE & g.ctrl_e) == g.cval_e;
        }
        emask = __ballot_sync(0xffffffffu, p);
      }

      unsigned rdn = rd[1];
      for (int o = 1; o < a.nops; ++o) {
        const unsigned rdc = rdn;
        rdn = rd[o + 1];               // next descriptor: in flight while this record computes
        if ((rdc & 0xffu) < 0x40u) {
          // ---- predecoded single-qubit block ---------------------------------------------------
          const unsigned slot = (rdc >> 8) & 0xffu;
          const unsigned e = ((emask >> slot) & 1u) ? (unsigned)ent[slot * THREADS + tid] : (rdc >> 16);
          const unsigned b0 = e & 0xffu, b1 = (e >> 8) & 0xffu;
          switch (rdc & 0xffu) {
#define RT_D1C(Q, CB) case Q * 8 + CB: d1_regctl<Q, CB>(A, mats_s, b0, b1); break;
            RT_D1C(0, 0) RT_D1C(0, 1) RT_D1C(0, 2) RT_D1C(0, 3) RT_D1C(0, 4)
            RT_D1C(1, 0) RT_D1C(1, 1) RT_D1C(1, 2) RT_D1C(1, 3) RT_D1C(1, 4)
            RT_D1C(2, 0) RT_D1C(2, 1) RT_D1C(2, 2) RT_D1C(2, 3) RT_D1C(2, 4)
            RT_D1C(3, 0) RT_D1C(3, 1) RT_D1C(3, 2) RT_D1C(3, 3) RT_D1C(3, 4)
            RT_D1C(4, 0) RT_D1C(4, 1) RT_D1C(4, 2) RT_D1C(4, 3) RT_D1C(4, 4)
#undef RT_D1C
            default: break;
          }
          continue;
        }
        const RtOp& op = ops[o];
        const int kind = op.kind & 0xff;
        if (kind <= RT_CX) {
          // ---- gates on register bits, with controls anywhere --------------------------------
          const RtGate& g = op.u.g;
          const bool tsel = ((tid & g.ctrl_t) == g.cval_t) && ((baseE & g.ctrl_e) == g.cval_e);
          const unsigned cr = g.ctrl_r, cv = g.cval_r;
          const bool has0 = (op.kind >> 8) & 1;
          const C* m = mats + op.mat_off;
          if (kind == RT_DENSE1) {
            if (has0) {
              switch (op.q0) {
                case 0: dense1<0, true>(A, m, tsel, cr, cv); break;
                case 1: if constexpr (RB > 1) dense1<1, true>(A, m, tsel, cr, cv); break;
                case 2: if constexpr (RB > 2) dense1<2, true>(A, m, tsel, cr, cv); break;
                case 3: if constexpr (RB > 3) dense1<3, true>(A, m, tsel, cr, cv); break;
                case 4: if constexpr (RB > 4) dense1<4, true>(A, m, tsel, cr, cv); break;
                default: break;
              }
            } else if (tsel) {
              switch (op.q0) {
                case 0: dense1<0, false>(A, m, true, cr, cv); break;
                case 1: if constexpr (RB > 1) dense1<1, false>(A, m, true, cr, cv); break;
                case 2: if constexpr (RB > 2) dense1<2, false>(A, m, true, cr, cv); break;
                case 3: if constexpr (RB > 3) dense1<3, false>(A, m, true, cr, cv); break;
                case 4: if constexpr (RB > 4) dense1<4, false>(A, m, true, cr, cv); break;
                default: break;
              }
            }
          } else if (kind == RT_CX) {
            switch (op.q0) {
              case 0: cx_gate<0>(A, tsel, cr, cv); break;
              case 1: if constexpr (RB > 1) cx_gate<1>(A, tsel, cr, cv); break;
              case 2: if constexpr (RB > 2) cx_gate<2>(A, tsel, cr, cv); break;
              case 3: if constexpr (RB > 3) cx_gate<3>(A, tsel, cr, cv); break;
              case 4: if constexpr (RB > 4) cx_gate<4>(A, tsel, cr, cv); break;
              default: break;
            }
          } else {   // RT_DENSE2
            switch (op.q0 * 8 + op.q1) {
#define RT_D2_CASE(Q0, Q1)                                                     \
  case Q0 * 8 + Q1:                                                            \
    if constexpr (Q0 < RB) {                                                   \
      if (has0) dense2<Q0, Q1, true>(A, m, tsel, cr, cv);                      \
      else if (tsel) dense2<Q0, Q1, false>(A, m, true, cr, cv);                \
    }                                                                          \
    break;
              RT_D2_CASE(1, 0)
              RT_D2_CASE(2, 0) RT_D2_CASE(2, 1)
              RT_D2_CASE(3, 0) RT_D2_CASE(3, 1) RT_D2_CASE(3, 2)
              RT_D2_CASE(4, 0) RT_D2_CASE(4, 1) RT_D2_CASE(4, 2) RT_D2_CASE(4, 3)
#undef RT_D2_CASE
              default: break;
            }
          }
          continue;
        }
        if (kind == RT_ROUND) {
          // transpose through shared memory: store in the old layout, load in the new one
          // (slots come from the per-CTA tables)
          const int nr = op.q0;
          unsigned tslot = tsl[cur * THREADS + tid], rs[RB];
#pragma unroll
          for (int b = 0; b < RB; ++b) rs[b] = rsl[cur * 8 + b];
          csync();
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int k = 0; k < NA; ++k) tile[v * tsize + (tslot ^ sel_xor(rs, k))] = A[v][k];
          tslot = tsl[nr * THREADS + tid];
#pragma unroll
          for (int b = 0; b < RB; ++b) rs[b] = rsl[nr * 8 + b];
          csync();
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int k = 0; k < NA; ++k) A[v][k] = tile[v * tsize + (tslot ^ sel_xor(rs, k))];
          cur = nr;
          if (pf && o == a.last_round && more) {
            // the buffer is idle from here to the next tile's first transposition: fetch the
            // next tile into it while this tile's last round computes and stores
            fetch_tile(a, vec[0], vec[1], tile, t + gridDim.x, bar, tm0, tm1);
            issued = true;
          }
          continue;
        }
        if (kind == RT_PARITY) {
          const RtGate& g = op.u.g;
          if (((tid & g.ctrl_t) != g.cval_t) || ((baseE & g.ctrl_e) != g.cval_e)) continue;
          const unsigned cr = g.ctrl_r, cv = g.cval_r;
          const unsigned tpar = (__popc(tid & g.par_t) + __popcll(baseE & g.par_e)) & 1u;
          // the thread's own parity decides which phase is "even" for it
          const C pe = mats[op.mat_off + tpar], po = mats[op.mat_off + (tpar ^ 1u)];
          const unsigned pr = g.par_r;
#pragma unroll
          for (int k = 0; k < NA; ++k) {
            if ((k & cr) != cv) continue;
            const bool odd = __popc((unsigned)k & pr) & 1u;       // uniform over the warp
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              if (odd) cmul_inplace(A[v][k], po);
              else cmul_inplace(A[v][k], pe);
            }
          }
          continue;
        }
        if (kind == RT_DIAG) {
          const C* tab = mats + op.mat_off;
          const int nd = op.q0;
          // table index = thread/external part | OR of per-register-bit contributions
          unsigned idx0 = 0, rc[RB];
#pragma unroll
          for (int b = 0; b < RB; ++b) rc[b] = 0;
          for (int b = 0; b < nd; ++b) {
            const int s = op.u.d.src[b];
            const unsigned bit = 1u << (nd - 1 - b);
            if (s >= 64) idx0 |= ((baseE >> (s - 64)) & 1ull) ? bit : 0u;
            else if (s >= 32) idx0 |= ((tid >> (s - 32)) & 1u) ? bit : 0u;
            else {
#pragma unroll
              for (int r = 0; r < RB; ++r)
                if (s == r) rc[r] |= bit;
            }
          }
#pragma unroll
          for (int k = 0; k < NA; ++k) {
            const C d = tab[idx0 | sel_xor(rc, k)];
#pragma unroll
            for (int v = 0; v < NV; ++v) cmul_inplace(A[v][k], d);
          }
          continue;
        }
        if (kind == RT_GEN) {
          if constexpr (NV > 1) {
            const unsigned tpar0 = (__popc(tid & op.u.p.zt) + __popcll(baseE & op.u.p.ze)) & 1u;
            const int ny = op.q1;
            const unsigned tpar = tpar0 ^ (unsigned)((ny >> 1) & 1);      // i^2 = -1, i^3 = -i
            const bool odd = ny & 1;
            const unsigned zr = op.u.p.zr;
            const unsigned xr = op.u.p.xr;
            double s = gen_dispatch(A, xr, zr, tpar, odd);
            s *= op.u.p.coef;
            s = warp_sum(s);
            if ((tid & 31u) == 0) accs[op.q0 * NW + (tid >> 5)] += s;
          }
          continue;
        }
      }

      // ---- store (last round layout) -----------------------------------------------------
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (v == 0 && !a.write0) continue;
        C* dst = vec[v] + base + toff[THREADS + tid];
        if (a.n <= 32) {
#pragma unroll
          for (int k = 0; k < NA; ++k) dst[kst32[k]] = A[v][k];
        } else {
#pragma unroll
          for (int k = 0; k < NA; ++k) dst[koff[NA + k]] = A[v][k];
        }
      }
      (void)issued;
    }

    if (a.nslots > 0) {
      csync();
      for (int s = tid; s < a.nslots; s += THREADS) {
        double acc = 0.0;
        for (int w = 0; w < NW; ++w) acc += accs[s * NW + w];
        partials[((size_t)blockIdx.y * a.nslots + s) * gridDim.x + blockIdx.x] = acc;
      }
    }
  }
};

template <typename T_, int RB, int NV, int THREADS, int MINB, bool WS = false>
__global__ void __launch_bounds__(THREADS + (WS ? 32 : 0), MINB)
k_rtile(const __grid_constant__ RtArgs a, cx<T_>* __restrict__ v0, cx<T_>* __restrict__ v1,
        const RtOp* __restrict__ ops_g, const double2* __restrict__ mats_g,
        const long long mat_bstride, double* __restrict__ partials) {
  RtKernel<T_, RB, NV, THREADS, WS>::run(a, v0, v1, ops_g, mats_g, mat_bstride, partials);
}

}  // namespace b200q
