// b200q — structure-specialised fused segment kernel (round 2; replaces the record interpreter
// of rtile.cuh on the hot path).
//
// One launch = ONE read and ONE write of the state (forward), or of the ket and one bra
// (adjoint reverse sweep), applying a whole *segment* of gates to 2^T-amplitude tiles held in
// REGISTERS (2^RB amplitudes per thread), exactly like rtile.cuh: rounds of register-resident
// gates separated by XOR-swizzled shared-memory transpositions, the next tile fetched by the TMA
// engine into the idle transpose buffer.  What differs is HOW the program reaches the SM:
//
//   rtile.cuh interprets 64-byte records (kind / register bit / control masks decoded per tile).
//   ncu on the 30-qubit ansatz: 53 % (forward) / 72 % (adjoint) of the issued instructions were
//   decode, address and predicate work, and on B200 a DFMA holds the issue port for two cycles
//   while every other instruction takes one (tools/micro/fp64_forms.cu: F12 + 8 SELs per pair
//   runs exactly as fast as F16), so those instructions are NOT free.
//
//   Here the segment's STRUCTURE — round layouts, record kinds, register bits, control
//   locations, Pauli masks of generator terms — is a set of compile-time constants: the host
//   (pennylane_b200/segjit.py) emits the tile body as a list of calls into the templates below
//   ("sk_body.inc") plus a configuration header ("sk_config.inc"), NVRTC compiles the pair once
//   per structure (cached by hash, in memory and on disk), and only the VALUES (matrix
//   coefficients, tile bit positions, tensor maps) are launch arguments.  A CNOT whose control
//   and target are register bits is a renaming of registers (zero instructions); controls on
//   thread bits or outside the tile are predicates evaluated once.
//
//   Single-qubit blocks are applied in a NORMALISED form.  U = s * diag(1,l) K(t) diag(1,r) with
//   K(t) = [[1,-t],[t,1]] (|t| <= 1; the "sin pivot" [[t,-1],[1,t]] otherwise, a run-time flag
//   so that parameter updates never change the structure) or its imaginary twin
//   [[1,-it],[-it,1]] (RX-like blocks).  The scalar s is common to every amplitude: the host
//   multiplies the s of all records of the segment together and the kernel applies the product
//   once, at the end of the segment (sk_scale).  Cost per amplitude pair: 4 FP64 instructions
//   for K, +4 per non-trivial phase (RZ.RY: 8, RY / RX: 4, generic: 12) against 16 for the
//   plain complex 2x2 product (measured: 0.29 / 0.16 / 0.42 vs 0.54 ms per record over 2^30
//   amplitudes).
//
// Reference analogue: none (default.qubit sweeps the state once per gate,
// simulate.py:214-235; adjoint_jacobian.py:121-137 once per gate and parameter).
// Algorithmic bytes per launch: 2*S*NV.
//
// This file is compiled by NVRTC only (no system headers).  "sk_config.inc" defines
//   SK_REAL (float|double) SK_RB SK_TB SK_NV SK_L SK_MINB SK_NROUNDS SK_NCOEF SK_NSLOTS SK_NEXT
//   SK_SWW SK_COEF_PARAM SK_RPOS {{...},...} SK_TPOS {{...},...}
//
// Coefficients.  SK_COEF_PARAM = 1: the table is a KERNEL PARAMETER (constant bank): every
// coefficient is an immediate-like operand of the FP64 instruction that uses it — no load, no
// register, no latency, and the pivot flags are uniform branches.  0 (tables larger than a few KB,
// or one table per batch element): the table is copied to shared memory once per CTA and read
// with volatile loads (the table is invariant over the tile loop and ptxas would otherwise hoist
// the coefficients of every record into registers).
#include "segk_args.h"
#include "sk_config.inc"

typedef SK_REAL real;
struct __align__(2 * sizeof(SK_REAL)) C { real x, y; };
struct __align__(64) SkTensorMap { unsigned long long opaque[16]; };
// the tile maps of vector 0 / 1: a KERNEL PARAMETER (no upload before the launch: a host-to-device
// copy on the compute stream would queue behind the exchange copies on the copy engines)
struct __align__(64) SkMaps { SkTensorMap m[2]; };

static constexpr int RB = SK_RB, TB = SK_TB, NV = SK_NV, T = SK_RB + SK_TB, L = SK_L;
static constexpr int NA = 1 << RB, THREADS = 1 << TB, NW = THREADS / 32;
static constexpr int SWW = SK_SWW;
static constexpr unsigned RSZ = sizeof(real);

__device__ constexpr int sk_rpos(int r, int b) { constexpr int t[SK_NROUNDS][SK_RB] = SK_RPOS; return t[r][b]; }
__device__ constexpr int sk_tpos(int r, int b) { constexpr int t[SK_NROUNDS][SK_TB] = SK_TPOS; return t[r][b]; }

// XOR swizzle of a tile-local index (linear over GF(2): slot(a | b) = slot(a) ^ slot(b) for
// disjoint a, b), same function as rtile.cuh
__device__ constexpr unsigned sk_sw(unsigned j) {
  unsigned s = 0;
  for (int sh = SWW; sh < 16; sh += SWW) s ^= (j >> sh);
  return s & ((1u << SWW) - 1u);
}
// tile-local index contributed by register index k / by the thread id in round r
__device__ constexpr unsigned sk_kj(int r, int k) {
  unsigned j = 0;
  for (int b = 0; b < RB; ++b) j |= ((unsigned)(k >> b) & 1u) << sk_rpos(r, b);
  return j;
}
template <int R> __device__ __forceinline__ unsigned sk_tj(const unsigned tid) {
  unsigned j = 0;
#pragma unroll
  for (int b = 0; b < TB; ++b) j |= ((tid >> b) & 1u) << sk_tpos(R, b);
  return j;
}
template <int R, int K> struct SkK {
  static constexpr unsigned j = sk_kj(R, K);
  static constexpr unsigned slot = sk_kj(R, K) ^ sk_sw(sk_kj(R, K));
};

// ---- shared-memory / TMA helpers ------------------------------------------------------------
__device__ __forceinline__ unsigned sk_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sk_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sk_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sk_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sk_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SK_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SK_DONE_%=;\n"
      "bra SK_WAIT_%=;\n"
      "SK_DONE_%=:\n"
      "}\n" ::"r"(sk_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void sk_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   sk_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(sk_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sk_tma_load(void* dst_smem, const SkTensorMap* tm, const int rank,
                                            const int (&c)[5], unsigned long long* bar) {
  const unsigned d = sk_smem_u32(dst_smem), b = sk_smem_u32(bar);
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
struct SkCoef { real v[SK_NCOEF]; };

// coefficient source (see the header): pair(i) = (table[i], table[i + 1])
#if SK_COEF_PARAM
struct Coefs {
  const SkCoef& cf;
  __device__ __forceinline__ C pair(const unsigned i) const { C r; r.x = cf.v[i]; r.y = cf.v[i + 1]; return r; }
  __device__ __forceinline__ bool flag(const unsigned i) const {
#if SK_IS_DOUBLE
    return __double2hiint(cf.v[i]) != 0;
#else
    return __float_as_int(cf.v[i]) != 0;
#endif
  }
};
#else
struct Coefs {
  unsigned base;          // shared-memory address of the table
  __device__ __forceinline__ C pair(const unsigned i) const {
    C r;
    const unsigned addr = base + i * RSZ;
#if SK_IS_DOUBLE
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
#else
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
#endif
    return r;
  }
  __device__ __forceinline__ bool flag(const unsigned i) const { return pair(i - 1).y != (real)0; }
};
#endif

__device__ __forceinline__ C sk_cmul(const C a, const C b) {
  C r;
  r.x = fma(a.x, b.x, -(a.y * b.y));
  r.y = fma(a.x, b.y, a.y * b.x);
  return r;
}
__device__ __forceinline__ double sk_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- records ----------------------------------------------------------------------------------
// Normalised single-qubit block on register bit Q (see the header).  Coefficients at `ca`:
// (t, pivot flag), then r if DR, then l if DL.  KERN 0: real kernel, 1: imaginary kernel.
template <int Q, int KERN, bool DL, bool DR, bool SINP>
__device__ __forceinline__ void sk_dk_body(C (&A)[NV][NA], const real t, const C r, const C l) {
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    if ((k >> Q) & 1) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const C x0 = A[v][k];
      C x1 = A[v][k | (1 << Q)];
      if (DR) x1 = sk_cmul(r, x1);
      C y0, y1;
      if (KERN == 0) {
        if (!SINP) {   // [[1,-t],[t,1]]
          y0.x = fma(-t, x1.x, x0.x); y0.y = fma(-t, x1.y, x0.y);
          y1.x = fma(t, x0.x, x1.x);  y1.y = fma(t, x0.y, x1.y);
        } else {       // [[t,-1],[1,t]]
          y0.x = fma(t, x0.x, -x1.x); y0.y = fma(t, x0.y, -x1.y);
          y1.x = fma(t, x1.x, x0.x);  y1.y = fma(t, x1.y, x0.y);
        }
      } else {
        if (!SINP) {   // [[1,-it],[-it,1]]
          y0.x = fma(t, x1.y, x0.x);  y0.y = fma(-t, x1.x, x0.y);
          y1.x = fma(t, x0.y, x1.x);  y1.y = fma(-t, x0.x, x1.y);
        } else {       // [[t,-i],[-i,t]]
          y0.x = fma(t, x0.x, x1.y);  y0.y = fma(t, x0.y, -x1.x);
          y1.x = fma(t, x1.x, x0.y);  y1.y = fma(t, x1.y, -x0.x);
        }
      }
      if (DL) y1 = sk_cmul(l, y1);
      A[v][k] = y0;
      A[v][k | (1 << Q)] = y1;
    }
  }
}
template <int Q, int KERN, bool DL, bool DR>
__device__ __forceinline__ void sk_dk(C (&A)[NV][NA], const Coefs& cs, const unsigned off) {
  const C tf = cs.pair(off);
  C r = {1, 0}, l = {1, 0};
  if (DR) r = cs.pair(off + 2);
  if (DL) l = cs.pair(off + (DR ? 4 : 2));
#if SK_COEF_PARAM
  const bool sinp = cs.flag(off + 1);
#else
  const bool sinp = tf.y != (real)0;
#endif
  if (!sinp) sk_dk_body<Q, KERN, DL, DR, false>(A, tf.x, r, l);
  else sk_dk_body<Q, KERN, DL, DR, true>(A, tf.x, r, l);
}

// in-place complex mat-vec on D amplitudes (every output born in the register it lives in)
template <int D>
__device__ __forceinline__ void sk_matvec(C (&A)[NA], const int (&ix)[D], const C (&m)[D * D]) {
  real tx[D], ty[D];
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
    const real d = m[r * D + r].x;
    A[ix[r]].x = fma(d, A[ix[r]].x, tx[r]);
    A[ix[r]].y = fma(d, A[ix[r]].y, ty[r]);
  }
}

// General 2x2 block on register bit Q.  Controls on register bits: (k & CR) == CV, resolved at
// compile time; `pred`: the thread / external part of the controls.  HAS0: a second matrix (at
// ca + 8 reals) acts where the controls fail (controlled-select), else those pairs are skipped.
template <int Q, unsigned CR, unsigned CV, bool HAS0>
__device__ __forceinline__ void sk_f16(C (&A)[NV][NA], const Coefs& cs, const unsigned off, const bool pred) {
  if (HAS0 || pred) {
    {
      const unsigned a1 = (HAS0 && !pred) ? off + 8 : off;
      const C m[4] = {cs.pair(a1), cs.pair(a1 + 2), cs.pair(a1 + 4), cs.pair(a1 + 6)};
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        if (((k >> Q) & 1) || ((unsigned)k & CR) != CV) continue;
        const int ix[2] = {k, k | (1 << Q)};
#pragma unroll
        for (int v = 0; v < NV; ++v) sk_matvec<2>(A[v], ix, m);
      }
    }
    if (HAS0 && CR != 0u) {
      const unsigned a0 = off + 8;
      const C m[4] = {cs.pair(a0), cs.pair(a0 + 2), cs.pair(a0 + 4), cs.pair(a0 + 6)};
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        if (((k >> Q) & 1) || ((unsigned)k & CR) == CV) continue;
        const int ix[2] = {k, k | (1 << Q)};
#pragma unroll
        for (int v = 0; v < NV; ++v) sk_matvec<2>(A[v], ix, m);
      }
    }
  }
}

// General 4x4 block on register bits Q0 (matrix MSB) > Q1, controls as sk_f16.  The 16 entries
// are re-read per quad (64 registers of amplitudes leave no room to pin them).
template <int Q0, int Q1, unsigned CR, unsigned CV>
__device__ __forceinline__ void sk_d2(C (&A)[NV][NA], const Coefs& cs, const unsigned off, const bool pred) {
  if (pred) {
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if (((k >> Q0) & 1) || ((k >> Q1) & 1) || ((unsigned)k & CR) != CV) continue;
      const int ix[4] = {k, k | (1 << Q1), k | (1 << Q0), k | (1 << Q0) | (1 << Q1)};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        C m[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = cs.pair(off + 2 * i);
        sk_matvec<4>(A[v], ix, m);
      }
    }
  }
}

// X on register bit Q.  Register-bit controls resolved at compile time: with no other control
// the exchange is a renaming of registers (no instructions); HASP: thread / external controls
// as a per-thread select.
template <int Q, unsigned CR, unsigned CV, bool HASP>
__device__ __forceinline__ void sk_cx(C (&A)[NV][NA], const bool pred) {
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    if (((k >> Q) & 1) || ((unsigned)k & CR) != CV) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const C a = A[v][k], b = A[v][k | (1 << Q)];
      if (HASP) {
        A[v][k].x = pred ? b.x : a.x; A[v][k].y = pred ? b.y : a.y;
        A[v][k | (1 << Q)].x = pred ? a.x : b.x; A[v][k | (1 << Q)].y = pred ? a.y : b.y;
      } else {
        A[v][k] = b;
        A[v][k | (1 << Q)] = a;
      }
    }
  }
}

// Parity phase: where the controls hold, amp *= parity ? m1 : m0.  PR: parity bits on register
// bits; RTPAR: the parity has a thread / external part `par_rt`; NORM: m0 == 1 was divided out
// on the host (uncontrolled phases; the quotient is part of the segment scalar) and only m1 is
// in the table.
template <unsigned CR, unsigned CV, unsigned PR, bool RTPAR, bool NORM>
__device__ __forceinline__ void sk_par(C (&A)[NV][NA], const Coefs& cs, const unsigned off, const bool cpred,
                                       const unsigned par_rt) {
  if (cpred) {
    C me = {1, 0}, mo;
    if (NORM) mo = cs.pair(off);
    else { me = cs.pair(off); mo = cs.pair(off + 2); }
    if (RTPAR) {
      const C a = me, b = mo;
      me.x = par_rt ? b.x : a.x; me.y = par_rt ? b.y : a.y;
      mo.x = par_rt ? a.x : b.x; mo.y = par_rt ? a.y : b.y;
    }
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if (((unsigned)k & CR) != CV) continue;
      const bool odd = __popc((unsigned)k & PR) & 1;
      if (NORM && !RTPAR && !odd) continue;
#pragma unroll
      for (int v = 0; v < NV; ++v) A[v][k] = sk_cmul(odd ? mo : me, A[v][k]);
    }
  }
}

// Diagonal table: amp *= tab[i0 | KI(k)], KI(k) = OR of the contributions RCb of the register
// bits set in k (compile time), i0 = the thread / external part.
template <unsigned RC0, unsigned RC1, unsigned RC2, unsigned RC3, unsigned RC4>
__device__ __forceinline__ void sk_diag(C (&A)[NV][NA], const Coefs& cs, const unsigned off, const unsigned i0) {
  constexpr unsigned rc[5] = {RC0, RC1, RC2, RC3, RC4};
  const unsigned base = off + i0 * 2u;
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    unsigned ki = 0;
#pragma unroll
    for (int b = 0; b < RB; ++b)
      if ((k >> b) & 1) ki |= rc[b];
    const C d = cs.pair(base + ki * 2u);
#pragma unroll
    for (int v = 0; v < NV; ++v) A[v][k] = sk_cmul(d, A[v][k]);
  }
}

// every amplitude times the scalar at `ca` (the product of the scalars the normalised records of
// this segment left out)
__device__ __forceinline__ void sk_scale(C (&A)[NV][NA], const Coefs& cs, const unsigned off) {
  const C s = cs.pair(off);
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int k = 0; k < NA; ++k) A[v][k] = sk_cmul(s, A[v][k]);
}

// Adjoint generator term: accs[SLOT] += coef * {Re|Im} sum_k sign_k conj(bra_k) ket_{k ^ XR},
// sign_k = (-1)^(popc((k ^ XR) & ZR) + tpar).  XR / ZR: X and Z parts of the Pauli term on
// register bits; `tpar`: parity of the thread / external Z part (+ the i^2 of two Y factors);
// ODD: an odd number of Y factors (the term is i * real Pauli: take Re instead of Im).
// The per-thread part of it: coef * sum over this thread's amplitudes.  The warp reduction is
// shared between up to four terms (sk_gen_flush*): a butterfly that halves the number of lanes
// per value at each step needs 6 double shuffles for four values (5 for two) where four separate
// reductions need 20 — ncu on the reverse sweep: SHFL 8.5 % of the issued instructions and
// mio_throttle the top stall.
template <unsigned XR, unsigned ZR, bool ODD>
__device__ __forceinline__ double sk_gen_val(const C (&A)[NV][NA], const Coefs& cs, const unsigned off,
                                             const unsigned tpar) {
  // two FMAs per amplitude into four independent chains (the sign of the term is a compile-time
  // operand negation); the first version formed the product (DMUL + DFMA) and added it (DADD):
  // 24 FP64 instructions per record and thread instead of 16
  double ac[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    const C b = A[NV - 1][k], x = A[0][k ^ XR];
    const bool neg = __popc((unsigned)(k ^ XR) & ZR) & 1;
    const double bx = neg ? -(double)b.x : (double)b.x, by = neg ? -(double)b.y : (double)b.y;
    double& a0 = ac[2 * (k & 1)];
    double& a1 = ac[2 * (k & 1) + 1];
    if (ODD) {
      a0 = fma(bx, (double)x.x, a0);
      a1 = fma(by, (double)x.y, a1);
    } else {
      a0 = fma(bx, (double)x.y, a0);
      a1 = fma(-by, (double)x.x, a1);
    }
  }
  const double acc = (ac[0] + ac[1]) + (ac[2] + ac[3]);
  const C cf = cs.pair(off);
  return acc * (tpar ? -(double)cf.x : (double)cf.x);
}
__device__ __forceinline__ double sk_shx(const double v, const int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
// warp sums of one / two / four per-thread values into accs[slot][warp]; the slots of one call
// are distinct (the host adds values of the same slot before), fixed order: deterministic
template <int S0>
__device__ __forceinline__ void sk_gen_flush1(double v0, double* accs, const unsigned tid) {
  v0 = sk_warp_sum(v0);
  if ((tid & 31u) == 0) accs[S0 * NW + (tid >> 5)] += v0;
}
template <int S0, int S1>
__device__ __forceinline__ void sk_gen_flush2(const double v0, const double v1, double* accs, const unsigned tid) {
  const bool hi = tid & 16u;
  double m = (hi ? v1 : v0) + sk_shx(hi ? v0 : v1, 16);      // lanes 0-15: v0, lanes 16-31: v1
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) m += sk_shx(m, o);
  if ((tid & 15u) == 0) accs[(hi ? S1 : S0) * NW + (tid >> 5)] += m;
}
template <int S0, int S1, int S2>
__device__ __forceinline__ void sk_gen_flush3(const double v0, const double v1, const double v2, double* accs,
                                              const unsigned tid) {
  const bool hi = tid & 16u, mid = tid & 8u;
  const double m0 = (hi ? v2 : v0) + sk_shx(hi ? v0 : v2, 16);   // lanes 0-15: (v0, v1), 16-31: (v2, -)
  const double m1 = (hi ? 0.0 : v1) + sk_shx(hi ? v1 : 0.0, 16);
  double m = (mid ? m1 : m0) + sk_shx(mid ? m0 : m1, 8);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) m += sk_shx(m, o);
  if ((tid & 7u) == 0 && !(hi && mid)) accs[(hi ? S2 : (mid ? S1 : S0)) * NW + (tid >> 5)] += m;
}
template <int S0, int S1, int S2, int S3>
__device__ __forceinline__ void sk_gen_flush4(const double v0, const double v1, const double v2, const double v3,
                                              double* accs, const unsigned tid) {
  const bool hi = tid & 16u, mid = tid & 8u;
  const double m0 = (hi ? v2 : v0) + sk_shx(hi ? v0 : v2, 16);   // lanes 0-15: (v0, v1), 16-31: (v2, v3)
  const double m1 = (hi ? v3 : v1) + sk_shx(hi ? v1 : v3, 16);
  double m = (mid ? m1 : m0) + sk_shx(mid ? m0 : m1, 8);         // 8-lane groups: v0, v1, v2, v3
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) m += sk_shx(m, o);
  if ((tid & 7u) == 0) accs[(hi ? (mid ? S3 : S2) : (mid ? S1 : S0)) * NW + (tid >> 5)] += m;
}

// ---- tile movement ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long sk_gscatter(unsigned j, const SkArgs& a) {
  unsigned long long off = j & ((1u << L) - 1u);
  const unsigned hi = j >> L;
#pragma unroll
  for (int b = 0; b < T - L; ++b) off |= (unsigned long long)((hi >> b) & 1u) << a.hi_bits[b];
  return off;
}
__device__ __forceinline__ unsigned long long sk_tile_base(const SkArgs& a, const unsigned long long t) {
  unsigned long long base = a.base_fix;
  for (int r = 0; r < a.nruns; ++r)
    base |= ((t >> a.run_s[r]) & ((1ull << a.run_len[r]) - 1ull)) << a.run_g[r];
  return base;
}

// Fetch tile `t` of every vector into the (idle) transpose buffer: one TMA box per vector issued
// by thread 0, or one bulk copy per contiguous run spread over the threads.  Every thread calls
// this once it no longer needs the buffer.
__device__ __forceinline__ void sk_fetch(const SkArgs& a, C* const (&vec)[2], C* tile, const unsigned long long t,
                                         unsigned long long* bar, const SkTensorMap* tm, const unsigned tid) {
  constexpr unsigned bytes = (unsigned)(NV * sizeof(C)) << T;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (a.tma_rank > 0) {
    __syncthreads();
    if (tid == 0) {
      const unsigned long long base = sk_tile_base(a, t);
      int c[5] = {0, 0, 0, 0, 0};
      for (int r = 1; r < a.tma_rank; ++r)
        if (a.tma_len[r]) c[r] = (int)((base >> a.tma_lo[r]) & ((1ull << a.tma_len[r]) - 1ull));
      sk_mbar_expect_tx(bar, bytes);
#pragma unroll
      for (int v = 0; v < NV; ++v) sk_tma_load(tile + ((unsigned long long)v << T), tm + v, a.tma_rank, c, bar);
    }
  } else {
    if (tid == 0) sk_mbar_expect_tx(bar, bytes);
    __syncthreads();
    const unsigned long long base = sk_tile_base(a, t);
    for (unsigned j = tid; j < (unsigned)NV << (T - L); j += THREADS) {
      const unsigned v = j >> (T - L), r = j & ((1u << (T - L)) - 1u);
      sk_bulk_g2s(tile + ((unsigned long long)v << T) + ((unsigned long long)r << L), vec[v] + base + sk_gscatter(r << L, a),
                  (unsigned)sizeof(C) << L, bar);
    }
  }
}

// registers <- landing buffer (natural order, as the TMA engine wrote it), layout of round R
template <int R, int K = 0>
__device__ __forceinline__ void sk_load(C (&A)[NV][NA], const C* tile, const unsigned tj) {
  if constexpr (K < NA) {
#pragma unroll
    for (int v = 0; v < NV; ++v) A[v][K] = tile[((unsigned)v << T) + (tj | SkK<R, K>::j)];
    sk_load<R, K + 1>(A, tile, tj);
  }
}
template <int R, int K = 0>
__device__ __forceinline__ void sk_sts(const C (&A)[NV][NA], C* tile, const unsigned tslot) {
  if constexpr (K < NA) {
#pragma unroll
    for (int v = 0; v < NV; ++v) tile[((unsigned)v << T) + (tslot ^ SkK<R, K>::slot)] = A[v][K];
    sk_sts<R, K + 1>(A, tile, tslot);
  }
}
template <int R, int K = 0>
__device__ __forceinline__ void sk_lds(C (&A)[NV][NA], const C* tile, const unsigned tslot) {
  if constexpr (K < NA) {
#pragma unroll
    for (int v = 0; v < NV; ++v) A[v][K] = tile[((unsigned)v << T) + (tslot ^ SkK<R, K>::slot)];
    sk_lds<R, K + 1>(A, tile, tslot);
  }
}
// transpose through shared memory: store in the layout of round R0, load in that of round R1
template <int R0, int R1>
__device__ __forceinline__ void sk_xpose(C (&A)[NV][NA], C* tile, const unsigned tid) {
  const unsigned t0 = sk_tj<R0>(tid), t1 = sk_tj<R1>(tid);
  __syncthreads();
  sk_sts<R0>(A, tile, t0 ^ sk_sw(t0));
  __syncthreads();
  sk_lds<R1>(A, tile, t1 ^ sk_sw(t1));
}

#define SK_LOAD(R) sk_load<R>(A, tile, sk_tj<R>(tid));
#define SK_XPOSE(R0, R1) sk_xpose<R0, R1>(A, tile, tid);
#define SK_FETCH_NEXT() if (more) sk_fetch(a, vec, tile, t + gridDim.x, bar, tm, tid);
#define SK_COEF(off) cs, (unsigned)(off)
#define SK_TBIT(b) ((tid >> (b)) & 1u)
#define SK_EBIT(e) ((ext >> (e)) & 1u)

// Register cap.  SK_MAXREG > 0: __maxnreg__ (it cannot be combined with __launch_bounds__, and
// launch bounds win over --maxrregcount): 120 registers x 512 resident threads leave 4096
// registers of the SM to the one-warp CTAs of the exchange's unpack kernel (remap.cu).
#if defined(SK_MAXREG) && SK_MAXREG > 0
#define SK_KERNEL_BOUNDS __maxnreg__(SK_MAXREG)
#else
#define SK_KERNEL_BOUNDS __launch_bounds__(1 << SK_TB, SK_MINB)
#endif
extern "C" __global__ void SK_KERNEL_BOUNDS
sk_kernel(const __grid_constant__ SkArgs a, C* __restrict__ v0, C* __restrict__ v1,
          const double* __restrict__ coef_g, const long long coef_bstride,
          const __grid_constant__ SkMaps tmaps, double* __restrict__ partials
#if SK_COEF_PARAM
          , const __grid_constant__ SkCoef cf
#endif
          ) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);                                            // NV << T
#if SK_COEF_PARAM
  unsigned long long* koff = reinterpret_cast<unsigned long long*>(tile + ((unsigned long long)NV << T));   // NA
#else
  real* coef = reinterpret_cast<real*>(tile + ((unsigned long long)NV << T));           // SK_NCOEF
  unsigned long long* koff = reinterpret_cast<unsigned long long*>(coef + SK_NCOEF);   // NA
#endif
  unsigned long long* bar = koff + NA;                                                 // 1
  double* accs = reinterpret_cast<double*>(bar + 1);                                   // NSLOTS * NW
  const unsigned tid = threadIdx.x;
  const SkTensorMap* tm = tmaps.m;

  {
#if !SK_COEF_PARAM
    const double* cg = coef_g + (long long)blockIdx.y * coef_bstride;
    for (int i = tid; i < SK_NCOEF; i += THREADS) coef[i] = (real)cg[i];
#endif
    for (int i = tid; i < SK_NSLOTS * NW; i += THREADS) accs[i] = 0.0;
    if (tid < NA) koff[tid] = sk_gscatter(sk_kj(SK_NROUNDS - 1, (int)tid), a);
    if (tid == 0) sk_mbar_init(bar, 1);
  }
  __syncthreads();
#if SK_COEF_PARAM
  const Coefs cs = {cf};
#else
  const Coefs cs = {sk_smem_u32(coef)};
#endif
  C* const vec[2] = {v0 + ((unsigned long long)blockIdx.y << a.n),
                     NV > 1 ? v1 + ((unsigned long long)blockIdx.y << a.n) : nullptr};
  const unsigned long long toff_st = sk_gscatter(sk_tj<SK_NROUNDS - 1>(tid), a);
  const unsigned koff_s = sk_smem_u32(koff);

  if (blockIdx.x < a.ntiles) sk_fetch(a, vec, tile, blockIdx.x, bar, tm, tid);
  unsigned phase = 0;
  C A[NV][NA];

  for (unsigned long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    const unsigned long long base = sk_tile_base(a, t);
    const bool more = t + gridDim.x < a.ntiles;
    unsigned ext = 0;
    {
      const unsigned long long baseE = base | a.base_hi;
#pragma unroll
      for (int e = 0; e < SK_NEXT; ++e) ext |= (unsigned)((baseE >> a.ext_pos[e]) & 1ull) << e;
    }
    sk_mbar_wait(bar, phase);
    phase ^= 1u;

#include "sk_body.inc"

    // ---- store (layout of the last round: tile positions 0..4 on the lanes) -----------------
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (v == 0 && NV > 1 && !a.write0) continue;
      C* dst = vec[v] + base + toff_st;
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        unsigned long long ko;     // volatile: 2^RB 64-bit offsets must not be hoisted out of the tile loop
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(ko) : "r"(koff_s + 8u * k));
        dst[ko] = A[v][k];
      }
    }
  }

  if (SK_NSLOTS > 0) {
    __syncthreads();
    for (int s = tid; s < SK_NSLOTS; s += THREADS) {
      double acc = 0.0;
      for (int w = 0; w < NW; ++w) acc += accs[s * NW + w];
      partials[((unsigned long long)blockIdx.y * SK_NSLOTS + s) * gridDim.x + blockIdx.x] = acc;
    }
  }
}
