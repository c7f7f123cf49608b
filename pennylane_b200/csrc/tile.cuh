// b200q — fused tile kernel (K5 of SURVEY.md section 2c; the device half of the host-side
// gate-fusion pass north_star asks for).
//
// One launch = ONE read and ONE write of the state, applying a whole run ("segment") of gates.
// A CTA stages a tile of 2^T amplitudes in shared memory.  The tile covers T index bits: the
// lowest L bits (contiguous in memory -> every global access is a run of 2^L amplitudes:
// 1 KiB at L = 6 in complex128) plus T-L arbitrary higher bits chosen by the host for this
// segment.  Every gate of the segment acts on tile bits only (other bits may appear as
// controls or in diagonal phases — those depend on the tile's base index, not on data), so the
// gates are applied with shared-memory passes and no further HBM traffic.
//
// Reference analogue: none (default.qubit applies one gate per full-state pass,
// simulate.py:214-235); algorithmic bytes per launch are 2*S regardless of the gate count.
#pragma once
#include "common.cuh"

namespace b200q {

enum TileOpKind : int {
  TILE_DENSE1 = 0,   // 2x2 on local bit t0, optional controls
  TILE_DENSE2 = 1,   // 4x4 on local bits (t0 = matrix MSB, t1 = LSB), optional controls
  TILE_CX = 2,       // X on local bit t0 with controls (pure swap, no flops)
  TILE_PARITY = 3,   // amp *= parity(local & ml, base & me) ? p1 : p0   (mat[0], mat[1])
  TILE_DIAG = 4,     // amp *= tab[idx], idx from up to 4 bits (local or external), table at mat
  TILE_SWAP = 5,     // swap local bits t0, t1 (controls allowed)
};

struct __align__(16) TileOp {
  int kind;
  int t0, t1;
  int mat_off;                 // offset (in complex entries) into the segment's matrix table
  unsigned ctrl_mask_l, ctrl_val_l;      // controls on tile-local bits
  unsigned par_mask_l;                    // TILE_PARITY: local bits in the parity
  int ndiag;                              // TILE_DIAG: number of index bits (<= 4)
  unsigned long long ctrl_mask_e, ctrl_val_e;   // controls on bits outside the tile (global positions)
  unsigned long long par_mask_e;          // TILE_PARITY: external bits in the parity
  signed char dbits[8];                   // TILE_DIAG: >= 0 local bit, < 0: -(global bit)-1, MSB first
};

struct TileArgs {
  int n, T, L, nops, nmat;
  int8_t hi_bits[24];          // global positions of tile bits L..T-1 (ascending)
  int8_t out_bits[B200Q_MAX_BITS];   // ascending global positions of the n-T non-tile bits
  unsigned long long ntiles;   // 2^(n-T)
};

__device__ __forceinline__ unsigned long long tile_scatter_hi(unsigned j_hi, const TileArgs& a) {
  unsigned long long off = 0;
  for (int b = 0; b < a.T - a.L; ++b) off |= (unsigned long long)((j_hi >> b) & 1u) << a.hi_bits[b];
  return off;
}

// insert a zero bit at position p (tile-local indices, 32-bit)
__device__ __forceinline__ unsigned ins0(unsigned g, int p) {
  return ((g >> p) << (p + 1)) | (g & ((1u << p) - 1u));
}

template <typename T_, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_tile(cx<T_>* __restrict__ state, const TileArgs a, const TileOp* __restrict__ ops_g,
       const double2* __restrict__ mats_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T_>* tile = reinterpret_cast<cx<T_>*>(smem_raw);                       // 2^T amplitudes
  cx<T_>* mats = tile + (1u << a.T);                                          // nmat entries
  TileOp* ops = reinterpret_cast<TileOp*>(mats + ((a.nmat + 1) & ~1));        // nops
  const unsigned tsize = 1u << a.T;
  const unsigned lowmask = (1u << a.L) - 1u;

  for (int i = threadIdx.x; i < a.nmat; i += THREADS)
    mats[i] = make_cx<T_>((T_)mats_g[i].x, (T_)mats_g[i].y);
  for (int i = threadIdx.x; i < a.nops; i += THREADS) ops[i] = ops_g[i];

  cx<T_>* st = state + ((unsigned long long)blockIdx.y << a.n);

  for (unsigned long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    // base index of this tile: deposit t into the non-tile bit positions
    unsigned long long base = 0;
    for (int b = 0; b < a.n - a.T; ++b) base |= ((t >> b) & 1ull) << a.out_bits[b];
    __syncthreads();                       // previous tile fully stored / tables loaded
    // ---- stage in -------------------------------------------------------------------
    for (unsigned j = threadIdx.x; j < tsize; j += THREADS) {
      const unsigned long long g = base | (unsigned long long)(j & lowmask) |
                                   tile_scatter_hi(j >> a.L, a);
      tile[j] = st[g];
    }
    __syncthreads();
    // ---- gates ------------------------------------------------------------------------
    for (int o = 0; o < a.nops; ++o) {
      const TileOp op = ops[o];
      // controls on bits outside the tile: uniform per tile
      if ((base & op.ctrl_mask_e) != op.ctrl_val_e) continue;     // uniform branch: no barrier skipped
      if (op.kind == TILE_DENSE1) {
        const cx<T_> m00 = mats[op.mat_off], m01 = mats[op.mat_off + 1],
                     m10 = mats[op.mat_off + 2], m11 = mats[op.mat_off + 3];
        const unsigned bit = 1u << op.t0;
        // lanes whose bit 2 is set visit their pair in the opposite order, which removes the
        // 2-way bank conflict of 16-byte accesses when t0 < 3
        const bool flip = (op.t0 < 3) && (threadIdx.x & 4);
        for (unsigned g = threadIdx.x; g < (tsize >> 1); g += THREADS) {
          const unsigned j0 = ins0(g, op.t0);
          if ((j0 & op.ctrl_mask_l) != op.ctrl_val_l) continue;
          const unsigned j1 = j0 | bit;
          cx<T_> x0, x1;
          if (flip) { x1 = tile[j1]; x0 = tile[j0]; } else { x0 = tile[j0]; x1 = tile[j1]; }
          cx<T_> y0 = make_cx<T_>(0, 0), y1 = make_cx<T_>(0, 0);
          cmac(y0, m00, x0); cmac(y0, m01, x1);
          cmac(y1, m10, x0); cmac(y1, m11, x1);
          if (flip) { tile[j1] = y1; tile[j0] = y0; } else { tile[j0] = y0; tile[j1] = y1; }
        }
      } else if (op.kind == TILE_DENSE2) {
        const cx<T_>* m = mats + op.mat_off;
        const int lo = op.t0 < op.t1 ? op.t0 : op.t1, hi = op.t0 < op.t1 ? op.t1 : op.t0;
        const unsigned b0 = 1u << op.t0, b1 = 1u << op.t1;       // b0: matrix MSB
        for (unsigned g = threadIdx.x; g < (tsize >> 2); g += THREADS) {
          const unsigned j = ins0(ins0(g, lo), hi);
          if ((j & op.ctrl_mask_l) != op.ctrl_val_l) continue;
          const unsigned idx[4] = {j, j | b1, j | b0, j | b0 | b1};
          cx<T_> x[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) x[r] = tile[idx[r]];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            cx<T_> y = make_cx<T_>(0, 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) cmac(y, m[r * 4 + c], x[c]);
            tile[idx[r]] = y;
          }
        }
      } else if (op.kind == TILE_CX) {
        const unsigned bit = 1u << op.t0;
        const bool flip = (op.t0 < 3) && (threadIdx.x & 4);
        for (unsigned g = threadIdx.x; g < (tsize >> 1); g += THREADS) {
          const unsigned j0 = ins0(g, op.t0);
          if ((j0 & op.ctrl_mask_l) != op.ctrl_val_l) continue;
          const unsigned j1 = j0 | bit;
          cx<T_> x0, x1;
          if (flip) { x1 = tile[j1]; x0 = tile[j0]; tile[j1] = x0; tile[j0] = x1; }
          else { x0 = tile[j0]; x1 = tile[j1]; tile[j0] = x1; tile[j1] = x0; }
        }
      } else if (op.kind == TILE_SWAP) {
        const int lo = op.t0 < op.t1 ? op.t0 : op.t1, hi = op.t0 < op.t1 ? op.t1 : op.t0;
        const unsigned b0 = 1u << op.t0, b1 = 1u << op.t1;
        for (unsigned g = threadIdx.x; g < (tsize >> 2); g += THREADS) {
          const unsigned j = ins0(ins0(g, lo), hi);
          if ((j & op.ctrl_mask_l) != op.ctrl_val_l) continue;
          const cx<T_> u = tile[j | b0], v = tile[j | b1];
          tile[j | b0] = v; tile[j | b1] = u;
        }
      } else if (op.kind == TILE_PARITY) {
        const cx<T_> p0 = mats[op.mat_off], p1 = mats[op.mat_off + 1];
        const unsigned pe = __popcll(base & op.par_mask_e) & 1u;
        for (unsigned j = threadIdx.x; j < tsize; j += THREADS) {
          if ((j & op.ctrl_mask_l) != op.ctrl_val_l) continue;
          const unsigned par = (__popc(j & op.par_mask_l) & 1u) ^ pe;
          tile[j] = cmul(par ? p1 : p0, tile[j]);
        }
      } else {  // TILE_DIAG
        const cx<T_>* tab = mats + op.mat_off;
        unsigned ext = 0;       // contribution of external bits to the table index
        for (int b = 0; b < op.ndiag; ++b) {
          const int d = op.dbits[b];
          if (d < 0) ext |= (unsigned)((base >> (-d - 1)) & 1ull) << (op.ndiag - 1 - b);
        }
        for (unsigned j = threadIdx.x; j < tsize; j += THREADS) {
          if ((j & op.ctrl_mask_l) != op.ctrl_val_l) continue;
          unsigned idx = ext;
          for (int b = 0; b < op.ndiag; ++b) {
            const int d = op.dbits[b];
            if (d >= 0) idx |= ((j >> d) & 1u) << (op.ndiag - 1 - b);
          }
          tile[j] = cmul(tab[idx], tile[j]);
        }
      }
      __syncthreads();
    }
    // ---- stage out ----------------------------------------------------------------------
    for (unsigned j = threadIdx.x; j < tsize; j += THREADS) {
      const unsigned long long g = base | (unsigned long long)(j & lowmask) |
                                   tile_scatter_hi(j >> a.L, a);
      st[g] = tile[j];
    }
  }
}

}  // namespace b200q
