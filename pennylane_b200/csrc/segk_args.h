// b200q — launch arguments of the structure-specialised segment kernel (segk.cuh).  Shared by
// the NVRTC-compiled kernel and the host launcher (segk_host.cu); built-in types only, so that
// NVRTC needs no system headers.
#pragma once

struct SkArgs {
  int n;                          // qubits of the (local) state
  int write0;                     // write vector 0 back (adjoint passes over several bras)
  int tma_rank;                   // > 0: the tile is ONE box of a rank-`tma_rank` tensor map
  int nruns;                      // runs of consecutive non-tile bits (tile number -> base)
  signed char tma_lo[5], tma_len[5];   // dims 1..rank-1: non-tile group -> coordinate = bits
                                       // [lo, lo+len) of the tile base; tile group -> len = 0
  signed char hi_bits[16];        // global positions of tile positions L..T-1 (ascending)
  signed char run_s[16], run_len[16], run_g[16];   // tile-number bits [s, s+len) -> global bits [g, g+len)
  signed char ext_pos[16];        // global positions of the external predicate bits
  unsigned long long ntiles;      // 2^(n-T)
  unsigned long long base_hi;     // OR-ed into the tile base for external predicates (rank bits)
  unsigned long long base_fix;    // partial launch: non-tile bits held fixed (OR-ed into every tile base);
                                  // ntiles / runs then enumerate the remaining non-tile bits only
};
