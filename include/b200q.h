/* b200q — C ABI of the B200-native statevector engine behind the `b200.qubit` PennyLane device.
 *
 * The reference (PennyLane, /root/reference) is pure Python and has NO FFI of its own: the
 * functions below are what its numerical engine `pennylane/devices/qubit/*.py` would bind if its
 * numpy calls were replaced by a native library.  Each entry point cites the reference routine
 * it stands in for.  `pennylane_b200/_lib.py` is the ctypes binding; INTEGRATION.md shows the
 * stub a PennyLane maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; b200q_last_error() describes
 *     the last failure on the calling thread;
 *   - `state` is a DEVICE pointer to batch * 2^n complex amplitudes (dtype 0 = complex64,
 *     1 = complex128), batch-major; the caller (torch) owns all device memory;
 *   - qubits are addressed by BIT POSITION q of the flat index (q = 0 has stride 1).
 *     default.qubit's wire w of an n-wire state is bit n-1-w (initialize_state.py:43-44);
 *   - matrices are row-major complex128 on the HOST (`*_host`) or dtype-typed on the DEVICE
 *     (`*_dev`); matrix-index bit (k-1-j) belongs to tgt_bits[j] (PennyLane: first wire = MSB);
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises unless documented;
 *   - `work` is a caller-provided device scratch buffer of >= b200q_workspace_bytes() bytes;
 *   - reductions are deterministic (fixed summation order).
 */
#ifndef B200Q_H
#define B200Q_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q_ABI_VERSION 1
#define B200Q_DTYPE_C64 0
#define B200Q_DTYPE_C128 1
#define B200Q_CDF_EXACT 0 /* numpy-ordered float64 additions: bit-exact shots      */
#define B200Q_CDF_FAST 1  /* blocked parallel scan: fastest, ~1e-16 CDF differences */
#define B200Q_CDF_EXACT_SERIAL 2 /* the same bits as EXACT from one dependent chain of additions
                                    (the checker of the parallel exact scan)               */

const char* b200q_last_error(void);
int b200q_version(void);
int b200q_sm_count(void);
size_t b200q_workspace_bytes(void);

/* |index> for every batch element.  initialize_state.py:41-45 (create_initial_state). */
int b200q_set_basis_state(void* state, int n, int dtype, int64_t batch, uint64_t index,
                          void* stream);

/* Dense 2^k x 2^k matrix on tgt_bits with controls.  apply_operation.py:151-255
 * (apply_operation_einsum / _tensordot), :645-723 (_apply_rotation_1q), :613-642 (Hadamard),
 * :521-524 (PauliX), :763-832 (CNOT, MultiControlledX: pass X with controls).
 * k <= 3: mat_host (by value) or mat_dev; 4 <= k <= 10: mat_dev only.
 * mat_bstride (complex elements) != 0 selects a different matrix per batch element
 * (parameter broadcasting, apply_operation.py:191-197); mat_dev only. */
int b200q_apply_matrix(void* state, int n, int dtype, int64_t batch, const int* tgt_bits, int k,
                       const int* ctrl_bits, const int* ctrl_vals, int nc, const void* mat_host,
                       const void* mat_dev, int64_t mat_bstride, void* stream);

/* General diagonal gate, 2^k entries.  apply_operation.py:528-609 and every op whose matrix is
 * diagonal.  k <= 6: diag_host or diag_dev; k <= 20: diag_dev. */
int b200q_apply_diag(void* state, int n, int dtype, int64_t batch, const int* bits, int k,
                     const void* diag_host, const void* diag_dev, int64_t diag_bstride,
                     void* stream);

/* Multiply the subspace selected by (ctrl_bits == ctrl_vals) by one complex scalar; nc = 0 is
 * GlobalPhase / rescale.  apply_operation.py:507-517, :528-609 (Z, PhaseShift, T, S), CZ, CCZ,
 * ControlledPhaseShift.  phase_dev (nullable): per-batch scalars of the state's dtype. */
int b200q_apply_phase(void* state, int n, int dtype, int64_t batch, const int* ctrl_bits,
                      const int* ctrl_vals, int nc, double phase_re, double phase_im,
                      const void* phase_dev, void* stream);

/* One block of a reduced density matrix, pennylane/math/quantum.py:386-487 (`reduce_statevector`,
 * the einsum behind qml.density_matrix / purity / vn_entropy / mutual_info):
 *   G[ai, bi] = sum_r psi[A, ai, r] * conj(psi[B, bi, r])
 * over all bits r that are neither inner nor outer.  inner_bits (mi <= 2, matrix MSB first) index
 * the block; outer_bits (mo, MSB first) are fixed to row_assign (A) and col_assign (B).
 * out_dev: 2 * 4^mi doubles, (re, im) interleaved, row-major.  One unbatched state per call.
 * Algorithmic bytes: S / 2^mo when A == B, else 2 S / 2^mo. */
int b200q_gram_block(const void* state, int n, int dtype, const int* inner_bits, int mi,
                     const int* outer_bits, int mo, uint64_t row_assign, uint64_t col_assign,
                     double* out_dev, void* work, size_t work_bytes, void* stream);

/* Mid-circuit measurement collapse, apply_operation.py:478-495 (projector, `state / norm`,
 * optional reset) as one sweep: amplitudes with (bit == sample) are multiplied by `scale`
 * (the host passes 1 / sqrt(p_sample), p from b200q_probs on that bit) and, when reset != 0 and
 * sample == 1, moved to the bit == 0 half; the other half is zeroed.  Unbatched states only
 * (the reference raises for batched ones, apply_operation.py:441-442).
 * Algorithmic bytes: S/2 read + S written. */
int b200q_collapse(void* state, int n, int dtype, int bit, int sample, int reset, double scale,
                   void* stream);

/* amp *= popcount(i & mask) odd ? p1 : p0.  RZ / IsingZZ / MultiRZ / PauliRot("Z..Z") of any
 * width (ops/qubit/parametric_ops_multi_qubit.py:93).  phases_dev (nullable): per-batch
 * (p0, p1) pairs of the state's dtype. */
int b200q_apply_parity_phase(void* state, int n, int dtype, int64_t batch, uint64_t mask,
                             double p0_re, double p0_im, double p1_re, double p1_im,
                             const void* phases_dev, void* stream);

/* exp(-i theta/2 P), P a Pauli word with >= 1 X/Y factor: xmask = X|Y bits, zmask = Z|Y bits,
 * ny = number of Y.  c = cos(theta/2), s = sin(theta/2).  parametric_ops_multi_qubit.py:380-436.
 * cs_dev (nullable): per-batch (c, s) packed as one complex of the state's dtype. */
int b200q_apply_pauli_rot(void* state, int n, int dtype, int64_t batch, uint64_t xmask,
                          uint64_t zmask, int ny, double c, double s, const void* cs_dev,
                          void* stream);

/* Marginal probabilities over bits[0..m) (bin MSB = bits[0]) -> out_dev[batch][2^m] float64.
 * measurements/probs.py:101-135 (ProbabilityMP.process_state). */
int b200q_probs(const void* state, int n, int dtype, int64_t batch, const int* bits, int m,
                double* out_dev, void* work, size_t work_bytes, void* stream);

/* <psi| sum_t coeff_t P_t |psi> -> out_dev[batch].  Terms are HOST arrays.
 * measure.py:74-99 (csr_dot_products, Pauli branch) + pauli_arithmetic.py:924-950. */
int b200q_expval_pauli_sum(const void* state, int n, int dtype, int64_t batch,
                           const uint64_t* xmasks, const uint64_t* zmasks, const int* nys,
                           const double* coeffs, int nterms, double* out_dev, void* work,
                           size_t work_bytes, void* stream);

/* <a|b> per batch element -> out_dev[0..batch) real parts, out_dev[batch..2*batch) imaginary.
 * measure.py:121-139 (full_dot_products), adjoint_jacobian.py:36-40. */
int b200q_inner(const void* a, const void* b, int n, int dtype, int64_t batch, double* out_dev,
                void* work, size_t work_bytes, void* stream);

/* out = scale * sum_t (cre_t + i cim_t) P_t |in>  (out != in).  adjoint_jacobian.py:113-115
 * (bras = 2 * obs|ket>), :315-317 (pauli_rep.dot). */
int b200q_pauli_sum_apply(const void* in, void* out, int n, int dtype, int64_t batch,
                          const uint64_t* xmasks, const uint64_t* zmasks, const int* nys,
                          const double* cre, const double* cim, int nterms, double scale,
                          void* work, size_t work_bytes, void* stream);

/* <bra|P|ket> for one Pauli word -> out_dev[0] = Re, out_dev[1] = Im. */
int b200q_pauli_braket(const void* bra, const void* ket, int n, int dtype, uint64_t xmask,
                       uint64_t zmask, int ny, double* out_dev, void* work, size_t work_bytes,
                       void* stream);

/* Shot sampler.  probs_dev: 2^m float64 probabilities, OVERWRITTEN with the normalised CDF.
 * uniforms_dev: `shots` float64 in [0,1) drawn by the host Generator.  Writes basis-state
 * indices to idx_out_dev (nullable, int64[shots]) and bit rows to bits_out_dev (nullable,
 * int64[shots][m], column 0 = most significant bit).  norm_out_dev[0] receives the numpy-order
 * sum of the probabilities (host checks |norm-1| <= 1e-6, sampling.py:514-519);
 * flags_dev[0] is set to 1 if any probability is NaN (sampling.py:322-325).
 * sampling.py:500-531 (_sample_probs_numpy) == Generator.choice(arange(2^m), shots, p). */
int b200q_sample(double* probs_dev, int m, const double* uniforms_dev, int64_t shots, int mode,
                 int64_t* idx_out_dev, int64_t* bits_out_dev, double* norm_out_dev,
                 int* flags_dev, void* work, size_t work_bytes, void* stream);

/* Sampler building blocks: b200q_sample == has_nan, np_sum, div_by, cumsum, div_by(last), search,
 * unpack_bits.  A sharded probability vector (rank r holds entries [r 2^m, (r+1) 2^m)) composes
 * them around its collectives: np_sum per rank + numpy's pairwise tree over ranks; cumsum with
 * carry_in_dev = the previous rank's last CDF entry (NULL on rank 0); search returns the number
 * of LOCAL cdf entries <= u, and the global index is the sum of those counts over ranks.
 * All vectors have 2^m entries.  sampling.py:500-531. */
int b200q_has_nan(const double* p_dev, int m, int* flag_dev, void* stream);
int b200q_np_sum(const double* p_dev, int m, double* out_dev, void* work, size_t work_bytes,
                 void* stream);
int b200q_div_by(double* p_dev, int m, const double* divisor_dev, void* stream);
int b200q_cumsum(double* p_dev, int m, int mode, const double* carry_in_dev, void* work,
                 size_t work_bytes, void* stream);
int b200q_search(const double* cdf_dev, int m, const double* uniforms_dev, int64_t shots,
                 int64_t* idx_out_dev, void* stream);
int b200q_unpack_bits(const int64_t* idx_dev, int64_t shots, int m, int64_t* bits_out_dev,
                      void* stream);

/* Fused gate segment: ONE read + ONE write of the state applies `nops` gates through a
 * shared-memory tile spanning `tile_bits` (T ascending bit positions, the first L equal to
 * 0..L-1).  ops_host: array of 64-byte b200q_tile_op records (layout below); mats_host: the
 * segment's complex128 matrix / phase table.  Replaces `nops` iterations of the gate loop
 * simulate.py:214-235 (the reference has no fusion).
 *
 *   struct b200q_tile_op {
 *     int32  kind;        // 0 dense 2x2, 1 dense 4x4, 2 controlled-X, 3 parity phase,
 *                         // 4 diagonal table (<= 4 bits), 5 swap
 *     int32  t0, t1;      // tile-local target bits (t0 = matrix MSB)
 *     int32  mat_off;     // offset into mats_host (complex entries)
 *     uint32 ctrl_mask_l, ctrl_val_l;   // controls on tile-local bits
 *     uint32 par_mask_l;  // parity phase: tile-local bits
 *     int32  ndiag;       // diagonal table: number of index bits
 *     uint64 ctrl_mask_e, ctrl_val_e;   // controls on bits outside the tile (global positions)
 *     uint64 par_mask_e;  // parity phase: bits outside the tile
 *     int8   dbits[8];    // diagonal table index bits, MSB first: >= 0 tile-local, < 0 -(global)-1
 *   };
 */
int b200q_apply_tile(void* state, int n, int dtype, int64_t batch, const int* tile_bits, int T,
                     int L, const void* ops_host, int nops, const void* mats_host, int nmat,
                     void* work, size_t work_bytes, void* stream);

/* out_dev[batch] = Re <psi| H |psi> for a CSR matrix H (2^k x 2^k, complex128 values, int64
 * indptr / indices, all on the device) acting on the k state bits tbits[0..k) (tbits[0] = most
 * significant bit of the matrix index).  Replaces measure.py:74-118 (csr_dot_products, scipy
 * branch: SparseHamiltonian) without expanding H to the full register. */
int b200q_expval_csr(const void* state, int n, int dtype, int64_t batch, const int* tbits, int k,
                     const int64_t* indptr_dev, const int64_t* indices_dev, const void* data_dev,
                     double* out_dev, void* work, size_t work_bytes, void* stream);

/* Register-tiled fused segment (the production fused path; b200q_apply_tile is the small-state
 * fallback).  ONE read + ONE write of vec0 (and of vec1 when given) applies `nops` records:
 * every thread keeps 2^RB amplitudes in registers, a segment is a sequence of ROUNDS (which tile
 * positions are register bits), gates on register bits cost no memory traffic, rounds are
 * separated by one swizzled shared-memory transpose.  (T, RB, threads) are fixed per
 * (dtype, nvec): query them with b200q_rtile_geometry.  ops_host: 64-byte records
 * (pennylane_b200/csrc/rtile.cuh `RtOp`, mirrored by pennylane_b200/compiler.py `RtOp`); the
 * first record must be a ROUND.  base_hi is OR-ed into the tile base index when evaluating
 * controls / parities on bits outside the tile: a sharded state passes its rank bits there.
 * vec1 != NULL selects adjoint mode: gates act on both vectors, GEN records accumulate
 * coef * Im<vec1| P |vec0> into slot q0; out_dev[batch][nslots] = scale * sums (deterministic
 * order); write0 = 0 leaves vec0 untouched (extra bras re-use the ket).
 * Replaces simulate.py:214-235 (gate loop) and adjoint_jacobian.py:121-137 (reverse sweep). */
int b200q_rtile_geometry(int dtype, int nvec, int* T_out, int* RB_out, int* threads_out);
int b200q_apply_rtile(void* vec0, void* vec1, int n, int dtype, int64_t batch, const int* tile_bits,
                      int T, int L, const void* ops_host, int nops, const void* mats_host, int nmat,
                      int nslots, int write0, uint64_t base_hi, double scale, double* out_dev,
                      void* work, size_t work_bytes, void* stream);
/* Same, for a broadcast state (simulate.py:235, apply_operation.py:186-197): mats_host holds
 * `batch` consecutive tables of nmat entries, batch element b reads table b (gates whose
 * parameters carry a leading batch axis; tables of unbatched records are simply repeated). */
int b200q_apply_rtile_bcast(void* vec0, void* vec1, int n, int dtype, int64_t batch,
                            const int* tile_bits, int T, int L, const void* ops_host, int nops,
                            const void* mats_host, int nmat, int nslots, int write0,
                            uint64_t base_hi, double scale, double* out_dev, void* work,
                            size_t work_bytes, void* stream);

/* Structure-specialised fused segment kernel (pennylane_b200/csrc/segk.cuh): the same tile
 * pipeline as b200q_apply_rtile, but the segment's STRUCTURE (round layouts, record kinds,
 * register bits, control locations) is compiled into the kernel and only the VALUES are launch
 * arguments.  The host (pennylane_b200/segjit.py) emits two small headers per structure
 * ("sk_config.inc", "sk_body.inc"); b200q_jit_compile turns segk.cuh + headers into an sm_100a
 * cubin through NVRTC (dlopen'ed libnvrtc.so.12; needs no GPU), b200q_seg_load makes it
 * launchable, b200q_seg_launch runs one segment:
 *   tile_bits[T] as for b200q_apply_rtile; RB / minb: register bits and CTAs per SM the kernel
 *   was compiled for; ext_pos[n_ext]: global positions of the bits outside the tile that the
 *   kernel's predicates read (base_hi is OR-ed into the tile base first: rank bits of a sharded
 *   state); coef_host[n_coef] doubles: matrix coefficients in the order the body consumes them;
 *   coef_mode 0: one table, staged in shared memory; 1: `batch` tables, batch element b reads
 *   table b (broadcast parameters); 2: the table is passed as a kernel parameter (kernels
 *   compiled with SK_COEF_PARAM: coefficients become constant-bank operands); vec1 / nslots /
 *   write0 / scale / out_dev: adjoint mode, as for b200q_apply_rtile;
 *   fix_mask / fix_val: PARTIAL launch — the non-tile index bits of fix_mask are held at
 *   fix_val and only those 2^(n-T-popc(fix_mask)) tiles are processed (0, 0: the whole state).
 *   The sharded engine runs the segments on either side of an exchange piece by piece so that
 *   the NVLink transfer of piece p overlaps the sweeps of the other pieces.
 * Replaces simulate.py:214-235 (gate loop) and adjoint_jacobian.py:121-137 (reverse sweep). */
int b200q_jit_available(void);
int b200q_jit_compile(const char* source, const char* const* header_names,
                      const char* const* header_sources, int n_headers, int lineinfo, int maxreg,
                      void** cubin_out, size_t* size_out);
void b200q_jit_free(void* cubin);
int b200q_seg_load(const void* cubin, size_t size, void** handle_out);
int b200q_seg_unload(void* handle);
int b200q_seg_launch(void* handle, void* vec0, void* vec1, int n, int dtype, int64_t batch,
                     const int* tile_bits, int T, int L, int RB, int minb, const int* ext_pos,
                     int n_ext, const double* coef_host, int n_coef, int coef_mode, int nslots,
                     int write0, uint64_t base_hi, uint64_t fix_mask, uint64_t fix_val, double scale,
                     double* out_dev, void* work, size_t work_bytes, void* stream);

/* Qubit-remapping exchange, data movement (K9 of SURVEY.md section 2c: the remap pack / unpack
 * pair).  One PIECE of an exchanged slab is `count` runs of `run_bytes` bytes, `src_pitch` /
 * `dst_pitch` bytes apart (pack: state -> dense staging buffer, dst_pitch == run_bytes; unpack:
 * a partner's staging buffer, mapped over NVLink, -> state, src_pitch == run_bytes).  Issued on
 * the copy engines (cudaMemcpy2DAsync; a loop of contiguous copies when the pitch exceeds the
 * 2 GiB pitch limit), so it runs beside the segment kernels, which occupy every SM.
 * The reference has no analogue (default.qubit is one array). */
int b200q_remap_copy(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                     size_t run_bytes, size_t count, void* stream);
/* The same copy as a KERNEL that shares the SMs with a running segment launch: the local half
 * of an exchange step (staging buffer -> state).  Beside a fused segment launch a copy-engine
 * device-to-device copy gets 0.4 TB/s and a pitched one waits for the launch boundary.
 *   mode B200Q_UNPACK_TMA : one thread per CTA keeps four 16 KiB cp.async.bulk copies in flight
 *                           (one warp, 64 KiB of shared memory: fits beside two resident segment
 *                           CTAs); 6.3 TB/s of traffic alone, 2 TB/s beside a two-round segment,
 *                           but starved by segments with many shared-memory transpositions;
 *   mode B200Q_UNPACK_REGS: through registers, no shared memory: ~1.1 TB/s beside any segment.
 * `ctas`: 0 = default grid.  Runs that are not multiples of 16 KiB (TMA) / 64 B fall back to
 * b200q_remap_copy. */
#define B200Q_UNPACK_TMA 0
#define B200Q_UNPACK_REGS 1
int b200q_remap_unpack(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                       size_t run_bytes, size_t count, int mode, int ctas, void* stream);
/* Flags of the exchange protocol as STREAM MEMORY OPERATIONS (cuStreamWriteValue32 /
 * cuStreamWaitValue32 with CU_STREAM_WAIT_VALUE_GEQ): `addr` is a 4-byte aligned device address,
 * local or a peer's (mapped over NVLink).  No kernel is launched: the exchange makes progress
 * while the segment kernels hold every SM. */
int b200q_stream_write32(void* addr, uint32_t value, void* stream);
int b200q_stream_wait_geq32(void* addr, uint32_t value, void* stream);

/* One reverse-sweep step of adjoint differentiation on vecs = [1 + n_bras][2^n] (row 0 = ket):
 *   z_b = <bra_b| G |ket>,  ket <- A ket,  bra_b <- A bra_b      (A = U^dagger, k <= 3)
 * out_dev[b] = -Im z_b  (= Re <bra_b| i G |ket>, the Jacobian entry when bras carry the factor
 * 2 of adjoint_jacobian.py:115).  gen_host == NULL: non-trainable op, only applies A.
 * adjoint_jacobian.py:121-137. */
int b200q_adjoint_step(void* vecs, int n, int dtype, int n_bras, const int* tgt_bits, int k,
                       const int* ctrl_bits, const int* ctrl_vals, int nc, const void* adj_host,
                       const void* gen_host, double* out_dev, void* work, size_t work_bytes,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H */
