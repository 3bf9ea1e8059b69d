/*
 * morfem_b200 -- C ABI of the B200 (sm_100a) reduced-order frequency-sweep path.
 *
 * The reference (SzymonKnopp/morfem) has no FFI layer: its boundary is the Python function API of
 * implementation.py / test_helpers.py.  Each entry point below replaces the third-party CPU call that the
 * reference reaches at the cited line; morfem_b200/implementation.py keeps the reference's Python signatures
 * and calls these through ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; complex128 is interleaved (re, im);
 *   - dense matrices are row-major with an explicit leading dimension in elements;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *   - return value: 0 ok, <0 = -(index of the offending argument, 1-based), >0 = cudaError_t of a failed
 *     launch / API call; mf_last_error() returns a thread-local description;
 *   - no call allocates or frees device memory: workspaces are sized by the *_ws_bytes queries and owned by
 *     the caller; nothing here synchronises the device.
 */
#ifndef MORFEM_B200_H
#define MORFEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } mf_c128;

#define MF_VERSION 100
#define MF_MAX_PORTS 16

int mf_version(void);
const char* mf_last_error(void);
/* Number of kernels launched by this library in the calling process so far (bench.py "gpu_launches"). */
int64_t mf_launch_count(void);

/* ---- stage 1 + 2 dense contractions ------------------------------------------------------------------
 * C (ra x rb) = op(A)^T B summed over n rows; A is n x ra, B is n x rb; conj_a != 0 selects A^H (Gram /
 * north-star Q^H), 0 selects the plain transpose of implementation.py:180 (`q_t = q.T`).
 * Replaces BLAS dgemm behind `(q_t @ a) @ q` (implementation.py:181-183) and the Gram products of the
 * CholeskyQR2 replacement of np.linalg.svd (implementation.py:226, :298, :210).
 * Split over row ranges; partials go to `ws` and are summed in a fixed order (deterministic). */
size_t mf_gemm_tn_ws_bytes(int ra, int rb, int64_t n);
int mf_gemm_tn_c128(const mf_c128* A, int64_t lda, int ra, const mf_c128* B, int64_t ldb, int rb, int64_t n,
                    int conj_a, mf_c128* C, int64_t ldc, void* ws, size_t ws_bytes, void* stream);

/* Out (n x rb) = A (n x ra) * W (ra x rb).  The `S R^-1` / `Q1 (R2^-1 U)` applications of CholeskyQR2 and
 * the lift Q x.  Out must not alias A. */
int mf_gemm_nn_c128(const mf_c128* A, int64_t lda, int64_t n, int ra, const mf_c128* W, int64_t ldw, int rb,
                    mf_c128* Out, int64_t ldo, void* stream);
/* Out (n x r) = A (n x r) * W with W (r x r) UPPER TRIANGULAR -- the `X R^-1` application of a Cholesky-QR pass
 * (stage 1, implementation.py:226).  Entries of W below the diagonal are taken as zero and whole 64-column tiles of
 * zero products are skipped (5/8 of the work at r = 256); results equal mf_gemm_nn on a W whose lower part is zero. */
int mf_trmm_nn_c128(const mf_c128* A, int64_t lda, int64_t n, int r, const mf_c128* W, int64_t ldw,
                    mf_c128* Out, int64_t ldo, void* stream);

/* ---- r x r factorisations used by the basis stage (all single-launch, device-resident) ---------------
 * mf_equilibrate:   d[j] = 1/sqrt(real(G[j][j])) (1 if the diagonal is not positive); G <- diag(d) G diag(d)
 *                   + shift * I.  stats[0] = max_ij |G_ij - delta_ij| of the INPUT (departure of the block whose
 *                   Gram matrix G is from orthonormality; the CholeskyQR pass counter tests it), stats may be NULL.
 * mf_potrf_upper:   Cholesky G = R^H R, R upper triangular written over G (strict lower part zeroed);
 *                   *info = 0 or the 1-based column at which a non-positive pivot appeared.
 * mf_trtri_upper:   Rinv = R^-1 (upper triangular), Rinv must not alias R.
 * mf_scale_cols / mf_scale_rows: X[:, j] *= d[j]^p / X[i, :] *= d[i]^p with p = +1 or -1.
 * mf_jacobi_svd:    one-sided Jacobi SVD of the r x r matrix X: on exit U holds the left singular vectors
 *                   (columns, sorted by descending sigma), sigma the singular values; X and work are
 *                   destroyed.  Replaces LAPACK gesdd on the small factor (np.linalg.svd of the snapshot
 *                   block = CholeskyQR2 + SVD of R).  Cooperative launch: needs r/2 <= co-resident CTAs. */
int mf_equilibrate_c128(mf_c128* G, int64_t ld, int r, double shift, double* d, double* stats, void* stream);
int mf_potrf_upper_c128(mf_c128* G, int64_t ld, int r, int* info, void* stream);
int mf_trtri_upper_c128(const mf_c128* R, int64_t ldr, int r, mf_c128* Rinv, int64_t ldi, void* stream);
/* Fused Cholesky + triangular inverse (one cooperative launch, 32 x 32 blocks, device-wide barriers between the block
 * steps): G = R^H R in place (R upper, zeros below) and Rinv = R^-1.  info as mf_potrf_upper_c128.  What the
 * Cholesky-QR passes use above r = 112 (np.linalg.svd replacement, implementation.py:226).  ws: mf_chol_inv_ws_bytes. */
size_t mf_chol_inv_ws_bytes(int r);
int mf_chol_inv_upper_c128(mf_c128* G, int64_t ldg, int r, mf_c128* Rinv, int64_t ldi, int* info,
                           void* ws, size_t ws_bytes, void* stream);
int mf_scale_cols_c128(mf_c128* X, int64_t ld, int rows, int cols, const double* d, int power, void* stream);
int mf_scale_rows_c128(mf_c128* X, int64_t ld, int rows, int cols, const double* d, int power, void* stream);
size_t mf_jacobi_svd_ws_bytes(int r);
int mf_jacobi_svd_c128(mf_c128* X, int64_t ld, int r, mf_c128* U, int64_t ldu, double* sigma, int max_sweeps,
                       double tol, int* sweeps_done, void* ws, size_t ws_bytes, void* stream);

/* ---- stage 2 sparse ------------------------------------------------------------------------------------
 * Y (nrows x r) = A Q with A in CSR (int32 indices; values real f64 when val_is_real != 0 else c128).
 * The reference's `q_t @ a` (implementation.py:181-183) is scipy's csr_matvecs over the CSR view of a^T, i.e.
 * the CSC arrays of `a` passed here unchanged give Y = a^T Q.  Row-sharded callers pass their slice of the
 * row pointer array rebased to start at 0 together with the matching slices of colidx/vals; column indices
 * stay global (they index rows of Q). */
int mf_spmm_csr_c128(const int32_t* rowptr, const int32_t* colidx, const void* vals, int val_is_real,
                     int64_t nrows, const mf_c128* Q, int64_t ldq, int r, mf_c128* Y, int64_t ldy, void* stream);
/* Two operators that share ONE sparsity pattern (the reference's Ct and Tt: stiffness and mass matrix of the same mesh,
 * projected back to back at implementation.py:181-183): Y0 = A0 Q and Y1 = A1 Q in one pass over the pattern -- every Q row is
 * loaded once for both.  vals0 / vals1 follow the same val_is_real. */
int mf_spmm_csr2_c128(const int32_t* rowptr, const int32_t* colidx, const void* vals0, const void* vals1, int val_is_real,
                      int64_t nrows, const mf_c128* Q, int64_t ldq, int r, mf_c128* Y0, int64_t ldy0, mf_c128* Y1, int64_t ldy1,
                      void* stream);
int mf_spmm_csr2_f64(const int32_t* rowptr, const int32_t* colidx, const double* vals0, const double* vals1, int64_t nrows,
                     const double* Q, int64_t ldq, int r, double* Y0, int64_t ldy0, double* Y1, int64_t ldy1, void* stream);

/* Row-grouped form of the same product for real operators with sorted column indices (the FEM operators of this path):
 * G = mf_spmm_group_size(r) consecutive rows are handled by one warp over the UNION of their columns, so that each
 * needed Q row is loaded once per group instead of once per non-zero (the CSR kernel is bound by L1 traffic, not HBM).
 * Build once per operator: counts[g] = union size of group g (mf_spmm_group_count); ustart = exclusive prefix sum of
 * counts (int64, ngroups + 1 entries, caller computed); mf_spmm_group_fill writes ucols[ustart[g] + k] and
 * uvals[(ustart[g] + k) * G + i] (coefficient of row g*G + i, 0 when absent).  Then mf_spmm_grouped_c128 == mf_spmm_csr_c128. */
int mf_spmm_group_size(int r);       /* complex128 Q: 4 rows up to r = 128, 2 above */
int mf_spmm_group_size_f64(int r);   /* float64 Q: 4 rows up to r = 256, 2 above */
int mf_spmm_group_count(const int32_t* rowptr, const int32_t* colidx, int64_t nrows, int G, int32_t* counts, void* stream);
int mf_spmm_group_fill(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t nrows, int G,
                       const int64_t* ustart, int32_t* ucols, double* uvals, void* stream);
int mf_spmm_grouped_c128(const int64_t* ustart, const int32_t* ucols, const double* uvals, int64_t nrows, int G,
                         const mf_c128* Q, int64_t ldq, int r, mf_c128* Y, int64_t ldy, void* stream);

/* Windowed SpMM with TMA staging (the north star's "Q tiles staged in shared memory by TMA"), same product as mf_spmm_csr_*
 * for REAL operator values.  One CTA per block of mf_spmm_window_rows_per_block() consecutive rows; the distinct Q rows the
 * block references (its window: ucol[wstart[b] .. wstart[b+1]), at most wmax <= mf_spmm_window_max_rows() of them) are fetched
 * slice by slice with bulk asynchronous copies (cp.async.bulk + mbarrier) and reused from shared memory by all rows of the
 * block; slot[k] is the position of non-zero k's column inside its block's window (8 bits instead of a 32-bit column index).
 * Rows may hold at most mf_spmm_window_max_nnz_per_row() non-zeros.  The window lists are built once per operator on the host
 * (morfem_b200.device.build_windows).  Kept as the measured alternative to the row-grouped kernel (profiles/r02_spmm.md). */
int mf_spmm_window_rows_per_block(void);
int mf_spmm_window_max_nnz_per_row(void);
int mf_spmm_window_max_rows(void);
int mf_spmm_window_c128(const int32_t* rowptr, const uint8_t* slot, const double* vals, const int32_t* wstart, const int32_t* ucol,
                        int64_t nrows, int wmax, const mf_c128* Q, int64_t ldq, int r, mf_c128* Y, int64_t ldy, void* stream);
int mf_spmm_window_f64(const int32_t* rowptr, const uint8_t* slot, const double* vals, const int32_t* wstart, const int32_t* ucol,
                       int64_t nrows, int wmax, const double* Q, int64_t ldq, int r, double* Y, int64_t ldy, void* stream);

/* B_r (r x m) = Q^T B (conj_q != 0: Q^H B) for B in CSC (n x m, int32 indices, real or complex values);
 * implementation.py:184 `q_t @ md.b`.  Rows outside [row0, row0 + nlocal) are skipped so that row-sharded
 * callers can all-reduce the partial results; Q points at the caller's first local row. */
int mf_project_rhs_c128(const int32_t* colptr, const int32_t* rowidx, const void* vals, int val_is_real, int m,
                        const mf_c128* Q, int64_t ldq, int r, int64_t row0, int64_t nlocal, int conj_q,
                        mf_c128* Br, int64_t ldb, void* stream);

/* ---- real float64 twins ----------------------------------------------------------------
 * The reference's data and its whole ROM path are real float64 (main.py:21-23, implementation.py:190).  For real
 * snapshots and operators these entries run stages 1 + 2 on 8-byte elements: half the bytes and a quarter of the flops
 * of the complex128 entries, with results that are bit-identical to them (every imaginary part would be an exact zero).
 * Same argument meaning as the _c128 entries above; there is no conjugation flag. */
size_t mf_gemm_tn_f64_ws_bytes(int ra, int rb, int64_t n);
int mf_gemm_tn_f64(const double* A, int64_t lda, int ra, const double* B, int64_t ldb, int rb, int64_t n,
                   double* C, int64_t ldc, void* ws, size_t ws_bytes, void* stream);
int mf_gemm_nn_f64(const double* A, int64_t lda, int64_t n, int ra, const double* W, int64_t ldw, int rb,
                   double* Out, int64_t ldo, void* stream);
int mf_trmm_nn_f64(const double* A, int64_t lda, int64_t n, int r, const double* W, int64_t ldw,
                   double* Out, int64_t ldo, void* stream);
int mf_spmm_csr_f64(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t nrows,
                    const double* Q, int64_t ldq, int r, double* Y, int64_t ldy, void* stream);
int mf_spmm_grouped_f64(const int64_t* ustart, const int32_t* ucols, const double* uvals, int64_t nrows, int G,
                        const double* Q, int64_t ldq, int r, double* Y, int64_t ldy, void* stream);
/* Batched reduced sweep on REAL reduced operators (same semantics as mf_sweep_lu_gsm_c128; X is real, S complex) -- the
 * reference's own arithmetic (implementation.py:190 allocates a float64 result; lu_factor / lu_solve on float64, :477-478).
 * variant 0 = auto (matrix resident in shared memory up to r = 128, the left-looking streamed LU of sweep_left.cu up to
 * r = 512), 3 / 5 force one of the two.  mf_sweep_f64_supported(r, m): any variant handles (r, m); workspace from
 * mf_sweep_f64_ws_bytes (the shared-memory kernel needs none). */
int mf_sweep_f64_supported(int r, int m);
int mf_sweep_f64_variant_supported(int r, int m, int variant);
size_t mf_sweep_f64_ws_bytes(int r, int m, int64_t F, int variant);
int mf_sweep_lu_gsm_f64(const double* A0, const double* A1, const double* A2, int64_t lda, const double* Br, int64_t ldb, int r, int m,
                        const double* c0, const double* c1, const double* c2, const double* cb, const double* zscale, int64_t F,
                        double* X, mf_c128* S, int* info, int variant, void* ws, size_t ws_bytes, void* stream);
/* One-sided Jacobi SVD of a REAL r x r matrix (r <= 64, see mf_jacobi_svd_f64_supported); X is not modified. */
int mf_jacobi_svd_f64_supported(int r);
int mf_jacobi_svd_f64(const double* X, int64_t ld, int r, double* U, int64_t ldu, double* sigma, int max_sweeps,
                      double tol, int* sweeps_done, void* stream);
int mf_project_rhs_f64(const int32_t* colptr, const int32_t* rowidx, const double* vals, int m,
                       const double* Q, int64_t ldq, int r, int64_t row0, int64_t nlocal, double* Br, int64_t ldb, void* stream);

/* As (r x r) = (A + A^T) / 2: the symmetrisation of implementation.py:528 hoisted out of the sweep loop
 * (linear in the operators, so applying it once to each reduced operator is the same mathematics). */
int mf_symmetrize_c128(const mf_c128* A, int64_t lda, int r, mf_c128* As, int64_t lds, void* stream);

/* ---- stages 3 + 4: the batched reduced sweep -------------------------------------------------------------
 * For every point i < F:
 *     A_i = c0[i] A0 + c1[i] A1 + c2[i] A2          (operators already symmetrised; NULL operator = zero)
 *     X_i = A_i^-1 (cb[i] Br)                        LU with partial pivoting (LAPACK getrf/getrs semantics,
 *                                                    implementation.py:477-478)
 *     Z_i = j zscale[i] X_i^T (cb[i] Br),  Y_i = Z_i^-1,  S_i = 2 (I + Y_i)^-1 - I     (test_helpers.py:9-14,
 *                                                    zscale[i] = 2 pi f_i eps0)
 * Outputs: S (F x m x m) when non-NULL, X (F x r x m) when non-NULL, info[i] = 0 or the 1-based column of the
 * first exactly-zero pivot (LAPACK convention).  `variant`: 0 = auto, 1 = generic (one CTA per point),
 * 2 = register-panel kernel (r <= 128), 3 = blocked DMMA kernel.  ws from mf_sweep_ws_bytes. */
size_t mf_sweep_ws_bytes(int r, int m, int64_t F, int variant);
/* 1 when `variant` (1..3) can run this (r, m), else 0; variant 0 (auto) is always supported. */
int mf_sweep_variant_supported(int r, int m, int variant);
int mf_sweep_lu_gsm_c128(const mf_c128* A0, const mf_c128* A1, const mf_c128* A2, int64_t lda,
                         const mf_c128* Br, int64_t ldb, int r, int m,
                         const double* c0, const double* c1, const double* c2, const double* cb,
                         const double* zscale, int64_t F,
                         mf_c128* X, mf_c128* S, int* info, int variant,
                         void* ws, size_t ws_bytes, void* stream);

/* Stage 4 on its own (test_helpers.py:9-14 applied to F points): S_i from given solutions X (F x r x m) and
 * port matrix Bmat (r x m): Z_i = j zscale[i] X_i^T (cb[i] Bmat), S_i = 2 (I + Z_i^-1)^-1 - I. */
int mf_gsm_c128(const mf_c128* X, const mf_c128* Bmat, int64_t ldb, int r, int m, const double* cb,
                const double* zscale, int64_t F, mf_c128* S, void* stream);

/* ---- "next" row N1: residual error estimator of the greedy basis search (implementation.py:348-452) --------
 * For every point i, with x_i = X[i] (r x m) from the sweep:
 *     E_i = sum_{a,b} c_a c_b x^H G_ab x - cb sum_a c_a x^H H_a - cb sum_a c_a H_a^H x + cb^2 BB      (m x m)
 *     err[i] = || E_i ||_F                                                     (implementation.py:424-441)
 * G_ab = (A_a Q)^H (A_b Q) (r x r), H_a = (A_a Q)^H B (r x m), BB = B^H B (m x m), all row-major, tightly packed.
 * G_host is a HOST array of 9 device pointers (index 3a + b), H_host a HOST array of 3; NULL entries are zero
 * blocks.  The reference forms the same blocks as sparse products h(a_i) @ a_j (implementation.py:370-402). */
int mf_estimator_c128(const mf_c128* X, int r, int m, int64_t F, const mf_c128* const* G_host,
                      const mf_c128* const* H_host, const mf_c128* BB,
                      const double* c0, const double* c1, const double* c2, const double* cb,
                      double* err, void* stream);

/* ---- measurement helper (bench.py, SURVEY 8d: "FP64 peak must be measured on the box") ---------------------
 * FP64 tensor-pipe (DMMA m8n8k4) issue rate of the current device in TFLOP/s, best of four timed launches of an
 * issue-loop kernel (8 independent accumulators per warp, `iters` rounds).  The one entry point that BLOCKS and
 * allocates (a few KB, freed before return): it times itself with CUDA events on `stream`. */
int mf_peak_dmma_tflops(int iters, double* tflops_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MORFEM_B200_H */
