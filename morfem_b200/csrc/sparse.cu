// Stage-2 sparse kernels: CSR SpMM  Y = A Q  (the reference's `q_t @ a`, implementation.py:181-183, which scipy
// evaluates as csr_matvecs over the CSR view of a^T -- i.e. over the CSC arrays of `a`), and the projection of
// the sparse port matrix  B_r = Q^T B  (implementation.py:184).
//
// SpMM layout: one warp per matrix row.  The 32 lanes first fetch up to 32 (column, value) pairs of the row in
// one coalesced load, then walk them with warp shuffles, so the dependent index->address chain is off the
// critical path; for every non-zero the warp streams the whole Q row (r * 16 B, contiguous) with 128-bit loads,
// CPL independent loads per lane in flight, and accumulates in registers.  Y rows are written once, coalesced.
// HBM-bound: algorithmic bytes nnz*(idx+val) + 4(N+1) + 2*N*r*16.
#include "common.cuh"

namespace {

__device__ __forceinline__ cplx ldg_q(const cplx* p) {
    // read-only path, keep in L1/L2: neighbouring rows reuse the same Q rows
    return __ldg(reinterpret_cast<const double2*>(p));
}

template <int CPL, bool REAL>
__global__ void __launch_bounds__(256)
spmm_csr_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const void* __restrict__ vals_v,
                long long nrows, const cplx* __restrict__ Q, long long ldq, int r, cplx* __restrict__ Y, long long ldy) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int s = rowptr[row], e = rowptr[row + 1];
    cplx acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = cmake(0.0, 0.0);
    for (int base = s; base < e; base += 32) {
        const int cnt = min(32, e - base);
        int my_col = 0; cplx my_val = cmake(0.0, 0.0);
        if (lane < cnt) {
            my_col = colidx[base + lane];
            if (REAL) my_val = cmake(reinterpret_cast<const double*>(vals_v)[base + lane], 0.0);
            else my_val = reinterpret_cast<const cplx*>(vals_v)[base + lane];
        }
        for (int k = 0; k < cnt; ++k) {
            const int col = __shfl_sync(0xffffffffu, my_col, k);
            cplx v;
            v.x = __shfl_sync(0xffffffffu, my_val.x, k);
            if (!REAL) v.y = __shfl_sync(0xffffffffu, my_val.y, k); else v.y = 0.0;
            const cplx* q = Q + (long long)col * ldq;
            cplx qv[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) { int idx = lane + 32 * c; qv[c] = idx < r ? ldg_q(q + idx) : cmake(0.0, 0.0); }
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (REAL) { acc[c].x = fma(v.x, qv[c].x, acc[c].x); acc[c].y = fma(v.x, qv[c].y, acc[c].y); }
                else cfma(acc[c], v, qv[c]);
            }
        }
    }
    cplx* y = Y + row * ldy;
#pragma unroll
    for (int c = 0; c < CPL; ++c) { int idx = lane + 32 * c; if (idx < r) y[idx] = acc[c]; }
}

template <bool REAL>
int spmm_dispatch(const int* rowptr, const int* colidx, const void* vals, long long nrows, const cplx* Q, long long ldq, int r,
                  cplx* Y, long long ldy, cudaStream_t st) {
    const int threads = 256;
    const long long warps = nrows;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return -5;
#define SPMM_LAUNCH(C) spmm_csr_kernel<C, REAL><<<(unsigned)blocks, threads, 0, st>>>(rowptr, colidx, vals, nrows, Q, ldq, r, Y, ldy)
    if (r <= 32) SPMM_LAUNCH(1);
    else if (r <= 64) SPMM_LAUNCH(2);
    else if (r <= 128) SPMM_LAUNCH(4);
    else if (r <= 256) SPMM_LAUNCH(8);
    else if (r <= 512) SPMM_LAUNCH(16);
    else return -9;
#undef SPMM_LAUNCH
    return 0;
}

// B_r[:, col] = sum over the non-zeros (row, v) of column `col` of B with row0 <= row < row0 + nlocal of
// v * op(Q[row - row0, :]).  One CTA per column; threads over the basis index.
template <bool REAL>
__global__ void project_rhs_kernel(const int* __restrict__ colptr, const int* __restrict__ rowidx, const void* __restrict__ vals_v,
                                   const cplx* __restrict__ Q, long long ldq, int r, long long row0, long long nlocal, int conj_q,
                                   cplx* __restrict__ Br, long long ldb) {
    const int col = blockIdx.x;
    const int s = colptr[col], e = colptr[col + 1];
    for (int i = threadIdx.x; i < r; i += blockDim.x) {
        cplx acc = cmake(0.0, 0.0);
        for (int p = s; p < e; ++p) {
            long long row = rowidx[p];
            if (row < row0 || row >= row0 + nlocal) continue;
            cplx v = REAL ? cmake(reinterpret_cast<const double*>(vals_v)[p], 0.0) : reinterpret_cast<const cplx*>(vals_v)[p];
            cplx q = Q[(row - row0) * ldq + i];
            if (conj_q) q.y = -q.y;
            cfma(acc, q, v);
        }
        Br[i * ldb + col] = acc;
    }
}

}  // namespace

extern "C" int mf_spmm_csr_c128(const int32_t* rowptr, const int32_t* colidx, const void* vals, int val_is_real,
                                int64_t nrows, const mf_c128* Q, int64_t ldq, int r, mf_c128* Y, int64_t ldy, void* stream) {
    if (!rowptr) MF_FAIL_ARG(1, "rowptr is NULL");
    if (!colidx) MF_FAIL_ARG(2, "colidx is NULL");
    if (!vals) MF_FAIL_ARG(3, "vals is NULL");
    if (nrows < 0) MF_FAIL_ARG(5, "nrows < 0");
    if (!Q || ldq < r) MF_FAIL_ARG(6, "Q is NULL or ldq < r");
    if (r <= 0 || r > 512) MF_FAIL_ARG(8, "need 0 < r <= 512");
    if (!Y || ldy < r) MF_FAIL_ARG(9, "Y is NULL or ldy < r");
    if (nrows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = val_is_real ? spmm_dispatch<true>(rowptr, colidx, vals, nrows, (const cplx*)Q, ldq, r, (cplx*)Y, ldy, st)
                         : spmm_dispatch<false>(rowptr, colidx, vals, nrows, (const cplx*)Q, ldq, r, (cplx*)Y, ldy, st);
    if (rc != 0) MF_FAIL_ARG(-rc, "size out of range for one launch");
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_project_rhs_c128(const int32_t* colptr, const int32_t* rowidx, const void* vals, int val_is_real, int m,
                                   const mf_c128* Q, int64_t ldq, int r, int64_t row0, int64_t nlocal, int conj_q,
                                   mf_c128* Br, int64_t ldb, void* stream) {
    if (!colptr) MF_FAIL_ARG(1, "colptr is NULL");
    if (!rowidx) MF_FAIL_ARG(2, "rowidx is NULL");
    if (!vals) MF_FAIL_ARG(3, "vals is NULL");
    if (m <= 0) MF_FAIL_ARG(5, "m <= 0");
    if (!Q || ldq < r) MF_FAIL_ARG(6, "Q is NULL or ldq < r");
    if (r <= 0) MF_FAIL_ARG(8, "r <= 0");
    if (!Br || ldb < m) MF_FAIL_ARG(12, "Br is NULL or ldb < m");
    cudaStream_t st = (cudaStream_t)stream;
    if (val_is_real) project_rhs_kernel<true><<<m, 256, 0, st>>>(colptr, rowidx, vals, (const cplx*)Q, ldq, r, row0, nlocal, conj_q, (cplx*)Br, ldb);
    else project_rhs_kernel<false><<<m, 256, 0, st>>>(colptr, rowidx, vals, (const cplx*)Q, ldq, r, row0, nlocal, conj_q, (cplx*)Br, ldb);
    MF_CHECK_LAUNCH();
    return 0;
}
