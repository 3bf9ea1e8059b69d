// Stage-2 sparse kernels: CSR SpMM  Y = A Q  (the reference's `q_t @ a`, implementation.py:181-183, which scipy
// evaluates as csr_matvecs over the CSR view of a^T -- i.e. over the CSC arrays of `a`), and the projection of
// the sparse port matrix  B_r = Q^T B  (implementation.py:184).
//
// SpMM layout: one warp per matrix row.  The 32 lanes first fetch up to 32 (column, value) pairs of the row in
// one coalesced load, then walk them with warp shuffles, so the dependent index->address chain is off the
// critical path; for every non-zero the warp streams the whole Q row (r * 16 B, contiguous) with 128-bit loads,
// CPL independent loads per lane in flight, and accumulates in registers.  Y rows are written once, coalesced.
// HBM-bound: algorithmic bytes nnz*(idx+val) + 4(N+1) + 2*N*r*16.
#include <stdlib.h>
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

// Two operators with ONE sparsity pattern (stiffness and mass matrix of the same mesh -- the reference's Ct and Tt): every
// Q row is pulled through L1 once and feeds both accumulator sets, which halves the register-fill traffic that bounds the
// single-operator kernel.  Y0 = A0 Q, Y1 = A1 Q.
template <int CPL, bool REAL>
__global__ void __launch_bounds__(256)
spmm_csr2_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const void* __restrict__ vals0_v,
                 const void* __restrict__ vals1_v, long long nrows, const cplx* __restrict__ Q, long long ldq, int r,
                 cplx* __restrict__ Y0, long long ldy0, cplx* __restrict__ Y1, long long ldy1) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int s = rowptr[row], e = rowptr[row + 1];
    cplx acc0[CPL], acc1[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { acc0[c] = cmake(0.0, 0.0); acc1[c] = cmake(0.0, 0.0); }
    for (int base = s; base < e; base += 32) {
        const int cnt = min(32, e - base);
        int my_col = 0; cplx my_v0 = cmake(0.0, 0.0), my_v1 = cmake(0.0, 0.0);
        if (lane < cnt) {
            my_col = colidx[base + lane];
            if (REAL) {
                my_v0 = cmake(reinterpret_cast<const double*>(vals0_v)[base + lane], 0.0);
                my_v1 = cmake(reinterpret_cast<const double*>(vals1_v)[base + lane], 0.0);
            } else {
                my_v0 = reinterpret_cast<const cplx*>(vals0_v)[base + lane];
                my_v1 = reinterpret_cast<const cplx*>(vals1_v)[base + lane];
            }
        }
#pragma unroll 2
        for (int k = 0; k < cnt; ++k) {
            const int col = __shfl_sync(0xffffffffu, my_col, k);
            cplx v0, v1;
            v0.x = __shfl_sync(0xffffffffu, my_v0.x, k); v1.x = __shfl_sync(0xffffffffu, my_v1.x, k);
            if (!REAL) { v0.y = __shfl_sync(0xffffffffu, my_v0.y, k); v1.y = __shfl_sync(0xffffffffu, my_v1.y, k); }
            else { v0.y = 0.0; v1.y = 0.0; }
            const cplx* q = Q + (long long)col * ldq;
            cplx qv[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; qv[c] = idx < r ? ldg_q(q + idx) : cmake(0.0, 0.0); }
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (REAL) {
                    acc0[c].x = fma(v0.x, qv[c].x, acc0[c].x); acc0[c].y = fma(v0.x, qv[c].y, acc0[c].y);
                    acc1[c].x = fma(v1.x, qv[c].x, acc1[c].x); acc1[c].y = fma(v1.x, qv[c].y, acc1[c].y);
                } else { cfma(acc0[c], v0, qv[c]); cfma(acc1[c], v1, qv[c]); }
            }
        }
    }
    cplx* y0 = Y0 + row * ldy0; cplx* y1 = Y1 + row * ldy1;
#pragma unroll
    for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; if (idx < r) { y0[idx] = acc0[c]; y1[idx] = acc1[c]; } }
}

// real float64 twin of the two-operator kernel
template <int CPL>
__global__ void __launch_bounds__(256)
spmm_csr2_f64_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals0,
                     const double* __restrict__ vals1, long long nrows, const double* __restrict__ Q, long long ldq, int r,
                     double* __restrict__ Y0, long long ldy0, double* __restrict__ Y1, long long ldy1) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int s = rowptr[row], e = rowptr[row + 1];
    double acc0[CPL], acc1[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { acc0[c] = 0.0; acc1[c] = 0.0; }
    for (int base = s; base < e; base += 32) {
        const int cnt = min(32, e - base);
        int my_col = 0; double my_v0 = 0.0, my_v1 = 0.0;
        if (lane < cnt) { my_col = colidx[base + lane]; my_v0 = vals0[base + lane]; my_v1 = vals1[base + lane]; }
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const int col = __shfl_sync(0xffffffffu, my_col, k);
            const double v0 = __shfl_sync(0xffffffffu, my_v0, k), v1 = __shfl_sync(0xffffffffu, my_v1, k);
            const double* q = Q + (long long)col * ldq;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int idx = lane + 32 * c;
                if (idx < r) { const double qv = __ldg(q + idx); acc0[c] = fma(v0, qv, acc0[c]); acc1[c] = fma(v1, qv, acc1[c]); }
            }
        }
    }
    double* y0 = Y0 + row * ldy0; double* y1 = Y1 + row * ldy1;
#pragma unroll
    for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; if (idx < r) { y0[idx] = acc0[c]; y1[idx] = acc1[c]; } }
}

// ------------------------------------------------------------------------------------------------------------------
// Row-grouped SpMM.  ncu shows the CSR kernel bound by the L1/TEX data path, not by HBM: every non-zero pulls a whole Q
// row (16 r bytes) into registers, 24 times per matrix row for the FEM stencils of this path.  Neighbouring rows of a
// bandwidth-reduced FEM operator share most of their columns, so G consecutive rows are processed by ONE warp over the
// UNION of their columns: each needed Q row is loaded once per group (2.1x fewer loads at G = 4 for the 27-point
// stencil) and feeds up to G accumulator sets; zero coefficients are skipped with warp-uniform branches.
// Format (built on the device once per operator by the two kernels below): group g owns union columns
// ustart[g] .. ustart[g+1]-1; ucols[k] is the column, uvals[k*G + i] the coefficient of row g*G + i (0 if absent).
// Requires sorted column indices within each row and real values.

template <int G>
__device__ __forceinline__ int group_merge(const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
                                           long long nrows, long long g, long long out0, int* __restrict__ ucols, double* __restrict__ uvals) {
    int p[G], e[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const long long row = g * G + i;
        p[i] = row < nrows ? rowptr[row] : 0;
        e[i] = row < nrows ? rowptr[row + 1] : 0;
    }
    int n = 0;
    while (true) {
        int c = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < G; ++i) if (p[i] < e[i]) c = min(c, colidx[p[i]]);
        if (c == 0x7fffffff) break;
#pragma unroll
        for (int i = 0; i < G; ++i) {
            double v = 0.0;
            if (p[i] < e[i] && colidx[p[i]] == c) { if (uvals) v = vals[p[i]]; ++p[i]; }
            if (uvals) uvals[(out0 + n) * G + i] = v;
        }
        if (ucols) ucols[out0 + n] = c;
        ++n;
    }
    return n;
}

template <int G>
__global__ void spmm_group_count_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, long long nrows, long long ngroups,
                                        int* __restrict__ counts) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ngroups) counts[g] = group_merge<G>(rowptr, colidx, nullptr, nrows, g, 0, nullptr, nullptr);
}

template <int G>
__global__ void spmm_group_fill_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
                                       long long nrows, long long ngroups, const long long* __restrict__ ustart, int* __restrict__ ucols,
                                       double* __restrict__ uvals) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ngroups) group_merge<G>(rowptr, colidx, vals, nrows, g, ustart[g], ucols, uvals);
}

template <int G, int CPL>
__global__ void __launch_bounds__(256)
spmm_grouped_kernel(const long long* __restrict__ ustart, const int* __restrict__ ucols, const double* __restrict__ uvals, long long nrows,
                    const cplx* __restrict__ Q, long long ldq, int r, cplx* __restrict__ Y, long long ldy) {
    const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    // blockIdx.y selects a slice of 32 CPL columns (wide bases: more rows per group fit the accumulator registers)
    Q += (long long)blockIdx.y * 32 * CPL; Y += (long long)blockIdx.y * 32 * CPL; r -= (int)blockIdx.y * 32 * CPL;
    __shared__ __align__(16) double cstage[8][32 * G];
    double* cw = cstage[(threadIdx.x >> 5) & 7];
    if (g * G >= nrows) return;
    const long long s = ustart[g], e = ustart[g + 1];
    cplx acc[G][CPL];
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[i][c] = cmake(0.0, 0.0);
    // columns and coefficients of 32 union entries are fetched coalesced, then broadcast from registers (column, by
    // shuffle) and from a per-warp shared staging area (coefficients): no dependent global load inside the inner loop
    for (long long base = s; base < e; base += 32) {
        const int cnt = (int)min((long long)32, e - base);
        int my_col = 0;
        if (lane < cnt) {
            my_col = ucols[base + lane];
            const double2* src = reinterpret_cast<const double2*>(uvals + (base + lane) * G);
            double2* dst = reinterpret_cast<double2*>(cw + lane * G);
            dst[0] = src[0];
            if (G == 4) dst[1] = src[1];
        }
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const int col = __shfl_sync(0xffffffffu, my_col, k);
            const cplx* q = Q + (long long)col * ldq;
            cplx qv[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; qv[c] = idx < r ? ldg_q(q + idx) : cmake(0.0, 0.0); }
            double v[G];
            {
                const double2* vp = reinterpret_cast<const double2*>(cw + k * G);      // broadcast shared-memory read
                const double2 t0 = vp[0];
                v[0] = t0.x; v[1] = t0.y;
                if (G == 4) { const double2 t1 = vp[1]; v[G > 2 ? 2 : 0] = t1.x; v[G > 3 ? 3 : 0] = t1.y; }
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
                if (v[i] != 0.0) {                               // warp-uniform
#pragma unroll
                    for (int c = 0; c < CPL; ++c) { acc[i][c].x = fma(v[i], qv[c].x, acc[i][c].x); acc[i][c].y = fma(v[i], qv[c].y, acc[i][c].y); }
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const long long row = g * G + i;
        if (row < nrows) {
            cplx* y = Y + row * ldy;
#pragma unroll
            for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; if (idx < r) y[idx] = acc[i][c]; }
        }
    }
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


// ------------------------------------------------------------------------------------------------------------------
// Real float64 twins (real operator values, real Q): same access pattern, 8-byte elements.
template <int CPL>
__global__ void __launch_bounds__(256)
spmm_csr_f64_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
                    long long nrows, const double* __restrict__ Q, long long ldq, int r, double* __restrict__ Y, long long ldy) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int s = rowptr[row], e = rowptr[row + 1];
    double acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
    for (int base = s; base < e; base += 32) {
        const int cnt = min(32, e - base);
        int my_col = 0; double my_val = 0.0;
        if (lane < cnt) { my_col = colidx[base + lane]; my_val = vals[base + lane]; }
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const int col = __shfl_sync(0xffffffffu, my_col, k);
            const double v = __shfl_sync(0xffffffffu, my_val, k);
            const double* q = Q + (long long)col * ldq;
#pragma unroll
            for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; if (idx < r) acc[c] = fma(v, __ldg(q + idx), acc[c]); }
        }
    }
    double* y = Y + row * ldy;
#pragma unroll
    for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; if (idx < r) y[idx] = acc[c]; }
}

template <int G, int CPL>
__global__ void __launch_bounds__(256)
spmm_grouped_f64_kernel(const long long* __restrict__ ustart, const int* __restrict__ ucols, const double* __restrict__ uvals, long long nrows,
                        const double* __restrict__ Q, long long ldq, int r, double* __restrict__ Y, long long ldy) {
    const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    Q += (long long)blockIdx.y * 32 * CPL; Y += (long long)blockIdx.y * 32 * CPL; r -= (int)blockIdx.y * 32 * CPL;    // column slice
    __shared__ __align__(16) double cstage[8][32 * G];
    double* cw = cstage[(threadIdx.x >> 5) & 7];
    if (g * G >= nrows) return;
    const long long s = ustart[g], e = ustart[g + 1];
    double acc[G][CPL];
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[i][c] = 0.0;
    for (long long base = s; base < e; base += 32) {
        const int cnt = (int)min((long long)32, e - base);
        int my_col = 0;
        if (lane < cnt) {
            my_col = ucols[base + lane];
            const double2* src = reinterpret_cast<const double2*>(uvals + (base + lane) * G);
            double2* dst = reinterpret_cast<double2*>(cw + lane * G);
            dst[0] = src[0];
            if (G == 4) dst[1] = src[1];
        }
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const int col = __shfl_sync(0xffffffffu, my_col, k);
            const double* q = Q + (long long)col * ldq;
            double qv[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; qv[c] = idx < r ? __ldg(q + idx) : 0.0; }
            double v[G];
            {
                const double2* vp = reinterpret_cast<const double2*>(cw + k * G);
                const double2 t0 = vp[0];
                v[0] = t0.x; v[1] = t0.y;
                if (G == 4) { const double2 t1 = vp[1]; v[G > 2 ? 2 : 0] = t1.x; v[G > 3 ? 3 : 0] = t1.y; }
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
                if (v[i] != 0.0) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[i][c] = fma(v[i], qv[c], acc[i][c]);
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const long long row = g * G + i;
        if (row < nrows) {
            double* y = Y + row * ldy;
#pragma unroll
            for (int c = 0; c < CPL; ++c) { const int idx = lane + 32 * c; if (idx < r) y[idx] = acc[i][c]; }
        }
    }
}

__global__ void project_rhs_f64_kernel(const int* __restrict__ colptr, const int* __restrict__ rowidx, const double* __restrict__ vals,
                                       const double* __restrict__ Q, long long ldq, int r, long long row0, long long nlocal,
                                       double* __restrict__ Br, long long ldb) {
    const int col = blockIdx.x;
    const int s = colptr[col], e = colptr[col + 1];
    for (int i = threadIdx.x; i < r; i += blockDim.x) {
        double acc = 0.0;
        for (int p = s; p < e; ++p) {
            const long long row = rowidx[p];
            if (row < row0 || row >= row0 + nlocal) continue;
            acc = fma(Q[(row - row0) * ldq + i], vals[p], acc);
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

// Rows per group (measured on B200, N = 1M, r = 256, profiles/r02_spmm.md): complex128 -- four rows up to r = 128, two above
// (four rows in column slices of 128 read every index and coefficient list twice and lost: 7.6 ms against 3.5 ms); float64 --
// four rows up to r = 256 (the accumulators of four rows of 256 real columns fit: 2.67 ms against 3.05 ms for two rows).
extern "C" int mf_spmm_group_size(int r) { return r <= 128 ? 4 : 2; }
extern "C" int mf_spmm_group_size_f64(int r) { return r <= 256 ? 4 : 2; }

extern "C" int mf_spmm_group_count(const int32_t* rowptr, const int32_t* colidx, int64_t nrows, int G, int32_t* counts, void* stream) {
    if (!rowptr) MF_FAIL_ARG(1, "rowptr is NULL");
    if (!colidx) MF_FAIL_ARG(2, "colidx is NULL");
    if (nrows < 0) MF_FAIL_ARG(3, "nrows < 0");
    if (G != 2 && G != 4) MF_FAIL_ARG(4, "group size must be 2 or 4");
    if (!counts) MF_FAIL_ARG(5, "counts is NULL");
    if (nrows == 0) return 0;
    const long long ngroups = (nrows + G - 1) / G;
    const unsigned blocks = (unsigned)((ngroups + 127) / 128);
    if (G == 4) spmm_group_count_kernel<4><<<blocks, 128, 0, (cudaStream_t)stream>>>(rowptr, colidx, nrows, ngroups, counts);
    else spmm_group_count_kernel<2><<<blocks, 128, 0, (cudaStream_t)stream>>>(rowptr, colidx, nrows, ngroups, counts);
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_spmm_group_fill(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t nrows, int G,
                                  const int64_t* ustart, int32_t* ucols, double* uvals, void* stream) {
    if (!rowptr) MF_FAIL_ARG(1, "rowptr is NULL");
    if (!colidx) MF_FAIL_ARG(2, "colidx is NULL");
    if (!vals) MF_FAIL_ARG(3, "vals is NULL");
    if (nrows < 0) MF_FAIL_ARG(4, "nrows < 0");
    if (G != 2 && G != 4) MF_FAIL_ARG(5, "group size must be 2 or 4");
    if (!ustart) MF_FAIL_ARG(6, "ustart is NULL");
    if (!ucols) MF_FAIL_ARG(7, "ucols is NULL");
    if (!uvals) MF_FAIL_ARG(8, "uvals is NULL");
    if (nrows == 0) return 0;
    const long long ngroups = (nrows + G - 1) / G;
    const unsigned blocks = (unsigned)((ngroups + 127) / 128);
    if (G == 4) spmm_group_fill_kernel<4><<<blocks, 128, 0, (cudaStream_t)stream>>>(rowptr, colidx, vals, nrows, ngroups, (const long long*)ustart, ucols, uvals);
    else spmm_group_fill_kernel<2><<<blocks, 128, 0, (cudaStream_t)stream>>>(rowptr, colidx, vals, nrows, ngroups, (const long long*)ustart, ucols, uvals);
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_spmm_grouped_c128(const int64_t* ustart, const int32_t* ucols, const double* uvals, int64_t nrows, int G,
                                    const mf_c128* Q, int64_t ldq, int r, mf_c128* Y, int64_t ldy, void* stream) {
    if (!ustart) MF_FAIL_ARG(1, "ustart is NULL");
    if (!ucols) MF_FAIL_ARG(2, "ucols is NULL");
    if (!uvals) MF_FAIL_ARG(3, "uvals is NULL");
    if (nrows < 0) MF_FAIL_ARG(4, "nrows < 0");
    if (r <= 0 || r > 512) MF_FAIL_ARG(8, "need 0 < r <= 512");
    if (G != mf_spmm_group_size(r)) MF_FAIL_ARG(5, "group size must equal mf_spmm_group_size(r)");
    if (!Q || ldq < r) MF_FAIL_ARG(6, "Q is NULL or ldq < r");
    if (!Y || ldy < r) MF_FAIL_ARG(9, "Y is NULL or ldy < r");
    if (nrows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long ngroups = (nrows + G - 1) / G;
    const long long blocks = (ngroups * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) MF_FAIL_ARG(4, "nrows too large for one launch");
#define GSPMM(GG, C, NY) spmm_grouped_kernel<GG, C><<<dim3((unsigned)blocks, NY), 256, 0, st>>>((const long long*)ustart, ucols, uvals, nrows, (const cplx*)Q, ldq, r, (cplx*)Y, ldy)
    if (r <= 32) GSPMM(4, 1, 1);
    else if (r <= 64) GSPMM(4, 2, 1);
    else if (r <= 128) GSPMM(4, 4, 1);
    else if (G == 4) GSPMM(4, 4, (r + 127) / 128);                 // column slices of 128
    else if (r <= 256) GSPMM(2, 8, 1);
    else GSPMM(2, 16, 1);
#undef GSPMM
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_spmm_csr_f64(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t nrows,
                               const double* Q, int64_t ldq, int r, double* Y, int64_t ldy, void* stream) {
    if (!rowptr) MF_FAIL_ARG(1, "rowptr is NULL");
    if (!colidx) MF_FAIL_ARG(2, "colidx is NULL");
    if (!vals) MF_FAIL_ARG(3, "vals is NULL");
    if (nrows < 0) MF_FAIL_ARG(4, "nrows < 0");
    if (!Q || ldq < r) MF_FAIL_ARG(5, "Q is NULL or ldq < r");
    if (r <= 0 || r > 512) MF_FAIL_ARG(7, "need 0 < r <= 512");
    if (!Y || ldy < r) MF_FAIL_ARG(8, "Y is NULL or ldy < r");
    if (nrows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (nrows * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) MF_FAIL_ARG(4, "nrows too large for one launch");
#define RSPMM(C) spmm_csr_f64_kernel<C><<<(unsigned)blocks, 256, 0, st>>>(rowptr, colidx, vals, nrows, Q, ldq, r, Y, ldy)
    if (r <= 32) RSPMM(1); else if (r <= 64) RSPMM(2); else if (r <= 128) RSPMM(4); else if (r <= 256) RSPMM(8); else RSPMM(16);
#undef RSPMM
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_spmm_grouped_f64(const int64_t* ustart, const int32_t* ucols, const double* uvals, int64_t nrows, int G,
                                   const double* Q, int64_t ldq, int r, double* Y, int64_t ldy, void* stream) {
    if (!ustart) MF_FAIL_ARG(1, "ustart is NULL");
    if (!ucols) MF_FAIL_ARG(2, "ucols is NULL");
    if (!uvals) MF_FAIL_ARG(3, "uvals is NULL");
    if (nrows < 0) MF_FAIL_ARG(4, "nrows < 0");
    if (r <= 0 || r > 512) MF_FAIL_ARG(8, "need 0 < r <= 512");
    if (G != mf_spmm_group_size_f64(r)) MF_FAIL_ARG(5, "group size must equal mf_spmm_group_size_f64(r)");
    if (!Q || ldq < r) MF_FAIL_ARG(6, "Q is NULL or ldq < r");
    if (!Y || ldy < r) MF_FAIL_ARG(9, "Y is NULL or ldy < r");
    if (nrows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long ngroups = (nrows + G - 1) / G;
    const long long blocks = (ngroups * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) MF_FAIL_ARG(4, "nrows too large for one launch");
#define GRSPMM(GG, C, NY) spmm_grouped_f64_kernel<GG, C><<<dim3((unsigned)blocks, NY), 256, 0, st>>>((const long long*)ustart, ucols, uvals, nrows, Q, ldq, r, Y, ldy)
    if (r <= 32) GRSPMM(4, 1, 1); else if (r <= 64) GRSPMM(4, 2, 1); else if (r <= 128) GRSPMM(4, 4, 1);
    else if (G == 4) GRSPMM(4, 8, (r + 255) / 256);                // float64: four rows of 256 columns fit; slices of 256 above
    else if (r <= 256) GRSPMM(2, 8, 1); else GRSPMM(2, 16, 1);
#undef GRSPMM
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_project_rhs_f64(const int32_t* colptr, const int32_t* rowidx, const double* vals, int m,
                                  const double* Q, int64_t ldq, int r, int64_t row0, int64_t nlocal, double* Br, int64_t ldb, void* stream) {
    if (!colptr) MF_FAIL_ARG(1, "colptr is NULL");
    if (!rowidx) MF_FAIL_ARG(2, "rowidx is NULL");
    if (!vals) MF_FAIL_ARG(3, "vals is NULL");
    if (m <= 0) MF_FAIL_ARG(4, "m <= 0");
    if (!Q || ldq < r) MF_FAIL_ARG(5, "Q is NULL or ldq < r");
    if (r <= 0) MF_FAIL_ARG(7, "r <= 0");
    if (!Br || ldb < m) MF_FAIL_ARG(10, "Br is NULL or ldb < m");
    project_rhs_f64_kernel<<<m, 256, 0, (cudaStream_t)stream>>>(colptr, rowidx, vals, Q, ldq, r, row0, nlocal, Br, ldb);
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_spmm_csr2_c128(const int32_t* rowptr, const int32_t* colidx, const void* vals0, const void* vals1, int val_is_real,
                                 int64_t nrows, const mf_c128* Q, int64_t ldq, int r, mf_c128* Y0, int64_t ldy0, mf_c128* Y1, int64_t ldy1,
                                 void* stream) {
    if (!rowptr) MF_FAIL_ARG(1, "rowptr is NULL");
    if (!colidx) MF_FAIL_ARG(2, "colidx is NULL");
    if (!vals0) MF_FAIL_ARG(3, "vals0 is NULL");
    if (!vals1) MF_FAIL_ARG(4, "vals1 is NULL");
    if (nrows < 0) MF_FAIL_ARG(6, "nrows < 0");
    if (!Q || ldq < r) MF_FAIL_ARG(7, "Q is NULL or ldq < r");
    if (r <= 0 || r > 512) MF_FAIL_ARG(9, "need 0 < r <= 512");
    if (!Y0 || ldy0 < r) MF_FAIL_ARG(10, "Y0 is NULL or ldy0 < r");
    if (!Y1 || ldy1 < r) MF_FAIL_ARG(12, "Y1 is NULL or ldy1 < r");
    if (nrows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (nrows * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) MF_FAIL_ARG(6, "nrows too large for one launch");
#define SPMM2(C, RL) spmm_csr2_kernel<C, RL><<<(unsigned)blocks, 256, 0, st>>>(rowptr, colidx, vals0, vals1, nrows, (const cplx*)Q, ldq, r, (cplx*)Y0, ldy0, (cplx*)Y1, ldy1)
#define SPMM2_R(RL) do { if (r <= 32) SPMM2(1, RL); else if (r <= 64) SPMM2(2, RL); else if (r <= 128) SPMM2(4, RL); else if (r <= 256) SPMM2(8, RL); else SPMM2(16, RL); } while (0)
    if (val_is_real) SPMM2_R(true); else SPMM2_R(false);
#undef SPMM2_R
#undef SPMM2
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_spmm_csr2_f64(const int32_t* rowptr, const int32_t* colidx, const double* vals0, const double* vals1, int64_t nrows,
                                const double* Q, int64_t ldq, int r, double* Y0, int64_t ldy0, double* Y1, int64_t ldy1, void* stream) {
    if (!rowptr) MF_FAIL_ARG(1, "rowptr is NULL");
    if (!colidx) MF_FAIL_ARG(2, "colidx is NULL");
    if (!vals0) MF_FAIL_ARG(3, "vals0 is NULL");
    if (!vals1) MF_FAIL_ARG(4, "vals1 is NULL");
    if (nrows < 0) MF_FAIL_ARG(5, "nrows < 0");
    if (!Q || ldq < r) MF_FAIL_ARG(6, "Q is NULL or ldq < r");
    if (r <= 0 || r > 512) MF_FAIL_ARG(8, "need 0 < r <= 512");
    if (!Y0 || ldy0 < r) MF_FAIL_ARG(9, "Y0 is NULL or ldy0 < r");
    if (!Y1 || ldy1 < r) MF_FAIL_ARG(11, "Y1 is NULL or ldy1 < r");
    if (nrows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (nrows * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) MF_FAIL_ARG(5, "nrows too large for one launch");
#define RSPMM2(C) spmm_csr2_f64_kernel<C><<<(unsigned)blocks, 256, 0, st>>>(rowptr, colidx, vals0, vals1, nrows, Q, ldq, r, Y0, ldy0, Y1, ldy1)
    if (r <= 32) RSPMM2(1); else if (r <= 64) RSPMM2(2); else if (r <= 128) RSPMM2(4); else if (r <= 256) RSPMM2(8); else RSPMM2(16);
#undef RSPMM2
    MF_CHECK_LAUNCH();
    return 0;
}
