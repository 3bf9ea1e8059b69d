// Real float64 twins of the stage-1/2 contractions (the reference's own dtype is real float64 --
// main.py:21-23, implementation.py:190 -- so for real operators and snapshots half the bytes and a quarter of the
// flops of the complex128 kernels suffice, with bit-identical results: every imaginary part would be an exact zero).
//
//   gemm_tn_f64 : C (ra x rb) = A^T B, reduction over the long row index n      (Gram matrices, q^T (a q))
//   gemm_nn_f64 : Out (n x rb) = A (n x ra) W (ra x rb)                          (S R^-1, x w)
//
// Same structure as dense_gemm.cu: DMMA.8x8x4 (one per k-step instead of four), cp.async 3-stage ring, shared leading
// dimensions chosen so that the 8-byte fragment loads of a warp fall into two conflict-free wavefronts, deterministic
// fixed-order reduction of the split-N partials.  Roofline: FP64 pipe for r >~ 100, HBM below (AI = r/8 flop/B).
#include "common.cuh"

namespace {

constexpr int KC = 16;
constexpr int STAGES = 3;

constexpr int TN_TI = 64, TN_TJ = 64, TN_THREADS = 128;
constexpr int TN_LDA = TN_TI + 8, TN_LDB = TN_TJ + 8;        // == 8 (mod 16) doubles: k-rows t land 64 B apart in the banks
constexpr int TN_STAGE_ELEMS = KC * (TN_LDA + TN_LDB);

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(s), "l"(gmem_src), "r"(bytes));
}

__global__ void __launch_bounds__(TN_THREADS, 2)
gemm_tn_f64_kernel(const double* __restrict__ A, long long lda, int ra, const double* __restrict__ B, long long ldb, int rb,
                   long long n, long long rows_per_split, double* __restrict__ part, int vec_ok, int symm) {
    extern __shared__ __align__(16) double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wi = warp >> 1, wj = warp & 1;
    const int ntj = (rb + TN_TJ - 1) / TN_TJ;
    const int ti = blockIdx.x / ntj, tj = blockIdx.x - ti * ntj;
    if (symm && ti > tj) return;               // Gram matrix A^T A: the lower tiles mirror the upper ones
    const int i0 = ti * TN_TI, j0 = tj * TN_TJ;
    const long long n0 = (long long)blockIdx.y * rows_per_split;
    long long n1 = n0 + rows_per_split; if (n1 > n) n1 = n;
    const int nchunks = n1 > n0 ? (int)((n1 - n0 + KC - 1) / KC) : 0;

    auto load_stage = [&](int stage, int chunk) {
        double* As = smem_d + stage * TN_STAGE_ELEMS;
        double* Bs = As + KC * TN_LDA;
        const long long row_base = n0 + (long long)chunk * KC;
        if (vec_ok) {                          // 16-byte copies: rows 16-byte aligned and ra, rb even
#pragma unroll
            for (int q = 0; q < (KC * TN_TI / 2) / TN_THREADS; ++q) {
                const int idx = tid + q * TN_THREADS;
                const int rr = idx / (TN_TI / 2), cc = 2 * (idx - rr * (TN_TI / 2));
                const long long row = row_base + rr;
                const bool ok = row < n1 && (i0 + cc) < ra;
                cp_async16(As + rr * TN_LDA + cc, A + (ok ? row * lda + i0 + cc : 0), ok);
            }
#pragma unroll
            for (int q = 0; q < (KC * TN_TJ / 2) / TN_THREADS; ++q) {
                const int idx = tid + q * TN_THREADS;
                const int rr = idx / (TN_TJ / 2), cc = 2 * (idx - rr * (TN_TJ / 2));
                const long long row = row_base + rr;
                const bool ok = row < n1 && (j0 + cc) < rb;
                cp_async16(Bs + rr * TN_LDB + cc, B + (ok ? row * ldb + j0 + cc : 0), ok);
            }
        } else {
#pragma unroll
            for (int q = 0; q < (KC * TN_TI) / TN_THREADS; ++q) {
                const int idx = tid + q * TN_THREADS;
                const int rr = idx / TN_TI, cc = idx - rr * TN_TI;
                const long long row = row_base + rr;
                const bool ok = row < n1 && (i0 + cc) < ra;
                cp_async8(As + rr * TN_LDA + cc, A + (ok ? row * lda + i0 + cc : 0), ok);
            }
#pragma unroll
            for (int q = 0; q < (KC * TN_TJ) / TN_THREADS; ++q) {
                const int idx = tid + q * TN_THREADS;
                const int rr = idx / TN_TJ, cc = idx - rr * TN_TJ;
                const long long row = row_base + rr;
                const bool ok = row < n1 && (j0 + cc) < rb;
                cp_async8(Bs + rr * TN_LDB + cc, B + (ok ? row * ldb + j0 + cc : 0), ok);
            }
        }
    };

    double c[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }

    for (int s = 0; s < STAGES - 1; ++s) { if (s < nchunks) load_stage(s, s); cp_async_commit(); }
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        { const int nc = ch + STAGES - 1; if (nc < nchunks) load_stage(nc % STAGES, nc); cp_async_commit(); }
        const double* As = smem_d + (ch % STAGES) * TN_STAGE_ELEMS;
        const double* Bs = As + KC * TN_LDA;
#pragma unroll
        for (int kk = 0; kk < KC / 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(kk * 4 + t) * TN_LDA + wi * 32 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(kk * 4 + t) * TN_LDB + wj * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    double* out = part + (long long)blockIdx.y * ra * rb;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = i0 + wi * 32 + i * 8 + g;
            const int col = j0 + wj * 32 + j * 8 + 2 * t;
            if (row < ra) {
                if (col < rb) out[(long long)row * rb + col] = c[i][j][0];
                if (col + 1 < rb) out[(long long)row * rb + col + 1] = c[i][j][1];
            }
        }
}

constexpr int RED_WARPS = 8;
__global__ void __launch_bounds__(RED_WARPS * 32)
reduce_partials_f64_kernel(const double* __restrict__ part, int nsplit, int ra, int rb, double* __restrict__ C, long long ldc, int symm) {
    __shared__ double red[RED_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long total = (long long)ra * rb;
    const long long oidx = (long long)blockIdx.x * 32 + lane;
    long long idx = oidx;
    if (symm && oidx < total) {
        const int i = (int)(oidx / rb), j = (int)(oidx - (long long)i * rb);
        if (i / TN_TI > j / TN_TJ) idx = (long long)j * rb + i;
    }
    double acc0 = 0.0, acc1 = 0.0;
    if (idx < total) {
        int s = warp;
        for (; s + RED_WARPS < nsplit; s += 2 * RED_WARPS) { acc0 += part[(long long)s * total + idx]; acc1 += part[(long long)(s + RED_WARPS) * total + idx]; }
        if (s < nsplit) acc0 += part[(long long)s * total + idx];
    }
    red[warp][lane] = acc0 + acc1;
    __syncthreads();
    if (warp == 0 && oidx < total) {
        double acc = red[0][lane];
#pragma unroll
        for (int w = 1; w < RED_WARPS; ++w) acc += red[w][lane];
        const int i = (int)(oidx / rb), j = (int)(oidx - (long long)i * rb);
        C[i * ldc + j] = acc;
    }
}

// `symm`: only the upper-triangular tiles do work (Gram matrices), so the row index is split into more, shorter ranges
void tn_plan(int ra, int rb, long long n, int& tiles, int& nsplit, long long& rows_per_split, bool symm = false) {
    tiles = ((ra + TN_TI - 1) / TN_TI) * ((rb + TN_TJ - 1) / TN_TJ);
    const int nt = (ra + TN_TI - 1) / TN_TI;
    const int work_tiles = symm ? nt * (nt + 1) / 2 : tiles;
    const int target = 148 * 2 * 2;
    long long max_split = (n + (long long)KC * 8 - 1) / ((long long)KC * 8);
    if (max_split < 1) max_split = 1;
    long long s = target / work_tiles; if (s < 1) s = 1; if (s > max_split) s = max_split;
    rows_per_split = (n + s - 1) / s;
    rows_per_split = (rows_per_split + KC - 1) / KC * KC;
    if (rows_per_split < KC) rows_per_split = KC;
    nsplit = (int)((n + rows_per_split - 1) / rows_per_split); if (nsplit < 1) nsplit = 1;
}

// ------------------------------------------------------------------------------------------------ gemm_nn
constexpr int NN_TM = 128, NN_TN = 64, NN_THREADS = 256;
constexpr int NN_LDA = KC + 4;          // == 4 (mod 16) doubles: the 8 rows of an A fragment hit distinct 32-byte bank groups
constexpr int NN_LDW = NN_TN + 8;
constexpr int NN_STAGE_ELEMS = NN_TM * NN_LDA + KC * NN_LDW;

__global__ void __launch_bounds__(NN_THREADS, 2)
gemm_nn_f64_kernel(const double* __restrict__ A, long long lda, long long n, int ra, const double* __restrict__ W, long long ldw, int rb,
                   double* __restrict__ Out, long long ldo, int w_upper) {
    extern __shared__ __align__(16) double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int ntn = (rb + NN_TN - 1) / NN_TN;
    const long long tm = blockIdx.x / ntn; const int tn = (int)(blockIdx.x - tm * ntn);
    const long long m0 = tm * NN_TM; const int j0 = tn * NN_TN;
    const int kmax = (w_upper && j0 + NN_TN < ra) ? j0 + NN_TN : ra;      // upper-triangular W: rows below the tile are zero
    const int nchunks = (kmax + KC - 1) / KC;

    auto load_stage = [&](int stage, int chunk) {
        double* As = smem_d + stage * NN_STAGE_ELEMS;
        double* Ws = As + NN_TM * NN_LDA;
        const int k0 = chunk * KC;
#pragma unroll
        for (int q = 0; q < (NN_TM * KC) / NN_THREADS; ++q) {
            const int idx = tid + q * NN_THREADS;
            const int rr = idx / KC, cc = idx - rr * KC;
            const bool ok = (m0 + rr) < n && (k0 + cc) < ra;
            cp_async8(As + rr * NN_LDA + cc, A + (ok ? (m0 + rr) * lda + k0 + cc : 0), ok);
        }
#pragma unroll
        for (int q = 0; q < (KC * NN_TN) / NN_THREADS; ++q) {
            const int idx = tid + q * NN_THREADS;
            const int rr = idx / NN_TN, cc = idx - rr * NN_TN;
            const bool ok = (k0 + rr) < ra && (j0 + cc) < rb;
            cp_async8(Ws + rr * NN_LDW + cc, W + (ok ? (long long)(k0 + rr) * ldw + j0 + cc : 0), ok);
        }
    };

    double c[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }

    for (int s = 0; s < STAGES - 1; ++s) { if (s < nchunks) load_stage(s, s); cp_async_commit(); }
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        { const int nc = ch + STAGES - 1; if (nc < nchunks) load_stage(nc % STAGES, nc); cp_async_commit(); }
        const double* As = smem_d + (ch % STAGES) * NN_STAGE_ELEMS;
        const double* Ws = As + NN_TM * NN_LDA;
#pragma unroll
        for (int kk = 0; kk < KC / 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(wm * 32 + i * 8 + g) * NN_LDA + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[(kk * 4 + t) * NN_LDW + wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long row = m0 + wm * 32 + i * 8 + g;
            const int col = j0 + wn * 32 + j * 8 + 2 * t;
            if (row < n) {
                double* dst = Out + row * ldo + col;
                if (col + 1 < rb && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) *reinterpret_cast<double2*>(dst) = make_double2(c[i][j][0], c[i][j][1]);
                else { if (col < rb) dst[0] = c[i][j][0]; if (col + 1 < rb) dst[1] = c[i][j][1]; }
            }
        }
}

}  // namespace

extern "C" size_t mf_gemm_tn_f64_ws_bytes(int ra, int rb, int64_t n) {
    if (ra <= 0 || rb <= 0 || n <= 0) return 16;
    int tiles, nsplit, nsplit2 = 0; long long rps;
    tn_plan(ra, rb, n, tiles, nsplit, rps);
    if (ra == rb) tn_plan(ra, rb, n, tiles, nsplit2, rps, true);
    return sizeof(double) * (size_t)(nsplit > nsplit2 ? nsplit : nsplit2) * ra * rb;
}

extern "C" int mf_gemm_tn_f64(const double* A, int64_t lda, int ra, const double* B, int64_t ldb, int rb, int64_t n,
                              double* C, int64_t ldc, void* ws, size_t ws_bytes, void* stream) {
    if (!A) MF_FAIL_ARG(1, "A is NULL");
    if (ra <= 0 || lda < ra) MF_FAIL_ARG(3, "need 0 < ra <= lda");
    if (!B) MF_FAIL_ARG(4, "B is NULL");
    if (rb <= 0 || ldb < rb) MF_FAIL_ARG(6, "need 0 < rb <= ldb");
    if (n < 0) MF_FAIL_ARG(7, "n < 0");
    if (!C || ldc < rb) MF_FAIL_ARG(8, "C is NULL or ldc < rb");
    if (!ws || ws_bytes < mf_gemm_tn_f64_ws_bytes(ra, rb, n)) MF_FAIL_ARG(10, "workspace too small (mf_gemm_tn_f64_ws_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    int tiles, nsplit; long long rps;
    const int symm = (A == B && lda == ldb && ra == rb) ? 1 : 0;
    tn_plan(ra, rb, n > 0 ? n : 1, tiles, nsplit, rps, symm != 0);
    const size_t smem = sizeof(double) * STAGES * TN_STAGE_ELEMS;
    MF_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int vec_ok = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 && (lda % 2) == 0 && (ldb % 2) == 0 &&
                       (ra % 2) == 0 && (rb % 2) == 0;
    dim3 grid(tiles, nsplit);
    gemm_tn_f64_kernel<<<grid, TN_THREADS, smem, st>>>(A, lda, ra, B, ldb, rb, n, rps, (double*)ws, vec_ok, symm);
    MF_CHECK_LAUNCH();
    const long long total = (long long)ra * rb;
    reduce_partials_f64_kernel<<<(unsigned)((total + 31) / 32), RED_WARPS * 32, 0, st>>>((const double*)ws, nsplit, ra, rb, C, ldc, symm);
    MF_CHECK_LAUNCH();
    return 0;
}

static int gemm_nn_f64_impl(const double* A, int64_t lda, int64_t n, int ra, const double* W, int64_t ldw, int rb,
                           double* Out, int64_t ldo, void* stream, int w_upper) {
    if (!A) MF_FAIL_ARG(1, "A is NULL");
    if (ra <= 0 || lda < ra) MF_FAIL_ARG(4, "need 0 < ra <= lda");
    if (n < 0) MF_FAIL_ARG(3, "n < 0");
    if (!W || ldw < rb) MF_FAIL_ARG(5, "W is NULL or ldw < rb");
    if (rb <= 0) MF_FAIL_ARG(7, "rb <= 0");
    if (!Out || ldo < rb) MF_FAIL_ARG(8, "Out is NULL or ldo < rb");
    if ((const void*)Out == (const void*)A) MF_FAIL_ARG(8, "Out must not alias A");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(double) * STAGES * NN_STAGE_ELEMS;
    MF_CHECK_CUDA(cudaFuncSetAttribute(gemm_nn_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles_m = (n + NN_TM - 1) / NN_TM;
    const int tiles_n = (rb + NN_TN - 1) / NN_TN;
    const long long grid = tiles_m * tiles_n;
    if (grid > 0x7fffffffLL) MF_FAIL_ARG(3, "n too large for one launch");
    gemm_nn_f64_kernel<<<(unsigned)grid, NN_THREADS, smem, st>>>(A, lda, n, ra, W, ldw, rb, Out, ldo, w_upper);
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_gemm_nn_f64(const double* A, int64_t lda, int64_t n, int ra, const double* W, int64_t ldw, int rb,
                              double* Out, int64_t ldo, void* stream) {
    return gemm_nn_f64_impl(A, lda, n, ra, W, ldw, rb, Out, ldo, stream, 0);
}

extern "C" int mf_trmm_nn_f64(const double* A, int64_t lda, int64_t n, int r, const double* W, int64_t ldw,
                              double* Out, int64_t ldo, void* stream) {
    return gemm_nn_f64_impl(A, lda, n, r, W, ldw, r, Out, ldo, stream, 1);
}
