// Windowed SpMM with TMA staging -- the variant the north star names for stage 2 ("Q tiles staged in shared memory by TMA"),
// for Y = A Q with A = the CSR view of a^T (implementation.py:181-183, scipy csr_matvecs).
//
// One CTA per block of RB = 8 consecutive matrix rows, one warp per row.  The union of the column indices of the block
// (its Q-row WINDOW: for the banded FEM operators nine runs of RB + 2 rows, ~90 rows) is built once per operator on the
// host; every non-zero carries the 8-bit slot of its Q row inside the window instead of a 32-bit column index.  The CTA
// walks the basis in column slices of 512 bytes (32 complex / 64 real columns): the window rows of a slice are fetched by
// bulk asynchronous copies (cp.async.bulk global -> shared, one 512-byte copy per window row, completion counted in bytes
// on an mbarrier; SASS UBLKCP), two slices in flight, and every Q row is then reused from shared memory by all rows of
// the block.  Each lane owns one 16-byte column chunk: with real operator values the update acc += v * q is the same
// arithmetic for a complex column (re, im) and for a pair of real columns, so one kernel serves both element types.
//
// Measured against the register-reuse kernels of sparse.cu in profiles/r02_spmm.md: the shared-memory data path moves
// nnz * r * 16 B either way (the staging adds a write per window row), so this variant trades L2 latency for capacity and
// does not beat the row-grouped kernel; it is kept behind mf_spmm_window_* as the measured alternative.
#include "common.cuh"

namespace {

constexpr int WIN_RB = 8;          // rows per CTA (= warps)
constexpr int WIN_MAXNNZ = 64;     // non-zeros per row held in shared memory
constexpr int WIN_CHUNKS = 32;     // 16-byte chunks per slice (one per lane)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bulk asynchronous copy global -> shared of `bytes` (multiple of 16), completion signalled on `bar` (TMA, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Q / Y are addressed in 16-byte chunks: ldq16 / ldy16 = leading dimension in chunks, r16 = chunks per row.
__global__ void __launch_bounds__(WIN_RB * 32, 2)
spmm_window_kernel(const int* __restrict__ rowptr, const unsigned char* __restrict__ slot, const double* __restrict__ vals,
                   const int* __restrict__ wstart, const int* __restrict__ ucol, long long nrows, int wmax,
                   const double2* __restrict__ Q, long long ldq16, int r16, double2* __restrict__ Y, long long ldy16) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw);            // two mbarriers (one per stage)
    double2* stage = reinterpret_cast<double2*>(smem_raw + 128);                          // 2 x wmax x 32 chunks
    double* rv = reinterpret_cast<double*>(stage + 2 * (size_t)wmax * WIN_CHUNKS);        // RB x MAXNNZ coefficients
    int* wcol = reinterpret_cast<int*>(rv + WIN_RB * WIN_MAXNNZ);                         // wmax window rows (Q row indices)
    unsigned char* rs = reinterpret_cast<unsigned char*>(wcol + wmax);                    // RB x MAXNNZ window slots
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long blk = blockIdx.x;
    const long long row = blk * WIN_RB + warp;
    const int w0 = wstart[blk], W = wstart[blk + 1] - w0;
    int cnt = 0;
    if (row < nrows) {
        const int s = rowptr[row];
        cnt = rowptr[row + 1] - s;
        for (int k = lane; k < cnt; k += 32) { rv[warp * WIN_MAXNNZ + k] = vals[s + k]; rs[warp * WIN_MAXNNZ + k] = slot[s + k]; }
    }
    for (int i = tid; i < W; i += WIN_RB * 32) wcol[i] = ucol[w0 + i];
    if (tid == 0) {
        mbar_init(bar, 1); mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nsl = (r16 + WIN_CHUNKS - 1) / WIN_CHUNKS;
    auto issue = [&](const int sl, const int st) {
        const unsigned bytes = (unsigned)min(WIN_CHUNKS, r16 - WIN_CHUNKS * sl) * 16u;
        if (tid == 0) mbar_expect_tx(bar + st, (unsigned)W * bytes);
        double2* dst = stage + (size_t)st * wmax * WIN_CHUNKS;
        for (int i = tid; i < W; i += WIN_RB * 32)
            bulk_g2s(dst + (size_t)i * WIN_CHUNKS, Q + (long long)wcol[i] * ldq16 + WIN_CHUNKS * sl, bytes, bar + st);
    };
    issue(0, 0);
    if (nsl > 1) issue(1, 1);
    for (int sl = 0; sl < nsl; ++sl) {
        const int st = sl & 1;
        mbar_wait(bar + st, (unsigned)(sl >> 1) & 1u);
        const double2* win = stage + (size_t)st * wmax * WIN_CHUNKS + lane;
        const double* v = rv + warp * WIN_MAXNNZ;
        const unsigned char* sidx = rs + warp * WIN_MAXNNZ;
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const double c = v[k];                                  // broadcast reads
            const double2 q = win[(int)sidx[k] * WIN_CHUNKS];       // conflict-free: one 16-byte chunk per lane
            acc.x = fma(c, q.x, acc.x); acc.y = fma(c, q.y, acc.y);
        }
        const int chunk = WIN_CHUNKS * sl + lane;
        if (row < nrows && chunk < r16) Y[row * ldy16 + chunk] = acc;
        __syncthreads();                                            // every warp is done with this stage
        if (sl + 2 < nsl) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads above before the async writes below
            issue(sl + 2, st);
        }
    }
}

size_t window_smem(int wmax) {
    return 128 + 2 * (size_t)wmax * WIN_CHUNKS * 16 + WIN_RB * WIN_MAXNNZ * 8 + (size_t)wmax * 4 + WIN_RB * WIN_MAXNNZ + 16;
}

int launch_window(const int* rowptr, const unsigned char* slot, const double* vals, const int* wstart, const int* ucol, long long nrows,
                  int wmax, const void* Q, long long ldq16, int r16, void* Y, long long ldy16, cudaStream_t st) {
    const size_t smem = window_smem(wmax);
    if (smem > 113 * 1024) MF_FAIL_ARG(7, "window too large for two CTAs per SM (mf_spmm_window_max_rows)");
    MF_CHECK_CUDA(cudaFuncSetAttribute(spmm_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (nrows + WIN_RB - 1) / WIN_RB;
    if (blocks > 0x7fffffffLL) MF_FAIL_ARG(6, "nrows too large for one launch");
    spmm_window_kernel<<<(unsigned)blocks, WIN_RB * 32, smem, st>>>(rowptr, slot, vals, wstart, ucol, nrows, wmax, (const double2*)Q, ldq16, r16,
                                                                   (double2*)Y, ldy16);
    MF_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" int mf_spmm_window_rows_per_block(void) { return WIN_RB; }
extern "C" int mf_spmm_window_max_nnz_per_row(void) { return WIN_MAXNNZ; }
// largest window (distinct Q rows referenced by one block of rows) the kernel can stage twice with two CTAs per SM
extern "C" int mf_spmm_window_max_rows(void) {
    int w = 255;                                                     // slots are 8 bits
    while (w > 0 && window_smem(w) > 113 * 1024) --w;
    return w;
}

#define MF_WINDOW_ARGS_CHECK()                                                                            \
    if (!rowptr) MF_FAIL_ARG(1, "rowptr is NULL");                                                        \
    if (!slot) MF_FAIL_ARG(2, "slot is NULL");                                                            \
    if (!vals) MF_FAIL_ARG(3, "vals is NULL");                                                            \
    if (!wstart) MF_FAIL_ARG(4, "wstart is NULL");                                                        \
    if (!ucol) MF_FAIL_ARG(5, "ucol is NULL");                                                            \
    if (nrows < 0) MF_FAIL_ARG(6, "nrows < 0");                                                           \
    if (wmax <= 0 || wmax > mf_spmm_window_max_rows()) MF_FAIL_ARG(7, "wmax out of range (mf_spmm_window_max_rows)"); \
    if (!Q || ldq < r) MF_FAIL_ARG(8, "Q is NULL or ldq < r");                                            \
    if (r <= 0) MF_FAIL_ARG(10, "r <= 0");                                                                \
    if (!Y || ldy < r) MF_FAIL_ARG(11, "Y is NULL or ldy < r");                                           \
    if (nrows == 0) return 0;

extern "C" int mf_spmm_window_c128(const int32_t* rowptr, const uint8_t* slot, const double* vals, const int32_t* wstart, const int32_t* ucol,
                                   int64_t nrows, int wmax, const mf_c128* Q, int64_t ldq, int r, mf_c128* Y, int64_t ldy, void* stream) {
    MF_WINDOW_ARGS_CHECK();
    return launch_window(rowptr, slot, vals, wstart, ucol, nrows, wmax, Q, ldq, r, Y, ldy, (cudaStream_t)stream);
}

extern "C" int mf_spmm_window_f64(const int32_t* rowptr, const uint8_t* slot, const double* vals, const int32_t* wstart, const int32_t* ucol,
                                  int64_t nrows, int wmax, const double* Q, int64_t ldq, int r, double* Y, int64_t ldy, void* stream) {
    MF_WINDOW_ARGS_CHECK();
    if ((r & 1) || (ldq & 1) || (ldy & 1)) MF_FAIL_ARG(10, "the float64 window kernel moves pairs of columns: r, ldq and ldy must be even");
    return launch_window(rowptr, slot, vals, wstart, ucol, nrows, wmax, Q, ldq / 2, r / 2, Y, ldy / 2, (cudaStream_t)stream);
}
