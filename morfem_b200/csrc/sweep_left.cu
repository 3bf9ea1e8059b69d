// Left-looking blocked LU sweep for large reduced models (r up to 512), templated on the element type (complex128 and the
// real float64 twin -- the reference's own arithmetic, implementation.py:190).  One CTA per frequency point.
//
// Why left-looking: at r = 256 the complex matrix is 1 MiB -- neither shared memory nor the register file holds it, and
// with two points in flight per SM the 296 matrices do not fit in L2 either.  A right-looking LU (sweep_stream.cu) reads
// AND writes the whole trailing matrix once per 16-column panel (~10 MB of DRAM traffic per point, every access a
// read-modify-write with its load latency in front of the DMMAs).  Here the trailing matrix never exists in memory:
//
//   for every block column j (16 columns):
//     LOAD    the block column of A(t) = c0 A0 + c1 A1 + c2 A2 (operators are L2 resident; rows taken in pivoted order
//             through perm[]) straight into DMMA accumulator registers: warp w owns the 16-row blocks w, w + NW, ...
//     CHAIN   for k < j:  C_b += (-L[b, k]) U[k, j]  for every owned block row b > k   (DMMA; the L panels are immutable
//             once written: they stream global -> a per-lane shared-memory FIFO by cp.async, prefetched across the steps);
//             block row k + 1 is final after step k: U[k+1, j] = L_{k+1,k+1}^-1 C_{k+1} (product with the inverted 16 x 16
//             unit-lower block, also DMMA) is published to shared memory for the next step.  One CTA barrier per step.
//     PANEL   the rows below (positions >= 16 j) go to shared memory and are factored with partial pivoting exactly as in
//             sweep_blocked.cu / sweep_stream.cu (CTA-wide arg-max per column, LAPACK's izamax magnitude and tie-break,
//             8-column register panels, negated multipliers, reciprocal pivots on the diagonal);
//     STORE   multipliers -> global panel j, INDEXED BY ORIGINAL ROW (row exchanges only swap two entries of perm[]: no data
//             ever moves, later block columns read "their" rows through perm[]); U blocks -> global in fragment order;
//             the 16 x 16 inverses of L_jj and U_jj (for the chain and for the back substitution) -> global.
//   The right-hand sides cb(t) Br are one more block column (no panel), so L is never revisited; the back substitution
//   is a 16-step chain of DMMA products with the inverted diagonal blocks of U.
//
// Traffic per point at r = 256, complex128: L panels re-read once per later block column, 5.6 MB of READS of immutable data
// (no read-modify-write, no hazards, 16 flop per byte) + 2 MB of operator reads from L2 + 1.1 MB written once.
// Reference semantics: implementation.py:468-480, :526-533 (lu_factor / lu_solve of the symmetrised system matrix);
// test_helpers.py:9-14 for the impedance matrix that gsm_finish_kernel completes.
#include <stdlib.h>
#include "sweep_blocked.cuh"

namespace {

// ---- element-type traits ---------------------------------------------------------------------------------------------
template <typename T> struct Num;

template <> struct Num<double> {
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ double one() { return 1.0; }
    static __device__ __forceinline__ double scale(double s, double a) { return s * a; }
    static __device__ __forceinline__ void axpy(double& acc, double s, double a) { acc = fma(s, a, acc); }      // acc += s a, s real
    static __device__ __forceinline__ void fma_(double& acc, double a, double b) { acc = fma(a, b, acc); }      // acc += a b
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double neg(double a) { return -a; }
    static __device__ __forceinline__ double abs1(double a) { return fabs(a); }
    static __device__ __forceinline__ double recip(double a) { return 1.0 / a; }
    // D(8x8) += A(8x4) B(4x8): one DMMA
    static __device__ __forceinline__ void mma(double& c0, double& c1, double a, double b) { dmma884(c0, c1, a, b); }
    static __device__ __forceinline__ double ldsv(unsigned addr) {
        double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v;
    }
    static __device__ __forceinline__ void cp16(void* dst, const void* src) {
        const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(src) : "memory");
    }
    // One FIFO stage (a 8 x 16 A operand in fragment order: element (g, 4 kk + t) at 32 kk + 4 g + t) from 8 matrix rows of 16 elements;
    // `src_tl` = row of this lane's g, plus t.  Two 16-byte copies per lane (pairs of neighbouring columns): the 8-byte cp.async this
    // replaces cost four shared-memory wavefronts per instruction -- 40 % of all shared-memory wavefronts of the float64 kernel
    // (profiles/r02_sweep_left_f64_ncu.md).
    static __device__ __forceinline__ void fifo_rows(double* stage, const double* src_tl, const int lane) {
        const int g = lane >> 2, t = lane & 3;
        const double* row = src_tl - t;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int kk = 2 * c + (t >> 1), tp = t & 1;
            cp16(stage + 32 * kk + 4 * g + 2 * tp, row + 4 * kk + 2 * tp);
        }
    }
    // One FIFO stage from 128 elements that are already in fragment order in global memory
    static __device__ __forceinline__ void fifo_flat(double* stage, const double* src, const int lane) {
#pragma unroll
        for (int c = 0; c < 2; ++c) cp16(stage + 64 * c + 2 * lane, src + 64 * c + 2 * lane);
    }
    // j zs acc  (the impedance matrix is purely imaginary for a real model)
    static __device__ __forceinline__ cplx jz(double zs, double acc) { return cmake(0.0, zs * acc); }
};

template <> struct Num<cplx> {
    static __device__ __forceinline__ cplx zero() { return cmake(0.0, 0.0); }
    static __device__ __forceinline__ cplx one() { return cmake(1.0, 0.0); }
    static __device__ __forceinline__ cplx scale(double s, cplx a) { return cmake(s * a.x, s * a.y); }
    static __device__ __forceinline__ void axpy(cplx& acc, double s, cplx a) { acc.x = fma(s, a.x, acc.x); acc.y = fma(s, a.y, acc.y); }
    static __device__ __forceinline__ void fma_(cplx& acc, cplx a, cplx b) { cfma(acc, a, b); }
    static __device__ __forceinline__ cplx mul(cplx a, cplx b) { return cmul(a, b); }
    static __device__ __forceinline__ cplx neg(cplx a) { return cmake(-a.x, -a.y); }
    static __device__ __forceinline__ double abs1(cplx a) { return fabs(a.x) + fabs(a.y); }
    static __device__ __forceinline__ cplx recip(cplx a) { return (a.y == 0.0) ? cmake(1.0 / a.x, 0.0) : crecip2(a); }
    // complex block product as four real DMMAs on the interleaved operands
    static __device__ __forceinline__ void mma(cplx& c0, cplx& c1, cplx a, cplx b) {
        dmma884(c0.x, c1.x, a.x, b.x); dmma884(c0.y, c1.y, a.x, b.y);
        dmma884(c0.x, c1.x, -a.y, b.y); dmma884(c0.y, c1.y, a.y, b.x);
    }
    static __device__ __forceinline__ cplx ldsv(unsigned addr) {
        cplx v; asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v;
    }
    static __device__ __forceinline__ void cp_async(void* dst, const void* src) {
        const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(src) : "memory");
    }
    static __device__ __forceinline__ void fifo_rows(cplx* stage, const cplx* src_tl, const int lane) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) cp_async(stage + 32 * kk + lane, src_tl + 4 * kk);
    }
    static __device__ __forceinline__ void fifo_flat(cplx* stage, const cplx* src, const int lane) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) cp_async(stage + 32 * kk + lane, src + 32 * kk + lane);
    }
    static __device__ __forceinline__ cplx jz(double zs, cplx acc) { return cmake(-zs * acc.y, zs * acc.x); }
};

template <typename T>
struct SweepParamsL {
    const T* A0; const T* A1; const T* A2; long long lda;      // symmetrised reduced operators (NULL = zero)
    const T* Br; long long ldb;                                // reduced port matrix r x m
    int r, m;
    const double* c0; const double* c1; const double* c2; const double* cb; const double* zs;
    long long F;
    T* X;         // F x r x m or NULL
    cplx* S;      // F x m x m or NULL
    int* info;    // F or NULL
    T* ws; long long ws_stride;   // per-CTA workspace slots (elements)
    unsigned long long* timing;   // debugging aid (MF_LEFT_TIMING): per-phase clock64 sums of CTA 0, or NULL
    int remap;                    // look-ahead body: panel warps on SM sub-partitions 0/1, run-ahead warps on 2/3 (MF_LEFT_REMAP)
};

struct CandKeyL { double v; int pos; int pad; };

// cp.async group bookkeeping with compiler memory barriers: the FIFO stages are read back with plain loads and no CTA
// barrier sits between the wait and those loads
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { v.x += __shfl_xor_sync(FULL, v.x, off); v.y += __shfl_xor_sync(FULL, v.y, off); }
    return v;
}

// fragment-major offsets inside a 16 x 16 block (256 elements):
//   B operand: element (k-row i, column c) -> slice kk = i / 4, column tile ct = c / 8, lane 4 (c % 8) + i % 4
//   A operand: element (row i, k-column c) -> row tile rb8 = i / 8, slice kk = c / 4, lane 4 (i % 8) + c % 4
__device__ __forceinline__ int bfrag_off(int i, int c) { return (((i >> 2) * 2 + (c >> 3)) << 5) + ((c & 7) << 2) + (i & 3); }
__device__ __forceinline__ int afrag_off(int i, int c) { return (((i >> 3) * 4 + (c >> 2)) << 5) + ((i & 7) << 2) + (c & 3); }

// ---- inner panel factorisation by the whole CTA (8 columns; see sweep_stream.cu panel_factor_mw for the description) ------
// barrier of the threads that factor the panel: the whole CTA (BARN = 0) or a named barrier of BARN threads (the look-ahead
// version of the kernel factors the panel with half of the warps while the other half runs ahead)
template <int BARN>
__device__ __forceinline__ void csync() {
    if (BARN == 0) __syncthreads();
    else asm volatile("bar.sync 1, %0;" :: "n"(BARN) : "memory");
}

#ifdef MF_PANEL_CLOCKS
__device__ unsigned long long g_panel_clk[8];     // diagnostic build only: clock64 sums of the stages of a panel column step (CTA 0, thread 0)
#define PCLK(i) do { if (blockIdx.x == 0 && tid == 0) { const long long now_ = clock64(); g_panel_clk[i] += (unsigned long long)(now_ - pc_); pc_ = now_; } } while (0)
#else
#define PCLK(i) do { } while (0)
#endif

template <typename T, int SL, int NW, int BARN = 0>
__device__ __forceinline__ void panel_factor_t(T* PB, const int LDp, const int rows, const int row0, const int tid,
                                               CandKeyL* candk, T* candrow, int* pvl, int* info_sh, const int info_base) {
    constexpr int NT = NW * 32;
    const int lane = tid & 31, warp = tid >> 5;
    // warps whose first row lies beyond the panel take part in the eight column barriers only
    const int nact = min(NW, (rows - row0 + 31) >> 5);
    if (warp >= nact) {
#pragma unroll
        for (int j = 0; j < 8; ++j) csync<BARN>();
        return;
    }
    T a[SL][8];
    int pos[SL];
    bool act[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int row = row0 + tid + NT * s;
        pos[s] = row;
        act[s] = row < rows;
        const int sw = swz(row & 7);
        const T* src = PB + row * LDp + row0;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[s][c] = act[s] ? src[c ^ sw] : Num<T>::zero();
    }
#ifdef MF_PANEL_CLOCKS
    long long pc_ = clock64();
#endif
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int Tg = row0 + j;
        PCLK(6);                                                 // tail of the previous column's update
        double vb = act[0] ? Num<T>::abs1(a[0][j]) : -1.0;
        int pbest = pos[0], bs = 0;
        T cand = a[0][j];
#pragma unroll
        for (int s = 1; s < SL; ++s) {
            const double v = act[s] ? Num<T>::abs1(a[s][j]) : -1.0;
            if (v > vb || (v == vb && pos[s] < pbest)) { vb = v; pbest = pos[s]; bs = s; cand = a[s][j]; }
        }
        T rc = Num<T>::zero();
        if (vb > 0.0) rc = Num<T>::recip(cand);                  // speculative reciprocal of this lane's candidate
        const int hi = __double2hiint(vb);
        const int hmax = __reduce_max_sync(FULL, hi);
        bool own = (hi == hmax);
        if (__popc(__ballot_sync(FULL, own)) != 1) {
            const unsigned lo = (unsigned)__double2loint(vb);
            const unsigned lmax = __reduce_max_sync(FULL, own ? lo : 0u);
            own = own && (lo == lmax);
            if (__popc(__ballot_sync(FULL, own)) != 1) {         // exact tie: first maximum (lowest position), as izamax / idamax
                const int pmin = __reduce_min_sync(FULL, own ? pbest : 0x7fffffff);
                own = own && (pbest == pmin);
            }
        }
        PCLK(0);                                                 // candidate + warp arg-max
        CandKeyL* ck = candk + (j & 1) * NW;
        T* cr = candrow + (j & 1) * NW * 8;
        if (own) {                                               // this warp's candidate: key and finished row
            ck[warp].v = vb; ck[warp].pos = pbest;
#pragma unroll
            for (int s = 0; s < SL; ++s)
                if (s == bs) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) cr[warp * 8 + c] = (c == j) ? rc : a[s][c];
                }
        }
        PCLK(1);                                                 // reciprocal + candidate row to shared memory
        csync<BARN>();
        PCLK(2);                                                 // barrier
        // global winner among the NW warp candidates: lane w looks at candidate w, then the same warp arg-max
        double cv = -2.0; int cp = 0x7fffffff;
        if (lane < nact) {
            const unsigned cks = (unsigned)__cvta_generic_to_shared(ck + lane);
            long long pbits;
            asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=d"(cv), "=l"(pbits) : "r"(cks));
            cp = (int)pbits;
        }
        const int chi = __double2hiint(cv);
        const int chmax = __reduce_max_sync(FULL, chi);
        bool cown = (chi == chmax);
        if (__popc(__ballot_sync(FULL, cown)) != 1) {
            const unsigned clo = (unsigned)__double2loint(cv);
            const unsigned clmax = __reduce_max_sync(FULL, cown ? clo : 0u);
            cown = cown && (clo == clmax);
            if (__popc(__ballot_sync(FULL, cown)) != 1) {
                const int cpmin = __reduce_min_sync(FULL, cown ? cp : 0x7fffffff);
                cown = cown && (cp == cpmin);
            }
        }
        const int gw = __ffs(__ballot_sync(FULL, cown)) - 1;
        const int P = __shfl_sync(FULL, cp, gw);
        const double gv = __shfl_sync(FULL, cv, gw);
        PCLK(3);                                                 // second-level arg-max
        if (own && warp == gw) {                                 // this lane held the pivot row: retire the slot
#pragma unroll
            for (int s = 0; s < SL; ++s) if (s == bs) act[s] = false;
        }
        const unsigned crs = (unsigned)__cvta_generic_to_shared(cr + gw * 8);
        T u[8];
        const T rcp = Num<T>::ldsv(crs + (unsigned)sizeof(T) * j);
#pragma unroll
        for (int c = j + 1; c < 8; ++c) u[c] = Num<T>::ldsv(crs + (unsigned)sizeof(T) * c);
        if (tid < 8) {                                           // the pivot row goes to its final place; (Tg & 7) == j
            const T v = Num<T>::ldsv(crs + (unsigned)sizeof(T) * tid);
            PB[Tg * LDp + row0 + (tid ^ swz(j))] = v;
        }
        if (tid == 8) { pvl[j] = P; if (!(gv > 0.0) && *info_sh == 0) *info_sh = info_base + Tg + 1; }
        PCLK(4);                                                 // pivot row back from shared memory
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const T nl = Num<T>::neg(Num<T>::mul(a[s][j], rcp));  // negated multiplier
            a[s][j] = nl;
#pragma unroll
            for (int c = j + 1; c < 8; ++c) Num<T>::fma_(a[s][c], nl, u[c]);
            if (pos[s] == Tg) pos[s] = P;
        }
        PCLK(5);                                                 // update issued
    }
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        if (act[s]) {
            const int q = pos[s];
            const int sw = swz(q & 7);
            T* dst = PB + q * LDp + row0;
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c ^ sw] = a[s][c];
        }
    }
}

// row exchanges of the first inner panel + U12 = L11^-1 A12 for one column c of the second (L11 is stored negated)
template <typename T>
__device__ __forceinline__ void stepb_column_t(T* M, const int LD, const int row0, const int c, const int* pv) {
    const int cbase = c & ~7, cin = c & 7;
    T* colp = M + row0 * LD + cbase;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int P = pv[j];
        if (P != row0 + j) {
            T* x = colp + j * LD + (cin ^ swz(j));
            T* y = M + P * LD + cbase + (cin ^ swz(P & 7));
            const T tmp = *x; *x = *y; *y = tmp;
        }
    }
    T u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = colp[j * LD + (cin ^ swz(j))];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        const T* lrow = M + (row0 + j) * LD + row0;
        const int sw = swz(j);
#pragma unroll
        for (int i = 0; i < j; ++i) Num<T>::fma_(u[j], lrow[i ^ sw], u[i]);
    }
#pragma unroll
    for (int j = 1; j < 8; ++j) colp[j * LD + (cin ^ swz(j))] = u[j];
}

// ======================================================================================================================
// The kernel, second version of the body (the first one -- block ownership, a CTA barrier per chain step, the product with
// the inverted diagonal block inside every chain link -- is in the history; its ncu capture, profiles/r02_sweep_left_v1_ncu.md,
// showed 1.23 M warp instructions per point of which 8 % DMMA and 24 % of the stall samples at the per-step barrier):
//   * accumulators live in the register form DMMA wants (re[2] / im[2] quads) -- no register shuffles around the MMAs;
//   * ownership by 8-row TILE, cyclic over the warps, and a static (unrolled) unit loop with a closed-form prefetch
//     iterator instead of the dynamic one;
//   * the chain is dataflow: U[k, j] is announced by one flag word per tile, consumers wait on the two flags they need,
//     there is no CTA barrier inside a block column's chain;
//   * the rows of a finished 16-row block are multiplied by the inverse of their unit-lower diagonal block ONCE, when the
//     block is factored ("L~[b, k] = L_bb^-1 (-L[b, k])", in place in global memory); a later block column then gets
//     U[b, j] = L_bb^-1 A[b, j] + sum_k L~[b, k] U[k, j] with the first product at load time, so a chain link is ONE
//     8 x 16 x 16 tile update plus the publication -- no triangular product on the critical path;
//   * operator loads are branch-free (all in flight at once); warps without rows skip the panel arithmetic.
// ======================================================================================================================
template <typename T> struct Acc;
template <> struct Acc<double> {                                  // one 8 x 8 tile: lane (g, t) holds row g, columns 2t, 2t + 1
    double v[2];
    __device__ __forceinline__ void zero() { v[0] = 0.0; v[1] = 0.0; }
    __device__ __forceinline__ void mma1(double a, double b) { dmma884(v[0], v[1], a, b); }
    __device__ __forceinline__ void mma2(double, double) {}
    __device__ __forceinline__ double get(int e) const { return v[e]; }
    __device__ __forceinline__ void set(int e, double x) { v[e] = x; }
};
template <> struct Acc<cplx> {
    double re[2], im[2];
    __device__ __forceinline__ void zero() { re[0] = re[1] = im[0] = im[1] = 0.0; }
    // the four real DMMAs of a complex block product, issued as two passes so that DMMAs on one accumulator are far apart
    __device__ __forceinline__ void mma1(cplx a, cplx b) { dmma884(re[0], re[1], a.x, b.x); dmma884(im[0], im[1], a.x, b.y); }
    __device__ __forceinline__ void mma2(cplx a, cplx b) { dmma884(re[0], re[1], -a.y, b.y); dmma884(im[0], im[1], a.y, b.x); }
    __device__ __forceinline__ cplx get(int e) const { return cmake(re[e], im[e]); }
    __device__ __forceinline__ void set(int e, cplx x) { re[e] = x.x; im[e] = x.y; }
};

// Inverse of a 16 x 16 triangular diagonal block of the factored panel (rows 0..15 of PB, swizzled, leading dimension 16), by ONE
// warp: the two 8 x 8 diagonal blocks by substitution (one column per lane, 7 dependent steps instead of 15), the off-diagonal
// block as two 8 x 8 x 8 DMMA products.  LOWER: L = I - S, S = the stored (negated) multipliers, inv = [[A^-1, 0], [C^-1 S_B A^-1, C^-1]];
// UPPER: U with the reciprocal pivots stored on the diagonal, inv = [[A^-1, -A^-1 B C^-1], [0, C^-1]].  The result goes to `outg`
// (global) and, if given, `outs` (shared) in A-fragment order; scr = 256 elements of shared scratch owned by the warp.
template <typename T, bool UPPER>
__device__ __forceinline__ void tri_inverse16(const T* PB, T* scr, T* outg, T* outs, const int lane) {
    T* D0 = scr;                                                  // inverse of the leading 8 x 8 block, row-major
    T* D1 = scr + 64;                                             // inverse of the trailing 8 x 8 block
    T* T1 = scr + 128;                                            // intermediate product
    T* W = scr + 192;                                             // off-diagonal block of the inverse
    const int g = lane >> 2, t = lane & 3;
    if (lane < 16) {
        const int blk = lane >> 3, c = lane & 7, o = 8 * blk;
        T x[8];
        if (!UPPER) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                T acc = (i == c) ? Num<T>::one() : Num<T>::zero();
#pragma unroll
                for (int k = 0; k < i; ++k) Num<T>::fma_(acc, PB[mphys(o + i, o + k, 16)], x[k]);
                x[i] = acc;
            }
        } else {
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                T acc = (i == c) ? Num<T>::one() : Num<T>::zero();
#pragma unroll
                for (int k = i + 1; k < 8; ++k) Num<T>::fma_(acc, Num<T>::neg(PB[mphys(o + i, o + k, 16)]), x[k]);
                x[i] = Num<T>::mul(acc, PB[mphys(o + i, o + i, 16)]);
            }
        }
        T* D = blk ? D1 : D0;
#pragma unroll
        for (int i = 0; i < 8; ++i) D[i * 8 + c] = x[i];
    }
    __syncwarp();
    {
        // first product: LOWER  T1 = S_B A^-1 (S_B = rows 8..15, columns 0..7);  UPPER  T1 = B C^-1 (B = rows 0..7, columns 8..15)
        const T* Dr = UPPER ? D1 : D0;
        Acc<T> acc; acc.zero();
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const T a = UPPER ? PB[mphys(g, 8 + 4 * kk + t, 16)] : PB[mphys(8 + g, 4 * kk + t, 16)];
            const T b = Dr[(4 * kk + t) * 8 + g];
            acc.mma1(a, b); acc.mma2(a, b);
        }
        T1[g * 8 + 2 * t] = acc.get(0); T1[g * 8 + 2 * t + 1] = acc.get(1);
    }
    __syncwarp();
    {
        // second product: LOWER  W = C^-1 T1;  UPPER  W = -(A^-1 T1)
        const T* Dl = UPPER ? D0 : D1;
        Acc<T> acc; acc.zero();
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const T a = Dl[g * 8 + 4 * kk + t];
            const T b = T1[(4 * kk + t) * 8 + g];
            acc.mma1(a, b); acc.mma2(a, b);
        }
        W[g * 8 + 2 * t] = UPPER ? Num<T>::neg(acc.get(0)) : acc.get(0);
        W[g * 8 + 2 * t + 1] = UPPER ? Num<T>::neg(acc.get(1)) : acc.get(1);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int e = lane + 32 * q, i = e >> 4, c = e & 15;
        T v;
        if (i < 8) v = (c < 8) ? D0[i * 8 + c] : (UPPER ? W[i * 8 + c - 8] : Num<T>::zero());
        else v = (c < 8) ? (UPPER ? Num<T>::zero() : W[(i - 8) * 8 + c]) : D1[(i - 8) * 8 + c - 8];
        outg[afrag_off(i, c)] = v;
        if (outs) outs[afrag_off(i, c)] = v;
    }
    __syncwarp();
}

// NW warps; warp w owns the 8-row tiles w, w + NW, ... (TPW of them at most); MINB CTAs per SM; NSTAGE = FIFO depth.
template <typename T, int NW, int TPW, int MINB, int NSTAGE>
__global__ void __launch_bounds__(NW * 32, MINB) sweep_left2_kernel(SweepParamsL<T> p, int R) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = NW * 32;
    constexpr int SL = (TPW + 3) / 4;                            // rows per thread in the panel factorisation (R <= 8 NW TPW)
    constexpr int RBW = (TPW + 1) / 2;                           // 16-row blocks per warp in the back substitution
    const int r = p.r, m = p.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tl = lane & 3;
    const int nb = R >> 4, ntiles = R >> 3;
    const int mct = (m + 7) >> 3;

    T* BC = reinterpret_cast<T*>(smem_raw);                      // R x 16: U blocks (fragment order) | panel rows (swizzled)
    T* ring = BC + (size_t)R * 16;                               // NW x NSTAGE x 128: per-lane FIFO of L fragments
    T* xch = ring + (size_t)NW * NSTAGE * 128;                   // 256: inverse of the current unit-lower block (A order) / exchange
    T* candrow = xch + 256;                                      // 2 x NW x 8
    CandKeyL* candk = reinterpret_cast<CandKeyL*>(candrow + 2 * NW * 8);   // 2 x NW
    int* perm = reinterpret_cast<int*>(candk + 2 * NW);          // R: position -> original row
    int* uflag = perm + R;                                       // R / 8: epoch at which tile t of the current block column was published
    int* lp = uflag + (R >> 3);                                  // 16 local pivot positions of the current panel
    int* info_sh = lp + 16;

    T* Lg = p.ws + (long long)blockIdx.x * p.ws_stride;          // [nb][R original rows][16]: negated multipliers (L~ for finished blocks)
    T* Ug = Lg + (long long)nb * R * 16;                         // [nb][nb][256]: U[b, k] in A-fragment order
    T* LIg = Ug + (long long)nb * nb * 256;                      // [nb][256]: inverse of the unit-lower diagonal blocks (A order)
    T* UIg = LIg + (long long)nb * 256;                          // [nb][256]: inverse of the upper diagonal blocks (A order)
    T* ringw = ring + (size_t)warp * NSTAGE * 128;
    const bool hasA0 = p.A0 != nullptr, hasA1 = p.A1 != nullptr, hasA2 = p.A2 != nullptr;
    const int imax = min(TPW, max(0, (ntiles - warp + NW - 1) / NW));        // owned tiles: t = warp + NW i, i < imax

    FragOff fop;
    {
        const int sg = swz(g);
        fop.g = g; fop.a0 = tl ^ sg; fop.a1 = (4 + tl) ^ sg; fop.c0 = (2 * tl) ^ sg; fop.c1 = (2 * tl + 1) ^ sg;
        fop.b0 = tl * 16 + (g ^ swz(tl)); fop.b1 = (4 + tl) * 16 + (g ^ swz(4 + tl));
    }
    for (int i = tid; i < (R >> 3); i += NT) uflag[i] = 0;
    int epoch = 0;

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        for (int i = tid; i < R; i += NT) perm[i] = i;
        if (tid == 0) *info_sh = 0;
        __syncthreads();

        Acc<T> C[TPW][2];                                        // [owned tile][column tile]

        for (int j = 0; j <= nb; ++j) {
            const bool isrhs = (j == nb);
            const int jj = isrhs ? nb : j;
            const int nct = isrhs ? mct : 2;
            ++epoch;

            // element (original row o, local column cl) of block column j of [A(t) | cb Br]; branch-free so that all loads of a
            // tile are in flight together (identity on the padded diagonal)
            auto elem = [&](const int o, const int cl) -> T {
                if (isrhs) {
                    const bool in = (o < r) & (cl < m);
                    const T x = __ldg(p.Br + (in ? (long long)o * p.ldb + cl : 0));
                    return in ? Num<T>::scale(cb, x) : Num<T>::zero();
                }
                const int cg = 16 * j + cl;
                const bool in = (o < r) & (cg < r);
                const long long off = in ? (long long)o * p.lda + cg : 0;
                T v = Num<T>::zero();
                if (hasA0) v = Num<T>::scale(c0, __ldg(p.A0 + off));
                if (hasA1) Num<T>::axpy(v, c1, __ldg(p.A1 + off));
                if (hasA2) Num<T>::axpy(v, c2, __ldg(p.A2 + off));
                return in ? v : ((o == cg) ? Num<T>::one() : Num<T>::zero());
            };

            // ---- LOAD ----
            int obase[TPW];                                      // per owned tile: 16 * (original row of lane's row) + tl
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                const int t = warp + NW * i;
                obase[i] = 0;
                C[i][0].zero(); C[i][1].zero();
                if (i < imax) {
                    const int o = perm[8 * t + g];
                    obase[i] = o * 16 + tl;
                    if (t >= 2 * jj) {                           // rows at or below the diagonal block: the plain matrix elements
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                            for (int e = 0; e < 2; ++e) C[i][ct].set(e, elem(o, 8 * ct + 2 * tl + e));
                    } else {                                     // rows of a finished block b: L_bb^-1 A[b, j] (lower triangular inverse)
                        const int b = t >> 1, h = t & 1;
                        const T* li = LIg + (long long)b * 256 + ((h * 4) << 5) + lane;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (kk < 2 || h == 1) {
                                const T a = li[kk << 5];
                                const int ok = perm[16 * b + 4 * kk + tl];
                                const T b0 = elem(ok, g), b1 = elem(ok, 8 + g);
                                C[i][0].mma1(a, b0); if (nct > 1) C[i][1].mma1(a, b1);
                                C[i][0].mma2(a, b0); if (nct > 1) C[i][1].mma2(a, b1);
                            }
                        }
                    }
                }
            }

            // U rows of a finished tile -> shared slot (B-fragment order), global (A-fragment order, for the back substitution), flag
            auto publish = [&](const Acc<T> (&Ct)[2], const int t) {
                const int b = t >> 1, h = t & 1;
                T* slot = BC + b * 256;
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) slot[bfrag_off(8 * h + g, 8 * ct + 2 * tl + e)] = Ct[ct].get(e);
                if (!isrhs) {
                    T* ub = Ug + ((long long)b * nb + j) * 256;
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct) {
                        T* dst = ub + afrag_off(8 * h + g, 8 * ct + 2 * tl);
                        dst[0] = Ct[ct].get(0); dst[1] = Ct[ct].get(1);
                    }
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) *reinterpret_cast<volatile int*>(uflag + t) = epoch;
            };

            // ---- CHAIN ----
            if (jj > 0) {
                // units of this warp: (step k, owned tile i) with t = warp + NW i >= 2 (k + 1); i runs from i0(k) to imax - 1
                auto i0 = [&](const int k) { const int d = 2 * (k + 1) - warp; return d <= 0 ? 0 : (d + NW - 1) / NW; };
                int pk = 0, pi = i0(0);
                auto pnorm = [&]() { while (pk < jj && pi >= imax) { ++pk; pi = i0(pk); } };
                auto issue = [&](const int stage) {
                    int ob = obase[0];
#pragma unroll
                    for (int i = 1; i < TPW; ++i) ob = (pi == i) ? obase[i] : ob;
                    const T* src = Lg + (long long)pk * R * 16 + ob;
                    Num<T>::fifo_rows(ringw + stage * 128, src, lane);
                };
                pnorm();
                int pstage = 0, cstage = 0;
#pragma unroll
                for (int st = 0; st < NSTAGE - 1; ++st) {
                    if (pk < jj) { issue(pstage); ++pi; pnorm(); }
                    cpa_commit();
                    pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                }
#pragma unroll
                for (int i = 0; i < TPW; ++i)                    // block 0 needs no update: U[0, j] = L_00^-1 A[0, j] is final
                    if (i < imax && warp + NW * i < 2) publish(C[i], warp + NW * i);
#pragma unroll 1
                for (int k = 0; k < jj; ++k) {
                    const int ifirst = i0(k);
                    if (ifirst >= imax) continue;                // no tile of this warp below block k any more
                    {                                            // wait for both halves of U[k, j]
                        volatile int* f = uflag + 2 * k;
                        while (f[0] != epoch || f[1] != epoch) { __nanosleep(20); }
                        __threadfence_block();
                    }
                    const T* Uk = BC + k * 256 + lane;
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        if (i >= ifirst && i < imax) {
                            if (pk < jj) { issue(pstage); ++pi; pnorm(); }
                            cpa_commit();
                            pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                            cpa_wait<NSTAGE - 1>();
                            const T* af = ringw + cstage * 128 + lane;
                            cstage = (cstage + 1 == NSTAGE) ? 0 : cstage + 1;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const T a = af[32 * kk];
                                const T b0 = Uk[(kk * 2) << 5], b1 = Uk[(kk * 2 + 1) << 5];
                                C[i][0].mma1(a, b0); if (nct > 1) C[i][1].mma1(a, b1);
                                C[i][0].mma2(a, b0); if (nct > 1) C[i][1].mma2(a, b1);
                            }
                            const int t = warp + NW * i;
                            if ((t >> 1) == k + 1 && k + 1 < jj) publish(C[i], t);      // the tile's last update: its U rows are final
                        }
                    }
                }
                cpa_wait<0>();
            }
            if (isrhs) break;

            // ---- PANEL: rows at positions >= 16 j -> shared memory (swizzled), LU with partial pivoting ----
            T* PB = BC + (size_t)j * 256;
            const int rows = R - 16 * j;
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                const int t = warp + NW * i;
                if (i < imax && t >= 2 * j) {
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                        for (int e = 0; e < 2; ++e) PB[mphys(8 * (t - 2 * j) + g, 8 * ct + 2 * tl + e, 16)] = C[i][ct].get(e);
                }
            }
            __syncthreads();
#pragma unroll 1
            for (int ip = 0; ip < 2; ++ip) {
                const int row0 = 8 * ip;
                int* pvl = lp + row0;
                if (SL > 1 && rows - row0 > NT) panel_factor_t<T, SL, NW>(PB, 16, rows, row0, tid, candk, candrow, pvl, info_sh, 16 * j);
                else panel_factor_t<T, 1, NW>(PB, 16, rows, row0, tid, candk, candrow, pvl, info_sh, 16 * j);
                __syncthreads();
                if (tid < row0) {                                // the exchanges also apply to the multipliers of the first inner panel
                    const int c = tid, cbase = c & ~7, cin = c & 7;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int P = pvl[q], Tg = row0 + q;
                        if (P != Tg) {
                            T* x = PB + Tg * 16 + cbase + (cin ^ swz(q));
                            T* y = PB + P * 16 + cbase + (cin ^ swz(P & 7));
                            const T tmp = *x; *x = *y; *y = tmp;
                        }
                    }
                }
                if (ip == 0) {
                    if (tid < 8) stepb_column_t<T>(PB, 16, 0, 8 + tid, pvl);
                    __syncthreads();
                    const int ntl = rows / 8 - 1;
                    const T b0 = PB[8 + fop.b0], b1 = PB[8 + fop.b1];
                    for (int ti = warp; ti < ntl; ti += NW) {
                        T* rowp = PB + (8 * (1 + ti) + g) * 16;
                        const T a0 = rowp[fop.a0], a1 = rowp[fop.a1];
                        Acc<T> v;
                        v.set(0, rowp[8 + fop.c0]); v.set(1, rowp[8 + fop.c1]);
                        v.mma1(a0, b0); v.mma2(a0, b0);
                        v.mma1(a1, b1); v.mma2(a1, b1);
                        rowp[8 + fop.c0] = v.get(0); rowp[8 + fop.c1] = v.get(1);
                    }
                    __syncthreads();
                }
            }
            __syncthreads();
            // ---- STORE: perm, multipliers (by original row), inverses of the diagonal blocks ----
            if (tid == 0) {
                for (int c = 0; c < 16; ++c) { const int P = 16 * j + lp[c]; const int tmp = perm[16 * j + c]; perm[16 * j + c] = perm[P]; perm[P] = tmp; }
            }
            __syncthreads();
            for (int e = tid; e < (rows - 16) * 16; e += NT) {
                const int lr = 16 + (e >> 4), c = e & 15;
                const int o = perm[16 * j + lr];
                Lg[((long long)j * R + o) * 16 + c] = PB[mphys(lr, c, 16)];
            }
            if (tid < 32) {
                const int c = tid & 15;
                T x[16];
                if (tid < 16) {                                  // column c of (I - S)^-1, S = stored (negated) multipliers of L11
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        T acc = (i == c) ? Num<T>::one() : Num<T>::zero();
#pragma unroll
                        for (int k = 0; k < i; ++k) Num<T>::fma_(acc, PB[mphys(i, k, 16)], x[k]);
                        x[i] = acc;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) { LIg[(long long)j * 256 + afrag_off(i, c)] = x[i]; xch[afrag_off(i, c)] = x[i]; }
                } else {                                         // column c of U11^-1 (reciprocal pivots on the stored diagonal)
#pragma unroll
                    for (int i = 15; i >= 0; --i) {
                        T acc = (i == c) ? Num<T>::one() : Num<T>::zero();
#pragma unroll
                        for (int k = i + 1; k < 16; ++k) Num<T>::fma_(acc, Num<T>::neg(PB[mphys(i, k, 16)]), x[k]);
                        x[i] = Num<T>::mul(acc, PB[mphys(i, i, 16)]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) UIg[(long long)j * 256 + afrag_off(i, c)] = x[i];
                }
            }
            __syncthreads();
            // ---- L~: rows of the finished block j in the earlier panels k < j are multiplied by L_jj^-1, in place ----
            for (int k = warp; k < j; k += NW) {
                T* Lk = Lg + (long long)k * R * 16;
                T bq[4][2];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const T* rowp = Lk + (long long)perm[16 * j + 4 * kk + tl] * 16;
                    bq[kk][0] = rowp[g]; bq[kk][1] = rowp[8 + g];
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    Acc<T> d0, d1;
                    d0.zero(); d1.zero();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (kk < 2 || h == 1) {
                            const T a = xch[((h * 4 + kk) << 5) + lane];
                            d0.mma1(a, bq[kk][0]); d1.mma1(a, bq[kk][1]);
                            d0.mma2(a, bq[kk][0]); d1.mma2(a, bq[kk][1]);
                        }
                    }
                    T* rowp = Lk + (long long)perm[16 * j + 8 * h + g] * 16 + 2 * tl;
                    rowp[0] = d0.get(0); rowp[1] = d0.get(1);
                    rowp[8] = d1.get(0); rowp[9] = d1.get(1);
                }
            }
            __syncthreads();
        }

        // ---- back substitution: x_k = U_kk^-1 (y_k - sum_{k' > k} U[k, k'] x_k'), blocks from the last to the first ----
        // block b belongs to warp b mod NW; its two row tiles of y are read back from the shared slots (B-fragment order)
        __syncthreads();
#pragma unroll
        for (int bi = 0; bi < RBW; ++bi) {
            const int b = warp + bi * NW;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (2 * bi + h < TPW) C[2 * bi + h][ct].set(e, b < nb ? BC[b * 256 + bfrag_off(8 * h + g, 8 * ct + 2 * tl + e)] : Num<T>::zero());
        }
        __syncthreads();
        for (int k = nb - 1; k >= 0; --k) {
#pragma unroll
            for (int bi = 0; bi < RBW; ++bi) {
                if (warp + bi * NW == k && 2 * bi + 1 < TPW) {   // x_k = U_kk^-1 (upper triangular inverse) times the finished block
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                            for (int e = 0; e < 2; ++e) xch[bfrag_off(8 * h + g, 8 * ct + 2 * tl + e)] = C[2 * bi + h][ct].get(e);
                    __syncwarp();
                    const T* ui = UIg + (long long)k * 256 + lane;
                    T* slot = BC + k * 256;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        Acc<T> d0, d1;
                        d0.zero(); d1.zero();
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (h == 0 || kk >= 2) {
                                const T a = ui[(h * 4 + kk) << 5];
                                const T b0 = xch[((kk * 2) << 5) + lane], b1 = xch[((kk * 2 + 1) << 5) + lane];
                                d0.mma1(a, b0); if (mct > 1) d1.mma1(a, b1);
                                d0.mma2(a, b0); if (mct > 1) d1.mma2(a, b1);
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            slot[bfrag_off(8 * h + g, 2 * tl + e)] = d0.get(e);
                            slot[bfrag_off(8 * h + g, 8 + 2 * tl + e)] = d1.get(e);
                        }
                    }
                }
            }
            __syncthreads();
            const T* Xk = BC + k * 256 + lane;
#pragma unroll
            for (int bi = 0; bi < RBW; ++bi) {
                const int b = warp + bi * NW;
                if (b < k && 2 * bi + 1 < TPW) {
                    const T* ub = Ug + ((long long)b * nb + k) * 256 + lane;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const T a = Num<T>::neg(ub[(h * 4 + kk) << 5]);
                            const T b0 = Xk[(kk * 2) << 5], b1 = Xk[(kk * 2 + 1) << 5];
                            C[2 * bi + h][0].mma1(a, b0); if (mct > 1) C[2 * bi + h][1].mma1(a, b1);
                            C[2 * bi + h][0].mma2(a, b0); if (mct > 1) C[2 * bi + h][1].mma2(a, b1);
                        }
                }
            }
        }
        __syncthreads();
        // ---- outputs: x (position k = original unknown k: there is no column pivoting), Z = j zs x^T (cb Br) -> S ----
        auto xs = [&](const int i, const int c) -> T { return BC[(i >> 4) * 256 + bfrag_off(i & 15, c)]; };
        if (p.X) for (int e = tid; e < r * m; e += NT) { const int i = e / m, c = e - i * m; p.X[pt * (long long)r * m + e] = xs(i, c); }
        if (p.info && tid == 0) p.info[pt] = *info_sh;
        if (p.S) {
            for (int e = warp; e < m * m; e += NW) {
                const int a = e / m, b = e - a * m;
                T acc = Num<T>::zero();
                for (int k = lane; k < r; k += 32) Num<T>::fma_(acc, xs(k, a), Num<T>::scale(cb, __ldg(p.Br + (long long)k * p.ldb + b)));
                acc = warp_sum(acc);
                if (lane == 0) p.S[pt * (long long)m * m + e] = Num<T>::jz(p.zs[pt], acc);
            }
        }
        __syncthreads();
    }
}


// ======================================================================================================================
// Look-ahead version.  The ncu capture of the version above (profiles/r02_sweep_left_v2_ncu.md) shows the DMMA pipe 45 % busy with
// two CTAs per SM: a block column is LOAD -> chain (latency bound: 2 j links) -> panel (16 column steps, a CTA barrier each) -> STORE,
// strictly one after the other, and the tensor pipe idles through the panel.  But block column j + 1 only needs panel j for its rows
// at positions >= 16 j: the rows of the finished blocks b < j -- the whole latency-bound chain U[0 .. j-1, j+1] -- depend on panels
// < j only.  So each step of the outer loop now runs
//   phase A   warps 0 .. NW/2-1: panel j + STORE (named barrier, half of the CTA)
//             warps NW/2 .. NW-1: the chain of column j + 1 over the finished blocks b < j, tile by tile (tile t = rows 8t..8t+7 of the
//             pivoted order): L_bb^-1 A[b, j+1] at load time, the b updates as the U blocks above are announced, publication;
//   phase B   all warps: the rows that needed panel j -- the two tiles of block j (their U rows are the last link) and the rows below,
//             which take all j + 1 updates with U already in shared memory, and go to shared memory as panel j + 1.
// ======================================================================================================================
template <typename T, int NW, int TPW, int MINB, int NSTAGE>
__global__ void __launch_bounds__(NW * 32, MINB) sweep_left3_kernel(SweepParamsL<T> p, int R) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = NW * 32;
    constexpr int NWP = NW / 2, NTP = NWP * 32;                  // panel warps / threads
    constexpr int NWE = NW - NWP;                                // warps that run ahead during the panel
    constexpr int SLP = (TPW + 1) / 2;                           // rows per panel thread (R <= 8 NW TPW, NTP threads)
    constexpr int RBW = (TPW + 1) / 2;                           // 16-row blocks per warp in the back substitution
    // warp roles: hardware warp w issues on SM sub-partition w % 4.  With p.remap the panel warps are the hardware warps with bit 1
    // clear (sub-partitions 0 and 1) and the run-ahead warps the others, so the scalar FP64 work of the panel does not queue behind
    // the DMMAs of this CTA's own chain
    const int lane = threadIdx.x & 31, hw = threadIdx.x >> 5;
    const int sp = hw & 3, qd = hw >> 2;
    const int warp = !p.remap ? hw : (sp < 2 ? qd * 2 + sp : NWP + qd * 2 + (sp - 2));
    const int tid = warp * 32 + lane;
    const int r = p.r, m = p.m;
    const int g = lane >> 2, tl = lane & 3;
    const int nb = R >> 4, ntiles = R >> 3;
    const int mct = (m + 7) >> 3;

    T* BC = reinterpret_cast<T*>(smem_raw);                      // R x 16: U blocks (fragment order) | panel rows (swizzled)
    T* ring = BC + (size_t)R * 16;                               // NW x NSTAGE x 128: per-lane FIFO of L fragments
    T* xch = ring + (size_t)NW * NSTAGE * 128;                   // 256: inverse of the current unit-lower block (A order) / exchange
    T* candrow = xch + 256;                                      // 2 x NW x 8
    CandKeyL* candk = reinterpret_cast<CandKeyL*>(candrow + 2 * NW * 8);   // 2 x NW
    int* perm = reinterpret_cast<int*>(candk + 2 * NW);          // R: position -> original row
    int* uflag = perm + R;                                       // R / 8: epoch at which tile t of the current block column was published
    int* lp = uflag + (R >> 3);                                  // 16 local pivot positions of the current panel
    int* info_sh = lp + 16;

    T* Lg = p.ws + (long long)blockIdx.x * p.ws_stride;          // [nb][R original rows][16]: negated multipliers (L~ for finished blocks)
    T* Ug = Lg + (long long)nb * R * 16;                         // [nb][nb][256]: U[b, k] in A-fragment order
    T* LIg = Ug + (long long)nb * nb * 256;                      // [nb][256]: inverse of the unit-lower diagonal blocks (A order)
    T* UIg = LIg + (long long)nb * 256;                          // [nb][256]: inverse of the upper diagonal blocks (A order)
    T* ringw = ring + (size_t)warp * NSTAGE * 128;
    const bool hasA0 = p.A0 != nullptr, hasA1 = p.A1 != nullptr, hasA2 = p.A2 != nullptr;

    FragOff fop;
    {
        const int sg = swz(g);
        fop.g = g; fop.a0 = tl ^ sg; fop.a1 = (4 + tl) ^ sg; fop.c0 = (2 * tl) ^ sg; fop.c1 = (2 * tl + 1) ^ sg;
        fop.b0 = tl * 16 + (g ^ swz(tl)); fop.b1 = (4 + tl) * 16 + (g ^ swz(4 + tl));
    }
    for (int i = tid; i < (R >> 3); i += NT) uflag[i] = 0;
    int epoch0 = 0;
    // phase timers (clock64 sums of CTA 0): 0 point, 1 panel, 2 store + inverses, 3 L~, 4 early chain, 5 phase A (CTA barrier to CTA barrier),
    // 6 phase B, 7 back substitution, 8 column-0 load
    unsigned long long* tim = (p.timing && blockIdx.x == 0) ? p.timing : nullptr;
    auto tick = [&]() -> long long { return tim ? clock64() : 0; };
    auto tock = [&](const int slot, const long long t0, const int who) { if (tim && tid == who) tim[slot] += (unsigned long long)(clock64() - t0); };                                              // epoch of column c of the current point = epoch0 + c + 1

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        const long long t_pt = tick();
        for (int i = tid; i < R; i += NT) perm[i] = i;
        if (tid == 0) *info_sh = 0;
        __syncthreads();

        // element (original row o, local column cl) of block column jc of [A(t) | cb Br] (jc == nb: the right-hand sides)
        auto elem = [&](const int jc, const int o, const int cl) -> T {
            if (jc == nb) {
                const bool in = (o < r) & (cl < m);
                const T x = __ldg(p.Br + (in ? (long long)o * p.ldb + cl : 0));
                return in ? Num<T>::scale(cb, x) : Num<T>::zero();
            }
            const int cg = 16 * jc + cl;
            const bool in = (o < r) & (cg < r);
            const long long off = in ? (long long)o * p.lda + cg : 0;
            T v = Num<T>::zero();
            if (hasA0) v = Num<T>::scale(c0, __ldg(p.A0 + off));
            if (hasA1) Num<T>::axpy(v, c1, __ldg(p.A1 + off));
            if (hasA2) Num<T>::axpy(v, c2, __ldg(p.A2 + off));
            return in ? v : ((o == cg) ? Num<T>::one() : Num<T>::zero());
        };
        // rows of a finished block b at load time: Ct = (L_bb^-1)[rows 8h .. 8h+7] A[b, jc]  (lower triangular inverse, A operand)
        auto init_u = [&](Acc<T> (&Ct)[2], const int t, const int jc, const int nct) {
            const int b = t >> 1, h = t & 1;
            const T* li = LIg + (long long)b * 256 + ((h * 4) << 5) + lane;
            Ct[0].zero(); Ct[1].zero();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (kk < 2 || h == 1) {
                    const T a = li[kk << 5];
                    const int ok = perm[16 * b + 4 * kk + tl];
                    const T b0 = elem(jc, ok, g), b1 = elem(jc, ok, 8 + g);
                    Ct[0].mma1(a, b0); if (nct > 1) Ct[1].mma1(a, b1);
                    Ct[0].mma2(a, b0); if (nct > 1) Ct[1].mma2(a, b1);
                }
            }
        };
        // U rows of a finished tile -> shared slot (B-fragment order), global (A-fragment order, for the back substitution), flag
        auto publish = [&](const Acc<T> (&Ct)[2], const int t, const int jc, const int ep) {
            const int b = t >> 1, h = t & 1;
            T* slot = BC + b * 256;
#pragma unroll
            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                for (int e = 0; e < 2; ++e) slot[bfrag_off(8 * h + g, 8 * ct + 2 * tl + e)] = Ct[ct].get(e);
            if (jc < nb) {
                T* ub = Ug + ((long long)b * nb + jc) * 256;
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
                    T* dst = ub + afrag_off(8 * h + g, 8 * ct + 2 * tl);
                    dst[0] = Ct[ct].get(0); dst[1] = Ct[ct].get(1);
                }
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) *reinterpret_cast<volatile int*>(uflag + t) = ep;
        };
        auto wait_u = [&](const int k, const int ep) {
            volatile int* f = uflag + 2 * k;
            while (f[0] != ep || f[1] != ep) { __nanosleep(20); }
            __threadfence_block();
        };
        // one 8 x 16 x 16 tile update: A fragments from a FIFO stage, B fragments = U[k, .] in its shared slot
        auto unit = [&](Acc<T> (&Ct)[2], const T* af, const T* Uk, const int nct) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const T a = af[32 * kk];
                const T b0 = Uk[(kk * 2) << 5], b1 = Uk[(kk * 2 + 1) << 5];
                Ct[0].mma1(a, b0); if (nct > 1) Ct[1].mma1(a, b1);
                Ct[0].mma2(a, b0); if (nct > 1) Ct[1].mma2(a, b1);
            }
        };

        Acc<T> C[TPW][2];

        // ---- block column 0: every row is below the (empty) factored part: plain load, straight to the panel buffer ----
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            const int t = warp + NW * i;
            if (t < ntiles) {
                const int o = perm[8 * t + g];
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) BC[mphys(8 * t + g, 8 * ct + 2 * tl + e, 16)] = elem(0, o, 8 * ct + 2 * tl + e);
            }
        }
        __syncthreads();
        tock(8, t_pt, 0);

        for (int j = 0; j < nb; ++j) {
            const long long t_a = tick();
            const int jn = j + 1;                                // the column whose chain overlaps with panel j (jn == nb: right-hand sides)
            const int nctn = (jn == nb) ? mct : 2;
            const int epn = epoch0 + jn + 1;
            // ================================================ phase A ================================================
            if (warp < NWP) {
                // ---- panel j (positions >= 16 j, in shared memory) by the panel warps, named barrier 1 ----
                T* PB = BC + (size_t)j * 256;
                const int rows = R - 16 * j;
#pragma unroll 1
                for (int ip = 0; ip < 2; ++ip) {
                    const int row0 = 8 * ip;
                    int* pvl = lp + row0;
                    if (SLP > 1 && rows - row0 > NTP) panel_factor_t<T, SLP, NWP, NTP>(PB, 16, rows, row0, tid, candk, candrow, pvl, info_sh, 16 * j);
                    else panel_factor_t<T, 1, NWP, NTP>(PB, 16, rows, row0, tid, candk, candrow, pvl, info_sh, 16 * j);
                    csync<NTP>();
                    if (tid < row0) {                            // the exchanges also apply to the multipliers of the first inner panel
                        const int c = tid, cbase = c & ~7, cin = c & 7;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int P = pvl[q], Tg = row0 + q;
                            if (P != Tg) {
                                T* x = PB + Tg * 16 + cbase + (cin ^ swz(q));
                                T* y = PB + P * 16 + cbase + (cin ^ swz(P & 7));
                                const T tmp = *x; *x = *y; *y = tmp;
                            }
                        }
                    }
                    if (ip == 0) {
                        if (tid < 8) stepb_column_t<T>(PB, 16, 0, 8 + tid, pvl);
                        csync<NTP>();
                        const int ntl = rows / 8 - 1;
                        const T b0 = PB[8 + fop.b0], b1 = PB[8 + fop.b1];
                        for (int ti = warp; ti < ntl; ti += NWP) {
                            T* rowp = PB + (8 * (1 + ti) + g) * 16;
                            const T a0 = rowp[fop.a0], a1 = rowp[fop.a1];
                            Acc<T> v;
                            v.set(0, rowp[8 + fop.c0]); v.set(1, rowp[8 + fop.c1]);
                            v.mma1(a0, b0); v.mma2(a0, b0);
                            v.mma1(a1, b1); v.mma2(a1, b1);
                            rowp[8 + fop.c0] = v.get(0); rowp[8 + fop.c1] = v.get(1);
                        }
                        csync<NTP>();
                    }
                }
                csync<NTP>();
                tock(1, t_a, 0);
                const long long t_s = tick();
                // ---- STORE: perm, multipliers (by original row), inverses of the diagonal blocks, L~ ----
                // warp 0: inverse of the unit-lower diagonal block (kept in shared memory for L~ below); warp 1: inverse of the upper one
                // (each uses its own idle FIFO stages as scratch); meanwhile one lane of the last panel warp applies the 16 exchanges to perm[]
                if (warp == 0) tri_inverse16<T, false>(PB, ringw, LIg + (long long)j * 256, xch, lane);
                else if (warp == 1) tri_inverse16<T, true>(PB, ringw, UIg + (long long)j * 256, nullptr, lane);
                if (tid == (NWP > 2 ? 64 : 32)) {                 // first lane of warp 2 (of warp 1, after its inverse, when there are two panel warps)
                    int lpr[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) lpr[c] = lp[c];
#pragma unroll
                    for (int c = 0; c < 16; ++c) { const int P = 16 * j + lpr[c]; const int tmp = perm[16 * j + c]; perm[16 * j + c] = perm[P]; perm[P] = tmp; }
                }
                csync<NTP>();
                for (int e = tid; e < (rows - 16) * 16; e += NTP) {
                    const int lr = 16 + (e >> 4), c = e & 15;
                    const int o = perm[16 * j + lr];
                    Lg[((long long)j * R + o) * 16 + c] = PB[mphys(lr, c, 16)];
                }
                tock(2, t_s, 0);
                const long long t_l = tick();
                for (int k = warp; k < j; k += NWP) {            // L~: rows of block j in the earlier panels, times L_jj^-1, in place
                    T* Lk = Lg + (long long)k * R * 16;
                    T bq[4][2];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const T* rowp = Lk + (long long)perm[16 * j + 4 * kk + tl] * 16;
                        bq[kk][0] = rowp[g]; bq[kk][1] = rowp[8 + g];
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        Acc<T> d0, d1;
                        d0.zero(); d1.zero();
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (kk < 2 || h == 1) {
                                const T a = xch[((h * 4 + kk) << 5) + lane];
                                d0.mma1(a, bq[kk][0]); d1.mma1(a, bq[kk][1]);
                                d0.mma2(a, bq[kk][0]); d1.mma2(a, bq[kk][1]);
                            }
                        }
                        T* rowp = Lk + (long long)perm[16 * j + 8 * h + g] * 16 + 2 * tl;
                        rowp[0] = d0.get(0); rowp[1] = d0.get(1);
                        rowp[8] = d1.get(0); rowp[9] = d1.get(1);
                    }
                }
                tock(3, t_l, 0);
                if (warp >= 1 && warp <= 3) tock(12 + warp, t_a, 32 * warp);     // end of phase A work of the other panel warps
            } else if (j > 0) {
                // ---- the chain of column jn over the finished blocks b < j, tile by tile (tiles e, e + NWE, ... < 2 j) ----
                const int ew = warp - NWP;
                const int nt = 2 * j;
                // FIFO producer: (tile pt_, step pk) over my tiles in order, k < b(tile)
                int ptile = ew, pk = 0;
                auto pnorm = [&]() { while (ptile < nt && pk >= (ptile >> 1)) { ptile += NWE; pk = 0; } };
                auto issue = [&](const int stage) {
                    const int o = perm[8 * ptile + g];
                    const T* src = Lg + ((long long)pk * R + o) * 16 + tl;
                    Num<T>::fifo_rows(ringw + stage * 128, src, lane);
                };
                pnorm();
                int pstage = 0, cstage = 0;
#pragma unroll
                for (int st = 0; st < NSTAGE - 1; ++st) {
                    if (ptile < nt) { issue(pstage); ++pk; pnorm(); }
                    cpa_commit();
                    pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                }
#pragma unroll 1
                for (int t = ew; t < nt; t += NWE) {
                    init_u(C[0], t, jn, nctn);
                    const int b = t >> 1;
#pragma unroll 1
                    for (int k = 0; k < b; ++k) {
                        if (ptile < nt) { issue(pstage); ++pk; pnorm(); }
                        cpa_commit();
                        pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                        wait_u(k, epn);
                        cpa_wait<NSTAGE - 1>();
                        unit(C[0], ringw + cstage * 128 + lane, BC + k * 256 + lane, nctn);
                        cstage = (cstage + 1 == NSTAGE) ? 0 : cstage + 1;
                    }
                    publish(C[0], t, jn, epn);
                }
                cpa_wait<0>();
                tock(4, t_a, NTP);
                if (warp - NWP >= 1 && warp - NWP <= 3) tock(8 + warp - NWP, t_a, 32 * warp);   // the other early warps
            }
            __syncthreads();
            tock(5, t_a, 0);
            const long long t_b = tick();
            // ================================================ phase B ================================================
            // rows at positions >= 16 j of column jn: tiles t = 2 j + warp + NW i.  Block j (t = 2j, 2j+1) takes j updates and is the
            // last link of the chain; the rows below take j + 1 updates and become panel jn.
            {
                const int nrem = ntiles - 2 * j;
                const int imax = min(TPW, max(0, (nrem - warp + NW - 1) / NW));
                int obase[TPW];
#pragma unroll
                for (int i = 0; i < TPW; ++i) {
                    const int t = 2 * j + warp + NW * i;
                    obase[i] = 0;
                    C[i][0].zero(); C[i][1].zero();
                    if (i < imax) {
                        const int o = perm[8 * t + g];
                        obase[i] = o * 16 + tl;
                        if ((t >> 1) == j) init_u(C[i], t, jn, nctn);
                        else {
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                for (int e = 0; e < 2; ++e) C[i][ct].set(e, elem(jn, o, 8 * ct + 2 * tl + e));
                        }
                    }
                }
                // units (k, i): k < j for every owned tile, k == j for the tiles below block j (i >= ilast)
                const int ilast = (warp < 2) ? 1 : 0;
                auto ifirst = [&](const int k) { return k < j ? 0 : ilast; };
                int pk = 0, pi = ifirst(0);
                auto pnorm = [&]() { while (pk <= j && pi >= imax) { ++pk; pi = ifirst(pk); } };
                auto issue = [&](const int stage) {
                    int ob = obase[0];
#pragma unroll
                    for (int i = 1; i < TPW; ++i) ob = (pi == i) ? obase[i] : ob;
                    const T* src = Lg + (long long)pk * R * 16 + ob;
                    Num<T>::fifo_rows(ringw + stage * 128, src, lane);
                };
                pnorm();
                int pstage = 0, cstage = 0;
#pragma unroll
                for (int st = 0; st < NSTAGE - 1; ++st) {
                    if (pk <= j) { issue(pstage); ++pi; pnorm(); }
                    cpa_commit();
                    pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                }
                if (j == 0 && warp < 2 && imax > 0) publish(C[0], warp, jn, epn);        // block 0 needs no update
#pragma unroll 1
                for (int k = 0; k <= j; ++k) {
                    const int i0 = ifirst(k);
                    if (i0 >= imax) continue;
                    if (k == j) wait_u(j, epn);                  // the last link; the blocks above were announced in phase A
                    const T* Uk = BC + k * 256 + lane;
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        if (i >= i0 && i < imax) {
                            if (pk <= j) { issue(pstage); ++pi; pnorm(); }
                            cpa_commit();
                            pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                            cpa_wait<NSTAGE - 1>();
                            unit(C[i], ringw + cstage * 128 + lane, Uk, nctn);
                            cstage = (cstage + 1 == NSTAGE) ? 0 : cstage + 1;
                            if (i == 0 && warp < 2 && k == j - 1) publish(C[0], 2 * j + warp, jn, epn);   // block j: U[j, jn] is final
                        }
                    }
                }
                cpa_wait<0>();
                if (jn < nb) {                                   // rows below block j -> panel jn (shared, swizzled; slots >= jn are free)
                    T* PBn = BC + (size_t)jn * 256;
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        const int t = 2 * j + warp + NW * i;
                        if (i < imax && t >= 2 * jn) {
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                for (int e = 0; e < 2; ++e) PBn[mphys(8 * (t - 2 * jn) + g, 8 * ct + 2 * tl + e, 16)] = C[i][ct].get(e);
                        }
                    }
                }
            }
            __syncthreads();
            tock(6, t_b, 0);
        }
        epoch0 += nb + 2;
        const long long t_bs = tick();

        // ---- back substitution: x_k = U_kk^-1 (y_k - sum_{k' > k} U[k, k'] x_k'), blocks from the last to the first ----
        // block b belongs to warp b mod NW; y comes back from the shared slots (B-fragment order).  The U blocks a warp multiplies
        // with stream through its FIFO one tile ahead (they were written a whole factorisation ago: DRAM latency otherwise).
        for (int e = tid; e < nb * 256 * (int)sizeof(T) / 128; e += NT)   // the inverted diagonal blocks: pull them into L2
            asm volatile("prefetch.global.L2 [%0];" :: "l"(UIg + (size_t)e * (128 / sizeof(T))));
#pragma unroll
        for (int bi = 0; bi < RBW; ++bi) {
            const int b = warp + bi * NW;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (2 * bi + h < TPW) C[2 * bi + h][ct].set(e, b < nb ? BC[b * 256 + bfrag_off(8 * h + g, 8 * ct + 2 * tl + e)] : Num<T>::zero());
        }
        // the U blocks of step kq (U[b, kq], b < kq) into L2, a couple of steps before the FIFO asks for them
        auto l2_prefetch_u = [&](const int kq) {
            if (kq < 1) return;
            constexpr int LPB = 256 * (int)sizeof(T) / 128;      // 128-byte lines per block
            for (int e = tid; e < kq * LPB; e += NT) {
                const int b = e / LPB, ln = e - b * LPB;
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Ug + ((long long)b * nb + kq) * 256 + (size_t)ln * (128 / sizeof(T))));
            }
        };
        l2_prefetch_u(nb - 1);
        l2_prefetch_u(nb - 2);
        {
            // units (k, bi, h): k from nb - 1 down to 1, owned blocks b = warp + NW bi < k, the two row tiles of the block
            int pk = nb - 1, pbi = 0, ph = 0;
            auto pnorm = [&]() { while (pk >= 1 && (pbi >= RBW || warp + pbi * NW >= pk)) { --pk; pbi = 0; } };
            auto issue = [&](const int stage) {
                Num<T>::fifo_flat(ringw + stage * 128, Ug + ((long long)(warp + pbi * NW) * nb + pk) * 256 + ph * 128, lane);
            };
            auto pnext = [&]() { ph ^= 1; if (ph == 0) ++pbi; pnorm(); };
            pnorm();
            int pstage = 0, cstage = 0;
#pragma unroll
            for (int st = 0; st < NSTAGE - 1; ++st) {
                if (pk >= 1) { issue(pstage); pnext(); }
                cpa_commit();
                pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
            }
            __syncthreads();
            for (int k = nb - 1; k >= 0; --k) {
#pragma unroll
                for (int bi = 0; bi < RBW; ++bi) {
                    if (warp + bi * NW == k && 2 * bi + 1 < TPW) {   // x_k = U_kk^-1 (upper triangular inverse) times the finished block
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                for (int e = 0; e < 2; ++e) xch[bfrag_off(8 * h + g, 8 * ct + 2 * tl + e)] = C[2 * bi + h][ct].get(e);
                        __syncwarp();
                        const T* ui = UIg + (long long)k * 256 + lane;
                        T* slot = BC + k * 256;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            Acc<T> d0, d1;
                            d0.zero(); d1.zero();
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                if (h == 0 || kk >= 2) {
                                    const T a = ui[(h * 4 + kk) << 5];
                                    const T b0 = xch[((kk * 2) << 5) + lane], b1 = xch[((kk * 2 + 1) << 5) + lane];
                                    d0.mma1(a, b0); if (mct > 1) d1.mma1(a, b1);
                                    d0.mma2(a, b0); if (mct > 1) d1.mma2(a, b1);
                                }
                            }
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                slot[bfrag_off(8 * h + g, 2 * tl + e)] = d0.get(e);
                                slot[bfrag_off(8 * h + g, 8 + 2 * tl + e)] = d1.get(e);
                            }
                        }
                    }
                }
                __syncthreads();
                l2_prefetch_u(k - 2);
                const T* Xk = BC + k * 256 + lane;
#pragma unroll
                for (int bi = 0; bi < RBW; ++bi) {
                    const int b = warp + bi * NW;
                    if (b < k && 2 * bi + 1 < TPW) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (pk >= 1) { issue(pstage); pnext(); }
                            cpa_commit();
                            pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                            cpa_wait<NSTAGE - 1>();
                            const T* af = ringw + cstage * 128 + lane;
                            cstage = (cstage + 1 == NSTAGE) ? 0 : cstage + 1;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const T a = Num<T>::neg(af[32 * kk]);
                                const T b0 = Xk[(kk * 2) << 5], b1 = Xk[(kk * 2 + 1) << 5];
                                C[2 * bi + h][0].mma1(a, b0); if (mct > 1) C[2 * bi + h][1].mma1(a, b1);
                                C[2 * bi + h][0].mma2(a, b0); if (mct > 1) C[2 * bi + h][1].mma2(a, b1);
                            }
                        }
                    }
                }
            }
            cpa_wait<0>();
        }
        __syncthreads();
        tock(7, t_bs, 0);
        auto xs = [&](const int i, const int c) -> T { return BC[(i >> 4) * 256 + bfrag_off(i & 15, c)]; };
        if (p.X) for (int e = tid; e < r * m; e += NT) { const int i = e / m, c = e - i * m; p.X[pt * (long long)r * m + e] = xs(i, c); }
        if (p.info && tid == 0) p.info[pt] = *info_sh;
        if (p.S) {
            for (int e = warp; e < m * m; e += NW) {
                const int a = e / m, b = e - a * m;
                T acc = Num<T>::zero();
                for (int k = lane; k < r; k += 32) Num<T>::fma_(acc, xs(k, a), Num<T>::scale(cb, __ldg(p.Br + (long long)k * p.ldb + b)));
                acc = warp_sum(acc);
                if (lane == 0) p.S[pt * (long long)m * m + e] = Num<T>::jz(p.zs[pt], acc);
            }
        }
        __syncthreads();
        tock(0, t_pt, 0);
    }
}


// ======================================================================================================================
// Two points per CTA in anti-phase, warps specialised by SM sub-partition ("ping-pong").
// Measured (tools/lat_bench.cu, profiles/r02_sweep_left_variants.md): hardware warp w issues on sub-partition w mod 4, and a scalar
// FP64 instruction on a sub-partition whose FP64 pipe is fed DMMAs waits ~60 cycles for it (8.7 -> 69.6 cycles per dependent DFMA) --
// that is the 2 k cycles of a pivot column step in the kernels above, where panel warps and DMMA warps share sub-partitions and
// the two CTAs of an SM drift into opposite phases.  Here ONE 16-warp CTA per SM holds two points:
//   P group: the 4 warps on sub-partition 0 -- panel factorisation, STORE, inverses, L~ (the code of the look-ahead body's panel warps);
//   D group: the 12 warps on sub-partitions 1..3 -- LOAD + CHAIN of a block column (the code of the plain body), back substitution, outputs.
// In half-step h the D group runs stage h/2 of point slot h mod 2 while the P group factors the panel the D group left in the other
// slot one half-step earlier; a CTA barrier separates half-steps.  Stages of a point: 0 = load block column 0, j = 1 .. nb-1 = block
// column j, nb = right-hand sides + back substitution + outputs; panel j follows stage j.  No scalar FP64 ever shares a pipe with a
// DMMA stream of another warp, and the DMMA warps always have a column of the other point to work on.
// ======================================================================================================================
template <typename T, int NSTAGE>
__global__ void __launch_bounds__(512, 1) sweep_left4_kernel(SweepParamsL<T> p, int R, int slot_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = 16, NP = 4, ND = 12, NTP = NP * 32, NTD = ND * 32;
    constexpr int TPW = 3;                                       // 8-row tiles per D warp (R <= 256: 32 tiles over 12 warps)
    constexpr int RBW = 2;                                       // 16-row blocks per D warp in the back substitution
    constexpr int NC = 4;                                        // accumulator tiles per D warp: max(TPW, 2 RBW)
    constexpr int SLP = 2;                                       // rows per P thread (R <= 256, 128 threads)
    const int lane = threadIdx.x & 31, hw = threadIdx.x >> 5;
    const bool isP = (hw & 3) == 0;
    const int gw = isP ? (hw >> 2) : ((hw >> 2) * 3 + (hw & 3) - 1);   // warp index inside the group
    const int gtid = gw * 32 + lane;
    const int r = p.r, m = p.m;
    const int g = lane >> 2, tl = lane & 3;
    const int nb = R >> 4, ntiles = R >> 3, nper = nb + 1;
    const int mct = (m + 7) >> 3;
    T* ring = reinterpret_cast<T*>(smem_raw + 2 * (size_t)slot_smem);   // ND x NSTAGE x 128: FIFOs of the D warps, then 2 x 256: scratch of P warps 0, 1
    T* ringw = isP ? ring + (size_t)ND * NSTAGE * 128 + (size_t)(gw & 1) * 256 : ring + (size_t)gw * NSTAGE * 128;
    const bool hasA0 = p.A0 != nullptr, hasA1 = p.A1 != nullptr, hasA2 = p.A2 != nullptr;
    FragOff fop;
    {
        const int sg = swz(g);
        fop.g = g; fop.a0 = tl ^ sg; fop.a1 = (4 + tl) ^ sg; fop.c0 = (2 * tl) ^ sg; fop.c1 = (2 * tl + 1) ^ sg;
        fop.b0 = tl * 16 + (g ^ swz(tl)); fop.b1 = (4 + tl) * 16 + (g ^ swz(4 + tl));
    }
    auto dsync = [&]() { asm volatile("bar.sync 2, %0;" :: "n"(NTD) : "memory"); };

    // points of slot s: first_s + k stride
    const long long stride = 2LL * gridDim.x;
    long long npt0, npt1;
    {
        const long long f0 = 2LL * blockIdx.x, f1 = f0 + 1;
        npt0 = f0 < p.F ? (p.F - f0 + stride - 1) / stride : 0;
        npt1 = f1 < p.F ? (p.F - f1 + stride - 1) / stride : 0;
    }
    long long hend = 0;
    if (npt0 > 0) hend = 2 * (npt0 * nper - 1) + 1;
    if (npt1 > 0) { const long long e1 = 2 * (npt1 * nper - 1) + 2; hend = e1 > hend ? e1 : hend; }
    for (int s = 0; s < 2; ++s) {
        int* uf = reinterpret_cast<int*>(smem_raw + (size_t)s * slot_smem + sizeof(T) * ((size_t)R * 16 + 256 + 2 * NP * 8) + sizeof(CandKeyL) * 2 * NP) + R;
        for (int i = threadIdx.x; i < (R >> 3); i += NW * 32) uf[i] = 0;
    }
    int epoch0 = 0, epoch1 = 0;
    __syncthreads();

    // MF_LEFT_TIMING (CTA 0): 0 = all half-steps, 1 = P group busy, 2 = D group busy in column stages, 3 = D in stage 0, 4 = D in the last stage,
    // 5 = P busy in the first half of the panels (j < nb / 2), 6 = D busy in the first half of the columns
    unsigned long long* tim = (p.timing && blockIdx.x == 0) ? p.timing : nullptr;
    for (long long h = 0; h < hend; ++h) {
        const long long t_h = tim ? clock64() : 0;
        // ---- which slot, point and stage this group works on in this half-step ----
        const int s = isP ? 1 - (int)(h & 1) : (int)(h & 1);
        const long long q = isP ? (h - 1 - s) / 2 : (h >> 1);    // stage counter of the slot (P: the panel after D stage q)
        const bool started = !isP || h >= 1 + s;
        const long long k = started ? q / nper : 0;
        const int js = started ? (int)(q - k * nper) : 0;
        const long long npts = s ? npt1 : npt0;
        const bool work = started && k < npts && (!isP || js < nb);
        if (work) {
            unsigned char* sbase = smem_raw + (size_t)s * slot_smem;
            T* BC = reinterpret_cast<T*>(sbase);                 // R x 16: U blocks (fragment order) | panel rows (swizzled)
            T* xch = BC + (size_t)R * 16;                        // 256
            T* candrow = xch + 256;                              // 2 x NP x 8
            CandKeyL* candk = reinterpret_cast<CandKeyL*>(candrow + 2 * NP * 8);
            int* perm = reinterpret_cast<int*>(candk + 2 * NP);  // R
            int* uflag = perm + R;                               // R / 8
            int* lp = uflag + (R >> 3);                          // 16
            int* info_sh = lp + 16;
            T* Lg = p.ws + (2LL * blockIdx.x + s) * p.ws_stride;
            T* Ug = Lg + (long long)nb * R * 16;
            T* LIg = Ug + (long long)nb * nb * 256;
            T* UIg = LIg + (long long)nb * 256;
            const long long pt = 2LL * blockIdx.x + s + k * stride;

            if (isP) {
                // =============================== P group: panel js, STORE, inverses, L~ ===============================
                const int j = js, tid = gtid, warp = gw;
                T* PB = BC + (size_t)j * 256;
                const int rows = R - 16 * j;
#pragma unroll 1
                for (int ip = 0; ip < 2; ++ip) {
                    const int row0 = 8 * ip;
                    int* pvl = lp + row0;
                    if (rows - row0 > NTP) panel_factor_t<T, SLP, NP, NTP>(PB, 16, rows, row0, tid, candk, candrow, pvl, info_sh, 16 * j);
                    else panel_factor_t<T, 1, NP, NTP>(PB, 16, rows, row0, tid, candk, candrow, pvl, info_sh, 16 * j);
                    csync<NTP>();
                    if (tid < row0) {                            // the exchanges also apply to the multipliers of the first inner panel
                        const int c = tid, cbase = c & ~7, cin = c & 7;
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) {
                            const int P = pvl[qq], Tg = row0 + qq;
                            if (P != Tg) {
                                T* x = PB + Tg * 16 + cbase + (cin ^ swz(qq));
                                T* y = PB + P * 16 + cbase + (cin ^ swz(P & 7));
                                const T tmp = *x; *x = *y; *y = tmp;
                            }
                        }
                    }
                    if (ip == 0) {
                        if (tid < 8) stepb_column_t<T>(PB, 16, 0, 8 + tid, pvl);
                        csync<NTP>();
                        const int ntl = rows / 8 - 1;
                        const T b0 = PB[8 + fop.b0], b1 = PB[8 + fop.b1];
                        for (int ti = warp; ti < ntl; ti += NP) {
                            T* rowp = PB + (8 * (1 + ti) + g) * 16;
                            const T a0 = rowp[fop.a0], a1 = rowp[fop.a1];
                            Acc<T> v;
                            v.set(0, rowp[8 + fop.c0]); v.set(1, rowp[8 + fop.c1]);
                            v.mma1(a0, b0); v.mma2(a0, b0);
                            v.mma1(a1, b1); v.mma2(a1, b1);
                            rowp[8 + fop.c0] = v.get(0); rowp[8 + fop.c1] = v.get(1);
                        }
                        csync<NTP>();
                    }
                }
                csync<NTP>();
                if (warp == 0) tri_inverse16<T, false>(PB, ringw, LIg + (long long)j * 256, xch, lane);
                else if (warp == 1) tri_inverse16<T, true>(PB, ringw, UIg + (long long)j * 256, nullptr, lane);
                if (tid == 64) {                                 // first lane of warp 2: the 16 exchanges of the panel applied to perm[]
                    int lpr[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) lpr[c] = lp[c];
#pragma unroll
                    for (int c = 0; c < 16; ++c) { const int P = 16 * j + lpr[c]; const int tmp = perm[16 * j + c]; perm[16 * j + c] = perm[P]; perm[P] = tmp; }
                }
                csync<NTP>();
                for (int e = tid; e < (rows - 16) * 16; e += NTP) {
                    const int lr = 16 + (e >> 4), c = e & 15;
                    const int o = perm[16 * j + lr];
                    Lg[((long long)j * R + o) * 16 + c] = PB[mphys(lr, c, 16)];
                }
                for (int kq = warp; kq < j; kq += NP) {          // L~: rows of block j in the earlier panels, times L_jj^-1, in place
                    T* Lk = Lg + (long long)kq * R * 16;
                    T bq[4][2];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const T* rowp = Lk + (long long)perm[16 * j + 4 * kk + tl] * 16;
                        bq[kk][0] = rowp[g]; bq[kk][1] = rowp[8 + g];
                    }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        Acc<T> d0, d1;
                        d0.zero(); d1.zero();
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (kk < 2 || hh == 1) {
                                const T a = xch[((hh * 4 + kk) << 5) + lane];
                                d0.mma1(a, bq[kk][0]); d1.mma1(a, bq[kk][1]);
                                d0.mma2(a, bq[kk][0]); d1.mma2(a, bq[kk][1]);
                            }
                        }
                        T* rowp = Lk + (long long)perm[16 * j + 8 * hh + g] * 16 + 2 * tl;
                        rowp[0] = d0.get(0); rowp[1] = d0.get(1);
                        rowp[8] = d1.get(0); rowp[9] = d1.get(1);
                    }
                }
            } else {
                // =============================== D group: stage js of the point ===============================
                const int warp = gw, tid = gtid;
                const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
                const bool isrhs = (js == nb);
                const int j = js;
                auto elem = [&](const int o, const int cl) -> T {
                    if (isrhs) {
                        const bool in = (o < r) & (cl < m);
                        const T x = __ldg(p.Br + (in ? (long long)o * p.ldb + cl : 0));
                        return in ? Num<T>::scale(cb, x) : Num<T>::zero();
                    }
                    const int cg = 16 * j + cl;
                    const bool in = (o < r) & (cg < r);
                    const long long off = in ? (long long)o * p.lda + cg : 0;
                    T v = Num<T>::zero();
                    if (hasA0) v = Num<T>::scale(c0, __ldg(p.A0 + off));
                    if (hasA1) Num<T>::axpy(v, c1, __ldg(p.A1 + off));
                    if (hasA2) Num<T>::axpy(v, c2, __ldg(p.A2 + off));
                    return in ? v : ((o == cg) ? Num<T>::one() : Num<T>::zero());
                };
                const int imax = min(TPW, max(0, (ntiles - warp + ND - 1) / ND));   // owned tiles: t = warp + ND i
                if (js == 0) {
                    // ---- a new point: identity permutation, block column 0 straight to the panel buffer ----
                    for (int i = tid; i < R; i += NTD) perm[i] = i;
                    if (tid == 0) *info_sh = 0;
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        const int t = warp + ND * i;
                        if (i < imax) {
                            const int o = 8 * t + g;
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                for (int e = 0; e < 2; ++e) BC[mphys(8 * t + g, 8 * ct + 2 * tl + e, 16)] = elem(o, 8 * ct + 2 * tl + e);
                        }
                    }
                } else {
                    const int nct = isrhs ? mct : 2;
                    const int epoch = s ? ++epoch1 : ++epoch0;
                    Acc<T> C[NC][2];
                    // ---- LOAD ----
                    int obase[TPW];
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        const int t = warp + ND * i;
                        obase[i] = 0;
                        C[i][0].zero(); C[i][1].zero();
                        if (i < imax) {
                            const int o = perm[8 * t + g];
                            obase[i] = o * 16 + tl;
                            if (t >= 2 * j) {
#pragma unroll
                                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                    for (int e = 0; e < 2; ++e) C[i][ct].set(e, elem(o, 8 * ct + 2 * tl + e));
                            } else {                             // rows of a finished block b: L_bb^-1 A[b, j]
                                const int b = t >> 1, hh = t & 1;
                                const T* li = LIg + (long long)b * 256 + ((hh * 4) << 5) + lane;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    if (kk < 2 || hh == 1) {
                                        const T a = li[kk << 5];
                                        const int ok = perm[16 * b + 4 * kk + tl];
                                        const T b0 = elem(ok, g), b1 = elem(ok, 8 + g);
                                        C[i][0].mma1(a, b0); if (nct > 1) C[i][1].mma1(a, b1);
                                        C[i][0].mma2(a, b0); if (nct > 1) C[i][1].mma2(a, b1);
                                    }
                                }
                            }
                        }
                    }
                    auto publish = [&](const Acc<T> (&Ct)[2], const int t) {
                        const int b = t >> 1, hh = t & 1;
                        T* slot = BC + b * 256;
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                            for (int e = 0; e < 2; ++e) slot[bfrag_off(8 * hh + g, 8 * ct + 2 * tl + e)] = Ct[ct].get(e);
                        if (!isrhs) {
                            T* ub = Ug + ((long long)b * nb + j) * 256;
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct) {
                                T* dst = ub + afrag_off(8 * hh + g, 8 * ct + 2 * tl);
                                dst[0] = Ct[ct].get(0); dst[1] = Ct[ct].get(1);
                            }
                        }
                        __syncwarp();
                        __threadfence_block();
                        if (lane == 0) *reinterpret_cast<volatile int*>(uflag + t) = epoch;
                    };
                    // ---- CHAIN ----
                    {
                        auto i0 = [&](const int kq) { const int d = 2 * (kq + 1) - warp; return d <= 0 ? 0 : (d + ND - 1) / ND; };
                        int pk = 0, pi = i0(0);
                        auto pnorm = [&]() { while (pk < j && pi >= imax) { ++pk; pi = i0(pk); } };
                        auto issue = [&](const int stage) {
                            int ob = obase[0];
#pragma unroll
                            for (int i = 1; i < TPW; ++i) ob = (pi == i) ? obase[i] : ob;
                            const T* src = Lg + (long long)pk * R * 16 + ob;
                            Num<T>::fifo_rows(ringw + stage * 128, src, lane);
                        };
                        pnorm();
                        int pstage = 0, cstage = 0;
#pragma unroll
                        for (int st = 0; st < NSTAGE - 1; ++st) {
                            if (pk < j) { issue(pstage); ++pi; pnorm(); }
                            cpa_commit();
                            pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                        }
#pragma unroll
                        for (int i = 0; i < TPW; ++i)            // block 0 needs no update
                            if (i < imax && warp + ND * i < 2) publish(C[i], warp + ND * i);
#pragma unroll 1
                        for (int kq = 0; kq < j; ++kq) {
                            const int ifirst = i0(kq);
                            if (ifirst >= imax) continue;
                            {
                                volatile int* f = uflag + 2 * kq;
                                while (f[0] != epoch || f[1] != epoch) { __nanosleep(20); }
                                __threadfence_block();
                            }
                            const T* Uk = BC + kq * 256 + lane;
#pragma unroll
                            for (int i = 0; i < TPW; ++i) {
                                if (i >= ifirst && i < imax) {
                                    if (pk < j) { issue(pstage); ++pi; pnorm(); }
                                    cpa_commit();
                                    pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                                    cpa_wait<NSTAGE - 1>();
                                    const T* af = ringw + cstage * 128 + lane;
                                    cstage = (cstage + 1 == NSTAGE) ? 0 : cstage + 1;
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk) {
                                        const T a = af[32 * kk];
                                        const T b0 = Uk[(kk * 2) << 5], b1 = Uk[(kk * 2 + 1) << 5];
                                        C[i][0].mma1(a, b0); if (nct > 1) C[i][1].mma1(a, b1);
                                        C[i][0].mma2(a, b0); if (nct > 1) C[i][1].mma2(a, b1);
                                    }
                                    const int t = warp + ND * i;
                                    if ((t >> 1) == kq + 1 && kq + 1 < j) publish(C[i], t);
                                }
                            }
                        }
                        cpa_wait<0>();
                    }
                    if (!isrhs) {
                        // ---- rows at positions >= 16 j -> the panel buffer (swizzled); the P group factors it in the next half-step ----
                        T* PB = BC + (size_t)j * 256;
#pragma unroll
                        for (int i = 0; i < TPW; ++i) {
                            const int t = warp + ND * i;
                            if (i < imax && t >= 2 * j) {
#pragma unroll
                                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                    for (int e = 0; e < 2; ++e) PB[mphys(8 * (t - 2 * j) + g, 8 * ct + 2 * tl + e, 16)] = C[i][ct].get(e);
                            }
                        }
                    } else {
                        // ---- every U row of the right-hand sides is published by now: back substitution and outputs ----
                        dsync();
                        for (int e = tid; e < nb * 256 * (int)sizeof(T) / 128; e += NTD)
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(UIg + (size_t)e * (128 / sizeof(T))));
#pragma unroll
                        for (int bi = 0; bi < RBW; ++bi) {
                            const int b = warp + bi * ND;
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                    for (int e = 0; e < 2; ++e)
                                        C[2 * bi + hh][ct].set(e, b < nb ? BC[b * 256 + bfrag_off(8 * hh + g, 8 * ct + 2 * tl + e)] : Num<T>::zero());
                        }
                        auto l2_prefetch_u = [&](const int kq) {
                            if (kq < 1) return;
                            constexpr int LPB = 256 * (int)sizeof(T) / 128;
                            for (int e = tid; e < kq * LPB; e += NTD) {
                                const int b = e / LPB, ln = e - b * LPB;
                                asm volatile("prefetch.global.L2 [%0];" :: "l"(Ug + ((long long)b * nb + kq) * 256 + (size_t)ln * (128 / sizeof(T))));
                            }
                        };
                        l2_prefetch_u(nb - 1);
                        l2_prefetch_u(nb - 2);
                        int pk = nb - 1, pbi = 0, ph = 0;
                        auto pnorm = [&]() { while (pk >= 1 && (pbi >= RBW || warp + pbi * ND >= pk)) { --pk; pbi = 0; } };
                        auto issue = [&](const int stage) {
                            Num<T>::fifo_flat(ringw + stage * 128, Ug + ((long long)(warp + pbi * ND) * nb + pk) * 256 + ph * 128, lane);
                        };
                        auto pnext = [&]() { ph ^= 1; if (ph == 0) ++pbi; pnorm(); };
                        pnorm();
                        int pstage = 0, cstage = 0;
#pragma unroll
                        for (int st = 0; st < NSTAGE - 1; ++st) {
                            if (pk >= 1) { issue(pstage); pnext(); }
                            cpa_commit();
                            pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                        }
                        dsync();
                        for (int kq = nb - 1; kq >= 0; --kq) {
#pragma unroll
                            for (int bi = 0; bi < RBW; ++bi) {
                                if (warp + bi * ND == kq) {      // x_k = U_kk^-1 times the finished block
#pragma unroll
                                    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                                        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                                            for (int e = 0; e < 2; ++e) xch[bfrag_off(8 * hh + g, 8 * ct + 2 * tl + e)] = C[2 * bi + hh][ct].get(e);
                                    __syncwarp();
                                    const T* ui = UIg + (long long)kq * 256 + lane;
                                    T* slot = BC + kq * 256;
#pragma unroll
                                    for (int hh = 0; hh < 2; ++hh) {
                                        Acc<T> d0, d1;
                                        d0.zero(); d1.zero();
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk) {
                                            if (hh == 0 || kk >= 2) {
                                                const T a = ui[(hh * 4 + kk) << 5];
                                                const T b0 = xch[((kk * 2) << 5) + lane], b1 = xch[((kk * 2 + 1) << 5) + lane];
                                                d0.mma1(a, b0); if (mct > 1) d1.mma1(a, b1);
                                                d0.mma2(a, b0); if (mct > 1) d1.mma2(a, b1);
                                            }
                                        }
#pragma unroll
                                        for (int e = 0; e < 2; ++e) {
                                            slot[bfrag_off(8 * hh + g, 2 * tl + e)] = d0.get(e);
                                            slot[bfrag_off(8 * hh + g, 8 + 2 * tl + e)] = d1.get(e);
                                        }
                                    }
                                }
                            }
                            dsync();
                            l2_prefetch_u(kq - 2);
                            const T* Xk = BC + kq * 256 + lane;
#pragma unroll
                            for (int bi = 0; bi < RBW; ++bi) {
                                const int b = warp + bi * ND;
                                if (b < kq) {
#pragma unroll
                                    for (int hh = 0; hh < 2; ++hh) {
                                        if (pk >= 1) { issue(pstage); pnext(); }
                                        cpa_commit();
                                        pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                                        cpa_wait<NSTAGE - 1>();
                                        const T* af = ringw + cstage * 128 + lane;
                                        cstage = (cstage + 1 == NSTAGE) ? 0 : cstage + 1;
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk) {
                                            const T a = Num<T>::neg(af[32 * kk]);
                                            const T b0 = Xk[(kk * 2) << 5], b1 = Xk[(kk * 2 + 1) << 5];
                                            C[2 * bi + hh][0].mma1(a, b0); if (mct > 1) C[2 * bi + hh][1].mma1(a, b1);
                                            C[2 * bi + hh][0].mma2(a, b0); if (mct > 1) C[2 * bi + hh][1].mma2(a, b1);
                                        }
                                    }
                                }
                            }
                        }
                        cpa_wait<0>();
                        dsync();
                        auto xs = [&](const int i, const int c) -> T { return BC[(i >> 4) * 256 + bfrag_off(i & 15, c)]; };
                        if (p.X) for (int e = tid; e < r * m; e += NTD) { const int i = e / m, c = e - i * m; p.X[pt * (long long)r * m + e] = xs(i, c); }
                        if (p.info && tid == 0) p.info[pt] = *info_sh;
                        if (p.S) {
                            for (int e = warp; e < m * m; e += ND) {
                                const int a = e / m, b = e - a * m;
                                T acc = Num<T>::zero();
                                for (int kq = lane; kq < r; kq += 32) Num<T>::fma_(acc, xs(kq, a), Num<T>::scale(cb, __ldg(p.Br + (long long)kq * p.ldb + b)));
                                acc = warp_sum(acc);
                                if (lane == 0) p.S[pt * (long long)m * m + e] = Num<T>::jz(p.zs[pt], acc);
                            }
                        }
                    }
                }
            }
            if (tim && gtid == 0) {
                const unsigned long long dt = (unsigned long long)(clock64() - t_h);
                if (isP) { atomicAdd(tim + 1, dt); if (js < nb / 2) atomicAdd(tim + 5, dt); }
                else { atomicAdd(tim + (js == 0 ? 3 : (js == nb ? 4 : 2)), dt); if (js > 0 && js <= nb / 2) atomicAdd(tim + 6, dt); }
            }
        }
        __syncthreads();
        if (tim && threadIdx.x == 0) atomicAdd(tim, (unsigned long long)(clock64() - t_h));
    }
}


struct LeftGeom { int R, nb, NW, RBW, MINB; size_t smem, slot_elems; int cfg; };

// Geometry per size (cfg): warps per CTA x owned 16-row blocks per warp must cover R / 16 blocks.
//   1: 4 warps x 2 blocks  (R <= 128)      2: 8 warps x 2 (R <= 256)      3: 16 warps x 2 (R <= 512)
//   4: 8 warps x 4 (R <= 512; real twin: two CTAs per SM)      5: 4 warps x 4 (float64, R <= 256: four CTAs per SM)
template <typename T>
LeftGeom left_geom(int r, int m) {
    LeftGeom gm;
    gm.R = (r + 15) / 16 * 16;
    gm.nb = gm.R / 16;
    // measured choice (profiles/r02_sweep_left_variants.md).  complex128: four 4-warp CTAs up to R = 112, two 8-warp CTAs up to 256, one
    // 16-warp CTA above; float64: four 4-warp CTAs up to 128 (two blocks per warp) and up to 160 (four blocks per warp), three 8-warp
    // CTAs up to 256 (at 100 points per CTA: r = 176 1.20 M against 0.97 M, r = 192 1.05 M against 0.76 M points/s), two 8-warp CTAs with
    // four blocks per warp above
    int cfg;
    if (sizeof(T) == 16) cfg = gm.nb <= 7 ? 1 : (gm.nb <= 16 ? 2 : 3);
    else cfg = gm.nb <= 8 ? 1 : (gm.nb <= 10 ? 5 : (gm.nb <= 16 ? 2 : 4));
    const int cfg_auto = cfg;
    if (const char* e = getenv("MF_LEFT_CFG")) {
        const int c = atoi(e);
        if ((c >= 1 && c <= (sizeof(T) == 8 ? 5 : 4)) || (c == 8 && sizeof(T) == 16 && gm.nb >= 2 && gm.nb <= 16)) cfg = c;
    }
    auto shape = [&](int c) {
        switch (c) {
            case 1:  gm.NW = 4; gm.RBW = 2; break;
            case 2:  gm.NW = 8; gm.RBW = 2; break;
            case 3:  gm.NW = 16; gm.RBW = 2; break;
            case 5:  gm.NW = 4; gm.RBW = 4; break;                // float64 only: four 4-warp CTAs per SM up to R = 256
            case 8:  gm.NW = 16; gm.RBW = 2; break;               // complex128, R <= 256: one 16-warp CTA, two points in anti-phase (sweep_left4_kernel)
            default: gm.NW = 8; gm.RBW = 4; break;
        }
    };
    shape(cfg);
    if (gm.NW * gm.RBW < gm.nb) { cfg = cfg_auto; shape(cfg); }   // a hand-picked geometry that does not cover R: fall back
    gm.cfg = cfg;
    constexpr int NSTAGE = 2;
    gm.smem = sizeof(T) * ((size_t)gm.R * 16 + (size_t)gm.NW * NSTAGE * 128 + 256 + 2 * (size_t)gm.NW * 8) + sizeof(CandKeyL) * 2 * gm.NW
            + sizeof(int) * ((size_t)gm.R + (size_t)gm.R / 8 + 16 + 4) + 64;
    gm.slot_elems = (size_t)gm.nb * gm.R * 16 + (size_t)gm.nb * gm.nb * 256 + 2 * (size_t)gm.nb * 256;
    (void)m;
    return gm;
}

template <typename K>
int left_occupancy(K kern, int threads, const LeftGeom& gm, int* per_sm) {
    MF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gm.smem));
    MF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kern, threads, gm.smem));
    return 0;
}

// Measured on B200 (profiles/r02_sweep_left_variants.md): complex128 runs the look-ahead body at every size -- at R = 208..256 it
// only wins while the two CTAs of an SM stay in step, i.e. together with short launches (left_chunk below; in one long launch the plain
// body is ~5..15 % ahead there); float64 (a quarter of the DMMA work, all latency): plain body.
template <typename T>
int left_version(const LeftGeom& gm) {
    (void)gm;
    return sizeof(T) == 16 ? 3 : 2;
}

// Points per CTA and launch (0 = the whole batch in one launch); see launch_left.  Measured (profiles/r02_sweep_left_chunks.md, launches of
// 100..150 points per CTA, points/s single launch -> 2 points per CTA and launch): complex128 look-ahead body R = 208 3.9e5 (plain body) ->
// 4.9e5, 224 3.4e5 -> 4.1e5, 240 3.0e5 -> 3.55e5, 256 2.7e5 -> 3.06e5; no change at R <= 192 and R >= 320 (one CTA per SM there);
// float64 plain body R = 192 1.05e6 -> 1.16e6, 224 7.3e5 -> 8.0e5, 256 5.7e5 -> 6.1e5, 512 9.1e4 -> 9.4e4; no gain at R = 160.
template <typename T>
long long left_chunk(const LeftGeom& gm, int ver) {
    (void)ver;                                                       // both bodies gain at these sizes (plain body at R = 256: 2.7e5 -> 2.9e5)
    if (sizeof(T) == 16) return (gm.R > 192 && gm.R <= 256) ? 2 : 0;
    return gm.R >= 192 ? 2 : 0;
}

template <typename T, int NW, int RBW, int MINB>
int launch_left(SweepParamsL<T> p, const LeftGeom& gm, size_t ws_bytes, cudaStream_t stream) {
    // which body: the look-ahead one (3) or the plain left-looking one (2); MF_LEFT_VER overrides the measured choice
    int ver = left_version<T>(gm);
    if (const char* e = getenv("MF_LEFT_VER")) { const int v = atoi(e); if (v == 2 || v == 3) ver = v; }
    if (getenv("MF_LEFT_LOOKAHEAD")) ver = 3;
    auto kern = ver == 3 ? sweep_left3_kernel<T, NW, 2 * RBW, MINB, 2> : sweep_left2_kernel<T, NW, 2 * RBW, MINB, 2>;
    int per_sm = 0;
    if (int rc = left_occupancy(kern, NW * 32, gm, &per_sm)) return rc;
    if (per_sm < 1) MF_FAIL_ARG(7, "left-looking sweep does not fit on an SM for this (r, m)");
    const size_t slot = sizeof(T) * gm.slot_elems;
    long long grid = (long long)mf_num_sms() * per_sm;
    if (grid > p.F) grid = p.F;
    if ((long long)(ws_bytes / slot) < grid) grid = (long long)(ws_bytes / slot);
    if (grid < 1 || !p.ws) MF_FAIL_ARG(21, "workspace too small for the left-looking blocked sweep (see mf_sweep_ws_bytes)");
    p.ws_stride = (long long)gm.slot_elems;
    p.timing = nullptr;
    // panel warps on their own SM sub-partitions: measured +10 % at R = 128, a loss from R = 192 up (profiles/r02_sweep_left_variants.md)
    static const int env_remap = getenv("MF_LEFT_REMAP") ? atoi(getenv("MF_LEFT_REMAP")) : -1;
    p.remap = env_remap >= 0 ? (env_remap != 0) : (NW == 8 && gm.R <= 128);
    static const bool want_timing = getenv("MF_LEFT_TIMING") != nullptr;     // debugging aid: blocks, prints the phase clocks of CTA 0
    if (want_timing) { MF_CHECK_CUDA(cudaMalloc(&p.timing, 16 * sizeof(unsigned long long))); MF_CHECK_CUDA(cudaMemsetAsync(p.timing, 0, 16 * 8, stream)); }
    // Points per CTA and launch.  The CTAs of an SM start a launch in step (both in their panel, then both in the DMMA phase) and drift
    // apart within ~20 points, after which one CTA's scalar-FP64 panel runs against the other's DMMA stream; cutting the
    // batch into short launches keeps them in step.  0 = one launch for the whole batch.
    long long chunk = left_chunk<T>(gm, ver);
    if (const char* e = getenv("MF_LEFT_CHUNK")) chunk = atoll(e);
    if (want_timing || chunk <= 0 || chunk * grid >= p.F) {
        kern<<<(unsigned)grid, NW * 32, gm.smem, stream>>>(p, gm.R);
        MF_CHECK_LAUNCH();
    } else {
        const long long per_launch = chunk * grid;
        for (long long f0 = 0; f0 < p.F; f0 += per_launch) {
            SweepParamsL<T> q = p;
            q.F = p.F - f0 < per_launch ? p.F - f0 : per_launch;
            q.c0 = p.c0 + f0; q.c1 = p.c1 + f0; q.c2 = p.c2 + f0; q.cb = p.cb + f0;
            if (p.zs) q.zs = p.zs + f0;
            if (p.X) q.X = p.X + f0 * (long long)p.r * p.m;
            if (p.S) q.S = p.S + f0 * (long long)p.m * p.m;
            if (p.info) q.info = p.info + f0;
            kern<<<(unsigned)(q.F < grid ? q.F : grid), NW * 32, gm.smem, stream>>>(q, gm.R);
            MF_CHECK_LAUNCH();
        }
    }
    if (want_timing) {
        unsigned long long h[16];
        MF_CHECK_CUDA(cudaMemcpyAsync(h, p.timing, sizeof(h), cudaMemcpyDeviceToHost, stream));
        MF_CHECK_CUDA(cudaStreamSynchronize(stream));
        cudaFree(p.timing);
        const double pts = (double)((p.F + grid - 1) / grid);
        fprintf(stderr, "[MF_LEFT_TIMING] r=%d m=%d grid=%lld points/CTA~%.0f  cycles per point: total %.0f | col0 load %.0f | panel %.0f | store+inv %.0f | L~ %.0f | "
                        "early chain %.0f | phase A %.0f | phase B %.0f | backsub %.0f\n", p.r, p.m, grid, pts, h[0] / pts, h[8] / pts, h[1] / pts, h[2] / pts,
                h[3] / pts, h[4] / pts, h[5] / pts, h[6] / pts, h[7] / pts);
        fprintf(stderr, "[MF_LEFT_TIMING]   end of phase A work per point: early warps +1..+3: %.0f %.0f %.0f | panel warps 1..3: %.0f %.0f %.0f\n",
                h[9] / pts, h[10] / pts, h[11] / pts, h[13] / pts, h[14] / pts, h[15] / pts);
#ifdef MF_PANEL_CLOCKS
        unsigned long long pc[8];
        MF_CHECK_CUDA(cudaMemcpyFromSymbol(pc, g_panel_clk, sizeof(pc)));
        const double cols = pts * gm.R;
        fprintf(stderr, "[MF_PANEL_CLOCKS] cycles per pivot column: arg-max %.0f | recip + candidate row %.0f | barrier %.0f | second arg-max %.0f | pivot row read %.0f | "
                        "update issue %.0f | update tail %.0f\n", pc[0] / cols, pc[1] / cols, pc[2] / cols, pc[3] / cols, pc[4] / cols, pc[5] / cols, pc[6] / cols);
        memset(pc, 0, sizeof(pc));
        MF_CHECK_CUDA(cudaMemcpyToSymbol(g_panel_clk, pc, sizeof(pc)));
#endif
    }
    if (p.S) return gsm_finish_launch(p.S, p.m, p.F, stream);
    return 0;
}

// the two-points-per-CTA kernel: one CTA per SM, two workspace slots and two shared-memory point slots each
template <typename T>
int launch_left4(SweepParamsL<T> p, const LeftGeom& gm, size_t ws_bytes, cudaStream_t stream) {
    constexpr int NSTAGE = 3, NP = 4;
    size_t slot_smem = sizeof(T) * ((size_t)gm.R * 16 + 256 + 2 * NP * 8) + sizeof(CandKeyL) * 2 * NP + sizeof(int) * ((size_t)gm.R + (size_t)gm.R / 8 + 16 + 4);
    slot_smem = (slot_smem + 127) / 128 * 128;
    const size_t smem = 2 * slot_smem + sizeof(T) * (12 * NSTAGE * 128 + 2 * 256);
    if (smem > 227 * 1024) MF_FAIL_ARG(7, "two-point left-looking sweep does not fit in shared memory for this r");
    auto kern = sweep_left4_kernel<T, NSTAGE>;
    MF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t slot = sizeof(T) * gm.slot_elems;
    long long grid = mf_num_sms();
    if (2 * grid > p.F) grid = (p.F + 1) / 2;
    if ((long long)(ws_bytes / slot) / 2 < grid) grid = (long long)(ws_bytes / slot) / 2;
    if (grid < 1 || !p.ws) MF_FAIL_ARG(21, "workspace too small for the left-looking blocked sweep (see mf_sweep_ws_bytes)");
    p.ws_stride = (long long)gm.slot_elems;
    p.timing = nullptr;
    p.remap = 0;
    static const bool want_timing = getenv("MF_LEFT_TIMING") != nullptr;
    if (want_timing) { MF_CHECK_CUDA(cudaMalloc(&p.timing, 16 * sizeof(unsigned long long))); MF_CHECK_CUDA(cudaMemsetAsync(p.timing, 0, 16 * 8, stream)); }
    kern<<<(unsigned)grid, 512, smem, stream>>>(p, gm.R, (int)slot_smem);
    MF_CHECK_LAUNCH();
    if (want_timing) {
        unsigned long long h[16];
        MF_CHECK_CUDA(cudaMemcpyAsync(h, p.timing, sizeof(h), cudaMemcpyDeviceToHost, stream));
        MF_CHECK_CUDA(cudaStreamSynchronize(stream));
        cudaFree(p.timing);
        const double pairs = (double)((p.F + 2 * grid - 1) / (2 * grid));
        fprintf(stderr, "[MF_LEFT_TIMING] two-point CTA r=%d m=%d grid=%lld  cycles per point pair: all half-steps %.0f | P busy %.0f (first half of the panels %.0f) | "
                        "D busy: columns %.0f (first half %.0f), stage 0 %.0f, last stage %.0f\n", p.r, p.m, grid, h[0] / pairs, h[1] / pairs, h[5] / pairs,
                h[2] / pairs, h[6] / pairs, h[3] / pairs, h[4] / pairs);
    }
    if (p.S) return gsm_finish_launch(p.S, p.m, p.F, stream);
    return 0;
}

template <typename T>
int dispatch_left(const SweepParamsL<T>& p, size_t ws_bytes, cudaStream_t stream) {
    const LeftGeom gm = left_geom<T>(p.r, p.m);
    constexpr bool REAL = sizeof(T) == 8;
    switch (gm.cfg) {
        case 1:  return launch_left<T, 4, 2, REAL ? 4 : 4>(p, gm, ws_bytes, stream);
        case 2:  return launch_left<T, 8, 2, REAL ? 3 : 2>(p, gm, ws_bytes, stream);
        case 3:  return launch_left<T, 16, 2, 1>(p, gm, ws_bytes, stream);
        case 5:  return launch_left<T, 4, 4, REAL ? 4 : 1>(p, gm, ws_bytes, stream);
        case 8:
            if constexpr (!REAL) return launch_left4<T>(p, gm, ws_bytes, stream);
            return launch_left<T, 8, 2, 3>(p, gm, ws_bytes, stream);
        default: return launch_left<T, 8, 4, REAL ? 2 : 1>(p, gm, ws_bytes, stream);
    }
}

template <typename T>
bool left_supports(int r, int m) {
    if (r < 1 || r > 512 || m < 1 || m > MF_MAX_PORTS) return false;
    const LeftGeom gm = left_geom<T>(r, m);
    return gm.smem <= 226 * 1024 && gm.NW * gm.RBW >= gm.nb;
}

template <typename T>
size_t left_ws_bytes(int r, int m, long long F) {
    const LeftGeom gm = left_geom<T>(r, m);
    // resident CTAs per SM of the geometry (the __launch_bounds__ the kernels are built with; the launcher clamps its grid to the slots it is given)
    constexpr bool REAL = sizeof(T) == 8;
    const int per_sm = gm.cfg == 1 ? 4 : gm.cfg == 2 ? (REAL ? 3 : 2) : gm.cfg == 3 ? 1 : gm.cfg == 8 ? 2 : gm.cfg == 5 ? 4 : (REAL ? 2 : 1);
    long long grid = (long long)mf_num_sms() * per_sm; if (grid > F) grid = F; if (grid < 1) grid = 1;
    return sizeof(T) * gm.slot_elems * (size_t)grid;
}

}  // namespace

// The Makefile compiles this file twice, once per element type (MF_LEFT_PART = 1: complex128, 2: float64), so that the two halves of the
// instantiations build side by side; without the macro one translation unit holds both.
#if !defined(MF_LEFT_PART) || MF_LEFT_PART == 1
bool sweep_left_supports_c128(int r, int m) { return left_supports<cplx>(r, m); }
size_t sweep_left_ws_bytes_c128(int r, int m, long long F) { return left_ws_bytes<cplx>(r, m, F); }

int sweep_left_launch_c128(const SweepParams& q, size_t ws_bytes, cudaStream_t stream) {
    SweepParamsL<cplx> p;
    p.A0 = q.A0; p.A1 = q.A1; p.A2 = q.A2; p.lda = q.lda; p.Br = q.Br; p.ldb = q.ldb; p.r = q.r; p.m = q.m;
    p.c0 = q.c0; p.c1 = q.c1; p.c2 = q.c2; p.cb = q.cb; p.zs = q.zs; p.F = q.F; p.X = q.X; p.S = q.S; p.info = q.info;
    p.ws = q.ws; p.ws_stride = 0; p.timing = nullptr;
    return dispatch_left<cplx>(p, ws_bytes, stream);
}
#endif

#if !defined(MF_LEFT_PART) || MF_LEFT_PART == 2
bool sweep_left_supports_f64(int r, int m) { return left_supports<double>(r, m); }
size_t sweep_left_ws_bytes_f64(int r, int m, long long F) { return left_ws_bytes<double>(r, m, F); }

int sweep_left_launch_f64(const double* A0, const double* A1, const double* A2, long long lda, const double* Br, long long ldb, int r, int m,
                          const double* c0, const double* c1, const double* c2, const double* cb, const double* zs, long long F,
                          double* X, cplx* S, int* info, void* ws, size_t ws_bytes, cudaStream_t stream) {
    SweepParamsL<double> p;
    p.A0 = A0; p.A1 = A1; p.A2 = A2; p.lda = lda; p.Br = Br; p.ldb = ldb; p.r = r; p.m = m;
    p.c0 = c0; p.c1 = c1; p.c2 = c2; p.cb = cb; p.zs = zs; p.F = F; p.X = X; p.S = S; p.info = info;
    p.ws = (double*)ws; p.ws_stride = 0; p.timing = nullptr;
    return dispatch_left<double>(p, ws_bytes, stream);
}
#endif
