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
    static __device__ __forceinline__ void cp_async(void* dst, const void* src) {
        const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(s), "l"(src) : "memory");
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
template <typename T, int SL, int NW>
__device__ __forceinline__ void panel_factor_t(T* PB, const int LDp, const int rows, const int row0, const int tid,
                                               CandKeyL* candk, T* candrow, int* pvl, int* info_sh, const int info_base) {
    constexpr int NT = NW * 32;
    const int lane = tid & 31, warp = tid >> 5;
    // warps whose first row lies beyond the panel take part in the eight column barriers only
    const int nact = min(NW, (rows - row0 + 31) >> 5);
    if (warp >= nact) {
#pragma unroll
        for (int j = 0; j < 8; ++j) __syncthreads();
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
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int Tg = row0 + j;
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
        __syncthreads();
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
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const T nl = Num<T>::neg(Num<T>::mul(a[s][j], rcp));  // negated multiplier
            a[s][j] = nl;
#pragma unroll
            for (int c = j + 1; c < 8; ++c) Num<T>::fma_(a[s][c], nl, u[c]);
            if (pos[s] == Tg) pos[s] = P;
        }
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

// ---- the kernel ------------------------------------------------------------------------------------------------------
// NW warps; warp w owns the 16-row blocks w, w + NW, ... (RBW of them at most); MINB CTAs per SM (register budget);
// NSTAGE = depth of the per-lane FIFO of L fragments.
template <typename T, int NW, int RBW, int MINB, int NSTAGE>
__global__ void __launch_bounds__(NW * 32, MINB) sweep_left_kernel(SweepParamsL<T> p, int R) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = NW * 32;
    constexpr int SL = (RBW + 1) / 2;                            // rows per thread in the panel factorisation (R <= 16 NW RBW)
    constexpr bool CACHE_B = sizeof(T) == 8;                     // real twin: keep the U fragments of a step in registers
    const int r = p.r, m = p.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nb = R >> 4;                                       // 16-row / 16-column blocks
    const int mct = (m + 7) >> 3;                                // column tiles of the right-hand-side block column

    T* BC = reinterpret_cast<T*>(smem_raw);                      // R x 16: U blocks (fragment order) | panel rows (swizzled)
    T* ring = BC + (size_t)R * 16;                               // NW x NSTAGE x 128: per-lane FIFO of L fragments
    T* xch = ring + (size_t)NW * NSTAGE * 128;                   // 256: C-layout -> B-fragment exchange of a chain link
    T* candrow = xch + 256;                                      // 2 x NW x 8
    CandKeyL* candk = reinterpret_cast<CandKeyL*>(candrow + 2 * NW * 8);   // 2 x NW
    int* perm = reinterpret_cast<int*>(candk + 2 * NW);          // R: position -> original row
    int* lp = perm + R;                                          // 16 local pivot positions of the current panel
    int* info_sh = lp + 16;

    T* Lg = p.ws + (long long)blockIdx.x * p.ws_stride;          // [nb][R original rows][16]: negated multipliers
    T* Ug = Lg + (long long)nb * R * 16;                         // [nb][nb][256]: U[b, k] in A-fragment order
    T* LIg = Ug + (long long)nb * nb * 256;                      // [nb][256]: inverse of the unit-lower diagonal blocks (A order)
    T* UIg = LIg + (long long)nb * 256;                          // [nb][256]: inverse of the upper diagonal blocks (A order)
    T* ringw = ring + (size_t)warp * NSTAGE * 128;
    const bool hasA0 = p.A0 != nullptr, hasA1 = p.A1 != nullptr, hasA2 = p.A2 != nullptr;

    FragOff fop;                                                 // fragment offsets in the swizzled panel (leading dimension 16)
    {
        const int sg = swz(g);
        fop.g = g; fop.a0 = t ^ sg; fop.a1 = (4 + t) ^ sg; fop.c0 = (2 * t) ^ sg; fop.c1 = (2 * t + 1) ^ sg;
        fop.b0 = t * 16 + (g ^ swz(t)); fop.b1 = (4 + t) * 16 + (g ^ swz(4 + t));
    }

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        for (int i = tid; i < R; i += NT) perm[i] = i;
        if (tid == 0) *info_sh = 0;
        __syncthreads();

        T C[RBW][2][2][2];                                       // [owned block][row tile][column tile][e]: rows 8 rb8 + g, columns 8 ct + 2 t + e

        // (U_b | y_b) = D^-1-like product of a finished block with an inverted 16 x 16 diagonal block (A operand, global),
        // published to BC slot b in B-fragment order; the C registers of the block are replaced by the result.
        auto convert = [&](T (&Cb)[2][2][2], const T* inv, const bool upper, const int nct, T* slot, T* ublk) {
#pragma unroll
            for (int rb8 = 0; rb8 < 2; ++rb8)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) xch[bfrag_off(8 * rb8 + g, 8 * ct + 2 * t + e)] = Cb[rb8][ct][e];
            __syncwarp();
            T D[2][2][2];
#pragma unroll
            for (int rb8 = 0; rb8 < 2; ++rb8)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) { D[rb8][ct][0] = Num<T>::zero(); D[rb8][ct][1] = Num<T>::zero(); }
#pragma unroll
            for (int rb8 = 0; rb8 < 2; ++rb8) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const bool need = upper ? (rb8 == 0 || kk >= 2) : (rb8 == 1 || kk < 2);   // triangular: skip the zero 8 x 8 block
                    if (need) {
                        const T a = inv[((rb8 * 4 + kk) << 5) + lane];
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct)
                            if (ct < nct) Num<T>::mma(D[rb8][ct][0], D[rb8][ct][1], a, xch[((kk * 2 + ct) << 5) + lane]);
                    }
                }
            }
#pragma unroll
            for (int rb8 = 0; rb8 < 2; ++rb8)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        slot[bfrag_off(8 * rb8 + g, 8 * ct + 2 * t + e)] = D[rb8][ct][e];
                        Cb[rb8][ct][e] = D[rb8][ct][e];
                    }
                    if (ublk) {                                  // U[b, j] for the back substitution, A-fragment order (two consecutive elements)
                        T* dst = ublk + afrag_off(8 * rb8 + g, 8 * ct + 2 * t);
                        dst[0] = D[rb8][ct][0]; dst[1] = D[rb8][ct][1];
                    }
                }
        };

        for (int j = 0; j <= nb; ++j) {
            const bool isrhs = (j == nb);
            const int jj = isrhs ? nb : j;                       // finished panels to the left
            const int nct = isrhs ? mct : 2;

            // ---- LOAD: block column j of [A(t) | cb Br], rows in pivoted order, into the accumulator registers ----
#pragma unroll
            for (int bi = 0; bi < RBW; ++bi) {
                const int b = warp + bi * NW;
#pragma unroll
                for (int rb8 = 0; rb8 < 2; ++rb8) {
                    const int o = (b < nb) ? perm[16 * b + 8 * rb8 + g] : 0;
#pragma unroll
                    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            T v = Num<T>::zero();
                            const int cl = 8 * ct + 2 * t + e;
                            if (b < nb) {
                                if (!isrhs) {
                                    const int cg = 16 * j + cl;
                                    if (o < r && cg < r) {
                                        const long long off = (long long)o * p.lda + cg;
                                        if (hasA0) v = Num<T>::scale(c0, __ldg(p.A0 + off));
                                        if (hasA1) Num<T>::axpy(v, c1, __ldg(p.A1 + off));
                                        if (hasA2) Num<T>::axpy(v, c2, __ldg(p.A2 + off));
                                    } else if (o == cg) v = Num<T>::one();      // identity on the padded diagonal
                                } else if (o < r && cl < m) v = Num<T>::scale(cb, __ldg(p.Br + (long long)o * p.ldb + cl));
                            }
                            C[bi][rb8][ct][e] = v;
                        }
                }
            }

            // ---- CHAIN: C_b += (-L[b, k]) U[k, j] for k < jj; U[k+1, j] published after step k ----
            if (jj > 0) {
                // unit = (step k, owned block bi, row tile rb8); the producer iterator runs NSTAGE - 1 units ahead
                int pk = 0, pbi = 0, prb = 0, ck = 0, cbi = 0, crb = 0;
                auto seek = [&](int& k, int& bi) {               // first unit at or after (k, bi): owned block b > k
                    while (k < jj) {
                        while (bi < RBW) {
                            const int b = warp + bi * NW;
                            if (b >= nb) { bi = RBW; break; }
                            if (b > k) return;
                            ++bi;
                        }
                        ++k; bi = 0;
                    }
                };
                auto next = [&](int& k, int& bi, int& rb) {
                    if (rb == 0) { rb = 1; return; }
                    rb = 0; ++bi;
                    seek(k, bi);
                };
                auto issue = [&](const int k, const int bi, const int rb, const int stage) {
                    const int b = warp + bi * NW;
                    const int o = perm[16 * b + 8 * rb + g];
                    const T* src = Lg + ((long long)k * R + o) * 16 + t;
                    T* dst = ringw + stage * 128 + lane;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) Num<T>::cp_async(dst + 32 * kk, src + 4 * kk);
                };
                seek(pk, pbi);
                seek(ck, cbi);
                int pstage = 0, cstage = 0;
#pragma unroll
                for (int s = 0; s < NSTAGE - 1; ++s) {
                    if (pk < jj) { issue(pk, pbi, prb, pstage); next(pk, pbi, prb); }
                    cpa_commit();
                    pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                }
                if (warp == 0) convert(C[0], LIg, false, nct, BC, isrhs ? nullptr : Ug + (long long)j * 256);   // block 0: U[0, j]
                __syncthreads();
                for (int k = 0; k < jj; ++k) {
                    const T* Uk = BC + k * 256;                  // U[k, j], B-fragment order
                    T bfr[CACHE_B ? 4 : 1][2];
                    if (CACHE_B) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct) bfr[CACHE_B ? kk : 0][ct] = Uk[((kk * 2 + ct) << 5) + lane];
                    }
                    while (ck == k) {
                        if (pk < jj) { issue(pk, pbi, prb, pstage); next(pk, pbi, prb); }
                        cpa_commit();
                        pstage = (pstage + 1 == NSTAGE) ? 0 : pstage + 1;
                        cpa_wait<NSTAGE - 1>();
                        const T* af = ringw + cstage * 128 + lane;
#pragma unroll
                        for (int bi = 0; bi < RBW; ++bi) {       // static register indexing of C
                            if (bi == cbi) {
#pragma unroll
                                for (int rb = 0; rb < 2; ++rb) {
                                    if (rb == crb) {
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk) {
                                            const T a = af[32 * kk];
#pragma unroll
                                            for (int ct = 0; ct < 2; ++ct)
                                                if (ct < nct) {
                                                    const T bq = CACHE_B ? bfr[CACHE_B ? kk : 0][ct] : Uk[((kk * 2 + ct) << 5) + lane];
                                                    Num<T>::mma(C[bi][rb][ct][0], C[bi][rb][ct][1], a, bq);
                                                }
                                        }
                                    }
                                }
                                // block row k + 1 is final after step k: publish U[k + 1, j]
                                if (crb == 1 && warp + bi * NW == k + 1 && k + 1 < jj)
                                    convert(C[bi], LIg + (long long)(k + 1) * 256, false, nct, BC + (k + 1) * 256,
                                            isrhs ? nullptr : Ug + ((long long)(k + 1) * nb + j) * 256);
                            }
                        }
                        cstage = (cstage + 1 == NSTAGE) ? 0 : cstage + 1;
                        next(ck, cbi, crb);
                    }
                    __syncthreads();
                }
                cpa_wait<0>();
            }
            if (isrhs) break;

            // ---- PANEL: rows at positions >= 16 j -> shared memory (swizzled), LU with partial pivoting ----
            T* PB = BC + (size_t)j * 256;
            const int rows = R - 16 * j;
#pragma unroll
            for (int bi = 0; bi < RBW; ++bi) {
                const int b = warp + bi * NW;
                if (b >= j && b < nb) {
#pragma unroll
                    for (int rb8 = 0; rb8 < 2; ++rb8)
#pragma unroll
                        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                PB[mphys(16 * (b - j) + 8 * rb8 + g, 8 * ct + 2 * t + e, 16)] = C[bi][rb8][ct][e];
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
                    // rows 8 .. rows-1, columns 8..15 += (-L21) U12 : one 8 x 8 tile per row tile, dealt to the warps
                    const int ntiles = rows / 8 - 1;
                    const T b0 = PB[8 + fop.b0], b1 = PB[8 + fop.b1];
                    for (int ti = warp; ti < ntiles; ti += NW) {
                        T* rowp = PB + (8 * (1 + ti) + g) * 16;
                        const T a0 = rowp[fop.a0], a1 = rowp[fop.a1];
                        T v0 = rowp[8 + fop.c0], v1 = rowp[8 + fop.c1];
                        Num<T>::mma(v0, v1, a0, b0);
                        Num<T>::mma(v0, v1, a1, b1);
                        rowp[8 + fop.c0] = v0; rowp[8 + fop.c1] = v1;
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
                    for (int i = 0; i < 16; ++i) LIg[(long long)j * 256 + afrag_off(i, c)] = x[i];
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
        }

        // ---- back substitution: x_k = U_kk^-1 (y_k - sum_{k' > k} U[k, k'] x_k'), blocks from the last to the first ----
        // y_b sits in the C registers of the warp that owns block b (left there by the right-hand-side block column).
        for (int k = nb - 1; k >= 0; --k) {
#pragma unroll
            for (int bi = 0; bi < RBW; ++bi)
                if (warp + bi * NW == k) convert(C[bi], UIg + (long long)k * 256, true, mct, BC + k * 256, nullptr);
            __syncthreads();
            const T* Xk = BC + k * 256;
#pragma unroll
            for (int bi = 0; bi < RBW; ++bi) {
                const int b = warp + bi * NW;
                if (b < k) {
                    const T* ub = Ug + ((long long)b * nb + k) * 256 + lane;
#pragma unroll
                    for (int rb8 = 0; rb8 < 2; ++rb8)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const T a = Num<T>::neg(ub[(rb8 * 4 + kk) << 5]);
#pragma unroll
                            for (int ct = 0; ct < 2; ++ct)
                                if (ct < mct) Num<T>::mma(C[bi][rb8][ct][0], C[bi][rb8][ct][1], a, Xk[((kk * 2 + ct) << 5) + lane]);
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
// Version 2 of the kernel body.  Same algorithm and storage as above, restructured after the first ncu capture
// (profiles/r02_sweep_left_v1_ncu.md: 1.23 M warp instructions per point of which 8 % DMMA; 24 % of the stall samples at the
// per-step CTA barrier of the chain):
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
                    T* dst = ringw + stage * 128 + lane;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) Num<T>::cp_async(dst + 32 * kk, src + 4 * kk);
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


struct LeftGeom { int R, nb, NW, RBW, MINB; size_t smem, slot_elems; int cfg; };

// Geometry per size (cfg): warps per CTA x owned 16-row blocks per warp must cover R / 16 blocks.
//   1: 4 warps x 2 blocks  (R <= 128)      2: 8 warps x 2 (R <= 256)      3: 16 warps x 2 (R <= 512)
//   4: 8 warps x 4 (R <= 512; real twin: two CTAs per SM)
template <typename T>
LeftGeom left_geom(int r, int m) {
    LeftGeom gm;
    gm.R = (r + 15) / 16 * 16;
    gm.nb = gm.R / 16;
    int cfg = gm.nb <= 8 ? 1 : (gm.nb <= 16 ? 2 : (sizeof(T) == 8 ? 4 : 3));
    if (const char* e = getenv("MF_LEFT_CFG")) { const int c = atoi(e); if (c >= 1 && c <= 4) cfg = c; }
    switch (cfg) {
        case 1:  gm.NW = 4; gm.RBW = 2; break;
        case 2:  gm.NW = 8; gm.RBW = 2; break;
        case 3:  gm.NW = 16; gm.RBW = 2; break;
        default: gm.NW = 8; gm.RBW = 4; break;
    }
    if (gm.NW * gm.RBW < gm.nb) {                                 // a hand-picked geometry that does not cover R: fall back
        cfg = gm.nb <= 8 ? 1 : (gm.nb <= 16 ? 2 : (sizeof(T) == 8 ? 4 : 3));
        gm.NW = cfg == 1 ? 4 : (cfg == 3 ? 16 : 8); gm.RBW = cfg == 4 ? 4 : 2;
    }
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

template <typename T, int NW, int RBW, int MINB>
int launch_left(SweepParamsL<T> p, const LeftGeom& gm, size_t ws_bytes, cudaStream_t stream) {
    // MF_LEFT_V1 selects the first version of the kernel body (kept for A/B measurements)
    static const bool v1 = getenv("MF_LEFT_V1") != nullptr;
    auto kern = v1 ? sweep_left_kernel<T, NW, RBW, MINB, 2> : sweep_left2_kernel<T, NW, 2 * RBW, MINB, 2>;
    int per_sm = 0;
    if (int rc = left_occupancy(kern, NW * 32, gm, &per_sm)) return rc;
    if (per_sm < 1) MF_FAIL_ARG(7, "left-looking sweep does not fit on an SM for this (r, m)");
    const size_t slot = sizeof(T) * gm.slot_elems;
    long long grid = (long long)mf_num_sms() * per_sm;
    if (grid > p.F) grid = p.F;
    if ((long long)(ws_bytes / slot) < grid) grid = (long long)(ws_bytes / slot);
    if (grid < 1 || !p.ws) MF_FAIL_ARG(21, "workspace too small for the left-looking blocked sweep (see mf_sweep_ws_bytes)");
    p.ws_stride = (long long)gm.slot_elems;
    kern<<<(unsigned)grid, NW * 32, gm.smem, stream>>>(p, gm.R);
    MF_CHECK_LAUNCH();
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
    long long grid = (long long)mf_num_sms() * 4; if (grid > F) grid = F; if (grid < 1) grid = 1;   // at most 4 CTAs per SM in any geometry
    return sizeof(T) * gm.slot_elems * (size_t)grid;
}

}  // namespace

bool sweep_left_supports_c128(int r, int m) { return left_supports<cplx>(r, m); }
bool sweep_left_supports_f64(int r, int m) { return left_supports<double>(r, m); }
size_t sweep_left_ws_bytes_c128(int r, int m, long long F) { return left_ws_bytes<cplx>(r, m, F); }
size_t sweep_left_ws_bytes_f64(int r, int m, long long F) { return left_ws_bytes<double>(r, m, F); }

int sweep_left_launch_c128(const SweepParams& q, size_t ws_bytes, cudaStream_t stream) {
    SweepParamsL<cplx> p;
    p.A0 = q.A0; p.A1 = q.A1; p.A2 = q.A2; p.lda = q.lda; p.Br = q.Br; p.ldb = q.ldb; p.r = q.r; p.m = q.m;
    p.c0 = q.c0; p.c1 = q.c1; p.c2 = q.c2; p.cb = q.cb; p.zs = q.zs; p.F = q.F; p.X = q.X; p.S = q.S; p.info = q.info;
    p.ws = q.ws; p.ws_stride = 0;
    return dispatch_left<cplx>(p, ws_bytes, stream);
}

int sweep_left_launch_f64(const double* A0, const double* A1, const double* A2, long long lda, const double* Br, long long ldb, int r, int m,
                          const double* c0, const double* c1, const double* c2, const double* cb, const double* zs, long long F,
                          double* X, cplx* S, int* info, void* ws, size_t ws_bytes, cudaStream_t stream) {
    SweepParamsL<double> p;
    p.A0 = A0; p.A1 = A1; p.A2 = A2; p.lda = lda; p.Br = Br; p.ldb = ldb; p.r = r; p.m = m;
    p.c0 = c0; p.c1 = c1; p.c2 = c2; p.cb = cb; p.zs = zs; p.F = F; p.X = X; p.S = S; p.info = info;
    p.ws = (double*)ws; p.ws_stride = 0;
    return dispatch_left<double>(p, ws_bytes, stream);
}
