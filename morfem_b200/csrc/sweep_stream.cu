// Blocked sweep for large reduced models (113 <= r <= 512): one CTA per frequency point (several CTAs per SM, see
// stream_geom), augmented matrix [A(t) | cb(t) Br] in a per-CTA slot of a global workspace, two-level blocked
// right-looking LU.
//
// Outer step (NBO = 16 columns; 32 in the one-CTA-per-SM geometry):
//   1. the (R - col0) x NBO outer panel is copied into shared memory (same XOR-swizzled layout as sweep_blocked.cu);
//   2. it is factored there in inner panels of 8 columns: ALL warps take part in the pivot search (rows split over the
//      CTA, one CTA barrier per column, candidates exchanged through shared memory), then the row exchanges, the
//      triangular solve and the DMMA update of the rest of the outer panel, exactly as in sweep_blocked.cu;
//   3. the composite row permutation of the NBO exchanges is formed once and applied to the trailing columns as a
//      gather (all loads before all stores) instead of NBO dependent row swaps through L2;
//   4. per chunk of ST_CW trailing columns: U12 = L11^-1 A12 by a blocked forward substitution (DMMA + 8 x 8 in-block
//      solves) in shared memory, written back as final U rows, then A22 -= L21 U12 with the C tiles streamed
//      global -> registers -> global, A fragments from the shared outer panel, B fragments from the shared U12 chunk.
// FUSE: there is no assembly pass.  The FIRST outer step reads its operands straight from the L2-resident reduced
// operators (aug(i, j) = c0 A0 + c1 A1 + c2 A2 | cb Br): the panel load, the rows that land in U12 and the C tiles
// of the first trailing update (row idx[x] of the operators = the row the exchanges brought to position x), so the
// matrix is written to its slot once, already updated (saves one write + one read of the slot per point).
// Traffic per point at r = 256: ~6 MB (NBO = 32) / ~11 MB (NBO = 16) for 47.9 MFLOP.
// Back substitution is row oriented (coalesced rows of U from L2, solution in shared memory); the impedance matrix
// goes to S and gsm_finish_kernel completes the S-parameter algebra (test_helpers.py:11-14).
// Reference semantics: implementation.py:468-480, :526-533 (lu_factor / lu_solve of the symmetrised system matrix).
#include <stdlib.h>
#include "sweep_blocked.cuh"

namespace {

constexpr int ST_MAXMOVED = 64;              // rows touched by the composite permutation of one outer panel (<= 2 NBO)
constexpr int ST_MAXNW = 8;
#ifndef MF_STREAM_SMALL_R
#define MF_STREAM_SMALL_R 128
#endif
#ifndef MF_STREAM_DEFAULT_FUSE
#define MF_STREAM_DEFAULT_FUSE true
#endif                  // most warps per CTA of any geometry (sizes the candidate exchange)

struct CandKey { double v; int pos; int pad; };

// ---- inner panel factorisation by the whole CTA ---------------------------------------------------------------
// Rows row0 .. rows-1, columns row0 .. row0+7 of the shared outer panel PB (leading dimension LDp).  Each thread
// holds SL rows in registers.  Per column: warp-level arg-max as in sweep_blocked.cu, the warp winners publish
// (magnitude, position, finished row with the reciprocal pivot) to shared memory, ONE CTA barrier, every thread
// picks the same global winner from the NW candidates.  Multipliers are stored negated.  pvl[j] = position (local
// row of PB) the j-th pivot row came from.  Ends with all rows written back; the caller synchronises.
template <int SL, int ST_NW>
__device__ __forceinline__ void panel_factor_mw(cplx* PB, const int LDp, const int rows, const int row0, const int tid,
                                                CandKey* candk, cplx* candrow, int* pvl, int* info_sh, const int info_base) {
    constexpr int ST_NT = ST_NW * 32;
    const int lane = tid & 31, warp = tid >> 5;
    cplx a[SL][8];
    int pos[SL];
    bool act[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int row = row0 + tid + ST_NT * s;
        pos[s] = row;
        act[s] = row < rows;
        const int sw = swz(row & 7);
        const cplx* src = PB + row * LDp + row0;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[s][c] = act[s] ? src[c ^ sw] : cmake(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int T = row0 + j;
        double vb = act[0] ? fabs(a[0][j].x) + fabs(a[0][j].y) : -1.0;
        int pbest = pos[0], bs = 0;
        cplx cand = a[0][j];
#pragma unroll
        for (int s = 1; s < SL; ++s) {
            const double v = act[s] ? fabs(a[s][j].x) + fabs(a[s][j].y) : -1.0;
            if (v > vb || (v == vb && pos[s] < pbest)) { vb = v; pbest = pos[s]; bs = s; cand = a[s][j]; }
        }
        cplx rc = cmake(0.0, 0.0);
        if (vb > 0.0) rc = (cand.y == 0.0) ? cmake(1.0 / cand.x, 0.0) : crecip2(cand);
        const int hi = __double2hiint(vb);
        const int hmax = __reduce_max_sync(FULL, hi);
        bool own = (hi == hmax);
        if (__popc(__ballot_sync(FULL, own)) != 1) {
            const unsigned lo = (unsigned)__double2loint(vb);
            const unsigned lmax = __reduce_max_sync(FULL, own ? lo : 0u);
            own = own && (lo == lmax);
            if (__popc(__ballot_sync(FULL, own)) != 1) {
                const int pmin = __reduce_min_sync(FULL, own ? pbest : 0x7fffffff);
                own = own && (pbest == pmin);
            }
        }
        CandKey* ck = candk + (j & 1) * ST_MAXNW;
        cplx* cr = candrow + (j & 1) * ST_MAXNW * 8;
        if (own) {                                   // this warp's candidate: key and finished row
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
        if (lane < ST_NW) {
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
        if (own && warp == gw) {                     // this lane held the pivot row: retire the slot
#pragma unroll
            for (int s = 0; s < SL; ++s) if (s == bs) act[s] = false;
        }
        const unsigned crs = (unsigned)__cvta_generic_to_shared(cr + gw * 8);
        cplx rcp, u[8];
        asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rcp.x), "=d"(rcp.y) : "r"(crs + 16u * j));
#pragma unroll
        for (int c = j + 1; c < 8; ++c)
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u[c].x), "=d"(u[c].y) : "r"(crs + 16u * c));
        if (tid < 8) {                               // the pivot row goes to its final place; (T & 7) == j
            cplx v;
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(crs + 16u * tid));
            PB[T * LDp + row0 + (tid ^ swz(j))] = v;
        }
        if (tid == 8) { pvl[j] = P; if (!(gv > 0.0) && *info_sh == 0) *info_sh = info_base + T + 1; }
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const cplx nl = cmake(-(a[s][j].x * rcp.x - a[s][j].y * rcp.y), -(a[s][j].x * rcp.y + a[s][j].y * rcp.x));
            a[s][j] = nl;
#pragma unroll
            for (int c = j + 1; c < 8; ++c) cfma(a[s][c], nl, u[c]);
            if (pos[s] == T) pos[s] = P;
        }
    }
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        if (act[s]) {
            const int q = pos[s];
            const int sw = swz(q & 7);
            cplx* dst = PB + q * LDp + row0;
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c ^ sw] = a[s][c];
        }
    }
}

__device__ __forceinline__ FragOff make_fragoff(int lane, int LD) {
    FragOff fo;
    const int g = lane >> 2, t = lane & 3, sg = swz(g);
    fo.g = g;
    fo.a0 = t ^ sg; fo.a1 = (4 + t) ^ sg;
    fo.c0 = (2 * t) ^ sg; fo.c1 = (2 * t + 1) ^ sg;
    fo.b0 = t * LD + (g ^ swz(t)); fo.b1 = (4 + t) * LD + (g ^ swz(4 + t));
    return fo;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// NBO = outer panel width (32 or 16), ST_NW = warps per CTA, ST_CW = trailing columns per chunk, RMAX = largest padded
// size the instance handles (sizes the per-thread register arrays), MINB = CTAs per SM the register budget is set for.
template <int NBO, int ST_NW, int ST_CW, int RMAX, int MINB, bool FUSE>
__global__ void __launch_bounds__(ST_NW * 32, MINB) sweep_stream_kernel(SweepParams p, int R, int LD) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int ST_NT = ST_NW * 32;
    constexpr int SL = RMAX / ST_NT;                             // rows per thread in the inner panel factorisation
    constexpr int NCBP = NBO / 8;
    const int r = p.r, m = p.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3, sg = swz(g);

    cplx* PB = reinterpret_cast<cplx*>(smem_raw);                 // R x NBO outer panel (later: the solution, R x m)
    cplx* UB = PB + (size_t)R * NBO;                              // NBO x ST_CW chunk of U12
    cplx* candrow = UB + NBO * ST_CW;                             // 2 x NW x 8
    CandKey* candk = reinterpret_cast<CandKey*>(candrow + 2 * ST_MAXNW * 8);   // 2 x NW
    int* idx = reinterpret_cast<int*>(candk + 2 * ST_MAXNW);         // R: composite permutation
    int* mv_dst = idx + R;                                        // ST_MAXMOVED
    int* mv_src = mv_dst + ST_MAXMOVED;                           // ST_MAXMOVED
    int* lp = mv_src + ST_MAXMOVED;                               // NBO local pivot positions of the outer panel
    int* nmoved = lp + NBO;                                       // 1
    int* info_sh = nmoved + 1;                                    // 1

    cplx* A = p.ws + (long long)blockIdx.x * p.ws_stride;         // this CTA's matrix slot (R x LD, plain row-major)
    const FragOff fop = make_fragoff(lane, NBO);
    const bool hasA0 = p.A0 != nullptr, hasA1 = p.A1 != nullptr, hasA2 = p.A2 != nullptr;

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        // ---- assemble [A(t) | cb Br] into the slot, identity on the padded diagonal ----
        // FUSE: no assembly pass -- the first outer step reads its operands straight from the operators (aug below)
        constexpr int RPLA = RMAX / 32;                           // row elements per lane (R <= 32 RPLA)
        constexpr int RGA = (RMAX <= 256 && MINB == 1) ? 2 : 1;   // operator rows per warp iteration (all their loads in flight)
        // element (i, j) of the augmented matrix [A(t) | cb Br], identity on the padded diagonal
        auto aug = [&](const int i, const int j) -> cplx {
            cplx v = cmake(0.0, 0.0);
            if (j < R) {
                if (i < r && j < r) {
                    const long long off = (long long)i * p.lda + j;
                    if (hasA0) { const cplx x = __ldg(p.A0 + off); v.x = c0 * x.x; v.y = c0 * x.y; }
                    if (hasA1) { const cplx x = __ldg(p.A1 + off); v.x = fma(c1, x.x, v.x); v.y = fma(c1, x.y, v.y); }
                    if (hasA2) { const cplx x = __ldg(p.A2 + off); v.x = fma(c2, x.x, v.x); v.y = fma(c2, x.y, v.y); }
                } else if (i == j) v.x = 1.0;
            } else if (i < r && j - R < m) {
                const cplx x = __ldg(p.Br + (long long)i * p.ldb + (j - R));
                v.x = cb * x.x; v.y = cb * x.y;
            }
            return v;
        };
        if (!FUSE)
        for (int i0 = warp * RGA; i0 < R; i0 += ST_NW * RGA) {
            cplx v[RGA][RPLA];
#pragma unroll
            for (int g2 = 0; g2 < RGA; ++g2) {
                const int i = i0 + g2;
                const long long rowoff = (long long)i * p.lda;
#pragma unroll
                for (int q = 0; q < RPLA; ++q) {
                    const int j = lane + 32 * q;
                    v[g2][q] = cmake(0.0, 0.0);
                    if (i < r && j < r) {
                        if (hasA0) { const cplx x = __ldg(p.A0 + rowoff + j); v[g2][q].x = c0 * x.x; v[g2][q].y = c0 * x.y; }
                        if (hasA1) { const cplx x = __ldg(p.A1 + rowoff + j); v[g2][q].x = fma(c1, x.x, v[g2][q].x); v[g2][q].y = fma(c1, x.y, v[g2][q].y); }
                        if (hasA2) { const cplx x = __ldg(p.A2 + rowoff + j); v[g2][q].x = fma(c2, x.x, v[g2][q].x); v[g2][q].y = fma(c2, x.y, v[g2][q].y); }
                    } else if (i == j) v[g2][q].x = 1.0;          // identity on the padded diagonal
                }
            }
#pragma unroll
            for (int g2 = 0; g2 < RGA; ++g2) {
                const int i = i0 + g2;
                if (i < R) {
                    cplx* Arow = A + (long long)i * LD;
#pragma unroll
                    for (int q = 0; q < RPLA; ++q) { const int j = lane + 32 * q; if (j < R) Arow[j] = v[g2][q]; }
                    for (int j = lane; j < LD - R; j += 32) {
                        cplx b = cmake(0.0, 0.0);
                        if (i < r && j < m) { const cplx x = __ldg(p.Br + (long long)i * p.ldb + j); b.x = cb * x.x; b.y = cb * x.y; }
                        Arow[R + j] = b;
                    }
                }
            }
        }
        if (tid == 0) *info_sh = 0;
        __syncthreads();

        // ---- outer steps ----
        for (int col0 = 0; col0 < R; col0 += NBO) {
            const int rows = R - col0;                           // rows (and local row count) of the outer panel
            // 1. outer panel -> shared memory
            const bool first = FUSE && col0 == 0;
            {                                                    // 32 / NBO rows per warp instruction, four of those in flight
                constexpr int RW = 32 / NBO;
                const int jl = lane % NBO, il = lane / NBO;
                for (int i0 = warp * RW * 4; i0 < rows; i0 += ST_NW * RW * 4) {
                    cplx v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i0 + q * RW + il;
                        if (i < rows) v[q] = first ? aug(i, jl) : A[(long long)(col0 + i) * LD + col0 + jl];
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i0 + q * RW + il;
                        if (i < rows) PB[i * NBO + (jl & ~7) + ((jl & 7) ^ swz(i & 7))] = v[q];
                    }
                }
            }
            for (int i = tid; i < rows; i += ST_NT) idx[i] = i;
            if (tid == 0) *nmoved = 0;
            __syncthreads();
            // 2. factor it in inner panels of 8 columns
#pragma unroll 1
            for (int ip = 0; ip < NCBP; ++ip) {
                const int row0 = 8 * ip;
                int* pvl = lp + row0;
                if (SL > 1 && rows - row0 > ST_NT) panel_factor_mw<SL, ST_NW>(PB, NBO, rows, row0, tid, candk, candrow, pvl, info_sh, col0);
                else panel_factor_mw<1, ST_NW>(PB, NBO, rows, row0, tid, candk, candrow, pvl, info_sh, col0);
                __syncthreads();
                // the exchanges also apply to the multipliers of the earlier inner panels (columns < row0): unlike in
                // sweep_blocked.cu, L21 of the whole outer panel is an operand of the trailing update
                if (tid < row0) {
                    const int c = tid, cbase = c & ~7, cin = c & 7;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int P = pvl[j], T = row0 + j;
                        if (P != T) {
                            cplx* x = PB + T * NBO + cbase + (cin ^ swz(j));
                            cplx* y = PB + P * NBO + cbase + (cin ^ swz(P & 7));
                            const cplx tmp = *x; *x = *y; *y = tmp;
                        }
                    }
                }
                if (ip + 1 < NCBP) {
                    for (int c = row0 + 8 + tid; c < NBO; c += ST_NT) stepb_column(PB, NBO, row0, c, pvl);
                    __syncthreads();
                    const int nrb = rows / 8 - (ip + 1), ncb = NCBP - (ip + 1);
                    const int ntiles = nrb * ncb;
                    update_tiles(PB, NBO, row0, ip + 1, nrb, ip + 1, (ntiles * warp) / ST_NW, (ntiles * (warp + 1)) / ST_NW, fop);
                    __syncthreads();
                }
            }
            __syncthreads();
            // U11 (with the reciprocal pivots) is final: write the NBO x NBO block back
            for (int e = tid; e < NBO * NBO; e += ST_NT) {
                const int i = e / NBO, j = e - i * NBO;
                A[(long long)(col0 + i) * LD + col0 + j] = PB[mphys(i, j, NBO)];
            }
            // 3. composite permutation of the NBO exchanges: idx[x] = local row that ends up at position x
            if (tid == 0) {
                for (int j = 0; j < NBO; ++j) { const int P = lp[j]; const int tmp = idx[j]; idx[j] = idx[P]; idx[P] = tmp; }
            }
            __syncthreads();
            for (int x = tid; x < rows; x += ST_NT) {
                if (x < NBO || idx[x] != x) { const int e = atomicAdd(nmoved, 1); mv_dst[e] = x; mv_src[e] = idx[x]; }
            }
            __syncthreads();
            const int ne = *nmoved;
            // 4. trailing columns in chunks
            for (int cc0 = col0 + NBO; cc0 < LD; cc0 += ST_CW) {
                const int cw = min(ST_CW, LD - cc0);             // multiple of 8
                {   // 4a. gather the moved rows (all loads, barrier, all stores); rows landing in the block go to UB
                    constexpr int NEG = ST_NT / ST_CW;             // row groups working side by side
                    const int c = tid % ST_CW, eg = tid / ST_CW;
                    constexpr int MAXMV = 2 * NBO;                 // rows the NBO exchanges can touch
                    cplx vals[MAXMV / NEG];
#pragma unroll
                    for (int i = 0; i < MAXMV / NEG; ++i) {
                        const int e = eg + NEG * i;
                        vals[i] = cmake(0.0, 0.0);
                        if (first) { if (e < ne && c < cw && mv_dst[e] < NBO) vals[i] = aug(mv_src[e], cc0 + c); }
                        else if (e < ne && c < cw) vals[i] = A[(long long)(col0 + mv_src[e]) * LD + cc0 + c];
                    }
                    __syncthreads();
#pragma unroll
                    for (int i = 0; i < MAXMV / NEG; ++i) {
                        const int e = eg + NEG * i;
                        if (e < ne && c < cw) {
                            const int d = mv_dst[e];
                            if (d < NBO) UB[mphys(d, c, ST_CW)] = vals[i];
                            else if (!first) A[(long long)(col0 + d) * LD + cc0 + c] = vals[i];
                        }
                    }
                }
                __syncthreads();
                // 4b. U12 = L11^-1 A12: blocked forward substitution in UB (L11 = -stored multipliers of PB rows < NBO)
                const FragOff fou = make_fragoff(lane, ST_CW);
#pragma unroll 1
                for (int bj = 0; bj < NCBP; ++bj) {
                    if (bj > 0) {
                        for (int ct = warp; ct < cw / 8; ct += ST_NW) {
                            cplx* pc0 = UB + (8 * bj + g) * ST_CW + 8 * ct + fou.c0;
                            cplx* pc1 = UB + (8 * bj + g) * ST_CW + 8 * ct + fou.c1;
                            const cplx v0 = *pc0, v1 = *pc1;
                            double cre0 = v0.x, cre1 = v1.x, cim0 = v0.y, cim1 = v1.y;
                            for (int bi = 0; bi < bj; ++bi) {
                                const cplx* arow = PB + (8 * bj + g) * NBO + 8 * bi;
                                const cplx a0 = arow[fop.a0], a1 = arow[fop.a1];
                                const cplx b0 = UB[8 * bi * ST_CW + 8 * ct + fou.b0], b1 = UB[8 * bi * ST_CW + 8 * ct + fou.b1];
                                dmma884(cre0, cre1, a0.x, b0.x); dmma884(cim0, cim1, a0.x, b0.y);
                                dmma884(cre0, cre1, -a0.y, b0.y); dmma884(cim0, cim1, a0.y, b0.x);
                                dmma884(cre0, cre1, a1.x, b1.x); dmma884(cim0, cim1, a1.x, b1.y);
                                dmma884(cre0, cre1, -a1.y, b1.y); dmma884(cim0, cim1, a1.y, b1.x);
                            }
                            *pc0 = cmake(cre0, cim0); *pc1 = cmake(cre1, cim1);
                        }
                        __syncthreads();
                    }
                    if (tid < cw) {                              // in-block unit-lower solve, one thread per column
                        const int c = tid, cbase = c & ~7, cin = c & 7;
                        cplx* colp = UB + 8 * bj * ST_CW + cbase;
                        cplx u[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) u[j] = colp[j * ST_CW + (cin ^ swz(j))];
#pragma unroll
                        for (int j = 1; j < 8; ++j) {
                            const cplx* lrow = PB + (8 * bj + j) * NBO + 8 * bj;
                            const int sw = swz(j);
#pragma unroll
                            for (int i = 0; i < j; ++i) cfma(u[j], lrow[i ^ sw], u[i]);
                        }
#pragma unroll
                        for (int j = 1; j < 8; ++j) colp[j * ST_CW + (cin ^ swz(j))] = u[j];
                    }
                    __syncthreads();
                }
                // 4c. final U rows of this chunk -> global
                for (int e = tid; e < NBO * cw; e += ST_NT) {
                    const int j = e / cw, c = e - j * cw;
                    A[(long long)(col0 + j) * LD + cc0 + c] = UB[mphys(j, c, ST_CW)];
                }
                // 4d. A22 += (-L21) U12.  Work unit = (row block, batch of <= 4 column tiles), dealt round-robin to the warps;
                // two register buffers: the C tiles of the next unit are in flight while the DMMAs of the current one run.
                {
                    const int nct = cw / 8, nbt = (nct + 3) >> 2;
                    const int nunits = (rows / 8 - NCBP) * nbt;
                    auto issue = [&](const int u, cplx (&v)[4][2]) {
                        const int rbi = u / nbt, ct0 = (u - rbi * nbt) * 4;
                        const int row = 8 * (NCBP + rbi) + g;    // local row of the outer panel
                        if (first) {                             // first step: the row the exchanges brought here, from the operators
                            const int srow = idx[row];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (ct0 + q < nct) { v[q][0] = aug(srow, cc0 + 2 * t + 8 * (ct0 + q)); v[q][1] = aug(srow, cc0 + 2 * t + 8 * (ct0 + q) + 1); }
                        } else {
                            const cplx* crow = A + (long long)(col0 + row) * LD + cc0 + 2 * t;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (ct0 + q < nct) { v[q][0] = crow[8 * (ct0 + q)]; v[q][1] = crow[8 * (ct0 + q) + 1]; }
                        }
                    };
                    auto compute = [&](const int u, cplx (&v)[4][2]) {
                        const int rbi = u / nbt, ct0 = (u - rbi * nbt) * 4;
                        const int row = 8 * (NCBP + rbi) + g;
                        const cplx* arow = PB + row * NBO;
                        cplx af[NBO / 4];
#pragma unroll
                        for (int kk = 0; kk < NBO / 4; ++kk) af[kk] = arow[(4 * kk + t) ^ sg];      // (4kk + t) ^ swz(g): same 8-block
                        cplx* crow = A + (long long)(col0 + row) * LD + cc0 + 2 * t;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (ct0 + q < nct) {
                                const int ct = ct0 + q;
                                double cre0 = v[q][0].x, cre1 = v[q][1].x, cim0 = v[q][0].y, cim1 = v[q][1].y;
#pragma unroll
                                for (int kk = 0; kk < NBO / 4; ++kk) {
                                    const cplx b = UB[(4 * kk + t) * ST_CW + 8 * ct + (g ^ swz((4 * kk + t) & 7))];
                                    dmma884(cre0, cre1, af[kk].x, b.x); dmma884(cim0, cim1, af[kk].x, b.y);
                                    dmma884(cre0, cre1, -af[kk].y, b.y); dmma884(cim0, cim1, af[kk].y, b.x);
                                }
                                crow[8 * ct] = cmake(cre0, cim0); crow[8 * ct + 1] = cmake(cre1, cim1);
                            }
                        }
                    };
                    if (MINB == 1) {                             // register budget of a lone CTA: two buffers
                        cplx va[4][2], vb[4][2];
                        int u = warp;
                        if (u < nunits) issue(u, va);
                        for (; u < nunits; u += 2 * ST_NW) {
                            if (u + ST_NW < nunits) issue(u + ST_NW, vb);
                            compute(u, va);
                            if (u + 2 * ST_NW < nunits) issue(u + 2 * ST_NW, va);
                            if (u + ST_NW < nunits) compute(u + ST_NW, vb);
                        }
                    } else {                                     // several CTAs per SM hide the load latency for each other:
                        // one row block per warp iteration, A fragments kept in registers across its column tiles
                        for (int rb = NCBP + warp; rb < rows / 8; rb += ST_NW) {
                            const cplx* arow = PB + (8 * rb + g) * NBO;
                            cplx af[NBO / 4];
#pragma unroll
                            for (int kk = 0; kk < NBO / 4; ++kk) af[kk] = arow[(4 * kk + t) ^ sg];
                            cplx* crow = A + (long long)(col0 + 8 * rb + g) * LD + cc0 + 2 * t;
                            const int srow = first ? idx[8 * rb + g] : 0;
                            for (int ct0 = 0; ct0 < nct; ct0 += 4) {
                                cplx v[4][2];
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if (ct0 + q < nct) {
                                        if (first) { v[q][0] = aug(srow, cc0 + 2 * t + 8 * (ct0 + q)); v[q][1] = aug(srow, cc0 + 2 * t + 8 * (ct0 + q) + 1); }
                                        else { v[q][0] = crow[8 * (ct0 + q)]; v[q][1] = crow[8 * (ct0 + q) + 1]; }
                                    }
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (ct0 + q < nct) {
                                        const int ct = ct0 + q;
                                        double cre0 = v[q][0].x, cre1 = v[q][1].x, cim0 = v[q][0].y, cim1 = v[q][1].y;
#pragma unroll
                                        for (int kk = 0; kk < NBO / 4; ++kk) {
                                            const cplx b = UB[(4 * kk + t) * ST_CW + 8 * ct + (g ^ swz((4 * kk + t) & 7))];
                                            dmma884(cre0, cre1, af[kk].x, b.x); dmma884(cim0, cim1, af[kk].x, b.y);
                                            dmma884(cre0, cre1, -af[kk].y, b.y); dmma884(cim0, cim1, af[kk].y, b.x);
                                        }
                                        crow[8 * ct] = cmake(cre0, cim0); crow[8 * ct + 1] = cmake(cre1, cim1);
                                    }
                                }
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }

        // ---- back substitution U x = y in blocks of 8 rows; y / x (R x m) live in shared memory (PB area) ----
        // Per block: lanes 0..m-1 of warp 0 solve the 8 x 8 triangular system (diagonal block staged in shared memory),
        // then every thread updates ITS row(s) above the block with the 8 new unknowns.  The 128-byte row segments of U and
        // the next diagonal block are prefetched into registers one block ahead, so the dependent chain per block is the
        // small solve plus two CTA barriers -- no DRAM round trip.
        cplx* xs = PB;                                            // R x m, xs[k * m + c]
        cplx* dblk = UB;                                          // 8 x 8 diagonal block (row-major)
        constexpr int RPT = RMAX / ST_NT;                         // rows per thread (R <= ST_NT RPT)
        for (int e = tid; e < R * m; e += ST_NT) { const int i = e / m, c = e - i * m; xs[e] = A[(long long)i * LD + R + c]; }
        cplx un[RPT][8], dn = cmake(0.0, 0.0);
        {
            const int kb8 = R - 8;
#pragma unroll
            for (int s = 0; s < RPT; ++s) {
                const int i = tid + ST_NT * s;
                const cplx* src = A + (long long)i * LD + kb8;
#pragma unroll
                for (int j = 0; j < 8; ++j) un[s][j] = (i < kb8) ? src[j] : cmake(0.0, 0.0);
            }
            if (tid < 64) dn = A[(long long)(kb8 + (tid >> 3)) * LD + kb8 + (tid & 7)];
        }
        for (int kb8 = R - 8; kb8 >= 0; kb8 -= 8) {
            cplx uc[RPT][8];
#pragma unroll
            for (int s = 0; s < RPT; ++s)
#pragma unroll
                for (int j = 0; j < 8; ++j) uc[s][j] = un[s][j];
            if (tid < 64) dblk[tid] = dn;
            if (kb8 >= 8) {                                       // prefetch the next block (one block above)
                const int nb8 = kb8 - 8;
#pragma unroll
                for (int s = 0; s < RPT; ++s) {
                    const int i = tid + ST_NT * s;
                    const cplx* src = A + (long long)i * LD + nb8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) un[s][j] = (i < nb8) ? src[j] : cmake(0.0, 0.0);
                }
                if (tid < 64) dn = A[(long long)(nb8 + (tid >> 3)) * LD + nb8 + (tid & 7)];
            }
            __syncthreads();                                      // dblk and the y values of this block are in place
            if (tid < m) {
                const int c = tid;
                cplx x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = xs[(kb8 + j) * m + c];
#pragma unroll
                for (int j = 7; j >= 0; --j) {
#pragma unroll
                    for (int jj = 7; jj > j; --jj) cfms(x[j], dblk[j * 8 + jj], x[jj]);   // newest unknown last (ztrsm order)
                    x[j] = cmul(x[j], dblk[j * 8 + j]);           // reciprocal pivot on the diagonal
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) xs[(kb8 + j) * m + c] = x[j];
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < RPT; ++s) {
                const int i = tid + ST_NT * s;
                if (i < kb8) {
                    for (int c = 0; c < m; ++c) {
                        cplx y = xs[i * m + c];
#pragma unroll
                        for (int j = 0; j < 8; ++j) cfms(y, uc[s][j], xs[(kb8 + j) * m + c]);
                        xs[i * m + c] = y;
                    }
                }
            }
        }
        __syncthreads();
        if (p.X) for (int e = tid; e < r * m; e += ST_NT) p.X[pt * (long long)r * m + e] = xs[e];
        if (p.info && tid == 0) p.info[pt] = *info_sh;
        // ---- impedance matrix Z = j zs x^T (cb Br) -> S (finished by gsm_finish_kernel) ----
        if (p.S) {
            for (int e = warp; e < m * m; e += ST_NW) {
                const int a = e / m, b = e - a * m;
                cplx acc = cmake(0.0, 0.0);
                for (int k = lane; k < r; k += 32) cfma(acc, xs[k * m + a], cscale(cb, __ldg(p.Br + (long long)k * p.ldb + b)));
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    acc.x += __shfl_xor_sync(FULL, acc.x, off);
                    acc.y += __shfl_xor_sync(FULL, acc.y, off);
                }
                if (lane == 0) { const double zs = p.zs[pt]; p.S[pt * (long long)m * m + e] = cmake(-zs * acc.y, zs * acc.x); }
            }
        }
        __syncthreads();
    }
}

struct StreamGeom { int R, LD, NBO, NW, CW, MINB; size_t smem; };

// Geometry per size.  Several CTAs per SM, so that one CTA's latency-bound phases (pivot search, gathers, U12 solves)
// run under another's DMMA update: R <= 128: three 4-warp CTAs (1.11 M points/s at r = 128 against 0.79 M for one 8-warp
// CTA), R <= 256: two 8-warp CTAs with 16-column outer panels (231 k points/s at r = 256 against 200 k), R > 256: one
// 8-warp CTA.  MF_STREAM_CFG selects a geometry by hand for measurements: 0 = one CTA per SM with 32-column panels,
// 1 = two 8-warp CTAs, 2 = three 4-warp CTAs, 3 = two 4-warp CTAs.
StreamGeom stream_geom(int r, int m) {
    StreamGeom gm;
    gm.R = (r + 31) / 32 * 32;                                   // multiple of both outer panel widths
    gm.LD = gm.R + (m + 7) / 8 * 8;
    int cfg = gm.R <= MF_STREAM_SMALL_R ? 2 : 1;
    if (const char* e = getenv("MF_STREAM_CFG")) cfg = atoi(e);
    if (gm.R > 256) cfg = 9;
    switch (cfg) {
        case 1:  gm.NBO = 16; gm.NW = 8; gm.CW = 64; gm.MINB = 2; break;
        case 2:  gm.NBO = 16; gm.NW = 4; gm.CW = 32; gm.MINB = 3; break;
        case 3:  gm.NBO = 16; gm.NW = 4; gm.CW = 64; gm.MINB = 2; break;
        case 9:  gm.NBO = 16; gm.NW = 8; gm.CW = 64; gm.MINB = 1; break;
        default: gm.NBO = 32; gm.NW = 8; gm.CW = 64; gm.MINB = 1; break;
    }
    gm.smem = sizeof(cplx) * ((size_t)gm.R * gm.NBO + (size_t)gm.NBO * gm.CW + 2 * ST_MAXNW * 8) + sizeof(CandKey) * 2 * ST_MAXNW
            + sizeof(int) * ((size_t)gm.R + 2 * ST_MAXMOVED + gm.NBO + 2) + 64;
    return gm;
}

template <int NBO, int NW, int CW, int RMAX, int MINB, bool FUSE>
int launch_stream(SweepParams p, const StreamGeom& gm, size_t ws_bytes, cudaStream_t stream) {
    auto kern = sweep_stream_kernel<NBO, NW, CW, RMAX, MINB, FUSE>;
    MF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gm.smem));
    const size_t slot = sizeof(cplx) * (size_t)gm.R * gm.LD;
    long long grid = (long long)mf_num_sms() * MINB;
    if (grid > p.F) grid = p.F;
    if ((long long)(ws_bytes / slot) < grid) grid = (long long)(ws_bytes / slot);
    if (grid < 1 || !p.ws) MF_FAIL_ARG(21, "workspace too small for the streamed blocked sweep (see mf_sweep_ws_bytes)");
    p.ws_stride = (long long)gm.R * gm.LD;
    kern<<<(unsigned)grid, NW * 32, gm.smem, stream>>>(p, gm.R, gm.LD);
    MF_CHECK_LAUNCH();
    if (p.S) return gsm_finish_launch(p.S, p.m, p.F, stream);
    return 0;
}

}  // namespace

bool sweep_stream_supports(int r, int m) {
    if (r < 1 || r > 512 || m < 1 || m > MF_MAX_PORTS) return false;
    const StreamGeom gm = stream_geom(r, m);
    return gm.smem <= 226 * 1024 && (size_t)gm.R * m * sizeof(cplx) <= sizeof(cplx) * (size_t)gm.R * gm.NBO;
}

size_t sweep_stream_ws_bytes(int r, int m, long long F) {
    const StreamGeom gm = stream_geom(r, m);
    long long grid = (long long)mf_num_sms() * gm.MINB; if (grid > F) grid = F; if (grid < 1) grid = 1;
    return sizeof(cplx) * (size_t)gm.R * gm.LD * (size_t)grid;
}

int sweep_stream_launch(const SweepParams& p, size_t ws_bytes, cudaStream_t stream) {
    const StreamGeom gm = stream_geom(p.r, p.m);
    bool fuse = MF_STREAM_DEFAULT_FUSE;
    if (const char* e = getenv("MF_STREAM_FUSE")) fuse = atoi(e) != 0;
#define MF_STREAM_CASE(NBO, NW, CW, RMAX, MINB)                                                        \
    return fuse ? launch_stream<NBO, NW, CW, RMAX, MINB, true>(p, gm, ws_bytes, stream)                 \
                : launch_stream<NBO, NW, CW, RMAX, MINB, false>(p, gm, ws_bytes, stream)
    if (gm.R > 256) { MF_STREAM_CASE(16, 8, 64, 512, 1); }
    if (gm.NW == 8 && gm.MINB == 2) { MF_STREAM_CASE(16, 8, 64, 256, 2); }
    if (gm.NW == 4 && gm.MINB == 3) { MF_STREAM_CASE(16, 4, 32, 256, 3); }
    if (gm.NW == 4) { MF_STREAM_CASE(16, 4, 64, 256, 2); }
    MF_STREAM_CASE(32, 8, 64, 256, 1);
#undef MF_STREAM_CASE
}
