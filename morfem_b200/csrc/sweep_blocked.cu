// Variant 3 of the batched reduced sweep: blocked right-looking LU with FP64 tensor-core (DMMA) trailing updates,
// augmented matrix [A(t) | cb(t) Br] resident in shared memory (r <= 112).
//
// One CTA per frequency point (persistent, grid-stride).  The matrix is padded to R = 8*ceil(r/8) rows (identity on the
// padded diagonal) and NCB = R/8 + ceil(m/8) column blocks of 8, stored row-major with an XOR swizzle of the column
// index inside every 8-column block (physical col = col ^ swz(row & 7)), which makes ALL access patterns used below
// free of shared-memory bank conflicts without padding: DMMA A-, B- and C-fragments, row-per-lane panel loads and
// column-per-thread sweeps.
//
// Per panel of 8 columns (LAPACK getrf order; implementation.py:477 `lu_factor`):
//   A. ONE warp factors the (R - 8k) x 8 panel in registers, one (or a few) rows per lane.  Partial pivoting uses
//      LAPACK's izamax magnitude |re| + |im| with first-maximum tie breaking, evaluated with warp REDUX/VOTE
//      operations (high word; low word and position only on ties) -- no CTA barrier inside the panel.  Row exchanges
//      are tracked as positions and materialise when the panel is written back.
//   B. one thread per trailing column applies the 8 row exchanges and the unit-lower triangular solve U12 = L11^-1 A12.
//   C. A22 -= L21 U12 as complex DMMA.8x8x4 block products (4 real DMMAs per complex k-step), fragments loaded
//      straight from the swizzled matrix.
// Schedule (look-ahead): ONE CTA barrier per panel.  After the barrier that publishes panel k, warp 0 alone applies
// B + C to column block k+1 and immediately factors panel k+1 (the critical path), while the other warps apply
// B + C of step k to all remaining column blocks in its shadow.
// The right-hand sides ride along as extra column blocks, so L is never needed again; back substitution
// (`lu_solve`, implementation.py:478) and the S-parameter algebra (test_helpers.py:9-14) form the epilogue.
// Roofline: FP64 pipe.  Operators are L2 resident; per point the kernel writes 16 m^2 bytes (+ 16 r m with X).
#include "sweep_blocked.cuh"
#include <stdlib.h>

namespace {

// ---- A. panel factorisation by one warp -------------------------------------------------------------------
// Rows row0 .. R-1, columns row0 .. row0+7, one (or a few) rows per lane, held in registers.  On exit the panel holds
// (at the exchanged row positions) -L11 \ U11 with the RECIPROCAL of each pivot on the diagonal, and -L21 below
// (multipliers are stored NEGATED so that the trailing update and the triangular solve are pure multiply-adds);
// piv[j] = position the j-th pivot row came from.
// Per column: every lane forms the magnitude of its best candidate and -- speculatively, overlapping the latency of
// the warp reduction -- its reciprocal; the winning lane publishes its finished row (L11 part, reciprocal pivot, U
// part) straight into the row's final place in shared memory and retires the slot; after one __syncwarp the other
// lanes read reciprocal and U entries from there.  `pbuf` = two ints for the pivot position (double buffered).
template <int SLOTS>
__device__ __forceinline__ void panel_factor(cplx* M, const int LD, const int R, const int row0, const int lane,
                                          int* pbuf, int* piv, int* info_sh) {
    cplx a[SLOTS][8];
    int pos[SLOTS];
    bool act[SLOTS];                        // slot holds a real row that has not been a pivot yet
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int row = row0 + lane + 32 * s;
        pos[s] = row;
        act[s] = row < R;
        const int sw = swz(row & 7);
        const cplx* src = M + row * LD + row0;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[s][c] = act[s] ? src[c ^ sw] : cmake(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int T = row0 + j;
        // this lane's best candidate: (magnitude descending, position ascending); retired slots count as -1
        double vb = act[0] ? fabs(a[0][j].x) + fabs(a[0][j].y) : -1.0;
        int pbest = pos[0], bs = 0;
        cplx cand = a[0][j];
#pragma unroll
        for (int s = 1; s < SLOTS; ++s) {
            const double v = act[s] ? fabs(a[s][j].x) + fabs(a[s][j].y) : -1.0;
            if (v > vb || (v == vb && pos[s] < pbest)) { vb = v; pbest = pos[s]; bs = s; cand = a[s][j]; }
        }
        // speculative reciprocal of this lane's candidate
        cplx rc = cmake(0.0, 0.0);
        if (vb > 0.0) rc = (cand.y == 0.0) ? cmake(1.0 / cand.x, 0.0) : crecip2(cand);
        // warp arg-max: high word first, low word and position only when needed
        const int hi = __double2hiint(vb);
        const int hmax = __reduce_max_sync(FULL, hi);
        bool own = (hi == hmax);
        if (__popc(__ballot_sync(FULL, own)) != 1) {
            const unsigned lo = (unsigned)__double2loint(vb);
            const unsigned lmax = __reduce_max_sync(FULL, own ? lo : 0u);
            own = own && (lo == lmax);
            if (__popc(__ballot_sync(FULL, own)) != 1) {       // exact tie: lowest position wins (first maximum, as izamax)
                const int pmin = __reduce_min_sync(FULL, own ? pbest : 0x7fffffff);
                own = own && (pbest == pmin);
            }
        }
        cplx* prow = M + T * LD + row0;     // final place of the pivot row; (T & 7) == j
        if (own) {
            pbuf[j & 1] = pbest;
            piv[j] = pbest;
            if (!(vb > 0.0) && *info_sh == 0) *info_sh = T + 1;   // exactly zero pivot (LAPACK info)
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
                if (s == bs) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) prow[c ^ swz(j)] = (c == j) ? rc : a[s][c];
                    act[s] = false;
                }
        }
        __syncwarp();
        const int P = *reinterpret_cast<volatile int*>(pbuf + (j & 1));
        const unsigned prow_s = (unsigned)__cvta_generic_to_shared(prow);
        cplx rcp, u[8];
        asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rcp.x), "=d"(rcp.y) : "r"(prow_s + 16u * (j ^ swz(j))));
#pragma unroll
        for (int c = j + 1; c < 8; ++c)
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u[c].x), "=d"(u[c].y) : "r"(prow_s + 16u * (c ^ swz(j))));
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            // negated multiplier (retired slots compute on stale values that are never used again)
            const cplx nl = cmake(-(a[s][j].x * rcp.x - a[s][j].y * rcp.y), -(a[s][j].x * rcp.y + a[s][j].y * rcp.x));
            a[s][j] = nl;
#pragma unroll
            for (int c = j + 1; c < 8; ++c) cfma(a[s][c], nl, u[c]);
            if (pos[s] == T) pos[s] = P;    // the row that sat at the target position moves to where the pivot came from
        }
    }
    // rows that never became a pivot: -L21, written to their (exchanged) positions
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        if (act[s]) {
            const int q = pos[s];
            const int sw = swz(q & 7);
            cplx* dst = M + q * LD + row0;
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c ^ sw] = a[s][c];
        }
    }
}

template <int SLOTS>
__device__ __forceinline__ void panel_dispatch(cplx* M, int LD, int R, int row0, int lane, int* pbuf, int* piv, int* info_sh) {
    const int left = R - row0;
    if (SLOTS >= 4 && left > 96) panel_factor<(SLOTS >= 4 ? 4 : SLOTS)>(M, LD, R, row0, lane, pbuf, piv, info_sh);
    else if (SLOTS >= 3 && left > 64) panel_factor<(SLOTS >= 3 ? 3 : SLOTS)>(M, LD, R, row0, lane, pbuf, piv, info_sh);
    else if (SLOTS >= 2 && left > 32) panel_factor<(SLOTS >= 2 ? 2 : SLOTS)>(M, LD, R, row0, lane, pbuf, piv, info_sh);
    else panel_factor<1>(M, LD, R, row0, lane, pbuf, piv, info_sh);
}

// ---- the kernel ----------------------------------------------------------------------------------------------
template <int SLOTS, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) sweep_blocked_kernel(SweepParams p, int R, int NCB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = NW * 32, NWK = NT - 32;                   // NWK worker threads (warps 1 .. NW-1)
    const int r = p.r, m = p.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int LD = NCB * 8, NRB = R >> 3;

    cplx* M = reinterpret_cast<cplx*>(smem_raw);                 // R x LD, swizzled
    int* pbuf = reinterpret_cast<int*>(M + (size_t)R * LD);      // 2 ints: pivot position staging (+ 2 pad)
    int* piv = pbuf + 4;                                         // 2 x 8, double buffered over the panel index
    int* info_sh = piv + 16;                                     // 1
    // the warp that runs the critical path rotates with the CTA index so that the (up to three) resident CTAs of
    // an SM do not all put it on the same scheduler
    const int pw = (int)(blockIdx.x % NW);
    const int wk = warp - (warp > pw ? 1 : 0);                   // worker index 0 .. NW-2 of the other warps

    FragOff fo;
    {
        const int g = lane >> 2, t = lane & 3, sg = swz(g);
        fo.g = g;
        fo.a0 = t ^ sg; fo.a1 = (4 + t) ^ sg;                    // A-fragment: row g, k = t / 4 + t
        fo.c0 = (2 * t) ^ sg; fo.c1 = (2 * t + 1) ^ sg;          // C-fragment: row g, cols 2t, 2t+1
        fo.b0 = t * LD + (g ^ swz(t)); fo.b1 = (4 + t) * LD + (g ^ swz(4 + t));   // B-fragment: row k, col g
    }
    const bool hasA0 = p.A0 != nullptr, hasA1 = p.A1 != nullptr, hasA2 = p.A2 != nullptr;

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        // ---- assemble [A(t) | cb Br], identity on the padded diagonal ----
        // (all loads of four rows are issued before the first use: the operators come from L2, ~1 us away)
        constexpr int RG = 4;
        for (int i0 = warp; i0 < r; i0 += NW * RG) {
            cplx x0[RG][SLOTS], x2[RG][SLOTS], xb[RG];
#pragma unroll
            for (int q = 0; q < RG; ++q) {
                const int i = i0 + q * NW;
                const long long rowoff = (long long)i * p.lda;
#pragma unroll
                for (int jj = 0; jj < SLOTS; ++jj) {
                    const int j = lane + 32 * jj;
                    const bool ok = i < r && j < r;
                    x0[q][jj] = (ok && hasA0) ? __ldg(p.A0 + rowoff + j) : cmake(0.0, 0.0);
                    x2[q][jj] = (ok && hasA2) ? __ldg(p.A2 + rowoff + j) : cmake(0.0, 0.0);
                }
                xb[q] = (i < r && lane < m) ? __ldg(p.Br + (long long)i * p.ldb + lane) : cmake(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < RG; ++q) {
                const int i = i0 + q * NW;
                if (i < r) {
                    const int sw = swz(i & 7);
                    cplx* Mrow = M + i * LD;
#pragma unroll
                    for (int jj = 0; jj < SLOTS; ++jj) {
                        const int j = lane + 32 * jj;
                        if (j < R) Mrow[(j & ~7) + ((j & 7) ^ sw)] = cmake(fma(c2, x2[q][jj].x, c0 * x0[q][jj].x), fma(c2, x2[q][jj].y, c0 * x0[q][jj].y));
                    }
                    if (lane < LD - R) Mrow[R + (lane & ~7) + ((lane & 7) ^ sw)] = cmake(cb * xb[q].x, cb * xb[q].y);
                }
            }
        }
        if (hasA1) {                        // rarely present (the reference's a1 is an empty matrix): same thread, same elements
            for (int i = warp; i < r; i += NW) {
                const int sw = swz(i & 7);
                for (int j = lane; j < r; j += 32) {
                    const cplx a = __ldg(p.A1 + (long long)i * p.lda + j);
                    cplx* e = M + i * LD + (j & ~7) + ((j & 7) ^ sw);
                    cplx v = *e; v.x = fma(c1, a.x, v.x); v.y = fma(c1, a.y, v.y); *e = v;
                }
            }
        }
        for (int i = r + warp; i < R; i += NW) {
            const int sw = swz(i & 7);
            for (int j = lane; j < LD; j += 32) M[i * LD + (j & ~7) + ((j & 7) ^ sw)] = cmake(j == i ? 1.0 : 0.0, 0.0);
        }
        if (tid == 0) *info_sh = 0;
        __syncthreads();

        // ---- blocked LU with look-ahead, right-hand sides eliminated alongside ----
        if (warp == pw) panel_dispatch<SLOTS>(M, LD, R, 0, lane, pbuf, piv, info_sh);
        for (int k = 0; k < NRB; ++k) {
            __syncthreads();                   // panel k is published; every warp has finished step k-1
            const int row0 = 8 * k;
            const int* pv = piv + 8 * (k & 1);
            const int nrb = NRB - (k + 1);
            if (warp == pw) {
                if (nrb > 0) {                 // critical path: column block k+1, then the next panel
                    if (lane < 8) stepb_column(M, LD, row0, row0 + 8 + lane, pv);
                    __syncwarp();
                    update_tiles(M, LD, row0, k + 1, nrb, k + 1, 0, nrb, fo);
                    __syncwarp();
                    panel_dispatch<SLOTS>(M, LD, R, row0 + 8, lane, pbuf, piv + 8 * ((k + 1) & 1), info_sh);
                }
            } else if (NW > 1) {               // everything to the right of column block k+1 (all of it after the last panel)
                const int cb0 = nrb > 0 ? k + 2 : k + 1;
                for (int c = 8 * cb0 + wk * 32 + lane; c < LD; c += NWK) stepb_column(M, LD, row0, c, pv);
                if (nrb > 0) {
                    asm volatile("bar.sync 1, %0;" :: "n"(NWK > 0 ? NWK : 32) : "memory");
                    const int ntiles = nrb * (NCB - cb0);
                    update_tiles(M, LD, row0, k + 1, nrb, cb0, (ntiles * wk) / (NW - 1), (ntiles * (wk + 1)) / (NW - 1), fo);
                }
            }
        }
        __syncthreads();

        // ---- back substitution U x = y (`lu_solve`), one warp per right-hand side, solution in registers (row i in
        //      lane i % 32, slot i / 32), processed in blocks of 8 rows: the 8 x 8 triangular system of a block is
        //      solved redundantly by every lane from broadcast loads, then each lane updates its own rows with the
        //      8 new unknowns -- 8 dependent steps per block instead of 8 shuffle round trips ----
        for (int c = warp; c < m; c += NW) {
            cplx y[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int i = s * 32 + lane;
                y[s] = i < R ? M[mphys(i, R + c, LD)] : cmake(0.0, 0.0);
            }
#pragma unroll
            for (int ks = SLOTS - 1; ks >= 0; --ks) {
#pragma unroll (SLOTS <= 2 ? 4 : 1)
                for (int kq = 3; kq >= 0; --kq) {
                    const int kb8 = ks * 32 + kq * 8;                    // first row of the block
                    if (kb8 < R) {
                        cplx x[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { x[j].x = __shfl_sync(FULL, y[ks].x, kq * 8 + j); x[j].y = __shfl_sync(FULL, y[ks].y, kq * 8 + j); }
#pragma unroll
                        for (int j = 7; j >= 0; --j) {
                            const cplx* urow = M + (kb8 + j) * LD + kb8;
                            const int sw = swz(j);
#pragma unroll
                            for (int jj = 7; jj > j; --jj) cfms(x[j], urow[jj ^ sw], x[jj]);   // newest unknown last (ztrsm order): short dependent chain
                            x[j] = cmul(x[j], urow[j ^ sw]);             // reciprocal pivot on the diagonal
                        }
#pragma unroll
                        for (int s = 0; s <= ks; ++s) {
                            const int i = s * 32 + lane;
                            if (i < kb8) {
                                const cplx* urow = M + i * LD + kb8;
                                const int sw = swz(i & 7);
#pragma unroll
                                for (int j = 0; j < 8; ++j) cfms(y[s], urow[j ^ sw], x[j]);
                            }
                        }
                        if ((lane >> 3) == kq) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) if ((lane & 7) == j) y[ks] = x[j];
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int i = s * 32 + lane;
                if (i < r) {
                    M[mphys(i, R + c, LD)] = y[s];
                    if (p.X) p.X[(pt * r + i) * m + c] = y[s];
                }
            }
        }
        __syncthreads();
        if (p.info && tid == 0) p.info[pt] = *info_sh;

        // ---- impedance matrix Z = j zs x^T (cb Br), written to S; the m x m algebra S = 2 (I + Z^-1)^-1 - I
        //      (test_helpers.py:11-14) is finished for all points by gsm_finish_kernel ----
        if (p.S) {
            for (int e = warp; e < m * m; e += NW) {
                const int a = e / m, b = e - a * m;
                cplx acc = cmake(0.0, 0.0);
                for (int k = lane; k < r; k += 32) cfma(acc, M[mphys(k, R + a, LD)], cscale(cb, __ldg(p.Br + (long long)k * p.ldb + b)));
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

struct BlockedGeom { int R, NCB; size_t smem; };

BlockedGeom blocked_geom(int r, int m) {
    BlockedGeom gm;
    gm.R = (r + 7) / 8 * 8;
    gm.NCB = gm.R / 8 + (m + 7) / 8;
    gm.smem = sizeof(cplx) * ((size_t)gm.R * gm.NCB * 8) + 128;
    return gm;
}

template <int SLOTS, int NW, int MINB>
int launch_blocked(const SweepParams& p, const BlockedGeom& gm, cudaStream_t stream) {
    auto kern = sweep_blocked_kernel<SLOTS, NW, MINB>;
    MF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gm.smem));
    int per_sm = 0;
    MF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, gm.smem));
    if (per_sm < 1) MF_FAIL_ARG(7, "blocked sweep does not fit on an SM for this (r, m)");
    long long grid = (long long)mf_num_sms() * per_sm;
    if (grid > p.F) grid = p.F;
    kern<<<(unsigned)grid, NW * 32, gm.smem, stream>>>(p, gm.R, gm.NCB);
    MF_CHECK_LAUNCH();
    if (p.S) return gsm_finish_launch(p.S, p.m, p.F, stream);
    return 0;
}

}  // namespace

bool sweep_blocked_fits_smem(int r, int m) {
    if (r < 1 || m < 1 || m > MF_MAX_PORTS) return false;
    const BlockedGeom gm = blocked_geom(r, m);
    return gm.R <= 128 && gm.smem <= 226 * 1024;
}

bool sweep_blocked_supports(int r, int m) { return sweep_blocked_fits_smem(r, m); }

size_t sweep_blocked_ws_bytes(int, int, long long) { return 0; }

int sweep_blocked_launch(const SweepParams& p_in, size_t ws_bytes, cudaStream_t stream) {
    SweepParams p = p_in;
    if (!sweep_blocked_fits_smem(p.r, p.m)) MF_FAIL_ARG(7, "matrix does not fit in shared memory (use the left-looking variant)");
    const BlockedGeom gm = blocked_geom(p.r, p.m);
#ifdef MF_BLOCKED_DEBUG
    if (getenv("MF_BLOCKED_DUMP") && p.ws && ws_bytes >= sizeof(cplx) * gm.R * gm.NCB * 8) p.ws_stride = -12345;
#endif
    (void)ws_bytes;
    if (gm.R <= 32) return launch_blocked<1, 2, 8>(p, gm, stream);
    if (gm.R <= 64) return launch_blocked<2, 4, 3>(p, gm, stream);
    if (gm.R <= 96) return launch_blocked<3, 8, 1>(p, gm, stream);
    return launch_blocked<4, 8, 1>(p, gm, stream);
}
