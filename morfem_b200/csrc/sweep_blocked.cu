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
//      LAPACK's izamax magnitude |re| + |im| with first-maximum tie breaking, evaluated with three warp REDUX
//      operations (high word, low word, position) -- no CTA barrier inside the panel.  Row exchanges are tracked as
//      positions and materialise when the panel is written back.
//   B. one thread per trailing column applies the 8 row exchanges and the unit-lower triangular solve U12 = L11^-1 A12.
//   C. all warps: A22 -= L21 U12 as complex DMMA.8x8x4 block products (4 real DMMAs per complex k-step), fragments
//      loaded straight from the swizzled matrix.
// The right-hand sides ride along as extra column blocks, so L is never needed again; back substitution
// (`lu_solve`, implementation.py:478) and the S-parameter algebra (test_helpers.py:9-14) form the epilogue.
// Roofline: FP64 pipe.  Operators are L2 resident; per point the kernel writes 16 m^2 bytes (+ 16 r m with X).
#include "sweep_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int swz(int g) { return (((g ^ (g >> 2)) & 1) << 2) | (g & 3); }
__device__ __forceinline__ int mphys(int row, int col, int LD) { return row * LD + (col & ~7) + ((col & 7) ^ swz(row & 7)); }

// 1/a by Smith's formula with reciprocals (two correctly rounded reciprocals instead of three divisions)
__device__ __forceinline__ cplx crecip2(cplx a) {
    if (fabs(a.x) >= fabs(a.y)) {
        const double ia = 1.0 / a.x, t = a.y * ia, d = fma(a.y, t, a.x), id = 1.0 / d;
        return cmake(id, -t * id);
    } else {
        const double ib = 1.0 / a.y, t = a.x * ib, d = fma(a.x, t, a.y), id = 1.0 / d;
        return cmake(t * id, -id);
    }
}

// ---- A. panel factorisation by one warp -------------------------------------------------------------------
// Rows row0 .. R-1, columns row0 .. row0+7.  On exit the panel holds (at the exchanged row positions) L11 \ U11 with
// the RECIPROCAL of each pivot on the diagonal, and L21 below; piv[j] = position the j-th pivot row came from.
template <int SLOTS>
__device__ __forceinline__ void panel_factor(cplx* __restrict__ M, const int LD, const int R, const int row0, const int lane,
                                             int* __restrict__ piv, int* __restrict__ info_sh) {
    double ar[SLOTS][8], ai[SLOTS][8];
    int pos[SLOTS];
    unsigned done = 0, valid = 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int row = row0 + lane + 32 * s;
        pos[s] = row;
        const bool v = row < R;
        if (v) valid |= 1u << s;
        const int sw = swz(row & 7);
        const cplx* src = M + row * LD + row0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            cplx x = cmake(0.0, 0.0);
            if (v) x = src[c ^ sw];
            ar[s][c] = x.x; ai[s][c] = x.y;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // this lane's best candidate: (magnitude descending, position ascending)
        unsigned bh = 0, bl = 0; int bp = 0x7fffffff, bs = 0;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const bool cnd = ((valid >> s) & 1u) && !((done >> s) & 1u);
            const double v = fabs(ar[s][j]) + fabs(ai[s][j]);
            const unsigned h = (unsigned)__double2hiint(v) + 1u, l = (unsigned)__double2loint(v);
            const bool better = (h > bh) || (h == bh && (l > bl || (l == bl && pos[s] < bp)));
            if (cnd && better) { bh = h; bl = l; bp = pos[s]; bs = s; }
        }
        const unsigned hmax = __reduce_max_sync(FULL, bh);
        const bool c1 = (bh == hmax);
        const unsigned lmax = __reduce_max_sync(FULL, c1 ? bl : 0u);
        const bool c2 = c1 && (bl == lmax);
        const int P = __reduce_min_sync(FULL, c2 ? bp : 0x7fffffff);
        const bool own = c2 && (bp == P);
        const int olane = __ffs(__ballot_sync(FULL, own)) - 1;
        const int oslot = __shfl_sync(FULL, bs, olane);
        // pivot row, columns j..7
        double ur[8], ui[8];
#pragma unroll
        for (int c = j; c < 8; ++c) {
            double tr = ar[0][c], ti = ai[0][c];
#pragma unroll
            for (int s = 1; s < SLOTS; ++s) if (oslot == s) { tr = ar[s][c]; ti = ai[s][c]; }
            ur[c] = __shfl_sync(FULL, tr, olane); ui[c] = __shfl_sync(FULL, ti, olane);
        }
        const bool zero = (hmax == 1u && lmax == 0u);          // pivot magnitude is exactly +0.0
        const cplx rcp = zero ? cmake(0.0, 0.0) : crecip2(cmake(ur[j], ui[j]));
        const int T = row0 + j;
        if (lane == 0) { piv[j] = P; if (zero && *info_sh == 0) *info_sh = T + 1; }
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (own && s == bs) {
                done |= 1u << s; pos[s] = T;
                ar[s][j] = rcp.x; ai[s][j] = rcp.y;            // reciprocal pivot on the diagonal
            } else {
                if (pos[s] == T) pos[s] = P;                   // the row that sat at the target position moves away
                if (((valid >> s) & 1u) && !((done >> s) & 1u)) {
                    const cplx l = cmul(cmake(ar[s][j], ai[s][j]), rcp);
                    ar[s][j] = l.x; ai[s][j] = l.y;
#pragma unroll
                    for (int c = j + 1; c < 8; ++c) {
                        ar[s][c] = fma(-l.x, ur[c], ar[s][c]); ar[s][c] = fma(l.y, ui[c], ar[s][c]);
                        ai[s][c] = fma(-l.x, ui[c], ai[s][c]); ai[s][c] = fma(-l.y, ur[c], ai[s][c]);
                    }
                }
            }
        }
    }
    // every lane has read its rows long ago (the REDUX/SHFL above synchronise the warp): write to the new positions
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        if ((valid >> s) & 1u) {
            const int q = pos[s];
            const int sw = swz(q & 7);
            cplx* dst = M + q * LD + row0;
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c ^ sw] = cmake(ar[s][c], ai[s][c]);
        }
    }
}

template <int SLOTS>
__device__ __forceinline__ void panel_dispatch(cplx* M, int LD, int R, int row0, int lane, int* piv, int* info_sh) {
    const int left = R - row0;
    if (SLOTS >= 4 && left > 96) panel_factor<(SLOTS >= 4 ? 4 : SLOTS)>(M, LD, R, row0, lane, piv, info_sh);
    else if (SLOTS >= 3 && left > 64) panel_factor<(SLOTS >= 3 ? 3 : SLOTS)>(M, LD, R, row0, lane, piv, info_sh);
    else if (SLOTS >= 2 && left > 32) panel_factor<(SLOTS >= 2 ? 2 : SLOTS)>(M, LD, R, row0, lane, piv, info_sh);
    else panel_factor<1>(M, LD, R, row0, lane, piv, info_sh);
}

// ---- the kernel ----------------------------------------------------------------------------------------------
template <int SLOTS, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) sweep_blocked_kernel(SweepParams p, int R, int NCB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = NW * 32;
    const int r = p.r, m = p.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int LD = NCB * 8, NRB = R >> 3;
    const int g = lane >> 2, t = lane & 3, sg = swz(g);

    cplx* M = reinterpret_cast<cplx*>(smem_raw);                 // R x LD, swizzled
    cplx* zmat = M + (size_t)R * LD;                             // m*m
    cplx* zscr = zmat + m * m;                                   // 2*m*m
    int* piv = reinterpret_cast<int*>(zscr + 2 * m * m);         // 8
    int* info_sh = piv + 8;                                      // 1

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        // ---- assemble [A(t) | cb Br], identity on the padded diagonal ----
        for (int i = warp; i < R; i += NW) {
            const int sw = swz(i & 7);
            for (int j = lane; j < LD; j += 32) {
                cplx v = cmake(0.0, 0.0);
                if (j < R) {
                    if (i < r && j < r) {
                        const long long off = (long long)i * p.lda + j;
                        if (p.A0) { const cplx a = __ldg(p.A0 + off); v.x = c0 * a.x; v.y = c0 * a.y; }
                        if (p.A1) { const cplx a = __ldg(p.A1 + off); v.x = fma(c1, a.x, v.x); v.y = fma(c1, a.y, v.y); }
                        if (p.A2) { const cplx a = __ldg(p.A2 + off); v.x = fma(c2, a.x, v.x); v.y = fma(c2, a.y, v.y); }
                    } else if (i == j) v.x = 1.0;
                } else if (i < r && j - R < m) {
                    const cplx b = __ldg(p.Br + (long long)i * p.ldb + (j - R));
                    v.x = cb * b.x; v.y = cb * b.y;
                }
                M[i * LD + (j & ~7) + ((j & 7) ^ sw)] = v;
            }
        }
        if (tid == 0) *info_sh = 0;
        __syncthreads();

        // ---- blocked LU, right-hand sides eliminated alongside ----
        for (int k = 0; k < NRB; ++k) {
            const int row0 = 8 * k;
            if (warp == 0) panel_dispatch<SLOTS>(M, LD, R, row0, lane, piv, info_sh);
            __syncthreads();
            // B. row exchanges + U12 = L11^-1 A12, one thread per trailing column
            const int c_lo = row0 + 8;
            for (int c = c_lo + tid; c < LD; c += NT) {
                const int cbase = c & ~7, cin = c & 7;
                cplx u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = M[(row0 + j) * LD + cbase + (cin ^ swz(j))];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int P = piv[j];
                    if (P >= c_lo) {
                        cplx* q = M + P * LD + cbase + (cin ^ swz(P & 7));
                        const cplx tmp = *q; *q = u[j]; u[j] = tmp;
                    } else {
#pragma unroll
                        for (int q = j + 1; q < 8; ++q) if (P == row0 + q) { const cplx tmp = u[q]; u[q] = u[j]; u[j] = tmp; }
                    }
                }
#pragma unroll
                for (int j = 1; j < 8; ++j) {
                    const cplx* lrow = M + (row0 + j) * LD + row0;
                    const int sw = swz(j);
#pragma unroll
                    for (int i = 0; i < j; ++i) cfms(u[j], lrow[i ^ sw], u[i]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) M[(row0 + j) * LD + cbase + (cin ^ swz(j))] = u[j];
            }
            __syncthreads();
            // C. trailing update A22 -= L21 U12 (complex DMMA)
            const int nrb = NRB - (k + 1), ncb = NCB - (k + 1);
            const int ntiles = nrb * ncb;
            for (int ti = warp; ti < ntiles; ti += NW) {
                const int cbk = ti / nrb;
                const int cbi = k + 1 + cbk, rbi = k + 1 + (ti - cbk * nrb);
                const cplx* arow = M + (8 * rbi + g) * LD + row0;
                const cplx a0 = arow[t ^ sg], a1 = arow[(4 + t) ^ sg];
                const cplx b0 = M[(row0 + t) * LD + 8 * cbi + (g ^ swz(t))];
                const cplx b1 = M[(row0 + 4 + t) * LD + 8 * cbi + (g ^ swz(4 + t))];
                cplx* crow = M + (8 * rbi + g) * LD + 8 * cbi;
                cplx* pc0 = crow + ((2 * t) ^ sg);
                cplx* pc1 = crow + ((2 * t + 1) ^ sg);
                const cplx v0 = *pc0, v1 = *pc1;
                double cre0 = v0.x, cre1 = v1.x, cim0 = v0.y, cim1 = v1.y;
                dmma884(cre0, cre1, -a0.x, b0.x); dmma884(cim0, cim1, -a0.x, b0.y);
                dmma884(cre0, cre1, a0.y, b0.y);  dmma884(cim0, cim1, -a0.y, b0.x);
                dmma884(cre0, cre1, -a1.x, b1.x); dmma884(cim0, cim1, -a1.x, b1.y);
                dmma884(cre0, cre1, a1.y, b1.y);  dmma884(cim0, cim1, -a1.y, b1.x);
                *pc0 = cmake(cre0, cim0); *pc1 = cmake(cre1, cim1);
            }
            __syncthreads();
        }

        // ---- back substitution U x = y, one warp per right-hand side, solution kept in registers ----
        for (int c = warp; c < m; c += NW) {
            double yr[SLOTS], yi[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int i = s * 32 + lane;
                cplx v = cmake(0.0, 0.0);
                if (i < R) v = M[mphys(i, R + c, LD)];
                yr[s] = v.x; yi[s] = v.y;
            }
#pragma unroll
            for (int ks = SLOTS - 1; ks >= 0; --ks) {
                for (int kl = 31; kl >= 0; --kl) {
                    const int k = ks * 32 + kl;
                    if (k >= R) continue;
                    const cplx inv = M[mphys(k, k, LD)];
                    const cplx xk = cmul(cmake(yr[ks], yi[ks]), inv);
                    const double xr = __shfl_sync(FULL, xk.x, kl), xi = __shfl_sync(FULL, xk.y, kl);
                    if (lane == kl) { yr[ks] = xr; yi[ks] = xi; }
#pragma unroll
                    for (int s = 0; s <= ks; ++s) {
                        const int i = s * 32 + lane;
                        if (i < k) {
                            const cplx u = M[mphys(i, k, LD)];
                            yr[s] = fma(-u.x, xr, yr[s]); yr[s] = fma(u.y, xi, yr[s]);
                            yi[s] = fma(-u.x, xi, yi[s]); yi[s] = fma(-u.y, xr, yi[s]);
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int i = s * 32 + lane;
                if (i < r) {
                    const cplx x = cmake(yr[s], yi[s]);
                    M[mphys(i, R + c, LD)] = x;
                    if (p.X) p.X[(pt * r + i) * m + c] = x;
                }
            }
        }
        __syncthreads();
        if (p.info && tid == 0) p.info[pt] = *info_sh;

        // ---- S-parameters: Z = j zs x^T (cb Br) ----
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
                if (lane == 0) { const double zs = p.zs[pt]; zmat[e] = cmake(-zs * acc.y, zs * acc.x); }
            }
            __syncthreads();
            if (tid == 0) gsm_from_impedance(zmat, zscr, m, p.S + pt * (long long)m * m);
        }
        __syncthreads();
    }
}

struct BlockedGeom { int R, NCB; size_t smem; };

BlockedGeom blocked_geom(int r, int m) {
    BlockedGeom gm;
    gm.R = (r + 7) / 8 * 8;
    gm.NCB = gm.R / 8 + (m + 7) / 8;
    gm.smem = sizeof(cplx) * ((size_t)gm.R * gm.NCB * 8 + 3 * (size_t)m * m) + 64;
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
    return 0;
}

}  // namespace

bool sweep_blocked_supports(int r, int m) {
    if (r < 1 || m < 1 || m > MF_MAX_PORTS) return false;
    const BlockedGeom gm = blocked_geom(r, m);
    return gm.R <= 128 && gm.smem <= 226 * 1024;
}

size_t sweep_blocked_ws_bytes(int, int, long long) { return 0; }

int sweep_blocked_launch(const SweepParams& p, size_t, cudaStream_t stream) {
    const BlockedGeom gm = blocked_geom(p.r, p.m);
    if (gm.R <= 32) return launch_blocked<1, 2, 8>(p, gm, stream);
    if (gm.R <= 64) return launch_blocked<2, 4, 3>(p, gm, stream);
    if (gm.R <= 96) return launch_blocked<3, 8, 1>(p, gm, stream);
    return launch_blocked<4, 8, 1>(p, gm, stream);
}
