// Real float64 twin of sweep_blocked.cu: the reference's reduced models are real (implementation.py:190
// allocates a float64 result), so for real reduced operators the batched LU runs on 8-byte elements -- half the shared
// memory per point (36 KiB at r = 64: up to six points in flight per SM instead of three), a quarter of the flops, and
// results bit-identical to the complex128 kernel (whose imaginary parts would all be exact zeros).  Same algorithm,
// schedule and storage scheme; see sweep_blocked.cu for the description.  The S-parameters stay complex: Z = j zs x^T b.
#include <stdlib.h>
#include "sweep_common.cuh"

namespace {

constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ int rswz(int g) { return (((g ^ (g >> 2)) & 1) << 2) | (g & 3); }
__device__ __forceinline__ int rphys(int row, int col, int LD) { return row * LD + (col & ~7) + ((col & 7) ^ rswz(row & 7)); }

struct SweepParamsR {
    const double* A0; const double* A1; const double* A2; long long lda;
    const double* Br; long long ldb;
    int r, m;
    const double* c0; const double* c1; const double* c2; const double* cb; const double* zs;
    long long F;
    double* X;    // F x r x m or NULL
    cplx* S;      // F x m x m or NULL
    int* info;
};

// ---- A. panel factorisation by one warp (see sweep_blocked.cu) -------------------------------------------------------
template <int SLOTS>
__device__ __forceinline__ void rpanel_factor(double* M, const int LD, const int R, const int row0, const int lane,
                                              int* pbuf, int* piv, int* info_sh) {
    double a[SLOTS][8];
    int pos[SLOTS];
    bool act[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int row = row0 + lane + 32 * s;
        pos[s] = row;
        act[s] = row < R;
        const int sw = rswz(row & 7);
        const double* src = M + row * LD + row0;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[s][c] = act[s] ? src[c ^ sw] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int T = row0 + j;
        double vb = act[0] ? fabs(a[0][j]) : -1.0;
        int pbest = pos[0], bs = 0;
        double cand = a[0][j];
#pragma unroll
        for (int s = 1; s < SLOTS; ++s) {
            const double v = act[s] ? fabs(a[s][j]) : -1.0;
            if (v > vb || (v == vb && pos[s] < pbest)) { vb = v; pbest = pos[s]; bs = s; cand = a[s][j]; }
        }
        const double rc = (vb > 0.0) ? 1.0 / cand : 0.0;          // speculative reciprocal of this lane's candidate
        const int hi = __double2hiint(vb);
        const int hmax = __reduce_max_sync(FULLM, hi);
        bool own = (hi == hmax);
        if (__popc(__ballot_sync(FULLM, own)) != 1) {
            const unsigned lo = (unsigned)__double2loint(vb);
            const unsigned lmax = __reduce_max_sync(FULLM, own ? lo : 0u);
            own = own && (lo == lmax);
            if (__popc(__ballot_sync(FULLM, own)) != 1) {        // exact tie: lowest position wins (first maximum, as idamax)
                const int pmin = __reduce_min_sync(FULLM, own ? pbest : 0x7fffffff);
                own = own && (pbest == pmin);
            }
        }
        double* prow = M + T * LD + row0;                          // final place of the pivot row; (T & 7) == j
        if (own) {
            pbuf[j & 1] = pbest;
            piv[j] = pbest;
            if (!(vb > 0.0) && *info_sh == 0) *info_sh = T + 1;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
                if (s == bs) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) prow[c ^ rswz(j)] = (c == j) ? rc : a[s][c];
                    act[s] = false;
                }
        }
        __syncwarp();
        const int P = *reinterpret_cast<volatile int*>(pbuf + (j & 1));
        const unsigned prow_s = (unsigned)__cvta_generic_to_shared(prow);
        double rcp, u[8];
        asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(rcp) : "r"(prow_s + 8u * (j ^ rswz(j))));
#pragma unroll
        for (int c = j + 1; c < 8; ++c) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(u[c]) : "r"(prow_s + 8u * (c ^ rswz(j))));
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const double nl = -(a[s][j] * rcp);                   // negated multiplier
            a[s][j] = nl;
#pragma unroll
            for (int c = j + 1; c < 8; ++c) a[s][c] = fma(nl, u[c], a[s][c]);
            if (pos[s] == T) pos[s] = P;
        }
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        if (act[s]) {
            const int q = pos[s];
            const int sw = rswz(q & 7);
            double* dst = M + q * LD + row0;
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c ^ sw] = a[s][c];
        }
    }
}

template <int SLOTS>
__device__ __forceinline__ void rpanel_dispatch(double* M, int LD, int R, int row0, int lane, int* pbuf, int* piv, int* info_sh) {
    const int left = R - row0;
    if (SLOTS >= 4 && left > 96) rpanel_factor<(SLOTS >= 4 ? 4 : SLOTS)>(M, LD, R, row0, lane, pbuf, piv, info_sh);
    else if (SLOTS >= 3 && left > 64) rpanel_factor<(SLOTS >= 3 ? 3 : SLOTS)>(M, LD, R, row0, lane, pbuf, piv, info_sh);
    else if (SLOTS >= 2 && left > 32) rpanel_factor<(SLOTS >= 2 ? 2 : SLOTS)>(M, LD, R, row0, lane, pbuf, piv, info_sh);
    else rpanel_factor<1>(M, LD, R, row0, lane, pbuf, piv, info_sh);
}

// ---- B. row exchanges + U12 = L11^-1 A12 for one trailing column c (L11 stored negated) ------------------------------
__device__ __forceinline__ void rstepb_column(double* M, const int LD, const int row0, const int c, const int* pv) {
    const int cbase = c & ~7, cin = c & 7;
    double* colp = M + row0 * LD + cbase;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int P = pv[j];
        if (P != row0 + j) {
            double* x = colp + j * LD + (cin ^ rswz(j));
            double* y = M + P * LD + cbase + (cin ^ rswz(P & 7));
            const double tmp = *x; *x = *y; *y = tmp;
        }
    }
    double u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = colp[j * LD + (cin ^ rswz(j))];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        const double* lrow = M + (row0 + j) * LD + row0;
        const int sw = rswz(j);
#pragma unroll
        for (int i = 0; i < j; ++i) u[j] = fma(lrow[i ^ sw], u[i], u[j]);
    }
#pragma unroll
    for (int j = 1; j < 8; ++j) colp[j * LD + (cin ^ rswz(j))] = u[j];
}

struct RFragOff { int a0, a1, c0, c1, b0, b1, g; };

// ---- C. tiles of the trailing update (one DMMA per k-step of 4) ------------------------------------------------------
__device__ __forceinline__ void rupdate_tiles(double* M, const int LD, const int row0, const int rb0, const int nrb, const int cb0,
                                              const int t_lo, const int t_hi, const RFragOff& fo) {
    if (t_hi <= t_lo) return;
    int cbk = t_lo / nrb, rbk = t_lo - cbk * nrb;
    const double* Ub = M + row0 * LD + 8 * cb0;
    double b0 = Ub[8 * cbk + fo.b0], b1 = Ub[8 * cbk + fo.b1];
    for (int ti = t_lo; ti < t_hi; ++ti) {
        double* rowp = M + (8 * (rb0 + rbk) + fo.g) * LD;
        const double a0 = rowp[row0 + fo.a0], a1 = rowp[row0 + fo.a1];
        double* pc0 = rowp + 8 * (cb0 + cbk) + fo.c0;
        double* pc1 = rowp + 8 * (cb0 + cbk) + fo.c1;
        double c0 = *pc0, c1 = *pc1;
        dmma884(c0, c1, a0, b0);
        dmma884(c0, c1, a1, b1);
        *pc0 = c0; *pc1 = c1;
        if (++rbk == nrb) {
            rbk = 0; ++cbk;
            if (ti + 1 < t_hi) { b0 = Ub[8 * cbk + fo.b0]; b1 = Ub[8 * cbk + fo.b1]; }
        }
    }
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
template <int SLOTS, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) sweep_blocked_f64_kernel(SweepParamsR p, int R, int NCB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = NW * 32, NWK = NT - 32;
    const int r = p.r, m = p.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int LD = NCB * 8, NRB = R >> 3;

    double* M = reinterpret_cast<double*>(smem_raw);             // R x LD, swizzled
    int* pbuf = reinterpret_cast<int*>(M + (size_t)R * LD);
    int* piv = pbuf + 4;
    int* info_sh = piv + 16;
    const int pw = (int)(blockIdx.x % NW);
    const int wk = warp - (warp > pw ? 1 : 0);

    RFragOff fo;
    {
        const int g = lane >> 2, t = lane & 3, sg = rswz(g);
        fo.g = g;
        fo.a0 = t ^ sg; fo.a1 = (4 + t) ^ sg;
        fo.c0 = (2 * t) ^ sg; fo.c1 = (2 * t + 1) ^ sg;
        fo.b0 = t * LD + (g ^ rswz(t)); fo.b1 = (4 + t) * LD + (g ^ rswz(4 + t));
    }
    const bool hasA0 = p.A0 != nullptr, hasA1 = p.A1 != nullptr, hasA2 = p.A2 != nullptr;

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        // ---- assemble [A(t) | cb Br], identity on the padded diagonal ----
        constexpr int RG = 4;
        for (int i0 = warp; i0 < r; i0 += NW * RG) {
            double x0[RG][SLOTS], x2[RG][SLOTS], xb[RG];
#pragma unroll
            for (int q = 0; q < RG; ++q) {
                const int i = i0 + q * NW;
                const long long rowoff = (long long)i * p.lda;
#pragma unroll
                for (int jj = 0; jj < SLOTS; ++jj) {
                    const int j = lane + 32 * jj;
                    const bool ok = i < r && j < r;
                    x0[q][jj] = (ok && hasA0) ? __ldg(p.A0 + rowoff + j) : 0.0;
                    x2[q][jj] = (ok && hasA2) ? __ldg(p.A2 + rowoff + j) : 0.0;
                }
                xb[q] = (i < r && lane < m) ? __ldg(p.Br + (long long)i * p.ldb + lane) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < RG; ++q) {
                const int i = i0 + q * NW;
                if (i < r) {
                    const int sw = rswz(i & 7);
                    double* Mrow = M + i * LD;
#pragma unroll
                    for (int jj = 0; jj < SLOTS; ++jj) {
                        const int j = lane + 32 * jj;
                        if (j < R) Mrow[(j & ~7) + ((j & 7) ^ sw)] = fma(c2, x2[q][jj], c0 * x0[q][jj]);
                    }
                    if (lane < LD - R) Mrow[R + (lane & ~7) + ((lane & 7) ^ sw)] = cb * xb[q];
                }
            }
        }
        if (hasA1) {
            for (int i = warp; i < r; i += NW) {
                const int sw = rswz(i & 7);
                for (int j = lane; j < r; j += 32) {
                    double* e = M + i * LD + (j & ~7) + ((j & 7) ^ sw);
                    *e = fma(c1, __ldg(p.A1 + (long long)i * p.lda + j), *e);
                }
            }
        }
        for (int i = r + warp; i < R; i += NW) {
            const int sw = rswz(i & 7);
            for (int j = lane; j < LD; j += 32) M[i * LD + (j & ~7) + ((j & 7) ^ sw)] = (j == i) ? 1.0 : 0.0;
        }
        if (tid == 0) *info_sh = 0;
        __syncthreads();

        // ---- blocked LU with look-ahead ----
        if (warp == pw) rpanel_dispatch<SLOTS>(M, LD, R, 0, lane, pbuf, piv, info_sh);
        for (int k = 0; k < NRB; ++k) {
            __syncthreads();
            const int row0 = 8 * k;
            const int* pv = piv + 8 * (k & 1);
            const int nrb = NRB - (k + 1);
            if (warp == pw) {
                if (nrb > 0) {
                    if (lane < 8) rstepb_column(M, LD, row0, row0 + 8 + lane, pv);
                    __syncwarp();
                    rupdate_tiles(M, LD, row0, k + 1, nrb, k + 1, 0, nrb, fo);
                    __syncwarp();
                    rpanel_dispatch<SLOTS>(M, LD, R, row0 + 8, lane, pbuf, piv + 8 * ((k + 1) & 1), info_sh);
                }
            } else if (NW > 1) {
                const int cb0 = nrb > 0 ? k + 2 : k + 1;
                for (int c = 8 * cb0 + wk * 32 + lane; c < LD; c += NWK) rstepb_column(M, LD, row0, c, pv);
                if (nrb > 0) {
                    asm volatile("bar.sync 1, %0;" :: "n"(NWK > 0 ? NWK : 32) : "memory");
                    const int ntiles = nrb * (NCB - cb0);
                    rupdate_tiles(M, LD, row0, k + 1, nrb, cb0, (ntiles * wk) / (NW - 1), (ntiles * (wk + 1)) / (NW - 1), fo);
                }
            }
        }
        __syncthreads();

        // ---- back substitution, blocks of 8 rows (see sweep_blocked.cu) ----
        for (int c = warp; c < m; c += NW) {
            double y[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int i = s * 32 + lane;
                y[s] = i < R ? M[rphys(i, R + c, LD)] : 0.0;
            }
#pragma unroll
            for (int ks = SLOTS - 1; ks >= 0; --ks) {
#pragma unroll (SLOTS <= 2 ? 4 : 1)
                for (int kq = 3; kq >= 0; --kq) {
                    const int kb8 = ks * 32 + kq * 8;
                    if (kb8 < R) {
                        double x[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = __shfl_sync(FULLM, y[ks], kq * 8 + j);
#pragma unroll
                        for (int j = 7; j >= 0; --j) {
                            const double* urow = M + (kb8 + j) * LD + kb8;
                            const int sw = rswz(j);
#pragma unroll
                            for (int jj = 7; jj > j; --jj) x[j] = fma(-urow[jj ^ sw], x[jj], x[j]);   // newest unknown last (dtrsm order)
                            x[j] *= urow[j ^ sw];                    // reciprocal pivot on the diagonal
                        }
#pragma unroll
                        for (int s = 0; s <= ks; ++s) {
                            const int i = s * 32 + lane;
                            if (i < kb8) {
                                const double* urow = M + i * LD + kb8;
                                const int sw = rswz(i & 7);
#pragma unroll
                                for (int j = 0; j < 8; ++j) y[s] = fma(-urow[j ^ sw], x[j], y[s]);
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
                    M[rphys(i, R + c, LD)] = y[s];
                    if (p.X) p.X[(pt * r + i) * m + c] = y[s];
                }
            }
        }
        __syncthreads();
        if (p.info && tid == 0) p.info[pt] = *info_sh;

        // ---- impedance matrix Z = j zs x^T (cb Br) -> S (finished by gsm_finish) ----
        if (p.S) {
            for (int e = warp; e < m * m; e += NW) {
                const int a = e / m, b = e - a * m;
                double acc = 0.0;
                for (int k = lane; k < r; k += 32) acc = fma(M[rphys(k, R + a, LD)], cb * __ldg(p.Br + (long long)k * p.ldb + b), acc);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULLM, acc, off);
                if (lane == 0) p.S[pt * (long long)m * m + e] = cmake(0.0, p.zs[pt] * acc);
            }
        }
        __syncthreads();
    }
}

template <int MMAX>
__global__ void __launch_bounds__(128) rgsm_finish_kernel(cplx* __restrict__ S, int m, long long F) {
    const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= F) return;
    cplx z[MMAX * MMAX], scratch[2 * MMAX * MMAX];
    cplx* sp = S + pt * (long long)m * m;
    for (int e = 0; e < m * m; ++e) z[e] = sp[e];
    gsm_from_impedance(z, scratch, m, sp);
}

struct RGeom { int R, NCB; size_t smem; };

RGeom rgeom(int r, int m) {
    RGeom gm;
    gm.R = (r + 7) / 8 * 8;
    gm.NCB = gm.R / 8 + (m + 7) / 8;
    gm.smem = sizeof(double) * ((size_t)gm.R * gm.NCB * 8) + 128;
    return gm;
}

template <int SLOTS, int NW, int MINB>
int rlaunch(const SweepParamsR& p, const RGeom& gm, cudaStream_t stream) {
    auto kern = sweep_blocked_f64_kernel<SLOTS, NW, MINB>;
    MF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gm.smem));
    int per_sm = 0;
    MF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, gm.smem));
    if (per_sm < 1) MF_FAIL_ARG(7, "real blocked sweep does not fit on an SM for this (r, m)");
    long long grid = (long long)mf_num_sms() * per_sm;
    if (grid > p.F) grid = p.F;
    kern<<<(unsigned)grid, NW * 32, gm.smem, stream>>>(p, gm.R, gm.NCB);
    MF_CHECK_LAUNCH();
    if (p.S) {
        const unsigned blocks = (unsigned)((p.F + 127) / 128);
        if (p.m <= 2) rgsm_finish_kernel<2><<<blocks, 128, 0, stream>>>(p.S, p.m, p.F);
        else if (p.m <= 4) rgsm_finish_kernel<4><<<blocks, 128, 0, stream>>>(p.S, p.m, p.F);
        else if (p.m <= 8) rgsm_finish_kernel<8><<<blocks, 128, 0, stream>>>(p.S, p.m, p.F);
        else rgsm_finish_kernel<MF_MAX_PORTS><<<blocks, 128, 0, stream>>>(p.S, p.m, p.F);
        MF_CHECK_LAUNCH();
    }
    return 0;
}

}  // namespace

// sweep_left.cu: the left-looking streamed LU on float64 elements (matrices that do not fit in shared memory, r up to 512)
bool sweep_left_supports_f64(int r, int m);
size_t sweep_left_ws_bytes_f64(int r, int m, long long F);
int sweep_left_launch_f64(const double* A0, const double* A1, const double* A2, long long lda, const double* Br, long long ldb, int r, int m,
                          const double* c0, const double* c1, const double* c2, const double* cb, const double* zs, long long F,
                          double* X, cplx* S, int* info, void* ws, size_t ws_bytes, cudaStream_t stream);

static bool rblocked_fits(int r, int m) {
    if (r < 1 || m < 1 || m > MF_MAX_PORTS) return false;
    const RGeom gm = rgeom(r, m);
    return gm.R <= 128 && gm.smem <= 226 * 1024;
}

static bool f64_uses_left(int r, int m, int variant) {
    if (variant == 5) return true;
    if (variant == 3) return false;
    if (getenv("MF_SWEEP_FORCE_SMEM") && rblocked_fits(r, m)) return false;
    // measured crossover (profiles/r02_sweep_crossover.md): shared-memory kernel up to r = 72, left-looking kernel above
    return !rblocked_fits(r, m) || ((r > 72 || getenv("MF_SWEEP_FORCE_LEFT")) && sweep_left_supports_f64(r, m));
}

extern "C" int mf_sweep_f64_supported(int r, int m) {
    return (rblocked_fits(r, m) || sweep_left_supports_f64(r, m)) ? 1 : 0;
}

extern "C" int mf_sweep_f64_variant_supported(int r, int m, int variant) {
    if (variant == 0) return mf_sweep_f64_supported(r, m);
    if (variant == 3) return rblocked_fits(r, m) ? 1 : 0;
    if (variant == 5) return sweep_left_supports_f64(r, m) ? 1 : 0;
    return 0;
}

extern "C" size_t mf_sweep_f64_ws_bytes(int r, int m, int64_t F, int variant) {
    if (r <= 0 || m <= 0 || F <= 0 || !mf_sweep_f64_variant_supported(r, m, variant)) return 256;
    const size_t need = f64_uses_left(r, m, variant) ? sweep_left_ws_bytes_f64(r, m, F) : 0;
    return need < 256 ? 256 : need;
}

extern "C" int mf_sweep_lu_gsm_f64(const double* A0, const double* A1, const double* A2, int64_t lda,
                                   const double* Br, int64_t ldb, int r, int m,
                                   const double* c0, const double* c1, const double* c2, const double* cb,
                                   const double* zscale, int64_t F, double* X, mf_c128* S, int* info, int variant,
                                   void* ws, size_t ws_bytes, void* stream) {
    if (!A0 && !A1 && !A2) MF_FAIL_ARG(1, "all three operators are NULL");
    if (lda < r) MF_FAIL_ARG(4, "lda < r");
    if (!Br || ldb < m) MF_FAIL_ARG(5, "Br is NULL or ldb < m");
    if (variant != 0 && variant != 3 && variant != 5) MF_FAIL_ARG(18, "variant must be 0 (auto), 3 (shared-memory blocked) or 5 (left-looking)");
    if (!mf_sweep_f64_variant_supported(r, m, variant)) MF_FAIL_ARG(7, "(r, m) not supported by this variant of the real sweep (mf_sweep_f64_variant_supported): use the complex128 entry");
    if (!c0 || !c1 || !c2) MF_FAIL_ARG(9, "coefficient arrays must not be NULL");
    if (!cb) MF_FAIL_ARG(12, "cb is NULL");
    if (S && !zscale) MF_FAIL_ARG(13, "zscale is NULL but S is requested");
    if (F < 0) MF_FAIL_ARG(14, "F < 0");
    if (!X && !S) MF_FAIL_ARG(15, "neither X nor S requested");
    if (F == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (f64_uses_left(r, m, variant))
        return sweep_left_launch_f64(A0, A1, A2, lda, Br, ldb, r, m, c0, c1, c2, cb, zscale, F, X, (cplx*)S, info, ws, ws_bytes, st);
    SweepParamsR p;
    p.A0 = A0; p.A1 = A1; p.A2 = A2; p.lda = lda; p.Br = Br; p.ldb = ldb; p.r = r; p.m = m;
    p.c0 = c0; p.c1 = c1; p.c2 = c2; p.cb = cb; p.zs = zscale; p.F = F; p.X = X; p.S = (cplx*)S; p.info = info;
    const RGeom gm = rgeom(r, m);
    if (gm.R <= 32) return rlaunch<1, 2, 12>(p, gm, st);
    if (gm.R <= 64) return rlaunch<2, 4, 5>(p, gm, st);
    if (gm.R <= 96) return rlaunch<3, 8, 2>(p, gm, st);
    return rlaunch<4, 8, 1>(p, gm, st);
}
