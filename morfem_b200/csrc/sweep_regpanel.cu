// Variant 2 of the batched reduced sweep: register-resident LU for r <= 64.
//
// One CTA per frequency point (persistent, grid-stride over points).  The T x T thread grid (T = 8 for r <= 32,
// T = 16 above) holds the augmented matrix [A(t) | cb(t) Br] in REGISTERS, distributed 2-D cyclically: thread
// (ti, tj) owns rows {a T + ti} x columns {b T + tj}.  The cyclic layout keeps every thread busy as the trailing
// matrix shrinks, and the row/column slot loops are compile-time so the triangular work saving is real.
//
// Per elimination step k (LAPACK getf2 order, implementation.py:477 `lu_factor`):
//   * the T threads that own column k sit in one warp: pivot search (max |re|+|im|, first maximum, as izamax) is a
//     quarter/half-warp shuffle reduction; pivot value and the swapped-out entry travel by shuffle as well;
//   * multipliers (scaled by the reciprocal pivot, as zgetf2 does) go to shared memory, double buffered;
//   * the pivot row is broadcast through shared memory -- and simply stays there: the rows of U (and of the
//     eliminated right-hand sides) accumulate in smem as a by-product, which is all back-substitution needs;
//     L is never needed again because the right-hand sides are eliminated alongside;
//   * rank-1 update from registers: RS + CS + MS shared-memory reads feed RS x (CS + MS) complex FMAs.
// Back-substitution (`lu_solve`, implementation.py:478) runs one warp per right-hand side out of shared memory with
// the solution vector in registers and shuffles for the broadcast; the S-parameter epilogue (test_helpers.py:9-14)
// is fused.  Roofline: FP64 pipe; inputs (three r x r operators) are L2 resident, output is 16 M^2 bytes per point.
#include "sweep_common.cuh"

namespace {

template <int T, int RS, int MS>
struct RegPanelCfg {
    static constexpr int CS = RS;
    static constexpr int NT = T * T;
    static constexpr int RMAX = T * RS;
    static constexpr int MMAX = T * MS;
};

template <int T, int RS, int MS>
__global__ void __launch_bounds__(T * T) sweep_regpanel_kernel(SweepParams p, int ldu) {
    constexpr int CS = RS, NT = T * T, RMAX = T * RS, NC = CS + MS;
    constexpr int PER = (RMAX + 31) / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int r = p.r, m = p.m, tid = threadIdx.x;
    const int ti = tid % T, tj = tid / T;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = (NT + 31) / 32;
    const unsigned gmask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane & ~(T - 1)));
    const int gbase = lane & ~(T - 1);

    cplx* Usm = reinterpret_cast<cplx*>(smem_raw);          // r x ldu : rows of U | eliminated rhs
    cplx* Krow = Usm + (size_t)r * ldu;                     // ldu
    cplx* Lcol = Krow + ldu;                                // 2 x RMAX
    cplx* zmat = Lcol + 2 * RMAX;                           // m*m
    cplx* zscr = zmat + m * m;                              // 2*m*m
    int* ish = reinterpret_cast<int*>(zscr + 2 * m * m);    // [0..1] pivot row (double buffered), [2] info

    double ar[RS][NC], ai[RS][NC];

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        // ---- assemble this thread's entries of [A(t) | cb Br] ----
#pragma unroll
        for (int a = 0; a < RS; ++a) {
            const int i = a * T + ti;
#pragma unroll
            for (int b = 0; b < CS; ++b) {
                const int j = b * T + tj;
                double re = 0.0, im = 0.0;
                if (i < r && j < r) {
                    const long long off = (long long)i * p.lda + j;
                    if (p.A0) { cplx v = __ldg(p.A0 + off); re = c0 * v.x; im = c0 * v.y; }
                    if (p.A1) { cplx v = __ldg(p.A1 + off); re = fma(c1, v.x, re); im = fma(c1, v.y, im); }
                    if (p.A2) { cplx v = __ldg(p.A2 + off); re = fma(c2, v.x, re); im = fma(c2, v.y, im); }
                }
                ar[a][b] = re; ai[a][b] = im;
            }
#pragma unroll
            for (int c = 0; c < MS; ++c) {
                const int j = c * T + tj;
                double re = 0.0, im = 0.0;
                if (i < r && j < m) { cplx v = __ldg(p.Br + (long long)i * p.ldb + j); re = cb * v.x; im = cb * v.y; }
                ar[a][CS + c] = re; ai[a][CS + c] = im;
            }
        }
        if (tid == 0) ish[2] = 0;
        __syncthreads();

        // ---- LU with partial pivoting; right-hand sides eliminated alongside ----
#pragma unroll
        for (int kb = 0; kb < CS; ++kb) {
            for (int kj = 0; kj < T; ++kj) {
                const int k = kb * T + kj;
                if (k >= r) break;
                const int buf = k & 1;
                if (tj == kj) {
                    // pivot search over rows >= k of column k (slots a >= kb; in slot kb only ti >= kj)
                    double best = -1.0; int bi = 0x7fffffff;
#pragma unroll
                    for (int a = kb; a < RS; ++a) {
                        const bool ok = (a > kb) || (ti >= kj);
                        const double v = fabs(ar[a][kb]) + fabs(ai[a][kb]);
                        if (ok && v > best) { best = v; bi = a * T + ti; }
                    }
#pragma unroll
                    for (int off = T / 2; off > 0; off >>= 1) {
                        const double ov = __shfl_xor_sync(gmask, best, off);
                        const int oi = __shfl_xor_sync(gmask, bi, off);
                        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                    }
                    const int prow = bi;
                    const int pslot = prow / T, plane = prow - pslot * T;
                    double sr = 0.0, si = 0.0;
#pragma unroll
                    for (int a = kb; a < RS; ++a) if (a == pslot) { sr = ar[a][kb]; si = ai[a][kb]; }
                    const double pr = __shfl_sync(gmask, sr, gbase + plane), pim = __shfl_sync(gmask, si, gbase + plane);
                    const double kr = __shfl_sync(gmask, ar[kb][kb], gbase + kj), kim = __shfl_sync(gmask, ai[kb][kb], gbase + kj);
                    if (prow != k && ti == plane) {
#pragma unroll
                        for (int a = kb; a < RS; ++a) if (a == pslot) { ar[a][kb] = kr; ai[a][kb] = kim; }
                    }
                    const bool zero = !(best > 0.0);
                    const cplx inv = crecip(cmake(pr, pim));
                    // multipliers of rows > k (zgetf2 scales by the reciprocal; a zero pivot leaves the column as is)
#pragma unroll
                    for (int a = kb; a < RS; ++a) {
                        const int i = a * T + ti;
                        if ((a > kb) || (ti > kj)) {
                            cplx e = cmake(ar[a][kb], ai[a][kb]);
                            if (!zero) e = cmul(e, inv);
                            ar[a][kb] = e.x; ai[a][kb] = e.y;
                            Lcol[buf * RMAX + i] = e;
                        }
                    }
                    if (ti == kj) {
                        Usm[(size_t)k * ldu + k] = inv;          // reciprocal diagonal for the back substitution
                        ish[buf] = prow;
                        if (zero && ish[2] == 0) ish[2] = k + 1;
                    }
                }
                __syncthreads();
                const int prow = ish[buf];
                const int pslot = prow / T, plane = prow - pslot * T;
                // pivot row -> Usm[k][k+1 ..] (stays there), old row k -> Krow (moves to the pivot's position)
                if (ti == plane) {
#pragma unroll
                    for (int a = kb; a < RS; ++a) {
                        if (a == pslot) {
#pragma unroll
                            for (int b = kb; b < CS; ++b) {
                                const int j = b * T + tj;
                                if (j > k && j < r) Usm[(size_t)k * ldu + j] = cmake(ar[a][b], ai[a][b]);
                            }
#pragma unroll
                            for (int c = 0; c < MS; ++c) {
                                const int j = c * T + tj;
                                if (j < m) Usm[(size_t)k * ldu + r + j] = cmake(ar[a][CS + c], ai[a][CS + c]);
                            }
                        }
                    }
                }
                if (prow != k && ti == kj) {
#pragma unroll
                    for (int b = kb; b < CS; ++b) {
                        const int j = b * T + tj;
                        if (j > k && j < r) Krow[j] = cmake(ar[kb][b], ai[kb][b]);
                    }
#pragma unroll
                    for (int c = 0; c < MS; ++c) {
                        const int j = c * T + tj;
                        if (j < m) Krow[r + j] = cmake(ar[kb][CS + c], ai[kb][CS + c]);
                    }
                }
                __syncthreads();
                if (prow != k && ti == plane) {
#pragma unroll
                    for (int a = kb; a < RS; ++a) {
                        if (a == pslot) {
#pragma unroll
                            for (int b = kb; b < CS; ++b) {
                                const int j = b * T + tj;
                                if (j > k && j < r) { const cplx v = Krow[j]; ar[a][b] = v.x; ai[a][b] = v.y; }
                            }
#pragma unroll
                            for (int c = 0; c < MS; ++c) {
                                const int j = c * T + tj;
                                if (j < m) { const cplx v = Krow[r + j]; ar[a][CS + c] = v.x; ai[a][CS + c] = v.y; }
                            }
                        }
                    }
                }
                // ---- rank-1 update of the trailing block ----
                double lr[RS], li[RS];
#pragma unroll
                for (int a = kb; a < RS; ++a) {
                    const int i = a * T + ti;
                    cplx l = cmake(0.0, 0.0);
                    if (((a > kb) || (ti > kj)) && i < r) l = Lcol[buf * RMAX + i];
                    lr[a] = l.x; li[a] = l.y;
                }
#pragma unroll
                for (int b = kb; b < CS; ++b) {
                    const int j = b * T + tj;
                    if (j > k && j < r) {
                        const cplx u = Usm[(size_t)k * ldu + j];
#pragma unroll
                        for (int a = kb; a < RS; ++a) {
                            ar[a][b] = fma(-lr[a], u.x, ar[a][b]); ar[a][b] = fma(li[a], u.y, ar[a][b]);
                            ai[a][b] = fma(-lr[a], u.y, ai[a][b]); ai[a][b] = fma(-li[a], u.x, ai[a][b]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < MS; ++c) {
                    const int j = c * T + tj;
                    if (j < m) {
                        const cplx u = Usm[(size_t)k * ldu + r + j];
#pragma unroll
                        for (int a = kb; a < RS; ++a) {
                            ar[a][CS + c] = fma(-lr[a], u.x, ar[a][CS + c]); ar[a][CS + c] = fma(li[a], u.y, ar[a][CS + c]);
                            ai[a][CS + c] = fma(-lr[a], u.y, ai[a][CS + c]); ai[a][CS + c] = fma(-li[a], u.x, ai[a][CS + c]);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- back substitution U x = y, one warp per right-hand side, solution kept in registers ----
        for (int c = warp; c < m; c += NWARPS) {
            double yr[PER], yi[PER];
#pragma unroll
            for (int s = 0; s < PER; ++s) {
                const int i = s * 32 + lane;
                cplx v = cmake(0.0, 0.0);
                if (i < r) v = Usm[(size_t)i * ldu + r + c];
                yr[s] = v.x; yi[s] = v.y;
            }
#pragma unroll
            for (int ks = PER - 1; ks >= 0; --ks) {
                for (int kl = 31; kl >= 0; --kl) {
                    const int k = ks * 32 + kl;
                    if (k >= r) continue;
                    const cplx inv = Usm[(size_t)k * ldu + k];
                    const cplx xk = cmul(cmake(yr[ks], yi[ks]), inv);
                    const double xr = __shfl_sync(0xffffffffu, xk.x, kl), xi = __shfl_sync(0xffffffffu, xk.y, kl);
                    if (lane == kl) { yr[ks] = xr; yi[ks] = xi; }
#pragma unroll
                    for (int s = 0; s <= ks; ++s) {
                        const int i = s * 32 + lane;
                        if (i < k) {
                            const cplx u = Usm[(size_t)i * ldu + k];
                            yr[s] = fma(-u.x, xr, yr[s]); yr[s] = fma(u.y, xi, yr[s]);
                            yi[s] = fma(-u.x, xi, yi[s]); yi[s] = fma(-u.y, xr, yi[s]);
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < PER; ++s) {
                const int i = s * 32 + lane;
                if (i < r) {
                    const cplx x = cmake(yr[s], yi[s]);
                    Usm[(size_t)i * ldu + r + c] = x;
                    if (p.X) p.X[(pt * r + i) * m + c] = x;
                }
            }
        }
        __syncthreads();
        if (p.info && tid == 0) p.info[pt] = ish[2];

        // ---- S-parameters: Z = j zs x^T (cb Br) ----
        if (p.S) {
            for (int e = warp; e < m * m; e += NWARPS) {
                const int a = e / m, b = e - a * m;
                cplx acc = cmake(0.0, 0.0);
                for (int k = lane; k < r; k += 32) cfma(acc, Usm[(size_t)k * ldu + r + a], cscale(cb, __ldg(p.Br + (long long)k * p.ldb + b)));
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                }
                if (lane == 0) { const double zs = p.zs[pt]; zmat[e] = cmake(-zs * acc.y, zs * acc.x); }
            }
            __syncthreads();
            if (tid == 0) gsm_from_impedance(zmat, zscr, m, p.S + pt * (long long)m * m);
        }
        __syncthreads();
    }
}

template <int T, int RS, int MS>
int launch_cfg(const SweepParams& p, cudaStream_t stream) {
    constexpr int RMAX = T * RS;
    int ldu = p.r + p.m; if ((ldu & 1) == 0) ldu += 1;
    const size_t smem = sizeof(cplx) * ((size_t)p.r * ldu + ldu + 2 * RMAX + 3 * p.m * p.m) + 16;
    auto kern = sweep_regpanel_kernel<T, RS, MS>;
    MF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T * T, smem));
    if (per_sm < 1) MF_FAIL_ARG(7, "register-panel sweep does not fit on an SM for this (r, m)");
    long long grid = (long long)mf_num_sms() * per_sm;
    if (grid > p.F) grid = p.F;
    kern<<<(unsigned)grid, T * T, smem, stream>>>(p, ldu);
    MF_CHECK_LAUNCH();
    return 0;
}

}  // namespace

bool sweep_regpanel_supports(int r, int m) { return r >= 1 && r <= 64 && m >= 1 && m <= 16; }

int sweep_regpanel_launch(const SweepParams& p, cudaStream_t stream) {
    const int r = p.r, m = p.m;
    if (r <= 32) {
        const int rs = (r + 7) / 8;
        if (m <= 8) {
            switch (rs) {
                case 1: return launch_cfg<8, 1, 1>(p, stream);
                case 2: return launch_cfg<8, 2, 1>(p, stream);
                case 3: return launch_cfg<8, 3, 1>(p, stream);
                default: return launch_cfg<8, 4, 1>(p, stream);
            }
        }
        switch (rs) {
            case 1: return launch_cfg<8, 1, 2>(p, stream);
            case 2: return launch_cfg<8, 2, 2>(p, stream);
            case 3: return launch_cfg<8, 3, 2>(p, stream);
            default: return launch_cfg<8, 4, 2>(p, stream);
        }
    }
    const int rs = (r + 15) / 16;
    if (rs <= 3) return launch_cfg<16, 3, 1>(p, stream);
    return launch_cfg<16, 4, 1>(p, stream);
}
