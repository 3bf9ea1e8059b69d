// r x r device-resident factorisations used by the basis stage (CholeskyQR2 + SVD of the triangular factor,
// the B200 replacement of `np.linalg.svd(S, full_matrices=False)[0]`, implementation.py:226/298/210) and the
// one-off symmetrisation of the reduced operators (implementation.py:528).  All of them touch at most a few
// MiB that stay in L2; they are latency-, not bandwidth- or flop-bound, so each is a single launch.
#include <stdlib.h>
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------- equilibrate
__global__ void equilibrate_kernel(cplx* G, long long ld, int r, double shift, double* d, double* stats) {
    extern __shared__ double ds[];
    __shared__ double wmax[32];
    double dev = 0.0;
    for (int j = threadIdx.x; j < r; j += blockDim.x) {
        double g = G[j * ld + j].x;
        double s = (g > 0.0 && isfinite(g)) ? 1.0 / sqrt(g) : 1.0;
        ds[j] = s; d[j] = s;
    }
    __syncthreads();
    for (long long idx = threadIdx.x; idx < (long long)r * r; idx += blockDim.x) {
        int i = (int)(idx / r), j = (int)(idx - (long long)i * r);
        cplx v = G[i * ld + j];
        dev = fmax(dev, fabs(v.x - (i == j ? 1.0 : 0.0)) + fabs(v.y));
        double s = ds[i] * ds[j];
        v.x *= s; v.y *= s;
        if (i == j) { v.x += shift; v.y = 0.0; }
        G[i * ld + j] = v;
    }
    for (int off = 16; off > 0; off >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, off));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = dev;
    __syncthreads();
    if (threadIdx.x == 0 && stats) {
        double mx = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, wmax[w]);
        stats[0] = isfinite(mx) ? mx : 1e300;
    }
}

// ------------------------------------------------------------------------------------------------- potrf
// Right-looking Cholesky G = R^H R (R upper, row-major), one CTA.  With SMEM the matrix is staged in shared memory
// (r <= POTRF_SMEM_MAX): every step of the factorisation is a dependent access, which costs an L2 round trip each
// when the matrix stays in global memory.
constexpr int POTRF_SMEM_MAX = 112;
constexpr int POTRF_THREADS = 1024;
template <bool SMEM>
__global__ void __launch_bounds__(POTRF_THREADS) potrf_upper_kernel(cplx* Gg, long long ldg, int r, int* info) {
    extern __shared__ __align__(16) cplx gsm[];
    __shared__ int bad;
    cplx* G = SMEM ? gsm : Gg;
    const long long ld = SMEM ? (long long)(r | 1) : ldg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = POTRF_THREADS / 32;
    if (SMEM) for (int i = warp; i < r; i += nwarps) for (int j = lane; j < r; j += 32) G[i * ld + j] = Gg[i * ldg + j];
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int k = 0; k < r; ++k) {
        const double piv = G[k * ld + k].x;
        if (!(piv > 0.0) || !isfinite(piv)) { if (tid == 0) bad = k + 1; }
        __syncthreads();
        if (bad) break;
        const double inv = rsqrt(piv), rkk = piv * inv;
        for (int j = k + tid; j < r; j += POTRF_THREADS) {
            cplx v = G[k * ld + j];
            if (j == k) v = cmake(rkk, 0.0); else { v.x *= inv; v.y *= inv; }
            G[k * ld + j] = v;
        }
        __syncthreads();
        // trailing update of the upper triangle: one row per warp, lanes along the row
        for (int i = k + 1 + warp; i < r; i += nwarps) {
            const cplx gki = cconj(G[k * ld + i]);
            for (int j = i + lane; j < r; j += 32) {
                cplx a = G[i * ld + j];
                cfms(a, gki, G[k * ld + j]);
                G[i * ld + j] = a;
            }
        }
        __syncthreads();
    }
    for (int i = warp; i < r; i += nwarps)
        for (int j = lane; j < r; j += 32) Gg[i * ldg + j] = (j < i) ? cmake(0.0, 0.0) : G[i * ld + j];
    if (tid == 0 && info) *info = bad;
}

// ------------------------------------------------------------------------------------------------- trtri
// One warp per column j of Rinv = R^-1 (R x = e_j).
//   SMEM (r <= 96): the CTA stages the block of R it needs and the reciprocal diagonal in shared memory; the warp keeps
//   x in registers (row i in lane i % 32, slot i / 32) and sweeps k = j .. 0: x_k <- x_k / r_kk, broadcast by shuffle,
//   then every lane updates its own rows -- no reduction on the dependent chain.
//   otherwise: dot-product form on global memory.
constexpr int TRTRI_SMEM_MAX = 96;
template <bool SMEM>
__global__ void trtri_upper_kernel(const cplx* __restrict__ Rg, long long ldg, int r, cplx* __restrict__ Rinv, long long ldi) {
    extern __shared__ __align__(16) cplx xs_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int j = blockIdx.x * wpb + warp;
    if (SMEM) {
        constexpr int SL = (TRTRI_SMEM_MAX + 31) / 32;
        cplx* Rs = xs_all;                                  // n x lds
        const int jmax = min(r - 1, (int)(blockIdx.x * wpb + wpb - 1)), n = jmax + 1;
        const int lds = r | 1;
        cplx* rd = Rs + (size_t)n * lds;                    // reciprocal diagonal
        for (int i = warp; i < n; i += wpb) for (int k = i + lane; k < n; k += 32) Rs[i * lds + k] = Rg[i * ldg + k];
        for (int i = threadIdx.x; i < n; i += blockDim.x) rd[i] = crecip(Rg[i * ldg + i]);
        __syncthreads();
        if (j >= r) return;
        double xr[SL], xi[SL];
#pragma unroll
        for (int s = 0; s < SL; ++s) { xr[s] = (s * 32 + lane == j) ? 1.0 : 0.0; xi[s] = 0.0; }
#pragma unroll
        for (int ks = SL - 1; ks >= 0; --ks) {
            for (int kl = 31; kl >= 0; --kl) {
                const int k = ks * 32 + kl;
                if (k > j) continue;
                const cplx xk0 = cmul(cmake(xr[ks], xi[ks]), rd[k]);
                const double xkr = __shfl_sync(0xffffffffu, xk0.x, kl), xki = __shfl_sync(0xffffffffu, xk0.y, kl);
                if (lane == kl) { xr[ks] = xkr; xi[ks] = xki; }
#pragma unroll
                for (int s = 0; s <= ks; ++s) {
                    const int i = s * 32 + lane;
                    if (i < k) {
                        const cplx u = Rs[i * lds + k];
                        xr[s] = fma(-u.x, xkr, xr[s]); xr[s] = fma(u.y, xki, xr[s]);
                        xi[s] = fma(-u.x, xki, xi[s]); xi[s] = fma(-u.y, xkr, xi[s]);
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SL; ++s) { const int i = s * 32 + lane; if (i < r) Rinv[i * ldi + j] = (i <= j) ? cmake(xr[s], xi[s]) : cmake(0.0, 0.0); }
        return;
    }
    if (j >= r) return;
    const cplx* R = Rg; const long long ldr = ldg;
    cplx* x = xs_all + (size_t)warp * r;
    if (lane == 0) x[j] = crecip(R[j * ldr + j]);
    __syncwarp();
    for (int i = j - 1; i >= 0; --i) {
        cplx acc = cmake(0.0, 0.0);
        for (int k = i + 1 + lane; k <= j; k += 32) cfma(acc, R[i * ldr + k], x[k]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (lane == 0) { cplx v = cmul(acc, crecip(R[i * ldr + i])); x[i] = cmake(-v.x, -v.y); }
        __syncwarp();
    }
    for (int i = lane; i < r; i += 32) Rinv[i * ldi + j] = (i <= j) ? x[i] : cmake(0.0, 0.0);
}

// ------------------------------------------------------------------------------------------ scale / sym
__global__ void scale_kernel(cplx* X, long long ld, int rows, int cols, const double* d, int power, int by_rows) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    int i = (int)(idx / cols), j = (int)(idx - (long long)i * cols);
    double s = d[by_rows ? i : j];
    if (power < 0) s = 1.0 / s;
    cplx v = X[i * ld + j]; v.x *= s; v.y *= s; X[i * ld + j] = v;
}

__global__ void symmetrize_kernel(const cplx* __restrict__ A, long long lda, int r, cplx* __restrict__ As, long long lds) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= r * r) return;
    int i = idx / r, j = idx - i * r;
    cplx a = A[i * lda + j], b = A[j * lda + i];
    As[i * lds + j] = cmake((a.x + b.x) / 2, (a.y + b.y) / 2);
}

// -------------------------------------------------------------------------------------------- Jacobi SVD
// One-sided (Hestenes) Jacobi on the ROWS of X: left rotations G with G X = Sigma W^H, hence X = G^H Sigma W^H
// and the left singular vectors are the columns of G^H.  r2 = r rounded up to even; CTA c owns one pair per
// round of the round-robin tournament; a device-wide barrier separates rounds (cooperative launch guarantees
// co-residency).  Layout of ws: Xw (r2 x r) | Gacc (r2 x r2) | sig (r2 doubles) | offmax (64 doubles) | counter.
constexpr int JS_THREADS = 128;

__device__ __forceinline__ double block_sum4(double& a, double& b, double& c, double& d, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off);
        c += __shfl_xor_sync(0xffffffffu, c, off); d += __shfl_xor_sync(0xffffffffu, d, off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { red[warp * 4 + 0] = a; red[warp * 4 + 1] = b; red[warp * 4 + 2] = c; red[warp * 4 + 3] = d; }
    __syncthreads();
    a = b = c = d = 0.0;
    for (int w = 0; w < nw; ++w) { a += red[w * 4 + 0]; b += red[w * 4 + 1]; c += red[w * 4 + 2]; d += red[w * 4 + 3]; }
    return a;
}

__global__ void __launch_bounds__(JS_THREADS)
jacobi_svd_kernel(const cplx* __restrict__ Xin, long long ld, int r, cplx* __restrict__ U, long long ldu, double* __restrict__ sigma,
                  int max_sweeps, double tol, int* sweeps_done, cplx* Xw, cplx* Gacc, double* sig, double* offmax, unsigned* counter) {
    __shared__ double red[4 * (JS_THREADS / 32)];
    const int r2 = (r + 1) & ~1;
    const int c = blockIdx.x, nblocks = gridDim.x, tid = threadIdx.x;
    unsigned epoch = 0;
    // init: copy X, identity accumulator
    for (int row = 2 * c; row < 2 * c + 2; ++row) {
        for (int k = tid; k < r; k += JS_THREADS) Xw[(long long)row * r + k] = row < r ? Xin[row * ld + k] : cmake(0.0, 0.0);
        for (int k = tid; k < r2; k += JS_THREADS) Gacc[(long long)row * r2 + k] = cmake(k == row ? 1.0 : 0.0, 0.0);
    }
    if (c == 0) for (int k = tid; k < 64; k += JS_THREADS) offmax[k] = 0.0;
    mf_grid_barrier_counter(counter, epoch, nblocks);

    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < r2 - 1; ++round) {
            int p, q;
            if (c == 0) { p = r2 - 1; q = round; }
            else { p = (round + c) % (r2 - 1); q = (round - c + (r2 - 1)) % (r2 - 1); }
            if (p > q) { int tmp = p; p = q; q = tmp; }
            cplx* xp = Xw + (long long)p * r; cplx* xq = Xw + (long long)q * r;
            double a = 0.0, b = 0.0, cr = 0.0, ci = 0.0;
            for (int k = tid; k < r; k += JS_THREADS) {
                cplx u = __ldcg(xp + k), v = __ldcg(xq + k);   // rows were last written by other CTAs: read at L2
                a += cnorm2(u); b += cnorm2(v);
                cr += u.x * v.x + u.y * v.y;      // u * conj(v)
                ci += u.y * v.x - u.x * v.y;
            }
            block_sum4(a, b, cr, ci, red);
            const double cabs = hypot(cr, ci);
            const double denom = sqrt(a) * sqrt(b);
            const double off = denom > 0.0 ? cabs / denom : 0.0;
            if (off > tol && cabs > 0.0) {
                if (tid == 0) atomicMax((unsigned long long*)&offmax[sweep & 63], (unsigned long long)__double_as_longlong(off));
                const double zeta = (b - a) / (2.0 * cabs);
                const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                const cplx ph = cmake(cr / cabs, ci / cabs);            // e^{i phi}
                const cplx sph = cscale(sn, ph), sphc = cscale(sn, cconj(ph));
                for (int k = tid; k < r; k += JS_THREADS) {
                    cplx u = __ldcg(xp + k), v = __ldcg(xq + k);
                    xp[k] = csub(cscale(cs, u), cmul(sph, v));
                    xq[k] = cadd(cmul(sphc, u), cscale(cs, v));
                }
                cplx* gp = Gacc + (long long)p * r2; cplx* gq = Gacc + (long long)q * r2;
                for (int k = tid; k < r2; k += JS_THREADS) {
                    cplx u = __ldcg(gp + k), v = __ldcg(gq + k);
                    gp[k] = csub(cscale(cs, u), cmul(sph, v));
                    gq[k] = cadd(cmul(sphc, u), cscale(cs, v));
                }
            }
            mf_grid_barrier_counter(counter, epoch, nblocks);
        }
        const double worst = *((volatile double*)&offmax[sweep & 63]);
        if (!(worst > tol)) { ++sweep; break; }
    }
    // singular values = row norms
    for (int row = 2 * c; row < 2 * c + 2; ++row) {
        double a = 0.0, z0 = 0.0, z1 = 0.0, z2 = 0.0;
        for (int k = tid; k < r; k += JS_THREADS) a += cnorm2(__ldcg(Xw + (long long)row * r + k));
        block_sum4(a, z0, z1, z2, red);
        if (tid == 0) sig[row] = sqrt(a);
    }
    mf_grid_barrier_counter(counter, epoch, nblocks);
    for (int row = 2 * c; row < 2 * c + 2; ++row) {
        if (row >= r) continue;
        const double s = __ldcg(sig + row);
        int rank = 0;
        for (int j = 0; j < r; ++j) { double sj = __ldcg(sig + j); rank += (sj > s) || (sj == s && j < row); }
        if (tid == 0) sigma[rank] = s;
        for (int k = tid; k < r; k += JS_THREADS) U[k * ldu + rank] = cconj(__ldcg(Gacc + (long long)row * r2 + k));
    }
    if (c == 0 && tid == 0 && sweeps_done) *sweeps_done = sweep;
}


// Block variant of the cooperative kernel (64 < r <= ~900): the device-wide barrier (~2 us) bounds the kernel above,
// 255 of them per sweep at r = 256.  Here a CTA owns a PAIR OF ROW BLOCKS (B rows each) per tournament round: it stages the
// 2B rows of X and of the rotation accumulator in shared memory, rotates all B*B cross pairs (B inner rounds of B disjoint
// pairs, one WARP per pair, CTA barriers only) and writes the rows back -- r2/B - 1 device-wide barriers per sweep instead
// of r2 - 1.  Pairs inside a block are rotated in round 0 of every sweep (every block is in exactly one pair then), so a
// sweep still visits every row pair exactly once (a cyclic Jacobi ordering; same convergence test as above).
constexpr int JB_THREADS = 256;
template <int B>
__global__ void __launch_bounds__(JB_THREADS)
jacobi_svd_block_kernel(const cplx* __restrict__ Xin, long long ld, int r, int r2, cplx* __restrict__ U, long long ldu,
                        double* __restrict__ sigma, int max_sweeps, double tol, int* sweeps_done, cplx* Xw, cplx* Gacc,
                        double* sig, double* offmax, unsigned* sync_words) {
    extern __shared__ __align__(16) cplx jb_sm[];
    cplx* Xs = jb_sm;                          // 2B x r
    cplx* Gs = Xs + (size_t)2 * B * r;         // 2B x r2
    const int nb = r2 / B, c = blockIdx.x, ncta = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned epoch = 0;
    auto barrier = [&]() {
        if (ncta <= 32) mf_grid_barrier_flags(sync_words, epoch, ncta);
        else mf_grid_barrier_counter(sync_words, epoch, ncta);
    };
    for (int l = 0; l < 2 * B; ++l) {
        const int row = c * 2 * B + l;
        for (int k = tid; k < r; k += JB_THREADS) Xw[(long long)row * r + k] = row < r ? Xin[row * ld + k] : cmake(0.0, 0.0);
        for (int k = tid; k < r2; k += JB_THREADS) Gacc[(long long)row * r2 + k] = cmake(k == row ? 1.0 : 0.0, 0.0);
    }
    if (c == 0) for (int k = tid; k < 64; k += JB_THREADS) offmax[k] = 0.0;
    barrier();

    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < nb - 1; ++round) {
            int P, Q;
            if (c == 0) { P = nb - 1; Q = round; }
            else { P = (round + c) % (nb - 1); Q = (round - c + (nb - 1)) % (nb - 1); }
            if (P > Q) { const int tmp = P; P = Q; Q = tmp; }
            auto grow = [&](const int l) { return l < B ? P * B + l : Q * B + (l - B); };   // local -> global row (ascending)
            for (int l = warp; l < 2 * B; l += JB_THREADS / 32) {
                const cplx* xsrc = Xw + (long long)grow(l) * r;
                const cplx* gsrc = Gacc + (long long)grow(l) * r2;
#pragma unroll 8
                for (int k = lane; k < r; k += 32) Xs[l * r + k] = __ldcg(reinterpret_cast<const double2*>(xsrc + k));
#pragma unroll 8
                for (int k = lane; k < r2; k += 32) Gs[l * r2 + k] = __ldcg(reinterpret_cast<const double2*>(gsrc + k));
            }
            __syncthreads();
            double offloc = 0.0;
            auto rotate_pair = [&](const int la, const int lb) {       // local rows la < lb, one warp
                cplx* xp = Xs + la * r; cplx* xq = Xs + lb * r;
                double a = 0.0, b = 0.0, cr = 0.0, ci = 0.0;
                for (int k = lane; k < r; k += 32) {
                    const cplx u = xp[k], v = xq[k];
                    a += cnorm2(u); b += cnorm2(v);
                    cr += u.x * v.x + u.y * v.y;      // u * conj(v)
                    ci += u.y * v.x - u.x * v.y;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off);
                    cr += __shfl_xor_sync(0xffffffffu, cr, off); ci += __shfl_xor_sync(0xffffffffu, ci, off);
                }
                const double cabs = hypot(cr, ci);
                const double denom = sqrt(a) * sqrt(b);
                const double off = denom > 0.0 ? cabs / denom : 0.0;
                if (off > tol && cabs > 0.0) {
                    offloc = fmax(offloc, off);
                    const double zeta = (b - a) / (2.0 * cabs);
                    const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                    const cplx ph = cmake(cr / cabs, ci / cabs);            // e^{i phi}
                    const cplx sph = cscale(sn, ph), sphc = cscale(sn, cconj(ph));
                    for (int k = lane; k < r; k += 32) {
                        const cplx u = xp[k], v = xq[k];
                        xp[k] = csub(cscale(cs, u), cmul(sph, v));
                        xq[k] = cadd(cmul(sphc, u), cscale(cs, v));
                    }
                    cplx* gp = Gs + la * r2; cplx* gq = Gs + lb * r2;
                    for (int k = lane; k < r2; k += 32) {
                        const cplx u = gp[k], v = gq[k];
                        gp[k] = csub(cscale(cs, u), cmul(sph, v));
                        gq[k] = cadd(cmul(sphc, u), cscale(cs, v));
                    }
                }
            };
            if (round == 0) {                  // pairs inside the two blocks: a (B - 1)-round tournament in each, side by side
                for (int ir = 0; ir < B - 1; ++ir) {
                    if (warp < B) {
                        const int h = warp / (B / 2), sl = warp % (B / 2);
                        int x, y;
                        if (sl == 0) { x = B - 1; y = ir; }
                        else { x = (ir + sl) % (B - 1); y = (ir - sl + (B - 1)) % (B - 1); }
                        if (x > y) { const int tmp = x; x = y; y = tmp; }
                        rotate_pair(h * B + x, h * B + y);
                    }
                    __syncthreads();
                }
            }
            for (int sft = 0; sft < B; ++sft) {                           // the B * B cross pairs
                if (warp < B) rotate_pair(warp, B + (warp + sft) % B);
                __syncthreads();
            }
            for (int l = warp; l < 2 * B; l += JB_THREADS / 32) {
                cplx* xdst = Xw + (long long)grow(l) * r;
                cplx* gdst = Gacc + (long long)grow(l) * r2;
                for (int k = lane; k < r; k += 32) xdst[k] = Xs[l * r + k];
                for (int k = lane; k < r2; k += 32) gdst[k] = Gs[l * r2 + k];
            }
            if (lane == 0 && offloc > 0.0) atomicMax((unsigned long long*)&offmax[sweep & 63], (unsigned long long)__double_as_longlong(offloc));
            barrier();
        }
        const double worst = *((volatile double*)&offmax[sweep & 63]);
        if (!(worst > tol)) { ++sweep; break; }
    }
    // singular values = row norms (this CTA: its 2B consecutive rows)
    for (int l = warp; l < 2 * B; l += JB_THREADS / 32) {
        const int row = c * 2 * B + l;
        double a = 0.0;
        for (int k = lane; k < r; k += 32) a += cnorm2(__ldcg(reinterpret_cast<const double2*>(Xw + (long long)row * r + k)));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) sig[row] = sqrt(a);
    }
    barrier();
    for (int l = warp; l < 2 * B; l += JB_THREADS / 32) {
        const int row = c * 2 * B + l;
        if (row >= r) continue;
        const double sv = __ldcg(sig + row);
        int rank = 0;
        for (int j = lane; j < r; j += 32) { const double sj = __ldcg(sig + j); rank += (sj > sv) || (sj == sv && j < row); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, off);
        if (lane == 0) sigma[rank] = sv;
        for (int k = lane; k < r; k += 32) U[k * ldu + rank] = cconj(__ldcg(reinterpret_cast<const double2*>(Gacc + (long long)row * r2 + k)));
    }
    if (c == 0 && tid == 0 && sweeps_done) *sweeps_done = sweep;
}


// Single-CTA variant for r <= 64: the whole problem (X and the rotation accumulator) lives in shared memory, one WARP
// per row pair, __syncthreads between tournament rounds instead of a device-wide barrier (~50 ns instead of ~2 us).
constexpr int JS_SMEM_MAX = 64;
__global__ void __launch_bounds__(1024)
jacobi_svd_smem_kernel(const cplx* __restrict__ Xin, long long ld, int r, cplx* __restrict__ U, long long ldu, double* __restrict__ sigma,
                       int max_sweeps, double tol, int* sweeps_done) {
    extern __shared__ __align__(16) cplx js_sm[];
    __shared__ double sig[JS_SMEM_MAX];
    __shared__ int rotated[2];
    const int r2 = (r + 1) & ~1, ldx = r | 1, ldg = r2 | 1;
    cplx* Xw = js_sm;                    // r2 x ldx
    cplx* Gacc = Xw + (size_t)r2 * ldx;  // r2 x ldg
    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5, nthreads = blockDim.x;
    for (int idx = tid; idx < r2 * r; idx += nthreads) { int i = idx / r, k = idx - i * r; Xw[i * ldx + k] = i < r ? Xin[i * ld + k] : cmake(0.0, 0.0); }
    for (int idx = tid; idx < r2 * r2; idx += nthreads) { int i = idx / r2, k = idx - i * r2; Gacc[i * ldg + k] = cmake(i == k ? 1.0 : 0.0, 0.0); }
    if (tid < 2) rotated[tid] = 0;
    __syncthreads();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < r2 - 1; ++round) {
            int p, q;
            if (c == 0) { p = r2 - 1; q = round; }
            else { p = (round + c) % (r2 - 1); q = (round - c + (r2 - 1)) % (r2 - 1); }
            if (p > q) { int tmp = p; p = q; q = tmp; }
            cplx* xp = Xw + p * ldx; cplx* xq = Xw + q * ldx;
            double a = 0.0, b = 0.0, cr = 0.0, ci = 0.0;
            for (int k = lane; k < r; k += 32) {
                const cplx u = xp[k], v = xq[k];
                a += cnorm2(u); b += cnorm2(v);
                cr += u.x * v.x + u.y * v.y;      // u * conj(v)
                ci += u.y * v.x - u.x * v.y;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off);
                cr += __shfl_xor_sync(0xffffffffu, cr, off); ci += __shfl_xor_sync(0xffffffffu, ci, off);
            }
            // rotation parameters with as few FP64 divisions / square roots as possible: all 32 warps of the CTA share
            // one SM's FP64 pipe, so this scalar chain is the cost of a round
            const double cabs = sqrt(fma(cr, cr, ci * ci));
            const double denom = sqrt(a * b);
            if (cabs > tol * denom && cabs > 0.0) {
                if (lane == 0) rotated[sweep & 1] = 1;
                const double icabs = 1.0 / cabs;
                const double zeta = 0.5 * (b - a) * icabs;
                const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
                const double cs = rsqrt(fma(tt, tt, 1.0)), sn = cs * tt;
                const cplx ph = cmake(cr * icabs, ci * icabs);          // e^{i phi}
                const cplx sph = cscale(sn, ph), sphc = cscale(sn, cconj(ph));
                for (int k = lane; k < r; k += 32) {
                    const cplx u = xp[k], v = xq[k];
                    xp[k] = csub(cscale(cs, u), cmul(sph, v));
                    xq[k] = cadd(cmul(sphc, u), cscale(cs, v));
                }
                cplx* gp = Gacc + p * ldg; cplx* gq = Gacc + q * ldg;
                for (int k = lane; k < r2; k += 32) {
                    const cplx u = gp[k], v = gq[k];
                    gp[k] = csub(cscale(cs, u), cmul(sph, v));
                    gq[k] = cadd(cmul(sphc, u), cscale(cs, v));
                }
            }
            __syncthreads();
        }
        const int any = rotated[sweep & 1];
        if (tid == 0) rotated[(sweep + 1) & 1] = 0;
        __syncthreads();
        if (!any) { ++sweep; break; }
    }
    // singular values = row norms; order by descending sigma (stable), U = G^H
    for (int row = c; row < r2; row += (nthreads >> 5)) {
        double a = 0.0;
        for (int k = lane; k < r; k += 32) a += cnorm2(Xw[row * ldx + k]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) sig[row] = sqrt(a);
    }
    __syncthreads();
    for (int row = c; row < r; row += (nthreads >> 5)) {
        const double sv = sig[row];
        int rank = 0;
        for (int j = 0; j < r; ++j) { const double sj = sig[j]; rank += (sj > sv) || (sj == sv && j < row); }
        if (lane == 0) sigma[rank] = sv;
        for (int k = lane; k < r; k += 32) U[k * ldu + rank] = cconj(Gacc[row * ldg + k]);
    }
    if (tid == 0 && sweeps_done) *sweeps_done = sweep;
}


// Real float64 twin of the single-CTA kernel (r <= 64): the triangular factor of a real snapshot block is real, and a
// real plane rotation costs half the FP64 work of the complex one -- which is what bounds this kernel (one SM's FP64
// pipe shared by 32 warps).
__global__ void __launch_bounds__(1024)
jacobi_svd_smem_f64_kernel(const double* __restrict__ Xin, long long ld, int r, double* __restrict__ U, long long ldu,
                           double* __restrict__ sigma, int max_sweeps, double tol, int* sweeps_done) {
    extern __shared__ __align__(16) double jsd_sm[];
    __shared__ double sig[JS_SMEM_MAX];
    __shared__ int rotated[2];
    const int r2 = (r + 1) & ~1, ldx = r | 1, ldg = r2 | 1;
    double* Xw = jsd_sm;                    // r2 x ldx
    double* Gacc = Xw + (size_t)r2 * ldx;   // r2 x ldg
    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5, nthreads = blockDim.x;
    for (int idx = tid; idx < r2 * r; idx += nthreads) { int i = idx / r, k = idx - i * r; Xw[i * ldx + k] = i < r ? Xin[i * ld + k] : 0.0; }
    for (int idx = tid; idx < r2 * r2; idx += nthreads) { int i = idx / r2, k = idx - i * r2; Gacc[i * ldg + k] = (i == k) ? 1.0 : 0.0; }
    if (tid < 2) rotated[tid] = 0;
    __syncthreads();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < r2 - 1; ++round) {
            int p, q;
            if (c == 0) { p = r2 - 1; q = round; }
            else { p = (round + c) % (r2 - 1); q = (round - c + (r2 - 1)) % (r2 - 1); }
            if (p > q) { int tmp = p; p = q; q = tmp; }
            double* xp = Xw + p * ldx; double* xq = Xw + q * ldx;
            double a = 0.0, b = 0.0, cr = 0.0;
            for (int k = lane; k < r; k += 32) { const double u = xp[k], v = xq[k]; a = fma(u, u, a); b = fma(v, v, b); cr = fma(u, v, cr); }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); cr += __shfl_xor_sync(0xffffffffu, cr, off);
            }
            const double cabs = fabs(cr);
            if (cabs > tol * sqrt(a * b) && cabs > 0.0) {
                if (lane == 0) rotated[sweep & 1] = 1;
                const double zeta = 0.5 * (b - a) / cabs;
                const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
                const double cs = rsqrt(fma(tt, tt, 1.0));
                const double sn = (cr >= 0.0 ? cs * tt : -cs * tt);     // sn * sign(cr): the complex kernel's sn * e^{i phi}
                for (int k = lane; k < r; k += 32) {
                    const double u = xp[k], v = xq[k];
                    xp[k] = fma(cs, u, -sn * v);
                    xq[k] = fma(sn, u, cs * v);
                }
                double* gp = Gacc + p * ldg; double* gq = Gacc + q * ldg;
                for (int k = lane; k < r2; k += 32) {
                    const double u = gp[k], v = gq[k];
                    gp[k] = fma(cs, u, -sn * v);
                    gq[k] = fma(sn, u, cs * v);
                }
            }
            __syncthreads();
        }
        const int any = rotated[sweep & 1];
        if (tid == 0) rotated[(sweep + 1) & 1] = 0;
        __syncthreads();
        if (!any) { ++sweep; break; }
    }
    for (int row = c; row < r2; row += (nthreads >> 5)) {
        double a = 0.0;
        for (int k = lane; k < r; k += 32) { const double v = Xw[row * ldx + k]; a = fma(v, v, a); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) sig[row] = sqrt(a);
    }
    __syncthreads();
    for (int row = c; row < r; row += (nthreads >> 5)) {
        const double sv = sig[row];
        int rank = 0;
        for (int j = 0; j < r; ++j) { const double sj = sig[j]; rank += (sj > sv) || (sj == sv && j < row); }
        if (lane == 0) sigma[rank] = sv;
        for (int k = lane; k < r; k += 32) U[k * ldu + rank] = Gacc[row * ldg + k];
    }
    if (tid == 0 && sweeps_done) *sweeps_done = sweep;
}

}  // namespace

extern "C" int mf_equilibrate_c128(mf_c128* G, int64_t ld, int r, double shift, double* d, double* stats, void* stream) {
    if (!G || ld < r) MF_FAIL_ARG(1, "G is NULL or ld < r");
    if (r <= 0 || r > 4096) MF_FAIL_ARG(3, "need 0 < r <= 4096");
    if (!d) MF_FAIL_ARG(5, "d is NULL");
    equilibrate_kernel<<<1, 1024, sizeof(double) * r, (cudaStream_t)stream>>>((cplx*)G, ld, r, shift, d, stats);
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_potrf_upper_c128(mf_c128* G, int64_t ld, int r, int* info, void* stream) {
    if (!G || ld < r) MF_FAIL_ARG(1, "G is NULL or ld < r");
    if (r <= 0) MF_FAIL_ARG(3, "r <= 0");
    if (r <= POTRF_SMEM_MAX) {
        const size_t smem = sizeof(cplx) * (size_t)r * (r | 1);
        MF_CHECK_CUDA(cudaFuncSetAttribute(potrf_upper_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        potrf_upper_kernel<true><<<1, POTRF_THREADS, smem, (cudaStream_t)stream>>>((cplx*)G, ld, r, info);
    } else {
        potrf_upper_kernel<false><<<1, POTRF_THREADS, 0, (cudaStream_t)stream>>>((cplx*)G, ld, r, info);
    }
    MF_CHECK_LAUNCH();
    return 0;
}

extern "C" int mf_trtri_upper_c128(const mf_c128* R, int64_t ldr, int r, mf_c128* Rinv, int64_t ldi, void* stream) {
    if (!R || ldr < r) MF_FAIL_ARG(1, "R is NULL or ldr < r");
    if (r <= 0 || r > 1024) MF_FAIL_ARG(3, "need 0 < r <= 1024");
    if (!Rinv || ldi < r) MF_FAIL_ARG(4, "Rinv is NULL or ldi < r");
    if ((const void*)R == (const void*)Rinv) MF_FAIL_ARG(4, "Rinv must not alias R");
    const int wpb = 4;
    if (r <= TRTRI_SMEM_MAX) {
        const size_t smem = sizeof(cplx) * ((size_t)r * (r | 1) + (size_t)r);
        MF_CHECK_CUDA(cudaFuncSetAttribute(trtri_upper_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trtri_upper_kernel<true><<<(r + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>((const cplx*)R, ldr, r, (cplx*)Rinv, ldi);
    } else {
        const size_t smem = sizeof(cplx) * (size_t)wpb * r;
        MF_CHECK_CUDA(cudaFuncSetAttribute(trtri_upper_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trtri_upper_kernel<false><<<(r + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>((const cplx*)R, ldr, r, (cplx*)Rinv, ldi);
    }
    MF_CHECK_LAUNCH();
    return 0;
}

static int scale_launch(mf_c128* X, int64_t ld, int rows, int cols, const double* d, int power, int by_rows, void* stream) {
    long long total = (long long)rows * cols;
    if (total == 0) return 0;
    scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((cplx*)X, ld, rows, cols, d, power, by_rows);
    g_mf_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_mf_err, sizeof(g_mf_err), "scale: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int mf_scale_cols_c128(mf_c128* X, int64_t ld, int rows, int cols, const double* d, int power, void* stream) {
    if (!X || ld < cols) MF_FAIL_ARG(1, "X is NULL or ld < cols");
    if (!d) MF_FAIL_ARG(5, "d is NULL");
    if (power != 1 && power != -1) MF_FAIL_ARG(6, "power must be +1 or -1");
    return scale_launch(X, ld, rows, cols, d, power, 0, stream);
}

extern "C" int mf_scale_rows_c128(mf_c128* X, int64_t ld, int rows, int cols, const double* d, int power, void* stream) {
    if (!X || ld < cols) MF_FAIL_ARG(1, "X is NULL or ld < cols");
    if (!d) MF_FAIL_ARG(5, "d is NULL");
    if (power != 1 && power != -1) MF_FAIL_ARG(6, "power must be +1 or -1");
    return scale_launch(X, ld, rows, cols, d, power, 1, stream);
}

extern "C" int mf_symmetrize_c128(const mf_c128* A, int64_t lda, int r, mf_c128* As, int64_t lds, void* stream) {
    if (!A || lda < r) MF_FAIL_ARG(1, "A is NULL or lda < r");
    if (r <= 0) MF_FAIL_ARG(3, "r <= 0");
    if (!As || lds < r) MF_FAIL_ARG(4, "As is NULL or lds < r");
    if ((const void*)A == (const void*)As) MF_FAIL_ARG(4, "As must not alias A");
    symmetrize_kernel<<<(r * r + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const cplx*)A, lda, r, (cplx*)As, lds);
    MF_CHECK_LAUNCH();
    return 0;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" size_t mf_jacobi_svd_ws_bytes(int r) {
    if (r <= 0) return 256;
    size_t r2 = (size_t)((r + 15) & ~15);            // the block kernel pads to a multiple of 2B = 16 rows
    return align_up(sizeof(cplx) * r2 * r, 256) + align_up(sizeof(cplx) * r2 * r2, 256) + align_up(sizeof(double) * r2, 256) + 64 * sizeof(double) + 256;
}

extern "C" int mf_jacobi_svd_c128(mf_c128* X, int64_t ld, int r, mf_c128* U, int64_t ldu, double* sigma, int max_sweeps,
                                  double tol, int* sweeps_done, void* ws, size_t ws_bytes, void* stream) {
    if (!X || ld < r) MF_FAIL_ARG(1, "X is NULL or ld < r");
    if (r <= 0 || r > 2048) MF_FAIL_ARG(3, "need 0 < r <= 2048");
    if (!U || ldu < r) MF_FAIL_ARG(4, "U is NULL or ldu < r");
    if (!sigma) MF_FAIL_ARG(6, "sigma is NULL");
    if (max_sweeps <= 0 || max_sweeps > 64) MF_FAIL_ARG(7, "need 0 < max_sweeps <= 64");
    if (!ws || ws_bytes < mf_jacobi_svd_ws_bytes(r)) MF_FAIL_ARG(10, "workspace too small (mf_jacobi_svd_ws_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t r2 = (size_t)((r + 1) & ~1);
    if (r <= JS_SMEM_MAX) {
        const size_t smem = sizeof(cplx) * (r2 * (size_t)(r | 1) + r2 * (r2 | 1));
        MF_CHECK_CUDA(cudaFuncSetAttribute(jacobi_svd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_svd_smem_kernel<<<1, (unsigned)(32 * (r2 / 2)), smem, st>>>((const cplx*)X, ld, r, (cplx*)U, ldu, sigma, max_sweeps, tol, sweeps_done);
        MF_CHECK_LAUNCH();
        return 0;
    }
    {   // block kernel: 8 rows per block while 16 rows of X and of the accumulator fit in shared memory, else 4
        const int B = ((size_t)16 * (2 * (size_t)r + 16) * sizeof(cplx) <= 220 * 1024) ? 8 : 4;
        const size_t r2b = ((size_t)r + 2 * B - 1) / (2 * B) * (2 * B);
        const size_t smem = sizeof(cplx) * 2 * B * ((size_t)r + r2b);
        const int grid = (int)(r2b / B / 2);
        const char* force_old = getenv("MF_JACOBI_PAIR_KERNEL");
        if (smem <= 220 * 1024 && grid <= mf_num_sms() && !(force_old && atoi(force_old))) {
            char* base = (char*)ws;
            cplx* Xw = (cplx*)base; base += align_up(sizeof(cplx) * r2b * r, 256);
            cplx* Gacc = (cplx*)base; base += align_up(sizeof(cplx) * r2b * r2b, 256);
            double* sig = (double*)base; base += align_up(sizeof(double) * r2b, 256);
            double* offmax = (double*)base; base += 64 * sizeof(double);
            unsigned* sync_words = (unsigned*)base;
            MF_CHECK_CUDA(cudaMemsetAsync(sync_words, 0, 256, st));
            const cplx* Xin = (const cplx*)X; long long ldl = ld, ldul = ldu; cplx* Uc = (cplx*)U; int r2i = (int)r2b;
            void* args[] = {(void*)&Xin, (void*)&ldl, (void*)&r, (void*)&r2i, (void*)&Uc, (void*)&ldul, (void*)&sigma, (void*)&max_sweeps,
                            (void*)&tol, (void*)&sweeps_done, (void*)&Xw, (void*)&Gacc, (void*)&sig, (void*)&offmax, (void*)&sync_words};
            const void* kern = B == 8 ? (const void*)jacobi_svd_block_kernel<8> : (const void*)jacobi_svd_block_kernel<4>;
            MF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MF_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(JB_THREADS), args, smem, st));
            g_mf_launches.fetch_add(1, std::memory_order_relaxed);
            return 0;
        }
    }
    char* base = (char*)ws;
    cplx* Xw = (cplx*)base; base += align_up(sizeof(cplx) * r2 * r, 256);
    cplx* Gacc = (cplx*)base; base += align_up(sizeof(cplx) * r2 * r2, 256);
    double* sig = (double*)base; base += align_up(sizeof(double) * r2, 256);
    double* offmax = (double*)base; base += 64 * sizeof(double);
    unsigned* counter = (unsigned*)base;
    const int grid = (int)(r2 / 2);
    int per_sm = 0;
    MF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_svd_kernel, JS_THREADS, 0));
    if (grid > per_sm * mf_num_sms()) MF_FAIL_ARG(3, "r too large for a co-resident Jacobi grid");
    MF_CHECK_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    const cplx* Xin = (const cplx*)X; long long ldl = ld, ldul = ldu; cplx* Uc = (cplx*)U;
    void* args[] = {(void*)&Xin, (void*)&ldl, (void*)&r, (void*)&Uc, (void*)&ldul, (void*)&sigma, (void*)&max_sweeps, (void*)&tol,
                    (void*)&sweeps_done, (void*)&Xw, (void*)&Gacc, (void*)&sig, (void*)&offmax, (void*)&counter};
    MF_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)jacobi_svd_kernel, dim3(grid), dim3(JS_THREADS), args, 0, st));
    g_mf_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

/* 1 when mf_jacobi_svd_f64 supports this r (single-CTA shared-memory kernel), else 0 (use the complex128 entry). */
extern "C" int mf_jacobi_svd_f64_supported(int r) { return (r >= 1 && r <= JS_SMEM_MAX) ? 1 : 0; }

extern "C" int mf_jacobi_svd_f64(const double* X, int64_t ld, int r, double* U, int64_t ldu, double* sigma, int max_sweeps,
                                 double tol, int* sweeps_done, void* stream) {
    if (!X || ld < r) MF_FAIL_ARG(1, "X is NULL or ld < r");
    if (r <= 0 || r > JS_SMEM_MAX) MF_FAIL_ARG(3, "need 0 < r <= 64 (mf_jacobi_svd_f64_supported)");
    if (!U || ldu < r) MF_FAIL_ARG(4, "U is NULL or ldu < r");
    if (!sigma) MF_FAIL_ARG(6, "sigma is NULL");
    if (max_sweeps <= 0 || max_sweeps > 64) MF_FAIL_ARG(7, "need 0 < max_sweeps <= 64");
    const size_t r2 = (size_t)((r + 1) & ~1);
    const size_t smem = sizeof(double) * (r2 * (size_t)(r | 1) + r2 * (r2 | 1));
    MF_CHECK_CUDA(cudaFuncSetAttribute(jacobi_svd_smem_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_svd_smem_f64_kernel<<<1, (unsigned)(32 * (r2 / 2)), smem, (cudaStream_t)stream>>>(X, ld, r, U, ldu, sigma, max_sweeps, tol, sweeps_done);
    MF_CHECK_LAUNCH();
    return 0;
}
