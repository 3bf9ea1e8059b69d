// Fused r x r Cholesky + triangular inverse for the basis stage (CholeskyQR passes, implementation.py:226/298/210):
//   G (Hermitian positive definite, equilibrated)  ->  R upper with G = R^H R (in place), Rinv = R^-1 (upper).
// ONE cooperative launch of CI_CTAS CTAs replaces the chain of ~80 small launches a blocked factorisation built from
// separate potrf / trtri / GEMM kernels needs at r = 256 (0.8 ms -> ~0.2 ms per Cholesky-QR pass).
//
// Right-looking blocked Cholesky, 32 x 32 blocks, three device-wide barriers (per-CTA flag words, ~1 us) per block step:
//   S1  CTA 0 factors the diagonal block in shared memory and inverts it (one warp per column);
//   S2  R[k][j] = Rinv_kk^H G[k][j] for the blocks right of the diagonal, one block per CTA;
//   S3  G[i][j] -= R[k][i]^H R[k][j] for the trailing block pairs i <= j, dealt round-robin to the CTAs.
// Then the off-diagonal blocks of the inverse, one CTA per block column (independent columns, bottom-up recurrence
//   Rinv[i][j] = -Rinv_ii sum_{l = i+1..j} R[i][l] Rinv[l][j]).
// A pivot that is not positive and finite is reported in *info (1-based column, LAPACK potrf convention; the first one
// wins) and replaced by 1 so that every CTA keeps in step with the barriers; the caller discards the factor.
// All data another CTA produced inside the launch is read with ld.global.cg (L2), never through L1.
#include "common.cuh"

namespace {

constexpr int CI_NB = 32, CI_T = 256, CI_LDS = CI_NB + 1, CI_CTAS = 32;

// Block (rows r0.., columns c0..) of M into S (32 x 33), zero padded outside the n x n matrix.
// ct: S[a][b] = conj(M[r0 + b][c0 + a])  (the conjugate transpose of the block).
__device__ __forceinline__ void ci_load(cplx* S, const cplx* M, long long ld, int r0, int c0, int n, bool ct) {
    for (int e = threadIdx.x; e < CI_NB * CI_NB; e += CI_T) {
        const int a = e >> 5, b = e & 31;
        const int gr = r0 + a, gc = c0 + b;
        cplx v = cmake(0.0, 0.0);
        if (gr < n && gc < n) v = __ldcg(reinterpret_cast<const double2*>(M + (long long)gr * ld + gc));
        if (ct) S[b * CI_LDS + a] = cconj(v); else S[a * CI_LDS + b] = v;
    }
}

// acc (rows ty, ty + 16; columns tx, tx + 16) += As (32 x 32) * Bs (32 x 32)
__device__ __forceinline__ void ci_mma(cplx (&acc)[2][2], const cplx* As, const cplx* Bs) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll 8
    for (int l = 0; l < CI_NB; ++l) {
        const cplx a0 = As[ty * CI_LDS + l], a1 = As[(ty + 16) * CI_LDS + l];
        const cplx b0 = Bs[l * CI_LDS + tx], b1 = Bs[l * CI_LDS + tx + 16];
        cfma(acc[0][0], a0, b0); cfma(acc[0][1], a0, b1); cfma(acc[1][0], a1, b0); cfma(acc[1][1], a1, b1);
    }
}

__global__ void __launch_bounds__(CI_T)
chol_inv_kernel(cplx* __restrict__ G, long long ldg, int n, cplx* __restrict__ Rinv, long long ldi, int* __restrict__ info, unsigned* flags) {
    __shared__ __align__(16) cplx As[CI_NB * CI_LDS];
    __shared__ __align__(16) cplx Bs[CI_NB * CI_LDS];
    __shared__ cplx rd[CI_NB];
    __shared__ int bad_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int nb = (n + CI_NB - 1) / CI_NB;
    unsigned epoch = 0;
    if (cta == 0 && tid == 0) { bad_sh = 0; }

    for (int k = 0; k < nb; ++k) {
        const int k0 = k * CI_NB, bk = min(CI_NB, n - k0);
        // ---- S1: diagonal block, CTA 0 ----
        if (cta == 0) {
            ci_load(As, G, ldg, k0, k0, n, false);
            __syncthreads();
            for (int c = 0; c < bk; ++c) {
                double piv = As[c * CI_LDS + c].x;
                if (!(piv > 0.0) || !isfinite(piv)) { if (tid == 0 && bad_sh == 0) bad_sh = k0 + c + 1; piv = 1.0; }
                const double inv = rsqrt(piv);
                __syncthreads();                                 // everyone has read the pivot
                if (tid >= c && tid < bk) {
                    cplx v = As[c * CI_LDS + tid];
                    if (tid == c) v = cmake(piv * inv, 0.0); else { v.x *= inv; v.y *= inv; }
                    As[c * CI_LDS + tid] = v;
                }
                __syncthreads();
                // trailing update of the upper triangle of the block: element (i, j), c < i <= j < bk
                for (int e = tid; e < CI_NB * CI_NB; e += CI_T) {
                    const int i = e >> 5, j = e & 31;
                    if (i > c && j >= i && j < bk) cfms(As[i * CI_LDS + j], cconj(As[c * CI_LDS + i]), As[c * CI_LDS + j]);
                }
                __syncthreads();
            }
            if (tid < CI_NB) rd[tid] = tid < bk ? crecip(As[tid * CI_LDS + tid]) : cmake(1.0, 0.0);
            __syncthreads();
            // inverse of the triangular block: warp w solves R x = e_j for columns j = w, w + 8, ...; x_i lives in lane i
            for (int j = warp; j < bk; j += CI_T / 32) {
                cplx x = cmake(lane == j ? 1.0 : 0.0, 0.0);
                for (int kk = j; kk >= 0; --kk) {
                    const cplx xk0 = cmul(x, rd[kk]);
                    const double xkr = __shfl_sync(0xffffffffu, xk0.x, kk), xki = __shfl_sync(0xffffffffu, xk0.y, kk);
                    if (lane == kk) x = cmake(xkr, xki);
                    else if (lane < kk) cfms(x, As[lane * CI_LDS + kk], cmake(xkr, xki));
                }
                Bs[lane * CI_LDS + j] = (lane <= j) ? x : cmake(0.0, 0.0);
            }
            __syncthreads();
            for (int e = tid; e < CI_NB * CI_NB; e += CI_T) {
                const int i = e >> 5, j = e & 31;
                if (i < bk && j < bk) {
                    G[(long long)(k0 + i) * ldg + k0 + j] = (j >= i) ? As[i * CI_LDS + j] : cmake(0.0, 0.0);
                    Rinv[(long long)(k0 + i) * ldi + k0 + j] = (j >= i) ? Bs[i * CI_LDS + j] : cmake(0.0, 0.0);
                }
            }
        }
        mf_grid_barrier_flags(flags, epoch, ncta);
        // ---- S2: R[k][j] = Rinv_kk^H G[k][j], j > k ----
        for (int j = k + 1 + cta; j < nb; j += ncta) {
            ci_load(As, Rinv, ldi, k0, k0, n, true);
            ci_load(Bs, G, ldg, k0, j * CI_NB, n, false);
            __syncthreads();
            cplx acc[2][2] = {{cmake(0.0, 0.0), cmake(0.0, 0.0)}, {cmake(0.0, 0.0), cmake(0.0, 0.0)}};
            ci_mma(acc, As, Bs);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int gr = k0 + ty + 16 * a, gc = j * CI_NB + tx + 16 * b;
                    if (gr < n && gc < n) G[(long long)gr * ldg + gc] = acc[a][b];
                }
            __syncthreads();
        }
        mf_grid_barrier_flags(flags, epoch, ncta);
        // ---- S3: G[i][j] -= R[k][i]^H R[k][j], k < i <= j ----
        {
            const int nt = nb - (k + 1);                         // trailing blocks per dimension
            const int ntask = nt * (nt + 1) / 2;
            for (int task = cta; task < ntask; task += ncta) {
                int ii = 0, rem = task;                          // row-major enumeration of the upper block triangle
                while (rem >= nt - ii) { rem -= nt - ii; ++ii; }
                const int i = k + 1 + ii, j = i + rem;
                ci_load(As, G, ldg, k0, i * CI_NB, n, true);
                ci_load(Bs, G, ldg, k0, j * CI_NB, n, false);
                __syncthreads();
                cplx acc[2][2] = {{cmake(0.0, 0.0), cmake(0.0, 0.0)}, {cmake(0.0, 0.0), cmake(0.0, 0.0)}};
                ci_mma(acc, As, Bs);
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int gr = i * CI_NB + ty + 16 * a, gc = j * CI_NB + tx + 16 * b;
                        if (gr < n && gc < n) {
                            cplx* dst = G + (long long)gr * ldg + gc;
                            const cplx v = __ldcg(reinterpret_cast<const double2*>(dst));
                            *dst = csub(v, acc[a][b]);
                        }
                    }
                __syncthreads();
            }
        }
        mf_grid_barrier_flags(flags, epoch, ncta);
    }
    if (cta == 0 && tid == 0 && info) *info = bad_sh;

    // ---- off-diagonal blocks of the inverse, one CTA per block column; strictly lower blocks are zeroed ----
    for (int j = 1 + cta; j < nb; j += ncta) {
        for (int i = j - 1; i >= 0; --i) {
            cplx acc[2][2] = {{cmake(0.0, 0.0), cmake(0.0, 0.0)}, {cmake(0.0, 0.0), cmake(0.0, 0.0)}};
            for (int l = i + 1; l <= j; ++l) {
                ci_load(As, G, ldg, i * CI_NB, l * CI_NB, n, false);
                ci_load(Bs, Rinv, ldi, l * CI_NB, j * CI_NB, n, false);
                __syncthreads();
                ci_mma(acc, As, Bs);
                __syncthreads();
            }
            // Rinv[i][j] = -Rinv_ii * acc
            ci_load(As, Rinv, ldi, i * CI_NB, i * CI_NB, n, false);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) Bs[(ty + 16 * a) * CI_LDS + tx + 16 * b] = acc[a][b];
            __syncthreads();
            cplx out[2][2] = {{cmake(0.0, 0.0), cmake(0.0, 0.0)}, {cmake(0.0, 0.0), cmake(0.0, 0.0)}};
            ci_mma(out, As, Bs);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int gr = i * CI_NB + ty + 16 * a, gc = j * CI_NB + tx + 16 * b;
                    if (gr < n && gc < n) Rinv[(long long)gr * ldi + gc] = cmake(-out[a][b].x, -out[a][b].y);
                }
            __syncthreads();                                     // the block just written is read back (through L2) next
        }
    }
    // zeros below the block diagonal of both outputs (the diagonal blocks were cleaned in S1)
    for (long long e = (long long)cta * CI_T + tid; e < (long long)n * n; e += (long long)ncta * CI_T) {
        const int i = (int)(e / n), j = (int)(e - (long long)i * n);
        if ((i / CI_NB) > (j / CI_NB)) { G[(long long)i * ldg + j] = cmake(0.0, 0.0); Rinv[(long long)i * ldi + j] = cmake(0.0, 0.0); }
    }
}

}  // namespace

extern "C" size_t mf_chol_inv_ws_bytes(int r) { (void)r; return 256; }

extern "C" int mf_chol_inv_upper_c128(mf_c128* G, int64_t ldg, int r, mf_c128* Rinv, int64_t ldi, int* info,
                                      void* ws, size_t ws_bytes, void* stream) {
    if (!G || ldg < r) MF_FAIL_ARG(1, "G is NULL or ldg < r");
    if (r <= 0 || r > 1024) MF_FAIL_ARG(3, "need 0 < r <= 1024");
    if (!Rinv || ldi < r) MF_FAIL_ARG(4, "Rinv is NULL or ldi < r");
    if ((const void*)G == (const void*)Rinv) MF_FAIL_ARG(4, "Rinv must not alias G");
    if (!ws || ws_bytes < mf_chol_inv_ws_bytes(r)) MF_FAIL_ARG(7, "workspace too small (mf_chol_inv_ws_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (r + CI_NB - 1) / CI_NB;
    int grid = nb * (nb - 1) / 2;                    // most parallel phase: the first trailing update
    if (grid < 1) grid = 1;
    if (grid > CI_CTAS) grid = CI_CTAS;
    MF_CHECK_CUDA(cudaMemsetAsync(ws, 0, 256, st));  // barrier flags (one word per CTA)
    cplx* Gc = (cplx*)G; cplx* Ri = (cplx*)Rinv; long long ldgl = ldg, ldil = ldi; unsigned* flags = (unsigned*)ws;
    void* args[] = {(void*)&Gc, (void*)&ldgl, (void*)&r, (void*)&Ri, (void*)&ldil, (void*)&info, (void*)&flags};
    MF_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)chol_inv_kernel, dim3(grid), dim3(CI_T), args, 0, st));
    g_mf_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}
