// Variant 1 of the batched reduced sweep: one CTA per frequency point, any r.
// The augmented matrix [A(t) | cb(t) Br] lives in shared memory when it fits (r <~ 110) and in a per-CTA
// global workspace slot (L2-resident) otherwise.  Unblocked right-looking LU with partial pivoting followed
// by back-substitution; the S-parameter epilogue is fused.  This is the simple, always-available path that
// the faster variants are checked against.
//
// Reference semantics: implementation.py:468-480 (solve_fem_point, dense branch), :526-533 (system_matrix,
// impulse_vector), test_helpers.py:9-14 (generalized_scattering_matrix).
#include "sweep_common.cuh"

namespace {

constexpr int GEN_THREADS = 256;

struct ArgMax { double v; int i; };

__device__ __forceinline__ ArgMax argmax_merge(ArgMax a, ArgMax b) {
    // larger value wins; ties go to the smaller index (first maximum, as izamax)
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(GEN_THREADS) sweep_generic_kernel(SweepParams p, int ldw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int r = p.r, m = p.m, tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;

    // shared carve-up: small arrays first, then (optionally) the matrix
    cplx* zmat = reinterpret_cast<cplx*>(smem_raw);            // m*m
    cplx* zscr = zmat + m * m;                                 // 2*m*m
    cplx* xk = zscr + 2 * m * m;                               // m   (broadcast of the solved row)
    double* redv = reinterpret_cast<double*>(xk + m);          // nwarps
    int* redi = reinterpret_cast<int*>(redv + 32);             // nwarps
    int* piv_sh = redi + 32;                                   // 1
    cplx* W = IN_SMEM ? reinterpret_cast<cplx*>(piv_sh + 4) : p.ws + (long long)blockIdx.x * p.ws_stride;
    const int ncol = r + m;

    for (long long pt = blockIdx.x; pt < p.F; pt += gridDim.x) {
        const double c0 = p.c0[pt], c1 = p.c1[pt], c2 = p.c2[pt], cb = p.cb[pt];
        // ---- assemble [A | b] ----
        for (int idx = tid; idx < r * r; idx += nthreads) {
            int i = idx / r, j = idx - i * r;
            cplx a = cmake(0.0, 0.0);
            if (p.A0) { cplx v = p.A0[i * p.lda + j]; a.x = c0 * v.x; a.y = c0 * v.y; }
            if (p.A1) { cplx v = p.A1[i * p.lda + j]; a.x = fma(c1, v.x, a.x); a.y = fma(c1, v.y, a.y); }
            if (p.A2) { cplx v = p.A2[i * p.lda + j]; a.x = fma(c2, v.x, a.x); a.y = fma(c2, v.y, a.y); }
            W[i * ldw + j] = a;
        }
        for (int idx = tid; idx < r * m; idx += nthreads) {
            int i = idx / m, j = idx - i * m;
            W[i * ldw + r + j] = cscale(cb, p.Br[i * p.ldb + j]);
        }
        int first_zero = 0;
        __syncthreads();

        // ---- LU with partial pivoting, right-hand sides eliminated alongside ----
        for (int k = 0; k < r; ++k) {
            ArgMax best; best.v = -1.0; best.i = k;
            for (int i = k + tid; i < r; i += nthreads) {
                ArgMax c; c.v = cabs1(W[i * ldw + k]); c.i = i;
                best = argmax_merge(best, c);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ArgMax o; o.v = __shfl_xor_sync(0xffffffffu, best.v, off); o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
                best = argmax_merge(best, o);
            }
            if (lane == 0) { redv[warp] = best.v; redi[warp] = best.i; }
            __syncthreads();
            if (warp == 0) {
                ArgMax b2; b2.v = lane < nwarps ? redv[lane] : -1.0; b2.i = lane < nwarps ? redi[lane] : 0x7fffffff;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    ArgMax o; o.v = __shfl_xor_sync(0xffffffffu, b2.v, off); o.i = __shfl_xor_sync(0xffffffffu, b2.i, off);
                    b2 = argmax_merge(b2, o);
                }
                if (lane == 0) { piv_sh[0] = b2.i; piv_sh[1] = (b2.v == 0.0); }
            }
            __syncthreads();
            const int pv = piv_sh[0];
            if (piv_sh[1] && first_zero == 0) first_zero = k + 1;
            if (pv != k) {
                for (int j = k + tid; j < ncol; j += nthreads) {
                    cplx t = W[k * ldw + j]; W[k * ldw + j] = W[pv * ldw + j]; W[pv * ldw + j] = t;
                }
            }
            __syncthreads();
            const cplx rp = crecip(W[k * ldw + k]);
            // multipliers
            for (int i = k + 1 + tid; i < r; i += nthreads) W[i * ldw + k] = cmul(W[i * ldw + k], rp);
            __syncthreads();
            // trailing update, columns fastest across threads
            const int tw = ncol - (k + 1);
            const int total = (r - (k + 1)) * tw;
            for (int idx = tid; idx < total; idx += nthreads) {
                int ii = idx / tw, jj = idx - ii * tw;
                int i = k + 1 + ii, j = k + 1 + jj;
                cplx a = W[i * ldw + j];
                cfms(a, W[i * ldw + k], W[k * ldw + j]);
                W[i * ldw + j] = a;
            }
            __syncthreads();
        }

        // ---- back substitution U x = y on the m right-hand sides ----
        for (int k = r - 1; k >= 0; --k) {
            if (tid < m) {
                cplx v = cmul(W[k * ldw + r + tid], crecip(W[k * ldw + k]));
                W[k * ldw + r + tid] = v;
                xk[tid] = v;
            }
            __syncthreads();
            for (int idx = tid; idx < k * m; idx += nthreads) {
                int i = idx / m, j = idx - i * m;
                cplx a = W[i * ldw + r + j];
                cfms(a, W[i * ldw + k], xk[j]);
                W[i * ldw + r + j] = a;
            }
            __syncthreads();
        }

        if (p.X) {
            cplx* xo = p.X + pt * (long long)r * m;
            for (int idx = tid; idx < r * m; idx += nthreads) { int i = idx / m, j = idx - i * m; xo[idx] = W[i * ldw + r + j]; }
        }
        if (p.info && tid == 0) p.info[pt] = first_zero;

        // ---- S-parameters: Z = j zs x^T (cb Br) ----
        if (p.S) {
            for (int e = warp; e < m * m; e += nwarps) {
                int a = e / m, b = e - a * m;
                cplx acc = cmake(0.0, 0.0);
                for (int k = lane; k < r; k += 32) cfma(acc, W[k * ldw + r + a], cscale(cb, p.Br[k * p.ldb + b]));
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                }
                if (lane == 0) { double zs = p.zs[pt]; zmat[e] = cmake(-zs * acc.y, zs * acc.x); }   // j*zs*acc
            }
            __syncthreads();
            if (tid == 0) gsm_from_impedance(zmat, zscr, m, p.S + pt * (long long)m * m);
        }
        __syncthreads();
    }
}

}  // namespace

size_t sweep_generic_small_smem(int m) { return sizeof(cplx) * (3 * m * m + m) + 32 * 8 + 32 * 4 + 16; }

int sweep_generic_launch(const SweepParams& p_in, size_t ws_bytes, cudaStream_t stream) {
    SweepParams p = p_in;
    const int r = p.r, m = p.m;
    int ldw = r + m; if ((ldw & 1) == 0) ldw += 1;
    const size_t small = sweep_generic_small_smem(m);
    const size_t mat_bytes = sizeof(cplx) * (size_t)r * ldw;
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const bool in_smem = small + mat_bytes + 1024 <= (size_t)max_smem;
    const int sms = mf_num_sms();
    if (in_smem) {
        size_t smem = small + mat_bytes;
        int per_sm = (int)((size_t)(max_smem) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 8) per_sm = 8;
        long long grid = (long long)sms * per_sm; if (grid > p.F) grid = p.F;
        MF_CHECK_CUDA(cudaFuncSetAttribute(sweep_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sweep_generic_kernel<true><<<(unsigned)grid, GEN_THREADS, smem, stream>>>(p, ldw);
    } else {
        const long long stride = (long long)r * ldw;
        long long slots = (long long)(ws_bytes / (sizeof(cplx) * stride));
        long long grid = (long long)sms * 2; if (grid > p.F) grid = p.F; if (grid > slots) grid = slots;
        if (grid < 1) MF_FAIL_ARG(21, "workspace too small for the generic sweep (see mf_sweep_ws_bytes)");
        if (!p.ws) MF_FAIL_ARG(21, "workspace pointer is NULL");
        p.ws_stride = stride;
        sweep_generic_kernel<false><<<(unsigned)grid, GEN_THREADS, small, stream>>>(p, ldw);
    }
    MF_CHECK_LAUNCH();
    return 0;
}

size_t sweep_generic_ws_bytes(int r, int m, long long F) {
    int ldw = r + m; if ((ldw & 1) == 0) ldw += 1;
    long long grid = (long long)mf_num_sms() * 2; if (grid > F) grid = F; if (grid < 1) grid = 1;
    return sizeof(cplx) * (size_t)r * ldw * grid;
}
