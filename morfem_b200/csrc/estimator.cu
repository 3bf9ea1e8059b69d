// Residual error estimator of the greedy basis search (implementation.py:348-452), evaluated on the device from
// the reduced-sweep solutions.  One CTA per frequency point; see include/morfem_b200.h for the formula.
#include "sweep_common.cuh"

namespace {

struct EstParams {
    const cplx* G[9]; const cplx* H[3]; const cplx* BB; const cplx* X;
    const double* c[3]; const double* cb; double* err; int r, m; long long F;
};

constexpr int EST_THREADS = 128;

__global__ void __launch_bounds__(EST_THREADS) estimator_kernel(EstParams p) {
    extern __shared__ __align__(16) cplx sm[];
    const int r = p.r, m = p.m, tid = threadIdx.x;
    cplx* xs = sm;                 // r*m
    cplx* E = xs + r * m;          // m*m
    const long long pt = blockIdx.x;
    const cplx* x = p.X + pt * (long long)r * m;
    for (int i = tid; i < r * m; i += EST_THREADS) xs[i] = x[i];
    __syncthreads();
    const double cc[3] = {p.c[0][pt], p.c[1][pt], p.c[2][pt]};
    const double cb = p.cb[pt];
    // T[i][q] = sum_ab c_a c_b G_ab[i, :] x  -  cb sum_a c_a H_a[i, :]   (r x m), one row per thread, kept in shared memory
    cplx* T = E + m * m;           // r*m
    for (int i = tid; i < r; i += EST_THREADS) {
        cplx t[MF_MAX_PORTS];
        for (int q = 0; q < m; ++q) t[q] = cmake(0.0, 0.0);
        for (int ab = 0; ab < 9; ++ab) {
            const cplx* G = p.G[ab];
            if (!G) continue;
            const double w = cc[ab / 3] * cc[ab % 3];
            for (int k = 0; k < r; ++k) {
                cplx g = cscale(w, G[(long long)i * r + k]);
                for (int q = 0; q < m; ++q) cfma(t[q], g, xs[k * m + q]);
            }
        }
        for (int a = 0; a < 3; ++a) {
            const cplx* H = p.H[a];
            if (!H) continue;
            const double w = -cb * cc[a];
            for (int q = 0; q < m; ++q) { cplx h = H[(long long)i * m + q]; t[q].x = fma(w, h.x, t[q].x); t[q].y = fma(w, h.y, t[q].y); }
        }
        for (int q = 0; q < m; ++q) T[i * m + q] = t[q];
    }
    __syncthreads();
    // E[pp][q] = sum_i conj(x[i][pp]) T[i][q]  - cb sum_a c_a (H_a^H x)[pp][q] + cb^2 BB[pp][q]: one entry per thread, summed in a fixed
    // order (deterministic estimator values, any m up to MF_MAX_PORTS)
    for (int e = tid; e < m * m; e += EST_THREADS) {
        const int pp = e / m, q = e - pp * m;
        cplx acc = cmake(0.0, 0.0);
        for (int i = 0; i < r; ++i) cfma(acc, cconj(xs[i * m + pp]), T[i * m + q]);
        for (int a = 0; a < 3; ++a) {
            const cplx* H = p.H[a];
            if (!H) continue;
            const double w = -cb * cc[a];
            cplx s = cmake(0.0, 0.0);
            for (int k = 0; k < r; ++k) cfma(s, cconj(H[(long long)k * m + pp]), xs[k * m + q]);
            acc.x = fma(w, s.x, acc.x); acc.y = fma(w, s.y, acc.y);
        }
        if (p.BB) { cplx bb = p.BB[e]; acc.x = fma(cb * cb, bb.x, acc.x); acc.y = fma(cb * cb, bb.y, acc.y); }
        E[e] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < m * m; ++i) s += cnorm2(E[i]);
        p.err[pt] = sqrt(s);
    }
}

}  // namespace

extern "C" int mf_estimator_c128(const mf_c128* X, int r, int m, int64_t F, const mf_c128* const* G_host,
                                 const mf_c128* const* H_host, const mf_c128* BB,
                                 const double* c0, const double* c1, const double* c2, const double* cb,
                                 double* err, void* stream) {
    if (!X) MF_FAIL_ARG(1, "X is NULL");
    if (r <= 0) MF_FAIL_ARG(2, "r <= 0");
    if (m <= 0 || m > MF_MAX_PORTS) MF_FAIL_ARG(3, "need 0 < m <= MF_MAX_PORTS");
    if (F < 0) MF_FAIL_ARG(4, "F < 0");
    if (!G_host) MF_FAIL_ARG(5, "G_host is NULL");
    if (!H_host) MF_FAIL_ARG(6, "H_host is NULL");
    if (!c0 || !c1 || !c2 || !cb) MF_FAIL_ARG(8, "coefficient arrays must not be NULL");
    if (!err) MF_FAIL_ARG(12, "err is NULL");
    if (F == 0) return 0;
    EstParams p;
    for (int i = 0; i < 9; ++i) p.G[i] = (const cplx*)G_host[i];
    for (int i = 0; i < 3; ++i) p.H[i] = (const cplx*)H_host[i];
    p.BB = (const cplx*)BB; p.X = (const cplx*)X; p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.cb = cb; p.err = err;
    p.r = r; p.m = m; p.F = F;
    const size_t smem = sizeof(cplx) * (2 * (size_t)r * m + (size_t)m * m);
    MF_CHECK_CUDA(cudaFuncSetAttribute(estimator_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    estimator_kernel<<<(unsigned)F, EST_THREADS, smem, (cudaStream_t)stream>>>(p);
    MF_CHECK_LAUNCH();
    return 0;
}

namespace {
__global__ void __launch_bounds__(64) gsm_kernel(const cplx* __restrict__ X, const cplx* __restrict__ B, long long ldb, int r, int m,
                                                 const double* __restrict__ cb, const double* __restrict__ zs, cplx* __restrict__ S) {
    __shared__ cplx zmat[MF_MAX_PORTS * MF_MAX_PORTS];
    __shared__ cplx zscr[2 * MF_MAX_PORTS * MF_MAX_PORTS];
    const long long pt = blockIdx.x;
    const cplx* x = X + pt * (long long)r * m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const double cbv = cb[pt], z = zs[pt];
    for (int e = warp; e < m * m; e += nwarps) {
        int a = e / m, b = e - a * m;
        cplx acc = cmake(0.0, 0.0);
        for (int k = lane; k < r; k += 32) cfma(acc, x[k * m + a], cscale(cbv, B[k * ldb + b]));
        for (int off = 16; off > 0; off >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (lane == 0) zmat[e] = cmake(-z * acc.y, z * acc.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) gsm_from_impedance(zmat, zscr, m, S + pt * (long long)m * m);
}
}  // namespace

extern "C" int mf_gsm_c128(const mf_c128* X, const mf_c128* Bmat, int64_t ldb, int r, int m, const double* cb,
                           const double* zscale, int64_t F, mf_c128* S, void* stream) {
    if (!X) MF_FAIL_ARG(1, "X is NULL");
    if (!Bmat || ldb < m) MF_FAIL_ARG(2, "Bmat is NULL or ldb < m");
    if (r <= 0) MF_FAIL_ARG(4, "r <= 0");
    if (m <= 0 || m > MF_MAX_PORTS) MF_FAIL_ARG(5, "need 0 < m <= MF_MAX_PORTS");
    if (!cb) MF_FAIL_ARG(6, "cb is NULL");
    if (!zscale) MF_FAIL_ARG(7, "zscale is NULL");
    if (F < 0) MF_FAIL_ARG(8, "F < 0");
    if (!S) MF_FAIL_ARG(9, "S is NULL");
    if (F == 0) return 0;
    gsm_kernel<<<(unsigned)F, 64, 0, (cudaStream_t)stream>>>((const cplx*)X, (const cplx*)Bmat, ldb, r, m, cb, zscale, (cplx*)S);
    MF_CHECK_LAUNCH();
    return 0;
}
