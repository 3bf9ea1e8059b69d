// Device helpers shared by the two blocked-LU sweep kernels (sweep_blocked.cu: matrix resident in shared memory;
// sweep_stream.cu: matrix streamed from an L2-resident workspace slot).  See sweep_blocked.cu for the layout notes.
#pragma once
#include "sweep_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int swz(int g) { return (((g ^ (g >> 2)) & 1) << 2) | (g & 3); }
__device__ __forceinline__ int mphys(int row, int col, int LD) { return row * LD + (col & ~7) + ((col & 7) ^ swz(row & 7)); }

// 1/a by Smith's formula with reciprocals (two correctly rounded reciprocals instead of three divisions)
__device__ __noinline__ cplx crecip2(cplx a) {
    if (fabs(a.x) >= fabs(a.y)) {
        const double ia = 1.0 / a.x, t = a.y * ia, d = fma(a.y, t, a.x), id = 1.0 / d;
        return cmake(id, -t * id);
    } else {
        const double ib = 1.0 / a.y, t = a.x * ib, d = fma(a.x, t, a.y), id = 1.0 / d;
        return cmake(t * id, -id);
    }
}

// ---- B. row exchanges + U12 = L11^-1 A12 for one trailing column c (L11 is stored negated) -----------------------
__device__ __forceinline__ void stepb_column(cplx* M, const int LD, const int row0, const int c, const int* pv) {
    const int cbase = c & ~7, cin = c & 7;
    cplx* colp = M + row0 * LD + cbase;
    // the (few) row exchanges of this panel, in order, straight in shared memory: compact code, uniform branches
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int P = pv[j];
        if (P != row0 + j) {
            cplx* x = colp + j * LD + (cin ^ swz(j));
            cplx* y = M + P * LD + cbase + (cin ^ swz(P & 7));
            const cplx tmp = *x; *x = *y; *y = tmp;
        }
    }
    cplx u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = colp[j * LD + (cin ^ swz(j))];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        const cplx* lrow = M + (row0 + j) * LD + row0;
        const int sw = swz(j);
#pragma unroll
        for (int i = 0; i < j; ++i) cfma(u[j], lrow[i ^ sw], u[i]);
    }
#pragma unroll
    for (int j = 1; j < 8; ++j) colp[j * LD + (cin ^ swz(j))] = u[j];
}

// per-lane fragment offsets inside 8 x 8 blocks of the swizzled matrix
struct FragOff { int a0, a1, c0, c1, b0, b1, g; };

// ---- C. tiles [t_lo, t_hi) of the trailing update of panel k over row blocks rb0.. (nrb of them) and column blocks
//         cb0.., column-block-major so that the B fragments are reloaded only when the column block changes -------
__device__ __forceinline__ void update_tiles(cplx* M, const int LD, const int row0, const int rb0, const int nrb, const int cb0,
                                             const int t_lo, const int t_hi, const FragOff& fo) {
    if (t_hi <= t_lo) return;
    int cbk = t_lo / nrb, rbk = t_lo - cbk * nrb;
    const cplx* Ub = M + row0 * LD + 8 * cb0;
    cplx b0 = Ub[8 * cbk + fo.b0], b1 = Ub[8 * cbk + fo.b1];
    for (int ti = t_lo; ti < t_hi; ++ti) {
        cplx* rowp = M + (8 * (rb0 + rbk) + fo.g) * LD;
        const cplx a0 = rowp[row0 + fo.a0], a1 = rowp[row0 + fo.a1];
        cplx* pc0 = rowp + 8 * (cb0 + cbk) + fo.c0;
        cplx* pc1 = rowp + 8 * (cb0 + cbk) + fo.c1;
        const cplx v0 = *pc0, v1 = *pc1;
        double cre0 = v0.x, cre1 = v1.x, cim0 = v0.y, cim1 = v1.y;
        dmma884(cre0, cre1, a0.x, b0.x); dmma884(cim0, cim1, a0.x, b0.y);
        dmma884(cre0, cre1, -a0.y, b0.y); dmma884(cim0, cim1, a0.y, b0.x);
        dmma884(cre0, cre1, a1.x, b1.x); dmma884(cim0, cim1, a1.x, b1.y);
        dmma884(cre0, cre1, -a1.y, b1.y); dmma884(cim0, cim1, a1.y, b1.x);
        *pc0 = cmake(cre0, cim0); *pc1 = cmake(cre1, cim1);
        if (++rbk == nrb) {
            rbk = 0; ++cbk;
            if (ti + 1 < t_hi) { b0 = Ub[8 * cbk + fo.b0]; b1 = Ub[8 * cbk + fo.b1]; }
        }
    }
}

// One thread per point: S <- 2 (I + Z^-1)^-1 - I in place (Z left in S by the sweep kernel).
template <int MMAX>
__global__ void __launch_bounds__(128) gsm_finish_kernel(cplx* __restrict__ S, int m, long long F) {
    const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= F) return;
    cplx z[MMAX * MMAX], scratch[2 * MMAX * MMAX];
    cplx* sp = S + pt * (long long)m * m;
    for (int e = 0; e < m * m; ++e) z[e] = sp[e];
    gsm_from_impedance(z, scratch, m, sp);
}


inline int gsm_finish_launch(cplx* S, int m, long long F, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((F + 127) / 128);
    if (m <= 2) gsm_finish_kernel<2><<<blocks, 128, 0, stream>>>(S, m, F);
    else if (m <= 4) gsm_finish_kernel<4><<<blocks, 128, 0, stream>>>(S, m, F);
    else if (m <= 8) gsm_finish_kernel<8><<<blocks, 128, 0, stream>>>(S, m, F);
    else gsm_finish_kernel<MF_MAX_PORTS><<<blocks, 128, 0, stream>>>(S, m, F);
    MF_CHECK_LAUNCH();
    return 0;
}

}  // namespace
