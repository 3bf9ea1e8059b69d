// Pieces shared by the batched-sweep kernel variants (stages 3 + 4).
#pragma once
#include "common.cuh"

struct SweepParams {
    const cplx* A0; const cplx* A1; const cplx* A2; long long lda;   // symmetrised reduced operators (NULL = zero)
    const cplx* Br; long long ldb;                                   // reduced port matrix r x m
    int r, m;
    const double* c0; const double* c1; const double* c2; const double* cb; const double* zs;
    long long F;
    cplx* X;      // F x r x m or NULL
    cplx* S;      // F x m x m or NULL
    int* info;    // F or NULL
    cplx* ws; long long ws_stride;  // per-CTA workspace slots (elements)
};

// In-place inverse of the m x m matrix Z (row-major, ld = m) held in shared/local memory, by Gaussian
// elimination with partial pivoting on [Z | I] -- the algorithm behind np.linalg.inv (LAPACK gesv with an
// identity right-hand side; test_helpers.py:11, :13).  `aug` is scratch of m*2m elements.  Single thread.
__device__ inline void small_inverse(cplx* Z, cplx* aug, int m) {
    const int w = 2 * m;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            aug[i * w + j] = Z[i * m + j];
            aug[i * w + m + j] = cmake(i == j ? 1.0 : 0.0, 0.0);
        }
    for (int k = 0; k < m; ++k) {
        int p = k; double best = cabs1(aug[k * w + k]);
        for (int i = k + 1; i < m; ++i) { double v = cabs1(aug[i * w + k]); if (v > best) { best = v; p = i; } }
        if (p != k) for (int j = 0; j < w; ++j) { cplx t = aug[k * w + j]; aug[k * w + j] = aug[p * w + j]; aug[p * w + j] = t; }
        cplx rp = crecip(aug[k * w + k]);
        for (int i = k + 1; i < m; ++i) {
            cplx l = cmul(aug[i * w + k], rp);
            aug[i * w + k] = l;
            for (int j = k + 1; j < w; ++j) cfms(aug[i * w + j], l, aug[k * w + j]);
        }
    }
    for (int j = m; j < w; ++j)
        for (int k = m - 1; k >= 0; --k) {
            cplx v = aug[k * w + j];
            for (int i = k + 1; i < m; ++i) cfms(v, aug[k * w + i], aug[i * w + j]);
            aug[k * w + j] = cmul(v, crecip(aug[k * w + k]));
        }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Z[i * m + j] = aug[i * w + m + j];
}

// Z (m x m, shared, already reduced over the basis index) -> S = 2 (I + Z^-1)^-1 - I, written to Sout.
// test_helpers.py:9-14.  Single thread; scratch needs 2*m*m elements.
__device__ inline void gsm_from_impedance(cplx* Z, cplx* scratch, int m, cplx* Sout) {
    small_inverse(Z, scratch, m);                       // gam = inv(gim)
    for (int i = 0; i < m; ++i) Z[i * m + i].x += 1.0;  // id + gam
    small_inverse(Z, scratch, m);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            cplx v = cscale(2.0, Z[i * m + j]);
            if (i == j) v.x -= 1.0;
            Sout[i * m + j] = v;
        }
}
