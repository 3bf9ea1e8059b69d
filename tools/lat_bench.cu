// Dependent-issue latencies of the instructions on the pivot-search path of the sweep kernels (one warp, one SM; B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_bench tools/lat_bench.cu && tools/lat_bench
// Each kernel runs N dependent operations between two clock64 reads and prints cycles per operation.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 512;

template <int OP>
__global__ void lat(double* out, long long* cyc, double seed, int nwarps_bar) {
    __shared__ int chase[64];
    __shared__ double keys[16];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 64) chase[threadIdx.x] = (threadIdx.x + 1) & 63;
    if (threadIdx.x < 16) keys[threadIdx.x] = seed * threadIdx.x;
    __syncthreads();
    double x = seed + lane * 1e-3, a = 1.0000001, b = 1e-9;
    int k = lane;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, a, b);                                                   // DFMA
        if (OP == 1) x = x + b;                                                          // DADD
        if (OP == 2) x = x * a;                                                          // DMUL
        if (OP == 3) x = 1.0 / x;                                                        // IEEE reciprocal
        if (OP == 4) k = __reduce_max_sync(0xffffffffu, k ^ i);                          // REDUX
        if (OP == 5) k = __popc(__ballot_sync(0xffffffffu, (k + i) & 1)) + lane;         // VOTE + POPC
        if (OP == 6) k = __shfl_sync(0xffffffffu, k, (lane + 1) & 31) + 1;               // SHFL
        if (OP == 7) k = chase[k & 63];                                                  // LDS (pointer chase)
        if (OP == 8) { __syncthreads(); k += i; }                                        // BAR.SYNC with blockDim threads
        if (OP == 9) { asm volatile("bar.sync 1, %0;" :: "r"(nwarps_bar * 32)); k += i; }   // named barrier (all warps of the block)
        if (OP == 10) {                                                                  // hi-word arg-max as in the panel: REDUX + VOTE + POPC
            const int hi = __double2hiint(x);
            const int hm = __reduce_max_sync(0xffffffffu, hi);
            const bool own = hi == hm;
            x += (__popc(__ballot_sync(0xffffffffu, own)) == 1) ? b : a;
        }
        if (OP == 11) {                                                                  // shared-memory round trip: STS, barrier, LDS
            keys[threadIdx.x >> 5] = x;
            __syncthreads();
            x = keys[(i + 1) & ((blockDim.x >> 5) - 1)] + b;
        }
        if (OP == 12) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + b; }   // MUFU.RCP64H
        if (OP == 13) x = fma(x, a, b) * fma(x, b, a);                                   // two independent DFMA + DMUL
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + k;
}

// DFMA chain of warp 0 while the warps selected by `mask` (bit = hardware warp id) run back-to-back DMMAs
__global__ void contend(double* out, long long* cyc, double seed, unsigned mask) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double x = seed + lane * 1e-3, a = 1.0000001, b = 1e-9;
    __shared__ volatile int done;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    if (warp == 0) {
        const long long t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = fma(x, a, b);
        const long long t1 = clock64();
        if (lane == 0) { cyc[0] = t1 - t0; done = 1; }
    } else if (mask >> warp & 1) {
        double c[4][2] = {};
        while (!done) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
        }
        x = c[0][0] + c[1][1] + c[2][0] + c[3][1];
    }
    out[threadIdx.x] = x;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const char* names[] = {"DFMA", "DADD", "DMUL", "1.0/x (IEEE)", "REDUX.MAX", "VOTE+POPC", "SHFL", "LDS chase", "__syncthreads", "bar.sync 1",
                           "hi-word arg-max (REDUX+VOTE+POPC+DADD)", "STS + barrier + LDS + DADD", "rcp.approx.f64 + DADD", "2 DFMA + DMUL"};
    for (int threads : {32, 128, 256}) {
        printf("== %d threads (one CTA)\n", threads);
        for (int op = 0; op < 14; ++op) {
            if (threads > 32 && !(op == 8 || op == 9 || op == 11 || op == 0)) continue;
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
#define CASE(n) case n: lat<n><<<1, threads>>>(out, cyc, 1.25, threads / 32); break;
                    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
                }
            }
            long long h = 0;
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-42s %7.1f cycles/op\n", names[op], (double)h / N);
        }
    }
    printf("== DFMA dependent chain of warp 0 with DMMA streams on other warps of the CTA (256 threads)\n");
    for (unsigned mask : {0x00u, 0x10u, 0x02u, 0x04u, 0x08u, 0x30u, 0xf0u, 0xfeu, 0xccu, 0x66u}) {
        for (int rep = 0; rep < 2; ++rep) contend<<<1, 256>>>(out, cyc, 1.25, mask);
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA warps mask 0x%02x: DFMA %7.1f cycles/op\n", mask, (double)h / N);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
