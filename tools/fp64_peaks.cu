// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) issue rates.
// These set the compute-roofline denominators for the dense contractions and the batched LU sweep
// ("FP64 peak must be measured on the box").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
    double av[4] = {a, a + 1, a + 2, a + 3};
    double bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
    double av[8] = {a, a + 1, a + 2, a + 3, a + 4, a + 5, a + 6, a + 7};
    double bv[4] = {b, b + 1, b + 2, b + 3};
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// smem-fed complex DMMA tile loop: each warp computes a 32x32 complex tile from smem operands with
// interleaved complex storage (one LDS.128 yields re+im fragment elements), 4 real DMMAs per complex block.
__global__ void zgemm_smem_kernel(double* out, int iters) {
    extern __shared__ double2 sm[];
    const int LD = 34;                       // complex elements per k-row (32 + 2 pad)
    double2* As = sm;                        // [64 k][LD]
    double2* Bs = sm + 64 * LD;
    for (int i = threadIdx.x; i < 2 * 64 * LD; i += blockDim.x) sm[i] = make_double2(1e-3 * (i % 7), 1e-3 * (i % 5));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    double cre[4][4][2], cim[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k0 = 0; k0 < 64; k0 += 4) {
            double2 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(k0 + t) * LD + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(k0 + t) * LD + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(cre[i][j][0], cre[i][j][1], a[i].x, b[j].x);
                    dmma884(cre[i][j][0], cre[i][j][1], -a[i].y, b[j].y);
                    dmma884(cim[i][j][0], cim[i][j][1], a[i].x, b[j].y);
                    dmma884(cim[i][j][0], cim[i][j][1], a[i].y, b[j].x);
                }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += cre[i][j][0] + cre[i][j][1] + cim[i][j][0] + cim[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32;
        int blocks = sms * (warps <= 8 ? 2 : 1);
        double nthreads = (double)blocks * threads;
        {
            float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = nthreads * iters * 8 * 2.0;
            printf("{\"bench\": \"dfma\", \"warps_per_cta\": %d, \"ctas\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", warps, blocks, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = nthreads / 32 * iters * 8 * (8 * 8 * 4 * 2.0);
            printf("{\"bench\": \"dmma_m8n8k4\", \"warps_per_cta\": %d, \"ctas\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", warps, blocks, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma1688_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = nthreads / 32 * iters * 8 * (16 * 8 * 8 * 2.0);
            printf("{\"bench\": \"dmma_m16n8k8\", \"warps_per_cta\": %d, \"ctas\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", warps, blocks, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma16816_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = nthreads / 32 * iters * 8 * (16 * 8 * 16 * 2.0);
            printf("{\"bench\": \"dmma_m16n8k16\", \"warps_per_cta\": %d, \"ctas\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", warps, blocks, ms, fl / ms * 1e-9);
        }
    }
    // smem-fed complex tile kernel
    for (int warps = 4; warps <= 8; warps *= 2) {
        int threads = warps * 32;
        int smem = 2 * 64 * 34 * 16;
        cudaFuncSetAttribute(zgemm_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int cps = 1; cps <= 2; ++cps) {
            int blocks = sms * cps;
            int it2 = 256;
            float ms = time_ms([&] { zgemm_smem_kernel<<<blocks, threads, smem>>>(out, it2); }, 5);
            double fl = (double)blocks * warps * it2 * 16.0 * 64.0 * (8 * 8 * 4 * 2.0);
            printf("{\"bench\": \"zgemm_smem_dmma_tile32x32\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", warps, cps, ms, fl / ms * 1e-9);
        }
    }
    CK(cudaDeviceSynchronize());
    cudaFree(out);
    return 0;
}
