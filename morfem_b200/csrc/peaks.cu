// Measurement helper (bench.py's roofline denominator): the FP64 tensor-pipe issue rate of the device the caller is on.
// MEASURED_PEAKS.json (driver-written) holds an HBM and a bf16 figure only; the kernels of this library are FP64, so the
// denominator of their compute roofline is measured here, in the run that reports the fraction.
#include "common.cuh"

namespace {

template <int ILP>
__global__ void dmma_issue_kernel(double* out, int iters, double a, double b) {
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
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // keeps the loop alive, never true in practice
}

}  // namespace

// Unlike every other entry point this one BLOCKS (it times its own launches with CUDA events on `stream`).
extern "C" int mf_peak_dmma_tflops(int iters, double* tflops_host, void* stream) {
    if (iters <= 0) MF_FAIL_ARG(1, "iters must be positive");
    if (!tflops_host) MF_FAIL_ARG(2, "tflops_host is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = mf_num_sms(), threads = 256, blocks = 2 * sms;
    double* out = nullptr;
    MF_CHECK_CUDA(cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    MF_CHECK_CUDA(cudaEventCreate(&e0));
    MF_CHECK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {                  // the first launches warm the clocks up; best of the rest
        cudaEventRecord(e0, st);
        dmma_issue_kernel<8><<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, st);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(out); MF_CHECK_CUDA(e); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    g_mf_launches.fetch_add(6, std::memory_order_relaxed);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    const double flops = (double)blocks * (threads / 32) * (double)iters * 8.0 * (8.0 * 8.0 * 4.0 * 2.0);
    *tflops_host = flops / ((double)best * 1e-3) * 1e-12;
    return 0;
}
