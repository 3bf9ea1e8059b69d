"""cuBLAS FP64 / complex128 GEMM throughput via torch.matmul (library reference point for the FP64 roofline)."""
import json, torch
def bench(dtype, n, flop_per_mac, reps=5):
    a = torch.randn(n, n, device="cuda", dtype=dtype); b = torch.randn(n, n, device="cuda", dtype=dtype)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return flop_per_mac * n ** 3 / best * 1e-9
print(json.dumps({"bench": "cublas_dgemm_8192", "tflops": bench(torch.float64, 8192, 2)}))
print(json.dumps({"bench": "cublas_zgemm_4096", "tflops": bench(torch.complex128, 4096, 8)}))
