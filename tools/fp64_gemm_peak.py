"""Measures the FP64 GEMM peaks MEASURED_PEAKS.json does not carry (cuBLAS DGEMM / ZGEMM through torch.matmul),
used as the denominator of the momentum-projection roofline ("of measured").  Prints one JSON object."""
import json
import torch

def bench(dtype, n, flops_per_mac, reps=5):
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return flops_per_mac * n ** 3 / best / 1e9

out = {"gpu": torch.cuda.get_device_name(0),
       "dgemm_8192_tflops": bench(torch.float64, 8192, 2),
       "zgemm_4096_tflops": bench(torch.complex128, 4096, 8),
       # the projection's shape: M = Lt*16*nLoop, N = Nmom, K = V3 (config 3: 25344 x 33 x 13824)
       }
M, N, K = 25344, 33, 13824
a = torch.randn(K, M, device="cuda", dtype=torch.complex128)   # column-major M x K
p = torch.randn(N, K, device="cuda", dtype=torch.complex128)   # column-major K x N
for _ in range(2):
    torch.matmul(p, a)
torch.cuda.synchronize()
best = 1e30
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(p, a); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["zgemm_momproj_shape_tflops"] = 8.0 * M * N * K / best / 1e9
out["zgemm_momproj_shape_ms"] = best
print(json.dumps(out))
