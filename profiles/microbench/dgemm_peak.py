# cuBLAS DGEMM 8192^3 through torch: the measured FP64 denominator for the posterior/Cholesky rooflines.
import torch, json
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(2): torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e30
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
# sustained: back to back for ~3 s
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
k = max(3, int(3000 / best)); e0.record()
for _ in range(k): torch.matmul(a, b, out=c)
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / k
print(json.dumps({"dgemm_n": n, "burst_ms": best, "burst_tflops": 2 * n**3 / best * 1e-9,
                  "sustained_ms": sus, "sustained_tflops": 2 * n**3 / sus * 1e-9}))
