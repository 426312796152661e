"""cuBLAS TF32 / bf16 GEMM throughput measured like MEASURED_PEAKS.json's bf16 number (torch.matmul 8192^3, best of 10 and a
4 s sustained loop) - the tensor-pipe roofline denominator for the kind::tf32 convolution kernels."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")
out = {}
for name, dt in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=dt); b = torch.randn(n, n, device=dev, dtype=dt)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20): a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out[name + "_tflops"] = 2 * n ** 3 / best / 1e9
    out[name + "_tflops_sustained"] = 2 * n ** 3 * k / e0.elapsed_time(e1) / 1e9
print(json.dumps(out))
