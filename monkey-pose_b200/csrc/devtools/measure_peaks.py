"""Development helper: dense-GEMM peaks of this box by the MEASURED_PEAKS.json method (torch.matmul 8192^3, 2*N^3
flops; best of 10 = burst, back to back for 4 s = sustained), for bf16 and for tf32 (fp32 inputs,
torch.backends.cuda.matmul.allow_tf32) -- BASELINE.md asks the builder to measure the tf32 peak.  Prints one JSON line."""
import json
import time

import torch

N = 8192
out = {"gpu": torch.cuda.get_device_name(0), "method": "torch.matmul %d^3, CUDA events" % N}
for name, dtype, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True), ("fp32", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(N, N, device="cuda", dtype=dtype)
    b = torch.randn(N, N, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, n = time.time(), 0
    e0.record()
    while time.time() - t0 < (4.0 if name != "fp32" else 1.5):
        for _ in range(10):
            a @ b
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    out[name + "_tflops_burst"] = round(2.0 * N ** 3 / (best * 1e-3) * 1e-12, 1)
    out[name + "_tflops_sustained"] = round(2.0 * N ** 3 * n / (e0.elapsed_time(e1) * 1e-3) * 1e-12, 1)
print(json.dumps(out))
