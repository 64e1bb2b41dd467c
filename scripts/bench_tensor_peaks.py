"""Dense matmul peaks of the box by input format (cuBLAS through torch, 8192^3, best of 10 and sustained over 2 s) --
the denominators behind the precision-mode ceilings of DESIGN.md 3.1 (SURVEY 8d asked for the TF32 figure).  One JSON line."""
import json
import time

import torch

N = 8192
out = {"n": N}
for name, dtype, tf32 in (("bf16", torch.bfloat16, False), ("fp16", torch.float16, False), ("tf32", torch.float32, True),
                          ("fp32", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(N, N, device="cuda", dtype=dtype)
    b = torch.randn(N, N, device="cuda", dtype=dtype)
    c = torch.empty(N, N, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_tflops_burst"] = 2 * N ** 3 / (best * 1e-3) / 1e12
    if name == "fp32":
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, n = time.time(), 0
    e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(10):
            torch.matmul(a, b, out=c)
        n += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out[name + "_tflops_sustained"] = n * 2 * N ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps(out))
