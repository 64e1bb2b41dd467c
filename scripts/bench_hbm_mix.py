"""HBM throughput of plain torch kernels by read : write mix (diagnostic behind DESIGN.md 3.3: what a store-heavy
epilogue can expect).  Prints one JSON line."""
import json
import torch

def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3

N = 1 << 28  # 1 GiB of fp32
x = torch.empty(N, device="cuda", dtype=torch.float32).normal_()
y = torch.empty_like(x)
z = torch.empty_like(x)
out = {}
out["write_only_fill_GBs"] = 4 * N / timed(lambda: y.fill_(1.0)) / 1e9
out["read_only_sum_GBs"] = 4 * N / timed(lambda: x.sum()) / 1e9
out["copy_1r1w_GBs"] = 8 * N / timed(lambda: y.copy_(x)) / 1e9
out["add_2r1w_GBs"] = 12 * N / timed(lambda: torch.add(x, y, out=z)) / 1e9
h = x.view(-1)[: N // 2]
out["cast_1r_half_w_GBs"] = (4 * N + 2 * N) / timed(lambda: x.to(torch.float16)) / 1e9
print(json.dumps(out))
