"""HBM bandwidth by direction on this box: write-only (fill), read-only (sum), copy — the GELU-forward GEMM writes 2 B/FLOP-free
bytes (two [M, 3072] bf16 outputs = 2 GB at M = 161,280) and the question is which roof that is."""
import torch
n = 1 << 30                                    # 2 GB of bf16
x = torch.empty(n, dtype=torch.bfloat16, device="cuda")
y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
b = 2.0 * n
s = t(lambda: x.fill_(1.0)); print("write-only  %.0f GB/s" % (b / s / 1e9))
s = t(lambda: x.zero_()); print("memset      %.0f GB/s" % (b / s / 1e9))
s = t(lambda: y.copy_(x)); print("copy        %.0f GB/s (read + write)" % (2 * b / s / 1e9))
xf = x.view(torch.int16)
s = t(lambda: torch.sum(xf[: n // 2].view(torch.int32))); print("read-only   %.0f GB/s" % ((b / 2) / s / 1e9))
