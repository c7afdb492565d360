"""tcgen05 wgrad vs the mma.sync split-M kernel at the C2 / full-fine-tune shapes (timed alone, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
M = 161280
def r(*s): return (torch.randn(*s, device="cuda") * 0.05).to(torch.bfloat16)
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for (N, K) in [(2304, 64), (64, 768), (768, 768), (3072, 768), (768, 3072), (768, 64), (64, 64)]:
    a, b = r(M, N), r(M, K)
    us_tc = t(lambda: ops.wgrad_tc(a, b))
    us_old = t(lambda: ops.wgrad_mma_sync(a, b), it=2)
    fl = 2.0 * M * N * K
    by = M * (N + K) * 2
    err = float((ops.wgrad_tc(a, b) - ops.wgrad_mma_sync(a, b)).abs().max())
    print("N=%5d K=%5d  tc %8.1f us (%7.1f TFLOP/s, %6.0f GB/s)   mma.sync %8.1f us   max|diff| %.3g" % (N, K, us_tc, fl / us_tc / 1e6, by / us_tc / 1e3, us_old, err), flush=True)
