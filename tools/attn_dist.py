import os, sys
sys.path.insert(0, "/root/repo")
import torch
from adapter4rec_b200 import ops
N, L, heads = 704, 197, 12
H = heads * 64
for scale in (0.05, 0.5, 1.0, 2.0):
    bufs = [(torch.randn(N * L, 3 * H, device="cuda") * scale).to(torch.bfloat16) for _ in range(3)]
    dctx = [torch.randn(N * L, H, device="cuda").to(torch.bfloat16) for _ in range(3)]
    for rot in (1, 3):
        for i in range(3): out, lse = ops.attn_small_fwd(bufs[i % rot], N, L, heads, 64, want_lse=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10): out, lse = ops.attn_small_fwd(bufs[i % rot], N, L, heads, 64, want_lse=True)
        e1.record(); torch.cuda.synchronize()
        tf = e0.elapsed_time(e1) * 100
        e0.record()
        for i in range(10): dq = ops.attn_small_bwd(bufs[i % rot], dctx[i % rot], N, L, heads, 64, lse=lse, ctx=out)
        e1.record(); torch.cuda.synchronize()
        tb = e0.elapsed_time(e1) * 100
        print("scale %.2f rot %d: fwd %.1f us  bwd %.1f us" % (scale, rot, tf, tb), flush=True)
