"""Repeated-launch determinism check of the tcgen05 mid-length attention (forward and backward): the kernels use no atomics, so
every launch on the same inputs must reproduce the first one bit for bit; a hand-over race (TMA box reuse, TMEM column reuse,
mbarrier phase slip) shows up as a differing launch.   python tools/attn_stress.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
for (N, L, heads) in [(704, 197, 12), (352, 207, 12), (96, 256, 12), (512, 129, 6), (1024, 64, 12)]:
    H = heads * 64
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = torch.randn(N * L, 3 * H, generator=g, device="cuda").to(torch.bfloat16)
    dctx = torch.randn(N * L, H, generator=g, device="cuda").to(torch.bfloat16)
    out0, lse0 = ops.attn_small_fwd(qkv, N, L, heads, 64, want_lse=True)
    dq0 = ops.attn_small_bwd(qkv, dctx, N, L, heads, 64, lse=lse0, ctx=out0)
    torch.cuda.synchronize()
    n = max(10, iters * 704 * 197 // (N * L))
    nf = nb = 0
    for i in range(n):
        out, lse = ops.attn_small_fwd(qkv, N, L, heads, 64, want_lse=True)
        dq = ops.attn_small_bwd(qkv, dctx, N, L, heads, 64, lse=lse0, ctx=out0)
        if not (torch.equal(out, out0) and torch.equal(lse, lse0)):
            nf += 1
        if not torch.equal(dq, dq0):
            nb += 1
    torch.cuda.synchronize()
    print("N=%d L=%d heads=%d: %d launches, forward differs %d x, backward differs %d x" % (N, L, heads, n, nf, nb), flush=True)
    bad += nf + nb
print("ALL OK" if bad == 0 else "FAILED")
sys.exit(0 if bad == 0 else 1)
