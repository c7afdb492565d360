import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
N, L, heads, dh = 352, 197, 12, 64
H = heads * dh
qkv = (torch.randn(N * L, 3 * H, device="cuda")).to(torch.bfloat16)
dctx = torch.randn(N * L, H, device="cuda").to(torch.bfloat16)
for it in range(3):
    out, lse = ops.attn_small_fwd(qkv, N, L, heads, dh, want_lse=True)
    dq = ops.attn_small_bwd(qkv, dctx, N, L, heads, dh, lse=lse, ctx=out)
torch.cuda.synchronize()
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record(); out, lse = ops.attn_small_fwd(qkv, N, L, heads, dh, want_lse=True); e1.record()
dq = ops.attn_small_bwd(qkv, dctx, N, L, heads, dh, lse=lse, ctx=out); e2.record(); torch.cuda.synchronize()
print("fwd %.3f ms  bwd %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2)))
