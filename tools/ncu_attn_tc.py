"""Short `ncu --set full` target: the tcgen05 mid-length attention at the C3 pass shape (704 images x 12 heads x 197 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
N, L, heads, dh = 704, 197, 12, 64
H = heads * dh
qkv = (torch.randn(N * L, 3 * H, device="cuda")).to(torch.bfloat16)
dctx = torch.randn(N * L, H, device="cuda").to(torch.bfloat16)
bwd = "--bwd" in sys.argv
for it in range(3):
    out, lse = ops.attn_small_fwd(qkv, N, L, heads, dh, want_lse=True)
    if bwd:
        dq = ops.attn_small_bwd(qkv, dctx, N, L, heads, dh, lse=lse, ctx=out)
torch.cuda.synchronize()
print("done")
