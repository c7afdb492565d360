"""ncu target: LayerNorm forward / backward at the C2 pass shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
M, H = 161280, 768
x = [(torch.randn(M, H, device="cuda")).to(torch.bfloat16) for _ in range(2)]
g, b = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
for i in range(3):
    y, _, mean, rstd = ops.layernorm_fwd(x[i % 2], g, b, 1e-12)
    dz = ops.layernorm_bwd(y, x[i % 2], mean, rstd, g)
torch.cuda.synchronize()
print("done")
