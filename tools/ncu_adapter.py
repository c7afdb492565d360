"""ncu target: the fused Houlsby adapter kernel (K5) at the C2/C3 pass shape, twice as warm-up, once captured."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
M, H = 161280, 768
def r(*s, sc=1.0): return (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)
h, inp = [r(M, H) for _ in range(2)], [r(M, H) for _ in range(2)]
wd, wu = r(64, H, sc=0.01), r(H, 64, sc=0.01)
bd, bu = torch.zeros(64, device="cuda"), torch.zeros(H, device="cuda")
g, b = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
for i in range(3):
    ops.adapter_ln_fwd(h[i % 2], inp[i % 2], wd, bd, wu, bu, g, b, 1e-12, act="relu", tail=0, save=True)
torch.cuda.synchronize()
print("done")
