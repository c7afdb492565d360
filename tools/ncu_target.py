"""Short target for `ncu --set full`: the four dominant GEMM shapes of one C2 pass, twice as warm-up, once captured."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
M = 161280
def r(*s): return (torch.randn(*s, device="cuda") * 0.05).to(torch.bfloat16)
x, w_qkv, t, bext = r(M, 768), r(2304, 768), r(M, 64), r(2304, 64)
w1, w2 = r(3072, 768), r(768, 3072)
b1, b2, bq = torch.randn(3072, device="cuda"), torch.randn(768, device="cuda"), torch.randn(2304, device="cuda")
u = torch.empty(M, 3072, dtype=torch.bfloat16, device="cuda")
for it in range(3):
    qkv = ops.gemm(x, w_qkv, bias=bq, a2=t, b2=bext)                       # QKV + LoRA K-extension
    f = ops.gemm(x, w1, bias=b1, epilogue=ops.EPI_GELU, aux=u)              # FFN1 + GELU (+ pre-activation)
    h = ops.gemm(f, w2, bias=b2, residual=x)                                # FFN2 + residual
    du = ops.gemm(h, w2.t().contiguous(), epilogue=ops.EPI_DGELU, aux=u)    # dFFN2 with fused GELU'
torch.cuda.synchronize()
print("done")
