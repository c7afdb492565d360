"""Short target for `ncu --set full`: the eight GEMM shapes of one C2 pass (LoRA, M = 161,280 token rows), twice as warm-up,
once captured; SHAPES lists (M, N, K + K2, epilogue id) in launch order — tools/ncu_gemm_traffic.py pairs it with the report."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M = 161280
SHAPES = [(M, 64, 768, 0), (M, 2304, 832, 0), (M, 768, 768, 0), (M, 3072, 768, 1), (M, 768, 3072, 0), (M, 3072, 768, 3),
          (M, 64, 1536, 0), (M, 768, 2368, 0)]
if __name__ == "__main__":
    import torch
    from adapter4rec_b200 import ops
    def r(*s): return (torch.randn(*s, device="cuda") * 0.05).to(torch.bfloat16)
    x, w_qkv, acat, bext = r(M, 768), r(2304, 768), r(64, 768), r(2304, 64)
    wo, w1, w2 = r(768, 768), r(3072, 768), r(768, 3072)
    b1, b2, bq = torch.randn(3072, device="cuda"), torch.randn(768, device="cuda"), torch.randn(2304, device="cuda")
    u = torch.empty(M, 3072, dtype=torch.bfloat16, device="cuda")
    for it in range(3):
        t = ops.gemm(x, acat)                                                   # T = x A_cat^T (LoRA intermediate)
        qkv = ops.gemm(x, w_qkv, bias=bq, a2=t, b2=bext)                        # QKV + LoRA K-extension
        y = ops.gemm(x, wo, bias=b2, residual=x)                                # attention.output dense + residual
        f = ops.gemm(y, w1, bias=b1, epilogue=ops.EPI_GELU, aux=u)              # FFN1 + GELU (+ pre-activation)
        h = ops.gemm(f, w2, bias=b2, residual=y)                                # FFN2 + residual
        du = ops.gemm(h, w2.t().contiguous(), epilogue=ops.EPI_DGELU, aux=u)    # dFFN2 with fused GELU'
        bt = bext.t().contiguous()
        dt = ops.gemm(qkv[:, :768], bt[:, :768], a2=qkv[:, 1536:], b2=bt[:, 1536:])    # dT = dq B_q + dv B_v (key third skipped)
        dx = ops.gemm(qkv, w_qkv.t().contiguous(), a2=dt, b2=acat.t().contiguous())   # dx = [dqkv | dT] [W ; A_cat]
    torch.cuda.synchronize()
    print("done")
