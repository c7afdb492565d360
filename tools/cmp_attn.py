import sys, torch
sys.path.insert(0, '/root/repo')
from adapter4rec_b200 import ops
torch.manual_seed(0)
for (N, L, heads, sc) in ((30, 37, 12, 0.2), (30, 37, 12, 3.0), (8, 197, 12, 3.0), (8, 197, 12, 0.2)):
    H = heads * 64
    qkv = (torch.randn(N * L, 3 * H, device='cuda') * sc).to(torch.bfloat16)
    dctx = (torch.randn(N * L, H, device='cuda') * 0.01).to(torch.bfloat16)
    q, k, v = [t.view(N, L, heads, 64).transpose(1, 2).double() for t in qkv.split(H, 1)]
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
    ref = (att @ v).transpose(1, 2).reshape(N * L, H)
    ref.backward(dctx.double())
    dref = torch.cat([t.grad.transpose(1, 2).reshape(N * L, H) for t in (q, k, v)], 1).detach()
    ref = ref.detach()
    ones = torch.ones(N, L, device='cuda')
    res = {}
    for name, mask in (("new", None), ("old", ones)):
        out, lse = ops.attn_small_fwd(qkv, N, L, heads, 64, mask=mask, want_lse=True)
        d = ops.attn_small_bwd(qkv, dctx, N, L, heads, 64, mask=mask, lse=lse, ctx=out)
        res[name] = (out, lse, d)
        e1 = float((out.double() - ref).norm() / ref.norm())
        e2 = [float((d[:, i*H:(i+1)*H].double() - dref[:, i*H:(i+1)*H]).norm() / dref[:, i*H:(i+1)*H].norm()) for i in range(3)]
        print(N, L, heads, sc, name, "fwd rel %.5f" % e1, "dq/dk/dv rel", ["%.5f" % x for x in e2], flush=True)
    print("   new-old max abs: out %.3g lse %.3g dqkv %.3g" % tuple(float((a.float() - b.float()).abs().max()) for a, b in zip(res["new"], res["old"])))
