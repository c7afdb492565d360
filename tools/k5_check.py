"""Correctness + timing of one K5 implementation (A4R_K5_IMPL = 2 | 3) against torch: all tails, relu / gelu, ragged M."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adapter4rec_b200 import ops
torch.manual_seed(0)
H, r = 768, 64
def rnd(*s, sc=1.0): return (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)
wd, wu = rnd(r, H, sc=0.05), rnd(H, r, sc=0.05)
bd, bu = torch.randn(r, device="cuda") * 0.1, torch.randn(H, device="cuda") * 0.1
g, b = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda") * 0.1
ok = True
for M in (128, 1000, 20000):
    h, inp = rnd(M, H), rnd(M, H)
    for tail in (0, 1, 2):
        for act in ("relu", "gelu"):
            for save in (False, True):
                gg, bb = (g, b) if tail == 0 else (None, None)
                out, z, mean, rstd, s, u = ops.adapter_ln_fwd(h, inp if tail != 2 else None, wd, bd, wu, bu, gg, bb, 1e-12, act=act, tail=tail, save=save)
                torch.cuda.synchronize()
                pre = h.float() @ wd.float().t() + bd
                sr = torch.relu(pre) if act == "relu" else torch.nn.functional.gelu(pre)
                zr = h.float() + sr.to(torch.bfloat16).float() @ wu.float().t() + bu + (inp.float() if tail != 2 else 0)
                ref = torch.nn.functional.layer_norm(zr.to(torch.bfloat16).float(), (H,), g, b, 1e-12) if tail == 0 else zr
                e = float((out.float() - ref).abs().max())
                line = "M=%d tail=%d %s save=%d out err %.4f" % (M, tail, act, save, e)
                good = e < 0.08
                if save:
                    es = float((s.float() - sr).abs().max()); good &= es < 0.03; line += " s %.4f" % es
                    if tail == 0:
                        ez = float((z.float() - zr).abs().max()); good &= ez < 0.05; line += " z %.4f" % ez
                        em = float((mean - zr.to(torch.bfloat16).float().mean(1)).abs().max()); good &= em < 1e-3; line += " mean %.5f" % em
                    ones = ops.s_ext(s)[:, r]
                    good &= bool((ones == 1).all())
                if not good:
                    ok = False
                    print("BAD ", line)
print("ALL OK" if ok else "FAILED")
