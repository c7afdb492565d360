"""Times the fused Houlsby kernel (K5) alone at the C1/C3 pass shape (M = 161,280, H = 768, r = 64) in its three modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if "--lib" in sys.argv:      # A/B against a variant build of the library (tools/variants/*.so)
    import adapter4rec_b200.lib as _lib
    _i = sys.argv.index("--lib")
    _lib.LIB_PATH = os.path.abspath(sys.argv[_i + 1])
    del sys.argv[_i:_i + 2]
from adapter4rec_b200 import ops
M, H = int(os.environ.get('K5_M', '161280')), 768
def r(*s, sc=1.0): return (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)
h, inp = [r(M, H) for _ in range(3)], [r(M, H) for _ in range(3)]
wd, wu = r(64, H, sc=0.05), r(H, 64, sc=0.05)
bd, bu = torch.randn(64, device="cuda") * 0.1, torch.randn(H, device="cuda") * 0.1
g, b = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda") * 0.1
peak = 6540e9
for name, kw, nbytes in (("train_ln", dict(tail=0, save=True, act="relu"), 4 * 1536 + 160), ("infer_ln", dict(tail=0, save=False, act="relu"), 3 * 1536),
                         ("train_res_gelu", dict(tail=1, save=True, act="gelu"), 3 * 1536 + 2 * 160)):
    gg, bb = (g, b) if kw["tail"] == 0 else (None, None)
    for i in range(3):
        ops.adapter_ln_fwd(h[i % 3], inp[i % 3], wd, bd, wu, bu, gg, bb, 1e-12, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        ops.adapter_ln_fwd(h[i % 3], inp[i % 3], wd, bd, wu, bu, gg, bb, 1e-12, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 12 * 1e3
    print("%s %.1f us  %.2f of HBM peak" % (name, us, M * nbytes / (us * 1e-6) / peak))
