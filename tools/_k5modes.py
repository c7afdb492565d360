import os, sys, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from adapter4rec_b200 import ops
    mode, M = sys.argv[1], int(sys.argv[2]); H = 768
    def r(*s, sc=1.0): return (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)
    h, inp = r(M, H), r(M, H)
    wd, wu = r(64, H, sc=0.05), r(H, 64, sc=0.05)
    bd, bu = torch.randn(64, device="cuda") * 0.1, torch.randn(H, device="cuda") * 0.1
    g, b = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda") * 0.1
    kw = {"train_ln": dict(tail=0, save=True), "infer_ln": dict(tail=0, save=False), "res": dict(tail=1, save=False), "none": dict(tail=2, save=False), "res_train": dict(tail=1, save=True)}[mode]
    gg, bb = (g, b) if kw["tail"] == 0 else (None, None)
    for i in range(20):
        ops.adapter_ln_fwd(h, inp if kw["tail"] != 2 else None, wd, bd, wu, bu, gg, bb, 1e-12, act="relu", **kw)
    torch.cuda.synchronize()
    print(mode, M, "ok")
else:
    for M in (100000, 161280):
        for mode in ("none", "res", "res_train", "infer_ln", "train_ln"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), mode, str(M)], capture_output=True, text=True, timeout=120)
            print((r.stdout.strip() or (mode + " " + str(M) + " FAILED: " + r.stderr.strip().splitlines()[-1][:120])))
