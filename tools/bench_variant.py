"""Runs one of bench.py's non-headline training configurations alone (for ncu launch lists / quick A-B timing):
    python tools/bench_variant.py c1 [--users 128] [--steps 2]
    python tools/bench_variant.py c3 [--users 32]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--lib" in sys.argv:      # A/B against a variant build of the library (tools/variants/*.so)
    import adapter4rec_b200.lib as _lib
    _i = sys.argv.index("--lib")
    _lib.LIB_PATH = os.path.abspath(sys.argv[_i + 1])
    del sys.argv[_i:_i + 2]
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("which", choices=["c1", "c3", "c4"])
ap.add_argument("--users", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
if a.which == "c4":
    out = bench.bench_c4_roberta_prompt_cpc(dev, 1, 0, a.steps, users=a.users or 256)
elif a.which == "c1":
    out = bench.bench_c1_houlsby(dev, 1, 0, a.steps, users=a.users or 256)
else:
    out = bench.bench_c3_vit(dev, 1, 0, a.steps, users=a.users or 64)
print(json.dumps(out))
