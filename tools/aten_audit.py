"""Which Python lines of the C2 train step still launch ATen kernels (copies, casts, adds, fills)?  torch.profiler with
stacks over one step; prints, per (ATen op, innermost adapter4rec_b200 frame), the launch count and device time."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import profile, ProfilerActivity
from adapter4rec_b200 import surgery
from adapter4rec_b200.model import BertModel, Model, TextConfigLite
from adapter4rec_b200.trainer import FlatAdamTrainer

which = sys.argv[1] if len(sys.argv) > 1 else "lora"
dev = torch.device("cuda", 0)
torch.manual_seed(1)
args = bench.make_args()
if which != "lora":
    args.adapter_type = "houslby"
    args.bert_adapter_down_size = 64
model = Model(args, bench.ITEMS, True, BertModel(TextConfigLite())).to(dev)
surgery.freeze_all(model)
surgery.insert_adapters(model, args)
model.train()
tr = FlatAdamTrainer(model, 1e-4, 1e-5, 1e-4, 1e-4, users_per_pass=128)
gen = torch.Generator().manual_seed(1)
cat = bench.synth_catalogue(gen)
x, m = bench.synth_batch(cat, 128, gen)
x, m = x.to(dev), m.to(dev)
for _ in range(2):
    tr.train_step(x, m)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.train_step(x, m)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type.name != "CPU" or not ev.name.startswith("aten::"):
        continue
    dt = sum(k.duration for k in ev.kernels) if ev.kernels else 0.0
    if not ev.kernels:
        continue
    frame = next((f for f in (ev.stack or []) if "adapter4rec_b200" in f or "bench.py" in f), "?")
    key = (ev.name, frame.split("adapter4rec_b200/")[-1][:90])
    agg[key][0] += len(ev.kernels)
    agg[key][1] += dt
tot = sum(v[1] for v in agg.values())
print("ATen kernels in one %s step (128 users): %d launches, %.0f us device time" % (which, sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6d  %8.0f us  %-28s %s" % (v[0], v[1], k[0], k[1]))
