"""Full fine-tuning step (fine_tune_to = all, pooler frozen as Pretraining/Text/run.py:48-64 does) of SASRec + BERT-base on
the C2 data shapes: forward + backward (data AND weight gradients of all 12 layers, embedding tables) + Adam over the
110 M parameters.  CUDA-event timing, 3 rotating resident batches (activations >> L2).  Under torchrun every rank trains its own
batch and the 440 MB of gradients are reduced either by ONE blocking all-reduce after the backward (--overlap 0) or in <= 25 MB
buckets issued from gradient hooks while the backward runs (--overlap 1, the default of FlatAdamTrainer); time = max over ranks.

    python tools/bench_full_ft.py [--users 128] [--steps 4]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_full_ft.py --overlap 1"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
from adapter4rec_b200 import lib, surgery  # noqa: E402
from adapter4rec_b200.model import BertModel, Model, TextConfigLite  # noqa: E402
from adapter4rec_b200.trainer import FlatAdamTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=128)
ap.add_argument("--users-per-pass", type=int, default=64)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--overlap", type=int, default=1)
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
lib.get_lib()
torch.manual_seed(123456)
args = B.make_args()
args.adding_adapter_to, args.fine_tune_to = "None", "all"
model = Model(args, B.ITEMS, True, BertModel(TextConfigLite())).to(dev)
for n, p in model.named_parameters():
    p.requires_grad = "pooler" not in n
model.train()
trainer = FlatAdamTrainer(model, args.lr, args.fine_tune_lr, args.adapter_bert_lr, args.adapter_sasrec_lr,
                          users_per_pass=a.users_per_pass, overlap=bool(a.overlap))
gen = torch.Generator().manual_seed(1 + rank)
cat = B.synth_catalogue(gen)
res = [tuple(t.to(dev) for t in B.synth_batch(cat, a.users, gen)) for _ in range(3)]
for i in range(a.warmup):
    loss = trainer.train_step(*res[i % 3])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
l0 = lib.get_lib().a4r_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    loss = trainer.train_step(*res[i % 3])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
tokens = a.users * 42 * B.L
flops = 3 * 12 * (14155776 + 92160) * tokens          # forward + dgrad + wgrad (no LoRA term)
if rank == 0:
  print(json.dumps({"what": "full fine-tune step, SASRec + BERT-base, C2 shapes", "n_gpus": world, "overlap": bool(a.overlap),
                  "buckets": len(getattr(trainer, "buckets", [])), "users_per_step": a.users * world,
                  "users_per_pass": a.users_per_pass, "ms_per_step": ms, "user_seqs_per_s": a.users * world / (ms / 1e3),
                  "model_tflops": flops / (ms / 1e3) / 1e12, "trainable_params": trainer.num_trainable,
                  "loss": float(loss), "launches_per_step": (lib.get_lib().a4r_launch_count() - l0) / a.steps,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
if world > 1:
    dist.destroy_process_group()
