"""C3 (BASELINE.json configs[2]): SASRec + ViT-B/16-224 with Houlsby adapters (r=64), synthetic HM-shape images,
seq_len 10 (22 images per user), bf16.  Prints one JSON line (not the driver's headline bench; recorded in profiles/)."""
import argparse, json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from adapter4rec_b200 import lib, ops, surgery
from adapter4rec_b200.cv import Model, ViTConfigLite, ViTForImageClassification
from adapter4rec_b200.model.layers import Linear
from adapter4rec_b200.trainer import FlatAdamTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=64)
ap.add_argument("--users-per-pass", type=int, default=32)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--adapter", default="houslby")
a = ap.parse_args()
S, D = 10, 64
args = types.SimpleNamespace(max_seq_len=S, l2_weight=0, embedding_dim=D, num_attention_heads=2, drop_rate=0.1,
                             transformer_block=2, CV_model_load="vit-base-patch16-224", cv_adapter_down_size=64,
                             adapter_down_size=16, adapter_dropout_rate=0.1, adapter_activation="RELU", n_tokens=10,
                             adapter_type=a.adapter, adding_adapter_to="all", is_serial="True", finetune_layernorm="None")
torch.manual_seed(12345)
dev = torch.device("cuda", 0)
net = ViTForImageClassification(ViTConfigLite())
net.classifier = Linear(768, D)
model = Model(args, 1000, True, net).to(dev)
surgery.freeze_all(model)
surgery.insert_adapters_cv(model, args)
model.train()
trainer = FlatAdamTrainer(model, 1e-4, 1e-5, 5e-4, 1e-4, users_per_pass=a.users_per_pass)
n_img = a.users * (S + 1) * 2
g = torch.Generator(device=dev).manual_seed(1)
batches = [(torch.rand((n_img, 3, 224, 224), generator=g, device=dev) * 2 - 1, torch.ones((a.users, S), device=dev)) for _ in range(2)]
for i in range(a.warmup):
    trainer.train_step(*batches[i % 2])
torch.cuda.synchronize()
ops.gemm_profile_start()
l0 = lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    loss = trainer.train_step(*batches[i % 2])
e1.record()
torch.cuda.synchronize()
fl, gms, calls = ops.gemm_profile_stop()
ms = e0.elapsed_time(e1) / a.steps
tok = 197 + (10 if "prompt" in a.adapter else 0)
fwd_flops_img = 12 * tok * (14155776 + 4 * tok * 768 + (393216 if "hous" in a.adapter else 0)) + 196 * 2 * 768 * 768
print(json.dumps({"config": "C3: SASRec + ViT-B/16-224 %s, S=10, bf16" % a.adapter, "users_per_step": a.users,
                  "images_per_step": n_img, "ms_per_step": ms, "users_per_s": a.users / ms * 1e3,
                  "images_per_s": n_img / ms * 1e3, "model_tflops": 2 * fwd_flops_img * n_img / ms / 1e9,
                  "gemm_tflops": fl / gms / 1e9, "gemm_share": gms / (ms * a.steps), "gpu_launches": lib.launch_count() - l0,
                  "trainable_params": trainer.num_trainable, "loss": float(loss),
                  "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
