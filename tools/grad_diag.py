import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests/golden")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, cases_cv, transrec_oracle as O
from test_model_cv_gpu import build_gpu_cv_model
kind = "cv_lora"
c = cases_cv.tiny_cv_case(kind); sd = cases_cv.build_state_dict(c)
model = build_gpu_cv_model(c, sd)
images, log_mask = cases_cv.build_batch(c)
cfg = O.VitConfig(hidden=c.hidden, layers=c.layers, heads=c.heads, patch=c.patch, eps=c.eps)
rec = O.RecConfig(max_seq_len=c.S, embedding_dim=c.D, heads=c.rec_heads, blocks=c.blocks)
train = sorted(set(cases_cv.trainable_keys(c, sd)))
def oracle_grads(round_w):
    osd = {k: (v.to(torch.bfloat16).float() if round_w and v.dim() >= 2 else v.clone()) for k, v in sd.items()}
    for k in train: osd[k].requires_grad_(True)
    O.cv_model_forward(images, log_mask, osd, cfg, rec).backward()
    return {k: osd[k].grad for k in train}
og = oracle_grads(False); og16 = oracle_grads(True)
model.eval(); model(images.cuda(), log_mask.cuda(), 0).backward()
params = dict(model.named_parameters())
tot = torch.cat([og[k].flatten() for k in train]).norm()
for k in train:
    if "user_encoder" not in k: continue
    g = params[k].grad.float().cpu(); o = og[k]; o16 = og16[k]
    print("%-70s |o|=%.3e (%.1e of total) rel(gpu,o)=%.3f rel(o16,o)=%.3f rel(gpu,o16)=%.3f" % (k[-70:], o.norm(), o.norm()/tot, (g-o).norm()/o.norm(), (o16-o).norm()/o.norm(), (g-o16).norm()/o16.norm()))
