"""Per-tensor gradient error of one CV golden case (GPU path vs fp32 oracle): python tools/grad_diag_cv.py cv_pfeiffer_ver2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import cases_cv
import transrec_oracle as O
from test_model_cv_gpu import build_gpu_cv_model

kind = sys.argv[1]
c = cases_cv.tiny_cv_case(kind)
sd = cases_cv.build_state_dict(c)
model = build_gpu_cv_model(c, sd)
images, log_mask = cases_cv.build_batch(c)
cfg = O.VitConfig(hidden=c.hidden, layers=c.layers, heads=c.heads, patch=c.patch, eps=c.eps)
rec = O.RecConfig(max_seq_len=c.S, embedding_dim=c.D, heads=c.rec_heads, blocks=c.blocks, parallel=c.parallel)
osd = {k: v.clone() for k, v in sd.items()}
train = sorted(set(cases_cv.trainable_keys(c, sd)))
for k in train:
    osd[k].requires_grad_(True)
O.cv_model_forward(images, log_mask, osd, cfg, rec).backward()
# same oracle with every weight rounded to bf16 (what the GPU path's cached operands are): the noise floor of rounding
bsd = {k: v.clone().to(torch.bfloat16).float() for k, v in sd.items()}
for k in train:
    bsd[k].requires_grad_(True)
O.cv_model_forward(images.to(torch.bfloat16).float(), log_mask, bsd, cfg, rec).backward()
model.eval()
model(images.cuda(), log_mask.cuda(), 0).backward()
params = dict(model.named_parameters())
tot = float(torch.cat([osd[k].grad.flatten() for k in train]).norm())
rows = []
for k in train:
    g, og, bg = params[k].grad.float().cpu(), osd[k].grad, bsd[k].grad
    rows.append((float((g - og).norm()) / tot, float((bg - og).norm()) / tot, float(og.norm()) / tot, k))
for r in sorted(rows, reverse=True)[:12]:
    print("gpu-err/total %.4f  bf16-weights-oracle-err/total %.4f  share %.3f  %s" % r)
allg = torch.cat([params[k].grad.float().cpu().flatten() for k in train])
allo = torch.cat([osd[k].grad.flatten() for k in train])
allb = torch.cat([bsd[k].grad.flatten() for k in train])
print("aggregate: gpu %.4f   oracle-with-bf16-weights %.4f" % (float((allg - allo).norm() / allo.norm()), float((allb - allo).norm() / allo.norm())))
