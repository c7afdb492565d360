"""Shared by the -m gpu parity tests and tools/measure_parity.py: run one seeded case through the CUDA path (drop-in
classes -> C ABI) and through the fp32 CPU oracle, and return the error figures the tests bound.

Tolerances are NOT free constants: tools/measure_parity.py records what this code measures on a B200 in
tests/golden/parity_measured.json, and a test allows `MARGIN` x the recorded figure (floored at `FLOOR`, the resolution
below which an fp32-vs-bf16 comparison carries no information).  Every trainable tensor is bounded individually — there
is no "negligible tensor" exemption and no noise-floor fallback."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.join(HERE, "golden"), os.path.join(os.path.dirname(HERE), "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
import cases_cv  # noqa: E402
import transrec_oracle as O  # noqa: E402

MEASURED_PATH = os.path.join(HERE, "golden", "parity_measured.json")
MARGIN = 2.0
# floors: a relative figure below 1e-3 is under one bf16 ulp (2^-8 = 3.9e-3) of the compared quantity; tensors whose oracle
# gradient is analytically zero (key-projection biases: softmax is invariant to a per-query constant) are bounded absolutely
FLOOR = {"loss_rel": 1e-3, "emb_max_abs": 4e-3, "emb_rel_l2": 1e-3, "grad_all_rel": 2e-3, "tensor_err_over_total": 1e-4}
# a tensor whose OWN relative error is under one bf16 epsilon (2^-7) is at the resolution of the activations its gradient is
# a sum of: its bound never drops below BF16_EPS x (its share of ||g_all||), whatever the recorded figure was
BF16_EPS = 2.0 ** -7


def measured():
    with open(MEASURED_PATH) as f:
        return json.load(f)


def bound(table, case, key):
    """allowed error for `key` of `case`: MARGIN x the recorded measurement, floored"""
    return max(MARGIN * float(table[case][key]), FLOOR[key])


def grad_figures(train, got, want):
    """per-tensor and aggregate gradient errors.  The per-tensor figure is ||g - g_oracle|| / ||g_oracle_all||: it bounds
    every tensor's absolute error (direction included) on the one scale that is meaningful for an optimizer step, and is
    well defined for tensors whose own gradient is (analytically) zero."""
    total = float(torch.cat([want[k].flatten() for k in train]).norm())
    per = {}
    for k in train:
        g, og = got[k].float().cpu(), want[k]
        per[k] = {"err_over_total": float((g - og).norm()) / total,
                  "rel": float((g - og).norm() / (og.norm() + 1e-30)),
                  "cos": float((g * og).sum() / (g.norm() * og.norm() + 1e-30)),
                  "share": float(og.norm()) / total}
    allg = torch.cat([got[k].float().cpu().flatten() for k in train])
    allo = torch.cat([want[k].flatten() for k in train])
    return per, float((allg - allo).norm() / allo.norm()), float((allg * allo).sum() / (allg.norm() * allo.norm()))


def summarise(per):
    """informational: over the tensors that carry at least 1 % of the gradient, the worst relative error and cosine"""
    big = [v for v in per.values() if v["share"] >= 1e-2]
    return {"tensor_rel_big": max(v["rel"] for v in big) if big else 0.0,
            "tensor_cos_big": min(v["cos"] for v in big) if big else 1.0}


def table_entry(fig):
    """what tools/measure_parity.py records for one case"""
    e = {k: fig[k] for k in ("loss", "oracle_loss", "loss_rel", "emb_max_abs", "emb_rel_l2")}
    if "per_tensor" in fig:
        e.update(grad_all_rel=fig["grad_all_rel"], grad_all_cos=fig["grad_all_cos"], tensor_rel_big=fig["tensor_rel_big"],
                 tensor_cos_big=fig["tensor_cos_big"],
                 per_tensor={k: v["err_over_total"] for k, v in fig["per_tensor"].items()})
    return e


def text_case(kind, full=False, unpad=False):
    """returns (figures, objects) for one text-tree case; figures = loss_rel, emb_max_abs, emb_rel_l2, grad_all_rel,
    grad_all_cos, per-tensor dict"""
    from test_model_gpu import build_gpu_model, oracle_setup
    c = cases.full_case(kind) if full else cases.tiny_case(kind)
    sd = cases.build_state_dict(c)
    items = cases.build_item_content(c)
    sample_items, log_mask, _ = cases.build_batch(c, items)
    rows = sample_items.view(-1, 2 * c.L)
    cfg, rec = oracle_setup(c)
    osd = {k: v.clone() for k, v in sd.items()}
    train = cases.trainable_keys(c, sd)
    for k in train:
        osd[k].requires_grad_(True)
    oloss = O.model_forward(rows, log_mask, osd, cfg, rec, cpc=c.cpc)
    if train:
        oloss.backward()
    model, args = build_gpu_model(c, sd)
    model.eval()    # parity is defined without dropout
    from adapter4rec_b200.data_utils.metrics import core_model
    core = core_model(model)
    if unpad:
        core.bert_encoder.text_encoders.title.bert_model.unpad = True
    loss = model(rows.cuda(), log_mask.cuda(), 0)
    fig = {"loss": float(loss.detach()), "oracle_loss": float(oloss.detach())}
    fig["loss_rel"] = abs(fig["loss"] - fig["oracle_loss"]) / abs(fig["oracle_loss"])
    if train:
        loss.backward()
        params = dict(model.named_parameters())
        missing = [k for k in train if params[k].grad is None]
        assert not missing, "no gradient for %s" % missing
        per, rel, cos = grad_figures(train, {k: params[k].grad for k in train}, {k: osd[k].grad for k in train})
        fig.update(grad_all_rel=rel, grad_all_cos=cos, per_tensor=per, **summarise(per))
        fig["frozen_clean"] = all(p.grad is None for n, p in params.items() if n not in train)
    with torch.no_grad():
        emb = core.bert_encoder(items.cuda()).float().cpu()
        oemb = O.item_embeddings(items, sd, cfg, rec, batch=16)
    d = emb[1:] - oemb[1:]       # row 0 = the padding item, never reaches the loss
    fig["emb_max_abs"] = float(d.abs().max())
    fig["emb_rel_l2"] = float(d.norm() / oemb[1:].norm())
    fig["emb_finite"] = bool(torch.isfinite(emb).all())
    return fig, dict(c=c, sd=sd, osd=osd, model=model, args=args, items=items, oracle_emb=oemb, emb=emb, train=train)


def cv_case(kind, full=False):
    from test_model_cv_gpu import build_gpu_cv_model
    c = cases_cv.full_cv_case(kind) if full else cases_cv.tiny_cv_case(kind)
    sd = cases_cv.build_state_dict(c)
    images, log_mask = cases_cv.build_batch(c)
    cfg = O.VitConfig(hidden=c.hidden, layers=c.layers, heads=c.heads, patch=c.patch, eps=c.eps)
    rec = O.RecConfig(max_seq_len=c.S, embedding_dim=c.D, heads=c.rec_heads, blocks=c.blocks, parallel=c.parallel)
    osd = {k: v.clone() for k, v in sd.items()}
    train = sorted(set(cases_cv.trainable_keys(c, sd)))
    for k in train:
        osd[k].requires_grad_(True)
    if kind == "cv_prompt":
        for suffix in ("weight", "bias"):
            osd[O.VIT_PREFIX + "embeddings.patch_embeddings.projection." + suffix] = \
                osd[O.VIT_PREFIX + "embeddings.wte.patch_embeddings.projection." + suffix]
    oloss = O.cv_model_forward(images, log_mask, osd, cfg, rec)
    if train:
        oloss.backward()
    model = build_gpu_cv_model(c, sd)
    model.eval()
    loss = model(images.cuda(), log_mask.cuda(), 0)
    fig = {"loss": float(loss.detach()), "oracle_loss": float(oloss.detach())}
    fig["loss_rel"] = abs(fig["loss"] - fig["oracle_loss"]) / abs(fig["oracle_loss"])
    if train:
        loss.backward()
        params = dict(model.named_parameters())
        missing = [k for k in train if params[k].grad is None]
        assert not missing, "no gradient for %s" % missing
        per, rel, cos = grad_figures(train, {k: params[k].grad for k in train}, {k: osd[k].grad for k in train})
        fig.update(grad_all_rel=rel, grad_all_cos=cos, per_tensor=per, **summarise(per))
        fig["frozen_clean"] = all(p.grad is None for n, p in params.items() if n not in train)
    with torch.no_grad():
        from adapter4rec_b200.data_utils.metrics import core_model
        emb = core_model(model).cv_encoder(images.cuda()).float().cpu()
        oemb = O.vit_encoder(images, {k: v.detach() for k, v in osd.items()}, cfg, rec)
    d = emb - oemb
    fig["emb_max_abs"] = float(d.abs().max())
    fig["emb_rel_l2"] = float(d.norm() / oemb.norm())
    fig["emb_finite"] = bool(torch.isfinite(emb).all())
    return fig, dict(c=c, sd=sd, osd=osd, model=model, images=images, oracle_emb=oemb, emb=emb, train=train)


def check_against_table(fig, table, case, has_grads):
    """the assertions every model-level parity test makes"""
    assert case in table, "no recorded measurement for %s: run tools/measure_parity.py on a B200" % case
    assert fig["emb_finite"]
    for key in ("loss_rel", "emb_max_abs", "emb_rel_l2"):
        assert fig[key] <= bound(table, case, key), "%s %s = %.3e > %.3e" % (case, key, fig[key], bound(table, case, key))
    if has_grads:
        assert fig["frozen_clean"], "a frozen parameter received a gradient"
        assert fig["grad_all_rel"] <= bound(table, case, "grad_all_rel"), \
            "%s aggregate gradient rel L2 %.4e > %.4e" % (case, fig["grad_all_rel"], bound(table, case, "grad_all_rel"))
        # every trainable tensor against ITS OWN recorded error (x MARGIN), on the ||g_all|| scale
        for k, v in fig["per_tensor"].items():
            lim = max(MARGIN * float(table[case]["per_tensor"][k]), FLOOR["tensor_err_over_total"], BF16_EPS * v["share"])
            assert v["err_over_total"] <= lim, "%s %s: ||dg|| / ||g_all|| = %.3e > %.3e (own share %.3e, rel %.3f, cos %.4f)" % (
                case, k, v["err_over_total"], lim, v["share"], v["rel"], v["cos"])
