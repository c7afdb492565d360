"""-m gpu: image tree (ViT item encoder + SASRec + loss, forward and backward) through the drop-in classes and the C ABI
against the oracle and the goldens of the unmodified Downstream/CV reference.  17-token cases run the short-sequence
attention kernel, 37/40-token cases the mid-length kernel (the one ViT-B/16-224's 197 tokens use)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import cases_cv  # noqa: E402
import transrec_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
KINDS = list(cases_cv.CV_ALL_KINDS)
# Same contract as tests/test_model_gpu.py, except the aggregate gradient bound: with 3 users per batch the LoRA
# gradient of the 37-token case is a sum of few bf16-rounded terms and its aggregate error sits at 4.5-5.7 % depending on
# where the attention kernel rounds (the masked and the unmasked mid-length kernels are both within 2.4e-3 of an fp64
# attention per kernel and differ from each other by one bf16 ulp: tools/cmp_attn.py), so the bound is 7 % here.
LOSS_RTOL, EMB_ATOL, GRAD_REL_L2, GRAD_ALL_REL_L2 = 2e-2, 3e-2, 0.15, 7e-2


def build_gpu_cv_model(c, sd):
    from adapter4rec_b200 import surgery
    from adapter4rec_b200.cv import Model, ViTConfigLite, ViTForImageClassification
    from adapter4rec_b200.model.layers import Linear
    args = cases_cv.reference_args(c)
    cfg = ViTConfigLite(hidden_size=c.hidden, num_hidden_layers=c.layers, num_attention_heads=c.heads,
                        intermediate_size=c.inter, image_size=c.image, patch_size=c.patch, layer_norm_eps=c.eps)
    net = ViTForImageClassification(cfg)
    net.classifier = Linear(c.hidden, args.embedding_dim)            # run_adapter.py:293-294
    model = Model(args, 100, True, net).cuda()
    surgery.freeze_all(model)
    if c.kind == "cv_full_ft":                                   # fine_tune_to = all: nothing frozen
        for p in model.parameters():
            p.requires_grad = True
    elif c.kind != "cv_base":
        model = surgery.insert_adapters_cv(model, args)          # compacter returns the CompacterModel wrapper
    assert set(model.state_dict().keys()) == set(sd.keys()), "state_dict keys must equal the reference's"
    model.load_state_dict(sd)
    got = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    assert got == sorted(set(cases_cv.trainable_keys(c, sd)))
    return model


@pytest.mark.parametrize("kind", KINDS)
def test_cv_train_step_matches_oracle_and_reference(kind):
    c = cases_cv.tiny_cv_case(kind)
    sd = cases_cv.build_state_dict(c)
    gold = torch.load(os.path.join(HERE, "golden", "transrec_%s.pt" % kind), weights_only=False)
    model = build_gpu_cv_model(c, sd)
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
    assert abs(float(oloss) - float(gold["loss"])) <= 2e-5 * abs(float(gold["loss"]))
    if train:
        oloss.backward()
    model.eval()    # parity is defined without dropout
    loss = model(images.cuda(), log_mask.cuda(), 0)
    lv, ov = float(loss.detach()), float(oloss.detach())
    assert abs(lv - ov) <= LOSS_RTOL * abs(ov), "loss %.6f vs oracle %.6f" % (lv, ov)
    with torch.no_grad():
        from adapter4rec_b200.data_utils.metrics import core_model
        emb = core_model(model).cv_encoder(images.cuda()).float().cpu()
    ref_emb = gold["item_emb"]   # ViT embeddings reach |x| ~ 2.5: absolute 3e-2 plus 2e-2 relative (bf16 activations)
    assert bool(((emb - ref_emb).abs() <= EMB_ATOL + 2e-2 * ref_emb.abs()).all()), float((emb - ref_emb).abs().max())
    if train:
        loss.backward()
        params = dict(model.named_parameters())
        total_norm = float(torch.cat([osd[k].grad.flatten() for k in train]).norm())
        for k in train:
            g, og = params[k].grad, osd[k].grad
            assert g is not None, "no gradient for " + k
            g = g.float().cpu()
            rel = float((g - og).norm() / (og.norm() + 1e-12))
            cos = float((g * og).sum() / (g.norm() * og.norm() + 1e-20))
            # tensors whose whole gradient is < 1 % of the total (e.g. the SASRec query-side LoRA factors: the oracle
            # itself moves them by 5-10 % when only the WEIGHTS are rounded to bf16) are bounded absolutely instead
            negligible = float((g - og).norm()) <= 5e-3 * total_norm
            # cos >= 0.985: a relative L2 error of 0.15 already implies cos >= sqrt(1 - 0.15^2) = 0.9887, so a tighter
            # cosine bound would silently override the stated L2 tolerance (3 users x 4 positions: the SASRec-side
            # rank-8 factors sit at rel 0.14 / cos 0.990 from bf16 rounding alone)
            assert (rel <= GRAD_REL_L2 and cos >= 0.985) or negligible, "%s: rel L2 %.4f cos %.5f" % (k, rel, cos)
        allg = torch.cat([params[k].grad.float().cpu().flatten() for k in train])
        allo = torch.cat([osd[k].grad.flatten() for k in train])
        err = float((allg - allo).norm() / allo.norm())
        if err > GRAD_ALL_REL_L2:
            # ill-conditioned case: measure the rounding-noise floor = the SAME fp32 oracle with nothing but the weights
            # and the images rounded to bf16 (what the GPU path's cached operands are).  cv_pfeiffer_ver2 moves by 6.2 %
            # under that rounding alone (tools/grad_diag_cv.py); the GPU path must stay within 1.25x of the floor.
            bsd = {k: v.clone().to(torch.bfloat16).float() for k, v in sd.items()}
            for k in train:
                bsd[k].requires_grad_(True)
            O.cv_model_forward(images.to(torch.bfloat16).float(), log_mask, bsd, cfg, rec).backward()
            allb = torch.cat([bsd[k].grad.flatten() for k in train])
            floor = float((allb - allo).norm() / allo.norm())
            assert err <= 1.25 * floor, "aggregate gradient error %.4f vs bf16-weight noise floor %.4f" % (err, floor)
        assert all(p.grad is None for n, p in params.items() if n not in train)


def test_patchify_and_assemble_kernels():
    from adapter4rec_b200 import ops
    g = torch.Generator().manual_seed(1)
    N, C, R, ps, H = 5, 3, 64, 16, 128
    img = (torch.rand((N, C, R, R), generator=g) * 2 - 1).cuda()
    w = (torch.randn((H, C, ps, ps), generator=g) * 0.05).cuda()
    b = torch.randn(H, generator=g).cuda()
    patches = ops.patchify(img, ps)
    ref_p = torch.nn.functional.unfold(img, ps, stride=ps).transpose(1, 2).reshape(N * (R // ps) ** 2, C * ps * ps)
    assert torch.equal(patches, ref_p.to(torch.bfloat16))
    pe = ops.gemm(patches, w.view(H, -1).to(torch.bfloat16), bias=b)
    ref = torch.nn.functional.conv2d(img, w, b, stride=ps).flatten(2).transpose(1, 2).reshape(-1, H)
    assert float((pe.float() - ref).abs().max()) <= 3e-2
    P, T = (R // ps) ** 2, 3
    cls, pos, prompt = [torch.randn(s, generator=g).to(torch.bfloat16).cuda() for s in ((H,), (P + 1, H), (T, H))]
    out = ops.vit_assemble(pe, cls, pos, prompt, N, P).view(N, 1 + P + T, H)
    exp = torch.cat([torch.cat([cls.float().expand(N, 1, H), pe.float().view(N, P, H)], 1) + pos.float(),
                     prompt.float().expand(N, T, H)], 1)
    assert float((out.float() - exp).abs().max()) <= 2e-2
    assert torch.equal(out[:, 1 + P:], prompt.expand(N, T, H))
