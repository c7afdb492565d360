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
# Tolerances: tests/parity_util.py (2 x the error recorded on a B200 in tests/golden/parity_measured.json).


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
    import parity_util as P
    fig, obj = P.cv_case(kind)
    gold = torch.load(os.path.join(HERE, "golden", "transrec_%s.pt" % kind), weights_only=False)
    assert abs(fig["oracle_loss"] - float(gold["loss"])) <= 2e-5 * abs(float(gold["loss"]))
    assert float((obj["oracle_emb"] - gold["item_emb"]).abs().max()) <= 2e-5 * float(gold["item_emb"].abs().max()) + 1e-6
    P.check_against_table(fig, P.measured(), "cv/" + kind, bool(obj["train"]))


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
