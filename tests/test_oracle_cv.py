"""CPU: pins the image-tree oracle functions against goldens of the unmodified Downstream/CV reference."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import cases_cv  # noqa: E402
import transrec_oracle as O  # noqa: E402

KINDS = list(cases_cv.CV_ALL_KINDS)


def load_case(kind):
    c = cases_cv.tiny_cv_case(kind)
    gold = torch.load(os.path.join(HERE, "golden", "transrec_%s.pt" % kind), weights_only=False)
    sd = cases_cv.build_state_dict(c)
    cfg = O.VitConfig(hidden=c.hidden, layers=c.layers, heads=c.heads, patch=c.patch, eps=c.eps)
    rec = O.RecConfig(max_seq_len=c.S, embedding_dim=c.D, heads=c.rec_heads, blocks=c.blocks, parallel=c.parallel)
    return c, gold, sd, cfg, rec


@pytest.mark.parametrize("kind", KINDS)
def test_cv_loss_grads_embeddings_match_reference(kind):
    c, gold, sd, cfg, rec = load_case(kind)
    train = sorted(set(cases_cv.trainable_keys(c, sd)))
    for k in train:
        sd[k] = sd[k].clone().requires_grad_(True)
    if kind == "cv_prompt":   # the shared patch projection appears under two names
        for suffix in ("weight", "bias"):
            sd[O.VIT_PREFIX + "embeddings.patch_embeddings.projection." + suffix] = \
                sd[O.VIT_PREFIX + "embeddings.wte.patch_embeddings.projection." + suffix]
    images, log_mask = cases_cv.build_batch(c)
    loss = O.cv_model_forward(images, log_mask, sd, cfg, rec)
    torch.testing.assert_close(loss.detach(), gold["loss"], rtol=2e-5, atol=2e-6)
    with torch.no_grad():
        torch.testing.assert_close(O.vit_encoder(images, sd, cfg, rec), gold["item_emb"], rtol=5e-5, atol=5e-6)
    if train:
        loss.backward()
        assert sorted(gold["grads"].keys()) == train
        for k in train:
            # fp32 summation order differs between the reference's einsum/bmm chain and the oracle's: the absolute
            # floor scales with the tensor's own magnitude (cv_compacter's loss of ~42 gives gradients of O(1))
            atol = max(2e-6, 1e-5 * float(gold["grads"][k].abs().max()))
            torch.testing.assert_close(sd[k].grad, gold["grads"][k], rtol=2e-4, atol=atol, msg=lambda m: k + ": " + m)
