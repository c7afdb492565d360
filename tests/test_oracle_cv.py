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

KINDS = ["cv_base", "cv_houlsby", "cv_lora", "cv_prompt"]


def load_case(kind):
    c = cases_cv.tiny_cv_case(kind)
    gold = torch.load(os.path.join(HERE, "golden", "transrec_%s.pt" % kind), weights_only=False)
    sd = cases_cv.build_state_dict(c)
    cfg = O.VitConfig(hidden=c.hidden, layers=c.layers, heads=c.heads, patch=c.patch, eps=c.eps)
    rec = O.RecConfig(max_seq_len=c.S, embedding_dim=c.D, heads=c.rec_heads, blocks=c.blocks)
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
            torch.testing.assert_close(sd[k].grad, gold["grads"][k], rtol=2e-4, atol=2e-6, msg=lambda m: k + ": " + m)
