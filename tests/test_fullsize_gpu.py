"""-m gpu: the train step at the REAL model sizes of BASELINE.json (BERT-base body, 30 tokens, S = 20, D = 64; Houlsby r = 64 /
16 and LoRA r = 8) against the fp32 CPU oracle on the same seeded weights — the tiny golden cases pin the oracle to the
reference, this case shows the kernels agree with it at the shapes the benchmark runs (12 layers deep, head width 64,
H = 768 fused adapter kernel, 768 / 2304 / 3072-wide GEMMs).  2 users = 2,520 tokens keep the oracle at a few seconds."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import cases  # noqa: E402,F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,unpad", [("houlsby", False), ("lora", False), ("lora", True), ("prompt_cpc", False)])
def test_full_size_step_matches_oracle(kind, unpad):
    """BERT-base Houlsby / LoRA (C1, C2) and RoBERTa-base + SoftEmbedding(n_tokens = 10) + ModelCPC (C4,
    Downstream/Text/model/model.py:113-135,586-630)."""
    import parity_util as P
    fig, _ = P.text_case(kind, full=True, unpad=unpad)
    print("full-size %s unpad=%s: loss %.5f vs %.5f, grad rel L2 %.4f cos %.5f, emb rel L2 %.4f max abs %.4f"
          % (kind, unpad, fig["loss"], fig["oracle_loss"], fig["grad_all_rel"], fig["grad_all_cos"], fig["emb_rel_l2"],
             fig["emb_max_abs"]))
    P.check_against_table(fig, P.measured(), "text_full/%s%s" % (kind, "/unpad" if unpad else ""), True)


@pytest.mark.parametrize("kind", ["cv_houlsby"])
def test_full_size_vit_step_matches_oracle(kind):
    """ViT-B/16-224 (12 layers, 197 tokens per image, r = 64 Houlsby wrappers on both sub-layer outputs; C3,
    Downstream/CV/model/model.py:54-77,182-212, encoders.py:31-32): 2 users = 24 images through the mid-length attention
    kernel, the fused adapter kernel's ViT tails and the patch-projection GEMM."""
    import parity_util as P
    fig, _ = P.cv_case(kind, full=True)
    print("full-size %s: loss %.5f vs %.5f, grad rel L2 %.4f cos %.5f, emb rel L2 %.4f max abs %.4f"
          % (kind, fig["loss"], fig["oracle_loss"], fig["grad_all_rel"], fig["grad_all_cos"], fig["emb_rel_l2"], fig["emb_max_abs"]))
    P.check_against_table(fig, P.measured(), "cv_full/" + kind, True)
