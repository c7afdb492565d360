"""Generates tests/golden/transrec_cv_*.pt from the UNMODIFIED image-tree reference (imported from
/root/reference/Downstream/CV/model; its data_utils needs lmdb, which is not installed, so only the model package is
used).  Run in the build container:   python tests/golden/make_golden_cv.py
The surgery repeats Downstream/CV/run_adapter.py:383-445 (inline in train() there).  `_LoraLinear` restates loralib
0.1.1 Linear (not installed), including r = 0 (a plain trainable Linear), for the 'cv_lora' case only."""
import math
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/Downstream/CV")

import cases_cv  # noqa: E402


class _LoraLinear(nn.Linear):
    def __init__(self, in_features, out_features, r=0, lora_alpha=1, **kw):
        super().__init__(in_features, out_features, **kw)
        self.r = r
        if r > 0:
            self.scaling = lora_alpha / r
            self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
            self.weight.requires_grad = False
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))

    def forward(self, x):
        y = nn.functional.linear(x, self.weight, self.bias)
        if self.r > 0:
            y = y + (x @ self.lora_A.t() @ self.lora_B.t()) * self.scaling
        return y


sys.modules["loralib"] = types.SimpleNamespace(Linear=_LoraLinear)

from transformers import ViTConfig, ViTForImageClassification  # noqa: E402

from model import (Model, SASRecAdaptedSelfOutput, SASRecCompacterAdaptedSelfOutput, SASRecParallelAdaptedSelfOutput,  # noqa: E402
                   SASRecPfeifferV2AdaptedSelfOutput, SoftPrompt, VITAdaptedOutput, VITAdaptedParallelOutput,
                   VITAdaptedSelfOutput, VITCompacterAdaptedOutput, VITCompacterAdaptedSelfOutput)
from model.layers import PHMLinear  # noqa: E402


class _CompacterModel(nn.Module):
    """Restates CompacterModel of Downstream/CV/run_adapter.py:85-98 (the entry script imports lmdb-backed data_utils at
    module level and cannot be imported here; the class is 12 lines: a shared phm_rule handed to every PHMLinear and a
    pass-through forward)."""

    def __init__(self, args, model):
        super().__init__()
        phm_dim = args.hypercomplex_division
        self.model = model
        self.phm_rule = nn.Parameter(torch.FloatTensor(phm_dim, phm_dim, phm_dim), requires_grad=True)
        self.phm_rule.data.normal_(mean=0, std=args.phm_init_range)
        for name, sub_module in model.named_modules():
            if isinstance(sub_module, PHMLinear):
                sub_module.set_phm_rule(phm_rule=self.phm_rule)

    def forward(self, sample_items, log_mask, local_rank):
        return self.model(sample_items, log_mask, local_rank)


def build_reference_model(c):
    args = cases_cv.reference_args(c)
    cfg = ViTConfig(hidden_size=c.hidden, num_hidden_layers=c.layers, num_attention_heads=c.heads,
                    intermediate_size=c.inter, image_size=c.image, patch_size=c.patch, layer_norm_eps=c.eps)
    cv_model = ViTForImageClassification(cfg)
    cv_model.classifier = nn.Linear(cv_model.classifier.in_features, args.embedding_dim)        # run_adapter.py:293-294
    model = Model(args, 100, True, cv_model)
    for p in model.parameters():
        p.requires_grad = False
    layers = model.cv_encoder.image_net.vit.encoder.layer
    blocks = model.user_encoder.transformer_encoder.transformer_blocks
    if c.kind == "cv_houlsby":
        for lm in layers:
            lm.attention.output = VITAdaptedSelfOutput(lm.attention.output, args)
            lm.output = VITAdaptedOutput(lm.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecAdaptedSelfOutput(tb, args)
    elif c.kind == "cv_parallel":                                  # run_adapter.py:236-247 (is_serial == "None")
        for lm in layers:
            lm.output = VITAdaptedParallelOutput(lm.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecParallelAdaptedSelfOutput(tb, args)
    elif c.kind == "cv_pfeiffer_ver2":                             # run_adapter.py:367-377
        for lm in layers:
            lm.attention.output = VITAdaptedSelfOutput(lm.attention.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecPfeifferV2AdaptedSelfOutput(tb, args)
    elif c.kind == "cv_compacter":                                 # run_adapter.py:396-411
        for lm in layers:
            lm.attention.output = VITCompacterAdaptedSelfOutput(lm.attention.output, args)
            lm.output = VITCompacterAdaptedOutput(lm.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecCompacterAdaptedSelfOutput(tb, args)
        model = _CompacterModel(args, model)
    elif c.kind == "cv_lora":
        import loralib as lora
        for lm in layers:
            lm.attention.attention.query = lora.Linear(768, 768, r=12)
            lm.attention.attention.value = lora.Linear(768, 768, r=12)
        for i in range(len(blocks)):
            blocks[i].multi_head_attention.w_Q = lora.Linear(args.embedding_dim, args.embedding_dim, r=4)
            blocks[i].multi_head_attention.w_V = lora.Linear(args.embedding_dim, args.embedding_dim)
    elif c.kind == "cv_full_ft":                                   # fine_tune_to = all: nothing frozen
        for p in model.parameters():
            p.requires_grad = True
    elif c.kind == "cv_prompt":
        s_wte = SoftPrompt(cv_model.vit.embeddings, n_tokens=args.n_tokens, embed_dim=768)
        model.cv_encoder.image_net.vit.embeddings = s_wte
        for name, param in model.named_parameters():
            if "cv_encoder.image_net.classifier" in name:
                param.requires_grad = True
    return model, args


def main():
    import transformers
    meta = {"torch": torch.__version__, "transformers": transformers.__version__}
    for kind in (sys.argv[1:] or cases_cv.CV_ALL_KINDS):
        c = cases_cv.tiny_cv_case(kind)
        model, args = build_reference_model(c)
        sd = cases_cv.build_state_dict(c)
        ref_keys = set(model.state_dict().keys())
        assert ref_keys == set(sd.keys()), (kind, sorted(ref_keys - set(sd.keys()))[:5], sorted(set(sd.keys()) - ref_keys)[:5])
        model.load_state_dict(sd)
        train_keys = cases_cv.trainable_keys(c, sd)
        got = sorted(n for n, p in model.named_parameters() if p.requires_grad)
        # named_parameters() de-duplicates the shared patch projection; compare on the de-duplicated set
        assert got == sorted(set(train_keys)), (kind, got[:6], sorted(train_keys)[:6])
        model.eval()
        images, log_mask = cases_cv.build_batch(c)
        loss = model(images, log_mask, "cpu")
        out = {"meta": meta, "kind": kind, "loss": loss.detach().clone()}
        if train_keys:
            loss.backward()
            out["grads"] = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
        with torch.no_grad():
            out["item_emb"] = (model.model if kind == "cv_compacter" else model).cv_encoder(images).clone()
        path = os.path.join(HERE, "transrec_%s.pt" % kind)
        torch.save(out, path)
        print(kind, "tokens", c.P + 1 + c.n_tokens, "loss %.6f" % float(loss), "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
