"""Seeded weights / inputs for the image tree (Downstream/CV), shared by make_golden_cv.py and the tests."""
import types

import torch

from cases import USER, _adapter, _linear, _ln, _lora, _n, _phm_adapter

VIT = "cv_encoder.image_net.vit."
CLS = "cv_encoder.image_net.classifier."


CV_ZOO_KINDS = ("cv_parallel", "cv_pfeiffer_ver2", "cv_compacter")          # SURVEY.md §8f-4 on the image tree
CV_ALL_KINDS = ("cv_base", "cv_houlsby", "cv_lora", "cv_prompt") + CV_ZOO_KINDS + ("cv_full_ft",)


def tiny_cv_case(kind):
    """kind: 'cv_base' | 'cv_houlsby' | 'cv_lora' | 'cv_prompt' | 'cv_parallel' (Houlsby, is_serial=None: layer.output
    only) | 'cv_pfeiffer_ver2' (attention.output only) | 'cv_compacter' | 'cv_full_ft' (nothing frozen: the source-domain
    stage Pretraining/CV/run.py and `fine_tune_to=all`, incl. the patch projection, cls token and position embeddings).
    hidden is 768 because the reference's ViT adapter
    wrappers hard-code it (Downstream/CV/model/model.py:186,202); depth / MLP width / image size are reduced."""
    c = types.SimpleNamespace()
    c.kind = kind
    c.hidden, c.heads, c.layers, c.inter = 768, 12, 2, 256
    if kind == "cv_full_ft":        # no adapter wrappers => no hard-coded 768; a narrow body keeps the golden at 1.7 MB
        c.hidden, c.heads = 128, 2
    c.patch = 16
    c.image = 96 if kind in ("cv_lora", "cv_prompt") else 64      # 37 tokens (mid-length kernel) / 17 tokens (short kernel)
    c.P = (c.image // c.patch) ** 2
    c.eps = 1e-12
    c.S, c.D, c.rec_heads, c.blocks = 4, 64, 2, 2
    c.cv_r, c.rec_r = 16, 8
    c.phm_dim = 4
    c.parallel = kind == "cv_parallel"
    c.n_tokens = 3 if kind == "cv_prompt" else 0
    c.cpc = False
    c.B = 3
    c.seed = {"cv_base": 21, "cv_houlsby": 22, "cv_lora": 23, "cv_prompt": 24, "cv_parallel": 25, "cv_pfeiffer_ver2": 26,
              "cv_compacter": 27, "cv_full_ft": 28}[kind]
    return c


def full_cv_case(kind):
    """The same generators at the REAL sizes of BASELINE.json configs[2] ("C3"): ViT-B/16-224 (768 / 12 layers / 12 heads /
    3072, 196 patches + cls = 197 tokens), adapter rank 64 (Downstream/CV/parameters.py), 2 users with max_seq_len 5 =
    24 images = 4,728 tokens (the oracle's forward + backward stays under a minute and a few GB)."""
    c = tiny_cv_case(kind)
    c.hidden, c.heads, c.layers, c.inter = 768, 12, 12, 3072
    c.image = 224
    c.P = (c.image // c.patch) ** 2
    c.S, c.B = 5, 2
    c.cv_r, c.rec_r = 64, 16
    if kind == "cv_prompt":
        c.n_tokens = 10
    c.seed += 100
    return c


def reference_args(c):
    return types.SimpleNamespace(
        max_seq_len=c.S, min_seq_len=2, l2_weight=0, embedding_dim=c.D, num_attention_heads=c.rec_heads, drop_rate=0.1,
        transformer_block=c.blocks, CV_model_load="vit-base-patch16-224", cv_adapter_down_size=c.cv_r,
        adapter_down_size=c.rec_r, adapter_dropout_rate=0.1, adapter_activation="RELU", n_tokens=c.n_tokens,
        adapter_type={"cv_base": "None", "cv_houlsby": "houslby", "cv_lora": "lora", "cv_prompt": "prompt",
                      "cv_parallel": "houslby", "cv_pfeiffer_ver2": "pfeiffer_ver2", "cv_compacter": "compacter",
                      "cv_full_ft": "None"}[c.kind],
        adding_adapter_to="all", is_serial="None" if c.kind == "cv_parallel" else "True", finetune_layernorm="None",
        hypercomplex_division=c.phm_dim, phm_init_range=0.0001)


def build_state_dict(c):
    g = torch.Generator().manual_seed(c.seed)
    sd = {}
    H, D = c.hidden, c.D
    rule = _n(g, (c.phm_dim,) * 3, 0.5) if c.kind == "cv_compacter" else None
    e = VIT + "embeddings."
    names = [e]
    if c.kind == "cv_prompt":
        names = [e + "wte."]
    for pre in names:
        sd[pre + "cls_token"] = _n(g, (1, 1, H), 0.05)
        sd[pre + "position_embeddings"] = _n(g, (1, c.P + 1, H), 0.05)
        sd[pre + "patch_embeddings.projection.weight"] = _n(g, (H, 3, c.patch, c.patch), 0.03)
        sd[pre + "patch_embeddings.projection.bias"] = _n(g, (H,), 0.05)
    if c.kind == "cv_prompt":   # SoftPrompt registers wte.patch_embeddings a second time as .patch_embeddings
        sd[e + "patch_embeddings.projection.weight"] = sd[e + "wte.patch_embeddings.projection.weight"]
        sd[e + "patch_embeddings.projection.bias"] = sd[e + "wte.patch_embeddings.projection.bias"]
        sd[e + "Prompt_Tokens"] = _n(g, (1, c.n_tokens, H), 0.05)
    for i in range(c.layers):
        p = VIT + "encoder.layer.%d." % i
        for nm in ("query", "key", "value"):
            if c.kind == "cv_lora" and nm != "key":
                _lora(sd, g, p + "attention.attention.%s." % nm, H, 12)      # run_adapter.py:387-388: r = 12
            else:
                _linear(sd, g, p + "attention.attention.%s." % nm, H, H, std=0.03)
        for out_name, in_f in (("attention.output.", H), ("output.", c.inter)):
            wrapped = (c.kind in ("cv_houlsby", "cv_compacter")
                       or (c.kind == "cv_parallel" and out_name == "output.")                  # run_adapter.py:240-241
                       or (c.kind == "cv_pfeiffer_ver2" and out_name == "attention.output."))  # run_adapter.py:369-372
            if wrapped:
                _linear(sd, g, p + out_name + "self_output.dense.", H, in_f, std=0.03)
                if c.kind == "cv_compacter":
                    _phm_adapter(sd, g, p + out_name + "adapter.", H, c.cv_r, c.phm_dim, rule)
                else:
                    _adapter(sd, g, p + out_name + "adapter.", H, c.cv_r)
            else:
                _linear(sd, g, p + out_name + "dense.", H, in_f, std=0.03)
            if out_name == "attention.output.":
                _linear(sd, g, p + "intermediate.dense.", c.inter, H, std=0.03)
        _ln(sd, g, p + "layernorm_before.", H)
        _ln(sd, g, p + "layernorm_after.", H)
    _ln(sd, g, VIT + "layernorm.", H)
    _linear(sd, g, CLS, D, H, std=0.05)
    sd[USER + "position_embedding.weight"] = _n(g, (c.S, D), 0.1)
    _ln(sd, g, USER + "layer_norm.", D)
    for j in range(c.blocks):
        p = USER + "transformer_blocks.%d." % j
        tb = p + ("transformer_block." if c.kind in ("cv_houlsby",) + CV_ZOO_KINDS else "")
        for nm in ("w_Q", "w_K", "w_V", "fc"):
            if c.kind == "cv_lora" and nm == "w_Q":
                _lora(sd, g, tb + "multi_head_attention.w_Q.", D, 4)          # run_adapter.py:392-393: r = 4
            elif c.kind == "cv_lora" and nm == "w_V":
                _linear(sd, g, tb + "multi_head_attention.w_V.", D, D, bias=True, std=0.1)   # lora.Linear(D, D): r = 0
            else:
                _linear(sd, g, tb + "multi_head_attention.%s." % nm, D, D, bias=False, std=0.1)
        _ln(sd, g, tb + "multi_head_attention.layer_norm.", D)
        _linear(sd, g, tb + "feed_forward.w_1.", 4 * D, D, std=0.1)
        _linear(sd, g, tb + "feed_forward.w_2.", D, 4 * D, std=0.1)
        _ln(sd, g, tb + "feed_forward.layer_norm.", D)
        if c.kind in ("cv_houlsby", "cv_parallel"):
            _adapter(sd, g, p + "adapter1.", D, c.rec_r)
            _adapter(sd, g, p + "adapter2.", D, c.rec_r)
        elif c.kind == "cv_pfeiffer_ver2":
            _adapter(sd, g, p + "adapter1.", D, c.rec_r)
        elif c.kind == "cv_compacter":
            _phm_adapter(sd, g, p + "adapter1.", D, c.rec_r, c.phm_dim, rule)
            _phm_adapter(sd, g, p + "adapter2.", D, c.rec_r, c.phm_dim, rule)
    if c.kind == "cv_compacter":      # CompacterModel (run_adapter.py:85-98) wraps the model and owns the shared rule
        sd = {"model." + k: v for k, v in sd.items()}
        sd["phm_rule"] = rule
    return sd


def trainable_keys(c, sd):
    if c.kind in ("cv_houlsby", "cv_parallel", "cv_pfeiffer_ver2"):
        return [k for k in sd if "adapter" in k]
    if c.kind == "cv_compacter":
        return [k for k in sd if "adapter" in k and not k.endswith("phm_rule")] + ["phm_rule"]
    if c.kind == "cv_lora":
        return [k for k in sd if "lora_" in k or (".query.bias" in k or ".value.bias" in k or ".w_Q.bias" in k)
                or ".w_V." in k]
    if c.kind == "cv_prompt":
        return [k for k in sd if k.endswith("Prompt_Tokens") or k.startswith(CLS)]
    if c.kind == "cv_full_ft":
        return list(sd)
    return []


def build_batch(c):
    """images as Build_Lmdb_Dataset returns them after Normalize(0.5, 0.5): values in (-1, 1)
    (Downstream/CV/data_utils/dataset.py:75-80), [B*(S+1)*2, 3, R, R]; left-padded log_mask [B, S]."""
    g = torch.Generator().manual_seed(c.seed + 2000)
    n = c.B * (c.S + 1) * 2
    images = torch.rand((n, 3, c.image, c.image), generator=g) * 2 - 1
    log_mask = torch.zeros((c.B, c.S))
    for b in range(c.B):
        ln = int(torch.randint(1, c.S + 1, (1,), generator=g))
        log_mask[b, c.S - ln:] = 1.0
    return images, log_mask
