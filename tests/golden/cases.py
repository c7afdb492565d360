"""Seeded weights and inputs shared by tests/golden/make_golden.py (which feeds them to the UNMODIFIED reference
modules) and by the tests (which feed the same tensors to the oracle and to the CUDA path).

State-dict keys are the reference's.  Everything is generated from torch CPU generators, so
the build container and the GPU box (same image) produce bit-identical tensors."""
import types

import torch

BERT = "bert_encoder.text_encoders.title.bert_model."
FC = "bert_encoder.text_encoders.title.fc."
USER = "user_encoder.transformer_encoder."


ZOO_KINDS = ("parallel", "pfeiffer", "pfeiffer_leaky", "pfeiffer_ver2", "compacter", "kadapter")
ALL_KINDS = ("base", "houlsby", "houlsby_gelu", "lora", "prompt_cpc") + ZOO_KINDS + ("full_ft", "full_ft_roberta")


def _houlsby_like(kind):
    """kinds whose BERT/SASRec wrappers carry an AdapterBlock bottleneck of rank 16"""
    return kind.startswith("houlsby") or kind in ZOO_KINDS


def tiny_case(kind):
    """kind: 'base' | 'houlsby' | 'lora' | 'prompt_cpc' (RoBERTa + SoftEmbedding + ModelCPC) | 'houlsby_gelu' |
    'parallel' (Houlsby, is_serial=None) | 'pfeiffer' (GELU) | 'pfeiffer_leaky' | 'pfeiffer_ver2'."""
    c = types.SimpleNamespace()
    c.kind = kind
    c.roberta = kind in ("prompt_cpc", "full_ft_roberta")
    c.cpc = kind == "prompt_cpc"
    c.hidden, c.layers, c.heads, c.inter = 128, 2, 2, 512
    c.vocab, c.max_pos = 200, 40
    c.eps = 1e-5 if c.roberta else 1e-12
    c.pad = 1 if c.roberta else 0
    c.L = 8                    # num_words_title
    c.S = 5                    # max_seq_len
    c.D = 64                   # embedding_dim
    c.rec_heads, c.blocks = 2, 2
    c.bert_r = 16 if _houlsby_like(kind) else 8   # bert_adapter_down_size (LoRA rank for kind == 'lora')
    c.rec_r = 16 if _houlsby_like(kind) else 8    # adapter_down_size
    c.n_tokens = 3 if kind == "prompt_cpc" else 0
    c.activation = {"houlsby_gelu": "GELU", "pfeiffer": "GELU", "pfeiffer_leaky": "leaky_relu"}.get(kind, "RELU")
    c.parallel = kind == "parallel"
    c.B = 6
    c.item_num = 40
    c.seed = {"base": 11, "houlsby": 12, "lora": 13, "prompt_cpc": 14, "houlsby_gelu": 15, "parallel": 16,
              "pfeiffer": 17, "pfeiffer_leaky": 18, "pfeiffer_ver2": 19, "compacter": 20, "kadapter": 21, "full_ft": 22, "full_ft_roberta": 23}[kind]
    # K-Adapter (parameters.py:68-71): adapters after BERT layers 0 and 1 of the 2-layer tiny body, width 64 with
    # 2 heads (head width 32); the SASRec-side adapters keep the reference default width 16 with 2 heads (head width 8)
    c.k_list, c.k_hidden, c.k_heads_bert, c.k_heads_rec = "0,1", 64, 2, 2
    c.phm_dim = 4              # hypercomplex_division
    return c


def full_case(kind):
    """The same generators at the REAL sizes of BASELINE.json's text configurations: BERT-base body (768 / 12 layers /
    12 heads / 3072, vocabulary 30,522), 30 title tokens, max_seq_len 20, embedding_dim 64, adapter ranks 64 (BERT) and
    16 (SASRec) as parameters.py:55,62 default them (LoRA: r = 8), 2 users = 84 item sequences = 2,520 tokens."""
    c = tiny_case(kind)
    c.hidden, c.layers, c.heads, c.inter = 768, 12, 12, 3072
    c.vocab, c.max_pos = 30522, 512
    if c.roberta:              # RoBERTa-base (BASELINE.json configs[3], "C4"): vocabulary 50,265, 514 positions
        c.vocab, c.max_pos = 50265, 514
    c.L, c.S, c.D = 30, 20, 64
    c.bert_r = 8 if kind == "lora" else 64
    c.rec_r = 8 if kind == "lora" else 16
    if kind == "prompt_cpc":
        c.n_tokens = 10        # SoftEmbedding with n_tokens = 10
    c.B, c.item_num = 2, 60
    c.seed += 100
    return c


def reference_args(c):
    """The argparse namespace the reference modules read (Downstream/Text/parameters.py)."""
    return types.SimpleNamespace(
        max_seq_len=c.S, min_seq_len=2, l2_weight=0, embedding_dim=c.D, num_attention_heads=c.rec_heads,
        drop_rate=0.1, transformer_block=c.blocks, num_words_title=c.L, num_words_abstract=50, num_words_body=50,
        news_attributes=["title"], word_embedding_dim=c.hidden,
        bert_model_load={128: "bert_tiny", 768: "bert_base_uncased"}[c.hidden],
        bert_adapter_down_size=c.bert_r, adapter_down_size=c.rec_r, adapter_dropout_rate=0.1,
        adapter_activation=c.activation, num_workers=0, adapter_type={"houlsby": "houslby", "houlsby_gelu": "houslby",
                                                                      "lora": "lora", "prompt_cpc": "prompt",
                                                                      "base": "None", "parallel": "houslby",
                                                                      "pfeiffer": "pfeiffer", "pfeiffer_leaky": "pfeiffer",
                                                                      "pfeiffer_ver2": "pfeiffer_ver2", "compacter": "compacter",
                                                                      "kadapter": "kadapter", "full_ft": "None",
                                                                      "full_ft_roberta": "None"}[c.kind],
        k_adapter_bert_list=c.k_list, k_adapter_bert_hidden_dim=c.k_hidden, num_adapter_heads_bert=c.k_heads_bert,
        num_adapter_heads_sasrec=c.k_heads_rec,
        hypercomplex_division=c.phm_dim, phm_init_range=0.0001,
        is_serial="None" if c.kind == "parallel" else "True", n_tokens=c.n_tokens)


def _n(g, shape, std):
    return torch.randn(shape, generator=g) * std


def _linear(sd, g, name, out_f, in_f, bias=True, std=0.05):
    sd[name + "weight"] = _n(g, (out_f, in_f), std)
    if bias:
        sd[name + "bias"] = _n(g, (out_f,), 0.05)


def _ln(sd, g, name, dim):
    sd[name + "weight"] = 1.0 + _n(g, (dim,), 0.1)
    sd[name + "bias"] = _n(g, (dim,), 0.1)


def _adapter(sd, g, name, dim, r):
    _linear(sd, g, name + "fc_down.", r, dim)
    _linear(sd, g, name + "fc_up.", dim, r)


def _phm(sd, g, name, in_f, out_f, n, rule):
    sd[name + "W_left"] = _n(g, (n, in_f // n, 1), 0.3)
    sd[name + "W_right"] = _n(g, (n, 1, out_f // n), 0.3)
    sd[name + "b"] = _n(g, (out_f,), 0.05)
    sd[name + "phm_rule"] = rule            # the ONE shared tensor, repeated under every PHMLinear by the reference


def _phm_adapter(sd, g, name, dim, r, n, rule):
    _phm(sd, g, name + "down_sampler.", dim, r, n, rule)
    _phm(sd, g, name + "up_sampler.", r, dim, n, rule)


def _tblock(sd, g, tb, D):
    """a plain TransformerBlock (modules.py:77-87) of width D"""
    for nm in ("w_Q", "w_K", "w_V", "fc"):
        _linear(sd, g, tb + "multi_head_attention.%s." % nm, D, D, bias=False, std=0.1)
    _ln(sd, g, tb + "multi_head_attention.layer_norm.", D)
    _linear(sd, g, tb + "feed_forward.w_1.", 4 * D, D, std=0.1)
    _linear(sd, g, tb + "feed_forward.w_2.", D, 4 * D, std=0.1)
    _ln(sd, g, tb + "feed_forward.layer_norm.", D)


def _kadapter(sd, g, name, dim, r):
    _linear(sd, g, name + "down_project.", r, dim, std=0.1)
    _linear(sd, g, name + "up_project.", dim, r, std=0.1)
    for j in range(2):
        _tblock(sd, g, name + "transformer_blocks.%d." % j, r)


def _lora(sd, g, name, dim, r):
    _linear(sd, g, name, dim, dim, bias=True)
    sd[name + "lora_A"] = _n(g, (r, dim), 0.1)
    sd[name + "lora_B"] = _n(g, (dim, r), 0.1)   # loralib initialises B to zero; non-zero here so the path is exercised


def build_state_dict(c):
    """Full model state dict (after adapter surgery) with the reference's key names."""
    g = torch.Generator().manual_seed(c.seed)
    sd = {}
    H = c.hidden
    rule = _n(g, (c.phm_dim,) * 3, 0.5) if c.kind == "compacter" else None
    e = BERT + "embeddings."
    if c.kind == "prompt_cpc":
        sd[e + "word_embeddings.wte.weight"] = _n(g, (c.vocab, H), 0.05)
        sd[e + "word_embeddings.learned_embedding"] = _n(g, (c.n_tokens, H), 0.05)
    else:
        sd[e + "word_embeddings.weight"] = _n(g, (c.vocab, H), 0.05)
    sd[e + "position_embeddings.weight"] = _n(g, (c.max_pos, H), 0.05)
    sd[e + "token_type_embeddings.weight"] = _n(g, (1 if c.roberta else 2, H), 0.05)
    _ln(sd, g, e + "LayerNorm.", H)
    for i in range(c.layers):
        p = BERT + "encoder.layer.%d." % i
        for nm in ("query", "key", "value"):
            if c.kind == "lora" and nm != "key":
                _lora(sd, g, p + "attention.self.%s." % nm, H, c.bert_r)
            else:
                _linear(sd, g, p + "attention.self.%s." % nm, H, H)
        for out_name, in_f in (("attention.output.", H), ("output.", c.inter)):
            # which of the two sub-layer outputs is wrapped: Houlsby serial/parallel both (run.py:456-460,468-474),
            # pfeiffer_ver2 attention.output only (:391-394), pfeiffer output only (:403-406)
            wrapped = (c.kind.startswith("houlsby") or c.kind in ("parallel", "compacter")
                       or (c.kind == "pfeiffer_ver2" and out_name == "attention.output.")
                       or (c.kind.startswith("pfeiffer") and c.kind != "pfeiffer_ver2" and out_name == "output."))
            if wrapped:
                _linear(sd, g, p + out_name + "self_output.dense.", H, in_f)
                _ln(sd, g, p + out_name + "self_output.LayerNorm.", H)
                if c.kind == "compacter":
                    _phm_adapter(sd, g, p + out_name + "adapter.", H, c.bert_r, c.phm_dim, rule)
                else:
                    _adapter(sd, g, p + out_name + "adapter.", H, c.bert_r)
                if c.kind in ("pfeiffer", "pfeiffer_leaky"):
                    _ln(sd, g, p + out_name + "LN.", H)
            else:
                _linear(sd, g, p + out_name + "dense.", H, in_f)
                _ln(sd, g, p + out_name + "LayerNorm.", H)
            if out_name == "attention.output.":
                _linear(sd, g, p + "intermediate.dense.", c.inter, H)
    _linear(sd, g, BERT + "pooler.dense.", H, H)
    _linear(sd, g, FC, c.D, H)
    D = c.D
    sd[USER + "position_embedding.weight"] = _n(g, (c.S, D), 0.1)
    _ln(sd, g, USER + "layer_norm.", D)
    for j in range(c.blocks):
        p = USER + "transformer_blocks.%d." % j
        tb = p + ("transformer_block." if (_houlsby_like(c.kind) and c.kind != "kadapter") else "")
        for nm in ("w_Q", "w_K", "w_V", "fc"):
            if c.kind == "lora" and nm in ("w_Q", "w_V"):
                _lora(sd, g, tb + "multi_head_attention.%s." % nm, D, c.rec_r)
            else:
                _linear(sd, g, tb + "multi_head_attention.%s." % nm, D, D, bias=False, std=0.1)
        _ln(sd, g, tb + "multi_head_attention.layer_norm.", D)
        _linear(sd, g, tb + "feed_forward.w_1.", 4 * D, D, std=0.1)
        _linear(sd, g, tb + "feed_forward.w_2.", D, 4 * D, std=0.1)
        _ln(sd, g, tb + "feed_forward.layer_norm.", D)
        if c.kind.startswith("houlsby") or c.kind == "parallel":
            _adapter(sd, g, p + "adapter1.", D, c.rec_r)
            _adapter(sd, g, p + "adapter2.", D, c.rec_r)
        elif c.kind == "pfeiffer_ver2":
            _adapter(sd, g, p + "adapter1.", D, c.rec_r)
        elif c.kind in ("pfeiffer", "pfeiffer_leaky"):
            _adapter(sd, g, p + "adapter.", D, c.rec_r)
            _ln(sd, g, p + "LN.", D)
        elif c.kind == "compacter":
            _phm_adapter(sd, g, p + "adapter1.", D, c.rec_r, c.phm_dim, rule)
            _phm_adapter(sd, g, p + "adapter2.", D, c.rec_r, c.phm_dim, rule)
    if c.kind == "kadapter":
        # BertKAdaptedBertModel replaces title.bert_model (body moves one level down); SASRecKAdaptedTransformerBlocks
        # replaces transformer_encoder.transformer_blocks (run.py:409-413)
        inner, blocks = BERT + "bert_model.", USER + "transformer_blocks."
        sd = {(inner + k[len(BERT):] if k.startswith(BERT) else
               blocks + "transformer_blocks." + k[len(blocks):] if k.startswith(blocks) else k): v for k, v in sd.items()}
        for i in range(len(c.k_list.split(","))):
            _kadapter(sd, g, BERT + "bert_adapter_list.%d." % i, H, c.k_hidden)
        _linear(sd, g, BERT + "com_dense.", H, 2 * H)
        for j in range(c.blocks):
            _kadapter(sd, g, blocks + "adapter_list.%d." % j, D, c.rec_r)
        _linear(sd, g, blocks + "com_dense2.", D, 2 * D, std=0.1)
    if c.kind == "compacter":       # CompacterModel wraps the model as `.model` and owns the shared rule (run.py:70-81)
        sd = {"model." + k: v for k, v in sd.items()}
        sd["phm_rule"] = rule
    return sd


def trainable_keys(c, sd):
    """Parameters left trainable by Downstream/Text/run.py:367-479 with fine_tune_to=None."""
    if c.kind.startswith("houlsby") or c.kind in ("parallel", "pfeiffer_ver2"):
        return [k for k in sd if "adapter" in k]
    if c.kind.startswith("full_ft"):
        # fine_tune_to = all with the pooler frozen, as the source-domain stage trains (Pretraining/Text/run.py:48-64)
        return [k for k in sd if "pooler" not in k]
    if c.kind == "kadapter":
        return [k for k in sd if "adapter_list" in k or "com_dense" in k]
    if c.kind == "compacter":     # named_parameters() lists the shared rule once, under the wrapper's own name
        return [k for k in sd if "adapter" in k and not k.endswith("phm_rule")] + ["phm_rule"]
    if c.kind in ("pfeiffer", "pfeiffer_leaky"):      # the new `LN` LayerNorms are created after the freeze
        return [k for k in sd if "adapter" in k or ".LN." in k]
    if c.kind == "lora":
        return [k for k in sd if "lora_" in k or (k.endswith("bias") and (
            ".query." in k or ".value." in k or ".w_Q." in k or ".w_V." in k))]
    if c.kind == "prompt_cpc":
        return [k for k in sd if k.endswith("learned_embedding")]
    return []


def build_item_content(c):
    """[item_num+1, 2L] int64 rows of ids | attention mask, as get_doc_input_bert builds them
    (Downstream/Text/data_utils/preprocess.py:114-151): row 0 all zeros; BERT: [CLS]=101->here 2, [SEP]->3, pad 0;
    RoBERTa: <s>=0, </s>=2, pad id 1 with mask 0."""
    g = torch.Generator().manual_seed(c.seed + 1000)
    rows = torch.zeros((c.item_num + 1, 2 * c.L), dtype=torch.long)
    for i in range(1, c.item_num + 1):
        n = int(torch.randint(3, c.L + 1, (1,), generator=g))
        ids = torch.randint(5, c.vocab, (n,), generator=g)
        if c.roberta:
            ids[0], ids[-1] = 0, 2
            rows[i, :c.L] = c.pad
        else:
            ids[0], ids[-1] = 2, 3
        rows[i, :n] = ids
        rows[i, c.L:c.L + n] = 1
    return rows


def build_batch(c, item_content):
    """A training batch as BuildTrainDataset.__getitem__ returns it (Downstream/Text/data_utils/dataset.py:24-49):
    sample_items [B, S+1, 2, 2L] int64, log_mask [B, S] float32 (left padded).  Also returns the id-level tensor."""
    g = torch.Generator().manual_seed(c.seed + 2000)
    S1 = c.S + 1
    ids = torch.zeros((c.B, S1, 2), dtype=torch.long)
    log_mask = torch.zeros((c.B, c.S))
    for b in range(c.B):
        n = int(torch.randint(2, S1 + 1, (1,), generator=g))   # sequence length incl. the last target
        seq = (torch.randperm(c.item_num, generator=g)[:n] + 1)
        ids[b, S1 - n:, 0] = seq
        neg = torch.randint(1, c.item_num + 1, (n - 1,), generator=g)
        ids[b, S1 - n:S1 - 1, 1] = neg
        log_mask[b, c.S - (n - 1):] = 1.0
    return item_content[ids], log_mask, ids


def build_eval_users(c):
    """Eval sequences (history + target) and the ids to mask, as read_behaviors builds them
    (Downstream/Text/data_utils/preprocess.py:51-59)."""
    g = torch.Generator().manual_seed(c.seed + 3000)
    seqs, hist = [], []
    for _ in range(9):
        n = int(torch.randint(2, c.S + 2, (1,), generator=g))
        seq = (torch.randperm(c.item_num, generator=g)[:n] + 1).tolist()
        seqs.append(seq)
        hist.append(seq[:-1])
    return seqs, hist
