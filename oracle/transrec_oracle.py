"""CPU oracle for the TransRec train + full-ranking-eval hot path of westlake-repl/Adapter4Rec.

TEST INFRASTRUCTURE ONLY.  Nothing under adapter4rec_b200/ imports this module; it may be imported by tests/,
by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, and only as the checker or the
timed CPU baseline — never as the product path.

What it is: a plain fp32 restatement, in functional PyTorch-on-CPU form, of the arithmetic the reference executes
for this path.  The reference itself is PyTorch, and the transformer bodies live in `transformers` (pinned 4.20.1 in
/root/reference/README.md:64; installed here: 5.5.0), so the restatement follows the reference's own call sites and
module code (cited per function, paths relative to /root/reference) and the published BERT/RoBERTa layer algebra.
loralib (pinned 0.1.1, README.md:65) is NOT installed: `lora_linear` restates its published Linear.forward.

Pinned: tests/golden/*.pt hold outputs of the UNMODIFIED reference modules (imported from /root/reference by
tests/golden/make_golden.py in the build container, with transformers' own BertModel/RobertaModel bodies) on seeded
weights and inputs; tests/test_oracle.py checks every function here against them.

Every function takes a `state_dict`-style mapping whose KEYS ARE THE REFERENCE'S, so the same
checkpoint dictionary drives the oracle and the CUDA path.  All functions are differentiable (torch autograd) so
gradient parity of the trainable adapter / LoRA / prompt parameters can be checked too.
"""
import math

import torch
import torch.nn.functional as F

BERT_PREFIX = "bert_encoder.text_encoders.title.bert_model."
FC_PREFIX = "bert_encoder.text_encoders.title.fc."
USER_PREFIX = "user_encoder.transformer_encoder."


class TextConfig:
    """The few numbers of the transformers config the path depends on."""

    def __init__(self, hidden=768, layers=12, heads=12, eps=1e-12, roberta=False, pad_token_id=0):
        self.hidden, self.layers, self.heads, self.eps = hidden, layers, heads, eps
        self.roberta, self.pad_token_id = roberta, pad_token_id


class RecConfig:
    """argparse fields read by the path (Downstream/Text/parameters.py:25-31,55,62,65,76)."""

    def __init__(self, max_seq_len=20, embedding_dim=64, heads=2, blocks=2, num_words_title=30,
                 adapter_activation="RELU", n_tokens=0, parallel=False, k_adapter_bert_list="0,11",
                 num_adapter_heads_bert=12, num_adapter_heads_sasrec=2):
        self.max_seq_len, self.embedding_dim, self.heads, self.blocks = max_seq_len, embedding_dim, heads, blocks
        self.num_words_title, self.adapter_activation, self.n_tokens = num_words_title, adapter_activation, n_tokens
        # is_serial == "None" (parameters.py:66, run.py:454/466): the parallel Houlsby wrappers have the SAME state_dict
        # keys as the serial ones, so the variant cannot be read off the checkpoint
        self.parallel = parallel
        # K-Adapter (parameters.py:68-71): which BERT layers feed adapters, and the adapters' own head counts
        self.k_adapter_bert_list = k_adapter_bert_list
        self.num_adapter_heads_bert, self.num_adapter_heads_sasrec = num_adapter_heads_bert, num_adapter_heads_sasrec


def layer_norm(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def adapter_block(x, sd, prefix, activation="RELU"):
    """AdapterBlock.forward, Downstream/Text/model/modules.py:131-134: fc_up(act(fc_down(x))) + x.
    (self.dropout is constructed at :129 but never applied.)  act = GELU iff args.adapter_activation == "GELU"
    (modules.py:122-125)."""
    h = F.linear(x, sd[prefix + "fc_down.weight"], sd[prefix + "fc_down.bias"])
    h = F.gelu(h) if activation == "GELU" else F.relu(h)
    return F.linear(h, sd[prefix + "fc_up.weight"], sd[prefix + "fc_up.bias"]) + x


def adapter_pfeiffer_block(x, sd, prefix, activation):
    """AdapterPfeifferBlock.forward, Downstream/Text/model/modules.py:155-158: fc_up(act(fc_down(x))), no residual;
    act in {GELU, leaky_relu (slope 0.01), relu} (modules.py:144-149)."""
    h = F.linear(x, sd[prefix + "fc_down.weight"], sd[prefix + "fc_down.bias"])
    h = {"GELU": F.gelu, "leaky_relu": F.leaky_relu, "relu": F.relu}[activation](h)
    return F.linear(h, sd[prefix + "fc_up.weight"], sd[prefix + "fc_up.bias"])


def unwrap_compacter(sd):
    """CompacterModel (Downstream/Text/run.py:70-81) wraps the model as `.model` and owns the shared `phm_rule`: its
    state_dict is {"phm_rule", "model.<inner key>"...} (every PHMLinear repeats the shared rule under its own name; only
    the top-level entry is read here).  Returns the inner-key view; a no-op for every other variant."""
    if "phm_rule" not in sd:
        return sd
    out = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    out["phm_rule"] = sd["phm_rule"]
    return out


def phm_linear(x, sd, prefix, phm_rule):
    """PHMLinear.forward with factorized_phm=True, shared_phm_rule=True (Downstream/Text/model/layers.py:153-166,
    matvec_product :10-22, kronecker.py:23-34): W = bmm(W_left, W_right); H = Σ_b rule[b] ⊗ W[b]; y = x·H + b."""
    W = torch.bmm(sd[prefix + "W_left"], sd[prefix + "W_right"])
    A = phm_rule
    H = torch.einsum('bac,bkp->bakcp', A, W).reshape(A.size(0), A.size(1) * W.size(1), A.size(2) * W.size(2)).sum(0)
    return torch.matmul(x, H) + sd[prefix + "b"]


def hypercomplex_adapter_block(x, sd, prefix):
    """HyperComplexAdapterBlock.forward, modules.py:247-250: up_sampler(gelu_new(down_sampler(x))), no residual;
    gelu_new = transformers' tanh approximation."""
    z = F.gelu(phm_linear(x, sd, prefix + "down_sampler.", sd["phm_rule"]), approximate="tanh")
    return phm_linear(z, sd, prefix + "up_sampler.", sd["phm_rule"])


def lora_linear(x, sd, prefix):
    """loralib 0.1.1 Linear.forward as used at Downstream/Text/run.py:414-428 (r > 0, lora_alpha = 1, lora_dropout = 0,
    not merged): F.linear(x, W, b) + (x @ lora_A.T @ lora_B.T) * (lora_alpha / r)."""
    y = F.linear(x, sd[prefix + "weight"], sd.get(prefix + "bias"))
    if prefix + "lora_A" in sd:
        a, b = sd[prefix + "lora_A"], sd[prefix + "lora_B"]
        y = y + (x @ a.t() @ b.t()) * (1.0 / a.shape[0])
    return y


def linear_or_lora(x, sd, prefix):
    return lora_linear(x, sd, prefix)


def self_output(hidden, input_tensor, sd, prefix, eps, activation, parallel=False):
    """BertSelfOutput / BertOutput (transformers) or, by the wrapper keys present, one of
    * BertAdaptedSelfOutput.forward, Downstream/Text/model/model.py:292-297 (Houlsby serial):
      dense -> dropout (eval: identity) -> adapter -> LayerNorm(h + input);
    * BertAdaptedParallelSelfOutput.forward, model.py:265-270 (`parallel`; same keys as the serial wrapper):
      LayerNorm(adapter(input) + dense(hidden) + input);
    * BertPfeifferAdaptedSelfOutput.forward, model.py:321-329 (has the extra `LN`): h = dense(hidden);
      t = LayerNorm(h + input); LN(adapter(t) + h + input), LN eps 1e-6."""
    if prefix + "adapter.down_sampler.W_left" in sd:
        # BertCompacterAdaptedSelfOutput.forward, model.py:715-720: LayerNorm(adapter(dense(hidden)) + input)
        h = F.linear(hidden, sd[prefix + "self_output.dense.weight"], sd[prefix + "self_output.dense.bias"])
        h = hypercomplex_adapter_block(h, sd, prefix + "adapter.")
        return layer_norm(h + input_tensor, sd, prefix + "self_output.LayerNorm.", eps)
    if prefix + "LN.weight" in sd:
        h = F.linear(hidden, sd[prefix + "self_output.dense.weight"], sd[prefix + "self_output.dense.bias"])
        t = layer_norm(h + input_tensor, sd, prefix + "self_output.LayerNorm.", eps)
        a = adapter_pfeiffer_block(t, sd, prefix + "adapter.", activation) + h
        return layer_norm(a + input_tensor, sd, prefix + "LN.", 1e-6)
    if prefix + "self_output.dense.weight" in sd:
        h = F.linear(hidden, sd[prefix + "self_output.dense.weight"], sd[prefix + "self_output.dense.bias"])
        if parallel:
            a = adapter_block(input_tensor, sd, prefix + "adapter.", activation)
            return layer_norm(a + h + input_tensor, sd, prefix + "self_output.LayerNorm.", eps)
        h = adapter_block(h, sd, prefix + "adapter.", activation)
        return layer_norm(h + input_tensor, sd, prefix + "self_output.LayerNorm.", eps)
    h = F.linear(hidden, sd[prefix + "dense.weight"], sd[prefix + "dense.bias"])
    return layer_norm(h + input_tensor, sd, prefix + "LayerNorm.", eps)


def bert_embeddings(ids, sd, cfg, n_tokens=0):
    """BertEmbeddings / RobertaEmbeddings.forward (transformers) with the SoftEmbedding substitution of
    Downstream/Text/model/model.py:620-630 when `embeddings.word_embeddings.learned_embedding` is present."""
    p = BERT_PREFIX + "embeddings."
    L = ids.shape[1]
    if p + "word_embeddings.learned_embedding" in sd:
        learned = sd[p + "word_embeddings.learned_embedding"]
        n = learned.shape[0]
        w = torch.cat([learned.unsqueeze(0).expand(ids.shape[0], -1, -1), sd[p + "word_embeddings.wte.weight"][ids[:, n:]]], 1)
    else:
        w = sd[p + "word_embeddings.weight"][ids]
    if cfg.roberta:
        m = (ids != cfg.pad_token_id).long()
        pos_ids = torch.cumsum(m, 1) * m + cfg.pad_token_id
    else:
        pos_ids = torch.arange(L).unsqueeze(0).expand_as(ids)
    x = w + sd[p + "token_type_embeddings.weight"][0] + sd[p + "position_embeddings.weight"][pos_ids]
    return layer_norm(x, sd, p + "LayerNorm.", cfg.eps)


def bert_layer(x, add_mask, sd, i, cfg, activation, parallel=False):
    """One BertLayer (post-LN).  q/k/v may be loralib Linears (run.py:416-421); attention.output / output may be
    Houlsby-wrapped (run.py:456-460).  Attention: softmax(q kᵀ / sqrt(d) + mask) v with the transformers additive
    mask (1 - m) * finfo(float32).min."""
    p = BERT_PREFIX + "encoder.layer.%d." % i
    N, L, H = x.shape
    dh = H // cfg.heads
    q = linear_or_lora(x, sd, p + "attention.self.query.").view(N, L, cfg.heads, dh).transpose(1, 2)
    k = linear_or_lora(x, sd, p + "attention.self.key.").view(N, L, cfg.heads, dh).transpose(1, 2)
    v = linear_or_lora(x, sd, p + "attention.self.value.").view(N, L, cfg.heads, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(dh) + add_mask
    ctx = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(N, L, H)
    y = self_output(ctx, x, sd, p + "attention.output.", cfg.eps, activation, parallel)
    f = F.gelu(F.linear(y, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    return self_output(f, y, sd, p + "output.", cfg.eps, activation, parallel)


class _Heads:
    """the two fields sasrec_block reads, for the TransformerBlocks inside a KAdapterBlock"""

    def __init__(self, heads):
        self.heads, self.adapter_activation, self.parallel = heads, "RELU", False


def kadapter_block(x, sd, prefix, heads):
    """KAdapterBlock.forward, Downstream/Text/model/modules.py:176-200: down_project -> 2 TransformerBlocks with the
    additive mask (1 - ones) * -10000 = 0 (no causal structure, no padding mask) -> up_project; + input."""
    h = F.linear(x, sd[prefix + "down_project.weight"], sd[prefix + "down_project.bias"])
    for j in range(2):
        h = sasrec_block(h, 0.0, sd, prefix + "transformer_blocks.%d." % j, _Heads(heads))
    return x + F.linear(h, sd[prefix + "up_project.weight"], sd[prefix + "up_project.bias"])


def bert_encoder(text, sd, cfg, rec):
    """Bert_Encoder.forward + Text_Encoder.forward, Downstream/Text/model/encoders.py:48-57,89-99:
    text [N, 2L] = ids | attention mask; returns GELU(fc(hidden[:, 0])) [N, D].
    With the K-Adapter wrapper (BertKAdaptedBertModel.forward, model.py:545-561; keys `bert_model.bert_model.*`,
    `bert_model.bert_adapter_list.*`, `bert_model.com_dense.*`): hidden_states = (embeddings, layer outputs...);
    last = adapter_i(hidden_states[k_i + 1] + last); hidden = com_dense([sequence_output | last])."""
    sd = unwrap_compacter(sd)
    kad = BERT_PREFIX + "com_dense.weight" in sd
    if kad:
        inner = BERT_PREFIX + "bert_model."
        body = {BERT_PREFIX + k[len(inner):]: v for k, v in sd.items() if k.startswith(inner)}
    else:
        body = sd
    L = text.shape[1] // 2
    ids, mask = text[:, :L], text[:, L:]
    x = bert_embeddings(ids, body, cfg, rec.n_tokens)
    add_mask = (1.0 - mask.float()).view(-1, 1, 1, L) * torch.finfo(torch.float32).min
    hidden_states = [x]
    for i in range(cfg.layers):
        x = bert_layer(x, add_mask, body, i, cfg, rec.adapter_activation, getattr(rec, "parallel", False))
        hidden_states.append(x)
    if kad:
        last = torch.zeros_like(x)
        for index, k in enumerate(int(i) + 1 for i in rec.k_adapter_bert_list.split(",")):
            last = kadapter_block(hidden_states[k] + last, sd, BERT_PREFIX + "bert_adapter_list.%d." % index,
                                  rec.num_adapter_heads_bert)
        x = F.linear(torch.cat([x, last], dim=2), sd[BERT_PREFIX + "com_dense.weight"], sd[BERT_PREFIX + "com_dense.bias"])
    cls = F.linear(x[:, 0], sd[FC_PREFIX + "weight"], sd[FC_PREFIX + "bias"])
    return F.gelu(cls)


def sasrec_block(x, att_mask, sd, p, rec):
    """TransformerBlock.forward (Downstream/Text/model/modules.py:45-87) or, by the wrapper keys present, one of
    * SASRecAdaptedSelfOutput.forward (model.py:341-376): adapter1 after fc, adapter2 after the FFN, both before the
      LayerNorms; SASRecPfeifferVer2AdaptedSelfOutput (model.py:389-423) is the same without adapter2;
    * SASRecParallelAdaptedSelfOutput.forward (model.py:484-520, `rec.parallel`): layer_norm(adapter1(x) + x + fc(..)),
      layer_norm(adapter2(y) + y + ffn(y));
    * SASRecPfeifferAdaptedSelfOutput.forward (model.py:437-471; has `adapter.` and `LN.`): plain attention half,
      t = layer_norm(y + ffn(y)); LN(adapter(t) + ffn(y) + y).
    w_Q / w_V may be loralib Linears with a bias (run.py:425-428)."""
    wrapped = p + "transformer_block.multi_head_attention.w_Q.weight" in sd
    pfeiffer = p + "LN.weight" in sd
    parallel = wrapped and getattr(rec, "parallel", False)
    act = rec.adapter_activation
    tb = p + ("transformer_block." if wrapped else "")
    a = tb + "multi_head_attention."
    B, S, D = x.shape
    dk = D // rec.heads
    q = linear_or_lora(x, sd, a + "w_Q.").view(B, S, rec.heads, dk).transpose(1, 2)
    k = linear_or_lora(x, sd, a + "w_K.").view(B, S, rec.heads, dk).transpose(1, 2)
    v = linear_or_lora(x, sd, a + "w_V.").view(B, S, rec.heads, dk).transpose(1, 2)
    attn = q @ k.transpose(-2, -1) / (dk ** 0.5) + att_mask
    h = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(B, S, D)
    h = F.linear(h, sd[a + "fc.weight"])
    compacter = p + "adapter1.down_sampler.W_left" in sd      # SASRecCompacterAdaptedSelfOutput, model.py:659-693
    if compacter:
        y = layer_norm(x + hypercomplex_adapter_block(h, sd, p + "adapter1."), sd, a + "layer_norm.", 1e-6)
    elif parallel:
        y = layer_norm(adapter_block(x, sd, p + "adapter1.", act) + x + h, sd, a + "layer_norm.", 1e-6)
    else:
        if p + "adapter1.fc_down.weight" in sd:
            h = adapter_block(h, sd, p + "adapter1.", act)
        y = layer_norm(x + h, sd, a + "layer_norm.", 1e-6)
    f = tb + "feed_forward."
    h = F.linear(F.relu(F.linear(y, sd[f + "w_1.weight"], sd[f + "w_1.bias"])), sd[f + "w_2.weight"], sd[f + "w_2.bias"])
    if compacter:
        return layer_norm(y + hypercomplex_adapter_block(h, sd, p + "adapter2."), sd, f + "layer_norm.", 1e-6)
    if parallel:
        return layer_norm(adapter_block(y, sd, p + "adapter2.", act) + y + h, sd, f + "layer_norm.", 1e-6)
    if pfeiffer:
        t = layer_norm(y + h, sd, f + "layer_norm.", 1e-6)
        return layer_norm(adapter_pfeiffer_block(t, sd, p + "adapter.", act) + h + y, sd, p + "LN.", 1e-6)
    if p + "adapter2.fc_down.weight" in sd:
        h = adapter_block(h, sd, p + "adapter2.", act)
    return layer_norm(y + h, sd, f + "layer_norm.", 1e-6)


def user_encoder(input_embs, log_mask, sd, rec):
    """User_Encoder.forward (encoders.py:24-29) + TransformerEncoder.forward (modules.py:101-113).
    Mask: 0 where (key j <= query i and log_mask[b, j] != 0) else -1e9, shape [B,1,S,S]."""
    sd = unwrap_compacter(sd)
    B, S, D = input_embs.shape
    valid = (log_mask != 0).view(B, 1, 1, S).expand(-1, -1, S, -1)
    att_mask = torch.where(torch.tril(valid), 0.0, -1e9)
    x = input_embs + sd[USER_PREFIX + "position_embedding.weight"][:S].unsqueeze(0)
    x = layer_norm(x, sd, USER_PREFIX + "layer_norm.", 1e-6)
    kp = USER_PREFIX + "transformer_blocks."
    if kp + "com_dense2.weight" in sd:
        # SASRecKAdaptedTransformerBlocks.forward, model.py:575-583 (dispatch: modules.py:108-109)
        last = torch.zeros_like(x)
        for j in range(rec.blocks):
            last = kadapter_block(x + last, sd, kp + "adapter_list.%d." % j, rec.num_adapter_heads_sasrec)
            x = sasrec_block(x, att_mask, sd, kp + "transformer_blocks.%d." % j, rec)
        return F.linear(torch.cat([x, last], dim=2), sd[kp + "com_dense2.weight"], sd[kp + "com_dense2.bias"])
    for j in range(rec.blocks):
        x = sasrec_block(x, att_mask, sd, USER_PREFIX + "transformer_blocks.%d." % j, rec)
    return x


def bce_with_logits_mean(x, target_one):
    """nn.BCEWithLogitsLoss (mean) against an all-ones / all-zeros target: mean softplus(-x) / mean softplus(x)."""
    return F.softplus(-x).mean() if target_one else F.softplus(x).mean()


def loss_from_embeddings(embs_all, log_mask, sd, rec, cpc=False):
    """Model.forward lines 53-68 / ModelCPC.forward lines 120-133 (Downstream/Text/model/model.py), given the item
    embeddings [B*(S+1)*2, D]."""
    S1 = rec.max_seq_len + 1
    e = embs_all.view(-1, S1, 2, rec.embedding_dim)
    pos, neg = e[:, :, 0], e[:, :, 1]
    prec = user_encoder(pos[:, :-1], log_mask, sd, rec)
    ps = (prec * pos[:, 1:]).sum(-1)
    ns = (prec * neg[:, :-1]).sum(-1)
    if cpc:
        return bce_with_logits_mean(ps[:, -1], True) + bce_with_logits_mean(ns[:, -1], False)
    idx = torch.where(log_mask != 0)
    return bce_with_logits_mean(ps[idx], True) + bce_with_logits_mean(ns[idx], False)


def inbatch_softmax_loss(prec, cand, item_ids, log_mask, cand_bias=None, masked_logit=-1e4):
    """In-batch softmax loss with duplicate-item masking (BASELINE.json north_star).

    PARITY UNPINNED: /root/reference has no such head (it trains with BCE, Downstream/Text/model/model.py:62-68), so
    there is no reference output to pin this function to.  It restates, loop for loop, the in-batch debiased
    cross-entropy of the same group's IDvs.MoRec trainer (not under /root/reference, no pinned version): logits of
    every position against all B*(S+1) history items of the batch, minus the log-popularity of the candidate; padding
    candidate slots and every candidate whose item id occurs in the user's own sequence — except the positive itself —
    are overwritten with -1e4; CrossEntropyLoss(mean) over the valid positions.

    prec [B,S,D], cand [B,S+1,D] (the encoded history items), item_ids [B,S+1] int64, log_mask [B,S]."""
    B, S, D = prec.shape
    S1 = S + 1
    score_embs = cand.reshape(B * S1, D)
    ce_label = torch.tensor([i * S + i + j for i in range(B) for j in range(1, S1)], dtype=torch.long)
    logits = torch.matmul(prec.reshape(B * S, D), score_embs.t())                     # [B*S, B*(S+1)]
    if cand_bias is not None:
        logits = logits - cand_bias.reshape(1, -1)
    col_pad = torch.cat((log_mask, torch.ones(B, 1)), dim=1).view(-1) == 0
    fill = torch.full_like(logits, masked_logit)
    logits = torch.where(col_pad.unsqueeze(0), fill, logits)
    logits = logits.view(B, S, -1)
    flat_ids = item_ids.reshape(-1)
    out = []
    for i in range(B):
        reject = item_ids[i]                                                          # the user's own S+1 items
        mask_row = (flat_ids.unsqueeze(0) == reject.unsqueeze(1)).any(dim=0)          # [B*(S+1)]
        mask_mat = mask_row.unsqueeze(0).repeat(S, 1)
        for j in range(S):
            mask_mat[j][i * S1 + j + 1] = False                                       # keep the positive
        out.append(torch.where(mask_mat, torch.full_like(logits[i], masked_logit), logits[i]))
    logits = torch.stack(out).view(B * S, -1)
    idx = torch.where(log_mask.reshape(-1) != 0)
    return F.cross_entropy(logits[idx], ce_label[idx])


def loss_from_embeddings_inbatch(embs_all, item_ids, log_mask, sd, rec, cand_bias=None):
    """Model.forward with the in-batch softmax head instead of BCE: same encoder outputs, same user encoder
    (model.py:53-60); the sampled negatives embs[:, :, 1] are not used by this head."""
    S1 = rec.max_seq_len + 1
    e = embs_all.view(-1, S1, 2, rec.embedding_dim)
    pos = e[:, :, 0]
    prec = user_encoder(pos[:, :-1], log_mask, sd, rec)
    return inbatch_softmax_loss(prec, pos, item_ids, log_mask, cand_bias)


def model_forward(sample_items, log_mask, sd, cfg, rec, cpc=False):
    """Model.forward / ModelCPC.forward: sample_items [B*(S+1)*2, 2L] int64, log_mask [B,S] -> scalar loss."""
    return loss_from_embeddings(bert_encoder(sample_items, sd, cfg, rec), log_mask, sd, rec, cpc)


# ------------------------------------------------------------------------------------------------------------------
# evaluation (Downstream/Text/data_utils/metrics.py:51-116, dataset.py:65-78)
# ------------------------------------------------------------------------------------------------------------------
def item_embeddings(item_content, sd, cfg, rec, batch=512):
    """get_item_embeddings, metrics.py:62-79: encoder over all I+1 item rows (row 0 = all-zero ids and mask)."""
    with torch.no_grad():
        return torch.cat([bert_encoder(item_content[i:i + batch], sd, cfg, rec) for i in range(0, len(item_content), batch)])


def eval_user_vectors(seqs, item_emb, sd, rec):
    """BuildEvalDataset.__getitem__ (dataset.py:65-78) + eval_model's user-encoder call (metrics.py:102-104):
    tokens = seq[:-1] left-padded with item 0 to max_seq_len, gather item embeddings, last-position output."""
    S = rec.max_seq_len
    B = len(seqs)
    tok = torch.zeros((B, S), dtype=torch.long)
    mask = torch.zeros((B, S))
    for b, seq in enumerate(seqs):
        t = seq[:-1]
        tok[b, S - len(t):] = torch.tensor(t)
        mask[b, S - len(t):] = 1.0
    with torch.no_grad():
        return user_encoder(item_emb[tok], mask, sd, rec)[:, -1], tok, mask


def rank_metrics(scores, history, target, topk=10):
    """metrics.py:105-111 + metrics_topK (:51-59) for ONE user.  scores [I+1] over ids 0..I; history ids are set to
    -inf; id 0 is dropped; rank = 1 + #{j : s_j > s_t} (the reference's argsort leaves ties unspecified; the
    tie-break fixed here and in the CUDA path is (score desc, id asc), i.e. equal scores with a smaller id rank
    ahead).  Returns (hit, ndcg)."""
    s = scores.clone()
    s[torch.as_tensor(history, dtype=torch.long)] = -float("inf")
    st = s[target]
    ids = torch.arange(s.shape[0])
    ahead = ((s > st) | ((s == st) & (ids < target)))[1:].sum().item()
    rank = 1 + ahead
    if rank <= topk:
        return 1.0, 1.0 / math.log2(rank + 1)
    return 0.0, 0.0


def topk_ids(scores, history, k=10):
    """Top-k item ids for one user under the total order (score desc, id asc), history masked, id 0 excluded."""
    s = scores.clone().double()
    s[torch.as_tensor(history, dtype=torch.long)] = -float("inf")
    s[0] = -float("inf")
    order = sorted(range(s.shape[0]), key=lambda j: (-s[j].item(), j))
    return order[:k]


def eval_model(seqs, histories, item_emb, sd, rec, topk=10):
    """eval_model, metrics.py:82-116 (single process): per-user (hit, ndcg) and their means."""
    u, _, _ = eval_user_vectors(seqs, item_emb, sd, rec)
    scores = u @ item_emb.t()
    res = [rank_metrics(scores[b], histories[b], seqs[b][-1], topk) for b in range(len(seqs))]
    hit = torch.tensor([r[0] for r in res])
    ndcg = torch.tensor([r[1] for r in res])
    return hit, ndcg


# ------------------------------------------------------------------------------------------------------------------
# train-batch assembly (Downstream/Text/data_utils/dataset.py:24-49)
# ------------------------------------------------------------------------------------------------------------------
def train_sample(seq, item_content, item_num, max_seq_len, rng):
    """BuildTrainDataset.__getitem__, dataset.py:24-49, for ONE user.  `rng` is Python's `random` module (or a
    random.Random): the rejection loop calls rng.randint(1, item_num) in the reference's order, so under the same seed it
    draws the reference's negatives.  Returns (sample_items [S+1, 2, 2L] int64, log_mask [S] f32, ids [S+1, 2] int64)."""
    S1 = max_seq_len + 1
    seq = list(seq)
    tokens_len = len(seq) - 1
    head = S1 - len(seq)
    log_mask = [0] * head + [1] * tokens_len
    neg_items = []
    for _ in range(tokens_len):
        sam_neg = rng.randint(1, item_num)
        while sam_neg in seq:
            sam_neg = rng.randint(1, item_num)
        neg_items.append(sam_neg)
    ids = torch.tensor([[0] * head + seq, [0] * head + neg_items + [0]], dtype=torch.long).t().contiguous()
    return torch.as_tensor(item_content).long()[ids], torch.tensor(log_mask, dtype=torch.float32), ids


# ------------------------------------------------------------------------------------------------------------------
# image tree (Downstream/CV): ViT item encoder with Houlsby / LoRA / soft-prompt variants
# ------------------------------------------------------------------------------------------------------------------
VIT_PREFIX = "cv_encoder.image_net.vit."
CLS_PREFIX = "cv_encoder.image_net.classifier."


class VitConfig:
    def __init__(self, hidden=768, layers=12, heads=12, patch=16, eps=1e-12):
        self.hidden, self.layers, self.heads, self.patch, self.eps = hidden, layers, heads, patch, eps


def vit_embeddings(images, sd, cfg):
    """ViTEmbeddings.forward (transformers) or SoftPrompt.forward (Downstream/CV/model/model.py:523-535) when
    `embeddings.Prompt_Tokens` is present: conv patch projection, [cls | patches] + positions, then the prompt tokens
    appended WITHOUT positions."""
    e = VIT_PREFIX + "embeddings."
    w = "wte." if e + "Prompt_Tokens" in sd else ""
    x = F.conv2d(images, sd[e + w + "patch_embeddings.projection.weight"], sd[e + w + "patch_embeddings.projection.bias"],
                 stride=cfg.patch).flatten(2).transpose(1, 2)
    x = torch.cat([sd[e + w + "cls_token"].expand(x.shape[0], -1, -1), x], 1) + sd[e + w + "position_embeddings"]
    if w:
        x = torch.cat([x, sd[e + "Prompt_Tokens"].expand(x.shape[0], -1, -1)], 1)
    return x


def vit_layer(x, sd, i, cfg, activation, parallel=False):
    """transformers' ViTLayer.forward (pre-LN) with the wrappers of Downstream/CV/model/model.py:182-212:
    VITAdaptedSelfOutput = adapter(dense(ctx)) (no residual), VITAdaptedOutput = adapter(dense(h)) + input."""
    p = VIT_PREFIX + "encoder.layer.%d." % i
    N, L, H = x.shape
    dh = H // cfg.heads
    xn = layer_norm(x, sd, p + "layernorm_before.", cfg.eps)
    q = linear_or_lora(xn, sd, p + "attention.attention.query.").view(N, L, cfg.heads, dh).transpose(1, 2)
    k = linear_or_lora(xn, sd, p + "attention.attention.key.").view(N, L, cfg.heads, dh).transpose(1, 2)
    v = linear_or_lora(xn, sd, p + "attention.attention.value.").view(N, L, cfg.heads, dh).transpose(1, 2)
    ctx = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1) @ v).transpose(1, 2).reshape(N, L, H)
    ao = p + "attention.output."
    x1 = _vit_sublayer_output(ctx, x, sd, ao, activation, parallel)
    f = F.gelu(F.linear(layer_norm(x1, sd, p + "layernorm_after.", cfg.eps), sd[p + "intermediate.dense.weight"],
                        sd[p + "intermediate.dense.bias"]))
    return _vit_sublayer_output(f, x1, sd, p + "output.", activation, parallel)


def _vit_sublayer_output(hidden, skip, sd, o, activation, parallel):
    """attention.output / output of one ViTLayer INCLUDING the layer's skip connection, by the wrapper keys present
    (Downstream/CV/model/model.py): plain dense(h) + skip; VITAdaptedSelfOutput / VITAdaptedOutput (:182-212)
    adapter(dense(h)) + skip; VITAdaptedParallelOutput (:165-179, `parallel`) dense(h) + skip + adapter(skip);
    VITCompacterAdapted{Self,}Output (:432-462) hypercomplex_adapter(dense(h)) + skip."""
    if o + "self_output.dense.weight" not in sd:
        return F.linear(hidden, sd[o + "dense.weight"], sd[o + "dense.bias"]) + skip
    h = F.linear(hidden, sd[o + "self_output.dense.weight"], sd[o + "self_output.dense.bias"])
    if o + "adapter.down_sampler.W_left" in sd:
        return hypercomplex_adapter_block(h, sd, o + "adapter.") + skip
    if parallel:
        return h + skip + adapter_block(skip, sd, o + "adapter.", activation)
    return adapter_block(h, sd, o + "adapter.", activation) + skip


def vit_encoder(images, sd, cfg, rec):
    """Vit_Encoder.forward (Downstream/CV/model/encoders.py:25-32): GELU(classifier(LN(h)[:, 0]))."""
    sd = unwrap_compacter(sd)
    x = vit_embeddings(images, sd, cfg)
    for i in range(cfg.layers):
        x = vit_layer(x, sd, i, cfg, rec.adapter_activation, getattr(rec, "parallel", False))
    x = layer_norm(x, sd, VIT_PREFIX + "layernorm.", cfg.eps)
    return F.gelu(F.linear(x[:, 0], sd[CLS_PREFIX + "weight"], sd[CLS_PREFIX + "bias"]))


def cv_model_forward(images, log_mask, sd, cfg, rec, cpc=False):
    """Model.forward / ModelCPC.forward of the image tree (Downstream/CV/model/model.py:54-77)."""
    return loss_from_embeddings(vit_encoder(images, sd, cfg, rec), log_mask, sd, rec, cpc)
