"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference/Downstream/Text)
on the seeded weights/inputs of tests/golden/cases.py.  Run in the build container (the GPU box has no
/root/reference):   python tests/golden/make_golden.py

The adapter surgery below repeats Downstream/Text/run.py:414-465 verbatim in effect (those lines live inside
train() and cannot be imported).  loralib is not installed; `_LoraLinear` restates loralib 0.1.1 Linear
(r > 0, lora_alpha = 1, no dropout, unmerged) — the only non-reference code on the golden path, used for the
'lora' case only."""
import logging
import math
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/Downstream/Text")

import cases  # noqa: E402


class _LoraLinear(nn.Linear):
    def __init__(self, in_features, out_features, r=0, lora_alpha=1, **kw):
        super().__init__(in_features, out_features, **kw)
        self.r, self.scaling = r, lora_alpha / r
        self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
        self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
        self.weight.requires_grad = False
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))

    def forward(self, x):
        return nn.functional.linear(x, self.weight, self.bias) + (x @ self.lora_A.t() @ self.lora_B.t()) * self.scaling


sys.modules["loralib"] = types.SimpleNamespace(Linear=_LoraLinear)

import torch.distributed as dist  # noqa: E402
from transformers import BertConfig, BertModel, RobertaConfig, RobertaModel  # noqa: E402

from data_utils.metrics import eval_model, get_item_embeddings, metrics_topK  # noqa: E402
from model import (BertAdaptedParallelSelfOutput, BertAdaptedSelfOutput, BertPfeifferAdaptedSelfOutput, Model,  # noqa: E402
                   ModelCPC, SASRecAdaptedSelfOutput, SASRecParallelAdaptedSelfOutput, SASRecPfeifferAdaptedSelfOutput,
                   SASRecPfeifferVer2AdaptedSelfOutput, SoftEmbedding)


def build_reference_model(c):
    args = cases.reference_args(c)
    kw = dict(vocab_size=c.vocab, hidden_size=c.hidden, num_hidden_layers=c.layers, num_attention_heads=c.heads,
              intermediate_size=c.inter, max_position_embeddings=c.max_pos, layer_norm_eps=c.eps,
              hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    if c.kind == "kadapter":
        # K-Adapter is the one variant where the all-masked padding item (row 0) reaches valid positions (its SASRec-side
        # adapters attend over the whole sequence with an all-ones mask), so the additive-mask softmax of the pinned
        # transformers 4.20.1 must be reproduced: "eager" attention IS that algorithm (scores + finfo.min mask -> uniform
        # softmax); the default sdpa path of the installed 5.5.0 returns a zero context for such rows instead
        kw["attn_implementation"] = "eager"
    if c.roberta:
        bert = RobertaModel(RobertaConfig(type_vocab_size=1, pad_token_id=1, **kw))
    else:
        bert = BertModel(BertConfig(**kw))
    model = (ModelCPC if c.cpc else Model)(args, c.item_num, True, bert)
    full_ft = c.kind.startswith("full_ft")                         # fine_tune_to = all: nothing frozen but the pooler
    for n, p in model.named_parameters():                          # run.py:369-371 (fine_tune_to=None)
        p.requires_grad = full_ft and "pooler" not in n
    layers = model.bert_encoder.text_encoders.title.bert_model.encoder.layer
    blocks = model.user_encoder.transformer_encoder.transformer_blocks
    if c.kind.startswith("houlsby"):                               # run.py:456-465
        for lm in layers:
            lm.attention.output = BertAdaptedSelfOutput(lm.attention.output, args)
            lm.output = BertAdaptedSelfOutput(lm.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecAdaptedSelfOutput(tb, args)
    elif c.kind == "parallel":                                     # run.py:466-479 (is_serial == "None")
        for lm in layers:
            lm.attention.output = BertAdaptedParallelSelfOutput(lm.attention.output, args)
            lm.output = BertAdaptedParallelSelfOutput(lm.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecParallelAdaptedSelfOutput(tb, args)
    elif c.kind == "pfeiffer_ver2":                                # run.py:389-399
        for lm in layers:
            lm.attention.output = BertAdaptedSelfOutput(lm.attention.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecPfeifferVer2AdaptedSelfOutput(tb, args)
    elif c.kind in ("pfeiffer", "pfeiffer_leaky"):                 # run.py:400-409
        for lm in layers:
            lm.output = BertPfeifferAdaptedSelfOutput(lm.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecPfeifferAdaptedSelfOutput(tb, args)
    elif c.kind == "compacter":                                    # run.py:435-450
        from model import BertCompacterAdaptedSelfOutput, SASRecCompacterAdaptedSelfOutput
        from run import CompacterModel                             # the reference's own wrapper class (entry script)
        for lm in layers:
            lm.attention.output = BertCompacterAdaptedSelfOutput(lm.attention.output, args)
            lm.output = BertCompacterAdaptedSelfOutput(lm.output, args)
        for i, tb in enumerate(blocks):
            blocks[i] = SASRecCompacterAdaptedSelfOutput(tb, args)
        model = CompacterModel(args, model)
    elif c.kind == "kadapter":                                     # run.py:409-413
        from model import BertKAdaptedBertModel, SASRecKAdaptedTransformerBlocks
        bert.config.output_hidden_states = True                    # the reference loads the config with it (run.py:293)
        title = model.bert_encoder.text_encoders.title
        title.bert_model = BertKAdaptedBertModel(title.bert_model, args)
        te = model.user_encoder.transformer_encoder
        te.transformer_blocks = SASRecKAdaptedTransformerBlocks(te.transformer_blocks, args)
    elif c.kind == "lora":                                         # run.py:414-428
        import loralib as lora
        for lm in layers:
            lm.attention.self.query = lora.Linear(args.word_embedding_dim, args.word_embedding_dim, r=args.bert_adapter_down_size)
            lm.attention.self.value = lora.Linear(args.word_embedding_dim, args.word_embedding_dim, r=args.bert_adapter_down_size)
        for i, tb in enumerate(blocks):
            blocks[i].multi_head_attention.w_Q = lora.Linear(args.embedding_dim, args.embedding_dim, r=args.adapter_down_size)
            blocks[i].multi_head_attention.w_V = lora.Linear(args.embedding_dim, args.embedding_dim, r=args.adapter_down_size)
    elif c.kind == "prompt_cpc":                                   # run.py:429-434
        s_wte = SoftEmbedding(bert.get_input_embeddings(), n_tokens=args.n_tokens, initialize_from_vocab=True)
        model.bert_encoder.text_encoders.title.bert_model.set_input_embeddings(s_wte)
    return model, args


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    dist.init_process_group("gloo", rank=0, world_size=1)
    import transformers
    meta = {"torch": torch.__version__, "transformers": transformers.__version__}
    only = sys.argv[1:]                                            # optional: regenerate just the named kinds
    for kind in (only or cases.ALL_KINDS):
        c = cases.tiny_case(kind)
        model, args = build_reference_model(c)
        sd = cases.build_state_dict(c)
        ref_keys = set(model.state_dict().keys())
        missing = ref_keys - set(sd.keys())
        extra = set(sd.keys()) - ref_keys
        # transformers registers position_ids / token_type_ids as (non-persistent or persistent) buffers: not weights
        missing = {k for k in missing if not k.endswith(("position_ids", "token_type_ids"))}
        assert not missing and not extra, (kind, sorted(missing), sorted(extra))
        model.load_state_dict(sd, strict=False)
        train_keys = cases.trainable_keys(c, sd)
        got_train = sorted(n for n, p in model.named_parameters() if p.requires_grad)
        assert got_train == sorted(train_keys), (kind, got_train, sorted(train_keys))
        model.eval()                                               # dropout off: parity is defined without dropout
        items = cases.build_item_content(c)
        sample_items, log_mask, id_batch = cases.build_batch(c, items)
        out = {"meta": meta, "kind": kind}
        loss = model(sample_items.view(-1, 2 * c.L), log_mask, "cpu")
        if train_keys:
            loss.backward()
            out["grads"] = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
        out["loss"] = loss.detach().clone()
        with torch.no_grad():
            inner = model.model if c.kind == "compacter" else model        # as metrics.py:72-73,101-102 reach in
            out["batch_item_emb"] = inner.bert_encoder(sample_items.view(-1, 2 * c.L)).clone()
            wrapper = types.SimpleNamespace(module=model, eval=model.eval)
            emb = get_item_embeddings(wrapper, items.numpy(), 16, args, True, "cpu")
            out["item_emb"] = emb.clone()
            seqs, hist = cases.build_eval_users(c)
            log = logging.getLogger("golden")
            hit10 = eval_model(wrapper, [torch.LongTensor(h) for h in hist], {i: s for i, s in enumerate(seqs)}, emb, 4,
                               args, c.item_num, log, "test", "cpu")
            out["eval_hit10_mean"] = float(hit10)
            # per-user values through the reference's own metrics_topK, exactly as eval_model's inner loop does
            S = c.S
            per_user = []
            for u, seq in enumerate(seqs):
                toks = seq[:-1]
                pad = [0] * (S - len(toks)) + toks
                mask = torch.FloatTensor([0] * (S - len(toks)) + [1] * len(toks))
                prec = inner.user_encoder(emb[pad].unsqueeze(0), mask.unsqueeze(0), "cpu")[:, -1]
                score = torch.matmul(prec, emb.t()).squeeze(0)
                score[torch.LongTensor(hist[u])] = -float("inf")
                score = score[1:]
                label = torch.zeros(c.item_num, dtype=torch.float64)
                label[seq[-1] - 1] = 1.0
                item_rank = torch.Tensor(range(1, c.item_num + 1))
                per_user.append(metrics_topK(score, label, item_rank, 10, "cpu").clone())
                if u == 0:
                    out["user0_scores"] = score.clone()
            out["eval_per_user"] = torch.stack(per_user)
        path = os.path.join(HERE, "transrec_%s.pt" % kind)
        torch.save(out, path)
        print(kind, "loss %.6f" % float(out["loss"]), "hit10 %.4f" % out["eval_hit10_mean"],
              "per-user mean", out["eval_per_user"].mean(0).tolist(), "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
