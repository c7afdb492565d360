"""-m gpu: the full train path (item encoder -> user encoder -> loss, forward AND backward) through the drop-in
model classes and the C ABI, against (a) the CPU oracle on the same seeded weights/inputs and (b) the golden outputs
of the unmodified reference.  bf16 activations with fp32 accumulation vs an fp32 oracle: tolerances stated inline."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import cases  # noqa: E402
import transrec_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
KINDS = list(cases.ALL_KINDS)

# Tolerances: tests/parity_util.py (2 x the error recorded on a B200 in tests/golden/parity_measured.json, per case and per
# trainable tensor; floors = the resolution of an fp32-vs-bf16 comparison).


def build_gpu_model(c, sd):
    from adapter4rec_b200 import surgery
    from adapter4rec_b200.model import BertModel, Model, ModelCPC, RobertaModel, TextConfigLite
    args = cases.reference_args(c)
    args.adding_adapter_to, args.finetune_layernorm = "all", "None"
    cfg = TextConfigLite(vocab_size=c.vocab, hidden_size=c.hidden, num_hidden_layers=c.layers,
                         num_attention_heads=c.heads, intermediate_size=c.inter, max_position_embeddings=c.max_pos,
                         layer_norm_eps=c.eps, type_vocab_size=1 if c.roberta else 2, pad_token_id=c.pad)
    bert = (RobertaModel if c.roberta else BertModel)(cfg)
    model = (ModelCPC if c.cpc else Model)(args, c.item_num, True, bert).cuda()
    surgery.freeze_all(model)
    if c.kind.startswith("full_ft"):                        # fine_tune_to = all, pooler frozen (Pretraining/Text/run.py:48-64)
        for n, p in model.named_parameters():
            p.requires_grad = "pooler" not in n
    elif c.kind != "base":
        model = surgery.insert_adapters(model, args)        # compacter returns the CompacterModel wrapper
    assert set(model.state_dict().keys()) == set(sd.keys()), "state_dict keys must equal the reference's"
    model.load_state_dict(sd)
    got_train = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    assert got_train == sorted(cases.trainable_keys(c, sd))
    return model, args


def oracle_setup(c):
    cfg = O.TextConfig(hidden=c.hidden, layers=c.layers, heads=c.heads, eps=c.eps, roberta=c.roberta, pad_token_id=c.pad)
    rec = O.RecConfig(max_seq_len=c.S, embedding_dim=c.D, heads=c.rec_heads, blocks=c.blocks, num_words_title=c.L,
                      adapter_activation=c.activation, n_tokens=c.n_tokens, parallel=c.parallel,
                      k_adapter_bert_list=c.k_list, num_adapter_heads_bert=c.k_heads_bert,
                      num_adapter_heads_sasrec=c.k_heads_rec)
    return cfg, rec


@pytest.mark.parametrize("kind", KINDS)
def test_train_step_matches_oracle_and_reference(kind):
    """loss, every trainable gradient tensor and the item embeddings of one seeded step: CUDA path vs the fp32 oracle (bounds
    from the recorded measurements), and the oracle vs the golden of the UNMODIFIED reference (fp32 vs fp32: 2e-5 / 1e-3)."""
    import parity_util as P
    fig, obj = P.text_case(kind)
    gold = torch.load(os.path.join(HERE, "golden", "transrec_%s.pt" % kind), weights_only=False)
    ov, gv = fig["oracle_loss"], float(gold["loss"])
    assert abs(ov - gv) <= 2e-5 * abs(gv)
    train, osd = obj["train"], obj["osd"]
    if train:
        total_norm = float(torch.cat([osd[k].grad.flatten() for k in train]).norm())
        for k in train:
            # (the key-projection biases have an analytically ZERO gradient — softmax is invariant to a per-query
            # constant — so both sides hold rounding noise of ~1e-9 there: absolute floor relative to the whole gradient)
            og, ref_g = osd[k].grad, gold["grads"][k]
            assert float((og - ref_g).norm()) < 1e-3 * float(ref_g.norm()) + 1e-6 * total_norm
    # the oracle's item table against the reference's (row 0 = the padding item is excluded: DESIGN.md §2)
    assert float((obj["oracle_emb"][1:] - gold["item_emb"][1:]).abs().max()) <= 2e-5 * float(gold["item_emb"][1:].abs().max()) + 1e-6
    P.check_against_table(fig, P.measured(), "text/" + kind, bool(train))


@pytest.mark.parametrize("kind", KINDS)
def test_item_encoder_matches_reference(kind):
    """the item encoder alone (what get_item_embeddings runs) against the golden table of the unmodified reference"""
    import parity_util as P
    c = cases.tiny_case(kind)
    sd = cases.build_state_dict(c)
    gold = torch.load(os.path.join(HERE, "golden", "transrec_%s.pt" % kind), weights_only=False)
    model, args = build_gpu_model(c, sd)
    model.eval()
    items = cases.build_item_content(c)
    with torch.no_grad():
        from adapter4rec_b200.data_utils.metrics import core_model
        emb = core_model(model).bert_encoder(items.cuda()).float().cpu()
    ref = gold["item_emb"]
    err = (emb[1:] - ref[1:]).abs().max()
    lim = P.bound(P.measured(), "text/" + kind, "emb_max_abs") + 1e-5     # golden == oracle to 1e-6 (test above)
    assert float(err) <= lim, "max abs err %.4f > %.4f" % (float(err), lim)
    assert torch.isfinite(emb).all()   # incl. the all-masked padding item (row 0)


@pytest.mark.parametrize("kind", ["lora", "houlsby", "pfeiffer_ver2", "compacter"])
def test_unpadded_token_layout_gives_the_same_step(kind):
    """bert_model.unpad = True runs the encoder on the kept tokens only (PackedTokens).  Every kept token's arithmetic is
    the same instruction sequence on the same inputs, so embeddings and loss must agree to the last bf16 bit; gradients are
    sums over tokens whose split order changes with the token count (fp32 reassociation only: 1e-3)."""
    from adapter4rec_b200.data_utils.metrics import core_model
    c = cases.tiny_case(kind)
    sd = cases.build_state_dict(c)
    items = cases.build_item_content(c)
    sample_items, log_mask, _ = cases.build_batch(c, items)
    rows = sample_items.view(-1, 2 * c.L).cuda()
    out = {}
    for unpad in (False, True):
        model, _ = build_gpu_model(c, sd)
        model.eval()
        core_model(model).bert_encoder.text_encoders.title.bert_model.unpad = unpad
        with torch.no_grad():
            emb = core_model(model).bert_encoder(items.cuda()).float().cpu()
        loss = model(rows, log_mask.cuda(), 0)
        loss.backward()
        grads = {n: p.grad.float().cpu() for n, p in model.named_parameters() if p.requires_grad}
        out[unpad] = (emb, float(loss.detach()), grads)
    assert torch.equal(out[True][0], out[False][0]), "item embeddings must be bit-identical"
    assert out[True][1] == out[False][1], "loss must be bit-identical"
    tot = float(torch.cat([g.flatten() for g in out[False][2].values()]).norm())
    for n, g in out[False][2].items():
        assert float((out[True][2][n] - g).norm()) <= 1e-3 * float(g.norm()) + 1e-5 * tot, n


def test_item_dedup_gives_the_same_loss_and_gradients():
    """Model.dedup_items: every distinct item row of the batch is encoded once (tiny case: 72 slots over 40 items + the
    padding row).  Forward values are identical (the encoder is independent across items), so the loss is bit-equal in
    eval mode; gradients differ by the summation order of the occurrences only."""
    c = cases.tiny_case("lora")
    sd = cases.build_state_dict(c)
    model, _ = build_gpu_model(c, sd)
    items = cases.build_item_content(c)
    sample_items, log_mask, _ = cases.build_batch(c, items)
    rows, lm = sample_items.view(-1, 2 * c.L).cuda(), log_mask.cuda()
    assert torch.unique(rows, dim=0).shape[0] < rows.shape[0], "the case must contain repeated items"
    model.eval()
    out = []
    for flag in (False, True):
        model.dedup_items = flag
        model.zero_grad(set_to_none=True)
        loss = model(rows, lm, 0)
        loss.backward()
        out.append((float(loss), torch.cat([p.grad.float().flatten() for p in model.parameters() if p.grad is not None])))
    model.dedup_items = False
    assert out[0][0] == out[1][0], (out[0][0], out[1][0])
    assert float((out[0][1] - out[1][1]).norm() / out[0][1].norm()) <= 2e-2
