"""-m gpu: K9-S, the in-batch softmax head with duplicate-item masking (a4r_inbatch_ce_fwd / _bwd through the C ABI),
against the CPU oracle (oracle/transrec_oracle.py:inbatch_softmax_loss, the loop-for-loop MoRec statement; parity
unpinned by the reference, which has no softmax head) on the same bf16-exact inputs.

Tolerances: the logits are fp32 accumulations of bf16 products on both sides, so the loss agrees to 1e-4 relative (the
kernel's exp is ex2.approx); gradients carry one bf16 rounding of P - onehot before the second contraction and one of
the output: 2^-6 relative of the row/column scale."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

BF16 = torch.bfloat16


def _case(B, S, D, n_items, seed, ragged=True, bias=False):
    g = torch.Generator().manual_seed(seed)
    prec = (torch.randn(B, S, D, generator=g) * (2.0 * D ** -0.5)).to(BF16)
    emb = torch.randn(B, S + 1, 2, D, generator=g).to(BF16)
    ids = torch.randint(1, n_items + 1, (B, S + 1), generator=g)
    log_mask = torch.ones(B, S)
    if ragged:
        lens = torch.randint(1, S + 1, (B,), generator=g)      # valid positions per user, left-padded like the reference
        for b in range(B):
            pad = S - int(lens[b])
            log_mask[b, :pad] = 0
            ids[b, :pad] = 0
    cb = torch.randn(B * (S + 1), generator=g) if bias else None
    return prec, emb, ids, log_mask, cb


def _oracle(prec, emb, ids, log_mask, cb):
    import transrec_oracle as O
    p = prec.float().requires_grad_(True)
    e = emb.float().requires_grad_(True)
    loss = O.inbatch_softmax_loss(p, e[:, :, 0], ids, log_mask, cb)
    loss.backward()
    return float(loss.detach()), p.grad, e.grad


@pytest.mark.parametrize("B,S,D,n_items,ragged,bias", [
    (8, 20, 64, 40, True, False),        # small catalogue: many cross-user duplicates
    (13, 7, 64, 1000, True, True),       # tails in rows and candidates, debias term
    (64, 20, 64, 500, False, False),     # several row / candidate tiles
    (5, 1, 64, 6, False, False),         # S = 1: every row of a tile is another user
    (9, 33, 128, 100, True, True),       # D = 128, S > 32
])
def test_inbatch_ce_matches_oracle(B, S, D, n_items, ragged, bias):
    from adapter4rec_b200 import ops
    prec, emb, ids, log_mask, cb = _case(B, S, D, n_items, 11 * B + S, ragged, bias)
    ref_loss, ref_dp, ref_de = _oracle(prec, emb, ids, log_mask, cb)
    dev = "cuda"
    cbd = None if cb is None else cb.to(dev)
    loss, count, lse = ops.inbatch_ce_fwd(prec.to(dev), emb.to(dev), ids.to(dev), log_mask.to(dev), cbd)
    assert float(count) == float(log_mask.sum())
    assert abs(float(loss) - ref_loss) <= 1e-4 * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    go = torch.full((1,), 3.0, device=dev)
    d_prec, d_emb = ops.inbatch_ce_bwd(prec.to(dev), emb.to(dev), ids.to(dev), log_mask.to(dev), lse, count, cand_bias=cbd,
                                       grad_out=go)
    d_prec, d_emb = d_prec.float().cpu() / 3.0, d_emb.float().cpu() / 3.0
    assert float(d_emb[:, :, 1].abs().max()) == 0.0            # the sampled negatives take no part in this head
    for got, ref, what in ((d_prec, ref_dp, "d_prec"), (d_emb, ref_de, "d_emb")):
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        assert err <= 2 ** -6 * scale + 1e-7, "%s: max abs err %.3g vs scale %.3g" % (what, err, scale)
        rel = float((got - ref).norm() / ref.norm())
        assert rel <= 1e-2, "%s: relative L2 error %.3g" % (what, rel)


def test_inbatch_ce_is_deterministic_and_masks_duplicates():
    """Two runs are bit-identical (no atomics), and a batch in which every user holds the same items — everything but
    the positive is a duplicate — gives loss == log(1 + (C-1) e^{-1e4 - s_t}) == 0 exactly."""
    from adapter4rec_b200 import ops
    prec, emb, ids, log_mask, _ = _case(32, 20, 64, 300, 5, True)
    a = [t.cuda() for t in (prec, emb, ids, log_mask)]
    l1, c1, s1 = ops.inbatch_ce_fwd(*a)
    g1 = ops.inbatch_ce_bwd(*a, s1, c1)
    l2, c2, s2 = ops.inbatch_ce_fwd(*a)
    g2 = ops.inbatch_ce_bwd(*a, s2, c2)
    assert torch.equal(l1, l2) and torch.equal(s1, s2) and torch.equal(g1[0], g2[0]) and torch.equal(g1[1], g2[1])
    B, S = 6, 20
    same = torch.arange(1, S + 2).repeat(B, 1)
    prec, emb, _, _, _ = _case(B, S, 64, 10, 9, False)
    loss, _, _ = ops.inbatch_ce_fwd(prec.cuda(), emb.cuda(), same.cuda(), torch.ones(B, S).cuda())
    assert float(loss) == 0.0


def test_model_inbatch_softmax_head():
    """Model.forward with args.loss_type = 'inbatch_softmax' against the oracle's loss_from_embeddings_inbatch on the
    encoder outputs of the same model (the encoder itself is covered by tests/test_model_gpu.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import transrec_oracle as O
    from test_model_gpu import build_gpu_model, oracle_setup
    c = cases.tiny_case("lora")
    sd = cases.build_state_dict(c)
    model, _ = build_gpu_model(c, sd)
    items = cases.build_item_content(c)
    sample_items, log_mask, _ = cases.build_batch(c, items)
    rows = sample_items.view(-1, 2 * c.L)
    B, S = log_mask.shape
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(1, 12, (B, S + 1), generator=g)
    ids[:, :-1][log_mask == 0] = 0
    model.eval()
    model.args.loss_type = "inbatch_softmax"
    loss = model(rows.cuda(), log_mask.cuda(), 0, sample_items_id=ids.cuda())
    loss.backward()
    cfg, rec = oracle_setup(c)
    embs = O.bert_encoder(rows, sd, cfg, rec)
    ref = O.loss_from_embeddings_inbatch(embs, ids, log_mask, sd, rec)
    assert abs(float(loss) - float(ref)) <= 2e-2 * abs(float(ref)), (float(loss), float(ref))
    grads = [p.grad for p in model.parameters() if p.requires_grad]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads) and any(float(gr.abs().max()) > 0 for gr in grads)
