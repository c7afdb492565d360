"""-m gpu: the full-ranking evaluator (K10-K13) through the C ABI against the CPU oracle and the reference goldens.
Integer results (top-k ids, HR, NDCG) must be bit-exact under the total order (score desc, id asc)."""
import logging
import math
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
BF16 = torch.bfloat16


def exact_embeddings(n, d, seed):
    """entries in {-2..2} * 2^-3: every dot product is exact in fp32 whatever the accumulation order"""
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(-2, 3, (n, d), generator=g).float() * 0.125)


def cpu_topk(users, items, hist, k, id_base=0):
    s = users.double() @ items.double().t()
    out = []
    for u in range(users.shape[0]):
        row = s[u].clone()
        ids = torch.arange(items.shape[0]) + id_base
        bad = torch.isin(ids, torch.as_tensor(hist[u])) | (ids == 0)
        row[bad] = -float("inf")
        order = sorted(range(items.shape[0]), key=lambda j: (-row[j].item(), j))[:k]
        out.append([(int(ids[j]), float(row[j])) for j in order if row[j] > -float("inf")])
    return out


@pytest.mark.parametrize("U,I,d,hl", [(5, 300, 64, 4), (200, 5000, 64, 21), (130, 1000, 768, 8), (1, 40, 64, 1),
                                       (300, 70000, 64, 20)])
def test_score_topk_bit_exact(U, I, d, hl):
    from adapter4rec_b200 import ops
    users, items = exact_embeddings(U, d, 1), exact_embeddings(I, d, 2)
    g = torch.Generator().manual_seed(3)
    hist = torch.randint(0, I, (U, hl), generator=g).int()
    sc, ids = ops.score_topk(users.to(BF16).cuda(), items.to(BF16).cuda(), id_base=0, history=hist.cuda(), k=10)
    msc, mid, _, _ = ops.topk_merge(sc, ids)
    check_users = range(U) if I <= 5000 else range(0, U, 37)
    ref = cpu_topk(users[list(check_users)], items, hist[list(check_users)].tolist(), 10)
    for r, u in enumerate(check_users):
        got = [(int(i), float(s)) for i, s in zip(mid[u].tolist(), msc[u].tolist()) if s > -float("inf")]
        assert got == ref[r], "user %d: %s vs %s" % (u, got[:4], ref[r][:4])


def _gpu_exact_topk(users, items, hist, id_base, k=10, chunk=1 << 19):
    """Chunked torch reference ON THE GPU for large shards: fp32 scores (exact for the {-2..2}/8 embeddings: every partial
    sum is a multiple of 1/64 below 2^24/64), history ids and id 0 removed, then the top-k under the total order (score desc,
    id asc) through ONE sortable int64 key = (64 * score) << 32 | (2^32 - 1 - id) — torch.topk of that key is exact and
    tie-free.  Returns ids [U, k] (int64) and scores [U, k] (f32)."""
    U = users.shape[0]
    best = torch.full((U, k), torch.iinfo(torch.int64).min, dtype=torch.int64, device=users.device)
    uf = users.float()
    for lo in range(0, items.shape[0], chunk):
        hi = min(items.shape[0], lo + chunk)
        sc = uf @ items[lo:hi].float().t()
        ids = torch.arange(lo, hi, device=users.device, dtype=torch.int64) + id_base
        key = ((sc * 64.0).round().to(torch.int64) << 32) | (0xFFFFFFFF - ids)[None, :]
        dead = (ids == 0)[None, :] | (ids[None, None, :] == hist.long()[:, :, None]).any(1)
        key = torch.where(dead, torch.full_like(key, torch.iinfo(torch.int64).min), key)
        best = torch.topk(torch.cat([best, key], 1), k, dim=1).values
    ids = 0xFFFFFFFF - (best & 0xFFFFFFFF)
    return ids, (best >> 32).float() / 64.0


@pytest.mark.parametrize("rows,id_base", [(3_000_001, 0), (3_000_000, 7_000_000)])
def test_score_topk_bit_exact_at_c5_shard_sizes(rows, id_base):
    """a4r_score_topk at the size of a real C5 shard (metrics.py:105-111 at I = 10 M, d = 768): >= 3 M rows x 768 bf16 =
    4.6 GB, i.e. row byte offsets far past 2^31, and item ids past 2^31 / 1536 (id_base 7 M: the last ranks of an 8-way
    shard).  Exactly representable embeddings make every score exact, and with 3 M items the top-10 of every user is FULL of
    exact ties — the (score desc, id asc) order is exercised across item splits and tile boundaries.  64 users of 300 are
    compared id for id and score for score with the chunked torch reference."""
    from adapter4rec_b200 import ops
    d, U, hl = 768, 300, 20
    g = torch.Generator(device="cuda").manual_seed(17 + id_base)
    items = torch.empty((rows, d), dtype=BF16, device="cuda")
    for lo in range(0, rows, 1 << 20):
        hi = min(rows, lo + (1 << 20))
        items[lo:hi] = (torch.randint(-2, 3, (hi - lo, d), generator=g, device="cuda").float() * 0.125).to(BF16)
    users = (torch.randint(-2, 3, (U, d), generator=g, device="cuda").float() * 0.125).to(BF16)
    hist = torch.randint(id_base, id_base + rows, (U, hl), generator=g, device="cuda", dtype=torch.int64).int()
    sc, ids = ops.score_topk(users, items, id_base=id_base, history=hist, k=10)
    msc, mid, _, _ = ops.topk_merge(sc, ids)
    # plant each sampled user's current best item into that user's history: it must disappear from the list
    sample = torch.arange(0, U, 5, device="cuda")[:64]
    ref_ids, ref_sc = _gpu_exact_topk(users[sample], items, hist[sample], id_base)
    assert torch.equal(mid[sample].long(), ref_ids), "top-10 ids differ"
    assert torch.equal(msc[sample], ref_sc), "top-10 scores differ"
    assert int(mid[sample].min()) >= max(1, id_base) and int(mid[sample].max()) < id_base + rows
    # ties must actually occur inside the lists for the tie-break to have been tested
    assert bool((ref_sc[:, 1:] == ref_sc[:, :-1]).any())
    hist2 = hist.clone()
    hist2[:, 0] = mid[:, 0]
    sc2, ids2 = ops.score_topk(users, items, id_base=id_base, history=hist2, k=10)
    _, mid2, _, _ = ops.topk_merge(sc2, ids2)
    assert not bool((mid2 == mid[:, :1]).any(1).any()), "an id placed in the history must leave the list"
    ref2, _ = _gpu_exact_topk(users[sample], items, hist2[sample], id_base)
    assert torch.equal(mid2[sample].long(), ref2)


def test_sharded_merge_equals_unsharded():
    """item table split into 3 id ranges (as 3 ranks would hold them): per-shard lists merged == single-shard result"""
    from adapter4rec_b200 import ops
    U, I, d = 150, 9001, 64
    users, items = exact_embeddings(U, d, 4).to(BF16).cuda(), exact_embeddings(I, d, 5).to(BF16).cuda()
    g = torch.Generator().manual_seed(6)
    hist = torch.randint(0, I, (U, 10), generator=g).int().cuda()
    tgt = torch.randint(1, I, (U,), generator=g).int().cuda()
    sc, ids = ops.score_topk(users, items, history=hist)
    full = ops.topk_merge(sc, ids, target=tgt)
    parts_s, parts_i = [], []
    for lo, hi in ((0, 3000), (3000, 6000), (6000, I)):
        s, i = ops.score_topk(users, items[lo:hi].contiguous(), id_base=lo, history=hist)
        ms, mi, _, _ = ops.topk_merge(s, i)
        parts_s.append(ms)
        parts_i.append(mi)
    merged = ops.topk_merge(torch.stack(parts_s).contiguous(), torch.stack(parts_i).contiguous(), target=tgt)
    for a, b in zip(full, merged):
        assert torch.equal(a, b)
    # HR / NDCG definition
    hit, ndcg = full[2].cpu(), full[3].cpu()
    for u in range(U):
        lst = full[1][u].tolist()
        if int(tgt[u]) in lst:
            assert hit[u] == 1 and abs(float(ndcg[u]) - 1 / math.log2(lst.index(int(tgt[u])) + 2)) < 1e-6
        else:
            assert hit[u] == 0 and ndcg[u] == 0


def test_gather_rows():
    from adapter4rec_b200 import ops
    t = torch.randn(1000, 64).to(BF16).cuda()
    ids = torch.randint(0, 1000, (37, 20)).cuda()
    assert torch.equal(ops.gather_rows(t, ids), t[ids])


@pytest.mark.parametrize("kind", ["base", "houlsby", "lora", "prompt_cpc"])
def test_eval_model_matches_reference(kind):
    """get_item_embeddings + eval_model (reference signatures) vs the golden HR@10 of the unmodified reference and the
    per-user oracle values.  Scores are bf16-level close, so a user whose target sits within tolerance of the rank-10
    boundary may flip; the test requires exact agreement for every user whose oracle margin exceeds the tolerance."""
    from test_model_gpu import build_gpu_model, oracle_setup
    from adapter4rec_b200.data_utils import eval_model, get_item_embeddings
    c = cases.tiny_case(kind)
    sd = cases.build_state_dict(c)
    gold = torch.load(os.path.join(HERE, "golden", "transrec_%s.pt" % kind), weights_only=False)
    model, args = build_gpu_model(c, sd)
    items = cases.build_item_content(c)
    seqs, hist = cases.build_eval_users(c)
    table = get_item_embeddings(model, items.numpy(), 16, args, True, 0)
    emb = table.shard.float().cpu()
    import parity_util as P
    assert float((emb[1:] - gold["item_emb"][1:]).abs().max()) <= P.bound(P.measured(), "text/" + kind, "emb_max_abs") + 1e-5
    log = logging.getLogger("eval_test")
    hit10 = eval_model(model, [torch.LongTensor(h) for h in hist], {i: s for i, s in enumerate(seqs)}, table, 4, args,
                       c.item_num, log, "test", 0)
    # per-user results of the CUDA evaluator against the oracle's, user by user.  A user is "decided" when the oracle's
    # margin between the target's score and the rank-10 boundary exceeds the bf16 score tolerance; every decided user must
    # agree exactly (hit AND ndcg), whatever the other users do.
    from adapter4rec_b200.data_utils.metrics import build_eval_arrays, eval_arrays
    tok, mask, tgt, hs = build_eval_arrays({i: s for i, s in enumerate(seqs)}, [torch.LongTensor(h) for h in hist],
                                           args.max_seq_len)
    hit, ndcg, top_ids = eval_arrays(model, torch.from_numpy(tok), torch.from_numpy(mask), torch.from_numpy(tgt),
                                     torch.from_numpy(hs), table, 4)
    hit, ndcg = hit.float().cpu(), ndcg.float().cpu()
    assert abs(float(hit.mean()) - hit10) < 1e-6
    cfg, rec = oracle_setup(c)
    u, _, _ = O.eval_user_vectors(seqs, gold["item_emb"], sd, rec)
    scores = u @ gold["item_emb"].t()
    SCORE_TOL = 5e-2        # bf16 user vector x bf16 table, D = 64, |score| <= ~4: measured max |ds| 2e-2
    n_decided, all_decided, exp_hits = 0, True, []
    for b in range(len(seqs)):
        s = scores[b].clone()
        s[torch.tensor(hist[b])] = -float("inf")
        s[0] = -float("inf")
        st = float(s[seqs[b][-1]])
        srt = torch.sort(s, descending=True).values
        ehit, endcg = O.rank_metrics(scores[b], hist[b], seqs[b][-1])
        exp_hits.append(ehit)
        # hit decided: the target is clear of the rank-10 boundary; ndcg decided: also clear of its rank neighbours
        boundary = float(srt[9]) if st < float(srt[9]) else float(srt[10])
        hit_decided = abs(st - boundary) > SCORE_TOL
        others = s.clone()
        others[seqs[b][-1]] = -float("inf")
        rank_decided = float((others[torch.isfinite(others)] - st).abs().min()) > SCORE_TOL
        all_decided &= hit_decided
        if hit_decided:
            n_decided += 1
            assert float(hit[b]) == ehit, "user %d: hit %s vs oracle %s" % (b, float(hit[b]), ehit)
            if rank_decided or ehit == 0.0:
                assert abs(float(ndcg[b]) - endcg) < 1e-6, "user %d: ndcg %s vs oracle %s" % (b, float(ndcg[b]), endcg)
    assert n_decided >= (len(seqs) + 1) // 2, "the case must decide most users (%d of %d)" % (n_decided, len(seqs))
    if all_decided:
        assert abs(hit10 - gold["eval_hit10_mean"]) < 1e-6
        assert abs(hit10 - sum(exp_hits) / len(exp_hits)) < 1e-6


def test_sharded_evaluator_under_nccl_two_ranks():
    """tools/dist_check_gpu.py under torchrun with 2 ranks (needs 2 visible GPUs, else skipped): the item-sharded evaluator
    (per-rank item table, users of a block split across the ranks for the encoder, all-gather of the user vectors and of the
    partial top-10 lists, merge) returns bit-identical ids / HR / NDCG to the unsharded run, and two data-parallel steps leave
    bit-identical parameters on both ranks."""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "dist_check_gpu.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])
