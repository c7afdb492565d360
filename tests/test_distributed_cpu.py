"""CPU (-m "not gpu"): the N > 1 host logic under a real 2-process gloo group (127.0.0.1): item-id sharding, the
ownership-masked gather + sum all-reduce (exactness), the all-gather + merge of per-shard top-k lists against the
unsharded oracle, and the single flat-gradient all-reduce of the trainer."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _merge_lists(lists, k):
    """reference merge of sorted (score, id) lists under (score desc, id asc) — what a4r_topk_merge implements"""
    allc = [c for l in lists for c in l]
    return sorted(allc, key=lambda c: (-c[0], c[1]))[:k]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import transrec_oracle as O
        from adapter4rec_b200.data_utils.metrics import ItemTable, shard_range
        from adapter4rec_b200.trainer import FlatAdamTrainer
        g = torch.Generator().manual_seed(5)
        I, d, U, k = 1001, 16, 12, 10
        items = (torch.randint(-2, 3, (I, d), generator=g).float() * 0.125)      # exact dot products
        users = (torch.randint(-2, 3, (U, d), generator=g).float() * 0.125)
        hist = [torch.randint(1, I, (5,), generator=g).tolist() for _ in range(U)]
        lo, hi = shard_range(I, rank, world)
        # (1) shards tile [0, I) without overlap
        bounds = [None] * world
        dist.all_gather_object(bounds, (lo, hi))
        assert bounds[0][0] == 0 and bounds[-1][1] == I and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        # (2) ownership-masked gather + sum all-reduce reproduces a plain gather exactly (bf16)
        table = ItemTable(items[lo:hi], lo, I, rank, world)
        ids = torch.randint(0, I, (U, 7), generator=g)
        part = table.table[table.local_index(ids)].clone()          # torch gather stands in for a4r_gather_rows on CPU
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        assert torch.equal(part, items.to(torch.bfloat16)[ids])
        # (3) per-shard top-k, all-gather, merge == unsharded oracle top-k ids
        scores = users @ items[lo:hi].t()
        mine = []
        for u in range(U):
            cand = []
            for j in range(hi - lo):
                gid = lo + j
                if gid != 0 and gid not in hist[u]:
                    cand.append((float(scores[u, j]), gid))
            mine.append(sorted(cand, key=lambda c: (-c[0], c[1]))[:k])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        full_scores = users @ items.t()
        for u in range(U):
            merged = [c[1] for c in _merge_lists([gathered[r][u] for r in range(world)], k)]
            assert merged == O.topk_ids(full_scores[u], hist[u], k)
        # (4) trainer: one flat all-reduce sums the ranks' gradients; parameter views stay attached to the flat buffer
        lin = torch.nn.Linear(4, 3)
        lin.weight.data.fill_(1.0)
        lin.bias.data.fill_(2.0)
        model = torch.nn.ModuleDict({"bert_encoder": torch.nn.ModuleDict({"adapter": lin})})
        tr = FlatAdamTrainer(model, 1e-3, 1e-3, 1e-3, 1e-3)
        assert tr.world == world and tr.num_trainable == 15
        assert lin.weight.data_ptr() == tr.flat_param.data_ptr() or lin.bias.data_ptr() == tr.flat_param.data_ptr()
        tr.zero_grad()
        lin.weight.grad += float(rank + 1)
        lin.bias.grad += 10.0 * (rank + 1)
        tr.reduce_gradients()
        assert torch.all(lin.weight.grad == 3.0) and torch.all(lin.bias.grad == 30.0)
        assert float(tr.flat_grad.sum()) == 12 * 3.0 + 3 * 30.0
        # (5) bucketed reduction issued from gradient hooks during the LAST pass of a multi-pass step: every bucket (incl.
        # one whose parameter receives no gradient) ends up holding the sum over ranks of the accumulated local gradients
        class Toy(torch.nn.Module):
            def __init__(self):
                super().__init__()
                torch.manual_seed(3)
                self.unused, self.a, self.b = torch.nn.Linear(2, 2), torch.nn.Linear(6, 5), torch.nn.Linear(5, 1)

            def forward(self, rows, log_mask, dev):
                return (self.b(torch.tanh(self.a(rows))).view(log_mask.shape[0], -1).mean(1) * log_mask.sum(1)).sum() / 7.0
        toy = Toy()
        tr2 = FlatAdamTrainer(toy, 1e-3, 1e-3, 1e-3, 1e-3, users_per_pass=3, bucket_bytes=64, overlap=True)
        assert len(tr2.buckets) >= 3 and sum(b[1] for b in tr2.buckets) == tr2.num_trainable
        gen = torch.Generator().manual_seed(100 + rank)
        rows, lm = torch.randn(8 * 2, 6, generator=gen), (torch.rand(8, 4, generator=gen) < 0.7).float()
        tr2.zero_grad()
        tr2.forward_backward(rows, lm)
        assert tr2._live and len(tr2._works) > 0      # some buckets left while the backward was running
        tr2.reduce_gradients()
        got = tr2.flat_grad.clone()
        ref_model = Toy()
        tr3 = FlatAdamTrainer(ref_model, 1e-3, 1e-3, 1e-3, 1e-3, users_per_pass=3, overlap=False)
        tr3.zero_grad()
        tr3.forward_backward(rows, lm)
        want = tr3.flat_grad.clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        assert torch.allclose(got, want, rtol=0, atol=1e-6), float((got - want).abs().max())
        assert float(got.abs().sum()) > 0 and not tr2._live and not tr2._works
        # overlap=None: blocking inside one node, bucketed when the job spans several (LOCAL_WORLD_SIZE < WORLD_SIZE)
        os.environ["LOCAL_WORLD_SIZE"] = str(world)
        assert len(FlatAdamTrainer(Toy(), 1e-3, 1e-3, 1e-3, 1e-3, bucket_bytes=64).buckets) == 0
        os.environ["LOCAL_WORLD_SIZE"] = "1"
        assert len(FlatAdamTrainer(Toy(), 1e-3, 1e-3, 1e-3, 1e-3, bucket_bytes=64).buckets) >= 3
        os.environ.pop("LOCAL_WORLD_SIZE")
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback
        ret[rank] = "FAIL: " + traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    world = 2
    port = 29600 + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


def test_parameter_groups_follow_reference_names():
    """run.py:505-523: 4 groups from ('bert_encoder' in name) x ('adapter' in name or 'lora' in name)."""
    from adapter4rec_b200.trainer import group_parameters
    m = torch.nn.ModuleDict({
        "bert_encoder": torch.nn.ModuleDict({"adapter": torch.nn.Linear(2, 2), "query": torch.nn.Linear(2, 2)}),
        "user_encoder": torch.nn.ModuleDict({"lora_x": torch.nn.Linear(2, 2), "fc": torch.nn.Linear(2, 2)})})
    g = group_parameters(m)
    assert [n for n, _ in g["adapter_bert"]] == ["bert_encoder.adapter.weight", "bert_encoder.adapter.bias"]
    assert [n for n, _ in g["bert"]] == ["bert_encoder.query.weight", "bert_encoder.query.bias"]
    assert [n for n, _ in g["adapter_recsys"]] == ["user_encoder.lora_x.weight", "user_encoder.lora_x.bias"]
    assert [n for n, _ in g["recsys"]] == ["user_encoder.fc.weight", "user_encoder.fc.bias"]


def test_eval_arrays_host_restatement():
    """build_eval_arrays == BuildEvalDataset.__getitem__ (dataset.py:65-78) for every user"""
    from adapter4rec_b200.data_utils.metrics import build_eval_arrays
    seqs = {0: [5, 9, 2], 1: [7, 1], 2: [3, 4, 6, 8, 10, 11]}
    hist = {u: torch.LongTensor(s[:-1]) for u, s in seqs.items()}
    tok, mask, tgt, h = build_eval_arrays(seqs, hist, 5)
    assert tok.tolist() == [[0, 0, 0, 5, 9], [0, 0, 0, 0, 7], [3, 4, 6, 8, 10]]
    assert mask.tolist() == [[0, 0, 0, 1, 1], [0, 0, 0, 0, 1], [1, 1, 1, 1, 1]]
    assert tgt.tolist() == [2, 1, 11]
    assert h.tolist() == [[5, 9, 0, 0, 0], [7, 0, 0, 0, 0], [3, 4, 6, 8, 10]]


@pytest.mark.parametrize("n_users,world,batch", [(129, 2, 64), (7, 4, 2), (64, 2, 64), (1, 2, 8), (1000, 8, 32)])
def test_rank_shard_gives_every_rank_the_same_number_of_steps(n_users, world, batch):
    """run.rank_shard == DistributedSampler's padded round-robin split (Downstream/Text/run.py:347): with 129 users on 2
    ranks at batch 64 an unpadded split gives 2 and 1 optimizer steps — one unmatched all-reduce; padded, both take 2."""
    from adapter4rec_b200.run import rank_shard
    users = list(range(100, 100 + n_users))
    shards = [rank_shard(users, r, world) for r in range(world)]
    per = (n_users + world - 1) // world
    assert all(len(s) == per for s in shards)
    steps = {(len(s) + batch - 1) // batch for s in shards}
    assert len(steps) == 1
    seen = [u for s in shards for u in s]
    assert set(seen) == set(users) and len(seen) - n_users < world          # everyone covered, < world repeats
    # identical to torch's sampler on the same (unshuffled) list
    from torch.utils.data import DistributedSampler
    for r in range(world):
        ref = list(DistributedSampler(users, num_replicas=world, rank=r, shuffle=False))
        assert [users[i] for i in ref] == shards[r]


def test_build_model_applies_the_reference_freeze_and_refuses_missing_checkpoints(tmp_path):
    """run.py:302-320 (prefix + pooler freeze before the Model is built, word_embedding_dim from the body's name) and
    :374-381 (--pretrained_model_name is loaded before the surgery; a missing file is an error, not a silent skip)."""
    from adapter4rec_b200 import run
    from adapter4rec_b200.model import TextConfigLite
    from adapter4rec_b200.parameters import parse_args
    cfg = TextConfigLite(vocab_size=100, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                         intermediate_size=256, max_position_embeddings=32)
    base = ["--embedding_dim", "64", "--bert_model_load", "bert_tiny", "--word_embedding_dim", "999", "--max_seq_len", "5",
            "--num_words_title", "8", "--adapter_type", "None", "--adding_adapter_to", "None", "--fine_tune_to", "all"]
    args = parse_args(base + ["--freeze_paras_before", "21", "--pretrained_model_name", "None"])
    model = run.build_model(args, 50, "cpu", cfg)
    assert args.word_embedding_dim == 128
    bert = model.bert_encoder.text_encoders.title.bert_model
    flags = [p.requires_grad for _, p in bert.named_parameters()]
    names = [n for n, _ in bert.named_parameters()]
    assert not any(flags[:21]) and all(f for i, f in enumerate(flags[21:], 21) if i not in (37, 38))
    assert names[37].startswith("pooler") and names[38].startswith("pooler") and not flags[37] and not flags[38]
    assert all(p.requires_grad for p in model.user_encoder.parameters())
    # checkpoint round trip through --pretrained_model_name
    torch.save({"model_state_dict": model.state_dict()}, tmp_path / "epoch-3.pt")
    args2 = parse_args(base + ["--pretrained_model_dir", str(tmp_path), "--pretrained_model_name", "epoch-3"])
    model2 = run.build_model(args2, 50, "cpu", cfg)
    for (n, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), n
    args3 = parse_args(base + ["--pretrained_model_dir", str(tmp_path), "--pretrained_model_name", "epoch-15"])
    with pytest.raises(FileNotFoundError):
        run.build_model(args3, 50, "cpu", cfg)


def _four_group_module():
    torch.manual_seed(3)
    return torch.nn.ModuleDict({
        "bert_encoder": torch.nn.ModuleDict({"adapter": torch.nn.Linear(3, 2), "query": torch.nn.Linear(3, 4)}),
        "user_encoder": torch.nn.ModuleDict({"lora_x": torch.nn.Linear(2, 2, bias=False), "fc": torch.nn.Linear(4, 3)})})


def _reference_optimizer(m, lr, fine_tune_lr, adapter_bert_lr, adapter_sasrec_lr):
    """the optimizer Downstream/Text/run.py:505-529 builds (same name tests, same group order)"""
    g = {"bert": [], "recsys": [], "adapter_bert": [], "adapter_recsys": []}
    for name, p in m.named_parameters():
        if p.requires_grad:
            ad = "adapter" in name or "lora" in name
            g[("adapter_bert" if ad else "bert") if 'bert_encoder' in name else ("adapter_recsys" if ad else "recsys")].append(p)
    return torch.optim.Adam([{'params': g["bert"], 'lr': fine_tune_lr}, {'params': g["recsys"], 'lr': lr},
                             {'params': g["adapter_bert"], 'lr': adapter_bert_lr},
                             {'params': g["adapter_recsys"], 'lr': adapter_sasrec_lr}])


def test_optimizer_state_is_interchangeable_with_the_reference_checkpoint_format(tmp_path):
    """SURVEY.md 8f-4 "checkpoint I/O compatibility": data_utils/utils.py:109-115 stores torch.optim.Adam's state_dict under
    'optimizer'.  (1) a state written by the reference's optimizer loads into FlatAdamTrainer (moments land at the right
    offsets of the flat buffers, the step count carries over); (2) FlatAdamTrainer.state_dict() loads into the reference's
    optimizer through torch's own load_state_dict and the two then take the same next step; (3) both directions survive
    torch.save / torch.load; (4) mismatched trainable sets are refused."""
    from adapter4rec_b200 import functional as Fn
    from adapter4rec_b200.trainer import FlatAdamTrainer
    lrs = dict(lr=1e-3, fine_tune_lr=2e-3, adapter_bert_lr=3e-3, adapter_sasrec_lr=4e-3)
    ref_model = _four_group_module()
    opt = _reference_optimizer(ref_model, **lrs)
    g = torch.Generator().manual_seed(9)
    for _ in range(3):
        for p in ref_model.parameters():
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
    torch.save({"optimizer": opt.state_dict()}, tmp_path / "ref.pt")
    mine = _four_group_module()
    mine.load_state_dict(ref_model.state_dict())
    tr = FlatAdamTrainer(mine, lrs["lr"], lrs["fine_tune_lr"], lrs["adapter_bert_lr"], lrs["adapter_sasrec_lr"])
    tr.load_state_dict(torch.load(tmp_path / "ref.pt", weights_only=False)["optimizer"])                      # (1)
    assert tr.step_count == 3
    ref_params = [p for grp in opt.param_groups for p in grp["params"]]
    by_name = dict(mine.named_parameters())
    assert [tuple(by_name[n].shape) for n, _, _ in tr.names] == [tuple(p.shape) for p in ref_params]     # same order
    for (n, off, k), p in zip(tr.names, ref_params):
        assert torch.equal(tr.exp_avg[off:off + k], opt.state[p]["exp_avg"].reshape(-1)), n
        assert torch.equal(tr.exp_avg_sq[off:off + k], opt.state[p]["exp_avg_sq"].reshape(-1)), n
    Fn.DropoutState.seed, Fn.DropoutState.counter = 4242, 17
    torch.save({"optimizer": tr.state_dict()}, tmp_path / "mine.pt")                                     # (2) + (3)
    sd = torch.load(tmp_path / "mine.pt", weights_only=False)["optimizer"]
    assert [grp["lr"] for grp in sd["param_groups"]] == [2e-3, 1e-3, 3e-3, 4e-3]
    assert [len(grp["params"]) for grp in sd["param_groups"]] == [2, 2, 2, 1]
    other = _four_group_module()
    other.load_state_dict(ref_model.state_dict())
    opt2 = _reference_optimizer(other, **lrs)
    opt2.load_state_dict(sd)                                                  # torch's own loader takes it
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    for p in ref_model.parameters():
        p.grad = torch.randn(p.shape, generator=g1)
    for p in other.parameters():
        p.grad = torch.randn(p.shape, generator=g2)
    opt.step(), opt2.step()
    for a, b in zip(ref_model.parameters(), other.parameters()):
        assert torch.equal(a, b)
    Fn.DropoutState.seed, Fn.DropoutState.counter = 1, 1
    sd = torch.load(tmp_path / "mine.pt", weights_only=False)["optimizer"]    # (opt2 adopted and advanced sd's step tensors)
    tr.load_state_dict(sd)
    assert (Fn.DropoutState.seed, Fn.DropoutState.counter) == (4242, 17) and tr.step_count == 3
    # a fresh trainer writes an empty state (torch does the same before the first step) and reads it back
    fresh = FlatAdamTrainer(_four_group_module(), 1e-3, 1e-3, 1e-3, 1e-3)
    assert fresh.state_dict()["state"] == {}
    fresh.load_state_dict(fresh.state_dict())
    assert fresh.step_count == 0
    # (4) another trainable set
    small = _four_group_module()
    small["user_encoder"]["fc"].weight.requires_grad = False
    with pytest.raises(ValueError):
        FlatAdamTrainer(small, 1e-3, 1e-3, 1e-3, 1e-3).load_state_dict(opt.state_dict())
    # the flat format this package wrote before still loads
    legacy = {"step": 7, "exp_avg": tr.exp_avg.clone() * 2, "exp_avg_sq": tr.exp_avg_sq.clone(), "names": list(tr.names)}
    tr.load_state_dict(legacy)
    assert tr.step_count == 7
