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
