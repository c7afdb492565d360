"""-m gpu: device train-batch assembly (a4r_sample_train_batch via data_utils.BuildTrainDataset) against the oracle
restatement of BuildTrainDataset.__getitem__ and the golden output of the unmodified reference dataset.
Integer work: bit-exact given the same negatives; the sampler itself is checked through the reference's admissibility
rule, its zero pattern, reproducibility and uniformity (the random stream is counter-based, not Python's)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from test_batch_oracle import oracle_batch  # noqa: E402
import cases  # noqa: E402

pytestmark = pytest.mark.gpu


def test_gather_and_mask_bit_exact_given_the_reference_negatives():
    from adapter4rec_b200.data_utils import BuildTrainDataset
    c = cases.tiny_case("houlsby")
    gold = torch.load(os.path.join(HERE, "golden", "train_batch.pt"), weights_only=False)
    sample_items, log_mask, ids, u2seq, items = oracle_batch(c, gold["seed"])
    ds = BuildTrainDataset(u2seq, items.numpy(), c.item_num, c.S, True, device="cuda")
    out, lm = ds.batch(sorted(u2seq), neg_items=ids[:, :, 1].contiguous())
    assert out.dtype == torch.int64 and tuple(out.shape) == tuple(gold["sample_items"].shape)
    assert torch.equal(out.cpu(), gold["sample_items"]) and torch.equal(out.cpu(), sample_items)
    assert torch.equal(lm.cpu(), gold["log_mask"])


def _ragged_users(n_users, item_num, S, seed):
    g = torch.Generator().manual_seed(seed)
    u2seq = {}
    for u in range(n_users):
        n = int(torch.randint(2, S + 2, (1,), generator=g))
        u2seq[u] = (torch.randperm(item_num, generator=g)[:n] + 1).tolist()
    return u2seq


@pytest.mark.parametrize("S,item_num,L", [(20, 300, 30), (40, 90, 5), (5, 8, 4)])
def test_sampler_properties(S, item_num, L):
    """ragged users; S+1 = 41 exercises two 32-lane chunks; item_num = 8 with S+1 = 6 forces many rejections"""
    from adapter4rec_b200.data_utils import BuildTrainDataset
    g = torch.Generator().manual_seed(7)
    content = torch.randint(0, 1000, (item_num + 1, 2 * L), generator=g)
    content[0] = 0
    u2seq = _ragged_users(64, item_num, S, 11)
    ds = BuildTrainDataset(u2seq, content, item_num, S, True, device="cuda", seed=99)
    users = sorted(u2seq)
    out, lm = ds.batch(users, check=True)
    neg = ds.last_negatives.cpu()
    seqs = ds.seqs.cpu()
    real = seqs != 0
    # zero pattern of dataset.py:41: negatives exactly under real, non-last slots
    want = real.clone()
    want[:, -1] = False
    assert torch.equal(neg != 0, want)
    # admissibility (dataset.py:37-39): in [1, item_num] and not in the user's own sequence
    assert int(neg[want].min()) >= 1 and int(neg[want].max()) <= item_num
    assert not bool((neg.unsqueeze(2) == seqs.unsqueeze(1))[want].any())
    # gather (dataset.py:46) and log_mask (:31) bit-exact
    assert torch.equal(out.cpu(), content[torch.stack([seqs, neg], 2)])
    assert torch.equal(lm.cpu(), real[:, :-1].float())
    # reproducible from (seed, batch counter); the next batch differs
    ds2 = BuildTrainDataset(u2seq, content, item_num, S, True, device="cuda", seed=99)
    out2, _ = ds2.batch(users)
    assert torch.equal(out2, out)
    ds2.batch(users)
    assert not torch.equal(ds2.last_negatives.cpu(), neg)


def test_sampler_is_uniform_over_admissible_items():
    """one user, many draws: every admissible item within 5 sigma of the uniform expectation, inadmissible never"""
    from adapter4rec_b200 import ops
    item_num, S1 = 50, 21
    seq = torch.arange(1, S1 + 1, dtype=torch.int64)              # items 1..21 are the user's own
    B = 20000
    seqs = seq.unsqueeze(0).repeat(B, 1).cuda()
    content = torch.zeros((item_num + 1, 2), dtype=torch.int64).cuda()
    _, _, neg, fail = ops.sample_train_batch(seqs, content, item_num, 5, 0)
    assert int(fail.item()) == 0
    draws = neg[:, :-1].flatten().cpu()
    counts = torch.bincount(draws, minlength=item_num + 1).double()
    assert counts[:S1 + 1].sum() == 0
    n, k = draws.numel(), item_num - S1
    exp, sd = n / k, (n * (1 / k) * (1 - 1 / k)) ** 0.5
    assert bool(((counts[S1 + 1:] - exp).abs() <= 5 * sd).all())


def test_impossible_sampling_is_reported():
    from adapter4rec_b200 import ops
    seqs = torch.tensor([[1, 2, 3]], dtype=torch.int64).cuda()     # item_num = 3: every item is in the sequence
    content = torch.zeros((4, 2), dtype=torch.int64).cuda()
    _, _, _, fail = ops.sample_train_batch(seqs, content, 3, 1, 0)
    assert int(fail.item()) == 1


def test_empty_batch():
    from adapter4rec_b200 import ops
    seqs = torch.zeros((0, 6), dtype=torch.int64).cuda()
    content = torch.zeros((4, 8), dtype=torch.int64).cuda()
    out, lm, neg, fail = ops.sample_train_batch(seqs, content, 3, 1, 0)
    assert out.shape == (0, 6, 2, 8) and lm.shape == (0, 5) and int(fail.item()) == 0
