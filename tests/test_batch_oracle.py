"""CPU: the oracle's train-sample restatement (oracle/transrec_oracle.py:train_sample) against what the unmodified
reference BuildTrainDataset.__getitem__ returned under the same Python RNG seed (tests/golden/train_batch.pt)."""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import cases  # noqa: E402
import transrec_oracle as O  # noqa: E402


def golden_users(c):
    seqs, _ = cases.build_eval_users(c)
    return {u: s for u, s in enumerate(seqs[:7])}


def oracle_batch(c, seed):
    items = cases.build_item_content(c)
    u2seq = golden_users(c)
    rng = random.Random(seed)
    out = [O.train_sample(u2seq[u], items, c.item_num, c.S, rng) for u in sorted(u2seq)]
    return (torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out]), torch.stack([o[2] for o in out]), u2seq, items)


def test_train_sample_is_bit_exact_with_the_reference_dataset():
    c = cases.tiny_case("houlsby")
    gold = torch.load(os.path.join(HERE, "golden", "train_batch.pt"), weights_only=False)
    sample_items, log_mask, ids, u2seq, _ = oracle_batch(c, gold["seed"])
    assert torch.equal(sample_items, gold["sample_items"])
    assert torch.equal(log_mask, gold["log_mask"])
    # structure of the id tensor (dataset.py:34-43): left padding, negatives only under real non-last slots, never in seq
    for r, u in enumerate(sorted(u2seq)):
        seq = u2seq[u]
        head = c.S + 1 - len(seq)
        assert ids[r, :head].abs().sum() == 0 and ids[r, -1, 1] == 0
        assert ids[r, head:, 0].tolist() == seq
        negs = ids[r, head:-1, 1].tolist()
        assert all(1 <= n <= c.item_num and n not in seq for n in negs)
