"""Generates tests/golden/train_batch.pt from the UNMODIFIED reference BuildTrainDataset
(/root/reference/Downstream/Text/data_utils/dataset.py:10-49) under random.seed(2024):
    python tests/golden/make_golden_batch.py
The fixture holds, for 7 users of the 'houlsby' tiny case, what __getitem__ returned (token rows and log_mask)."""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/Downstream/Text")

import cases  # noqa: E402
from data_utils.dataset import BuildTrainDataset  # noqa: E402


def users_for(c):
    seqs, _ = cases.build_eval_users(c)          # 9 sequences of 2..S+1 distinct items
    return {u: s for u, s in enumerate(seqs[:7])}


def main():
    c = cases.tiny_case("houlsby")
    items = cases.build_item_content(c).numpy()
    u2seq = users_for(c)
    ds = BuildTrainDataset(u2seq=u2seq, item_content=items, item_num=c.item_num, max_seq_len=c.S, use_modal=True)
    random.seed(2024)
    out = [ds[u] for u in sorted(u2seq)]
    torch.save({"sample_items": torch.stack([o[0] for o in out]), "log_mask": torch.stack([o[1] for o in out]),
                "seed": 2024}, os.path.join(HERE, "train_batch.pt"))
    print("wrote train_batch.pt", torch.stack([o[0] for o in out]).shape)


if __name__ == "__main__":
    main()
