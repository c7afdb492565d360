"""CPU: host logic of the unpadded token layout (adapter4rec_b200.model.bert.PackedTokens) — which padded rows are kept,
the cumulative sequence offsets, the per-token key mask and the [CLS] rows — incl. an item whose mask is all zero (the
padding item, row 0 of item_content), which must keep its L tokens with a zero mask (uniform attention, as the
reference's additive mask gives)."""
import torch

from adapter4rec_b200.model import PackedTokens


def test_packed_tokens_layout():
    mask = torch.tensor([[0, 0, 0, 0],        # padding item: kept entirely, mask stays 0
                         [1, 1, 1, 0],
                         [1, 0, 0, 0],
                         [1, 1, 1, 1]])
    p = PackedTokens(mask)
    assert p.cu_seqlens.tolist() == [0, 4, 7, 8, 12] and p.cu_seqlens.dtype == torch.int32
    assert p.token_rows.tolist() == [0, 1, 2, 3, 4, 5, 6, 8, 12, 13, 14, 15]
    assert p.cls_rows.tolist() == [0, 4, 7, 8]
    assert p.token_mask.tolist() == [0, 0, 0, 0] + [1] * 8 and p.token_mask.dtype == torch.float32
    assert p.num_tokens == 12


def test_packed_tokens_accepts_a_strided_mask_view_and_non_prefix_masks():
    rows = torch.zeros((3, 10), dtype=torch.int64)
    rows[:, 5:] = torch.tensor([[1, 1, 0, 1, 0], [1, 0, 0, 0, 0], [1, 1, 1, 1, 1]])
    p = PackedTokens(torch.narrow(rows, 1, 5, 5))          # the reference's `text_attmask` view (encoders.py:52)
    assert p.cu_seqlens.tolist() == [0, 3, 4, 9]
    assert p.token_rows.tolist() == [0, 1, 3, 5, 10, 11, 12, 13, 14]
    assert bool((p.token_mask == 1).all())


def test_unique_rows_matches_torch_unique():
    """functional.unique_rows (hash + verification) returns a valid (unique, inverse, counts) triple"""
    import torch
    from adapter4rec_b200 import functional as Fn
    g = torch.Generator().manual_seed(0)
    rows = torch.randint(0, 4, (500, 6), generator=g)
    u, inv, cnt = Fn.unique_rows(rows)
    ref = torch.unique(rows, dim=0)
    assert u.shape == ref.shape and bool((u[inv] == rows).all()) and int(cnt.sum()) == rows.shape[0]
    assert torch.equal(torch.unique(u, dim=0), ref)
    assert torch.equal(cnt, torch.bincount(inv, minlength=u.shape[0]))
