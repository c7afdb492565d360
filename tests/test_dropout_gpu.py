"""-m gpu: counter-based dropout (train-mode semantics of nn.Dropout in the reference's BERT / SASRec blocks).
The RNG is restated on the host (splitmix64) so the kernels can be checked against torch with the SAME mask."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
M64 = (1 << 64) - 1


def rng64(seed, counter):
    """host restatement of rng64() in csrc/a4r_common.cuh (splitmix64 finaliser)"""
    z = (counter.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def keep_mask(seed, offset, n, p):
    """mask of n consecutive elements starting at counter `offset` (4 elements per counter)"""
    thr = int(p * 65536.0 + 0.5)
    with np.errstate(over="ignore"):
        bits = rng64(seed, np.arange((n + 3) // 4, dtype=np.uint64) + np.uint64(offset))
    lanes = np.stack([(bits >> np.uint64(16 * k)) & np.uint64(0xFFFF) for k in range(4)], 1).reshape(-1)[:n]
    return torch.from_numpy((lanes >= thr).astype(np.float32)), 65536.0 / (65536 - thr)


def test_dropout_kernel_matches_host_rng_and_is_unbiased():
    from adapter4rec_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((4096, 768), generator=g, device="cuda").to(BF16)
    res = torch.randn((4096, 768), generator=g, device="cuda").to(BF16)
    p, seed, off = 0.1, 12345, 777
    out = ops.dropout(x, res, p, seed, off)
    mask, scale = keep_mask(seed, off, x.numel(), p)
    ref = (x.float().cpu().flatten() * mask * scale + res.float().cpu().flatten()).view_as(x)
    assert float((out.float().cpu() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())
    keep = float(mask.mean())
    assert abs(keep - 0.9) < 2e-3, keep
    assert torch.equal(ops.dropout(x, res, p, seed, off), out)                 # deterministic
    assert not torch.equal(ops.dropout(x, res, p, seed, off + 1), out)         # a different counter range
    assert torch.equal(ops.dropout(x, None, 0.0, seed, off), x)                # p = 0 is the identity


def test_dropout_add_function_backward_reuses_the_mask():
    from adapter4rec_b200 import functional as Fn
    Fn.DropoutState.manual_seed(99)
    x = torch.randn((512, 64), device="cuda").to(BF16).requires_grad_(True)
    r = torch.randn((512, 64), device="cuda").to(BF16).requires_grad_(True)
    out = Fn.dropout_add(x, r, 0.25)
    mask, scale = keep_mask(99, 0, x.numel(), 0.25)
    dy = torch.randn_like(out)
    out.backward(dy)
    ref_dx = (dy.float().cpu().flatten() * mask * scale).view_as(x)
    assert float((x.grad.float().cpu() - ref_dx).abs().max()) <= 2 ** -7 * float(ref_dx.abs().max())
    assert torch.equal(r.grad, dy)


@pytest.mark.parametrize("N,L,heads,dh,causal", [(6, 30, 12, 64, False), (9, 20, 2, 32, True)])
def test_attention_probability_dropout(N, L, heads, dh, causal):
    from adapter4rec_b200 import ops
    H = heads * dh
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn((N * L, 3 * H), generator=g, device="cuda").to(BF16)
    dctx = torch.randn((N * L, H), generator=g, device="cuda").to(BF16)
    p, seed, off = 0.1, 4242, 1000
    # mask[n, h, i, j] with nt = j/8, t = (j%8)/2, e = j%2:
    #   lane ((nt&1)*2 + e) of rng64(seed, off + ((n*heads + h)*32 + i)*8 + (nt>>1)*4 + t)      (attention_small.cu)
    n_i, h_i, i_i, j_i = np.meshgrid(np.arange(N), np.arange(heads), np.arange(L), np.arange(L), indexing="ij")
    nt, tt, ee = j_i // 8, (j_i % 8) // 2, j_i % 2
    ctr = (off + ((n_i * heads + h_i) * 32 + i_i) * 8 + (nt >> 1) * 4 + tt).astype(np.uint64)
    with np.errstate(over="ignore"):
        bits = rng64(seed, ctr)
    lane = (bits >> (16 * ((nt & 1) * 2 + ee)).astype(np.uint64)) & np.uint64(0xFFFF)
    thr = int(p * 65536.0 + 0.5)
    mask = torch.from_numpy((lane >= thr).astype(np.float32)).cuda()
    scale = 65536.0 / (65536 - thr)
    qf = qkv.float().requires_grad_(True)
    q, k, v = [t.view(N, L, heads, dh).transpose(1, 2) for t in qf.view(N * L, 3, H).unbind(1)]
    s = (q @ k.transpose(-1, -2)) * dh ** -0.5
    if causal:
        s = s + torch.where(torch.tril(torch.ones(L, L, dtype=torch.bool, device="cuda")), 0.0, -1e9)
    ref = ((torch.softmax(s, -1) * mask * scale) @ v).transpose(1, 2).reshape(N * L, H)
    got = ops.attn_small_fwd(qkv, N, L, heads, dh, causal=causal, mask_neg=-1e9, dropout=(p, seed, off))
    err = (got.float() - ref.detach()).abs()
    assert bool((err <= 2e-2 + 2 ** -6 * ref.detach().abs()).all()), float(err.max())
    ref.backward(dctx.float())
    dqkv = ops.attn_small_bwd(qkv, dctx, N, L, heads, dh, causal=causal, mask_neg=-1e9, dropout=(p, seed, off))
    for j, nm in enumerate(("dq", "dk", "dv")):
        a, b = dqkv[:, j * H:(j + 1) * H].float(), qf.grad[:, j * H:(j + 1) * H]
        rel = float((a - b).norm() / b.norm())
        assert rel <= 2e-2, "%s relative L2 error %.4f" % (nm, rel)


def test_train_mode_step_is_deterministic_given_the_seed_and_close_to_eval():
    import cases
    from adapter4rec_b200 import functional as Fn
    from test_model_gpu import build_gpu_model
    c = cases.tiny_case("houlsby")
    sd = cases.build_state_dict(c)
    model, _ = build_gpu_model(c, sd)
    items = cases.build_item_content(c)
    sample_items, log_mask, _ = cases.build_batch(c, items)
    rows, lm = sample_items.view(-1, 2 * c.L).cuda(), log_mask.cuda()
    model.eval()
    base = float(model(rows, lm, 0))
    model.train()
    losses, grads = [], []
    for _ in range(2):
        Fn.DropoutState.manual_seed(7)
        model.zero_grad(set_to_none=True)
        loss = model(rows, lm, 0)
        loss.backward()
        losses.append(float(loss))
        grads.append(torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]))
    assert losses[0] == losses[1] and torch.equal(grads[0], grads[1])        # same seed -> bit-identical step
    assert losses[0] != base and abs(losses[0] - base) < 0.5 * abs(base)      # dropout changes the loss, moderately
    assert torch.isfinite(grads[0]).all() and float(grads[0].norm()) > 0
    Fn.DropoutState.manual_seed(8)
    assert float(model(rows, lm, 0)) != losses[0]                             # another seed, another mask
