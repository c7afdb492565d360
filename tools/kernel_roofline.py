"""Per-kernel roofline table: every kernel of the hot path timed ALONE at its C2 / C3 / C5 shape with CUDA events on
the launching stream (3 warm-ups, then `--iters` launches over rotating buffers that together exceed the 126 MB L2),
algorithmic bytes / FLOPs as stated in DESIGN.md §4, against MEASURED_PEAKS.json (burst figures: kernels timed alone).

    python tools/kernel_roofline.py > gpurun_out/kernel_roofline.json

One JSON object per line; the last line is a summary table.  Copy into profiles/ for the record."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--lib" in sys.argv:      # A/B against a variant build of the library (tools/variants/*.so)
    import adapter4rec_b200.lib as _lib
    _i = sys.argv.index("--lib")
    _lib.LIB_PATH = os.path.abspath(sys.argv[_i + 1])
    del sys.argv[_i:_i + 2]
import torch  # noqa: E402

from adapter4rec_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--users", type=int, default=128, help="users per pass of C2 (M = users*42*30 tokens)")
ap.add_argument("--only", default="")
a = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
try:
    PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PK_SRC = "measured"
except Exception:
    PK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    PK_SRC = "fallback"
HBM = PK["hbm_gbs"]
TF = PK["bf16_tflops"]
BF16 = torch.bfloat16
H, L, HEADS = 768, 30, 12
NSEQ = a.users * 42
M = NSEQ * L
rows = []


def r16(*s, scale=0.05):
    return (torch.randn(*s, device=dev) * scale).to(BF16)


def timeit(fn, nrot):
    for i in range(3):
        fn(i % nrot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        fn(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters * 1e3    # us per launch


def report(name, us, nbytes=None, flops=None, bound="hbm", note=""):
    rec = {"kernel": name, "us_per_launch": round(us, 2), "bound": bound}
    if nbytes is not None:
        rec["algorithmic_bytes"] = int(nbytes)
        rec["GBps"] = round(nbytes / us / 1e3, 1)
    if flops is not None:
        rec["algorithmic_flops"] = float(flops)
        rec["TFLOPs"] = round(flops / us / 1e6, 1)
    if bound == "hbm":
        rec["frac"] = round(rec["GBps"] / HBM, 3)
    elif bound == "tensor":
        rec["frac"] = round(rec["TFLOPs"] / TF, 3)
    if note:
        rec["note"] = note
    rows.append(rec)
    print(json.dumps(rec), flush=True)


def want(name):
    return not a.only or any(t in name for t in a.only.split(","))


NR = 3  # rotating buffer sets (each kernel's working set is >= 250 MB: 3 sets never fit the 126 MB L2)

# ------------------------------------------------------------------ K3 short-sequence attention (C2: L = 30)
if want("attn_small"):
    qkv = [r16(M, 3 * H, scale=0.5) for _ in range(NR)]
    dctx = [r16(M, H) for _ in range(NR)]
    ids = torch.ones((NSEQ, L), dtype=torch.int64, device=dev)
    us = timeit(lambda i: ops.attn_small_fwd(qkv[i], NSEQ, L, HEADS, 64, mask=ids), NR)
    report("attn_small_fwd L=30 (BERT, key mask)", us, nbytes=M * 4 * H * 2, flops=4.0 * L * H * M)
    us = timeit(lambda i: ops.attn_small_bwd(qkv[i], dctx[i], NSEQ, L, HEADS, 64, mask=ids), NR)
    report("attn_small_bwd L=30", us, nbytes=M * 7 * H * 2, flops=10.0 * L * H * M)
    drop = (0.1, 1234, 0)
    us = timeit(lambda i: ops.attn_small_fwd(qkv[i], NSEQ, L, HEADS, 64, mask=ids, dropout=drop), NR)
    report("attn_small_fwd L=30 + probability dropout", us, nbytes=M * 4 * H * 2)
    us = timeit(lambda i: ops.attn_small_bwd(qkv[i], dctx[i], NSEQ, L, HEADS, 64, mask=ids, dropout=drop), NR)
    report("attn_small_bwd L=30 + probability dropout", us, nbytes=M * 7 * H * 2)
    del qkv, dctx

# ------------------------------------------------------------------ K6 LayerNorm
if want("layernorm"):
    x = [r16(M, H, scale=1.0) for _ in range(NR)]
    dy = [r16(M, H) for _ in range(NR)]
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    us = timeit(lambda i: ops.layernorm_fwd(x[i], g, b, 1e-12), NR)
    report("layernorm_fwd H=768 (+stats)", us, nbytes=M * (2 * H * 2 + 8))
    _, _, mean, rstd = ops.layernorm_fwd(x[0], g, b, 1e-12)
    us = timeit(lambda i: ops.layernorm_bwd(dy[i], x[i], mean, rstd, g), NR)
    report("layernorm_bwd H=768", us, nbytes=M * (3 * H * 2 + 8))
    us = timeit(lambda i: ops.layernorm_bwd(dy[i], x[i], mean, rstd, g, masked=(0.1, 77, 0)), NR)
    report("layernorm_bwd H=768 + dropout-masked second output", us, nbytes=M * (4 * H * 2 + 8))
    dg, db = torch.empty(H, device=dev), torch.empty(H, device=dev)
    us = timeit(lambda i: ops.layernorm_bwd(dy[i], x[i], mean, rstd, g, dgamma=dg, dbeta=db), NR)
    report("layernorm_bwd H=768 + dgamma/dbeta", us, nbytes=M * (3 * H * 2 + 8))
    us = timeit(lambda i: ops.dropout(x[i], dy[i], 0.1, 5, 0), NR)
    report("dropout + residual add", us, nbytes=M * 3 * H * 2)
    del x, dy

# ------------------------------------------------------------------ K1 embedding gather + LayerNorm
if want("embed"):
    word = r16(30522, H)
    pos = r16(512, H)
    typ = r16(2, H)
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    idr = [torch.randint(0, 30522, (NSEQ, 2 * L), device=dev) for _ in range(NR)]
    us = timeit(lambda i: ops.embed_ln_fwd(idr[i], L, word, pos, typ, g, b, 1e-12), NR)
    report("embed_ln_fwd (word+pos+type gather + LN)", us, nbytes=M * (8 + H * 2 + H * 2))
    del word, idr

# ------------------------------------------------------------------ skinny weight gradients (LoRA)
if want("wgrad"):
    dqkv = [r16(M, 3 * H) for _ in range(NR)]
    T = [r16(M, 64) for _ in range(NR)]
    x = [r16(M, H) for _ in range(NR)]
    us = timeit(lambda i: ops.wgrad(dqkv[i], T[i]), NR)
    report("wgrad dqkv^T.[T|1]  (N=2304, K=64)", us, nbytes=M * (3 * H + 64) * 2, flops=2.0 * M * 3 * H * 64)
    us = timeit(lambda i: ops.wgrad(T[i], x[i]), NR)
    report("wgrad dT^T.x  (N=64, K=768)", us, nbytes=M * (H + 64) * 2, flops=2.0 * M * H * 64)
    us = timeit(lambda i: ops.colsum(x[i]), NR)
    report("colsum H=768", us, nbytes=M * H * 2)
    del dqkv, T, x

# ------------------------------------------------------------------ K5 fused Houlsby adapter block
if want("adapter"):
    h = [r16(M, H, scale=1.0) for _ in range(NR)]
    inp = [r16(M, H, scale=1.0) for _ in range(NR)]
    wd, wu = r16(64, H, scale=0.01), r16(H, 64, scale=0.01)
    bd, bu = torch.zeros(64, device=dev), torch.zeros(H, device=dev)
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    us = timeit(lambda i: ops.adapter_ln_fwd(h[i], inp[i], wd, bd, wu, bu, g, b, 1e-12, act="relu", tail=0), NR)
    report("adapter_ln_fwd r=64 (down->relu->up->+h+input->LN), inference", us, nbytes=M * 3 * H * 2,
           flops=4.0 * M * H * 64)
    us = timeit(lambda i: ops.adapter_ln_fwd(h[i], inp[i], wd, bd, wu, bu, g, b, 1e-12, act="relu", tail=0, save=True), NR)
    report("adapter_ln_fwd r=64, training (also writes z, s, stats)", us, nbytes=M * (4 * H * 2 + 64 * 2 + 8),
           flops=4.0 * M * H * 64)
    us = timeit(lambda i: ops.adapter_ln_fwd(h[i], inp[i], wd, bd, wu, bu, act="gelu", tail=1, save=True), NR)
    report("adapter_ln_fwd r=64 gelu, tail=+input (ViT output), training", us, nbytes=M * (3 * H * 2 + 2 * 64 * 2),
           flops=4.0 * M * H * 64)
    del h, inp

# ------------------------------------------------------------------ K9 losses
if want("loss"):
    B, S, D = 4096, 20, 64
    prec = [r16(B, S, D, scale=0.3) for _ in range(NR)]
    emb = [r16(B, S + 1, 2, D, scale=0.3) for _ in range(NR)]
    lm = torch.ones((B, S), device=dev)
    us = timeit(lambda i: ops.bce_loss_fwd(prec[i], emb[i], lm), NR)
    report("bce_loss_fwd B=4096 S=20 D=64", us, nbytes=B * S * 3 * D * 2, note="tiny: launch/latency-bound")
    loss, count, pos, neg = ops.bce_loss_fwd(prec[0], emb[0], lm)
    us = timeit(lambda i: ops.bce_loss_bwd(prec[i], emb[i], lm, pos, neg, count), NR)
    report("bce_loss_bwd B=4096", us, nbytes=B * S * 3 * D * 2 + B * (S + (S + 1) * 2) * D * 2)
    B = 512
    prec = [r16(B, S, D, scale=0.3) for _ in range(NR)]
    emb = [r16(B, S + 1, 2, D, scale=0.3) for _ in range(NR)]
    lm = torch.ones((B, S), device=dev)
    item_ids = torch.randint(1, 80000, (B, S + 1), device=dev)
    us = timeit(lambda i: ops.inbatch_ce_fwd(prec[i], emb[i], item_ids, lm), NR)
    report("inbatch_ce_fwd B=512 (10,240 queries x 10,752 candidates)", us, flops=2.0 * B * S * B * (S + 1) * D,
           bound="on-chip", note="logits never leave the SM; candidates are L2-resident")
    loss, count, lse = ops.inbatch_ce_fwd(prec[0], emb[0], item_ids, lm)
    us = timeit(lambda i: ops.inbatch_ce_bwd(prec[i], emb[i], item_ids, lm, lse, count), NR)
    report("inbatch_ce_bwd B=512", us, flops=6.0 * B * S * B * (S + 1) * D, bound="on-chip")

# ------------------------------------------------------------------ K14 Adam
if want("adam"):
    n = 64 << 20
    p, gr, m, v = [torch.randn(n, device=dev) * 0.01 for _ in range(4)]
    us = timeit(lambda i: ops.adam_step(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-8, 0.0, i + 1), 1)
    report("adam_step 64 Mi params (4 reads + 3 writes of f32)", us, nbytes=n * 28)
    del p, gr, m, v

# ------------------------------------------------------------------ K10 / K13 evaluator side kernels
if want("eval"):
    tab = r16(1_000_001, 64, scale=0.3)
    idx = [torch.randint(0, 1_000_001, (65536, 20), device=dev) for _ in range(NR)]
    us = timeit(lambda i: ops.gather_rows(tab, idx[i]), NR)
    report("gather_rows 65,536 users x 20 x D=64", us, nbytes=65536 * 20 * (8 + 2 * 64 * 2))
    P, U = 16, 65536
    sc = [torch.randn((P, U, 10), device=dev).sort(dim=2, descending=True)[0].contiguous() for _ in range(NR)]
    ii = [torch.randint(1, 10_000_000, (P, U, 10), device=dev, dtype=torch.int32) for _ in range(NR)]
    tgt = torch.randint(1, 10_000_000, (U,), device=dev, dtype=torch.int32)
    us = timeit(lambda i: ops.topk_merge(sc[i], ii[i], target=tgt), NR)
    report("topk_merge 16 partial lists x 65,536 users (+HR/NDCG)", us, nbytes=P * U * 80 + U * 92)
    del tab, idx, sc, ii

# ------------------------------------------------------------------ train-batch assembly (negative sampling + row gather)
if want("batch"):
    I, B, S1, W = 80_000, 4096, 21, 60
    content = torch.randint(0, 30522, (I + 1, W), device=dev)
    seqs = [torch.stack([torch.randperm(I, device=dev)[:S1] + 1 for _ in range(64)]).repeat(B // 64, 1).contiguous()
            for _ in range(NR)]
    us = timeit(lambda i: ops.sample_train_batch(seqs[i], content, I, 7, i << 40), NR)
    report("sample_train_batch B=4096 users (S=20, L=30): sampler + 2(S+1) int64 rows/user", us,
           nbytes=B * S1 * (2 * W * 8 * 2 + 8 + 8 + 4), note="reads are random 480 B rows of a 38 MB table (L2-resident)")
    del content, seqs

# ------------------------------------------------------------------ K11/K12 score GEMM + top-k  (C5)
if want("score"):
    I, d, U = 2_000_000, 768, 9472
    table = r16(I, d, scale=d ** -0.5)
    users = [r16(U, d, scale=1.0) for _ in range(2)]
    hist = torch.randint(1, I, (U, 20), device=dev, dtype=torch.int32)
    us = timeit(lambda i: ops.score_topk(users[i], table, id_base=0, history=hist, k=10), 2)
    report("score_topk U=9472 x I=2M x d=768 (+history mask + top-10)", us, flops=2.0 * U * I * d, bound="tensor")
    del table
    I, d, U = 4_000_000, 64, 8192
    table = r16(I, d, scale=0.3)
    users = [r16(U, d, scale=1.0) for _ in range(2)]
    hist = torch.randint(1, I, (U, 20), device=dev, dtype=torch.int32)
    us = timeit(lambda i: ops.score_topk(users[i], table, id_base=0, history=hist, k=10), 2)
    report("score_topk U=8192 x I=4M x d=64", us, flops=2.0 * U * I * d, bound="epilogue",
           note="d=64: 128 FLOP per score, the top-k epilogue (1 compare per score) is the bound; scores/s = %.3g"
                % (U * I / us * 1e6))
    del table

# ------------------------------------------------------------------ tcgen05 GEMM shapes of one C2 pass
if want("gemm"):
    x = [r16(M, H) for _ in range(2)]
    w_qkv, t, bext = r16(3 * H, H), r16(M, 64), r16(3 * H, 64)
    w1, w2, wo = r16(4 * H, H), r16(H, 4 * H), r16(H, H)
    b1, b2, bq = torch.randn(4 * H, device=dev), torch.randn(H, device=dev), torch.randn(3 * H, device=dev)
    u = torch.empty(M, 4 * H, dtype=BF16, device=dev)
    f = r16(M, 4 * H)
    dq = r16(M, 3 * H)
    us = timeit(lambda i: ops.gemm(x[i], w_qkv, bias=bq, a2=t, b2=bext), 2)
    report("gemm QKV + LoRA K-extension  N=2304 K=768+64", us, flops=2.0 * M * 3 * H * (H + 64), bound="tensor")
    us = timeit(lambda i: ops.gemm(x[i], wo, bias=b2, residual=x[1 - i]), 2)
    report("gemm attention.output + residual  N=768 K=768", us, flops=2.0 * M * H * H, bound="tensor")
    us = timeit(lambda i: ops.gemm(x[i], w1, bias=b1, epilogue=ops.EPI_GELU, aux=u), 2)
    report("gemm FFN1 + GELU (+pre-activation out)  N=3072 K=768", us, flops=2.0 * M * 4 * H * H, bound="tensor")
    us = timeit(lambda i: ops.gemm(f, w2, bias=b2, residual=x[i]), 2)
    report("gemm FFN2 + residual  N=768 K=3072", us, flops=2.0 * M * 4 * H * H, bound="tensor")
    us = timeit(lambda i: ops.gemm(x[i], w1, epilogue=ops.EPI_DGELU, aux=u), 2)
    report("gemm dFFN2 + GELU' epilogue  N=3072 K=768", us, flops=2.0 * M * 4 * H * H, bound="tensor")
    us = timeit(lambda i: ops.gemm(x[i], w1, bias=b1, epilogue=ops.EPI_GELU), 2)
    report("gemm FFN1 + GELU, inference (one output)  N=3072 K=768", us, flops=2.0 * M * 4 * H * H, bound="tensor")
    us = timeit(lambda i: ops.gemm(x[i], w1, bias=b1), 2)
    report("gemm N=3072 K=768 plain linear + bias (one output)", us, flops=2.0 * M * 4 * H * H, bound="tensor")
    wt = r16(H, 3 * H)
    us = timeit(lambda i: ops.gemm(dq, wt), 1)
    report("gemm dQKV (dx = dqkv.W)  N=768 K=2304", us, flops=2.0 * M * H * 3 * H, bound="tensor")
    a_cat = r16(64, H)
    us = timeit(lambda i: ops.gemm(x[i], a_cat), 2)
    report("gemm LoRA down T = x.A_cat^T  N=64 K=768", us, nbytes=M * (H + 64) * 2, flops=2.0 * M * 64 * H, bound="hbm")
    del x, u, f, dq, t

# ------------------------------------------------------------------ C3: ViT kernels (32 users per pass = 704 images)
if want("vit"):
    NI, LV = 704, 197
    MV = NI * LV
    qkv = [r16(MV, 3 * H, scale=0.5) for _ in range(NR)]
    dctx = [r16(MV, H) for _ in range(NR)]
    us = timeit(lambda i: ops.attn_small_fwd(qkv[i], NI, LV, HEADS, 64, want_lse=True), NR)
    report("attn_mid_fwd L=197 (ViT)", us, nbytes=MV * (4 * H * 2 + HEADS * 4), flops=4.0 * LV * H * MV, bound="mma.sync",
           note="FLOPs = 4*L*768 per token")
    out, lse = ops.attn_small_fwd(qkv[0], NI, LV, HEADS, 64, want_lse=True)
    us = timeit(lambda i: ops.attn_small_bwd(qkv[i], dctx[i], NI, LV, HEADS, 64, lse=lse, ctx=out), NR)
    report("attn_mid_bwd L=197 (ViT)", us, nbytes=MV * (8 * H * 2 + HEADS * 4), flops=10.0 * LV * H * MV, bound="mma.sync")
    del qkv, dctx
    img = [torch.rand((NI, 3, 224, 224), device=dev) for _ in range(2)]
    us = timeit(lambda i: ops.patchify(img[i], 16), 2)
    report("patchify 704 images (f32 -> bf16 im2col)", us, nbytes=NI * 3 * 224 * 224 * 6)
    pe = r16(NI * 196, H)
    cls, pos = r16(1, H), r16(197, H)
    us = timeit(lambda i: ops.vit_assemble(pe, cls, pos, None, NI, 196), 1)
    report("vit_assemble (cls|patches + pos)", us, nbytes=NI * (196 + 197) * H * 2)

print(json.dumps({"summary": rows, "peaks": {"hbm_gbs": HBM, "bf16_tflops": TF, "source": PK_SRC},
                  "shape": {"tokens": M, "sequences": NSEQ}}))
