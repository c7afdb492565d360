"""Debug aid: first op whose output differs between the padded and the unpadded token layout (tiny cases).  Every
adapter4rec_b200.ops entry point is wrapped to record its bf16/f32 outputs; the two traces are compared call by call
(padded rows are mapped through PackedTokens.token_rows / cls rows where the row counts differ)."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tests/golden"); sys.path.insert(0, "/root/repo/oracle")
import torch, cases
from test_model_gpu import build_gpu_model
from adapter4rec_b200 import ops
from adapter4rec_b200.model import bert as bert_mod
from adapter4rec_b200.data_utils.metrics import core_model

NAMES = ["gemm", "attn_small_fwd", "layernorm_fwd", "embed_ln_fwd", "adapter_ln_fwd", "gather_rows", "act_fwd", "dropout"]
trace = []
orig = {n: getattr(ops, n) for n in NAMES}


def wrap(name):
    def f(*a, **k):
        out = orig[name](*a, **k)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        info = ""
        if name == "gemm":
            info = "M=%d N=%d K=%d epi=%s res=%s strideA=%s" % (a[0].shape[0], a[1].shape[0], a[0].shape[1], k.get("epilogue", "lin"),
                                                               k.get("residual") is not None, tuple(a[0].stride()))
        trace.append((name + " " + info, [o.detach().float().cpu().clone() for o in outs if torch.is_tensor(o) and o.dtype in (torch.bfloat16, torch.float32)]))
        return out
    return f


for n in NAMES:
    setattr(ops, n, wrap(n))

kinds = sys.argv[1:] or ["pfeiffer_ver2"]
for kind in kinds:
    c = cases.tiny_case(kind); sd = cases.build_state_dict(c); items = cases.build_item_content(c)
    runs = {}
    packed_holder = {}
    P0 = bert_mod.PackedTokens

    class Spy(P0):
        def __init__(self, m):
            super().__init__(m)
            packed_holder["p"] = self
    bert_mod.PackedTokens = Spy
    for unpad in (False, True):
        model, _ = build_gpu_model(c, sd); model.eval()
        core_model(model).bert_encoder.text_encoders.title.bert_model.unpad = unpad
        trace.clear()
        with torch.no_grad():
            core_model(model).bert_encoder(items.cuda())
        runs[unpad] = list(trace)
    p = packed_holder["p"]
    rows = p.token_rows.cpu(); cls = p.cls_rows.cpu()
    N = items.shape[0]; T = rows.numel(); NL = None
    print(kind, "calls padded %d unpadded %d, N=%d T=%d" % (len(runs[False]), len(runs[True]), N, T))
    i = j = 0
    A, B = runs[False], runs[True]
    while i < len(A) and j < len(B):
        na, oa = A[i]; nb, ob = B[j]
        if na.split()[0] != nb.split()[0]:
            # the unpadded trace has extra gather_rows calls
            if nb.startswith("gather_rows"):
                j += 1; continue
            if na.startswith("gather_rows"):
                i += 1; continue
            print("  trace mismatch", na, "|", nb); break
        for x, y in zip(oa, ob):
            if x.dim() < 2: continue
            x2, y2 = x.reshape(-1, x.shape[-1]), y.reshape(-1, y.shape[-1])
            if x2.shape[0] != y2.shape[0]:
                if y2.shape[0] == T and x2.shape[0] % N == 0:
                    x2 = x2[rows]
                else:
                    print("  [%d/%d] %s: shapes %s vs %s (skipped)" % (i, j, na, tuple(x.shape), tuple(y.shape))); continue
            if x2.shape != y2.shape:
                print("  [%d/%d] %s: shapes %s vs %s (skipped)" % (i, j, na, tuple(x.shape), tuple(y.shape))); continue
            d = (x2 != y2)
            if d.any():
                rr = d.any(1).nonzero().flatten().tolist()
                print("  [%d/%d] %s | %s: DIFF rows %s (of %d) cols %s max %.3g" % (i, j, na, nb, rr[:8], x2.shape[0], d.any(0).nonzero().flatten().tolist()[:8],
                                                                              float((x2 - y2).abs().max())))
            else:
                print("  [%d/%d] %s: equal %s" % (i, j, na, tuple(x2.shape)))
        i += 1; j += 1
