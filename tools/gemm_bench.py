"""GPU micro-benchmark of the tcgen05 GEMM on the shapes of one C2 pass (M = 128 users x 42 items x 30 tokens)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if "--lib" in sys.argv:      # A/B against a variant build of the library (tools/variants/*.so)
    import adapter4rec_b200.lib as _lib
    i = sys.argv.index("--lib")
    _lib.LIB_PATH = os.path.abspath(sys.argv[i + 1])
    del sys.argv[i:i + 2]
from adapter4rec_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 161280
BN = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = "cuda"
def r(*s): return (torch.randn(*s, device=dev) * 0.05).to(torch.bfloat16)
shapes = [
    ("T=x*Acat      ", 64, 768, {}),
    ("QKV+lora ext  ", 2304, 768, {"ext": 64, "bias": True}),
    ("QKV plain     ", 2304, 768, {"bias": True}),
    ("out-proj+res  ", 768, 768, {"bias": True, "res": True}),
    ("out-proj+res+drop", 768, 768, {"bias": True, "res": True, "drop": True}),
    ("FFN2+res+drop ", 768, 3072, {"bias": True, "res": True, "drop": True}),
    ("FFN1 gelu+aux ", 3072, 768, {"bias": True, "epi": ops.EPI_GELU, "aux": True}),
    ("FFN1 gelu     ", 3072, 768, {"bias": True, "epi": ops.EPI_GELU}),
    ("FFN1 linear   ", 3072, 768, {"bias": True}),
    ("FFN2+res      ", 768, 3072, {"bias": True, "res": True}),
    ("dFFN2 dgelu   ", 3072, 768, {"epi": ops.EPI_DGELU, "aux": True}),
    ("dFFN1+res     ", 768, 3072, {"res": True}),
    ("dQKV->dx +ext ", 768, 2304, {"ext": 64}),
    ("dT            ", 64, 2304, {}),
    ("plain 768x768 ", 768, 768, {}),
]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, N, K, o in shapes:
    a, b = r(M, K), r(N, K)
    kw = {"block_n": BN if N >= 256 else 0}
    if o.get("ext"): kw.update(a2=r(M, o["ext"]), b2=r(N, o["ext"]))
    if o.get("bias"): kw["bias"] = torch.randn(N, device=dev)
    if o.get("res"): kw["residual"] = r(M, N)
    if o.get("epi") is not None: kw["epilogue"] = o["epi"]
    if o.get("aux"): kw["aux"] = r(M, N)
    if o.get("drop"): kw["dropout"] = (0.1, 0x5EED, 1234)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(3): ops.gemm(a, b, out=out, **kw)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, b, out=out, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * M * N * (K + o.get("ext", 0))
    print("%s M=%d N=%4d K=%4d  %.3f ms  %.0f TFLOP/s" % (name, M, N, K, ms, fl / ms / 1e9), flush=True)
    del a, b, out, kw
# torch reference (cuBLAS) for context
a, b = r(M, 768), r(3072, 768)
for _ in range(3): torch.matmul(a, b.t())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); torch.matmul(a, b.t()); e1.record(); torch.cuda.synchronize()
print("cuBLAS  M=%d N=3072 K=768: %.3f ms %.0f TFLOP/s" % (M, e0.elapsed_time(e1), 2.0 * M * 3072 * 768 / e0.elapsed_time(e1) / 1e9))
