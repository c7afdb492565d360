"""profiles/r02_ncu_full_gemm_summary.json (tools/ncu_summarise.py over the capture of tools/ncu_target.py) ->
profiles/r02_ncu_full_gemm_traffic.json: per GEMM shape the DRAM bytes of one launch (dram__bytes_read + write), which
bench.py scales by the row count and averages over the launches of its timed region for `roofline.traffic`."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_target import SHAPES
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_ncu_full_gemm_summary.json")
recs = json.load(open(src))
assert len(recs) == len(SHAPES), (len(recs), len(SHAPES))
out = []
for shape, r in zip(SHAPES, recs):
    M, N, K, epi = shape
    # algorithmic bytes: A once, C once, + the streamed epilogue tensor (residual / pre-activation in or out); weights are L2-resident
    extra = M * N * 2 if (epi in (1, 3) or (epi == 0 and N == 768 and K in (768, 3072))) else 0
    out.append({"shape": list(shape), "kernel": r["kernel"][:64], "dram_bytes_per_launch": r["dram_bytes_per_launch"],
                "algorithmic_bytes": M * K * 2 + M * N * 2 + extra,
                "duration_ms_under_ncu": r["duration_us"] / 1e3, "tensor_pipe_active_pct": r["tensor_pipe_pct"],
                "note": "ncu --set full --clock-control none, tools/ncu_target.py at HEAD (round 2); summary in profiles/r02_ncu_full_gemm_summary.json"})
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_full_gemm_traffic.json"), "w"), indent=1)
for o in out:
    print(o["shape"], "%.0f MB DRAM vs %.0f MB algorithmic (x%.2f), tensor pipe %.1f%%" % (
        o["dram_bytes_per_launch"] / 1e6, o["algorithmic_bytes"] / 1e6, o["dram_bytes_per_launch"] / o["algorithmic_bytes"], o["tensor_pipe_active_pct"]))
