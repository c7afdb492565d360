"""Records the parity figures the -m gpu model tests bound (tests/parity_util.py) — run ON A B200:

    python tools/measure_parity.py            # writes tests/golden/parity_measured.json (+ a copy under gpurun_out/)

For every seeded case (tiny text kinds, tiny image kinds, the full-size BERT-base / RoBERTa-base / ViT-B/16 cases) it runs
the CUDA path and the fp32 CPU oracle on the same tensors and stores: loss relative error, item-embedding max-abs and
relative-L2 error, aggregate gradient relative-L2 error, and for EVERY trainable tensor ||g - g_oracle|| / ||g_oracle_all||.
The tests allow 2 x these figures.  Re-run after any change to a kernel's rounding behaviour and commit the new table."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import parity_util as P  # noqa: E402
import cases  # noqa: E402
import cases_cv  # noqa: E402


def main():
    assert torch.cuda.is_available(), "run on a B200"
    only = set(sys.argv[1:])
    table = {}
    if only and os.path.exists(P.MEASURED_PATH):
        table = P.measured()
    jobs = [("text/" + k, lambda k=k: P.text_case(k)) for k in cases.ALL_KINDS]
    jobs += [("cv/" + k, lambda k=k: P.cv_case(k)) for k in cases_cv.CV_ALL_KINDS]
    jobs += [("text_full/houlsby", lambda: P.text_case("houlsby", full=True)),
             ("text_full/lora", lambda: P.text_case("lora", full=True)),
             ("text_full/lora/unpad", lambda: P.text_case("lora", full=True, unpad=True)),
             ("text_full/prompt_cpc", lambda: P.text_case("prompt_cpc", full=True)),
             ("cv_full/cv_houlsby", lambda: P.cv_case("cv_houlsby", full=True))]
    for name, job in jobs:
        if only and name not in only and name.split("/")[0] not in only:
            continue
        t0 = time.time()
        fig, _ = job()
        table[name] = P.table_entry(fig)
        e = table[name]
        print("%-28s loss_rel %.2e emb max %.2e rel %.2e grad_all %.3e big-tensor rel %.3f cos %.5f  (%.1f s)" % (
            name, e["loss_rel"], e["emb_max_abs"], e["emb_rel_l2"], e.get("grad_all_rel", 0.0), e.get("tensor_rel_big", 0.0),
            e.get("tensor_cos_big", 1.0), time.time() - t0), flush=True)
        torch.cuda.empty_cache()
    table["_meta"] = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__,
                      "note": "written by tools/measure_parity.py; tests allow MARGIN = 2 x these figures (tests/parity_util.py)"}
    text = json.dumps(table, indent=1, sort_keys=True)
    with open(P.MEASURED_PATH, "w") as f:
        f.write(text + "\n")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_measured.json"), "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
