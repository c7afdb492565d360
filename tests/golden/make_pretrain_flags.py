"""Dumps the defaults of the reference's pre-training command lines (Pretraining/Text/parameters.py:4-51,
Pretraining/CV/parameters.py:4-45) to tests/golden/pretrain_{text,cv}_flags.json.

Both modules start with `from data_utils.utils import *` (the CV package imports lmdb, not installed here); the only name the
parsers need from there is `argparse`, so that one line is replaced — the parsers themselves run unmodified."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

for tree, out in (("Text", "pretrain_text_flags.json"), ("CV", "pretrain_cv_flags.json")):
    ref = "/root/reference/Pretraining/%s/parameters.py" % tree
    src = open(ref).read().replace("from data_utils.utils import *", "import argparse", 1)
    ns = {"__name__": "reference_pretraining_parameters"}
    exec(compile(src, ref, "exec"), ns)
    argv, sys.argv = sys.argv, ["run.py"]
    try:
        args = ns["parse_args"]()
    finally:
        sys.argv = argv
    with open(os.path.join(HERE, out), "w") as f:
        json.dump(vars(args), f, indent=1, sort_keys=True)
        f.write("\n")
    print(tree, len(vars(args)), "flags")
