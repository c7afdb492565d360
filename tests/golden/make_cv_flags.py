"""Dumps the defaults of the reference's CV command line (Downstream/CV/parameters.py:4-70) to tests/golden/cv_flags.json.

The reference module starts with `from data_utils.utils import *`, whose package imports lmdb (not installed here); the only
name it needs from there is `argparse`, so the source is executed with that one line replaced — the parser itself runs unmodified."""
import json
import os
import sys

REF = "/root/reference/Downstream/CV/parameters.py"
HERE = os.path.dirname(os.path.abspath(__file__))

src = open(REF).read().replace("from data_utils.utils import *", "import argparse", 1)
ns = {"__name__": "reference_cv_parameters"}
exec(compile(src, REF, "exec"), ns)
argv, sys.argv = sys.argv, ["run_adapter.py"]
try:
    args = ns["parse_args"]()
finally:
    sys.argv = argv
with open(os.path.join(HERE, "cv_flags.json"), "w") as f:
    json.dump(vars(args), f, indent=1, sort_keys=True)
    f.write("\n")
print(len(vars(args)), "flags")
