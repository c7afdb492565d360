"""Per-kernel counts of the SASS mnemonics that tell a Blackwell-native kernel from a recompiled one (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), HMMA (mma.sync), LDGSTS (cp.async), MUFU.
Reads the in-tree library with cuobjdump (no GPU needed):

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.json"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adapter4rec_b200", "libadapter4rec_sm100.so")
WANT = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "MUFU", "SYNCS")


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in WANT}
            kernels[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["instructions"] += 1
        base = op.split(".")[0]
        if base in kernels[cur]:
            kernels[cur][base] += 1
    names = list(kernels)
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except Exception:
        demangle = {n: n for n in names}
    rows = []
    for n in names:
        d = re.sub(r"\(anonymous namespace\)::", "", demangle.get(n, n))
        d = re.sub(r"\(CUtensorMap_st.*", "", d)
        c = kernels[n]
        rows.append({"kernel": d, "instructions": c["instructions"], "tcgen05_mma": c["UTCHMMA"] + c["UTCQMMA"], "tcgen05_commit": c["UTCBAR"],
                     "tcgen05_ld": c["LDTM"], "tcgen05_st": c["STTM"], "tma_load": c["UTMALDG"], "tma_store": c["UTMASTG"],
                     "mma_sync_hmma": c["HMMA"], "cp_async_ldgsts": c["LDGSTS"], "mufu": c["MUFU"], "mbarrier_syncs": c["SYNCS"]})
    rows.sort(key=lambda r: (-r["tcgen05_mma"], r["kernel"]))
    json.dump({"library": os.path.relpath(LIB, ROOT), "how": "cuobjdump -sass, mnemonics counted per function (static counts)",
               "kernels": rows}, sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main()
