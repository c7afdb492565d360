#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (SURVEY.md §5 "race detection"): memcheck, racecheck (shared-memory
# hazards), synccheck (barrier / .aligned misuse).  Each tool is time-boxed; logs land in gpurun_out/.
O=gpurun_out
T=${1:-420}
SEL="gemm or adapter or layernorm or attn or score or wgrad or embed or topk"
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  timeout $T compute-sanitizer --tool $tool --print-limit 20 --log-file $O/r02_sanitizer_$tool.log \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" -p no:cacheprovider > $O/r02_sanitizer_${tool}_pytest.log 2>&1
  echo "== $tool: exit $?"; tail -3 $O/r02_sanitizer_${tool}_pytest.log; grep -c "=========" $O/r02_sanitizer_$tool.log; grep "SUMMARY" $O/r02_sanitizer_$tool.log | tail -2
done
