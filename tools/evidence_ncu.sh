#!/bin/bash
# Round-2 ncu evidence: one `--set full` capture per hot kernel (run under gpurun; reports land in gpurun_out/).
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:gemm_tn --launch-skip 16 -c 8 -o $O/r02_ncu_gemm python tools/ncu_target.py > $O/r02_ncu_gemm.log 2>&1
$NCU -k regex:adapter_rows --launch-skip 2 -c 1 -o $O/r02_ncu_adapter_rows python tools/ncu_adapter.py > $O/r02_ncu_adapter.log 2>&1
$NCU -k regex:attn --launch-skip 6 -c 4 -o $O/r02_ncu_attn_tc python tools/ncu_attn_tc.py --bwd > $O/r02_ncu_attn_tc.log 2>&1
$NCU -k regex:score_topk --launch-skip 1 -c 1 -o $O/r02_ncu_score_topk python tools/ncu_score.py > $O/r02_ncu_score.log 2>&1
$NCU -k regex:ln_ --launch-skip 4 -c 2 -o $O/r02_ncu_layernorm python tools/ncu_ln.py > $O/r02_ncu_ln.log 2>&1
$NCU -k "regex:attn_fwd|attn_bwd|wgrad_tc_kernel" --launch-skip 5 -c 5 -o $O/r02_ncu_misc python tools/ncu_misc.py > $O/r02_ncu_misc.log 2>&1
ls -la $O/*.ncu-rep
tail -2 $O/r02_ncu_*.log
