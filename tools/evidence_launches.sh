#!/bin/bash
# Round-2 launch lists (ncu gpu__time_duration per launch; cold-cache, serialised: the SHARES are what count).
O=gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCU -c 8000 --log-file $O/r02_launches_c2_128users.csv python bench.py --users 128 --users-per-pass 128 --steps 2 --warmup 1 --no-cpu-baseline --no-eval --no-variants > $O/r02_launches_c2.log 2>&1
$NCU -c 6000 --log-file $O/r02_launches_c3_32users.csv python tools/bench_c3.py --users 32 --users-per-pass 32 --steps 1 --warmup 1 > $O/r02_launches_c3.log 2>&1
$NCU -c 6000 --log-file $O/r02_launches_c1_128users.csv python tools/bench_variant.py c1 --users 128 --steps 1 > $O/r02_launches_c1.log 2>&1
tail -2 $O/r02_launches_c3.log $O/r02_launches_c1.log
