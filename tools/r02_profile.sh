#!/bin/bash
# Round-2 ncu captures (run through gpurun; every command below first ran to completion WITHOUT ncu in the same call).
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_step.py --config C2 > $O/r02_prof_c2.log 2>&1 || exit 1
python tools/prof_step.py --config C2 --option fused=1 > $O/r02_prof_c2_fused.log 2>&1 || exit 1
python tools/prof_step.py --config C3 --candidates 18944 > $O/r02_prof_c3.log 2>&1 || exit 1
python tools/prof_step.py --config C4 --candidates 18944 > $O/r02_prof_c4.log 2>&1 || exit 1
$NCU -k regex:"tri_kernel|eval_kernel" -s 4 -c 4 -o $O/r02_c2_tri_eval -f python tools/prof_step.py --config C2 > $O/r02_ncu_c2.log 2>&1
$NCU -k regex:"fused_kernel" -s 1 -c 1 -o $O/r02_c2_fused -f python tools/prof_step.py --config C2 --option fused=1 > $O/r02_ncu_c2_fused.log 2>&1
$NCU -k regex:"tri_kernel|eval_kernel|gram_kernel" -s 5 -c 5 -o $O/r02_c3 -f python tools/prof_step.py --config C3 --candidates 18944 > $O/r02_ncu_c3.log 2>&1
$NCU -k regex:"tri_kernel|eval_kernel" -s 4 -c 4 -o $O/r02_c4 -f python tools/prof_step.py --config C4 --candidates 18944 > $O/r02_ncu_c4.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-stepwise > $O/r02_launchlist_bench.json 2> $O/r02_launchlist_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-stepwise > $O/r02_ncu_launches.log 2>&1
ls -la $O/*.ncu-rep
