#!/bin/bash
# Round-2 ncu captures, part 2: C4 and C5 at the reduced N of the bench's extra blocks (each command ran without ncu first).
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_step.py --config C4 --candidates 18944 > $O/r02_prof_c4.log 2>&1 || exit 1
python tools/prof_step.py --config C5 --candidates 47360 > $O/r02_prof_c5.log 2>&1 || exit 1
$NCU -k regex:"tri_kernel|eval_kernel" -s 2 -c 2 -o $O/r02_c4 -f python tools/prof_step.py --config C4 --candidates 18944 > $O/r02_ncu_c4.log 2>&1
$NCU -k regex:"tri_kernel|eval_kernel" -s 2 -c 2 -o $O/r02_c5 -f python tools/prof_step.py --config C5 --candidates 47360 > $O/r02_ncu_c5.log 2>&1
ls -la $O/r02_c4.ncu-rep $O/r02_c5.ncu-rep
