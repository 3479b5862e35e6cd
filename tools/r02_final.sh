#!/bin/bash
# Final round-2 evidence on one GPU: full -m gpu suite, smoke, bench (both arms), ncu refresh of the C2 kernels.
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_final_pytest.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_final_ref.json 2> /dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $O/r02_final_bench.json 2> $O/r02_final_bench.err; echo "bench rc=$?"
python tools/prof_step.py --config C2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tri_kernel|eval_kernel" -s 4 -c 4 -o $O/r02_final_c2 -f python tools/prof_step.py --config C2 > $O/r02_final_ncu.log 2>&1
python tools/ncu_summary.py $O/r02_final_c2.ncu-rep > $O/r02_final_c2_ncu.txt 2>&1; grep -E "^==|time_duration|fp64_cycles|dmma_cycles|dram__bytes" $O/r02_final_c2_ncu.txt | head -24
rm -f $O/r02_final_c2.ncu-rep
