#!/bin/bash
# usage: tools/sweep_tri.sh "C2:100000 C5:50000"  -> both stage depths per config
for spec in $1; do
  c=${spec%%:*}; N=${spec##*:}
  for cps in 1 2; do
    timeout 900 python bench.py --config $c --candidates $N --steps 3 --warmup 3 --no-cpu-baseline --tri-cps $cps > gpurun_out/sw_$c_$cps.json 2>gpurun_out/sw.err || tail -3 gpurun_out/sw.err
    python - <<PY
import json
d=json.load(open("gpurun_out/sw_$c_$cps.json"))
print("$c cps=$cps  %.4g designs/s  tri frac %.4f  tri ms %.3f" % (d["value"], d["roofline"]["frac"], d["roofline"]["launch_ms"]))
PY
  done
done
