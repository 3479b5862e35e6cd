#!/bin/bash
# usage: tools/run_configs.sh "C3:20000 C5:100000"   (reduced-N bench of the other BASELINE configs)
for spec in $1; do
  c=${spec%%:*}; N=${spec##*:}
  start=$(date +%s)
  timeout 1200 python bench.py --config $c --candidates $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2>gpurun_out/bench_$c.err || tail -5 gpurun_out/bench_$c.err
  echo "$c wall $(( $(date +%s) - start )) s"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$c.json"))
    print("$c", "%.4g designs/s" % d["value"], "ms/step %.2f" % d["ms_per_step"], "tri frac %.3f" % d["roofline"]["frac"], "pipeline frac %.3f" % d["roofline"]["pipeline_frac"], d["roofline"]["kernel_ms_per_step"], d.get("e2e",{}).get("value"), d["clocks"])
except Exception as e:
    print("$c failed", e)
PY
done
