#!/bin/bash
# Round-2 regression run on one GPU: full -m gpu suite, smoke, both bench arms with their wall time.
O=gpurun_out; T=${1:-s3}
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
s=$(date +%s); python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_ref.json 2> /dev/null; echo "ref rc=$? wall $(( $(date +%s) - s )) s"
s=$(date +%s); python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$? wall $(( $(date +%s) - s )) s"
python - <<PY
import json
d = json.loads(open("$O/${T}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("C2 value %.4g e2e %.4g ms %.3f tri frac %.4f pipeline %.4f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["pipeline_frac"], d["parity"]["max_rel_err"]))
for k, v in d["configs"].items():
    print(k, "value %.4g tri %.4f pipeline %.4f setup_s %.2f parity %.2e" % (v["value"], v["roofline"]["frac"], v["roofline"]["pipeline_frac"], v["setup_s"], v["cpu_baseline"]["parity"]["max_rel_err"]))
print("sweep", d["scaling_sweep"]["value"], "stepwise", d["e2e_stepwise"]["value"], d["e2e_stepwise_resident"]["value"])
PY
