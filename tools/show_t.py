import json
for f in ("t1", "t2", "t3"):
    d = json.load(open("gpurun_out/%s.json" % f)); r = d["roofline"]
    print(f, round(d["ms_per_step"], 3), round(r["frac"], 4), round(r["pipeline_frac"], 4), {k: round(v, 3) for k, v in r["kernel_ms_per_step"].items()})
