"""BASELINE configs at (or near) their full candidate counts on one GPU: 1 warm-up + 2 timed resident steps each, with the
per-family device times (builder-run evidence next to the driver-run reduced-N blocks of bench.py)."""
import json
import sys
import numpy as np
sys.path.insert(0, '.')
import bench
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic

ctx = g.get_context(0)
ctx.set_scratch_bytes(64 << 30)
peak = ctx.probe_fp64_peak()
out = {}
for name, N in [(a.split('=')[0], int(float(a.split('=')[1]))) for a in sys.argv[1:]]:
    cfg = dict(synthetic.CONFIGS[name]); cfg['N'] = N
    p = bench.Problem(g, ctx, cfg, 0, None, N, 0, want_host=False)
    p.step_resident()
    t, res, prof = bench.timed_device(ctx, p.step_resident, 2)
    ms = t / 2
    r = bench.roofline_of(cfg, N, 2, ms, prof, peak, None)
    out[name] = {'workload': bench.workload_name(cfg, name), 'N': N, 'value': N / (ms * 1e-3), 'ms_per_step': ms,
                 'tri_frac': r['frac'], 'pipeline_frac': r['pipeline_frac'], 'kernel_ms_per_step': r['kernel_ms_per_step'],
                 'argmax': {c: int(res[c][0]) for c in p.crits}}
    print(name, json.dumps(out[name]), flush=True)
    p.close()
json.dump(out, open('gpurun_out/r02_full_size.json', 'w'), indent=1)
