"""bench.py output contract: the committed round bench line carries every key the driver and the judge read, and the
reference arm (CPU oracle port on a bounded sample) runs without a GPU and prints the same schema."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config'}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d['unit'] == 'designs/s' and d['higher_is_better'] is True and d['scaling'] == 'weak'
    assert d['dtype'] == 'f64' and d['data'] == 'synthetic' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(d['e2e'])
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(d['cpu_baseline'])
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1


def test_committed_bench_line_has_every_contract_key():
    d = json.load(open(os.path.join(ROOT, 'profiles', 'r01_bench_c2.json')))
    _check_common(d)
    assert d['n_gpus'] == 1 and d['warmup'] >= 3 and d['gpu_launches'] > 0
    r = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r)
    assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert d['e2e']['value'] != d['value']
    assert 'C2' in d['config']['workload'] and 'N=100000' in d['config']['workload']


def test_reference_arm_runs_on_the_cpu_and_prints_one_json_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d['impl'] == 'reference' and d['value'] > 0 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
