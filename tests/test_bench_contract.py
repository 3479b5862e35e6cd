"""bench.py output contract, checked on FRESH output of both arms: the reference arm (CPU oracle port on a bounded
sample) runs without a GPU; the GPU arm (-m gpu) must carry every key the driver and the judge read, including the
in-run parity block, the C3/C4/C5 blocks and the C5 strong-scaling block."""
import json
import os
import subprocess
import sys

import pytest

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


def _run(args, env=None, timeout=1500):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                         env=env, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, 'bench.py must print exactly one line on stdout'
    return json.loads(lines[0])


def test_reference_arm_runs_on_the_cpu_and_prints_one_json_line():
    d = _run(['--impl', 'reference', '--steps', '2', '--warmup', '1'], env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
    _check_common(d)
    assert d['impl'] == 'reference' and d['value'] > 0 and d['cpu_baseline']['value'] == d['value']
    assert d['steps'] == 2 and d['warmup'] == 1          # K and W are honoured (the driver compares them)
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # same `config` as the GPU arm prints for the same command line
    sys.path.insert(0, ROOT)
    import bench
    from gpdoemd_b200 import synthetic
    assert d['config'] == bench.config_dict(dict(synthetic.CONFIGS['C2']), 'C2', 1)


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''


@pytest.mark.gpu
def test_gpu_arm_fresh_line_has_every_contract_key():
    d = _run(['--steps', '3', '--warmup', '3', '--extra-steps', '1'])
    _check_common(d)
    assert d['n_gpus'] == 1 and d['warmup'] >= 3 and d['steps'] == 3 and d['gpu_launches'] > 0
    r = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'pipeline_frac', 'configs'} <= set(r)
    assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    assert 0.5 < r['frac'] < 1.0 and 0.5 < r['pipeline_frac'] < 1.0
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert d['e2e']['value'] != d['value']
    assert 'C2' in d['config']['workload'] and 'N=100000' in d['config']['workload']
    # parity of the timed configuration, measured in this very run
    par = d['cpu_baseline']['parity']
    assert par['ok'] and par['argmax_match'] and par['max_rel_err'] < 1e-7 and par['n'] >= 256
    # the big configs: driver-visible roofline + parity + CPU baseline each
    for name in ('C3', 'C4', 'C5'):
        e = r['configs'][name]
        assert name in e['workload'] and e['value'] > 0
        assert 0.5 < e['roofline']['frac'] < 1.0 and 0.5 < e['roofline']['pipeline_frac'] < 1.0
        assert e['cpu_baseline']['parity']['ok'], (name, e['cpu_baseline']['parity'])
    sw = d['roofline']['scaling_sweep']
    sys.path.insert(0, ROOT)
    import bench
    from gpdoemd_b200 import synthetic
    assert d['config'] == bench.config_dict(dict(synthetic.CONFIGS['C2']), 'C2', 1)      # same dict as the reference arm
    assert sw['scaling'] == 'strong' and sw['config_id'] == 'C5' and sw['n_gpus'] == 1 and sw['value'] > 0
