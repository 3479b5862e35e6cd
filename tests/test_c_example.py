"""examples/score_c_abi.c: the design step from plain C through the C ABI alone (what a cgo / JNI / N-API binding or a
C host would call).  CPU: it compiles, links against the in-tree library and fails loudly without a GPU.  GPU: it runs,
and with two GPUs every rank thread reports the same global argmax as the one-GPU run."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / 'score_c_abi')
    lib = os.path.join(ROOT, 'gpdoemd_b200')
    cmd = ['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
           os.path.join(ROOT, 'examples', 'score_c_abi.c'), '-L', lib, '-lgpdoemd_b200', '-Wl,-rpath,' + lib, '-lm',
           '-lpthread', '-o', exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    out = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
    assert out.returncode != 0 and 'gpd_create' in out.stderr          # no CPU fallback


@pytest.mark.gpu
def test_c_example_runs(tmp_path):
    import torch
    exe = _build(tmp_path)
    one = subprocess.run([exe, '1'], capture_output=True, text=True, timeout=300)
    assert one.returncode == 0, one.stderr
    m = re.search(r'BH = ([0-9.e+-]+), index (\d+)', one.stdout)
    assert m, one.stdout
    if torch.cuda.device_count() >= 2:
        two = subprocess.run([exe, '2'], capture_output=True, text=True, timeout=300)
        assert two.returncode == 0, two.stderr
        got = re.findall(r'BH = ([0-9.e+-]+), index (\d+)', two.stdout)
        assert len(got) == 2 and all(g == (m.group(1), m.group(2)) for g in got), (one.stdout, two.stdout)
