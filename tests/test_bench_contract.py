"""bench.py prints exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'cpu_baseline'}


def _run(args, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines            # ONE line, nothing else on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--workload', '512x128'])
    assert d['impl'] == 'reference' and BASE_KEYS <= set(d)
    # like-for-like: the reference arm prints the very `config` the B200 arm prints for the same workload
    sys.path.insert(0, ROOT)
    import bench
    assert d['config'] == bench.workload_config('mesh', 512, 128, 1)
    assert d['cpu_baseline_splu']['value'] > 0 and d['run']['pcg_iterations_per_step'] > 0
    assert d['value'] > 0 and d['unit'] == 'cell-timesteps/s' and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] == 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['dtype'] == 'f64' and d['vs_baseline'] is None


@pytest.mark.gpu
def test_gpu_arm_line_small_workload():
    d = _run(['--workload', '256x128', '--steps', '3', '--warmup', '3', '--no-cpu-baseline'])
    assert BASE_KEYS | {'roofline', 'clocks'} <= set(d)
    assert d['n_gpus'] == 1 and d['steps'] == 3 and d['value'] > 0 and d['gpu_launches'] > 0
    assert d['e2e']['h2d_bytes_per_step'] == d['e2e']['d2h_bytes_per_step'] == 8 * 256 * 128 and d['e2e']['value'] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert d['config'] == bench.workload_config('mesh', 256, 128, 1)
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and r['peak'] > 0 and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
