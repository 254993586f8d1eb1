"""bench.py's reference arm runs on the CPU: check the JSON-line contract of the driver on the
reduced workload (one line on stdout, required keys, config shared with the product arm)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload',
           'sr3_48', '--steps', '2', '--warmup', '1'] + list(extra)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    return [l for l in res.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'cg_iterations_per_sec'
    assert d['unit'] == 'CG-it/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['vs_baseline'] is None and d['dtype'] == 'f32' and d['data'] == 'synthetic'
    assert d['n_gpus'] == 1 and d['steps'] == 2
    assert d['config']['workload'].startswith('sr3_48: 3-channel thick-slice super-resolution')
    assert d['config']['channels'] == 3 and d['warmup'] >= 1
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_other_ranks_print_nothing_and_all_cores_are_used():
    """Under torchrun every rank gets OMP_NUM_THREADS=1: rank 0 still uses the host's cores,
    the other ranks exit 0 without work."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1', OMP_NUM_THREADS='1')
    assert _run('--gpus', '2', env=env) == []
    env.update(RANK='0', LOCAL_RANK='0')
    d = json.loads(_run('--gpus', '2', env=env)[0])
    assert d['n_gpus'] == 2
    assert d['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))
