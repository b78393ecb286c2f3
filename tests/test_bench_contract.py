"""bench.py's reference arm (`--impl reference`: the oracle port timed on the host cores) runs
without a GPU; check the JSON line it prints against the driver's contract."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True,
                         timeout=900, cwd=ROOT,
                         env=dict(os.environ, TB_BENCH_CPU_ENVS='32'))   # small sample: this is a format check
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.strip().splitlines() if l.startswith('{')]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['higher_is_better'] is True
    assert line['metric'].startswith('env-steps/sec') and line['unit'] == 'env-steps/s'
    assert line['n_gpus'] == 1 and line['steps'] == 1 and line['value'] > 0
    base = line['cpu_baseline']
    assert base['kind'] in ('port', 'reference') and base['cores'] >= 1 and base['value'] == line['value']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['h2d_bytes_per_step'] == 0 == e2e['d2h_bytes_per_step']
    assert 'workload' in line['config']
    # thread-pinned sweep incl. the forked --parallel C --sequential N/C grid
    grid = base['grid']
    assert any(not g['mode'].startswith('--parallel 1 ') for g in grid) and all(g['torch_threads'] >= 1 for g in grid)
