"""Every bench workload runs end to end at a reduced path count and prints the
contract's JSON line (keeps `bench.py` in step with the package's internals)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
        'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
        'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline']


@pytest.mark.parametrize('workload,paths', [('c1', 20000), ('c2', 200000), ('c2_qe', 200000),
                                            ('c3', 200000), ('c4', 20000), ('c5', 100000),
                                            ('materialise', 100000)])
def test_bench_workload_runs(workload, paths):
  out = subprocess.run(
      [sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', workload,
       '--paths', str(paths), '--steps', '1', '--warmup', '1'],
      capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  line = [l for l in out.stdout.splitlines() if l.startswith('{')][-1]
  d = json.loads(line)
  for k in KEYS:
    assert k in d, k
  assert d['value'] > 0 and d['gpu_launches'] > 0
  assert d['e2e']['value'] > 0 and d['roofline']['frac'] > 0
