"""bench.py and __graft_entry__.smoke() must keep running: a small invocation of each on the GPU (the driver runs both at
round end; a host-side refactor that breaks them would otherwise only show up there)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
        'dtype', 'data', 'config', 'clocks', 'gpu_launches', 'e2e', 'roofline'}


def test_bench_line_small():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--videos', '4', '--steps', '2', '--warmup', '3',
                          '--no-cpu-baseline', '--e2e-split', '2'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert KEYS <= set(line), KEYS - set(line)
    assert line['value'] > 0 and line['e2e']['value'] > 0 and line['gpu_launches'] > 100
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(line['roofline'])
    assert line['e2e']['h2d_bytes_per_step'] == 4 * 4096 * 2048 * 4


def test_smoke_entry():
    out = subprocess.run([sys.executable, '-c', 'import __graft_entry__ as g; g.smoke()'], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
