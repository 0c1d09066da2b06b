"""bench.py on a box without a GPU: the byte model of SURVEY.md section 8(d), the mesh
sizes, and the --impl reference arm's JSON line (the driver parses it)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from p1afempy_b200 import meshgen  # noqa: E402


@pytest.mark.parametrize('level', [0, 1, 3, 5])
def test_mesh_counts_match_the_generator(level):
    c, e, _ = meshgen.unit_square(level)
    m, n, nnz = bench.mesh_counts(level)
    assert (m, n) == (e.shape[0], c.shape[0])
    # nnz = nodes + 2 * edges; edges = nodes + triangles - 1 on a simply connected mesh
    assert nnz == n + 2 * (n + m - 1)


def test_algorithmic_bytes_follow_the_survey():
    m, n, nnz = bench.mesh_counts(12)
    assert (m, n, nnz) == (67108864, 33562625, 234905601)
    stiff = bench.algorithmic_bytes('stiffness', m, n, nnz, 0)
    assert stiff == 12 * m + 16 * n + 12 * nnz + 4 * (n + 1)
    assert abs(stiff / m - 64.0) < 0.01                      # 64 B per triangle on S(k)
    assert bench.algorithmic_bytes('mass', m, n, nnz, 0) == stiff
    assert bench.algorithmic_bytes('general', m, n, nnz, 16) == stiff + 4 * 16 * 8 * m
    assert bench.algorithmic_bytes('weighted', m, n, nnz, 16) == stiff + 8 * n + 16 * 8 * m
    assert bench.algorithmic_bytes('load', m, n, nnz, 16) == 12 * m + 16 * n + 8 * n + 16 * 8 * m
    assert abs(bench.algorithmic_bytes('eta', m, n, nnz, 0) / m - 40.0) < 0.01


def test_reference_level_respects_the_budget():
    # one call of S(12) takes ~27 s on one core: 23 calls do not fit 300 s, 4 do
    assert bench.reference_level('stiffness', 12, 23, 300.) < 12
    assert bench.reference_level('stiffness', 5, 23, 300.) == 5


@pytest.mark.parametrize('op', ['stiffness', 'load'])
def test_reference_arm_prints_one_json_line(op):
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference',
                          '--op', op, '--level', '5', '--steps', '2', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC[op] and d['unit'] == bench.UNIT
    assert d['steps'] == 2 and d['higher_is_better'] is True and d['dtype'] == 'f64'
    assert d['value'] > 0 and np.isclose(d['value'],
                                         bench.mesh_counts(5)[0] / d['ms_per_step'] / 1e3, rtol=1e-6)
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] == 1
    assert d['e2e'] == {'value': d['value'], 'unit': bench.UNIT, 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert d['config']['same_mesh_as_gpu_arm'] is True


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference',
                          '--level', '5', '--steps', '1', '--gpus', '2'],
                         capture_output=True, text=True, timeout=300, cwd=REPO, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('this box has a GPU')
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--steps', '1',
                          '--level', '5'], capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode != 0 and 'needs a GPU' in (out.stderr + out.stdout)
