"""Repository contracts: the product package never touches oracle/, bench.py's
reference arm prints the agreed JSON line, entry points exist."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_the_oracle():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|from\s+\.\.?oracle|oracle\.', re.M)
    for base, _, files in os.walk(os.path.join(ROOT, 'omg_tools_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.h')):
                src = open(os.path.join(base, f)).read()
                assert not pat.search(src), os.path.join(base, f)


def test_bench_reference_arm_json_line():
    out = subprocess.check_output(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
         '--warmup', '0', '--cpu-sample', '4'], cwd=ROOT, timeout=600,
        stderr=subprocess.DEVNULL).decode().strip().splitlines()[-1]
    line = json.loads(out)
    assert line['impl'] == 'reference' and line['metric'] == 'mpc_solves_per_sec'
    assert line['unit'] == 'solves/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['steps'] == 1
    cb = line['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == line['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config']


def test_entry_points_exist():
    import __graft_entry__ as ge
    assert callable(ge.build) and callable(ge.smoke)
    for name in ('DESIGN.md', 'INTEGRATION.md', 'include/omg_b200.h', 'bench.py',
                 'oracle/ipm.c', 'oracle/ipm_ref.py', 'tests/golden/spline_golden.npz',
                 'tests/golden/p2p_golden.npz'):
        assert os.path.exists(os.path.join(ROOT, name)), name
