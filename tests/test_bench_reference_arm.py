"""bench.py --impl reference: the reference's own CPU implementation of the path, timed on the host cores (no GPU involved).
Checks the JSON contract of that line (impl, metric / unit of our arm, cpu_baseline describing the run, e2e with zero copies)
and that the timed code is the UNMODIFIED reference whenever a copy of it is present (/root/reference here, the staged
baseline/_ref on the GPU box) - the oracle port is only the fallback."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('workload,metric_word', [('cfg3', 'posterior+EI'), ('cfg4', 'NLML')])
def test_reference_arm_prints_one_contract_line(workload, metric_word):
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--workload', workload, '--steps', '1',
                          '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and metric_word in d['metric'] and d['unit'] == 'evals/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['gpu_launches'] == 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['value'] == d['value'] and cb['cores'] >= 1 and 'sample' in cb
    sys.path.insert(0, REPO)
    from oracle import ref_cpu_arm
    assert cb['kind'] == ('reference' if ref_cpu_arm.available() else 'port')
    with open(os.path.join(REPO, 'BASELINE.json')) as fh:
        if workload == 'cfg3':
            assert d['metric'] == json.load(fh)['metric']
