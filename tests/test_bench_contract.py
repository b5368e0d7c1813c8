"""bench.py's driver contract, as far as a box without a GPU can check it: the reference arm (`--impl reference`, the
CPU port timed on the host cores) prints ONE JSON line with the agreed keys, and the product arm fails loudly instead of
falling back to the CPU when there is no device."""
import json
import pathlib
import subprocess
import sys

import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]


def run_bench(*args):
    return subprocess.run([sys.executable, str(ROOT / 'bench.py'), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '1', '--shard', '512')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'reads classified/sec' and d['unit'] == 'reads/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] >= 1 and d['data'] == 'synthetic'
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'model' not in d['config']
    cpu = d['cpu_baseline']
    assert cpu['kind'] in ('port', 'reference') and cpu['cores'] >= 1 and cpu['sample'] and cpu['value'] == d['value']
    e2e = d['e2e']
    assert e2e['value'] == d['value'] and e2e['unit'] == d['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='a GPU is present: the product arm would simply run')
def test_product_arm_fails_loudly_without_a_gpu():
    r = run_bench('--steps', '1', '--warmup', '1', '--shard', '512')
    assert r.returncode != 0
    assert not any(l.lstrip().startswith('{') for l in r.stdout.splitlines()), 'no bench line may be printed without a GPU'
