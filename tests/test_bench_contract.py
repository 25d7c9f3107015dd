"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm prints ONE JSON
line with the agreed keys (it times the CPU port of the reference loop, so it runs anywhere), and the product arm
fails loudly -- no CPU fallback -- when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run('--impl', 'reference', '--steps', '2', '--warmup', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['metric'] == 'classified atom-pairs/s' and d['unit'] == 'pairs/s' and d['higher_is_better'] is True
    assert d['steps'] == 2 and d['warmup'] == 1 and d['value'] > 0 and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'configs[2]' in d['config']['workload']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '1'], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_fails_loudly_without_a_device():
    from arpeggio_b200 import _lib
    if _lib.lib().arp_device_count() > 0:
        pytest.skip('a CUDA device is present')
    r = _run('--steps', '1', '--warmup', '1', '--no-cpu', timeout=300)
    assert r.returncode != 0
    assert 'no CUDA device' in (r.stderr + r.stdout) or 'no CPU fallback' in (r.stderr + r.stdout)
    assert not any(ln.startswith('{') for ln in r.stdout.splitlines())


@pytest.mark.gpu
def test_product_arm_prints_the_contract_line():
    r = _run('--steps', '5', '--warmup', '3', '--batch-structures', '12', '--large-atoms', '150000', timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert 'impl' not in d and d['metric'] == 'classified atom-pairs/s' and d['n_gpus'] == 1 and d['steps'] == 5
    assert d['value'] > 1e9 and d['scaling'] == 'weak' and d['data'] == 'synthetic' and d['vs_baseline'] is None
    assert d['gpu_launches'] == d['kernels_per_step'] * d['steps'] and d['kernels_per_step'] >= 3
    e = d['e2e']
    # the end-to-end stream is the packed sorted one: 4 bytes per record + 4 per atom (+ 8); the inputs travel in wire form
    n, atoms = d['config']['pairs_per_structure'], d['config']['atoms_per_gpu']
    assert 0 < e['value'] < d['value'] and e['sorted'] is True
    assert e['d2h_bytes_per_step'] == 4 * n + 4 * (atoms + 2)
    legs = e['legs']
    assert 0 < e['h2d_bytes_per_step'] < legs['plain_inputs']['h2d_bytes_per_step'] and legs['plain_inputs']['value'] > 0
    assert legs['records16']['d2h_bytes_per_step'] == 16 * n and legs['records16']['value'] > 0
    assert legs['compact']['d2h_bytes_per_step'] == 8 * n + 4 * (atoms + 1) and legs['compact']['value'] > 0
    assert legs['with_distances']['d2h_bytes_per_step'] == e['d2h_bytes_per_step'] + 4 * n
    assert 0 < legs['with_distances']['value'] <= e['value'] * 1.2 and d['resident_pipelined']['value'] > 0
    assert legs['wire_h_fix']['h2d_bytes_per_step'] < e['h2d_bytes_per_step'] and legs['wire_h_fix']['value'] > 0
    assert e['pcie']['h2d_gbs'] > 1 and e['pcie']['d2h_gbs'] > 1 and e['serial_value'] > 0
    assert 0.05 < e['pcie']['e2e_fraction_of_floor'] < 1.3          # the end-to-end step against the slower copy direction alone
    rf = d['roofline']
    assert rf['bound'] == 'hbm' and rf['unit'] == 'GB/s' and abs(rf['frac'] - rf['achieved'] / rf['peak']) < 1e-9
    assert rf['algorithmic_bytes'] > 16 * d['config']['pairs_per_structure']
    assert abs(rf['kernel_ms'] - d['ms_per_step']) < 1e-12 and rf['pair_kernels']['kernel_ms'] <= rf['kernel_ms'] * 1.2
    assert rf['large']['atoms'] == 150000 and rf['large']['frac'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] > 0
    assert 'sm_mhz' in d['clocks'] and 'reasons' in d['clocks']
    assert d['batch']['structures'] == 12 and d['batch']['value'] > 0
