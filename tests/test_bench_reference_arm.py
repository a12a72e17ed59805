"""The CPU arm of bench.py (`--impl reference`) runs without a GPU and prints the contract's JSON line: checked here end to end on a
minimal run (1 timed step), because the driver executes that arm on the GPU box and a broken line there cannot be repaired afterwards."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', SFB_REF_BUDGET_S='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '0'], capture_output=True, text=True,
                         cwd=ROOT, env=env, timeout=580)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'steps/s' and d['higher_is_better'] is True and d['scaling'] == 'weak'
    assert d['metric'].startswith('distillation-steps/sec') and d['steps'] == 2 and d['n_gpus'] == 1
    assert d['value'] > 0 and abs(d['ms_per_step'] - 1e3 / d['value']) < 1e-2 * d['ms_per_step']
    assert d['e2e'] == {'value': d['value'], 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and 1 <= cb['cores'] <= (os.cpu_count() or 1) and 'sample' in cb
    assert 'workload' in d['config']
    # `config` is the workload only and is built by the SAME function the GPU arm uses (the driver compares the two arms' configs);
    # what differs between the arms lives in `engine`
    sys.path.insert(0, ROOT)
    import bench
    assert d['config'] == bench.workload_config(1, d['config']['unet_evals_per_step_mean'])
    assert set(d['config']) == {'workload', 'parallelism', 'unet_evals_per_step_mean', 'l2', 'lpips'} and 'host' in d['engine']
    # at least one WHOLE step is really run; what is filled in from measured components is declared
    assert d['whole_steps_measured'] == 1 and d['extrapolated_steps'] == 1 and d['extrapolated'] is True
    assert d['whole_step_s']['min'] > 0 and d['components_s']['per_unet_eval'] > 0


def test_other_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith('{')]
