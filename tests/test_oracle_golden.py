"""CPU: the self-contained oracle (oracle/prosim_oracle.py) against vectors generated from the
reference's own code (tests/golden/make_golden.py), and the weight generator's determinism."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.prosim_oracle import ProSimOracle
from prosim_b200 import synthetic, weights
from tests.helpers import ALL_COND, CASES, GOLDEN, cond_suffix, load_golden, per_tick_max, stack_rollout, to_double


@pytest.mark.parametrize('goal', [False, True, ALL_COND])
def test_state_dict_table_matches_reference_keys(goal):
    with open(os.path.join(GOLDEN, f'state_dict_keys{cond_suffix(goal)}.json')) as f:
        ref = json.load(f)
    mine = [[n, list(s)] for n, s, _ in weights.param_specs(goal)]
    assert mine == ref['keys']
    sd = weights.random_state_dict(0, goal)
    assert float(sum(v.double().sum() for v in sd.values())) == pytest.approx(ref['checksum'], rel=1e-12)
    assert float(sum(v.double().abs().sum() for v in sd.values())) == pytest.approx(ref['abs_checksum'], rel=1e-12)


@pytest.mark.parametrize('name', [n for n in CASES if 'cfg3' not in n and 'cfg4' not in n] + ['cfg3_a128_m512_s80'])
def test_oracle_matches_reference_golden(name):
    kw, goal = CASES[name]
    gold = load_golden(name)
    out = ProSimOracle(weights.random_state_dict(0, goal), goal).forward(synthetic.make_batch(**kw))['motion_pred']
    names, traj, vel = stack_rollout(out)
    assert names == gold['agent_names'].tolist()
    assert out['pair_names'] == gold['pair_names'].tolist()
    # same machine + same torch build -> bit equal; other host CPUs may round GEMMs differently, so the
    # gate is the closed-loop protocol of SURVEY section 8d: no further from fp64 than the reference + 1e-4.
    gap32 = per_tick_max(gold['traj'], gold['traj64'])
    mine64 = per_tick_max(traj, gold['traj64'])
    assert np.all(mine64 <= gap32 + 1e-4), (mine64, gap32)
    assert np.abs(out['motion_pred'].numpy()[: len(names)] - gold['motion_pred'][: len(names)]).max() < 1e-5
    assert per_tick_max(traj, gold['traj'])[0] < 1e-5
    assert per_tick_max(vel, gold['vel'])[0] < 1e-5
    assert np.array_equal(out['motion_prob'].numpy(), gold['motion_prob'])


def test_oracle_fp64_matches_reference_fp64():
    name = 'ragged_b3_s30'
    kw, goal = CASES[name]
    gold = load_golden(name)
    sd = {k: v.double() for k, v in weights.random_state_dict(0, goal).items()}
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        out = ProSimOracle(sd, goal, dtype=torch.float64).forward(to_double(synthetic.make_batch(**kw)))['motion_pred']
    finally:
        torch.set_default_dtype(prev)
    _, traj, _ = stack_rollout(out)
    assert np.abs(traj - gold['traj64']).max() < 1e-9
