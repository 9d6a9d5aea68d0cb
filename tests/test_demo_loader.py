"""BASELINE configs[0], literal: a scene of the reference's demo_dataset through prosim_b200/demo_loader.py (SURVEY.md section
8f-1).  CPU tests: the loader's invariants on the shipped files (needs the mounted reference tree), the fixture round trip,
and the oracle on the committed fixture against the reference's own rollout of that batch; GPU test: ProSimB200 on it."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle.prosim_oracle import ProSimOracle
from prosim_b200 import demo_loader, weights
from tests.helpers import DEMO_CASES, demo_batch, load_golden, per_tick_max, stack_rollout

NAME = 'cfg1_demo_scene6_a16_s20'


@pytest.mark.ref_tree
def test_loader_on_the_shipped_demo_scenes():
    root = os.path.join(ref_shim.REF_ROOT, 'demo_dataset')
    if not os.path.isdir(root):
        pytest.skip('demo_dataset is not part of the staged reference copy')
    b = demo_loader.load_demo_scene(root, **DEMO_CASES[NAME])
    ex = b.extras
    obs, mp, pr = ex['init_obs'], ex['init_map'], ex['prompt']['motion_pred']
    A, M = obs['input'].shape[1], mp['input'].shape[1]
    assert obs['input'].shape == (1, A, 11, 24) and A == 16 and mp['input'].shape == (1, M, 19, 11) and M >= 256
    assert ex['all_t_indices'].tolist() == [0, 10] and list(ex['fut_obs'].keys()) == [10]
    assert not torch.isnan(obs['input'][obs['mask']]).any() and not torch.isnan(mp['input'][mp['mask'][..., None].expand_as(mp['input'])]).any()
    # every agent is observed in its own last-step frame: x = y = 0, sin = 0, cos = 1 at the last history step
    last = obs['input'][0, :, -1]
    assert last[:, :3].abs().max() == 0 and (last[:, 3] == 1).all()
    assert (obs['input'][0, :, :, 13:24] == torch.eye(11)).all() and (obs['input'][0, :, 0, 10:13].sum(-1) == 1).all()
    # the centre agent is the ego and sits at the origin of the centred frame
    assert obs['agent_ids'][0][0] == 'ego' and obs['position'][0, 0].abs().max() < 1e-6 and abs(float(obs['heading'][0, 0])) < 1e-6
    # polylines in symmetric coordinates: first start point and last end point mirror each other on the x axis
    cnt = mp['mask'][0].sum(-1)
    v = mp['input'][0]
    start, end = v[:, 0, :2], v[torch.arange(M), cnt - 1, 2:4]
    assert (start + end).abs().max() < 1e-3 and end[:, 1].abs().max() < 1e-3 and (end[:, 0] >= 0).all()
    assert set(v[..., 4][mp['mask'][0]].unique().tolist()) <= {1.0, 2.0, 3.0}
    assert sorted(pr['agent_ids'][0]) == sorted(set(pr['agent_ids'][0])) and set(pr['agent_ids'][0]) <= set(obs['agent_ids'][0])
    assert pr['prompt'].shape == (1, len(pr['agent_ids'][0]), 7)
    # the loaded batch IS the committed fixture
    fix = demo_batch(NAME)
    for key in ('init_obs', 'init_map'):
        for k in ('input', 'mask', 'position', 'heading'):
            assert torch.equal(torch.nan_to_num(ex[key][k].float()), torch.nan_to_num(fix.extras[key][k].float())), (key, k)
    assert fix.extras['prompt']['motion_pred']['agent_ids'] == pr['agent_ids']
    other = demo_loader.load_demo_scene(root, scene='scene_11', ts=10, steps=20, max_agents=16)
    assert other.extras['init_obs']['input'].shape[1] == 8          # a scene with fewer agents than the cap


def test_oracle_on_the_demo_fixture_matches_the_reference_rollout():
    gold = load_golden(NAME)
    out = ProSimOracle(weights.random_state_dict(0)).forward(demo_batch(NAME))['motion_pred']
    names, traj, _ = stack_rollout(out)
    assert names == gold['agent_names'].tolist() and out['pair_names'] == gold['pair_names'].tolist()
    gap_ref = per_tick_max(gold['traj'], gold['traj64'])
    assert np.all(per_tick_max(traj, gold['traj64']) <= gap_ref + 1e-4)


@pytest.mark.gpu
def test_gpu_rollout_of_the_demo_scene_matches_the_reference_golden():
    from prosim_b200.model import ProSimB200
    gold = load_golden(NAME)
    model = ProSimB200(state_dict=weights.random_state_dict(0), device='cuda')
    with torch.no_grad():
        out = model.forward(demo_batch(NAME).to('cuda'), 'val')['motion_pred']
    names, traj, vel = stack_rollout(out)
    assert names == gold['agent_names'].tolist() and out['pair_names'] == gold['pair_names'].tolist()
    P = len(names)
    # real data is worse conditioned than the synthetic scenes (the reference's own fp32 run is 1e-4 from its fp64 run after ONE
    # tick): every gate is relative to the reference's fp64 evaluation
    mp, g32, g64 = out['motion_pred'].cpu().numpy()[:P], gold['motion_pred'][:P], gold['motion_pred64'][:P]
    assert np.abs(mp - g64).max() <= np.abs(g32 - g64).max() + 1e-5
    gap_ref, gap_gpu = per_tick_max(gold['traj'], gold['traj64']), per_tick_max(traj, gold['traj64'])
    print('ref32-vs-64', gap_ref, 'gpu-vs-64', gap_gpu)
    assert np.all(gap_gpu <= gap_ref + 1e-4)
    assert np.all(per_tick_max(vel, gold['vel64']) <= per_tick_max(gold['vel'], gold['vel64']) + 1e-4)
