"""CPU, only where /root/reference is mounted: the oracle restatement equals the reference's own
code (run behind oracle/ref_shim.py) bit for bit, on ragged / permuted / goal-conditioned batches."""
import pytest
import torch

from oracle import ref_shim
from oracle.prosim_oracle import ProSimOracle
from prosim_b200 import synthetic, weights

pytestmark = pytest.mark.ref_tree


ALL = ('goal', 'v_action_tag', 'drag_point')      # PROMPT.CONDITION.TYPES of prosim_demo/cfg/no_text.yaml


@pytest.mark.parametrize('goal', [False, True, ALL])
def test_param_table_equals_live_reference(goal):
    m, _ = ref_shim.build_reference_model(weights.cond_types(goal))
    ref = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert ref == [(n, tuple(s)) for n, s, _ in weights.param_specs(goal)]
    assert sum(p.numel() for p in m.parameters()) == {False: 10454196, True: 11314740, ALL: 11399220}[goal]


@pytest.mark.parametrize('goal', [False, True, ALL, ('drag_point', 'v_action_tag')])
def test_oracle_bit_equal_to_reference(goal):
    m, _ = ref_shim.build_reference_model(weights.cond_types(goal))
    sd = weights.random_state_dict(0, goal)
    m.load_state_dict(sd)
    kw = dict(agents_per_scene=[20, 13], map_per_scene=[48, 37], steps=30, goal=bool(goal), permute_obs=True,
              tags='v_action_tag' in weights.cond_types(goal), drag='drag_point' in weights.cond_types(goal))
    b_ref, b_orc = synthetic.make_batch(**kw), synthetic.make_batch(**kw)
    with torch.no_grad():
        ref = m.forward(b_ref, 'val')['motion_pred']
    orc = ProSimOracle(sd, goal).forward(b_orc)['motion_pred']
    for k in ('motion_pred', 'motion_prob', 'reconst_pred'):
        assert torch.equal(ref[k], orc[k]), k
    assert ref['pair_names'] == orc['pair_names']
    assert list(ref['rollout_trajs']) == list(orc['rollout_trajs'])
    for name, r in ref['rollout_trajs'].items():
        for key, val in r.items():
            assert torch.equal(val, orc['rollout_trajs'][name][key]), (name, key)
    for t in b_ref.extras['fut_obs'].keys():
        for key in ('input', 'mask', 'position', 'heading'):
            a, b = b_ref.extras['fut_obs'][t][key], b_orc.extras['fut_obs'][t][key]
            assert torch.equal(torch.nan_to_num(a.float()), torch.nan_to_num(b.float())), (t, key)


@pytest.mark.parametrize('fusion,attn', [('mlp', False), ('replace', True), ('mlp', True)])
def test_oracle_bit_equal_to_reference_obs_update_variants(fusion, attn):
    """MODEL.OBS_UPDATE.{FUSION: 'mlp', ATTN_UPDATE: True} (scene_encoder/attn_fusion.py:136-203): the per-tick scene update
    with the old / new token MLP and with the re-run agent / map->agent attention, tier B == the reference's own code."""
    opts = ['MODEL.OBS_UPDATE.FUSION', fusion, 'MODEL.OBS_UPDATE.ATTN_UPDATE', str(attn)]
    m, cfg = ref_shim.build_reference_model((), opts=opts)
    assert cfg.MODEL.OBS_UPDATE.FUSION == fusion and bool(cfg.MODEL.OBS_UPDATE.ATTN_UPDATE) == attn
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(n, tuple(sh)) for n, sh, _ in weights.param_specs(False, obs_fusion=fusion)]
    sd = weights.random_state_dict(0, obs_fusion=fusion)
    m.load_state_dict(sd)
    kw = dict(agents_per_scene=[20, 13], map_per_scene=[48, 37], steps=40, permute_obs=True)
    with torch.no_grad():
        ref = m.forward(synthetic.make_batch(**kw), 'val')['motion_pred']
    orc = ProSimOracle(sd)
    assert orc.obs_fusion == fusion
    orc.attn_update = attn
    out = orc.forward(synthetic.make_batch(**kw))['motion_pred']
    assert torch.equal(ref['motion_pred'], out['motion_pred'])
    for name, r in ref['rollout_trajs'].items():
        assert torch.equal(r['traj'], out['rollout_trajs'][name]['traj']), name
    plain = ProSimOracle(weights.random_state_dict(0)).forward(synthetic.make_batch(**kw))['motion_pred']
    assert not torch.equal(plain['motion_pred'], out['motion_pred'])          # the variants really change the rollout


def test_oracle_bit_equal_to_reference_with_action_noise():
    """MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD > 0 (act_decoder.py:113-115): same torch generator state -> same draws
    (the oracle also consumes the degenerate TOP_K = 1 draw of traj_sam.py:313) -> bit-equal noisy rollouts."""
    m, _ = ref_shim.build_reference_model((), opts=['MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD', '0.05'])
    sd = weights.random_state_dict(0)
    m.load_state_dict(sd)
    kw = dict(agents_per_scene=[9, 6], map_per_scene=[30, 24], steps=30)
    torch.manual_seed(123)
    with torch.no_grad():
        ref = m.forward(synthetic.make_batch(**kw), 'val')['motion_pred']
    orc = ProSimOracle(sd)
    orc.noise_std = 0.05
    torch.manual_seed(123)
    out = orc.forward(synthetic.make_batch(**kw))['motion_pred']
    assert torch.equal(ref['motion_pred'], out['motion_pred'])
    quiet = ProSimOracle(sd).forward(synthetic.make_batch(**kw))['motion_pred']
    d = (out['motion_pred'][:15, 0, :, :2] - quiet['motion_pred'][:15, 0, :, :2])        # first tick: cumsum of N(0, 0.05^2)
    assert 0.02 < float(d[:, 0].std()) < 0.1 and float(d[:, -1].std()) > float(d[:, 0].std())
    for name, r in ref['rollout_trajs'].items():
        assert torch.equal(r['traj'], out['rollout_trajs'][name]['traj']), name


def test_goal_sampler_helpers_equal_reference():
    """rollout/gpu_utils.py:125-177 sample_M_goal_cond_to_batch and traj_sam.py:402-439 _select_k_emd_from_batch: same host
    generator state -> the same draws -> identical tensors (pure index bookkeeping, no device work)."""
    import copy
    ref_shim.install_shims()
    from prosim.rollout.gpu_utils import sample_M_goal_cond_to_batch as ref_sample
    from prosim_b200.model import select_k_emd_from_batch
    from prosim_b200.rollout import sample_M_goal_cond_to_batch
    batch = synthetic.make_batch(n_scenes=1, n_agents=9, n_map=12, steps=20, goal=True)
    ids = batch.extras['prompt']['motion_pred']['agent_ids'][0]
    g = torch.Generator().manual_seed(3)
    K = 6
    sample = {'motion_pred': {'pair_names': [f'0-{a}-0' for a in ids[:7]],          # two agents without predictions
                              'goal_point': torch.randn(7, K, 2, generator=g) * 20, 'goal_prob': torch.rand(7, K, generator=g)}}
    sample['motion_pred']['goal_point'][2] *= 0.01                                     # snapped to (0, 0): inside smooth_dist
    b1, b2 = copy.deepcopy(batch), copy.deepcopy(batch)
    torch.manual_seed(11)
    b1 = ref_sample(b1, copy.deepcopy(sample), top_K=3, M=4, stop_smooth_num=5.0)
    torch.manual_seed(11)
    b2 = sample_M_goal_cond_to_batch(b2, copy.deepcopy(sample), 3, 4, stop_smooth_num=5.0)
    c1, c2 = b1.extras['condition'].all_cond['goal'], b2.extras['condition'].all_cond['goal']
    assert sorted(c1.keys()) == sorted(c2.keys())
    for k in ('input', 'mask', 'prompt_idx', 'prompt_mask'):
        assert torch.equal(c1[k], c2[k]), k
    assert (c2['input'][:, 2, :2] == 0).all() and c2['input'].shape == (4, 7, 3)
    # _select_k_emd_from_batch on K goal-conditioned policy tokens per agent
    m, _ = ref_shim.build_reference_model(())
    emds = {'emd': torch.randn(2, 5, K, 128, generator=g), 'goal_prob': torch.rand(2, 5, K, generator=g),
            'goal_point': torch.randn(2, 5, K, 2, generator=g), 'agent_type': torch.ones(2, 5, dtype=torch.long)}
    m.mode = 'val'
    torch.manual_seed(5)
    r = m._select_k_emd_from_batch(dict(emds), None)
    torch.manual_seed(5)
    mine = select_k_emd_from_batch(dict(emds), None, 'val', m.rollout_top_k, m.rollout_top_k_train, torch.device('cpu'))
    for k in ('emd', 'goal', 'select_idx'):
        assert torch.equal(r[k], mine[k]), k
    m.rollout_top_k = 3
    torch.manual_seed(6)
    r = m._select_k_emd_from_batch(dict(emds), None)
    torch.manual_seed(6)
    mine = select_k_emd_from_batch(dict(emds), None, 'val', 3, 1, torch.device('cpu'))
    assert torch.equal(r['emd'], mine['emd']) and mine['emd'].shape == (2, 5, 128)
    assert not torch.equal(mine['select_idx'], emds['goal_prob'].argmax(-1))            # top-3 sampling really samples


def test_world_transform_equals_reference():
    """oracle.rollout_trajs_in_world == the reference's obtain_rollout_trajs_in_world (rollout/gpu_utils.py:230-281)."""
    import math
    import types
    import numpy as np
    ref_shim.install_shims()
    from prosim.rollout.gpu_utils import obtain_rollout_trajs_in_world
    from oracle.prosim_oracle import rollout_trajs_in_world
    sd = weights.random_state_dict(0)
    out = ProSimOracle(sd).forward(synthetic.make_batch(agents_per_scene=[6, 4], map_per_scene=[20, 16], steps=20))
    th = -1.1
    tf = torch.tensor([[math.cos(th), -math.sin(th), 33.0], [math.sin(th), math.cos(th), 7.5], [0.0, 0.0, 1.0]])
    fake = types.SimpleNamespace(centered_world_from_agent_tf=tf[None].repeat(2, 1, 1))
    trajs_M, ids_M = obtain_rollout_trajs_in_world(fake, out)
    mine, names = rollout_trajs_in_world(out['motion_pred'], tf)
    assert np.array_equal(np.concatenate(trajs_M, axis=0), mine.numpy())
    assert [i for ids in ids_M for i in ids] == [n.split('-')[1] for n in names]
