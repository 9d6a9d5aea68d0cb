"""GPU (-m gpu): the full closed-loop rollout through the reference-facing API (ProSimB200.forward) against
the CPU oracle and the reference-generated golden vectors.  Protocol: SURVEY.md section 8d --
(1) teacher-forced per-tick parity <= 1e-5, (2) bit-exact index bookkeeping, (3) closed-loop parity judged
against the reference's own fp64 evaluation where its fp32 noise exceeds 5e-5."""
import copy
import math

import numpy as np
import pytest
import torch

from oracle.prosim_oracle import ProSimOracle
from prosim_b200 import synthetic, weights
from tests.helpers import BENCH_CASES, CASES, edge_set, load_golden, make_case_batch, per_tick_max, stack_rollout

pytestmark = pytest.mark.gpu

_models = {}


def _model(goal):
    from prosim_b200.config import get_config
    from prosim_b200.model import ProSimB200
    if goal not in _models:
        cfg = get_config(opts=['PROMPT.CONDITION.TYPES', list(weights.cond_types(goal))] if goal else None)
        _models[goal] = ProSimB200(cfg, weights.random_state_dict(0, goal), device='cuda')
    return _models[goal]


def _run_gpu(kw, goal, keep_edges=False, name=None):
    """name: a golden case -- its inputs are checked against the checksum recorded with the golden (helpers.make_case_batch)."""
    model = _model(goal)
    model.keep_tick_edges = keep_edges
    batch = (make_case_batch(name) if name else synthetic.make_batch(**kw)).to('cuda')
    with torch.no_grad():
        out = model.forward(batch, 'val')['motion_pred']
    torch.cuda.synchronize()
    model.keep_tick_edges = False
    return out, batch


@pytest.mark.parametrize('name', list(CASES))
def test_closed_loop_matches_reference_golden(name):
    kw, goal = CASES[name]
    gold = load_golden(name)
    out, _ = _run_gpu(kw, goal, name=name)
    names, traj, vel = stack_rollout(out)
    # bit-exact bookkeeping: agent order, pair names, shapes, probabilities
    assert names == gold['agent_names'].tolist()
    assert out['pair_names'] == gold['pair_names'].tolist()
    assert tuple(out['motion_pred'].shape) == gold['motion_pred'].shape
    assert np.array_equal(out['motion_prob'].cpu().numpy(), gold['motion_prob'])
    P = len(names)
    # first tick is open loop: identical inputs -> 1e-5 gate on every output
    assert np.abs(out['motion_pred'].cpu().numpy()[:P] - gold['motion_pred'][:P]).max() < 1e-5
    assert np.abs(out['reconst_pred'].cpu().numpy() - gold['reconst_pred']).max() < 1e-5
    # closed loop: no further from the reference's fp64 evaluation than the reference's own fp32 run + 1e-4,
    # and within 1e-4 of the fp32 reference wherever the reference's own noise is below 5e-5
    gap_ref = per_tick_max(gold['traj'], gold['traj64'])
    gap_gpu = per_tick_max(traj, gold['traj64'])
    d32 = per_tick_max(traj, gold['traj'])
    print(name, 'ref32-vs-64', gap_ref, 'gpu-vs-64', gap_gpu, 'gpu-vs-ref32', d32)
    assert np.all(gap_gpu <= gap_ref + 1e-4), (gap_gpu, gap_ref)
    assert np.all(d32[gap_ref < 5e-5] <= 1e-4), (d32, gap_ref)
    assert np.all(per_tick_max(vel, gold['vel64']) <= per_tick_max(gold['vel'], gold['vel64']) + 1e-4)
    init_pos = torch.stack([out['rollout_trajs'][n]['init_pos'] for n in names]).cpu().numpy()
    assert np.array_equal(init_pos, gold['init_pos'])


def _flip_candidates(batch_cpu, gold, traj_gpu, scenes, n_agents, steps):
    """Discrete-decision bookkeeping for the closed-loop comparison ("in closed loop report any
    edge flips with the pair's distance-to-threshold").  At every tick the policy builds radius graphs from the CURRENT
    positions (agent-agent 100 m, agent-map 50 m, policy/act_decoder.py:250,259); a pair whose distance is closer to the radius
    than twice the position difference between the two runs may be an edge in one run and not in the other -- from that tick on
    the two runs are different dynamical systems and a coordinate-wise tolerance says nothing.  Returns first_flagged[scene]
    = the first tick index whose graph may differ (len(ticks) if none) and a printable report."""
    ex = batch_cpu.extras
    init_pos = gold['init_pos'].astype(np.float64).reshape(len(scenes), n_agents, 2)
    ref = gold['traj'].astype(np.float64).reshape(len(scenes), n_agents, steps, 4)
    gpu = np.asarray(traj_gpu, dtype=np.float64).reshape(len(scenes), n_agents, steps, 4)
    n_ticks = steps // 10
    first, report = {}, []
    for si, sc in enumerate(scenes):
        mpos = ex['init_map']['position'][sc, :, 0].double().numpy()
        first[sc] = n_ticks
        for k in range(n_ticks):
            if k == 0:
                p_ref = p_gpu = init_pos[si]
            else:   # traj_sam.py:213: world position = init_pos + traj_xy of the last rolled-out step
                p_ref, p_gpu = init_pos[si] + ref[si, :, 10 * k - 1, :2], init_pos[si] + gpu[si, :, 10 * k - 1, :2]
            err = np.abs(p_ref - p_gpu).max()
            d_aa = np.linalg.norm(p_ref[:, None] - p_ref[None], axis=-1)
            d_am = np.linalg.norm(p_ref[:, None] - mpos[None], axis=-1)
            margin = min(np.abs(d_aa - 100.0).min(), np.abs(d_am - 50.0).min())
            if margin <= 2.0 * err + 1e-6:
                first[sc] = k
                report.append(f'scene {sc}: tick {k} may flip an edge (pair {margin:.2e} m from its radius, runs {err:.2e} m apart)')
                break
    return first, report


@pytest.mark.parametrize('name', list(BENCH_CASES))
def test_benchmarked_shape_matches_reference_golden(name):
    """The path bench.py measures, at its own shape: >= 1024 policy rows per launch, so the tcgen05 node kernel
    (attn_post_sw_kernel), the tensor-core PointNet / K'|V' kernels and policy_head2_kernel run -- an 80-step closed loop of
    8 scenes x 128 agents x 512 polylines against the reference's own B = 8 run (fp32 and fp64) under the SURVEY section 8d
    protocol, on the scenes the golden keeps.  Ticks after a possible edge flip (see _flip_candidates) are reported, not gated."""
    from prosim_b200 import lib
    kw, goal, scenes = BENCH_CASES[name]
    gold = load_golden(name)
    n_sw, n_post = lib.launch_count(lib.KERNEL_CLASSES['attn_post_sw']), lib.launch_count(lib.KERNEL_CLASSES['attn_post'])
    out, _ = _run_gpu(kw, goal, name=name)
    n_sw = lib.launch_count(lib.KERNEL_CLASSES['attn_post_sw']) - n_sw
    n_post = lib.launch_count(lib.KERNEL_CLASSES['attn_post']) - n_post
    assert n_sw >= 12 * 8 + 12 and n_post == 12 * 8 + 12 + 12          # ticks + generator (+ encoder) on the 32-row kernel
    names, traj, vel = stack_rollout(out)
    rows, P = gold['rows'], int(gold['n_rows'])
    assert len(names) == P and [names[i] for i in rows] == gold['agent_names'].tolist()
    n_ticks = len(out['pair_names']) // P
    pair_rows = np.concatenate([rows + k * P for k in range(n_ticks)])
    assert [out['pair_names'][i] for i in pair_rows] == gold['pair_names'].tolist()
    mp = out['motion_pred'].cpu().numpy()[pair_rows]
    R, A, S = len(rows), kw['n_agents'], kw['steps']
    assert np.abs(mp[:R] - gold['motion_pred'][:R]).max() < 1e-5           # first tick is open loop
    assert np.abs(out['reconst_pred'].cpu().numpy()[pair_rows] - gold['reconst_pred']).max() < 1e-5
    traj, vel = traj[rows], vel[rows]
    first, report = _flip_candidates(synthetic.make_batch(**kw), gold, traj, scenes, A, S)
    print('\n'.join(report) if report else 'no edge-flip candidates')
    gap_ref = per_tick_max(gold['traj'], gold['traj64'])
    gated = 0
    for si, sc in enumerate(scenes):
        sl = slice(si * A, (si + 1) * A)
        k_ok = first[sc]                    # ticks 0 .. k_ok-1 were decoded from provably identical graphs ... and tick k_ok's
        if k_ok == 0:                       # OUTPUT is the first that may differ
            continue
        g_gpu = per_tick_max(traj[sl], gold['traj64'][sl])[:k_ok]
        d32 = per_tick_max(traj[sl], gold['traj'][sl])[:k_ok]
        gv = per_tick_max(vel[sl], gold['vel64'][sl])[:k_ok]
        print(name, 'scene', sc, 'gated ticks', k_ok, 'gpu-vs-64', g_gpu, 'gpu-vs-ref32', d32)
        assert np.all(g_gpu <= gap_ref[:k_ok] + 1e-4), (sc, g_gpu, gap_ref)
        assert np.all(d32[gap_ref[:k_ok] < 5e-5] <= 1e-4), (sc, d32, gap_ref)
        assert np.all(gv <= per_tick_max(gold['vel'], gold['vel64'])[:k_ok] + 1e-4)
        gated += k_ok
    assert gated >= len(scenes) * n_ticks // 2, 'too few (scene, tick) pairs without a possible edge flip: the test is vacuous'


@pytest.mark.parametrize('name', list(BENCH_CASES))
def test_benchmarked_shape_teacher_forced_ticks(name):
    """SURVEY section 8d (1) at the bench shape: the reference's own state (its fp32 B = 8 run) is fed to every GPU tick of the
    kept scenes; each tick's motion_pred must match the reference's within 1e-5.  Same launches as the bench workload
    (>= 1024 rows: tcgen05 node kernel, policy_head2_kernel)."""
    kw, goal, scenes = BENCH_CASES[name]
    gold = load_golden(name)
    model = _model(goal)
    batch = make_case_batch(name).to('cuda')
    rows, P = torch.as_tensor(gold['rows']).cuda(), int(gold['n_rows'])
    R = len(rows)
    with torch.no_grad():
        scene = model.encode_scene(batch)
        policy = model.generate_policy(batch, scene, model.encode_prompt(batch))
        ids = {'motion_pred': batch.extras['prompt']['motion_pred']['agent_ids']}
        g_traj, g_vel = torch.as_tensor(gold['traj']).cuda(), torch.as_tensor(gold['vel']).cuda()
        all_t = list(range(0, kw['steps'], 10))
        worst = 0.0
        for k in range(kw['steps'] // 10):
            trajs = model.init_agent_trajs(ids, batch)
            st = trajs['motion_pred']
            T = st['traj'].shape[2]
            st['traj'].view(-1, T, 4)[rows, 11:11 + 10 * k] = g_traj[:, :10 * k]
            st['vel'].view(-1, T, 2)[rows, 11:11 + 10 * k] = g_vel[:, :10 * k]
            st['last_step'] = 11 + 10 * k
            scene_t, a_pos = model.step_env(scene, trajs, batch, ids, 10 * k, all_t)
            out = model.decode_output(policy, scene_t, ids, batch, a_pos, 10 * k, None)['motion_pred']
            err = float((out['motion_pred'][rows].cpu() - torch.as_tensor(gold['motion_pred'][k * R:(k + 1) * R])).abs().max())
            print(name, 'tick', 10 * k, 'teacher-forced max err', err)
            worst = max(worst, err)
    assert worst < 1e-5


@pytest.mark.parametrize('name', ['cfg2_a64_m256_s40', 'ragged_b3_s30', 'ragged_goal_b2_s20', 'ragged_mixed_b2_s20'])
def test_teacher_forced_ticks_and_edge_sets(name):
    """Feed the oracle's state at every tick to the GPU tick: motion_pred within 1e-5, neighbour sets identical,
    fut_obs written in place like the reference does."""
    kw, goal = CASES[name]
    sd = weights.random_state_dict(0, goal)
    orc = ProSimOracle(sd, goal)
    orc.trace = []
    b_cpu = synthetic.make_batch(**kw)
    ref = orc.forward(b_cpu)['motion_pred']
    model = _model(goal)
    batch = synthetic.make_batch(**kw).to('cuda')
    scene = model.encode_scene(batch)
    prompt = model.encode_prompt(batch)
    policy = model.generate_policy(batch, scene, prompt)
    ids = {'motion_pred': batch.extras['prompt']['motion_pred']['agent_ids']}
    # one-time phases
    emd_ref = orc.generate_policy(b_cpu, orc.encode_scene(synthetic.make_batch(**kw)), orc.encode_prompt(b_cpu))['emd']
    assert (policy['motion_pred']['emd'].cpu() - emd_ref).abs().max() < 5e-5
    pl = scene['_plan']
    assert edge_set(pl.edges_enc[0].to_edge_index()) == edge_set(orc._dbg_enc['e_a'])
    assert edge_set(pl.edges_enc[1].to_edge_index()) == edge_set(orc._dbg_enc['e_s'])
    assert edge_set(pl.edges_gen[0].to_edge_index()) == edge_set(orc._dbg_gen['e_pp'])
    assert edge_set(pl.edges_gen[1].to_edge_index()) == edge_set(orc._dbg_gen['e_sp'])
    P = len(ref['pair_names']) // len(orc.trace)
    model.keep_tick_edges = True
    all_t = [x['t'] for x in orc.trace]
    for k, (tick, state) in enumerate(zip(orc.trace, orc.trace_states)):
        trajs = model.init_agent_trajs(ids, batch)
        st = trajs['motion_pred']
        ls = state['last_step']
        st['traj'][:, :, :ls] = state['traj'].cuda()
        st['vel'][:, :, :ls] = state['vel'].cuda()
        st['last_step'] = ls
        pl.tick_edges = []
        with torch.no_grad():      # one tick driven from outside through the reference's own methods (traj_sam.py:159-170)
            scene_t, a_pos = model.step_env(scene, trajs, batch, ids, tick['t'], all_t)
            out = model.decode_output(policy, scene_t, ids, batch, a_pos, tick['t'], None)['motion_pred']
        torch.cuda.synchronize()
        err = (out['motion_pred'].cpu() - tick['motion_pred']).abs().max()
        print(name, 'tick', tick['t'], 'teacher-forced max err', float(err))
        assert err < 1e-5
        e_a, e_m = pl.tick_edges[0]
        assert edge_set(e_a) == edge_set(tick['e_ap']), f'a2p edge set differs at t={tick["t"]}'
        assert edge_set(e_m) == edge_set(tick['e_mp']), f'm2p edge set differs at t={tick["t"]}'
        if tick['t'] > 0:
            f_gpu, f_ref = batch.extras['fut_obs'][tick['t']], b_cpu.extras['fut_obs'][tick['t']]
            assert torch.equal(f_gpu['mask'].cpu(), f_ref['mask'])
            a = torch.nan_to_num(f_gpu['input'].cpu())
            b = torch.nan_to_num(f_ref['input'])
            assert (a - b).abs().max() < 2e-5
            assert (f_gpu['position'].cpu() - f_ref['position']).abs().max() < 1e-5
    model.keep_tick_edges = False


def test_tick_loop_driven_from_outside_equals_rollout_batch():
    """The reference's per-tick method surface (traj_sam.py:144-175, 178, 205, 276, 635; scene_encoder.update_scene_emb
    attn_fusion.py:238; policy(...) policy/base.py:19): a caller that runs the loop itself gets the bits rollout_batch gives,
    and the dictionaries it sees have the reference's keys and shapes."""
    kw = dict(agents_per_scene=[14, 9, 21], map_per_scene=[40, 32, 25], steps=30, permute_obs=True)
    ref, _ = _run_gpu(kw, False)
    model = _model(False)
    batch = synthetic.make_batch(**kw).to('cuda')
    with torch.no_grad():
        scene = model.encode_scene(batch)
        policy = model.generate_policy(batch, scene, model.encode_prompt(batch))
        ids = {'motion_pred': batch.extras['prompt']['motion_pred']['agent_ids']}
        trajs = model.init_agent_trajs(ids, batch)
        all_t = sorted(batch.extras['all_t_indices'].tolist())
        S = 40 + 32 + 25 + 14 + 9 + 21
        assert scene['scene_tokens'].shape == (S, 128) and scene['scene_pos'].shape == (S, 2) and scene['scene_ori'].shape == (S, 1)
        assert scene['scene_type'].sum() == 44 and scene['scene_batch_idx'].shape == (S,)
        outs = []
        for t in all_t:
            scene, a_pos = model.step_env(scene, trajs, batch, ids, t, all_t)
            st = trajs['motion_pred']
            tidx = st['last_step']
            # traj_sam.py:211-215: the quirky world pose (no rotation of the offset by init_heading)
            want = st['init_pos'] + st['traj'][..., tidx - 1, :2]
            assert a_pos['position'].shape == (3, 21, 2) and a_pos['heading'].shape == (3, 21, 1)
            valid = batch.extras['prompt']['motion_pred']['prompt_mask']
            assert torch.equal(a_pos['position'][valid], want[valid])
            assert scene['scene_tokens'].shape == (S, 128)
            b_emd, b_obs, b_map, b_pos, names = model._get_policy_batch_input(batch, ids['motion_pred'], policy['motion_pred'],
                                                                              scene, a_pos, t)
            out = {'motion_pred': model.get_action(b_emd, b_obs, b_map, b_pos, [names])}
            out['motion_pred']['pair_names'] = names
            assert out['motion_pred']['motion_pred'].shape == (44, 1, 10, 5) and names[0] == f'0-{ids["motion_pred"][0][0]}-{t}'
            trajs = model.step_agent_traj(trajs, out, ids, t, 'val')
            outs.append(out)
        res = model._process_rollout(trajs, outs, ids)['motion_pred']
    assert res['pair_names'] == ref['pair_names'] and torch.equal(res['motion_pred'], ref['motion_pred'])
    assert isinstance(res['rollout_trajs'], dict) and list(res['rollout_trajs'].keys()) == list(ref['rollout_trajs'].keys())
    for name, r in ref['rollout_trajs'].items():
        assert torch.equal(r['traj'], res['rollout_trajs'][name]['traj']) and torch.equal(r['vel'], res['rollout_trajs'][name]['vel'])
    # the policy also accepts the reference's dense token layout (traj_sam.py:356-400: [B, S, 128] + mask)
    with torch.no_grad():
        pl = scene['_plan']
        x_a, a_p, a_o = scene['_agent']
        B, A = 3, 21
        dense = lambda flat, w: torch.zeros(B * A, w, device='cuda').index_copy_(0, pl.i['agent_rows2'].long(), flat.reshape(-1, w)).view(B, A, w)
        mask = torch.zeros(B * A, dtype=torch.bool, device='cuda').index_fill_(0, pl.i['agent_rows2'].long(), True).view(B, A)
        d_obs = {'input': dense(x_a, 128), 'mask': mask, 'pos': dense(a_p, 2), 'ori': dense(a_o, 1)}
        again = model.policy(b_emd, d_obs, b_map, b_pos, names, None)
    assert torch.equal(again['motion_pred'], outs[-1]['motion_pred']['motion_pred'])


@pytest.mark.parametrize('fusion,attn', [('mlp', False), ('replace', True), ('mlp', True)])
def test_obs_update_variants_match_oracle(fusion, attn):
    """MODEL.OBS_UPDATE.{FUSION: 'mlp', ATTN_UPDATE: True} (scene_encoder/attn_fusion.py:136-203): the per-tick scene update
    with the old / new token MLP (obs_fuse_kernel) and with the re-run agent / map -> agent attention, against the oracle
    (itself bit-equal to the reference's own code for these variants: tests/test_oracle_vs_reference.py)."""
    from prosim_b200.config import get_config
    from prosim_b200.model import ProSimB200
    sd = weights.random_state_dict(0, obs_fusion=fusion)
    model = ProSimB200(get_config(opts=['MODEL.OBS_UPDATE.FUSION', fusion, 'MODEL.OBS_UPDATE.ATTN_UPDATE', attn]), sd, device='cuda')
    kw = dict(agents_per_scene=[20, 13, 31], map_per_scene=[48, 37, 60], steps=40, permute_obs=True)
    orc = ProSimOracle(sd)
    orc.attn_update = attn
    ref = orc.forward(synthetic.make_batch(**kw))['motion_pred']
    with torch.no_grad():
        out = model.forward(synthetic.make_batch(**kw).to('cuda'), 'val')['motion_pred']
    assert out['pair_names'] == ref['pair_names']
    P = 64
    assert (out['motion_pred'][:P].cpu() - ref['motion_pred'][:P]).abs().max() < 1e-5          # tick 0: no update yet
    assert (out['motion_pred'][P:2 * P].cpu() - ref['motion_pred'][P:2 * P]).abs().max() < 5e-5  # tick 1: first updated scene
    _, traj, _ = stack_rollout(out)
    _, traj_ref, _ = stack_rollout(ref)
    print(fusion, attn, 'per-tick max |traj - oracle|', per_tick_max(traj, traj_ref))
    assert np.abs(traj - traj_ref).max() < 1e-4
    if attn:
        pl = model._last_plan
        assert edge_set(pl.edges_upd[0].to_edge_index()) == edge_set(orc._dbg_upd['e_a'])
        assert edge_set(pl.edges_upd[1].to_edge_index()) == edge_set(orc._dbg_upd['e_ma'])
    plain = ProSimOracle(weights.random_state_dict(0)).forward(synthetic.make_batch(**kw))['motion_pred']
    assert np.abs(traj - stack_rollout(plain)[1]).max() > 1e-3                                  # the variant is really active


def test_fused_tick_entry_point_is_bit_identical():
    """prosim_policy_tick (the whole tick in one C call) against the same tick issued kernel family by
    kernel family through the per-step entry points: identical bits, for FFMA-sized and tensor-core-sized launches."""
    model = _model(False)
    for kw in (dict(agents_per_scene=[14, 9, 21], map_per_scene=[40, 32, 25], steps=30, permute_obs=True),
               dict(n_scenes=9, n_agents=120, n_map=70, steps=20)):
        try:
            model.fused_tick = True
            a, _ = _run_gpu(kw, False)
            model.fused_tick = False
            b, _ = _run_gpu(kw, False)
        finally:
            model.fused_tick = True
        assert torch.equal(a['motion_pred'], b['motion_pred']) and torch.equal(a['_state']['traj'], b['_state']['traj'])


def test_batch_invariance_and_replicas():
    """A scene rolled out alone equals the same scene inside a batch bit for bit (fixed-order reductions,
    no cross-scene coupling) -- the property the multi-GPU scene sharding relies on."""
    kw = dict(agents_per_scene=[20, 31, 12], map_per_scene=[50, 64, 40], steps=30)
    out_b, _ = _run_gpu(kw, False)
    for s, (na, nm) in enumerate(zip(kw['agents_per_scene'], kw['map_per_scene'])):
        out_1, _ = _run_gpu(dict(n_scenes=1, n_agents=na, n_map=nm, steps=30, first_scene=s), False)
        for name, r in out_1['rollout_trajs'].items():
            other = out_b['rollout_trajs'][f'{s}-{name.split("-", 1)[1]}']
            assert torch.equal(r['traj'], other['traj']) and torch.equal(r['vel'], other['vel']), (s, name)


def test_batch_invariance_across_kernel_variants():
    """FFMA node kernels: large batches run the 32/64-row gemm_tile kernels, single scenes the 2..16-row latency kernels;
    every output is accumulated in the same (ascending k) order, so the two must still agree bit for bit."""
    from prosim_b200 import lib
    lib.set_tensor_core(False)
    try:
        kw = dict(n_scenes=12, n_agents=100, n_map=80, steps=20)
        out_b, _ = _run_gpu(kw, False)
        for s in (0, 7):
            out_1, _ = _run_gpu(dict(n_scenes=1, n_agents=100, n_map=80, steps=20, first_scene=s), False)
            for name, r in out_1['rollout_trajs'].items():
                other = out_b['rollout_trajs'][f'{s}-{name.split("-", 1)[1]}']
                assert torch.equal(r['traj'], other['traj']) and torch.equal(r['vel'], other['vel']), (s, name)
    finally:
        lib.set_tensor_core(True)


def test_batch_invariance_tensor_core_kernel():
    """tcgen05 node kernel: a scene's result does not depend on what else is in the batch (bit exact between a 12-scene and a
    20-scene batch, and -- since the 16-row and 32-row CTAs, the one-launch and the three-kernel edge paths share their
    arithmetic -- with the scene run alone; test_single_scene_equals_large_batch_bit_for_bit asserts the latter exactly)."""
    kw = dict(n_agents=100, n_map=80, steps=20)
    out_a, _ = _run_gpu(dict(kw, n_scenes=12), False)
    out_b, _ = _run_gpu(dict(kw, n_scenes=20), False)
    for name, r in out_a['rollout_trajs'].items():
        other = out_b['rollout_trajs'][name]
        assert torch.equal(r['traj'], other['traj']) and torch.equal(r['vel'], other['vel']), name
    worst = 0.0
    for s in (0, 7):
        out_1, _ = _run_gpu(dict(kw, n_scenes=1, first_scene=s), False)
        for name, r in out_1['rollout_trajs'].items():
            other = out_a['rollout_trajs'][f'{s}-{name.split("-", 1)[1]}']
            worst = max(worst, float((r['traj'] - other['traj']).abs().max()))
    print('tensor-core batch vs FFMA single scene, 20 steps: max |traj diff|', worst)
    assert worst < 1e-4


def test_tensor_core_path_with_capped_generator_graph():
    """>= 1024 policy rows with 450 map polylines per scene: the generator's scene->prompt graph hits its 512-neighbour
    cap (16 z tiles per row in the edge kernel) on the tensor-core path; checked against the same scenes run one by one
    (one-launch edge kernel with four warps per row, itself pinned to the reference goldens)."""
    kw = dict(n_agents=100, n_map=450, steps=20)
    out_a, _ = _run_gpu(dict(kw, n_scenes=11), False)
    worst = 0.0
    for s in (3, 10):
        out_1, _ = _run_gpu(dict(kw, n_scenes=1, first_scene=s), False)
        for name, r in out_1['rollout_trajs'].items():
            other = out_a['rollout_trajs'][f'{s}-{name.split("-", 1)[1]}']
            worst = max(worst, float((r['traj'] - other['traj']).abs().max()))
    assert worst < 1e-4, worst


def test_row_split_schedule_is_bit_identical():
    """prosim_set_stack_split: the fixed-source stacks run as independent >= 1024-row chains on side streams; every
    setting must give the same bits (rows never interact; chains are cut at multiples of 128 rows)."""
    from prosim_b200 import lib
    kw = dict(n_scenes=24, n_agents=100, n_map=80, steps=20)
    try:
        lib.set_stack_split(1)
        ref, _ = _run_gpu(kw, False)
        for parts in (2, 3, 4):
            lib.set_stack_split(parts)
            out, _ = _run_gpu(kw, False)
            assert torch.equal(out['motion_pred'], ref['motion_pred'])
            for name, r in ref['rollout_trajs'].items():
                assert torch.equal(r['traj'], out['rollout_trajs'][name]['traj']), (parts, name)
    finally:
        lib.set_stack_split(1)     # the library default


def test_plan_cache_hits_only_on_identical_bookkeeping():
    """The host index plan is reused when shapes, validity masks and id lists repeat, and only then."""
    model = _model(False)
    kw = dict(agents_per_scene=[9, 14], map_per_scene=[20, 31], steps=20)
    b1 = synthetic.make_batch(**kw).to('cuda')
    b2 = synthetic.make_batch(**kw).to('cuda')                     # same bookkeeping, different tensors
    b3 = synthetic.make_batch(**kw, permute_obs=True).to('cuda')   # same shapes, other id order
    b4 = synthetic.make_batch(agents_per_scene=[9, 13], map_per_scene=[20, 31], steps=20).to('cuda')
    with torch.no_grad():
        p1, p2, p3, p4 = (model._plan(b) for b in (b1, b2, b3, b4))
        assert p1 is p2 and p3 is not p1 and p4 is not p1 and p4 is not p3
        o1 = model.forward(b1, 'val')['motion_pred']
        o3 = model.forward(b3, 'val')['motion_pred']
    for name, r in o1['rollout_trajs'].items():
        assert (r['traj'] - o3['rollout_trajs'][name]['traj']).abs().max() < 1e-4


def test_run_to_run_determinism_on_the_tensor_core_path():
    """Dynamic row queue, TMA rings and elected-lane MMA issue must not leak scheduling into the results: repeated
    forwards of a >= 1024-row batch are bit identical."""
    model = _model(False)
    b0 = synthetic.make_batch(n_scenes=10, n_agents=110, n_map=90, steps=20).to('cuda')
    ref = None
    with torch.no_grad():
        for _ in range(6):
            b = synthetic.clone_batch(b0)[0]
            traj = model.forward(b, 'val')['motion_pred']['_state']['traj'].clone()
            if ref is None:
                ref = traj
            assert torch.equal(ref, traj)


def test_action_noise_rollout_matches_oracle_with_injected_draws():
    """MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD > 0 (act_decoder.py:113-115): with the SAME standard-normal draws fed
    to the GPU head kernel and to the oracle, the noisy closed loop agrees like the noise-free one does; with its own
    CUDA generator the model is reproducible under torch.manual_seed and actually noisy."""
    from prosim_b200.config import get_config
    from prosim_b200.model import ProSimB200
    kw = dict(agents_per_scene=[14, 9], map_per_scene=[40, 32], steps=20)
    sd = weights.random_state_dict(0)
    model = ProSimB200(get_config(opts=['MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD', 0.05]), sd, device='cuda')
    assert model.noise_std == pytest.approx(0.05)
    g = torch.Generator().manual_seed(77)
    draws = [torch.randn(23, 1, 10, 2, generator=g) for _ in range(2)]
    it_gpu, it_cpu = iter(draws), iter(draws)
    model.noise_fn = lambda shape: next(it_gpu).cuda().contiguous()
    orc = ProSimOracle(sd)
    orc.noise_std = 0.05
    orc.noise_fn = lambda like: next(it_cpu).to(like.dtype)
    ref = orc.forward(synthetic.make_batch(**kw))['motion_pred']
    with torch.no_grad():
        out = model.forward(synthetic.make_batch(**kw).to('cuda'), 'val')['motion_pred']
    assert (out['motion_pred'][:23].cpu() - ref['motion_pred'][:23]).abs().max() < 1e-5       # open-loop tick
    _, traj, _ = stack_rollout(out)
    _, traj_ref, _ = stack_rollout(ref)
    assert np.abs(traj - traj_ref).max() < 1e-4
    quiet = ProSimOracle(sd).forward(synthetic.make_batch(**kw))['motion_pred']
    assert np.abs(traj - stack_rollout(quiet)[1]).max() > 0.05                                  # the noise is really there
    # own generator: seeded runs repeat, differently seeded runs differ
    model.noise_fn = lambda shape: torch.randn(shape, device='cuda', dtype=torch.float32)
    runs = []
    for seed in (5, 5, 6):
        torch.manual_seed(seed)
        with torch.no_grad():
            runs.append(model.forward(synthetic.make_batch(**kw).to('cuda'), 'val')['motion_pred']['motion_pred'].clone())
    assert torch.equal(runs[0], runs[1]) and not torch.equal(runs[0], runs[2])


def test_agent_permutation_equivariance():
    """Storing the observation slots in another order must not change any agent's trajectory beyond
    summation-order rounding (edges are visited in ascending slot index)."""
    kw = dict(n_scenes=1, n_agents=24, n_map=60, steps=20)
    out_a, _ = _run_gpu(kw, False)
    out_b, _ = _run_gpu(dict(kw, permute_obs=True), False)
    for name, r in out_a['rollout_trajs'].items():
        assert (r['traj'] - out_b['rollout_trajs'][name]['traj']).abs().max() < 1e-4


def test_cpu_batch_is_rejected():
    from prosim_b200 import lib
    with pytest.raises(lib.ProSimLibError):
        _model(False).forward(synthetic.make_batch(n_scenes=1, n_agents=4, n_map=8, steps=10), 'val')


def test_parallel_rollout_batch_and_world_transform():
    """The reference's M-replica rollout helpers (rollout/gpu_utils.py:59-281): every replica equals the single rollout
    bit for bit, and the local->world transform matches the oracle restatement."""
    from oracle.prosim_oracle import rollout_trajs_in_world
    from prosim_b200.rollout import obtain_rollout_trajs_in_world, parallel_rollout_batch
    kw = dict(n_scenes=1, n_agents=20, n_map=48, steps=30)
    single, _ = _run_gpu(kw, False)
    model = _model(False)
    batch = synthetic.make_batch(**kw).to('cuda')
    M = 3
    res = parallel_rollout_batch(batch, M, model)
    assert model.mode == 'rollout'
    rt = res['motion_pred']['rollout_trajs']
    assert len(rt) == M * 20
    for name, r in single['rollout_trajs'].items():
        aid = name.split('-', 1)[1]
        for m in range(M):
            assert torch.equal(r['traj'], rt[f'{m}-{aid}']['traj']) and torch.equal(r['vel'], rt[f'{m}-{aid}']['vel'])
    th = 0.7
    tf = torch.tensor([[math.cos(th), -math.sin(th), 120.5], [math.sin(th), math.cos(th), -40.25], [0.0, 0.0, 1.0]])
    batch.centered_world_from_agent_tf = tf[None].repeat(M, 1, 1)
    trajs_M, ids_M = obtain_rollout_trajs_in_world(batch, res)
    assert len(trajs_M) == M and ids_M[0] == [n.split('-')[1] for n in list(rt.keys())[:20]]
    cpu_res = {'rollout_trajs': {n: {k: v.cpu() for k, v in r.items()} for n, r in rt.items()}}
    ref, _ = rollout_trajs_in_world(cpu_res, tf)
    got = np.concatenate(trajs_M, axis=0)
    assert np.abs(got[..., :2] - ref[..., :2].numpy()).max() < 2e-5
    dh = np.abs(got[..., 2] - ref[..., 2].numpy())
    assert np.minimum(dh, 2 * math.pi - dh).max() < 1e-5


def test_goal_sampler_parallel_rollout_matches_oracle():
    """parallel_rollout_batch with a sampler model (rollout/gpu_utils.py:203-216): M goal conditions are sampled from the
    sampler's top_K goal predictions, the policy tokens are generated on the M-replica batch and the replicas rolled out
    together.  Checked against the oracle run on a hand-replicated CPU batch carrying the SAME sampled goals; also
    obtain_rollout_trajs_in_world(noise_std > 0) and K goal-conditioned tokens per agent through rollout_batch."""
    import copy
    from prosim_b200.containers import BatchCondition, BatchDataDict, BatchPrompt
    from prosim_b200.rollout import _rep, _rep_imd, obtain_rollout_trajs_in_world, parallel_rollout_batch
    model = _model(True)
    kw = dict(n_scenes=1, n_agents=12, n_map=30, steps=20, goal=True)
    M, K = 3, 4
    ids = synthetic.make_batch(**kw).extras['prompt']['motion_pred']['agent_ids'][0]
    g = torch.Generator().manual_seed(9)

    class Sampler:
        def forward(self, batch, mode):
            return {'motion_pred': {'pair_names': [f'0-{a}-0' for a in ids], 'goal_point': torch.randn(12, K, 2, generator=g) * 30,
                                    'goal_prob': torch.rand(12, K, generator=g)}}

    batch = synthetic.make_batch(**kw).to('cuda')
    torch.manual_seed(21)
    res = parallel_rollout_batch(batch, M, model, top_K=2, sampler_model=Sampler())['motion_pred']
    goal = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in batch.extras['condition'].all_cond['goal'].items()}
    assert goal['input'].shape == (M, 12, 3) and not torch.equal(goal['input'][0], goal['input'][1])
    # the oracle on a CPU batch replicated by hand, with the same goals
    cpu = synthetic.make_batch(**kw)
    ex = cpu.extras
    ex['init_obs'], ex['init_map'] = _rep_imd(ex['init_obs'], M), _rep_imd(ex['init_map'], M)
    ex['fut_obs'] = BatchDataDict({t: _rep_imd(ex['fut_obs'][t], M) for t in ex['fut_obs'].keys()})
    ex['prompt'] = BatchPrompt({'motion_pred': {k: (v * M if isinstance(v, list) else _rep(v, M))
                                                for k, v in ex['prompt']['motion_pred'].items()}})
    ex['condition'] = BatchCondition({'goal': goal})
    cpu.scene_ids = list(cpu.scene_ids) * M
    ref = ProSimOracle(weights.random_state_dict(0, True), True).forward(cpu)['motion_pred']
    assert res['pair_names'] == ref['pair_names']
    _, traj, _ = stack_rollout(res)
    _, traj_ref, _ = stack_rollout(ref)
    assert np.abs(traj - traj_ref).max() < 1e-4
    assert np.abs(traj[:12] - traj[12:24]).max() > 1e-2                  # replicas with different goals really differ
    # world transform with evaluation noise (gpu_utils.py:254-256): same generator state -> same draws
    batch.centered_world_from_agent_tf = torch.eye(3)[None].repeat(M, 1, 1)
    quiet, _ = obtain_rollout_trajs_in_world(batch, {'motion_pred': res})
    torch.manual_seed(4)
    noisy, _ = obtain_rollout_trajs_in_world(batch, {'motion_pred': res}, noise_std=0.5)
    d = np.concatenate(noisy)[..., :2] - np.concatenate(quiet)[..., :2]
    assert 0.3 < d.std() < 0.7 and np.abs(np.concatenate(noisy)[..., 2] - np.concatenate(quiet)[..., 2]).max() < 1e-6
    # K policy tokens per agent: rollout_batch selects one per agent (traj_sam.py:159, 402-439)
    b1 = synthetic.make_batch(**kw).to('cuda')
    with torch.no_grad():
        scene = model.encode_scene(b1)
        policy = model.generate_policy(b1, scene, model.encode_prompt(b1))['motion_pred']
        idsd = {'motion_pred': b1.extras['prompt']['motion_pred']['agent_ids']}
        emd = policy['emd']
        multi = {'emd': torch.stack([emd + 0.1 * k for k in range(K)], dim=2), 'agent_type': policy['agent_type'],
                 'goal_prob': torch.rand(1, 12, K, device='cuda'), 'goal_point': torch.randn(1, 12, K, 2, device='cuda')}
        top1 = multi['goal_prob'].argmax(-1)
        out = model.rollout_batch(b1, scene, {'motion_pred': multi}, idsd, model.init_agent_trajs(idsd, b1), [0, 10], 'val')['motion_pred']
        b2 = synthetic.make_batch(**kw).to('cuda')
        scene2 = model.encode_scene(b2)
        chosen = {'emd': torch.gather(multi['emd'], 2, top1[..., None, None].repeat(1, 1, 1, 128)).squeeze(2), 'agent_type': policy['agent_type']}
        want = model.rollout_batch(b2, scene2, {'motion_pred': chosen}, idsd, model.init_agent_trajs(idsd, b2), [0, 10], 'val')['motion_pred']
    assert out['goal_prob'].shape == (24, K) and out['goal'].shape == (24, 2)    # carried through like the reference's policy does
    assert torch.equal(out['motion_pred'], want['motion_pred'])          # ROLLOUT.POLICY.TOP_K = 1: the most probable token


def test_graphed_forward_matches_eager_and_accepts_host_batches():
    """CUDA-graph replay (graph_runner.GraphedForward): bit-identical to the eager forward, reusable for a new batch of
    the same shape, fed straight from pinned host memory."""
    from prosim_b200.graph_runner import GraphedForward
    model = _model(False)
    runner = GraphedForward(model)
    kw = dict(n_scenes=2, n_agents=24, n_map=40, steps=30)
    for first in (0, 5, 0):
        eager, _ = _run_gpu(dict(kw, first_scene=first), False)
        host = synthetic.make_batch(**kw, first_scene=first, pin_memory=True)
        out = runner(host, 'val')['motion_pred']
        torch.cuda.synchronize()
        assert out['pair_names'] == eager['pair_names']
        assert torch.equal(out['motion_pred'], eager['motion_pred'])
        for name, r in eager['rollout_trajs'].items():
            assert torch.equal(r['traj'], out['rollout_trajs'][name]['traj']), (first, name)
    assert len(runner._cache) == 1


def test_single_scene_equals_large_batch_bit_for_bit():
    """The latency path and the throughput path are the same arithmetic: a scene rolled out alone (16-row node-kernel CTAs,
    one-launch edge kernel with four warps per row, small-launch head) equals the same scene inside a 30-scene batch (3000 rows:
    32-row node-kernel CTAs, the three-kernel persistent edge path, 32-row head kernel) bit for bit."""
    kw = dict(n_agents=100, n_map=80, steps=30)
    out_b, _ = _run_gpu(dict(kw, n_scenes=30), False)
    for s in (0, 17, 29):
        out_1, _ = _run_gpu(dict(kw, n_scenes=1, first_scene=s), False)
        for name, r in out_1['rollout_trajs'].items():
            other = out_b['rollout_trajs'][f'{s}-{name.split("-", 1)[1]}']
            assert torch.equal(r['traj'], other['traj']) and torch.equal(r['vel'], other['vel']), (s, name)
