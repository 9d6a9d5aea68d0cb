"""Host-side mirror of the reference's batch-parallel rollout helpers (prosim/rollout/gpu_utils.py).

  replica_batch_for_parallel_rollout   :59-123   one encoded scene replicated M times along the batch
  parallel_rollout_batch               :179-228  encode once, replicate, roll out M replicas together
  obtain_rollout_trajs_in_world        :230-281  agent-t0-frame trajectories -> world (x, y, heading)

  sample_M_goal_cond_to_batch          :125-177  M sampled goal conditions from a sampler model's goal predictions

Same names, argument order and return values.  ``sampler_model`` is any object whose ``forward(batch, 'val')`` returns
``{'motion_pred': {'pair_names', 'goal_point' [P, K, 2], 'goal_prob' [P, K]}}`` (the reference uses a second ProSim with goal
prediction heads; the released config ships none, so the branch is exercised with a stand-in sampler in the tests).
"""
import torch

from . import ops
from .containers import BatchCondition, BatchDataDict, BatchPrompt, InputMaskData

HIST = 11


def _rep(t, M):
    return t.repeat(M, *([1] * (t.ndim - 1)))


def _rep_imd(d, M):
    out = InputMaskData.__new__(InputMaskData)
    out.input, out.mask = _rep(d.input, M), _rep(d.mask, M)
    out.position = _rep(d.position, M) if d.position is not None else None
    out.heading = _rep(d.heading, M) if d.heading is not None else None
    out.agent_ids = (d.agent_ids * M) if d.agent_ids is not None else None
    return out


def replica_batch_for_parallel_rollout(scene_embs, policy_emds, prompt_encs, policy_agent_ids, agent_trajs, batch, M):
    """gpu_utils.py:59-123.  The batch's inputs are replicated in place (the reference replaces ``fut_obs``; here the
    observation / map / prompt tensors are replicated too so that the integer bookkeeping of the M-scene batch can be
    rebuilt from it)."""
    from .model import _SceneEmbs
    pl = scene_embs['_plan']
    NM = pl.NM
    tok, pos, ori = scene_embs['scene_tokens'], scene_embs['scene_pos'], scene_embs['scene_ori'].reshape(-1)
    ex = batch.extras
    ex['init_obs'] = _rep_imd(ex['init_obs'], M)
    ex['init_map'] = _rep_imd(ex['init_map'], M)
    ex['fut_obs'] = BatchDataDict({t: _rep_imd(ex['fut_obs'][t], M) for t in ex['fut_obs'].keys()})
    ex['prompt'] = BatchPrompt({task: {k: (v * M if isinstance(v, list) else _rep(v, M)) for k, v in p.items()
                                       if not k.startswith('_') and k != 'prompt_emd'}
                                for task, p in ex['prompt'].all_prompts.items()})
    ex['condition'] = BatchCondition({c: {k: (v if isinstance(v, (list, str)) else _rep(v, M)) for k, v in d.items()}
                                      for c, d in ex['condition'].all_cond.items()})
    batch.scene_ids = list(batch.scene_ids) * M
    batch._b200_plan = None
    # token layout of the model: all map tokens of all scenes first, then all agent tokens
    tok_M = torch.cat([tok[:NM].repeat(M, 1), tok[NM:].repeat(M, 1)])
    pos_M = torch.cat([pos[:NM].repeat(M, 1), pos[NM:].repeat(M, 1)])
    ori_M = torch.cat([ori[:NM].repeat(M), ori[NM:].repeat(M)])
    scene_embs_M = _SceneEmbs({'obs_mask': scene_embs['obs_mask'].repeat(M, 1), 'map_mask': scene_embs['map_mask'].repeat(M, 1),
                               'max_map_num': scene_embs['max_map_num'], 'max_agent_num': scene_embs['max_agent_num'],
                               '_tok': tok_M, '_tok_pos': pos_M, '_tok_ori': ori_M,
                               '_agent': (tok_M[M * NM:], pos_M[M * NM:], ori_M[M * NM:]), '_slot': 0, '_shared': {}})
    policy_emds_M = None
    if policy_emds is not None:
        policy_emds_M = {'motion_pred': {}}
        for name, val in policy_emds['motion_pred'].items():
            if isinstance(val, torch.Tensor) and not (name.startswith('_') and name != '_emd_flat'):
                policy_emds_M['motion_pred'][name] = _rep(val, M)
    policy_agent_ids_M = {'motion_pred': [policy_agent_ids['motion_pred'][0]] * M}
    st = agent_trajs['motion_pred']
    agent_trajs_M = {'motion_pred': {k: (_rep(v, M) if isinstance(v, torch.Tensor) else v) for k, v in st.items()}}
    # gpu_utils.py:99-106 (the replicated prompt encodings feed decode_policy on the M-scene batch in the sampler branch)
    pe = prompt_encs['motion_pred']
    prompt_encs_M = {'motion_pred': {k: _rep(pe[k], M) for k in ('prompt', 'prompt_mask', 'position', 'heading', 'agent_type',
                                                                  'prompt_emd') if k in pe}}
    if '_emd_flat' in pe:
        prompt_encs_M['motion_pred']['_emd_flat'] = _rep(pe['_emd_flat'], M)
    prompt_encs_M['motion_pred']['agent_ids'] = [pe['agent_ids'][0]] * M
    return scene_embs_M, policy_emds_M, prompt_encs_M, policy_agent_ids_M, agent_trajs_M, batch


def sample_M_goal_cond_to_batch(batch, sample_result, top_K, M, stop_smooth_num=5.0):
    """gpu_utils.py:125-177: for each of the M replicas and each prompt agent the sampler predicted goals for, pick one of
    its top_K goal points at random (``torch.randperm`` on the host generator, like the reference), snap goals closer than
    ``stop_smooth_num`` to the origin on both axes to (0, 0), and install them as THE goal condition of the batch
    ({'input' [M, n, 3] = (gx, gy, 80), 'mask', 'prompt_idx' [M, n, 1], 'prompt_mask'}).  One scene (batch size 1) is assumed."""
    device = batch.extras['prompt']['motion_pred']['prompt'].device
    res = sample_result['motion_pred']
    index = {name: i for i, name in enumerate(res['pair_names'])}      # the reference's list.index search, built once
    goal_point, goal_prob = res['goal_point'].detach().cpu(), res['goal_prob'].detach().cpu()
    goal_inputs_M, prompt_idxs_M = [], []
    for b in range(M):
        goal_inputs_b, prompt_idxs_b = [], []
        for pidx, aname in enumerate(batch.extras['prompt']['motion_pred']['agent_ids'][0]):
            pred_idx = index.get(f'0-{aname}-0')
            if pred_idx is None:
                continue
            top_k_idx = torch.argsort(-goal_prob[pred_idx])[:top_K]
            select_goal = goal_point[pred_idx][top_k_idx[torch.randperm(top_K)[0]]].clone()
            if torch.abs(select_goal[0]) < stop_smooth_num and torch.abs(select_goal[1]) < stop_smooth_num:
                select_goal[0] = 0.0
                select_goal[1] = 0.0
            goal_inputs_b.append(torch.tensor([select_goal[0], select_goal[1], 80.0]))
            prompt_idxs_b.append(torch.tensor([pidx]))
        goal_inputs_M.append(torch.stack(goal_inputs_b))
        prompt_idxs_M.append(torch.stack(prompt_idxs_b))
    goal_inputs_M = torch.stack(goal_inputs_M).to(device)
    prompt_idxs_M = torch.stack(prompt_idxs_M).to(device)
    n = goal_inputs_M.shape[1]
    goal_cond_M = {'input': goal_inputs_M, 'mask': torch.ones(M, n, dtype=torch.bool, device=device),
                   'prompt_idx': prompt_idxs_M, 'prompt_mask': torch.ones(M, n, dtype=torch.bool, device=device),
                   'caption_str': 'show as green cross'}
    batch.extras['condition'].all_cond = {'goal': goal_cond_M}
    return batch


def parallel_rollout_batch(batch, M, model, top_K=3, sampler_model=None, smooth_dist=5.0):
    """gpu_utils.py:179-228: encode the (single) scene once, replicate it M times, roll the replicas out together."""
    with torch.no_grad():
        scene_embs = model.encode_scene(batch)
        prompt_encs = model.encode_prompt(batch)
        policy_agent_ids = {task: batch.extras['prompt'][task]['agent_ids'] for task in ['motion_pred']}
        all_t_indices = sorted(batch.extras['all_t_indices'].cpu().numpy().tolist())
        agent_trajs = model.init_agent_trajs(policy_agent_ids, batch, all_t_indices)
        if sampler_model is None:
            policy_emds = model.decode_policy(batch, scene_embs, prompt_encs)
            scene_embs_M, policy_emds_M, _, policy_agent_ids_M, agent_trajs_M, batch = replica_batch_for_parallel_rollout(
                scene_embs, policy_emds, prompt_encs, policy_agent_ids, agent_trajs, batch, M)
            scene_embs_M['_plan'] = model._plan(batch)
        else:
            # gpu_utils.py:203-216: M goal conditions sampled from the sampler's top_K goal predictions, then the policy
            # tokens are generated on the M-replica batch (the goal condition differs per replica)
            sample_result = sampler_model.forward(batch, 'val')
            scene_embs_M, _, prompt_encs_M, policy_agent_ids_M, agent_trajs_M, batch = replica_batch_for_parallel_rollout(
                scene_embs, None, prompt_encs, policy_agent_ids, agent_trajs, batch, M)
            batch = sample_M_goal_cond_to_batch(batch, sample_result, top_K, M, stop_smooth_num=smooth_dist)
            scene_embs_M['_plan'] = model._plan(batch)
            policy_emds_M = model.decode_policy(batch, scene_embs_M, prompt_encs_M)
        model.mode = 'rollout'
        result_M = model.rollout_batch(batch, scene_embs_M, policy_emds_M, policy_agent_ids_M, agent_trajs_M,
                                       all_t_indices, 'rollout')
    return result_M


def obtain_rollout_trajs_in_world(batch, result_M, noise_std=0.0):
    """gpu_utils.py:230-281.  Returns (list over batch ids of numpy [n_agents, steps, 3] (x, y, heading) in world
    coordinates, list of object-id lists).  Like the reference, every agent uses ``batch.centered_world_from_agent_tf[0]``."""
    res = result_M['motion_pred']
    st = res['_state']
    names = list(res['rollout_trajs'].keys())
    B, N, T = st['traj'].shape[:3]
    dev = st['traj'].device
    tf = torch.as_tensor(batch.centered_world_from_agent_tf[0], dtype=torch.float32).to(dev).contiguous()
    rows = torch.tensor(res['rollout_trajs']._rows, dtype=torch.int32, device=dev)
    traj = st['traj'].view(-1, T, 4)
    if noise_std > 0.0:     # gpu_utils.py:254-256: Gaussian noise on the rolled-out (x, y) before the frame transforms
        traj = traj.clone()
        traj[rows.long(), HIST:, :2] += torch.randn(len(names), T - HIST, 2, device=dev) * noise_std
    world = ops.rollout_to_world(traj, st['init_pos'].view(-1, 2), st['init_heading'].view(-1),
                                 rows, T, HIST, T - HIST, tf).cpu().numpy()
    batch_ids = [int(n.split('-')[0]) for n in names]
    object_ids = [n.split('-')[1] for n in names]
    trajs_M, ids_M = [], []
    for b in sorted(set(batch_ids)):
        sel = [i for i, x in enumerate(batch_ids) if x == b]
        trajs_M.append(world[sel])
        ids_M.append([object_ids[i] for i in sel])
    return trajs_M, ids_M
