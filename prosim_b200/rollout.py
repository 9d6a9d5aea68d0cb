"""Host-side mirror of the reference's batch-parallel rollout helpers (prosim/rollout/gpu_utils.py).

  replica_batch_for_parallel_rollout   :59-123   one encoded scene replicated M times along the batch
  parallel_rollout_batch               :179-228  encode once, replicate, roll out M replicas together
  obtain_rollout_trajs_in_world        :230-281  agent-t0-frame trajectories -> world (x, y, heading)

Same names, argument order and return values.  The goal-sampler branch (``sampler_model``) needs the goal-prediction
heads, which the released config disables (DECODER.GOAL_PRED.ENABLE = False); it raises NotImplementedError.
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
    scene_embs_M = dict(scene_embs)
    scene_embs_M['scene_tokens'] = torch.cat([tok[:NM].repeat(M, 1), tok[NM:].repeat(M, 1)])
    scene_embs_M['scene_pos'] = torch.cat([pos[:NM].repeat(M, 1), pos[NM:].repeat(M, 1)])
    scene_embs_M['scene_ori'] = torch.cat([ori[:NM].repeat(M), ori[NM:].repeat(M)]).view(-1, 1)
    for name in ('obs_mask', 'map_mask'):
        scene_embs_M[name] = scene_embs[name].repeat(M, 1)
    scene_embs_M.pop('_plan')
    policy_emds_M = None
    if policy_emds is not None:
        policy_emds_M = {'motion_pred': {}}
        for name, val in policy_emds['motion_pred'].items():
            policy_emds_M['motion_pred'][name] = _rep(val, M)
    policy_agent_ids_M = {'motion_pred': [policy_agent_ids['motion_pred'][0]] * M}
    st = agent_trajs['motion_pred']
    agent_trajs_M = {'motion_pred': {k: (_rep(v, M) if isinstance(v, torch.Tensor) else v) for k, v in st.items()}}
    prompt_encs_M = None
    return scene_embs_M, policy_emds_M, prompt_encs_M, policy_agent_ids_M, agent_trajs_M, batch


def parallel_rollout_batch(batch, M, model, top_K=3, sampler_model=None, smooth_dist=5.0):
    """gpu_utils.py:179-228: encode the (single) scene once, replicate it M times, roll the replicas out together."""
    if sampler_model is not None:
        raise NotImplementedError('goal-sampler rollouts need DECODER.GOAL_PRED, which the released config disables')
    with torch.no_grad():
        scene_embs = model.encode_scene(batch)
        prompt_encs = model.encode_prompt(batch)
        policy_agent_ids = {task: batch.extras['prompt'][task]['agent_ids'] for task in ['motion_pred']}
        all_t_indices = sorted(batch.extras['all_t_indices'].cpu().numpy().tolist())
        agent_trajs = model.init_agent_trajs(policy_agent_ids, batch, all_t_indices)
        policy_emds = model.decode_policy(batch, scene_embs, prompt_encs)
        scene_embs_M, policy_emds_M, _, policy_agent_ids_M, agent_trajs_M, batch = replica_batch_for_parallel_rollout(
            scene_embs, policy_emds, prompt_encs, policy_agent_ids, agent_trajs, batch, M)
        scene_embs_M['_plan'] = model._plan(batch)
        model.mode = 'rollout'
        result_M = model.rollout_batch(batch, scene_embs_M, policy_emds_M, policy_agent_ids_M, agent_trajs_M,
                                       all_t_indices, 'rollout')
    return result_M


def obtain_rollout_trajs_in_world(batch, result_M, noise_std=0.0):
    """gpu_utils.py:230-281.  Returns (list over batch ids of numpy [n_agents, steps, 3] (x, y, heading) in world
    coordinates, list of object-id lists).  Like the reference, every agent uses ``batch.centered_world_from_agent_tf[0]``."""
    if noise_std > 0.0:
        raise NotImplementedError('noise_std > 0 is an evaluation-time perturbation outside the rollout path')
    res = result_M['motion_pred']
    st = res['_state']
    names = list(res['rollout_trajs'].keys())
    B, N, T = st['traj'].shape[:3]
    dev = st['traj'].device
    tf = torch.as_tensor(batch.centered_world_from_agent_tf[0], dtype=torch.float32).to(dev).contiguous()
    rows = torch.tensor(res['rollout_trajs']._rows, dtype=torch.int32, device=dev)
    world = ops.rollout_to_world(st['traj'].view(-1, T, 4), st['init_pos'].view(-1, 2), st['init_heading'].view(-1),
                                 rows, T, HIST, T - HIST, tf).cpu().numpy()
    batch_ids = [int(n.split('-')[0]) for n in names]
    object_ids = [n.split('-')[1] for n in names]
    trajs_M, ids_M = [], []
    for b in sorted(set(batch_ids)):
        sel = [i for i, x in enumerate(batch_ids) if x == b]
        trajs_M.append(world[sel])
        ids_M.append([object_ids[i] for i in sel])
    return trajs_M, ids_M
