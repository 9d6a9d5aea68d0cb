"""ProSimB200 -- the reference's closed-loop rollout model behind its own interface.

Host-side mirror of ``ProSim`` (prosim/models/traj_sam.py:13-643): same public methods, argument
meaning, output dict and ``state_dict`` key names, so it drops in wherever the reference model is
called (``model.forward(batch, 'val')``, ``parallel_rollout_batch``).  All arithmetic runs in the
hand-written sm_100a kernels of libprosim_b200.so; this file only does what the reference does in
Python too -- bookkeeping -- but once per batch and with integer index maps instead of the
reference's per-tick f-string / ``list.index`` searches (traj_sam.py:245-249, 289-298, 478).

Device data layout (all fp32 / int32, allocated once per batch shape):
  tokens      [NM + NA][128]  valid map tokens first, then valid agent tokens (attn_fusion.py:90-105)
  state       traj [B*N][11+steps][4], vel [..][2], init_pos [B*N][2], init_heading [B*N]
  edges       fixed-stride neighbour lists + z = LayerNorm(rel PE) [rows*stride][128]
  K'|V'       [layers][rows][256]; the map side of the policy is computed once per scene and reused
              by every tick (the reference recomputes it 8 times)
"""
from collections.abc import Mapping

import numpy as np
import torch
import torch.nn as nn

from . import lib, ops, weights
from .config import check_supported, get_config
from .registry import registry

HIST = 11
STEP = 10
D = 128


class _Plan:
    """Integer bookkeeping of one batch, built once on the host and uploaded in a single copy."""
    pass


class _RolloutTrajs(Mapping):
    """``rollout_trajs`` of the reference (traj_sam.py:582-593): {"b-id": {traj [steps,4], init_pos [2], init_heading [1],
    vel [steps,2]}} as views of the state buffers.  The per-agent views are created on first access (all of them with four
    ``unbind`` calls when iterated) instead of 8 Python indexing ops per agent per forward -- 35 ms at 4096 agents."""

    def __init__(self, names, rows, st):
        self._names, self._rows, self._st = names, [int(r) for r in rows], st
        self._index = None
        self._all = None

    def _materialise(self):
        if self._all is None:
            st = self._st
            T = st['traj'].shape[2]
            traj = st['traj'].view(-1, T, 4)[:, HIST:].unbind(0)
            vel = st['vel'].view(-1, T, 2)[:, HIST:].unbind(0)
            pos = st['init_pos'].view(-1, 2).unbind(0)
            head = st['init_heading'].view(-1, 1).unbind(0)
            self._all = {n: {'traj': traj[r], 'init_pos': pos[r], 'init_heading': head[r], 'vel': vel[r]}
                         for n, r in zip(self._names, self._rows)}
        return self._all

    def __getitem__(self, name):
        if self._all is not None:
            return self._all[name]
        if self._index is None:
            self._index = {n: r for n, r in zip(self._names, self._rows)}
        r = self._index[name]
        st = self._st
        T = st['traj'].shape[2]
        return {'traj': st['traj'].view(-1, T, 4)[r, HIST:], 'init_pos': st['init_pos'].view(-1, 2)[r],
                'init_heading': st['init_heading'].view(-1, 1)[r], 'vel': st['vel'].view(-1, T, 2)[r, HIST:]}

    def __iter__(self):
        return iter(self._names)

    def __len__(self):
        return len(self._names)

    def items(self):
        return self._materialise().items()

    def values(self):
        return self._materialise().values()


def _pad4(a):
    a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
    pad = (-a.size) % 4
    return np.concatenate([a, np.zeros(pad, np.int32)]) if pad else a


@registry.register_model(name='prosim_b200')
class ProSimB200(nn.Module):
    def __init__(self, config=None, state_dict=None, device=None):
        super().__init__()
        self.config = config if config is not None else get_config()
        check_supported(self.config)
        cfg = self.config
        self.tasks = list(cfg.TASK.TYPES)
        self.use_condition = len(cfg.PROMPT.CONDITION.TYPES) > 0
        self.cond_types = tuple(cfg.PROMPT.CONDITION.TYPES)
        self.rollout_steps = cfg.ROLLOUT.POLICY.REPLAN_FREQ
        self.hist_step = cfg.DATASET.FORMAT.HISTORY.STEPS
        assert self.rollout_steps == STEP and self.hist_step == HIST and cfg.DATASET.FORMAT.TARGET.STEPS == STEP
        self.num_layers = cfg.MODEL.POLICY.ACT_DECODER.ATTN.NUM_LAYER
        self.cond_layers = cfg.MODEL.CONDITION_TRANSFORMER.NLAYER
        self.mode = 'val'
        # act_decoder.py:113-115: Gaussian noise on the predicted per-step displacements; the draws come from torch's CUDA
        # generator with the reference's shapes ([P, 1, 10, 2] per tick), noise_fn can be replaced to inject given draws
        self.noise_std = float(cfg.MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD)
        self.noise_fn = lambda shape: torch.randn(shape, device=self._device, dtype=torch.float32)
        self._device = torch.device(device if device is not None else 'cuda')
        self._sd = None
        self._arena = None
        self._off = None
        self._bufs = {}
        self._plan_cache = []
        used = [t for t in cfg.PROMPT.CONDITION.MOTION_TAG.USED_TAGS if t in weights.V_ACTION_TAG_ID]   # condition_encoders.py:58
        if 'v_action_tag' in self.cond_types and tuple(used) != weights.V_ACTION_TAGS:
            raise NotImplementedError('PROMPT.CONDITION.MOTION_TAG.USED_TAGS differs from the released list')
        lut = -torch.ones(16, dtype=torch.int64)        # tag id (motion_tag_utils.py:4-15) -> position in USED_TAGS
        for pos, tag in enumerate(used):
            lut[weights.V_ACTION_TAG_ID[tag]] = pos
        self._tag_slot = lut
        if state_dict is None:
            state_dict = weights.random_state_dict(0, self.cond_types)
        self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ weights
    @property
    def device(self):
        return self._device

    def state_dict(self, *args, **kwargs):
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        want = [n for n, _, _ in weights.param_specs(self.cond_types, self.num_layers, self.cond_layers)]
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in set(want)]
        if missing or (strict and unexpected):
            raise RuntimeError(f'load_state_dict: missing {missing[:5]} unexpected {unexpected[:5] if strict else []}')
        self._sd = {k: state_dict[k].detach().float().cpu().clone() for k in want}
        arena, self._off = weights.pack_model(self._sd, self.num_layers, self.cond_layers)
        lib.load()  # fail loudly here, not at the first kernel call, if the native library is absent
        self._arena = arena.to(self._device)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device) != self._device:
            self._device = torch.device(device)
            self._arena = self._arena.to(self._device)
            self._bufs = {}
        return self

    def eval(self):
        return self

    def _note_edges(self, kind, edges, n_layers):
        """Measurement support (bench.py roofline): remember the edge count of every attention launch of a forward."""
        if getattr(self, 'edge_log', None) is not None:
            self.edge_log.append((kind, edges.deg.sum(), edges.n_dst, n_layers))

    def _buf(self, name, shape, dtype=torch.float32):
        shape = tuple(int(s) for s in shape)
        b = self._bufs.get(name)
        n = int(np.prod(shape)) if shape else 1
        if b is None or b.dtype != dtype or b.numel() < n:
            b = torch.empty(max(n, 1), device=self._device, dtype=dtype)
            self._bufs[name] = b
        return b[:n].view(shape)

    # ------------------------------------------------------------------ bookkeeping
    def _plan(self, batch):
        if getattr(batch, '_b200_plan', None) is not None:
            return batch._b200_plan
        ex = batch.extras
        obs, mp = ex['init_obs'], ex['init_map']
        prm = ex['prompt'][self.tasks[0]]
        pl = _Plan()
        pl.all_t = sorted(int(t) for t in ex['all_t_indices'].cpu().numpy().tolist())
        pl.B, pl.A = obs['input'].shape[:2]
        pl.M = mp['input'].shape[1]
        pl.N = prm['prompt_mask'].shape[1]
        B, A, M, N = pl.B, pl.A, pl.M, pl.N
        futs = [ex['fut_obs'][t] for t in pl.all_t[1:]]
        # one device->host copy of every validity mask the bookkeeping needs
        ov = torch.stack([obs['mask'].all(-1).any(-1)] + [f['mask'].all(-1).any(-1) for f in futs])
        flat = torch.cat([ov.reshape(-1), mp['mask'].any(-1).reshape(-1), prm['prompt_mask'].reshape(-1)]).cpu().numpy()
        # Plans are pure functions of (shapes, validity masks, id lists, ticks): batches that repeat them -- replicas of a
        # scene, a fixed cast of agents re-simulated from new states -- reuse the device-resident index maps instead of
        # rebuilding and re-uploading them (1.4 ms of host work during which the GPU idles at the start of a forward)
        id_lists = [obs['agent_ids'], prm['agent_ids']] + [f['agent_ids'] for f in futs]
        sig = (B, A, M, N, tuple(pl.all_t), str(self._device))
        mask_bytes = flat.tobytes()
        for ent in self._plan_cache:
            if ent[0] == sig and ent[1] == mask_bytes and len(ent[2]) == len(id_lists) and \
                    all(a is b or a == b for a, b in zip(ent[2], id_lists)):
                batch._b200_plan = ent[3]
                return ent[3]
        nt = len(pl.all_t)
        ov = flat[:nt * B * A].reshape(nt, B, A).copy()
        mv = flat[nt * B * A:nt * B * A + B * M].reshape(B, M)
        pm = flat[nt * B * A + B * M:].reshape(B, N)

        ids = prm['agent_ids']
        pl.policy_ids = ids
        n_b = np.array([len(x) for x in ids], dtype=np.int64)
        for b in range(B):
            if not (pm[b, :n_b[b]].all() and not pm[b, n_b[b]:].any()):
                raise ValueError('prompt_mask must be True exactly on the first len(agent_ids[b]) slots')
        p_b = np.repeat(np.arange(B), n_b)
        p_n = np.concatenate([np.arange(k) for k in n_b]) if B else np.zeros(0, np.int64)
        pl.P = int(n_b.sum())
        p_off = np.concatenate([[0], np.cumsum(n_b)])
        pl.p_b, pl.p_n = p_b, p_n

        map_rows = np.flatnonzero(mv.reshape(-1))
        m_cnt = mv.sum(axis=1)
        m_off = np.concatenate([[0], np.cumsum(m_cnt)])
        pl.NM = int(map_rows.size)

        def slots(id_lists):
            out = np.empty(pl.P, dtype=np.int64)
            k = 0
            for b in range(B):
                lut = {a: i for i, a in enumerate(id_lists[b])}
                for a in ids[b]:
                    out[k] = b * A + lut[a]
                    k += 1
            return out

        ints = {}
        ints['map_rows'] = map_rows
        ints['p_row'] = p_b * N + p_n
        ints['p_scene'] = p_b
        ints['seg_prompt'] = np.stack([p_off[:-1], n_b, np.zeros(B), np.zeros(B)], axis=1)
        ints['seg_map'] = np.stack([m_off[:-1], m_cnt, np.zeros(B), np.zeros(B)], axis=1)
        lut = -np.ones(B * N + 1, dtype=np.int64)
        lut[p_b * N + p_n] = np.arange(pl.P)
        ints['prow_lut'] = lut
        slot0 = slots(obs['agent_ids'])
        pl.NA, pl.max_a = [], 0
        for i in range(nt):
            if i == 0:
                slot = slot0
            else:
                f_ids = futs[i - 1]['agent_ids']
                slot = slot0 if f_ids == obs['agent_ids'] else slots(f_ids)
                ov[i].reshape(-1)[slot] = True
            valid = ov[i]
            a_cnt = valid.sum(axis=1)
            a_off = np.concatenate([[0], np.cumsum(a_cnt)])
            ints[f'p_slot{i}'] = slot
            ints[f'agent_rows{i}'] = np.flatnonzero(valid.reshape(-1))
            ints[f'seg_agent{i}'] = np.stack([a_off[:-1], a_cnt, np.zeros(B), np.zeros(B)], axis=1)
            pl.NA.append(int(a_cnt.sum()))
            pl.max_a = max(pl.max_a, int(a_cnt.max()) if B else 0)
            if i == 0:
                a_cnt0, a_off0 = a_cnt, a_off
        pl.max_m = int(m_cnt.max()) if B else 0
        pl.max_p = int(n_b.max()) if B else 0
        pl.max_tok = int((m_cnt + a_cnt0).max()) if B else 0
        ints['tok_scene'] = np.concatenate([np.repeat(np.arange(B), m_cnt), np.repeat(np.arange(B), a_cnt0)])
        ints['seg_scene'] = np.stack([m_off[:-1], m_cnt, pl.NM + a_off0[:-1], a_cnt0], axis=1)
        pl.scene_batch_idx_host = ints['tok_scene']

        offs, chunks, pos = {}, [], 0
        for k, v in ints.items():
            v = _pad4(v)
            offs[k] = (pos, int(np.asarray(ints[k]).size))
            chunks.append(v)
            pos += v.size
        host = torch.from_numpy(np.concatenate(chunks)).pin_memory()
        dev = host.to(self._device, non_blocking=True)
        pl.int_host, pl.int_dev = host, dev
        pl.i = {k: dev[o:o + n] for k, (o, n) in offs.items()}
        pl.key = (B, A, M, N, pl.P, pl.NM, tuple(pl.NA), pl.max_a, pl.max_m, pl.max_p, pl.max_tok, tuple(pl.all_t),
                  int(host.numel()))
        pl.steps = len(pl.all_t) * STEP
        pl.T = HIST + pl.steps
        batch._b200_plan = pl
        self._plan_cache.insert(0, (sig, mask_bytes, id_lists, pl))
        del self._plan_cache[4:]
        return pl

    # ------------------------------------------------------------------ reference API
    def forward(self, batch, mode):
        """traj_sam.py:59-71."""
        if not batch.extras['init_obs']['input'].is_cuda:
            raise lib.ProSimLibError('ProSimB200 needs the batch on the GPU (batch.to(device)); there is no CPU path')
        self.mode = mode
        scene_embs = self.encode_scene(batch)
        prompt_encs = self.encode_prompt(batch)
        return self.decode_batch(scene_embs, prompt_encs, batch, mode)

    def decode_batch(self, scene_embs, prompt_encs, batch, mode):
        """traj_sam.py:103-116."""
        policy_emds = self.generate_policy(batch, scene_embs, prompt_encs)
        policy_agent_ids = {task: batch.extras['prompt'][task]['agent_ids'] for task in self.tasks}
        all_t_indices = self._plan(batch).all_t
        agent_trajs = self.init_agent_trajs(policy_agent_ids, batch)
        return self.rollout_batch(batch, scene_embs, policy_emds, policy_agent_ids, agent_trajs, all_t_indices, mode)

    def encode_scene(self, batch):
        """traj_sam.py:73-77 -> scene_encoder/base.py:31-46 + attn_fusion.py:78-134."""
        pl = self._plan(batch)
        ex = batch.extras
        obs, mp = ex['init_obs'], ex['init_map']
        ar, off = self._arena, self._off
        NM, NA = pl.NM, pl.NA[0]
        S = NM + NA
        tok = self._buf('tok', (S, D))
        tok_pos = self._buf('tok_pos', (S, 2))
        tok_ori = self._buf('tok_ori', (S,))
        ops.pointnet(1, mp['input'], mp['mask'], pl.i['map_rows'], ar, off['map_enc'], out=tok[:NM], tc_off=off['map_enc_tc'])
        ops.pointnet(0, obs['input'], obs['mask'], pl.i['agent_rows0'], ar, off['obs_enc'], out=tok[NM:], tc_off=off['obs_enc_tc'])
        ops.gather_pose(mp['position'], mp['heading'], pl.i['map_rows'], tok_pos[:NM], tok_ori[:NM])
        ops.gather_pose(obs['position'], obs['heading'], pl.i['agent_rows0'], tok_pos[NM:], tok_ori[NM:])
        dim_t = ar[off['dim_t16']:off['dim_t16'] + 16]
        k_a = min(self.config.MODEL.SCENE_ENCODER.ATTN.MAX_NUM_NEIGH * 4, 100)
        k_s = self.config.MODEL.SCENE_ENCODER.ATTN.MAX_NUM_NEIGH
        a_pos, a_ori = tok_pos[NM:], tok_ori[NM:]
        e_a = ops.knn_edges(a_pos, pl.i['tok_scene'][NM:], a_pos, pl.i['seg_agent0'].view(-1, 4), k_a, max(pl.max_a, 1),
                            min(k_a, max(pl.max_a, 1)))
        e_s = ops.knn_edges(tok_pos, pl.i['tok_scene'], tok_pos, pl.i['seg_scene'].view(-1, 4), k_s, max(pl.max_tok, 1),
                            min(k_s, max(pl.max_tok, 1)))
        ops.edge_pe(e_a, a_pos, a_ori, a_pos, a_ori, dim_t, z=self._buf('z_enc_a', (NA * e_a.stride, 96)))
        ops.edge_pe(e_s, tok_pos, tok_ori, tok_pos, tok_ori, dim_t, z=self._buf('z_enc_s', (S * e_s.stride, 96)))
        self._note_edges('enc_a2a', e_a, self.num_layers)
        self._note_edges('enc_s2s', e_s, self.num_layers)
        ws = self._buf('attn_ws', (lib.load().prosim_attn_workspace_floats(S, S, max(e_a.stride, e_s.stride)),))
        lf = weights.ATTN_LAYER_FLOATS
        xa = tok[NM:]
        for i in range(self.num_layers):
            ops.attn_layer(xa, xa, e_a, ar, off['enc_a2a'] + i * lf, out=xa, workspace=ws)
            ops.attn_layer(tok, tok, e_s, ar, off['enc_s2s'] + i * lf, out=tok, workspace=ws)
        # built from the already-uploaded index arrays: no pageable H2D copy (it would stall the host behind the encoder)
        sb = pl.i['tok_scene'].long()
        st = torch.cat([torch.zeros(NM, dtype=torch.long, device=self._device),
                        torch.ones(NA, dtype=torch.long, device=self._device)])
        pl.edges_enc = (e_a, e_s)
        return {'obs_mask': obs['mask'].all(-1).any(-1), 'map_mask': mp['mask'].any(-1), 'scene_batch_idx': sb,
                'scene_type': st, 'scene_pos': tok_pos, 'scene_ori': tok_ori.view(-1, 1), 'scene_tokens': tok,
                'max_map_num': pl.M, 'max_agent_num': pl.A, '_plan': pl}

    def encode_prompt(self, batch, prompt_dict={}):
        """traj_sam.py:79-101 -> prompt_encoder/base.py:37-50 (MLP on the valid prompt rows)."""
        pl = self._plan(batch)
        out = {}
        for task in (self.tasks if len(prompt_dict) == 0 else prompt_dict.keys()):
            data = prompt_dict[task] if task in prompt_dict else batch.extras['prompt'][task]
            rows = pl.i['p_row'].long()
            feat = data['prompt'].reshape(pl.B * pl.N, -1)[rows].contiguous()
            emd_flat = ops.mlp2(feat, feat.shape[1], True, self._arena, self._off['prompt_mlp'])
            emd = torch.zeros(pl.B * pl.N, D, device=self._device)
            emd[rows] = emd_flat
            data['prompt_emd'] = emd.view(pl.B, pl.N, D)
            data['_emd_flat'] = emd_flat
            out[task] = data
        return out

    def generate_policy(self, batch, scene_embs, prompt_encs):
        """traj_sam.py:118-142 -> decoder/sym_coord.py:63-140 (+ goal condition: condition_transformer/*)."""
        pl = self._plan(batch)
        ar, off = self._arena, self._off
        dcfg = self.config.MODEL.DECODER.ATTN
        lf = weights.ATTN_LAYER_FLOATS
        dim_t = ar[off['dim_t16']:off['dim_t16'] + 16]
        result = {}
        for task, enc in prompt_encs.items():
            rows = pl.i['p_row'].long()
            P = pl.P
            x_p = enc['_emd_flat'] if '_emd_flat' in enc else enc['prompt_emd'].reshape(-1, D)[rows].contiguous()
            p_pos = enc['position'].reshape(-1, 2)[rows].contiguous()
            p_ori = enc['heading'].reshape(-1)[rows].contiguous()
            tok, tok_pos = scene_embs['scene_tokens'], scene_embs['scene_pos']
            tok_ori = scene_embs['scene_ori'].reshape(-1)
            S = tok.shape[0]
            cap = dcfg.MAX_NUM_NEIGH
            e_pp = ops.radius_edges(p_pos, pl.i['p_scene'], p_pos, pl.i['seg_prompt'].view(-1, 4), dcfg.PROMPT_RADIUS, cap,
                                    min(cap + 1, max(pl.max_p, 1)), drop_self=True)
            e_sp = ops.radius_edges(p_pos, pl.i['p_scene'], tok_pos, pl.i['seg_scene'].view(-1, 4), dcfg.SCENE_RADIUS, cap,
                                    min(cap, max(pl.max_tok, 1)))
            ops.edge_pe(e_pp, p_pos, p_ori, p_pos, p_ori, dim_t, z=self._buf('z_pp', (P * e_pp.stride, 96)))
            ops.edge_pe(e_sp, p_pos, p_ori, tok_pos, tok_ori, dim_t, z=self._buf('z_sp', (P * e_sp.stride, 96)))
            kv_s = ops.attn_kv(tok, ar, off['dec_s2p'], self.num_layers, lf, kv=self._buf('kv_dec', (self.num_layers, S, 2 * D)))
            ws = self._buf('attn_ws', (lib.load().prosim_attn_workspace_floats(P, P, max(e_pp.stride, e_sp.stride)),))
            emd_flat = ops.attn_stack(x_p, self.num_layers, ops.stack_side(ar, off['dec_p2p'], e_pp),
                                      ops.stack_side(ar, off['dec_s2p'], e_sp, kv_s), workspace=ws)
            pl.edges_gen = (e_pp, e_sp)
            self._note_edges('gen_p2p', e_pp, self.num_layers)
            self._note_edges('gen_s2p', e_sp, self.num_layers)
            if self.use_condition and 'policy_decoder' in self.config.MODEL.CONDITION_TRANSFORMER.CONDITION_LOCATIONS:
                emd_flat = self._conditions(batch.extras['condition'], emd_flat, p_pos, p_ori, pl, ws)
            emd = torch.zeros(pl.B * pl.N, D, device=self._device)
            emd[rows] = emd_flat
            result[task] = {'emd': emd.view(pl.B, pl.N, D), 'agent_type': enc['agent_type'], '_emd_flat': emd_flat}
        return result

    decode_policy = generate_policy  # name used by the reference's rollout helpers (rollout/gpu_utils.py:200,216)

    def _conditions(self, cond, emd_flat, p_pos, p_ori, pl, ws):
        """condition_transformer/base.py:38-60.  Every condition entry is encoded by the encoder of its type (goal MLP +
        time PE, tag vector + interval PE, drag-point PointNet: condition_encoders.py), the embeddings of the types present
        on an agent are mean-pooled onto its self edge (condition_attns.py:114-189), 3 GNN layers run over those edges
        and the result is added to the policy embedding of EVERY valid prompt row (condition_attns.py:226).
        Two entries of the same type (same tag name) on one agent: the reference keeps whichever index_put wrote last;
        here as well one of them wins (which one is unspecified)."""
        ar, off = self._arena, self._off
        dev = self._device
        P, B, N = pl.P, pl.B, pl.N
        dim_t = ar[off['dim_t16']:off['dim_t16'] + 16]
        n_tags = len(weights.V_ACTION_TAGS)
        embs, puts, n_emb, n_slots = [], [], 0, 0
        for t in self.cond_types:
            width = n_tags if t == 'v_action_tag' else 1
            if t in cond.keys() and cond[t]['input'].shape[1] > 0:
                c = cond[t]
                C = c['input'].shape[1]
                if t == 'goal':
                    x = c['input'].reshape(-1, 3).float().contiguous()
                    emb = ops.mlp2(x, 2, False, ar, off['goal_mlp'], tpe_col=2,
                                   dim_t128=ar[off['dim_t128']:off['dim_t128'] + 128])
                    col = torch.zeros(B * C, dtype=torch.int64, device=dev)
                elif t == 'v_action_tag':
                    tags = c['input'].reshape(-1, 3).contiguous()
                    emb = ops.tag_embed(tags, ar[off['tag_vec']:off['tag_vec'] + 16 * D],
                                        ar[off['dim_t64']:off['dim_t64'] + 64], n_tags)
                    col = self._tag_slot.to(dev)[tags[:, 0].clamp(0, 15)]   # pooling order = USED_TAGS order
                else:
                    T = c['input'].shape[2]
                    if T not in (16, 8):
                        raise NotImplementedError(f'drag_point with {T} points per agent (16 or 8 are built)')
                    x = c['input'].reshape(B * C, T, 2).float().contiguous()
                    emb = ops.pointnet(2 if T == 16 else 3, x, None, torch.arange(B * C, device=dev, dtype=torch.int32),
                                       ar, off['drag_enc'])
                    col = torch.zeros(B * C, dtype=torch.int64, device=dev)
                bidx = torch.arange(B, device=dev)[:, None].expand(B, C)
                nidx = c['prompt_idx'][..., 0].clamp(min=0)
                row = pl.i['prow_lut'].long()[(bidx * N + nidx).reshape(-1)]
                valid = c['mask'].reshape(-1) & (row >= 0)
                if t == 'v_action_tag':
                    valid = valid & (tags[:, 0] >= 0) & (tags[:, 0] < 16) & (col >= 0)
                row = torch.where(valid, row, torch.full_like(row, P))
                puts.append((row, n_slots + col.clamp(min=0), n_emb + torch.arange(B * C, device=dev, dtype=torch.int32)))
                embs.append(emb)
                n_emb += B * C
            n_slots += width
        if not embs:
            return emd_flat
        slot = torch.full((P + 1, n_slots), -1, device=dev, dtype=torch.int32)
        for row, col, ent in puts:
            slot.index_put_((row, col), ent)
        extra, has = ops.cond_pool(torch.cat(embs) if len(embs) > 1 else embs[0], slot[:P].contiguous())
        nbr = torch.arange(P, device=dev, dtype=torch.int32)
        e = ops.EdgeList(nbr, has, 1, 1)
        ops.edge_pe(e, p_pos, p_ori, p_pos, p_ori, dim_t, extra=extra)
        x_c = ops.attn_stack(emd_flat, self.cond_layers, ops.stack_side(ar, off['cond_attn'], e), None, workspace=ws)
        return emd_flat + x_c

    def init_agent_trajs(self, policy_agent_ids, batch, all_t_indices=None):
        """traj_sam.py:597-633 (3-argument form accepted for rollout/gpu_utils.py:196)."""
        pl = self._plan(batch)
        obs = batch.extras['init_obs']
        R = pl.B * pl.N
        traj = torch.zeros(R, pl.T, 4, device=self._device)
        vel = torch.zeros(R, pl.T, 2, device=self._device)
        init_pos = torch.zeros(R, 2, device=self._device)
        init_heading = torch.zeros(R, device=self._device)
        ops.init_traj(obs['input'], obs['position'], obs['heading'], pl.i['p_slot0'], pl.i['p_row'], pl.T, traj, vel,
                      init_pos, init_heading)
        st = {'traj': traj.view(pl.B, pl.N, pl.T, 4), 'vel': vel.view(pl.B, pl.N, pl.T, 2),
              'init_pos': init_pos.view(pl.B, pl.N, 2), 'init_heading': init_heading.view(pl.B, pl.N, 1),
              'last_step': HIST}
        return {task: st for task in self.tasks}

    def rollout_batch(self, batch, scene_embs, policy_emds, policy_agent_ids, agent_trajs, all_t_indices, mode):
        """traj_sam.py:144-175 (tick loop) + 205-274 step_env + 178-202 decode_output + 276-349 step_agent_traj
        + 562-595 _process_rollout."""
        task = self.tasks[0]
        pl = self._plan(batch)
        ar, off = self._arena, self._off
        ex = batch.extras
        obs = ex['init_obs']
        lf = weights.ATTN_LAYER_FLOATS
        L = self.num_layers
        acfg = self.config.MODEL.POLICY.ACT_DECODER.ATTN
        P, NM, T = pl.P, pl.NM, pl.T
        st = agent_trajs[task]
        traj, vel = st['traj'].view(-1, T, 4), st['vel'].view(-1, T, 2)
        init_pos, init_heading = st['init_pos'].view(-1, 2), st['init_heading'].view(-1)
        tok = scene_embs['scene_tokens']
        tok_pos, tok_ori = scene_embs['scene_pos'], scene_embs['scene_ori'].reshape(-1)
        pe = policy_emds[task]
        rows = pl.i['p_row'].long()
        emd_flat = pe['_emd_flat'] if '_emd_flat' in pe else pe['emd'].reshape(-1, D)[rows].contiguous()
        a_type = pe['agent_type'].reshape(-1)[rows].to(torch.int32).contiguous()
        dim_t = ar[off['dim_t16']:off['dim_t16'] + 16]
        n_ticks = len(all_t_indices)
        max_na = max(pl.NA)

        x_m, m_pos, m_ori = tok[:NM], tok_pos[:NM], tok_ori[:NM]
        kv_m = ops.attn_kv(x_m, ar, off['pol_m2p'], L, lf, kv=self._buf('kv_m', (L, NM, 2 * D)))
        stride_a = min(acfg.MAX_NUM_NEIGH, max(pl.max_a, 1))
        stride_m = min(acfg.MAX_NUM_NEIGH, max(pl.max_m, 1))
        nbr_a, deg_a = self._buf('nbr_a', (P * stride_a,), torch.int32), self._buf('deg_a', (P,), torch.int32)
        nbr_m, deg_m = self._buf('nbr_m', (P * stride_m,), torch.int32), self._buf('deg_m', (P,), torch.int32)
        z_a, z_m = self._buf('z_a', (P * stride_a, 96)), self._buf('z_m', (P * stride_m, 96))
        self._buf('kv_a', (L, max_na, 2 * D))
        x_a_buf = self._buf('x_a', (max_na, D))
        a_pos_buf, a_ori_buf = self._buf('a_pos', (max_na, 2)), self._buf('a_ori', (max_na,))
        p_pos, p_ori = self._buf('p_pos', (P, 2)), self._buf('p_ori', (P,))
        fuse = self._buf('fuse', (P, D))
        ws = self._buf('attn_ws', (lib.load().prosim_attn_workspace_floats(P, 0, max(stride_a, stride_m)),))
        motion_pred = torch.empty(n_ticks, P, 1, STEP, 5, device=self._device)
        tidx = int(st['last_step'])
        pl.tick_edges = []

        for k, t in enumerate(all_t_indices):
            i = pl.all_t.index(int(t))  # plan slot of this tick (a caller may roll out a sub-range of ticks)
            na = pl.NA[i]
            if i == 0:
                ops.step_env(traj, vel, init_pos, init_heading, pl.i['p_row'], pl.i['p_slot0'], T, tidx, p_pos, p_ori)
                x_a, a_pos, a_ori = tok[NM:], tok_pos[NM:], tok_ori[NM:]
            else:
                fut = ex['fut_obs'][t]
                ops.step_env(traj, vel, init_pos, init_heading, pl.i['p_row'], pl.i[f'p_slot{i}'], T, tidx, p_pos, p_ori,
                             fut=(fut['input'], fut['mask'], fut['position'], fut['heading']))
                x_a, a_pos, a_ori = x_a_buf[:na], a_pos_buf[:na], a_ori_buf[:na]
                ops.pointnet(0, fut['input'], fut['mask'], pl.i[f'agent_rows{i}'], ar, off['obs_enc'], out=x_a, tc_off=off['obs_enc_tc'])
                ops.gather_pose(fut['position'], fut['heading'], pl.i[f'agent_rows{i}'], a_pos, a_ori)
            e_a = ops.radius_edges(p_pos, pl.i['p_scene'], a_pos, pl.i[f'seg_agent{i}'].view(-1, 4), acfg.AGENT_RADIUS,
                                   acfg.MAX_NUM_NEIGH, stride_a, nbr=nbr_a, deg=deg_a)
            e_m = ops.radius_edges(p_pos, pl.i['p_scene'], m_pos, pl.i['seg_map'].view(-1, 4), acfg.MAP_RADIUS,
                                   acfg.MAX_NUM_NEIGH, stride_m, nbr=nbr_m, deg=deg_m)
            e_m.warps_per_row = 2     # map radius 50 m: a policy row typically sees ~40 of the scene's polylines
            ops.edge_pe(e_a, p_pos, p_ori, a_pos, a_ori, dim_t, z=z_a)
            ops.edge_pe(e_m, p_pos, p_ori, m_pos, m_ori, dim_t, z=z_m)
            kva = ops.attn_kv(x_a, ar, off['pol_a2p'], L, lf, kv=self._buf('kv_a', (L, na, 2 * D)))
            ops.attn_stack(emd_flat, L, ops.stack_side(ar, off['pol_a2p'], e_a, kva),
                           ops.stack_side(ar, off['pol_m2p'], e_m, kv_m), out=fuse, workspace=ws)
            self._note_edges('pol_a2p', e_a, L)
            self._note_edges('pol_m2p', e_m, L)
            noise = self.noise_fn((P, 1, STEP, 2)) if self.noise_std > 0 else None
            ops.policy_head(fuse, a_type, ar, off['head'], motion_pred=motion_pred[k], noise=noise, noise_std=self.noise_std)
            if self.noise_std > 0:       # traj_sam.py:313: the (degenerate, TOP_K = 1) mode draw still advances the generator
                torch.randint(0, 1, (P,), device=self._device)
            ops.step_agent_traj(motion_pred[k], pl.i['p_row'], T, tidx, traj, vel)
            tidx += STEP
            if getattr(self, 'keep_tick_edges', False):
                pl.tick_edges.append((e_a.to_edge_index(), e_m.to_edge_index()))
        st['last_step'] = tidx
        reconst = ops.reconst(emd_flat, ar, off['head'])
        return self._process_rollout(pl, st, motion_pred, reconst, all_t_indices)

    def _process_rollout(self, pl, st, motion_pred, reconst, all_t):
        """traj_sam.py:562-595: concatenate per-tick outputs, per-agent views of the state buffers."""
        task = self.tasks[0]
        n_ticks, P = motion_pred.shape[:2]
        res = {'motion_pred': motion_pred.view(n_ticks * P, 1, STEP, 5),
               'motion_prob': torch.ones(n_ticks * P, 1, device=self._device),
               'reconst_pred': reconst.repeat(n_ticks, 1)}
        names, agent_names = [], []
        for b, ids in enumerate(pl.policy_ids):
            agent_names += [f'{b}-{a}' for a in ids]
        for t in all_t:
            names += [f'{n}-{t}' for n in agent_names]
        res['pair_names'] = names
        res['rollout_trajs'] = _RolloutTrajs(agent_names, pl.p_b * pl.N + pl.p_n, st)
        res['_state'] = st
        return {task: res}
