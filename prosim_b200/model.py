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


class _RolloutTrajs(dict):
    """``rollout_trajs`` of the reference (traj_sam.py:582-593): {"b-id": {traj [steps, 4], init_pos [2], init_heading [1],
    vel [steps, 2]}} as views of the state buffers.  A real ``dict`` (the reference returns one) that fills itself on first
    use: the per-agent views come from four ``unbind`` calls instead of 8 Python indexing ops per agent per forward (35 ms
    at 4096 agents), and a caller that only reads ``_state`` never pays for them."""

    def __init__(self, names, rows, st, last_step=None):
        super().__init__()
        self._names, self._rows, self._st = list(names), [int(r) for r in rows], st
        self._last = int(last_step) if last_step is not None else st['traj'].shape[2]
        self._filled = False

    def _fill(self):
        if not self._filled:
            self._filled = True
            st = self._st
            T = st['traj'].shape[2]
            traj = st['traj'].view(-1, T, 4)[:, HIST:self._last].unbind(0)
            vel = st['vel'].view(-1, T, 2)[:, HIST:self._last].unbind(0)
            pos = st['init_pos'].view(-1, 2).unbind(0)
            head = st['init_heading'].view(-1, 1).unbind(0)
            dict.update(self, {n: {'traj': traj[r], 'init_pos': pos[r], 'init_heading': head[r], 'vel': vel[r]}
                               for n, r in zip(self._names, self._rows)})
        return self

    def __getitem__(self, name):
        return dict.__getitem__(self._fill(), name)

    def __iter__(self):
        return iter(self._names)

    def __len__(self):
        return len(self._names)

    def __contains__(self, name):
        return dict.__contains__(self._fill(), name)

    def get(self, name, default=None):
        return dict.get(self._fill(), name, default)

    def keys(self):
        return dict.keys(self._fill())

    def items(self):
        return dict.items(self._fill())

    def values(self):
        return dict.values(self._fill())

    def __repr__(self):
        return dict.__repr__(self._fill())


def _norm_device(device):
    """'cuda' -> 'cuda:<current>' (tensor.device always carries an index; comparisons need one too)."""
    d = torch.device(device)
    if d.type == 'cuda' and d.index is None and torch.cuda.is_available():
        d = torch.device('cuda', torch.cuda.current_device())
    return d


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
        self.rollout_top_k = cfg.ROLLOUT.POLICY.TOP_K                  # traj_sam.py:25-26
        self.rollout_top_k_train = cfg.ROLLOUT.POLICY.TOP_K_TRAIN
        self.hist_step = cfg.DATASET.FORMAT.HISTORY.STEPS
        assert self.rollout_steps == STEP and self.hist_step == HIST and cfg.DATASET.FORMAT.TARGET.STEPS == STEP
        self.num_layers = cfg.MODEL.POLICY.ACT_DECODER.ATTN.NUM_LAYER
        self.obs_fusion = cfg.MODEL.OBS_UPDATE.FUSION                  # 'replace' | 'mlp' (attn_fusion.py:238-251)
        self.attn_update = bool(cfg.MODEL.OBS_UPDATE.ATTN_UPDATE)
        self.cond_layers = cfg.MODEL.CONDITION_TRANSFORMER.NLAYER
        self.mode = 'val'
        # act_decoder.py:113-115: Gaussian noise on the predicted per-step displacements; the draws come from torch's CUDA
        # generator with the reference's shapes ([P, 1, 10, 2] per tick), noise_fn can be replaced to inject given draws
        self.noise_std = float(cfg.MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD)
        self.noise_fn = lambda shape: torch.randn(shape, device=self._device, dtype=torch.float32)
        self._device = _norm_device(device if device is not None else 'cuda')
        self._sd = None
        self._arena = None
        self._off = None
        self._bufs = {}
        self._plan_cache = []
        # the reference's sub-module handles (traj_sam.py:28-57) as far as the rollout drives them from outside
        self.scene_encoder = SceneEncoderB200(self)
        self.policy = PolicyB200(self)
        self.fused_tick = True      # policy ticks through prosim_policy_tick (one C call); False = one call per kernel family
        used = [t for t in cfg.PROMPT.CONDITION.MOTION_TAG.USED_TAGS if t in weights.V_ACTION_TAG_ID]   # condition_encoders.py:58
        if 'v_action_tag' in self.cond_types and tuple(used) != weights.V_ACTION_TAGS:
            raise NotImplementedError('PROMPT.CONDITION.MOTION_TAG.USED_TAGS differs from the released list')
        lut = -torch.ones(16, dtype=torch.int64)        # tag id (motion_tag_utils.py:4-15) -> position in USED_TAGS
        for pos, tag in enumerate(used):
            lut[weights.V_ACTION_TAG_ID[tag]] = pos
        self._tag_slot = lut
        if state_dict is None:
            state_dict = weights.random_state_dict(0, self.cond_types, obs_fusion=self.obs_fusion)
        self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ weights
    @property
    def device(self):
        return self._device

    def state_dict(self, *args, **kwargs):
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Same contract as nn.Module.load_state_dict: strict raises on missing / unexpected keys; strict=False (how the
        reference loads its checkpoints, trainer.py:162-163) keeps the current value of a missing parameter (the seeded
        initialisation before the first load) and ignores unexpected keys; both lists are returned."""
        specs = weights.param_specs(self.cond_types, self.num_layers, self.cond_layers, self.obs_fusion)
        want = [n for n, _, _ in specs]
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in set(want)]
        if strict and (missing or unexpected):
            raise RuntimeError(f'load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}')
        bad = [k for k, shp, _ in specs if k in state_dict and tuple(state_dict[k].shape) != tuple(shp)]
        if bad:
            raise RuntimeError(f'load_state_dict: size mismatch for {bad[:5]}')
        base = self._sd if self._sd is not None else (weights.random_state_dict(0, self.cond_types, obs_fusion=self.obs_fusion)
                                                      if missing else {})
        self._sd = {k: (state_dict[k] if k in state_dict else base[k]).detach().float().cpu().clone() for k in want}
        arena, self._off = weights.pack_model(self._sd, self.num_layers, self.cond_layers)
        lib.load()  # fail loudly here, not at the first kernel call, if the native library is absent
        self._arena = arena.to(self._device)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def to(self, device=None, *a, **k):
        """Moves the packed weight arena; dtype arguments are rejected (the path computes in fp32 only)."""
        if isinstance(device, torch.dtype) or k.get('dtype') is not None or any(isinstance(x, torch.dtype) for x in a):
            raise NotImplementedError('ProSimB200 computes in fp32; no other dtype is built')
        if device is not None and _norm_device(device) != self._device:
            if torch.device(device).type != 'cuda':
                raise lib.ProSimLibError('ProSimB200 runs on CUDA devices only (no CPU path)')
            self._device = _norm_device(device)
            self._arena = self._arena.to(self._device)
            self._bufs = {}
            self._plan_cache = []
        return self

    def cuda(self, device=None):
        return self.to(torch.device('cuda', device) if isinstance(device, int) else (device or 'cuda'))

    def half(self):
        return self.to(torch.float16)

    def bfloat16(self):
        return self.to(torch.bfloat16)

    def double(self):
        return self.to(torch.float64)

    def float(self):
        return self

    def eval(self):
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError('ProSimB200 is an inference (rollout) model; training is out of scope')
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
    def _check_batch(self, batch):
        """The kernels hard-code the released data format (24 features x 11 steps per agent, 11 features x 19 vectors per
        polyline; config.check_supported pins the config side): any other layout is rejected here instead of being read
        out of bounds on the device."""
        ex = batch.extras
        obs, mp = ex['init_obs'], ex['init_map']
        prm = ex['prompt'][self.tasks[0]]

        def chk(t, name, tail, dtype, lead=None):
            if not isinstance(t, torch.Tensor):
                raise TypeError(f'{name} must be a tensor')
            if tuple(t.shape[len(t.shape) - len(tail):]) != tuple(tail) or (lead is not None and tuple(t.shape[:len(lead)]) != tuple(lead)):
                raise ValueError(f'{name}: shape {tuple(t.shape)} does not end in {tuple(tail)} (leading {lead})')
            if t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
                raise TypeError(f'{name}: dtype {t.dtype}, expected {dtype}')
            if not t.is_contiguous():
                raise ValueError(f'{name} must be contiguous')
            if t.device != self._device:
                raise lib.ProSimLibError(f'{name} is on {t.device}, the model on {self._device} (batch.to(device) first; no CPU path)')

        B, A = obs['input'].shape[:2]
        M = mp['input'].shape[1]
        N = prm['prompt_mask'].shape[1]
        bmask = (torch.bool, torch.uint8)
        for name, d in [('init_obs', obs)] + [(f'fut_obs[{t}]', ex['fut_obs'][t]) for t in ex['fut_obs'].keys()]:
            chk(d['input'], f'{name}.input', (HIST, 24), torch.float32, (B, A))
            chk(d['mask'], f'{name}.mask', (HIST, 24), bmask, (B, A))
            chk(d['position'], f'{name}.position', (2,), torch.float32, (B, A))
            chk(d['heading'], f'{name}.heading', (A,), torch.float32, (B,))
        chk(mp['input'], 'init_map.input', (19, 11), torch.float32, (B, M))
        chk(mp['mask'], 'init_map.mask', (19,), bmask, (B, M))
        chk(mp['position'], 'init_map.position', (1, 2), torch.float32, (B, M))
        chk(mp['heading'], 'init_map.heading', (1,), torch.float32, (B, M))
        chk(prm['prompt'], 'prompt.prompt', (7,), torch.float32, (B, N))
        chk(prm['prompt_mask'], 'prompt.prompt_mask', (N,), bmask, (B,))
        chk(prm['position'], 'prompt.position', (2,), torch.float32, (B, N))
        chk(prm['heading'], 'prompt.heading', (1,), torch.float32, (B, N))
        chk(prm['agent_type'], 'prompt.agent_type', (N,), torch.int64, (B,))

    def _plan(self, batch, private=False):
        """private: build a plan that is neither looked up in nor added to the plan cache and owns its device index maps
        (GraphedForward overwrites them in place on every replay)."""
        if getattr(batch, '_b200_plan', None) is not None and not private:
            return batch._b200_plan
        ex = batch.extras
        obs, mp = ex['init_obs'], ex['init_map']
        prm = ex['prompt'][self.tasks[0]]
        if obs['input'].is_cuda:
            self._check_batch(batch)
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
        for ent in ([] if private else self._plan_cache):
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
        pl.obs_id_lists = [obs['agent_ids']] + [f['agent_ids'] for f in futs]
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
            if self.attn_update:
                ints[f'agent_scene{i}'] = np.repeat(np.arange(B), a_cnt)
            if self.obs_fusion == 'mlp':
                # token row of every observation slot at this tick, and (i >= 1) the agents observed at tick i - 1 too:
                # pairs (old token row, new token row) for obs_update_mlp (attn_fusion.py:180-189)
                pos_of = -np.ones(B * A, dtype=np.int64)
                pos_of[ints[f'agent_rows{i}']] = np.arange(int(a_cnt.sum()))
                if i > 0:
                    new_ids, old_ids = pl.obs_id_lists[i], pl.obs_id_lists[i - 1]
                    f_old, f_new = [], []
                    for b in range(B):
                        lut = {a: k for k, a in enumerate(old_ids[b])}
                        for k_new, a in enumerate(new_ids[b]):
                            if a in lut and pos_of[b * A + k_new] >= 0 and prev_pos_of[b * A + lut[a]] >= 0:
                                f_old.append(prev_pos_of[b * A + lut[a]])
                                f_new.append(pos_of[b * A + k_new])
                    ints[f'fuse_old{i}'], ints[f'fuse_new{i}'] = np.array(f_old, dtype=np.int64), np.array(f_new, dtype=np.int64)
                prev_pos_of = pos_of
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
        if not private:
            self._plan_cache.insert(0, (sig, mask_bytes, id_lists, pl))
            del self._plan_cache[4:]
        return pl

    # ------------------------------------------------------------------ reference API
    def forward(self, batch, mode):
        """traj_sam.py:59-71."""
        if not batch.extras['init_obs']['input'].is_cuda:
            raise lib.ProSimLibError('ProSimB200 needs the batch on the GPU (batch.to(device)); there is no CPU path')
        self.mode = mode
        with torch.cuda.device(self._device):        # every launch goes to this device's current stream
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
        pl.edges_enc = (e_a, e_s)
        # the reference's flat arrays (scene_tokens / scene_pos / scene_ori / scene_batch_idx / scene_type, map rows first)
        # are views / concatenations made on first access (_SceneEmbs); the model itself uses the private entries
        return _SceneEmbs({'obs_mask': obs['mask'].all(-1).any(-1), 'map_mask': mp['mask'].any(-1),
                           'max_map_num': pl.M, 'max_agent_num': pl.A, '_plan': pl, '_tok': tok, '_tok_pos': tok_pos,
                           '_tok_ori': tok_ori, '_agent': (tok[NM:], tok_pos[NM:], tok_ori[NM:]), '_slot': 0, '_shared': {}})

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
            if scene_embs.get('_slot', 0) == 0 and '_tok' in scene_embs:
                tok, tok_pos, tok_ori = scene_embs['_tok'], scene_embs['_tok_pos'], scene_embs['_tok_ori']
            else:
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
        """traj_sam.py:597-633 (3-argument form accepted for rollout/gpu_utils.py:196).  The trajectory buffers are
        preallocated for the whole rollout ([B, N, 11 + steps, 4]; the reference grows them with torch.cat each tick):
        ``last_step`` says how many steps are valid, exactly like the reference's ``a_traj[task]['last_step']``."""
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

    def _select_k_emd_from_batch(self, policy_emds, batch):
        """traj_sam.py:402-439."""
        return select_k_emd_from_batch(policy_emds, batch, self.mode, self.rollout_top_k, self.rollout_top_k_train, self._device)

    def rollout_batch(self, batch, scene_embs, policy_emds, policy_agent_ids, agent_trajs, all_t_indices, mode):
        """traj_sam.py:144-175: the closed-loop tick loop, written against the same four methods as the reference
        (step_env -> decode_output -> step_agent_traj, then _process_rollout), so a caller may drive the ticks itself."""
        self.mode = mode if mode is not None else self.mode
        task = self.tasks[0]
        policy_emds = dict(policy_emds)
        policy_emds[task] = self._select_k_emd_from_batch(policy_emds[task], batch)
        all_t = [int(t) for t in all_t_indices]
        pl = self._plan(batch)
        pl.tick_edges = []
        self._last_plan = pl
        model_outputs = []
        for t in all_t:
            scene_embs, agent_positions = self.step_env(scene_embs, agent_trajs, batch, policy_agent_ids, t, all_t)
            model_output = self.decode_output(policy_emds, scene_embs, policy_agent_ids, batch, agent_positions, t, None)
            agent_trajs = self.step_agent_traj(agent_trajs, model_output, policy_agent_ids, t, mode)
            model_outputs.append(model_output)
        return self._process_rollout(agent_trajs, model_outputs, policy_agent_ids)

    def step_env(self, scene_embs, a_traj, batch, policy_agent_ids, t, all_t_indices):
        """traj_sam.py:205-274: current world pose of every policy agent; for every tick but the first of
        ``all_t_indices`` the agents' 11-step observation windows are rebuilt in the frame of their last step, written
        into ``batch.extras['fut_obs'][t]`` in place (like the reference) and re-encoded (``_update_scene_emb``).
        Returns (scene_embs of this tick, agent positions {'position' [B, N, 2], 'heading' [B, N, 1]})."""
        task = self.tasks[0]
        pl = self._plan(batch)
        self._last_plan = pl
        st = a_traj[task]
        T = st['traj'].shape[2]
        tidx = int(st['last_step'])
        traj, vel = st['traj'].view(-1, T, 4), st['vel'].view(-1, T, 2)
        init_pos, init_heading = st['init_pos'].view(-1, 2), st['init_heading'].view(-1)
        p_pos, p_ori = self._buf('p_pos', (pl.P, 2)), self._buf('p_ori', (pl.P,))
        all_t = [int(x) for x in all_t_indices]
        t_idx = all_t.index(int(t))                  # the reference decides "first tick" by position in the list it was given
        if t_idx == 0:
            ops.step_env(traj, vel, init_pos, init_heading, pl.i['p_row'], pl.i['p_slot0'], T, tidx, p_pos, p_ori)
            scene_embs_next = scene_embs
        else:
            i = pl.all_t.index(int(t))               # plan slot of this tick's fut_obs (index maps, valid-agent rows)
            fut = batch.extras['fut_obs'][t]
            ops.step_env(traj, vel, init_pos, init_heading, pl.i['p_row'], pl.i[f'p_slot{i}'], T, tidx, p_pos, p_ori,
                         fut=(fut['input'], fut['mask'], fut['position'], fut['heading']))
            old_obs = batch.extras['fut_obs'][all_t[t_idx - 1]] if t_idx > 1 else batch.extras['init_obs']
            scene_embs_next = self._update_scene_emb(scene_embs, fut, old_obs['agent_ids'], _slot=i)
        return scene_embs_next, _AgentPositions(p_pos, p_ori, pl)

    def _update_scene_emb(self, last_scene_embs, batch_obs_new, old_obs_agent_ids, _slot=None):
        """traj_sam.py:541-550 (the reference clones every tensor of the dict first; nothing is modified in place here)."""
        return self.scene_encoder.update_scene_emb(last_scene_embs, batch_obs_new, old_obs_agent_ids, _slot=_slot)

    def decode_output(self, policy_emds, scene_embs, policy_agent_ids, batch, agent_positions=None, target_t=None,
                      latent_state_dict=None):
        """traj_sam.py:178-202: gather the per-row policy inputs of this tick and run the policy."""
        task = list(policy_emds.keys())[0]
        batch_policy_emd, batch_obs, batch_map, batch_pos, batch_pair_names = self._get_policy_batch_input(
            batch, policy_agent_ids[task], policy_emds[task], scene_embs, agent_positions, target_t)
        latent_state = None if latent_state_dict is None else self.policy.format_latent_state(latent_state_dict, [batch_pair_names])
        all_output = self.get_action(batch_policy_emd, batch_obs, batch_map, batch_pos, [batch_pair_names], latent_state=latent_state)
        all_output['pair_names'] = batch_pair_names
        all_output['_p_row'] = self._plan(batch).i['p_row']
        return {task: all_output}

    def _get_policy_batch_input(self, batch, policy_agent_ids, policy_emds, scene_embs, agent_positions, target_t):
        """traj_sam.py:441-525 for the rollout case (agent_positions given): one row per (scene, policy agent).  The
        reference scatters the flat scene tokens into dense [B, S, 128] tensors which the policy immediately re-flattens
        (act_decoder.py:224-237); here the flat tokens are handed over as they are, with per-scene ranges."""
        if agent_positions is None:
            raise NotImplementedError('open-loop decoding from io_pairs_batch is the training path (out of scope)')
        pl = self._plan(batch)
        rows = pl.i['p_row'].long()
        emd_flat = policy_emds['_emd_flat'] if '_emd_flat' in policy_emds else policy_emds['emd'].reshape(-1, D)[rows].contiguous()
        if '_atype_flat' not in policy_emds:
            policy_emds['_atype_flat'] = policy_emds['agent_type'].reshape(-1)[rows].to(torch.int32).contiguous()
        batch_policy_emd = {'emd': emd_flat, 'agent_type': policy_emds['_atype_flat'], 'batch_idx': pl.i['p_scene'],
                            '_cache': policy_emds}
        for key in ('goal', 'goal_prob', 'goal_point', 'select_idx'):
            if key in policy_emds:
                batch_policy_emd[key] = policy_emds[key].reshape((-1,) + tuple(policy_emds[key].shape[2:]))[rows]
        if isinstance(agent_positions, _AgentPositions):
            batch_pos = {'position': agent_positions.p_pos, 'heading': agent_positions.p_ori.view(-1, 1)}
        else:
            batch_pos = {key: agent_positions[key].reshape(pl.B * pl.N, -1)[rows].contiguous() for key in ('position', 'heading')}
        slot = scene_embs.get('_slot', 0)
        x_a, a_pos, a_ori = scene_embs['_agent']
        NM = pl.NM
        batch_obs = {'tokens': x_a, 'pos': a_pos, 'ori': a_ori, 'seg': pl.i[f'seg_agent{slot}'].view(-1, 4), 'max_per_scene': pl.max_a}
        batch_map = {'tokens': scene_embs['_tok'][:NM], 'pos': scene_embs['_tok_pos'][:NM], 'ori': scene_embs['_tok_ori'][:NM],
                     'seg': pl.i['seg_map'].view(-1, 4), 'max_per_scene': pl.max_m, '_cache': scene_embs['_shared']}
        if '_names' not in policy_emds:
            policy_emds['_names'] = [f'{b}-{a}' for b, ids in enumerate(policy_agent_ids) for a in ids]
        batch_pair_names = [f'{n}-{target_t}' for n in policy_emds['_names']]
        return batch_policy_emd, batch_obs, batch_map, batch_pos, batch_pair_names

    def get_action(self, policy_emb, obs_data, map_data, pos_data, pair_names, latent_state=None):
        """traj_sam.py:635-640."""
        pair_names_all = []
        for pair_name in pair_names:
            pair_names_all += pair_name
        return self.policy(policy_emb, obs_data, map_data, pos_data, pair_names_all, latent_state)

    def step_agent_traj(self, a_traj, model_output, policy_agent_ids, t, mode):
        """traj_sam.py:276-349: rotate the 10 predicted steps into each agent's t0 frame and append them.  Rows are matched
        by position (row p of the policy output is policy agent p of the plan; the reference searches the pair-name strings)."""
        task = self.tasks[0]
        out = model_output[task]
        st = a_traj[task]
        B, N, T = st['traj'].shape[:3]
        tidx = int(st['last_step'])
        if tidx + STEP > T:            # a caller-made buffer without room: grow it like the reference's torch.cat
            grow = tidx + STEP - T
            st['traj'] = torch.cat([st['traj'], torch.zeros(B, N, grow, 4, device=self._device)], dim=2)
            st['vel'] = torch.cat([st['vel'], torch.zeros(B, N, grow, 2, device=self._device)], dim=2)
            T = tidx + STEP
        motion_pred = out['motion_pred']
        P = motion_pred.shape[0]
        rollout_k = self.rollout_top_k_train if mode == 'train' else self.rollout_top_k
        rollout_k = min(rollout_k, motion_pred.shape[1])
        if rollout_k > 1 or self.noise_std > 0:       # traj_sam.py:311-313 (the draw also keeps the generator stream aligned)
            top = torch.topk(out['motion_prob'], rollout_k, dim=1)[1]
            rand_idxs = torch.randint(0, rollout_k, (P,), device=self._device)
            if motion_pred.shape[1] > 1:
                sel = top[torch.arange(P, device=self._device), rand_idxs]
                motion_pred = motion_pred[torch.arange(P, device=self._device), sel][:, None]
        p_row = out.get('_p_row')
        if p_row is None:
            p_row = self._buf_rows(policy_agent_ids[task], N)
        ops.step_agent_traj(motion_pred.contiguous(), p_row, T, tidx, st['traj'].view(-1, T, 4), st['vel'].view(-1, T, 2))
        st['last_step'] = tidx + STEP
        return a_traj

    def _buf_rows(self, agent_ids, N):
        key = tuple(len(x) for x in agent_ids) + (N,)
        cache = self.__dict__.setdefault('_rows_cache', {})
        if key not in cache:
            rows = np.concatenate([b * N + np.arange(len(ids)) for b, ids in enumerate(agent_ids)]).astype(np.int32)
            cache[key] = torch.from_numpy(rows).to(self._device)
        return cache[key]

    def _process_rollout(self, agent_trajs, model_outputs, policy_agent_ids):
        """traj_sam.py:562-595: concatenate the per-tick outputs; per-agent views of the state buffers."""
        task = self.tasks[0]
        st = agent_trajs[task]
        res = {}
        for key in ('motion_pred', 'motion_prob', 'goal', 'pair_names', 'goal_prob', 'goal_point', 'select_idx', 'reconst_pred'):
            if key not in model_outputs[0][task]:
                continue
            if key == 'pair_names':
                res[key] = [n for mo in model_outputs for n in mo[task][key]]
            else:
                res[key] = torch.cat([mo[task][key] for mo in model_outputs], dim=0)
        N = st['traj'].shape[1]
        agent_names = [f'{b}-{a}' for b, ids in enumerate(policy_agent_ids[task]) for a in ids]
        rows = np.concatenate([b * N + np.arange(len(ids)) for b, ids in enumerate(policy_agent_ids[task])]) \
            if agent_names else np.zeros(0, np.int64)
        res['rollout_trajs'] = _RolloutTrajs(agent_names, rows, st, int(st['last_step']))
        res['_state'] = st
        return {task: res}


def select_k_emd_from_batch(policy_emds, batch, mode, rollout_top_k, rollout_top_k_train, device):
    """traj_sam.py:402-439.  One policy token per agent (DECODER.GOAL_PRED disabled, emd [B, N, D]): nothing to select.
    With K goal-conditioned tokens per agent (emd [B, N, K, D] + goal_prob [B, N, K] / goal_point [B, N, K, 2]): inference picks
    uniformly among the top-``ROLLOUT.POLICY.TOP_K`` goal probabilities (one ``torch.randint`` on the host generator, like the
    reference), training the token whose goal point is closest to the ground-truth goal.  Index bookkeeping in torch."""
    if policy_emds['emd'].ndim == 3:
        return policy_emds
    goal_prob, goal_point = policy_emds['goal_prob'], policy_emds['goal_point']
    B, N, K, Dm = policy_emds['emd'].shape
    if mode == 'train':
        gt_goal = batch.extras['io_pairs_batch']['goal'][:, 0, :]
        goal_idxs = torch.min(torch.norm(goal_point - gt_goal[:, :, None, :], dim=-1), dim=-1)[1]
    else:
        rollout_k = min(rollout_top_k, K)
        top = torch.topk(goal_prob, rollout_k, dim=-1)[1]
        rand_idxs = torch.randint(0, rollout_k, (B, N,)).to(device=device)
        goal_idxs = torch.gather(top, -1, rand_idxs[..., None]).squeeze(-1)
    policy_emds = {k: v for k, v in policy_emds.items() if not k.startswith('_')}
    policy_emds['select_idx'] = goal_idxs
    policy_emds['emd'] = torch.gather(policy_emds['emd'], -2, goal_idxs[..., None, None].repeat(1, 1, 1, Dm)).squeeze(-2)
    policy_emds['goal'] = torch.gather(goal_point, -2, goal_idxs[..., None, None].repeat(1, 1, 1, 2)).squeeze(-2)
    return policy_emds


class _AgentPositions(Mapping):
    """``a_pos`` of traj_sam.py:209-215: {'position': [B, N, 2], 'heading': [B, N, 1]}.  The kernels produce the poses of the P
    policy rows; the dense [B, N] tensors are scattered on first access (only a caller outside the model ever asks)."""

    def __init__(self, p_pos, p_ori, pl):
        self.p_pos, self.p_ori, self._pl, self._dense = p_pos, p_ori, pl, {}

    def __getitem__(self, key):
        if key not in ('position', 'heading'):
            raise KeyError(key)
        if key not in self._dense:
            pl = self._pl
            w = 2 if key == 'position' else 1
            out = torch.zeros(pl.B * pl.N, w, device=self.p_pos.device)
            out[pl.i['p_row'].long()] = (self.p_pos if key == 'position' else self.p_ori.view(-1, 1))
            self._dense[key] = out.view(pl.B, pl.N, w)
        return self._dense[key]

    def __iter__(self):
        return iter(('position', 'heading'))

    def __len__(self):
        return 2


class _SceneEmbs(dict):
    """Scene-embedding dict of the reference (scene_encoder/base.py:31-46, attn_fusion.py:205-236).  The model keeps map
    tokens and agent tokens in separate buffers (map tokens are never copied between ticks); the reference's flat
    ``scene_tokens / scene_pos / scene_ori / scene_batch_idx / scene_type`` arrays (map rows first, then agent rows) are
    concatenated on first access."""
    _LAZY = ('scene_tokens', 'scene_pos', 'scene_ori', 'scene_batch_idx', 'scene_type', 'obs_mask')

    def __missing__(self, key):
        if key not in self._LAZY:
            raise KeyError(key)
        pl, NM = self['_plan'], self['_plan'].NM
        x_a, a_pos, a_ori = self['_agent']
        slot = self.get('_slot', 0)
        dev = x_a.device
        if key == 'obs_mask':
            val = self['_obs_mask_src'].all(-1).any(-1)
        elif key == 'scene_tokens':
            val = torch.cat([self['_tok'][:NM], x_a])
        elif key == 'scene_pos':
            val = torch.cat([self['_tok_pos'][:NM], a_pos])
        elif key == 'scene_ori':
            val = torch.cat([self['_tok_ori'][:NM], a_ori]).view(-1, 1)
        elif key == 'scene_batch_idx':
            seg = pl.i[f'seg_agent{slot}'].view(-1, 4)[:, 1].long()
            val = torch.cat([pl.i['tok_scene'][:NM].long(), torch.repeat_interleave(torch.arange(pl.B, device=dev), seg)])
        else:
            val = torch.cat([torch.zeros(NM, dtype=torch.long, device=dev), torch.ones(x_a.shape[0], dtype=torch.long, device=dev)])
        self[key] = val
        return val

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._LAZY


@registry.register_scene_encoder(name='attn_fusion_relpe_b200')
class SceneEncoderB200(nn.Module):
    """``model.scene_encoder`` of the reference (AttentionSceneEncoderRelPE, scene_encoder/attn_fusion.py:13-251) as far as
    the rollout drives it: ``forward(batch)`` = encode_scene, ``update_scene_emb`` = the per-tick re-encoding of the agent
    histories.  Holds no parameters of its own (the model owns the packed weight arena)."""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, '_m', model)

    def forward(self, batch):
        return self._m.encode_scene(batch)

    def update_scene_emb(self, scene_emds, batch_obs, old_obs_agent_ids, _slot=None):
        """attn_fusion.py:238-251 with OBS_UPDATE.FUSION = 'replace': the agent tokens become the PointNet encoding of the
        new observation windows, the map tokens are untouched (attn_fusion.py:205-236 re-concatenates the flat arrays; here
        the new agent tokens simply live in their own buffer)."""
        m = self._m
        pl = scene_emds['_plan']
        if _slot is None:           # called from outside the model: find the tick this observation belongs to
            _slot = next((i for i, ids in enumerate(pl.obs_id_lists) if i > 0 and (ids is batch_obs['agent_ids'] or ids == batch_obs['agent_ids'])), None)
            if _slot is None:
                raise ValueError('update_scene_emb: batch_obs is not one of the fut_obs entries this batch was planned with')
        na = pl.NA[_slot]
        max_na = max(pl.NA)
        # two alternating agent-token buffers: the 'mlp' fusion reads the previous tick's tokens while it writes this tick's
        x_a = m._buf(f'x_a{_slot & 1}', (max_na, D))[:na]
        a_pos, a_ori = m._buf(f'a_pos{_slot & 1}', (max_na, 2))[:na], m._buf(f'a_ori{_slot & 1}', (max_na,))[:na]
        rows = pl.i[f'agent_rows{_slot}']
        ar, off = m._arena, m._off
        ops.pointnet(0, batch_obs['input'], batch_obs['mask'], rows, ar, off['obs_enc'], out=x_a, tc_off=off['obs_enc_tc'])
        ops.gather_pose(batch_obs['position'], batch_obs['heading'], rows, a_pos, a_ori)
        if m.obs_fusion == 'mlp':
            # attn_fusion.py:177-203: agents observed at the previous tick too get MLP([old token | new token])
            if scene_emds.get('_slot', 0) != _slot - 1:
                raise ValueError("OBS_UPDATE.FUSION = 'mlp' needs the scene embedding of the previous tick")
            if pl.i[f'fuse_new{_slot}'].shape[0] > 0:
                ops.obs_fuse(scene_emds['_agent'][0], pl.i[f'fuse_old{_slot}'], x_a, pl.i[f'fuse_new{_slot}'], ar, off['obs_fuse'])
        if m.attn_update:
            self._update_scene_emb_attn(scene_emds, pl, _slot, x_a, a_pos, a_ori)
        out = _SceneEmbs({k: v for k, v in dict.items(scene_emds) if k not in _SceneEmbs._LAZY})
        out['_agent'] = (x_a, a_pos, a_ori)
        out['_slot'] = _slot
        out['_obs_mask_src'] = batch_obs['mask']
        out['max_agent_num'] = batch_obs['input'].shape[1]
        return out


    def _update_scene_emb_attn(self, scene_emds, pl, slot, x_a, a_pos, a_ori):
        """attn_fusion.py:136-175 (OBS_UPDATE.ATTN_UPDATE): redo the encoder's agent self-attention and map -> agent attention
        on radius graphs over the new agent tokens, in place: 6 x (a2a layer, s2s layer with the map tokens as sources).  The
        map-side K'|V' of the s2s layers do not change during a rollout: computed once per scene encoding."""
        m = self._m
        ar, off = m._arena, m._off
        lf = weights.ATTN_LAYER_FLOATS
        L = m.num_layers
        ecfg = m.config.MODEL.SCENE_ENCODER.ATTN
        NM = pl.NM
        na = x_a.shape[0]
        x_m, m_pos, m_ori = scene_emds['_tok'][:NM], scene_emds['_tok_pos'][:NM], scene_emds['_tok_ori'][:NM]
        dim_t = ar[off['dim_t16']:off['dim_t16'] + 16]
        cap = ecfg.MAX_NUM_NEIGH
        a_scene = pl.i[f'agent_scene{slot}']
        seg_a = pl.i[f'seg_agent{slot}'].view(-1, 4)
        e_a = ops.radius_edges(a_pos, a_scene, a_pos, seg_a, ecfg.AGENT_RADIUS, cap, min(cap + 1, max(pl.max_a, 1)), drop_self=True)
        e_m = ops.radius_edges(a_pos, a_scene, m_pos, pl.i['seg_map'].view(-1, 4), ecfg.SCENE_RADIUS, cap, min(cap, max(pl.max_m, 1)))
        ops.edge_pe(e_a, a_pos, a_ori, a_pos, a_ori, dim_t, z=m._buf('z_upd_a', (na * e_a.stride, 96)))
        ops.edge_pe(e_m, a_pos, a_ori, m_pos, m_ori, dim_t, z=m._buf('z_upd_m', (na * e_m.stride, 96)))
        shared = scene_emds['_shared']
        kv_m = shared.get('kv_m_enc')
        if kv_m is None:
            kv_m = shared['kv_m_enc'] = ops.attn_kv(x_m, ar, off['enc_s2s'], L, lf, kv=m._buf('kv_m_enc', (L, NM, 2 * D)))
        ws = m._buf('attn_ws_upd', (lib.load().prosim_attn_workspace_floats(na, na, max(e_a.stride, e_m.stride)),))
        ops.attn_stack(x_a, L, ops.stack_side(ar, off['enc_a2a'], e_a), ops.stack_side(ar, off['enc_s2s'], e_m, kv_m), out=x_a,
                       workspace=ws)
        pl.edges_upd = (e_a, e_m)


@registry.register_policy(name='rel_pe_temporal_b200')
class PolicyB200(nn.Module):
    """``model.policy`` of the reference (Policy_RelPE_Temporal -> PolicyNoRNN -> AttnRelPE -> ActDecoder; policy/base.py:9-23,
    temporal_ar.py:66-92, act_decoder.py:78-135, :239-279): one policy tick for P rows."""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, '_m', model)

    def format_latent_state(self, latent_state_dict, all_batch_pair_names):
        return None                 # PolicyNoRNN keeps no recurrent state (temporal_ar.py:71-73)

    def forward(self, policy_emd, batch_obs, batch_map, batch_pos, pair_names, latent_state):
        m = self._m
        ar, off = m._arena, m._off
        lf = weights.ATTN_LAYER_FLOATS
        L = m.num_layers
        acfg = m.config.MODEL.POLICY.ACT_DECODER.ATTN
        dev = m._device
        batch_obs, batch_map = _flat_tokens(batch_obs, dev), _flat_tokens(batch_map, dev)
        emd = policy_emd['emd']
        P = emd.shape[0]
        p_scene = policy_emd['batch_idx'].to(torch.int32)
        a_type = policy_emd['agent_type'].to(torch.int32)
        p_pos = batch_pos['position'].reshape(P, 2).contiguous()
        p_ori = batch_pos['heading'].reshape(P).contiguous()
        dim_t = ar[off['dim_t16']:off['dim_t16'] + 16]
        x_a, a_pos, a_ori = batch_obs['tokens'], batch_obs['pos'], batch_obs['ori'].reshape(-1)
        x_m, m_pos, m_ori = batch_map['tokens'], batch_map['pos'], batch_map['ori'].reshape(-1)
        na, nm = x_a.shape[0], x_m.shape[0]
        stride_a = min(acfg.MAX_NUM_NEIGH, max(int(batch_obs['max_per_scene']), 1))
        stride_m = min(acfg.MAX_NUM_NEIGH, max(int(batch_map['max_per_scene']), 1))
        # map-side K'|V' of all layers: the map tokens and the weights never change during a rollout, so they are computed
        # once per scene encoding and reused by every tick (the reference recomputes them 8 times)
        cache = batch_map.get('_cache')
        kv_m = None if cache is None else cache.get('kv_m')
        if kv_m is None:
            kv_m = ops.attn_kv(x_m, ar, off['pol_m2p'], L, lf, kv=m._buf('kv_m', (L, nm, 2 * D)))
            if cache is not None:
                cache['kv_m'] = kv_m
        fuse = m._buf('fuse', (P, D))
        noise = m.noise_fn((P, 1, STEP, 2)) if m.noise_std > 0 else None
        debug = getattr(m, 'keep_tick_edges', False) or getattr(m, 'edge_log', None) is not None or not m.fused_tick
        if not debug:
            # the whole tick through ONE C call (prosim_policy_tick): same kernels, same bits as the step-by-step path below
            cfg = ops.tick_cfg(P, na, nm, batch_obs['max_per_scene'], batch_map['max_per_scene'], acfg.MAX_NUM_NEIGH, L)
            tick_ws = m._buf('tick_ws', (ops.tick_workspace_bytes(cfg) + 256,), torch.uint8)
            tick_ws = tick_ws[(-tick_ws.data_ptr()) % 256:]
            motion_pred = ops.policy_tick(cfg, emd, a_type, p_scene, p_pos, p_ori, x_a, a_pos, a_ori, batch_obs['seg'], m_pos, m_ori,
                                          batch_map['seg'], kv_m, ar, off['pol_a2p'], off['pol_m2p'], off['head'], dim_t,
                                          acfg.AGENT_RADIUS, acfg.MAP_RADIUS, tick_ws, fuse, noise=noise, noise_std=m.noise_std)
            return self._result(policy_emd, emd, motion_pred, latent_state)
        nbr_a, deg_a = m._buf('nbr_a', (P * stride_a,), torch.int32), m._buf('deg_a', (P,), torch.int32)
        nbr_m, deg_m = m._buf('nbr_m', (P * stride_m,), torch.int32), m._buf('deg_m', (P,), torch.int32)
        z_a, z_m = m._buf('z_a', (P * stride_a, 96)), m._buf('z_m', (P * stride_m, 96))
        e_a = ops.radius_edges(p_pos, p_scene, a_pos, batch_obs['seg'], acfg.AGENT_RADIUS, acfg.MAX_NUM_NEIGH, stride_a,
                               nbr=nbr_a, deg=deg_a)
        e_m = ops.radius_edges(p_pos, p_scene, m_pos, batch_map['seg'], acfg.MAP_RADIUS, acfg.MAX_NUM_NEIGH, stride_m,
                               nbr=nbr_m, deg=deg_m)
        e_m.warps_per_row = 2     # map radius 50 m: a policy row typically sees ~40 of the scene's polylines
        ops.edge_pe(e_a, p_pos, p_ori, a_pos, a_ori, dim_t, z=z_a)
        ops.edge_pe(e_m, p_pos, p_ori, m_pos, m_ori, dim_t, z=z_m)
        kva = ops.attn_kv(x_a, ar, off['pol_a2p'], L, lf, kv=m._buf('kv_a', (L, na, 2 * D)))
        ws = m._buf('attn_ws', (lib.load().prosim_attn_workspace_floats(P, 0, max(stride_a, stride_m)),))
        ops.attn_stack(emd, L, ops.stack_side(ar, off['pol_a2p'], e_a, kva), ops.stack_side(ar, off['pol_m2p'], e_m, kv_m),
                       out=fuse, workspace=ws)
        m._note_edges('pol_a2p', e_a, L)
        m._note_edges('pol_m2p', e_m, L)
        motion_pred = ops.policy_head(fuse, a_type, ar, off['head'], noise=noise, noise_std=m.noise_std)
        if getattr(m, 'keep_tick_edges', False):
            pl = getattr(m, '_last_plan', None)
            if pl is not None:
                pl.tick_edges.append((e_a.to_edge_index(), e_m.to_edge_index()))
        return self._result(policy_emd, emd, motion_pred, latent_state)

    def _result(self, policy_emd, emd, motion_pred, latent_state):
        m = self._m
        ar, off = m._arena, m._off
        P = emd.shape[0]
        result = {'motion_pred': motion_pred, 'motion_prob': torch.ones(P, 1, device=m._device)}
        if m.config.LOSS.ROLLOUT_TRAJ.USE_GOAL_PRED_LOSS:
            pc = policy_emd.get('_cache')
            rc = None if pc is None else pc.get('_reconst')     # pred_mlp(emd) does not depend on the tick
            if rc is None:
                rc = ops.reconst(emd, ar, off['head'])
                if pc is not None:
                    pc['_reconst'] = rc
            result['reconst_pred'] = rc
        for key in ('goal', 'goal_prob', 'goal_point'):
            if key in policy_emd:
                result[key] = policy_emd[key]
        result['latent_state'] = latent_state
        return result


def _flat_tokens(d, dev):
    """Policy-side token set: the model's flat form {'tokens' [n, 128], 'pos' [n, 2], 'ori' [n], 'seg' int32 [B, 4],
    'max_per_scene'} passes through; the reference's dense form {'input' [B, S, 128], 'mask' [B, S], 'pos' [B, S, 2],
    'ori' [B, S, 1]} (traj_sam.py:356-400) is flattened by its mask the way act_decoder.py:224-237 does."""
    if 'tokens' in d:
        return d
    mask = d['mask'].bool()
    cnt = mask.sum(dim=1).cpu()
    off = torch.cumsum(cnt, 0) - cnt
    seg = torch.stack([off, cnt, torch.zeros_like(cnt), torch.zeros_like(cnt)], dim=1).to(torch.int32).to(dev)
    return {'tokens': d['input'][mask].contiguous(), 'pos': d['pos'][mask].contiguous(), 'ori': d['ori'][mask].reshape(-1).contiguous(),
            'seg': seg, 'max_per_scene': int(cnt.max()) if cnt.numel() else 0}
