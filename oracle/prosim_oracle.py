"""TEST INFRASTRUCTURE (oracle tier B) -- never imported by the product path.

Self-contained CPU restatement, in plain PyTorch, of the reference's closed-loop rollout
(``ProSim.forward``, prosim/models/traj_sam.py:59-175) for the released model shape
(prosim_demo/cfg/waymo_demo.yaml).  It works straight off a ``state_dict`` with the
reference's key names.  Each function cites the reference lines it follows.

PINNING: ``tests/test_oracle_vs_reference.py`` runs this against the reference's own code
(oracle/ref_shim.py, only where /root/reference is mounted) and ``tests/test_oracle_golden.py``
against vectors generated from the reference and committed under tests/golden/ (generator:
tests/golden/make_golden.py).  The reference itself ships no tests, golden vectors or weights
(SURVEY.md section 4), so those reference-generated vectors are the pin.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.
"""
import math

import torch
import torch.nn.functional as F

from . import graph

HIST = 11
STEP = 10
DT = 0.1


# ----------------------------------------------------------------------------- geometry
def wrap_angle(a):
    """models/utils/geometry.py:13-17: -pi + (a + pi) % 2pi (sign-of-divisor modulo)."""
    return -math.pi + (a + math.pi) % (2 * math.pi)


def rotate2d(xy, theta):
    """geometry.py:19-22 batch_rotate_2D."""
    x1 = xy[..., 0] * torch.cos(theta) - xy[..., 1] * torch.sin(theta)
    y1 = xy[..., 1] * torch.cos(theta) + xy[..., 0] * torch.sin(theta)
    return torch.stack([x1, y1], dim=-1)


def fourier_fix(x, num_pos_feats, temperature=10000):
    """layers/fourier_embedding.py:56-79: parameter-free interleaved sin/cos embedding of each scalar."""
    pos = x * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=x.device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    parts = []
    for i in range(pos.shape[-1]):
        p = pos[..., i, None] / dim_t
        parts.append(torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2))
    return torch.cat(parts, dim=-1)


def rel_pe_input(edge_index, ori_dst, pos_dst, ori_src, pos_src):
    """policy/act_decoder.py:203-221 (= decoder/sym_coord.py:44-61, scene_encoder/attn_fusion.py:44-58):
    [|dp|, wrap(theta_src - theta_dst), phi, phi], phi = angle of dp seen from the destination heading."""
    src, dst = edge_index[0], edge_index[1]
    ori_vec = torch.stack([ori_dst.cos().squeeze(-1), ori_dst.sin().squeeze(-1)], dim=-1)
    rel_pos = pos_src[src] - pos_dst[dst]
    rel_ori = wrap_angle(ori_src[src] - ori_dst[dst]).squeeze(-1)
    c = ori_vec[dst]
    phi = torch.atan2(c[..., 0] * rel_pos[..., 1] - c[..., 1] * rel_pos[..., 0],
                      (c[..., :2] * rel_pos[..., :2]).sum(dim=-1))
    return torch.stack([torch.norm(rel_pos, dim=-1), rel_ori, phi, phi], dim=-1)


def segment_softmax(src, index, n):
    """torch_geometric.utils.softmax (call site attention_layer.py:91)."""
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    seg_max = torch.full((n,) + tuple(src.shape[1:]), float('-inf'), dtype=src.dtype)
    seg_max = seg_max.scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    out = (src - seg_max.gather(0, idx)).exp()
    seg_sum = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, out)
    return out / (seg_sum.gather(0, idx) + 1e-16)


class ProSimOracle:
    def __init__(self, state_dict, goal_condition=None, dtype=torch.float32, faithful_bookkeeping=False):
        self.w = {k: v.detach().to(dtype).cpu() for k, v in state_dict.items()}
        self.dtype = dtype
        ce = 'condition_transformers.policy_decoder.condition_encoders.'
        if goal_condition is None or goal_condition is True:        # types in PROMPT.CONDITION.TYPES order
            goal_condition = tuple(t for t in ('goal', 'v_action_tag', 'drag_point') if any(k.startswith(ce + t) for k in self.w))
        self.cond_types = tuple(goal_condition or ())
        self.goal_condition = bool(self.cond_types)
        self.faithful_bookkeeping = faithful_bookkeeping
        self.noise_std = 0.0      # MODEL.POLICY.ACT_DECODER.RANDOM_NOISE_STD (act_decoder.py:113-115)
        self.noise_fn = torch.randn_like
        self.trace = None  # set to [] to record per-tick state for teacher-forced tests
        self.trace_states = []
        # MODEL.OBS_UPDATE (attn_fusion.py:15-19, 238-251): 'mlp' fusion iff the checkpoint carries obs_update_mlp
        self.obs_fusion = 'mlp' if 'scene_encoder.obs_update_mlp.mlp.0.weight' in self.w else 'replace'
        self.attn_update = False

    # ------------------------------------------------------------------------- building blocks
    def lin(self, name, x, bias=True):
        return F.linear(x, self.w[name + '.weight'], self.w.get(name + '.bias') if bias else None)

    def ln(self, name, x):
        return F.layer_norm(x, (x.shape[-1],), self.w[name + '.weight'], self.w[name + '.bias'])

    def mlp(self, prefix, x, n_layers, ret_before_act=False, without_norm=False):
        """layers/mlp.py:475-494."""
        idx = 0
        for i in range(n_layers):
            x = self.lin(f'{prefix}.mlp.{idx}', x)
            idx += 1
            if i < n_layers - 1:
                if not without_norm:
                    x = self.ln(f'{prefix}.mlp.{idx}', x)
                    idx += 1
                x = F.relu(x)
                idx += 1
        if not ret_before_act:
            x = F.relu(x)
        return x

    def pointnet(self, prefix, polylines, mask, n_pre, n_mlp):
        """scene_encoder/pointnet_encoder.py:24-62.  mask [B,N,P]; masked points are zero BEFORE each max."""
        B, N, P, _ = polylines.shape
        feat_valid = self.mlp(f'{prefix}.pre_mlps', polylines[mask], n_pre)
        feat = polylines.new_zeros(B, N, P, feat_valid.shape[-1])
        feat[mask] = feat_valid
        pooled = feat.max(dim=2)[0]
        feat = torch.cat((feat, pooled[:, :, None, :].repeat(1, 1, P, 1)), dim=-1)
        feat_valid = self.mlp(f'{prefix}.mlps', feat[mask], n_mlp)
        buf = feat.new_zeros(B, N, P, feat_valid.shape[-1])
        buf[mask] = feat_valid
        buf = buf.max(dim=2)[0]
        valid = mask.sum(dim=-1) > 0
        out_valid = self.mlp(f'{prefix}.out_mlps', buf[valid], 2, ret_before_act=True, without_norm=True)
        out = buf.new_zeros(B, N, out_valid.shape[-1])
        out[valid] = out_valid
        return out

    def obs_encoder(self, obs):
        """scene_encoder/obs_encoder.py:75-86."""
        pmask = obs['mask'].all(dim=-1)
        return self.pointnet('scene_encoder.obs_encoder', obs['input'].to(self.dtype), pmask, 1, 2), pmask.any(dim=-1)

    def map_encoder(self, mp):
        """scene_encoder/map_encoder.py:67-88."""
        return (self.pointnet('scene_encoder.map_encoder', mp['input'].to(self.dtype), mp['mask'], 3, 2),
                mp['mask'].any(dim=-1))

    def attention_layer(self, prefix, x_src, x_dst, r, edge_index, bipartite):
        """layers/attention_layer.py:56-118 (eval mode: dropout inert)."""
        H, Dh = 8, 16
        x = x_dst
        xs = self.ln(f'{prefix}.attn_prenorm_x_src', x_src)
        xd = self.ln(f'{prefix}.attn_prenorm_x_dst' if bipartite else f'{prefix}.attn_prenorm_x_src', x_dst)
        r = self.ln(f'{prefix}.attn_prenorm_r', r)
        q = self.lin(f'{prefix}.to_q', xd).view(-1, H, Dh)
        k = self.lin(f'{prefix}.to_k', xs, bias=False).view(-1, H, Dh)
        v = self.lin(f'{prefix}.to_v', xs).view(-1, H, Dh)
        src, dst = edge_index[0], edge_index[1]
        k_j = k[src] + self.lin(f'{prefix}.to_k_r', r, bias=False).view(-1, H, Dh)
        v_j = v[src] + self.lin(f'{prefix}.to_v_r', r).view(-1, H, Dh)
        sim = (q[dst] * k_j).sum(dim=-1) * (Dh ** -0.5)
        attn = segment_softmax(sim, dst, x_dst.shape[0])
        msg = v_j * attn.unsqueeze(-1)
        agg = torch.zeros((x_dst.shape[0], H, Dh), dtype=msg.dtype).index_add_(0, dst, msg)
        agg = agg.view(-1, H * Dh)
        g = torch.sigmoid(self.lin(f'{prefix}.to_g', torch.cat([agg, xd], dim=-1)))
        upd = agg + g * (self.lin(f'{prefix}.to_s', xd) - agg)
        x = x + self.ln(f'{prefix}.attn_postnorm', self.lin(f'{prefix}.to_out', upd))
        ff = self.lin(f'{prefix}.ff_mlp.3', F.relu(self.lin(f'{prefix}.ff_mlp.0', self.ln(f'{prefix}.ff_prenorm', x))))
        return x + self.ln(f'{prefix}.ff_postnorm', ff)

    def rel_pe(self, edge_index, ori_dst, pos_dst, ori_src, pos_src):
        return fourier_fix(rel_pe_input(edge_index, ori_dst, pos_dst, ori_src, pos_src), 128 / 4)

    # ------------------------------------------------------------------------- one-time phases
    def encode_scene(self, batch):
        """scene_encoder/base.py:31-46 + attn_fusion.py:78-134."""
        obs, mp = batch.extras['init_obs'], batch.extras['init_map']
        map_emb, map_mask = self.map_encoder(mp)
        obs_emb, obs_mask = self.obs_encoder(obs)
        B = map_emb.shape[0]
        map_b = torch.arange(B).unsqueeze(1).repeat(1, map_emb.shape[1]).view(-1)[map_mask.view(-1)]
        obs_b = torch.arange(B).unsqueeze(1).repeat(1, obs_emb.shape[1]).view(-1)[obs_mask.view(-1)]
        scene_b = torch.cat([map_b, obs_b])
        scene_type = torch.cat([torch.zeros_like(map_b), torch.ones_like(obs_b)])
        x = torch.cat([map_emb.view(-1, 128)[map_mask.view(-1)], obs_emb.view(-1, 128)[obs_mask.view(-1)]])
        f = lambda t: t.to(self.dtype)
        obs_pos = f(obs['position']).view(-1, 2)[obs_mask.view(-1)]
        obs_ori = f(obs['heading']).view(-1, 1)[obs_mask.view(-1)]
        pos = torch.cat([f(mp['position']).view(-1, 2)[map_mask.view(-1)], obs_pos])
        ori = torch.cat([f(mp['heading']).view(-1, 1)[map_mask.view(-1)], obs_ori])
        e_a = graph.knn_graph(obs_pos, k=min(32 * 4, 100), batch=obs_b, loop=True)
        e_s = graph.knn_graph(pos, k=32, batch=scene_b, loop=True)
        pe_a = self.rel_pe(e_a, obs_ori, obs_pos, obs_ori, obs_pos)
        pe_s = self.rel_pe(e_s, ori, pos, ori, pos)
        a_mask = scene_type == 1
        for i in range(6):
            xa = x[a_mask]
            x[a_mask] = self.attention_layer(f'scene_encoder.a2a_attn_layers.{i}', xa, xa, pe_a, e_a, False)
            x = self.attention_layer(f'scene_encoder.s2s_attn_layers.{i}', x, x, pe_s, e_s, False)
        self._dbg_enc = dict(e_a=e_a, e_s=e_s)
        return dict(obs_mask=obs_mask, map_mask=map_mask, scene_batch_idx=scene_b, scene_type=scene_type,
                    scene_pos=pos, scene_ori=ori, scene_tokens=x,
                    max_map_num=map_emb.shape[1], max_agent_num=obs_emb.shape[1])

    def encode_prompt(self, batch):
        """traj_sam.py:79-101 + prompt_encoder/base.py:37-50."""
        p = dict(batch.extras['prompt']['motion_pred'])
        p['prompt_emd'] = self.mlp('prompt_encoder.motion_pred.state_encoder', p['prompt'].to(self.dtype), 2,
                                   ret_before_act=True)
        return p

    def generate_policy(self, batch, scene, prompt):
        """traj_sam.py:118-142 + decoder/sym_coord.py:63-140 (+ goal condition, condition_transformer/*)."""
        mask = prompt['prompt_mask']
        B, N = mask.shape
        pb = torch.arange(B).unsqueeze(1).repeat(1, N).view(-1)[mask.view(-1)]
        x_p = prompt['prompt_emd'].view(-1, 128)[mask.view(-1)]
        p_pos = prompt['position'].to(self.dtype).view(-1, 2)[mask.view(-1)]
        p_ori = prompt['heading'].to(self.dtype).view(-1, 1)[mask.view(-1)]
        e_pp = graph.radius_graph(p_pos, r=300, batch=pb, max_num_neighbors=512)
        pe_pp = self.rel_pe(e_pp, p_ori, p_pos, p_ori, p_pos)
        e_ps = graph.radius(x=scene['scene_pos'], y=p_pos, r=300, batch_x=scene['scene_batch_idx'], batch_y=pb,
                            max_num_neighbors=512)
        e_sp = e_ps[[1, 0]]
        pe_sp = self.rel_pe(e_sp, p_ori, p_pos, scene['scene_ori'], scene['scene_pos'])
        x_s = scene['scene_tokens']
        for i in range(6):
            x_p = self.attention_layer(f'decoder.p2p_attn_layers.{i}', x_p, x_p, pe_pp, e_pp, False)
            x_p = self.attention_layer(f'decoder.s2p_attn_layers.{i}', x_s, x_p, pe_sp, e_sp, True)
        emd = torch.zeros_like(prompt['prompt_emd'])
        emd[mask] = x_p
        self._dbg_gen = dict(e_pp=e_pp, e_sp=e_sp)
        out = dict(emd=emd, agent_type=prompt['agent_type'])
        cond = batch.extras['condition']
        if self.goal_condition:
            out['emd'] = self.condition_attn(cond, emd, mask, prompt['position'].to(self.dtype),
                                                  prompt['heading'].to(self.dtype))
        return out

    def encode_conditions(self, cond):
        """condition_transformer/base.py:38-48 + condition_encoders.py: one {'emd','mask','prompt_idx'} dict per
        condition type, in encoder (= PROMPT.CONDITION.TYPES) order; v_action_tag expands into one entry per used tag
        that occurs in the batch (condition_encoders.py:107-143)."""
        ct = 'condition_transformers.policy_decoder.condition_encoders'
        out = {}
        for t in self.cond_types:
            if t not in cond.keys() or cond[t]['input'].shape[1] == 0:
                continue
            c = cond[t]
            if t == 'goal':                                              # condition_encoders.py:21-51
                gi = c['input'].to(self.dtype)
                emd = self.mlp(f'{ct}.goal.goal_encoder', gi[..., :2], 2, ret_before_act=True, without_norm=True)
                out['goal'] = dict(emd=emd + fourier_fix(gi[..., 2:], 128), mask=c['mask'], prompt_idx=c['prompt_idx'])
            elif t == 'drag_point':                                      # condition_encoders.py:162-191
                pts = c['input'].to(self.dtype)
                emd = self.pointnet(f'{ct}.drag_point.pointnet_encoder', pts, ~(pts.isnan().any(-1)), 1, 2)
                out['drag_point'] = dict(emd=emd, mask=c['mask'], prompt_idx=c['prompt_idx'])
            elif t == 'v_action_tag':                                    # condition_encoders.py:76-145
                from prosim_b200.weights import V_ACTION_TAGS, V_ACTION_TAG_ID
                inp = c['input']
                B, C = inp.shape[:2]
                bi = torch.arange(B)[:, None].expand(-1, C)
                ci = torch.arange(C)[None, :].expand(B, -1)
                for tag in V_ACTION_TAGS:
                    sel = inp[..., 0] == V_ACTION_TAG_ID[tag]
                    if int(sel.sum()) == 0:
                        continue
                    cnt = sel.sum(1)
                    T = int(cnt.max())
                    tmask = torch.arange(T)[None, :].expand(B, -1) < cnt[:, None]
                    emd = torch.zeros(B, T, 128, dtype=self.dtype)
                    vmask = torch.zeros(B, T, dtype=torch.bool)
                    pidx = -torch.ones(B, T, 1, dtype=torch.long)
                    tb, tc = bi[sel], ci[sel]
                    vmask[tmask] = c['mask'][tb, tc]
                    pidx[tmask] = c['prompt_idx'][tb, tc]
                    e = self.w[f'{ct}.v_action_tag.tag_encoder.{tag}'][None, :].expand(int(sel.sum()), -1)
                    e = e + fourier_fix(inp[tb, tc, 1:3], 64).to(self.dtype)
                    emd[tmask] = e
                    out[tag] = dict(emd=emd, mask=vmask, prompt_idx=pidx)
            else:
                raise NotImplementedError(t)
        return out

    def condition_attn(self, cond, emd, mask, position, heading):
        """condition_transformer/base.py:38-60, condition_attns.py:114-228.  A unary condition on agent n is a self
        edge n->n; its attribute is the MEAN over the condition types present on that agent (COND_POOL_FUNC 'mean',
        condition_attns.py:186-189) plus the rel-PE of a zero offset.  The 3-layer GNN output is added to every valid
        prompt row (condition_attns.py:226)."""
        ct = 'condition_transformers.policy_decoder'
        emds = self.encode_conditions(cond)
        if len(emds) == 0:
            return emd
        B, N = mask.shape
        M = len(emds)
        edge_attr = torch.zeros(B, N, M, 128, dtype=self.dtype)          # the [B,N,N,M,D] matrix is diagonal in (N,N)
        edge_mask = torch.zeros(B, N, M, dtype=torch.bool)
        for m, d in enumerate(emds.values()):
            cmask = d['mask']
            C = d['emd'].shape[1]
            bidx = torch.arange(B).unsqueeze(-1).expand(B, C)[cmask]
            nidx = d['prompt_idx'][..., 0][cmask]
            edge_attr[bidx, nidx, m] = d['emd'][cmask]
            edge_mask[bidx, nidx, m] = True
        edge_attr = edge_attr.sum(dim=-2) / edge_mask.sum(dim=-1).clamp(min=1)[..., None]
        edge_mask = edge_mask.any(dim=-1)
        node_idx = -torch.ones(B, N, dtype=torch.long)
        node_idx[mask] = torch.arange(int(mask.sum()))
        ve = edge_mask.nonzero()
        e = torch.stack([node_idx[ve[:, 0], ve[:, 1]], node_idx[ve[:, 0], ve[:, 1]]], dim=0)
        r = edge_attr[edge_mask] + self.rel_pe(e, heading[mask], position[mask], heading[mask], position[mask])
        self._dbg_cond = dict(attr=edge_attr, has=edge_mask)
        x_p = emd[mask]
        for i in range(3):
            x_p = self.attention_layer(f'{ct}.condition_attn.attn_layers.{i}', x_p, x_p, r, e, False)
        emd = emd.clone()
        emd[mask] += x_p
        return emd

    # ------------------------------------------------------------------------- rollout state
    @staticmethod
    def _slot_maps(policy_ids, obs_ids):
        b, n, o = [], [], []
        for bi, ids in enumerate(policy_ids):
            lut = {a: i for i, a in enumerate(obs_ids[bi])}
            for ni, a in enumerate(ids):
                b.append(bi), n.append(ni), o.append(lut[a])
        return b, n, o

    def init_agent_trajs(self, policy_ids, batch):
        """traj_sam.py:597-633."""
        obs = batch.extras['init_obs']
        B, N = len(policy_ids), max(len(x) for x in policy_ids)
        b, n, o = self._slot_maps(policy_ids, obs['agent_ids'])
        st = dict(traj=torch.zeros(B, N, HIST, 4, dtype=self.dtype), vel=torch.zeros(B, N, HIST, 2, dtype=self.dtype),
                  init_pos=torch.zeros(B, N, 2, dtype=self.dtype), init_heading=torch.zeros(B, N, 1, dtype=self.dtype),
                  last_step=HIST)
        inp = obs['input'].to(self.dtype)
        st['traj'][b, n] = torch.nan_to_num(inp[b, o, :, :4], nan=0.0)
        st['vel'][b, n] = torch.nan_to_num(inp[b, o, :, 4:6], nan=0.0)
        st['init_pos'][b, n] = obs['position'].to(self.dtype)[b, o]
        st['init_heading'][b, n] = obs['heading'].to(self.dtype)[b, o, None]
        return st

    def step_env(self, scene, st, batch, policy_ids, t, all_t):
        """traj_sam.py:205-274 (+ _update_scene_emb :541-550, attn_fusion.py:205-251).  Note the reference
        quirk kept here: position = init_pos + traj_xy WITHOUT rotating by init_heading (:213)."""
        tidx = st['last_step']
        pos = st['init_pos'] + st['traj'][..., tidx - 1, :2]
        theta = torch.arctan2(st['traj'][..., tidx - 1, 2], st['traj'][..., tidx - 1, 3])
        a_pos = dict(position=pos, heading=wrap_angle(theta[:, :, None] + st['init_heading']))
        ti = all_t.index(t)
        if ti == 0:
            return scene, a_pos
        fut = batch.extras['fut_obs'][t]
        b, n, o = self._slot_maps(policy_ids, fut['agent_ids'])
        abs_traj = st['traj'][b, n, tidx - HIST - 2:tidx]
        th = torch.atan2(abs_traj[..., 2], abs_traj[..., 3])
        xy = rotate2d(abs_traj[..., :2] - abs_traj[..., -1:, :2], -th[..., -1:])
        dth = wrap_angle(th - th[..., -1:])
        rel = torch.cat([xy, torch.sin(dth)[..., None], torch.cos(dth)[..., None]], dim=-1)
        vel = rotate2d(st['vel'][b, n, tidx - HIST - 1:tidx], -th[..., -1:])
        acc = torch.diff(vel, dim=1) / DT
        vel_acc = torch.cat([vel[:, 1:, :], acc], dim=-1)
        inp = fut['input']
        inp[b, o, :HIST, :4] = rel[:, -HIST:].to(inp.dtype)
        inp[b, o, :HIST, 4:8] = vel_acc.to(inp.dtype)
        fut['position'][b, o] = a_pos['position'][b, n].to(inp.dtype)
        fut['heading'][b, o] = a_pos['heading'][b, n].squeeze(-1).to(inp.dtype)
        fut['mask'][b, o, :HIST] = True
        obs_emb, obs_mask = self.obs_encoder(fut)
        if self.obs_fusion == 'mlp':
            # attn_fusion.py:177-203: agents present in the old AND the new observation get MLP([old token | new token])
            old_ids = (batch.extras['fut_obs'][all_t[ti - 1]] if ti > 1 else batch.extras['init_obs'])['agent_ids']
            bi_, new_o, old_o = [], [], []
            for bidx in range(obs_emb.shape[0]):
                lut = {a: i for i, a in enumerate(old_ids[bidx])}
                for new_oidx, agent_id in enumerate(fut['agent_ids'][bidx]):
                    if agent_id in lut:
                        bi_.append(bidx), new_o.append(new_oidx), old_o.append(lut[agent_id])
            old_mask = scene['obs_mask']
            old_emb = torch.zeros(old_mask.shape[:2] + (128,), dtype=self.dtype)
            old_emb[old_mask] = scene['scene_tokens'][scene['scene_type'] == 1]
            fused = self.mlp('scene_encoder.obs_update_mlp', torch.cat([old_emb[bi_, old_o], obs_emb[bi_, new_o]], dim=-1), 2,
                             ret_before_act=True)
            obs_emb[bi_, new_o] = fused
        mt = scene['scene_type'] == 0
        B = obs_emb.shape[0]
        map_b = scene['scene_batch_idx'][mt]
        obs_b = torch.arange(B).unsqueeze(1).repeat(1, obs_emb.shape[1]).view(-1)[obs_mask.view(-1)]
        new = dict(scene)
        new['scene_batch_idx'] = torch.cat([map_b, obs_b])
        new['scene_type'] = torch.cat([torch.zeros_like(map_b), torch.ones_like(obs_b)])
        new['scene_tokens'] = torch.cat([scene['scene_tokens'][mt], obs_emb.view(-1, 128)[obs_mask.view(-1)]])
        new['scene_pos'] = torch.cat([scene['scene_pos'][mt], fut['position'].to(self.dtype).view(-1, 2)[obs_mask.view(-1)]])
        new['scene_ori'] = torch.cat([scene['scene_ori'][mt], fut['heading'].to(self.dtype).view(-1, 1)[obs_mask.view(-1)]])
        new['obs_mask'] = obs_mask
        new['max_agent_num'] = obs_emb.shape[1]
        if self.attn_update:
            new = self.update_scene_emb_attn(new)
        return new, a_pos

    def update_scene_emb_attn(self, scene):
        """attn_fusion.py:136-175 (OBS_UPDATE.ATTN_UPDATE): after the agent tokens were replaced, redo the encoder's agent
        self-attention and map -> agent attention on radius graphs (100 m / 50 m, 32 neighbours), agent rows only."""
        mt, at = scene['scene_type'] == 0, scene['scene_type'] == 1
        m_pos, m_ori, a_pos, a_ori = scene['scene_pos'][mt], scene['scene_ori'][mt], scene['scene_pos'][at], scene['scene_ori'][at]
        mb, ab = scene['scene_batch_idx'][mt], scene['scene_batch_idx'][at]
        e_a = graph.radius_graph(a_pos, r=100, batch=ab, loop=False, max_num_neighbors=32)
        e_am = graph.radius(x=m_pos, y=a_pos, r=50, batch_x=mb, batch_y=ab, max_num_neighbors=32)
        e_ma = e_am[[1, 0]]
        pe_a = self.rel_pe(e_a, a_ori, a_pos, a_ori, a_pos)
        pe_ma = self.rel_pe(e_ma, a_ori, a_pos, m_ori, m_pos)
        x_a, x_m = scene['scene_tokens'][at], scene['scene_tokens'][mt]
        for i in range(6):
            x_a = self.attention_layer(f'scene_encoder.a2a_attn_layers.{i}', x_a, x_a, pe_a, e_a, False)
            # a non-bipartite module called with (x_src, x_dst): both pre-norms are the one shared LayerNorm (attention_layer.py:48-49)
            x_a = self.attention_layer(f'scene_encoder.s2s_attn_layers.{i}', x_m, x_a, pe_ma, e_ma, True)
        scene = dict(scene)
        tok = scene['scene_tokens'].clone()
        tok[at] = x_a
        scene['scene_tokens'] = tok
        self._dbg_upd = dict(e_a=e_a, e_ma=e_ma)
        return scene

    def policy_tick(self, policy, scene, policy_ids, a_pos, t):
        """traj_sam.py:178-202,441-525 (row gather) + policy/act_decoder.py:239-279 (attn_fuse) +
        act_decoder.py:78-135 (_compute_traj).  The dense [B,S,128] round trip of _scene_emd_to_batch /
        _process_scene_token is an identity on the flat valid tokens and is skipped."""
        b, n = [], []
        for bi, ids in enumerate(policy_ids):
            b += [bi] * len(ids)
            n += list(range(len(ids)))
        names = [f'{bi}-{policy_ids[bi][ni]}-{t}' for bi, ni in zip(b, n)]
        x_p = policy['emd'][b, n]
        a_type = policy['agent_type'][b, n]
        p_pos, p_ori = a_pos['position'][b, n], a_pos['heading'][b, n]
        pb = torch.tensor(b, dtype=torch.long)
        at = scene['scene_type'] == 1
        mt = ~at
        sb = scene['scene_batch_idx']
        e_pa = graph.radius(x=scene['scene_pos'][at], y=p_pos, r=100, batch_x=sb[at], batch_y=pb, max_num_neighbors=768)
        e_ap = e_pa[[1, 0]]
        pe_ap = self.rel_pe(e_ap, p_ori, p_pos, scene['scene_ori'][at], scene['scene_pos'][at])
        e_pm = graph.radius(x=scene['scene_pos'][mt], y=p_pos, r=50, batch_x=sb[mt], batch_y=pb, max_num_neighbors=768)
        e_mp = e_pm[[1, 0]]
        pe_mp = self.rel_pe(e_mp, p_ori, p_pos, scene['scene_ori'][mt], scene['scene_pos'][mt])
        x_a, x_m = scene['scene_tokens'][at], scene['scene_tokens'][mt]
        emd_rows = x_p
        for i in range(6):
            x_p = self.attention_layer(f'policy.act_decoder.a2p_attn_layers.{i}', x_a, x_p, pe_ap, e_ap, True)
            x_p = self.attention_layer(f'policy.act_decoder.m2p_attn_layers.{i}', x_m, x_p, pe_mp, e_mp, True)
        out = self.policy_head(x_p, a_type, emd_rows)
        out['pair_names'] = names
        if self.trace is not None:
            self.trace.append(dict(t=t, e_ap=e_ap, e_mp=e_mp, x_a=x_a, x_m=x_m, p_pos=p_pos, p_ori=p_ori,
                                   a_pos=scene['scene_pos'][at], a_ori=scene['scene_ori'][at],
                                   m_pos=scene['scene_pos'][mt], m_ori=scene['scene_ori'][mt],
                                   fuse=x_p, motion_pred=out['motion_pred']))
        return out

    def policy_head(self, feat, agent_type, emd_rows):
        """act_decoder.py:78-135 with PRED_MODE=anchor, K=1: anchor = Embedding[type-1]; CG_stacked(3)
        (layers/mlp.py:207-241; max over K=1 is the identity); motion_head; cumsum / wrap."""
        pa = 'policy.act_decoder'
        anchor = self.w[f'{pa}.motion_anchors.weight'][agent_type - 1][:, None, :]
        ctx = feat

        def cg(i, inp, c):
            y = F.relu(self.ln(f'{pa}.CG_decode.CGs.{i}.MLP.1', self.lin(f'{pa}.CG_decode.CGs.{i}.MLP.0', inp)))
            y = y * c.unsqueeze(1)
            return y, torch.max(y, dim=1)[0]

        inp_, ctx_ = cg(0, anchor, ctx)
        for i in range(1, 3):
            inp, c = cg(i, inp_, ctx_)
            inp_ = (inp_ * i + inp) / (i + 1)
            ctx_ = (ctx_ * i + c) / (i + 1)
        motion = self.mlp(f'{pa}.motion_head', inp_, 3, ret_before_act=True).view(feat.shape[0], 1, STEP, 5)
        if self.noise_std > 0:
            motion[..., :2] += self.noise_fn(motion[..., :2]) * self.noise_std
        xy = motion[..., :2].cumsum(dim=-2)
        hd = wrap_angle(motion[..., 2:3].cumsum(dim=-2))
        pred = torch.cat([xy, hd, motion[..., 3:]], dim=-1)
        return dict(motion_pred=pred, motion_prob=torch.ones_like(motion[..., 0, 0]),
                    reconst_pred=self.mlp(f'{pa}.pred_mlp', emd_rows, 3, ret_before_act=True))

    def step_agent_traj(self, st, out, policy_ids):
        """traj_sam.py:276-349 with TOP_K=1 (the randint draw is degenerate)."""
        b, n = [], []
        for bi, ids in enumerate(policy_ids):
            b += [bi] * len(ids)
            n += list(range(len(ids)))
        tidx = st['last_step']
        if self.noise_std > 0:     # traj_sam.py:313: the degenerate mode draw still advances the generator
            torch.randint(0, 1, (len(b),))
        cur = st['traj'][b, n, :tidx]
        pred = out['motion_pred'][:, 0, :STEP]
        last = torch.arctan2(cur[:, -1, 2], cur[:, -1, 3])[:, None]
        xy = rotate2d(pred[:, :, :2], last) + cur[:, -1:, :2]
        th = wrap_angle(last + pred[:, :, 2])
        fut = torch.cat([xy, torch.sin(th)[..., None], torch.cos(th)[..., None]], dim=-1)
        B, N = st['traj'].shape[:2]
        new_traj = torch.zeros(B, N, STEP, 4, dtype=self.dtype)
        new_traj[b, n] = fut
        new_vel = torch.zeros(B, N, STEP, 2, dtype=self.dtype)
        new_vel[b, n] = rotate2d(pred[..., 3:5], last)
        st['traj'] = torch.cat([st['traj'], new_traj], dim=2)
        st['vel'] = torch.cat([st['vel'], new_vel], dim=2)
        st['last_step'] = tidx + STEP
        return st

    # ------------------------------------------------------------------------- entry point
    @torch.no_grad()
    def forward(self, batch, mode='val'):
        """traj_sam.py:59-71, 103-116, 144-175, 562-595."""
        scene = self.encode_scene(batch)
        prompt = self.encode_prompt(batch)
        policy = self.generate_policy(batch, scene, prompt)
        policy_ids = batch.extras['prompt']['motion_pred']['agent_ids']
        all_t = sorted(batch.extras['all_t_indices'].cpu().numpy().tolist())
        st = self.init_agent_trajs(policy_ids, batch)
        outs = []
        for t in all_t:
            if self.trace is not None:
                self.trace_states.append(dict(t=t, traj=st['traj'].clone(), vel=st['vel'].clone(), last_step=st['last_step']))
            scene, a_pos = self.step_env(scene, st, batch, policy_ids, t, all_t)
            out = self.policy_tick(policy, scene, policy_ids, a_pos, t)
            if self.faithful_bookkeeping:
                self._quadratic_name_matching(out['pair_names'], policy_ids, t)
            st = self.step_agent_traj(st, out, policy_ids)
            outs.append(out)
        res = {k: torch.cat([o[k] for o in outs], dim=0) for k in ('motion_pred', 'motion_prob', 'reconst_pred')}
        res['pair_names'] = [x for o in outs for x in o['pair_names']]
        res['rollout_trajs'] = {}
        for bi, ids in enumerate(policy_ids):
            for ni, aid in enumerate(ids):
                res['rollout_trajs'][f'{bi}-{aid}'] = dict(
                    traj=st['traj'][bi, ni, HIST:], vel=st['vel'][bi, ni, HIST:],
                    init_pos=st['init_pos'][bi, ni], init_heading=st['init_heading'][bi, ni])
        self.final_state = st
        return {'motion_pred': res}

    @staticmethod
    def _quadratic_name_matching(pair_names, policy_ids, t):
        """The reference's O(P^2) per-tick string search (traj_sam.py:289-298; temporal_ar.py:22-35),
        kept behind a flag so the CPU baseline can include it."""
        for bi, ids in enumerate(policy_ids):
            for aid in ids:
                name = f'{bi}-{aid}-{t}'
                if name in pair_names:
                    pair_names.index(name)
        agent_names = ['-'.join(nm.split('-')[:-1]) for nm in pair_names]
        uniq = sorted(set(agent_names))
        [uniq.index(nm) for nm in agent_names]


def rollout_trajs_in_world(result, tf):
    """rollout/gpu_utils.py:230-281 (obtain_rollout_trajs_in_world) + rollout/utils.py:347-392.
    result: the 'motion_pred' dict of a forward; tf: [3, 3] centre->world.  Returns ([P, steps, 3], names)."""
    names = list(result['rollout_trajs'].keys())
    trajs = torch.stack([result['rollout_trajs'][n]['traj'] for n in names])
    init_pos = torch.stack([result['rollout_trajs'][n]['init_pos'] for n in names])
    init_heads = torch.stack([result['rollout_trajs'][n]['init_heading'] for n in names])
    xy_c = rotate2d(trajs[:, :, :2], init_heads) + init_pos[:, None]
    hs = torch.arctan2(trajs[:, :, 2], trajs[:, :, 3])
    hs_c = wrap_angle(hs + init_heads)
    mat = torch.transpose(tf, -1, -2)
    xy_w = (xy_c[..., None, :] @ mat[None, None, :2, :2]).squeeze(-2) + mat[None, None, -1:, :2].squeeze(-2)
    rot = torch.arctan2(tf[1, 0], tf[0, 0])
    hs_w = (hs_c + rot + math.pi) % (2 * math.pi) - math.pi
    return torch.cat([xy_w, hs_w[:, :, None]], dim=-1), names
