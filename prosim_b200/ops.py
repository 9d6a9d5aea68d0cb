"""Thin tensor-level wrappers over the C ABI (one per entry point of include/prosim_b200.h).

Plumbing only: shape/dtype/device checks, output allocation with torch, pointer extraction.  Every
function enqueues on ``torch.cuda.current_stream()`` and returns device tensors; nothing here computes.
"""
import ctypes

import torch

from . import lib
from .lib import Graph, StackSide, ptr

D = 128
HEADS = 8
ATTN_WS_PER_DST = 2 * (D + HEADS * D + D + D) + HEADS * D + D + 2 * D


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise lib.ProSimLibError(f'{name} must be a CUDA tensor (no CPU fallback on this path)')
    if t.device.index != torch.cuda.current_device():
        raise lib.ProSimLibError(f'{name} lives on {t.device} but the current device is cuda:{torch.cuda.current_device()} '
                                 '(the kernels launch on the current device\'s stream)')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be contiguous')


def as_u8(mask):
    """torch.bool and torch.uint8 share a 1-byte layout: reinterpret without copying."""
    return mask.view(torch.uint8) if mask.dtype == torch.bool else mask


class EdgeList:
    """Fixed-stride neighbour lists (+ the per-edge normalised relative PE once computed)."""

    def __init__(self, nbr, deg, stride, max_deg, z=None, zd=128, warps_per_row=0):
        self.nbr, self.deg, self.stride, self.max_deg, self.z = nbr, deg, int(stride), int(max_deg), z
        self.zd, self.warps_per_row = int(zd), int(warps_per_row)

    @property
    def n_dst(self):
        return self.deg.shape[0]

    def c_struct(self):
        return Graph(ptr(self.z), ptr(self.nbr), ptr(self.deg), self.stride, self.max_deg, self.zd, self.warps_per_row)

    def to_edge_index(self):
        """[2, E] (row0 = source, row1 = destination) on the host, for bit-exact comparison with the oracle."""
        deg = self.deg.cpu().long()
        nbr = self.nbr.view(self.n_dst, self.stride).cpu().long()
        j = torch.arange(self.stride)[None, :] < deg[:, None]
        dst = torch.arange(self.n_dst)[:, None].expand_as(nbr)[j]
        return torch.stack([nbr[j], dst], dim=0)


def pointnet(kind, x, mask, rows, w_arena, w_off, out=None, tc_off=None):
    """kind 0 obs / 1 map / 2, 3 drag points (16 / 8 per agent; mask None = validity from NaN).
    tc_off: offset of the tensor-core operand block (weights.pack_pointnet_tc); None = fp32 FFMA kernel."""
    _chk(x, torch.float32, 'x'), _chk(rows, torch.int32, 'rows')
    if mask is not None or kind < 2:
        mask = as_u8(mask)
        _chk(mask, torch.uint8, 'mask')
    n = rows.shape[0]
    if out is None:
        out = torch.empty(n, D, device=x.device, dtype=torch.float32)
    lib.call('prosim_pointnet_fwd', kind, ptr(x), ptr(mask), ptr(rows), n, ptr(w_arena, w_off),
             ptr(w_arena, tc_off) if tc_off is not None else None, ptr(out), _stream())
    return out


def radius_edges(qpos, qscene, spos, seg, r, cap, stride, drop_self=False, nbr=None, deg=None):
    for t, n in ((qpos, 'qpos'), (spos, 'spos')):
        _chk(t, torch.float32, n)
    _chk(qscene, torch.int32, 'qscene'), _chk(seg, torch.int32, 'seg')
    nq = qpos.shape[0]
    stride = max(int(stride), 1)
    if nbr is None:
        nbr = torch.empty(nq * stride, device=qpos.device, dtype=torch.int32)
        deg = torch.empty(nq, device=qpos.device, dtype=torch.int32)
    lib.call('prosim_build_radius_edges', ptr(qpos), ptr(qscene), nq, ptr(spos), ptr(seg), float(r), int(cap),
             int(bool(drop_self)), ptr(nbr), ptr(deg), stride, _stream())
    return EdgeList(nbr, deg, stride, stride)


def knn_edges(qpos, qscene, spos, seg, k, nmax, stride):
    for t, n in ((qpos, 'qpos'), (spos, 'spos')):
        _chk(t, torch.float32, n)
    _chk(qscene, torch.int32, 'qscene'), _chk(seg, torch.int32, 'seg')
    nq = qpos.shape[0]
    stride = max(int(stride), 1)
    nbr = torch.empty(nq * stride, device=qpos.device, dtype=torch.int32)
    deg = torch.empty(nq, device=qpos.device, dtype=torch.int32)
    lib.call('prosim_build_knn_edges', ptr(qpos), ptr(qscene), nq, ptr(spos), ptr(seg), int(k), int(nmax), ptr(nbr),
             ptr(deg), stride, _stream())
    return EdgeList(nbr, deg, stride, stride)


def edge_pe(edges, dpos, dori, spos, sori, dim_t16, extra=None, z=None):
    for t, n in ((dpos, 'dpos'), (dori, 'dori'), (spos, 'spos'), (sori, 'sori'), (dim_t16, 'dim_t16'), (extra, 'extra')):
        _chk(t, torch.float32, n)
    n_dst = edges.n_dst
    zd = 128 if extra is not None else 96     # pure rel-PE rows carry 96 distinct features (phi is embedded twice)
    if z is None:
        z = torch.empty(n_dst * edges.stride, zd, device=dpos.device, dtype=torch.float32)
    lib.call('prosim_edge_pe', ptr(dpos), ptr(dori), n_dst, ptr(spos), ptr(sori), ptr(edges.nbr), ptr(edges.deg),
             edges.stride, ptr(dim_t16), ptr(extra), zd, ptr(z), _stream())
    edges.z, edges.zd = z, zd
    return edges


def attn_kv(x_src, w_arena, w_off, n_layers, layer_floats, kv=None):
    _chk(x_src, torch.float32, 'x_src')
    n = x_src.shape[0]
    if kv is None:
        kv = torch.empty(n_layers, n, 2 * D, device=x_src.device, dtype=torch.float32)
    lib.call('prosim_attn_kv', ptr(x_src), n, ptr(w_arena, w_off), layer_floats, n_layers, ptr(kv), n * 2 * D, _stream())
    return kv


def attn_workspace(n_dst, n_src, device, max_stride=768):
    n = lib.load().prosim_attn_workspace_floats(int(n_dst), int(n_src), int(max_stride))
    return torch.empty(n, device=device, dtype=torch.float32)


def attn_layer(x_src, x_dst, edges, w_arena, w_off, out=None, workspace=None):
    _chk(x_src, torch.float32, 'x_src'), _chk(x_dst, torch.float32, 'x_dst')
    n_src, n_dst = x_src.shape[0], x_dst.shape[0]
    if workspace is None:
        workspace = attn_workspace(n_dst, n_src, x_dst.device, edges.stride)
    if out is None:
        out = torch.empty_like(x_dst)
    g = edges.c_struct()
    lib.call('prosim_attn_layer_fwd', ptr(x_src), n_src, ptr(x_dst), n_dst, ctypes.byref(g), ptr(w_arena, w_off),
             ptr(workspace), workspace.numel(), ptr(out), _stream())
    return out


def stack_side(w_arena, w_off, edges, kv=None):
    side = StackSide(ptr(w_arena, w_off), ptr(kv), (kv.shape[1] * kv.shape[2]) if kv is not None else 0, edges.c_struct())
    side._keepalive = (w_arena, kv, edges)   # the struct only holds raw addresses
    return side


def attn_stack(x, n_layers, side_a, side_b=None, out=None, workspace=None):
    _chk(x, torch.float32, 'x')
    n = x.shape[0]
    if workspace is None:
        workspace = attn_workspace(n, n, x.device, max(side_a.graph.stride, side_b.graph.stride if side_b is not None else 1))
    if out is None:
        out = torch.empty_like(x)
    lib.call('prosim_attn_stack_fwd', ptr(x), n, int(n_layers), ctypes.byref(side_a),
             ctypes.byref(side_b) if side_b is not None else None, ptr(workspace), workspace.numel(), ptr(out), _stream())
    return out


def policy_head(feat, agent_type, w_arena, w_off, motion_pred=None, noise=None, noise_std=0.0):
    """noise: optional standard-normal draws [P, 1, 10, 2] (RANDOM_NOISE_STD > 0, act_decoder.py:113-115)."""
    _chk(feat, torch.float32, 'feat'), _chk(agent_type, torch.int32, 'agent_type'), _chk(noise, torch.float32, 'noise')
    P = feat.shape[0]
    if noise is not None and noise.numel() != P * 20:
        raise ValueError('noise must hold [P, 1, 10, 2] values')
    if motion_pred is None:
        motion_pred = torch.empty(P, 1, 10, 5, device=feat.device, dtype=torch.float32)
    lib.call('prosim_policy_head_fwd', ptr(feat), ptr(agent_type), P, ptr(w_arena, w_off), ptr(noise), float(noise_std),
             ptr(motion_pred), _stream())
    return motion_pred


def tick_cfg(P, n_agent, n_map, max_a, max_m, max_neigh, n_layers):
    return lib.Cfg(int(P), int(n_agent), int(n_map), int(max_a), int(max_m), int(max_neigh), int(n_layers))


def tick_workspace_bytes(cfg):
    return int(lib.load().prosim_workspace_bytes(ctypes.byref(cfg)))


def policy_tick(cfg, emd, agent_type, p_scene, p_pos, p_ori, x_a, a_pos, a_ori, seg_a, m_pos, m_ori, seg_m, kv_m, w_arena,
                off_a2p, off_m2p, off_head, dim_t16, agent_radius, map_radius, workspace, fuse, motion_pred=None, noise=None,
                noise_std=0.0):
    """One policy tick through the single C entry point (include/prosim_b200.h: prosim_policy_tick); workspace: uint8 tensor
    of at least tick_workspace_bytes(cfg) bytes.  Returns motion_pred [P, 1, 10, 5]."""
    for t, n in ((emd, 'emd'), (p_pos, 'p_pos'), (p_ori, 'p_ori'), (x_a, 'x_a'), (a_pos, 'a_pos'), (a_ori, 'a_ori'), (m_pos, 'm_pos'),
                 (m_ori, 'm_ori'), (kv_m, 'kv_m'), (fuse, 'fuse'), (noise, 'noise'), (dim_t16, 'dim_t16')):
        _chk(t, torch.float32, n)
    for t, n in ((agent_type, 'agent_type'), (p_scene, 'p_scene'), (seg_a, 'seg_agent'), (seg_m, 'seg_map')):
        _chk(t, torch.int32, n)
    _chk(workspace, torch.uint8, 'workspace')
    P = cfg.n_policy_rows
    if emd.shape[0] != P or x_a.shape[0] != cfg.n_agent_tokens or m_pos.shape[0] != cfg.n_map_tokens:
        raise ValueError('policy_tick: tensor sizes do not match the tick configuration')
    if noise is not None and noise.numel() != P * 20:
        raise ValueError('noise must hold [P, 1, 10, 2] values')
    if motion_pred is None:
        motion_pred = torch.empty(P, 1, 10, 5, device=emd.device, dtype=torch.float32)
    t = lib.Tick()
    t.cfg = cfg
    for name, val in (('emd', emd), ('agent_type', agent_type), ('p_scene', p_scene), ('p_pos', p_pos), ('p_ori', p_ori),
                      ('x_agent', x_a), ('agent_pos', a_pos), ('agent_ori', a_ori), ('seg_agent', seg_a), ('map_pos', m_pos),
                      ('map_ori', m_ori), ('seg_map', seg_m), ('kv_map', kv_m), ('dim_t16', dim_t16), ('noise', noise),
                      ('fuse', fuse), ('motion_pred', motion_pred)):
        setattr(t, name, ptr(val))
    t.w_a2p, t.w_m2p, t.w_head = ptr(w_arena, off_a2p), ptr(w_arena, off_m2p), ptr(w_arena, off_head)
    t.agent_radius, t.map_radius, t.noise_std = float(agent_radius), float(map_radius), float(noise_std)
    t.p_row = t.traj = t.vel = None
    t.T = t.tidx = 0
    lib.call('prosim_policy_tick', ctypes.byref(t), ptr(workspace), workspace.numel(), _stream())
    return motion_pred


def reconst(emd, w_arena, w_off):
    _chk(emd, torch.float32, 'emd')
    out = torch.empty(emd.shape[0], 2, device=emd.device, dtype=torch.float32)
    lib.call('prosim_reconst_fwd', ptr(emd), emd.shape[0], ptr(w_arena, w_off), ptr(out), _stream())
    return out


def mlp2(x, k0, use_ln, w_arena, w_off, tpe_col=None, dim_t128=None):
    _chk(x, torch.float32, 'x')
    n, ld = x.shape
    out = torch.empty(n, D, device=x.device, dtype=torch.float32)
    tpe = ptr(x, tpe_col) if tpe_col is not None else None
    lib.call('prosim_mlp2_fwd', ptr(x), ld, int(k0), n, int(bool(use_ln)), ptr(w_arena, w_off), tpe, ld,
             ptr(dim_t128) if tpe_col is not None else None, ptr(out), _stream())
    return out


def obs_fuse(x_old, idx_old, x_new, idx_new, w_arena, w_off):
    """x_new[idx_new] <- obs_update_mlp([x_old[idx_old] | x_new[idx_new]]) in place (attn_fusion.py:177-203)."""
    _chk(x_old, torch.float32, 'x_old'), _chk(x_new, torch.float32, 'x_new')
    _chk(idx_old, torch.int32, 'idx_old'), _chk(idx_new, torch.int32, 'idx_new')
    if idx_old.shape[0] != idx_new.shape[0]:
        raise ValueError('obs_fuse: index lists differ in length')
    lib.call('prosim_obs_fuse_fwd', ptr(x_old), ptr(idx_old), ptr(x_new), ptr(idx_new), idx_new.shape[0], ptr(w_arena, w_off), _stream())
    return x_new


def tag_embed(tags, table, dim_t64, n_tags=11):
    """tags: int64 [n, 3] (tag id, start, end) -> [n, 128] (condition_encoders.py:76-145)."""
    _chk(tags, torch.int64, 'tags'), _chk(table, torch.float32, 'table'), _chk(dim_t64, torch.float32, 'dim_t64')
    n = tags.shape[0]
    out = torch.empty(n, D, device=tags.device, dtype=torch.float32)
    lib.call('prosim_tag_embed_fwd', ptr(tags), n, int(n_tags), ptr(table), ptr(dim_t64), ptr(out), _stream())
    return out


def cond_pool(emb, slot):
    """emb [n, 128], slot int32 [P, n_slots] -> (extra [P, 128], has int32 [P]) (condition_attns.py:114-189)."""
    _chk(emb, torch.float32, 'emb'), _chk(slot, torch.int32, 'slot')
    P, ns = slot.shape
    extra = torch.empty(P, D, device=emb.device, dtype=torch.float32)
    has = torch.empty(P, device=emb.device, dtype=torch.int32)
    lib.call('prosim_cond_pool_fwd', ptr(emb), ptr(slot), P, ns, ptr(extra), ptr(has), _stream())
    return extra, has


def init_traj(obs_in, obs_pos, obs_head, p_slot, p_row, T, traj, vel, init_pos, init_heading):
    for t, n in ((obs_in, 'obs input'), (obs_pos, 'obs position'), (obs_head, 'obs heading'), (traj, 'traj'), (vel, 'vel'),
                 (init_pos, 'init_pos'), (init_heading, 'init_heading')):
        _chk(t, torch.float32, n)
    _chk(p_slot, torch.int32, 'p_slot'), _chk(p_row, torch.int32, 'p_row')
    lib.call('prosim_init_traj', ptr(obs_in), ptr(obs_pos), ptr(obs_head), ptr(p_slot), ptr(p_row), p_row.shape[0], int(T),
             ptr(traj), ptr(vel), ptr(init_pos), ptr(init_heading), _stream())


def step_env(traj, vel, init_pos, init_heading, p_row, p_slot, T, tidx, p_pos, p_ori, fut=None):
    """fut: None on the first tick, else (input, mask, position, heading) tensors of fut_obs[t] (written in place)."""
    for t, n in ((traj, 'traj'), (vel, 'vel'), (init_pos, 'init_pos'), (init_heading, 'init_heading'), (p_pos, 'p_pos'),
                 (p_ori, 'p_ori')):
        _chk(t, torch.float32, n)
    _chk(p_row, torch.int32, 'p_row'), _chk(p_slot, torch.int32, 'p_slot')
    f_in = f_mask = f_pos = f_head = None
    if fut is not None:
        f_in, f_mask, f_pos, f_head = fut
        f_mask = as_u8(f_mask)
        _chk(f_mask, torch.uint8, 'fut mask')
        for t, n in ((f_in, 'fut input'), (f_pos, 'fut position'), (f_head, 'fut heading')):
            _chk(t, torch.float32, n)
    lib.call('prosim_step_env', ptr(traj), ptr(vel), ptr(init_pos), ptr(init_heading), ptr(p_row), ptr(p_slot),
             p_row.shape[0], int(T), int(tidx), ptr(p_pos), ptr(p_ori), ptr(f_in), ptr(f_mask), ptr(f_pos), ptr(f_head),
             _stream())


def gather_pose(pos, head, rows, out_pos, out_ori):
    _chk(pos, torch.float32, 'pos'), _chk(head, torch.float32, 'head'), _chk(rows, torch.int32, 'rows')
    _chk(out_pos, torch.float32, 'out_pos'), _chk(out_ori, torch.float32, 'out_ori')
    lib.call('prosim_gather_pose', ptr(pos), ptr(head), ptr(rows), rows.shape[0], ptr(out_pos), ptr(out_ori), _stream())


def step_agent_traj(motion_pred, p_row, T, tidx, traj, vel):
    _chk(motion_pred, torch.float32, 'motion_pred'), _chk(traj, torch.float32, 'traj'), _chk(vel, torch.float32, 'vel')
    _chk(p_row, torch.int32, 'p_row')
    if motion_pred.numel() != p_row.shape[0] * 50:
        raise ValueError('motion_pred must hold [P, 1, 10, 5] values')
    lib.call('prosim_step_agent_traj', ptr(motion_pred), ptr(p_row), p_row.shape[0], int(T), int(tidx), ptr(traj), ptr(vel),
             _stream())


def rollout_to_world(traj, init_pos, init_heading, p_row, T, t0, steps, tf):
    """[P, steps, 3] world (x, y, heading) of the rolled-out steps; tf: [3, 3] centre->world transform on the device."""
    for t, n in ((traj, 'traj'), (init_pos, 'init_pos'), (init_heading, 'init_heading'), (tf, 'tf')):
        _chk(t, torch.float32, n)
    P = p_row.shape[0]
    out = torch.empty(P, steps, 3, device=traj.device, dtype=torch.float32)
    lib.call('prosim_rollout_to_world', ptr(traj), ptr(init_pos), ptr(init_heading), ptr(p_row), P, int(T), int(t0),
             int(steps), ptr(tf), ptr(out), _stream())
    return out
