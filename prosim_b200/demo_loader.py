"""Mini-loader for the reference's ``demo_dataset`` (trajdata cache format) -> the batch the rollout path consumes.

SURVEY.md section 8f-1: makes BASELINE configs[0] literal ("demo_dataset single scene, 16 agents, 20-step unconditional
rollout") without trajdata.  CPU-side data formatting, numpy / torch only; nothing here runs on the hot path.

What is restated from the reference (file:line in each function):
  * lane vectorisation            prosim/dataset/data_utils.py:155-257   (_get_vectorized_lanes_from_vector_map, COLLATE_MODE 'lane')
  * local map, symmetric coords   prosim/dataset/format_utils.py:153-263 (get_local_vec_map, local_map_to_sym_coord, get_center_vec_init_map)
  * agent-centred histories       prosim/dataset/format_utils.py:357-447 (get_center_obs) and :667-687 (get_future_obs, 'latest')
  * agent-status prompt           prosim/dataset/prompt_utils.py:24-150
  * tick indices, target ranking  prosim/dataset/format_utils.py:699-713, :765-769
What lives in the UN-VENDORED trajdata (not under /root/reference, not installed) and is restated from its file formats and
documented behaviour -- PARITY UNPINNED for these conventions, they can only be checked against the files the reference ships:
  * the cache files: ``agent_data_dt0.10.feather`` (agent_id, scene_ts, x, y, z, vx, vy, ax, ay, heading, length, width,
    height), ``scene_metadata_dt0.10.dill`` (a pickled trajdata Scene: agent names and AgentType values) and the
    ``maps/<map>.pb`` VectorizedMap protobuf (wire format decoded by hand: element{1 id, 2 road_lane{1 center, 2 left_edge,
    3 right_edge: polyline{1 dx_mm, 2 dy_mm, 3 dz_mm packed sint32 deltas, 4 h_rad packed double}}}, 5 shifted_origin)
  * SceneBatch conventions: the scene is centred on the ``ego`` agent's pose at the current step (``standardize_data``:
    positions translated and rotated, headings relative, velocities / accelerations rotated), agents are those with a state at
    the current step, ordered by distance to the centre, at most ``max_agents``; lanes are those with a centre point within
    sqrt(2) * map_range of the centre (``VectorMap.get_lanes_within``); a lane without traffic-light data has status -1
    (trajdata TrafficLightStatus.NO_DATA).
"""
import io
import math
import os
import pickle

import numpy as np
import torch

from .containers import BatchCondition, BatchDataDict, BatchPrompt, InputMaskData, SceneBatch

HIST = 11
STATE_COLS = ('x', 'y', 'z', 'vx', 'vy', 'ax', 'ay', 'heading')


# ----------------------------------------------------------------------------------------------- files
class _Stub:
    def __init__(self, *a, **k):
        self.args = a

    def __setstate__(self, st):
        self.__dict__.update(st if isinstance(st, dict) else {'state': st})


class _Unpickler(pickle.Unpickler):
    """Reads trajdata's pickled metadata without trajdata: unknown classes become attribute bags (enum members keep their value)."""

    def find_class(self, module, name):
        if not module.startswith('trajdata'):
            try:
                obj = super().find_class(module, name)
                if isinstance(obj, type) or callable(obj):
                    return obj
            except Exception:
                pass
        return type(name, (_Stub,), {'__module__': module})


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, i


def _fields(b):
    """Protobuf wire format: [(field, wire type, value)] of one message."""
    i, out = 0, []
    while i < len(b):
        k, i = _varint(b, i)
        f, w = k >> 3, k & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v, i = b[i:i + 8], i + 8
        elif w == 2:
            n, i = _varint(b, i)
            v, i = b[i:i + n], i + n
        elif w == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(f'unsupported wire type {w}')
        out.append((f, w, v))
    return out


def _packed_sint(b):
    i, out = 0, []
    while i < len(b):
        u, i = _varint(b, i)
        out.append((u >> 1) ^ -(u & 1))
    return np.asarray(out, dtype=np.int64)


def _polyline(msg, origin):
    d = {f: v for f, w, v in _fields(msg) if w == 2}
    dx, dy = _packed_sint(d.get(1, b'')), _packed_sint(d.get(2, b''))
    xy = np.stack([np.cumsum(dx), np.cumsum(dy)], axis=1) / 1000.0 + origin[None, :2]
    return xy


def read_vector_map(pb_path):
    """maps/<name>.pb -> list of lanes {'id', 'center' [n, 2], 'left_edge' [m, 2] | None, 'right_edge' | None} in world metres."""
    top = _fields(open(pb_path, 'rb').read())
    origin = np.zeros(3)
    for f, w, v in top:
        if f == 5 and w == 2:
            pt = {ff: np.frombuffer(vv, dtype='<f8')[0] for ff, ww, vv in _fields(v) if ww == 1}
            origin = np.array([pt.get(1, 0.0), pt.get(2, 0.0), pt.get(3, 0.0)])
    lanes = []
    for f, w, v in top:
        if f != 2 or w != 2:
            continue
        el = _fields(v)
        lane_msg = next((vv for ff, ww, vv in el if ff == 2 and ww == 2), None)      # oneof: road_lane
        if lane_msg is None:
            continue
        parts = {ff: vv for ff, ww, vv in _fields(lane_msg) if ww == 2 and ff in (1, 2, 3)}
        if 1 not in parts:
            continue
        lanes.append({'id': next((vv for ff, ww, vv in el if ff == 1), b'').decode(),
                      'center': _polyline(parts[1], origin),
                      'left_edge': _polyline(parts[2], origin) if 2 in parts else None,
                      'right_edge': _polyline(parts[3], origin) if 3 in parts else None})
    return lanes


def read_scene(scene_dir):
    """agent table + metadata of one cached scene -> (names, types [n], states [n, T, 8] (NaN = absent), extents [n, T, 2], meta)."""
    import pandas as pd
    df = pd.read_feather(os.path.join(scene_dir, 'agent_data_dt0.10.feather'))
    with open(os.path.join(scene_dir, 'scene_metadata_dt0.10.dill'), 'rb') as fh:
        meta = _Unpickler(io.BytesIO(fh.read())).load()
    T = int(meta.length_timesteps)
    names = [a.name for a in meta.agents]
    types = []
    for a in meta.agents:
        t = a.type
        types.append(int(t.args[0]) if getattr(t, 'args', None) else int(getattr(t, 'value', 0)))
    states = np.full((len(names), T, 8), np.nan, dtype=np.float64)
    extents = np.full((len(names), T, 2), np.nan, dtype=np.float64)
    idx = {n: i for i, n in enumerate(names)}
    rows = df['agent_id'].map(lambda x: idx[str(x)]).to_numpy()
    ts = df['scene_ts'].to_numpy()
    states[rows, ts] = df[list(STATE_COLS)].to_numpy(dtype=np.float64)
    extents[rows, ts] = df[['length', 'width']].to_numpy(dtype=np.float64)
    tls = pd.read_feather(os.path.join(scene_dir, 'tls_data_dt0.10.feather')) if os.path.exists(
        os.path.join(scene_dir, 'tls_data_dt0.10.feather')) else None
    return names, np.asarray(types), states, extents, {'meta': meta, 'tls': tls}


# ----------------------------------------------------------------------------------------------- frames
def _rot(x, y, th):
    """prosim/dataset/data_utils.py:82-84 (rotate)."""
    c, s = np.cos(th), np.sin(th)
    return x * c - y * s, x * s + y * c


def _to_frame(states, pos0, h0):
    """trajdata ``transform_to_frame_offset_rot`` / ``standardize_data``: positions translated and rotated into the frame at
    (pos0, h0), velocities and accelerations rotated, heading relative.  states [..., 8] -> same layout."""
    out = np.array(states, dtype=np.float64, copy=True)
    out[..., 0], out[..., 1] = _rot(states[..., 0] - pos0[..., 0], states[..., 1] - pos0[..., 1], -h0)
    out[..., 3], out[..., 4] = _rot(states[..., 3], states[..., 4], -h0)
    out[..., 5], out[..., 6] = _rot(states[..., 5], states[..., 6], -h0)
    out[..., 7] = states[..., 7] - h0
    return out


# ----------------------------------------------------------------------------------------------- map
def vectorize_lanes(lanes, centre_xy, centre_h, tls_of, map_range=200.0, center_rate=1, edge_rate=4, max_lane_points=20,
                    include=('center', 'right_edge', 'left_edge')):
    """prosim/dataset/data_utils.py:155-257 with COLLATE_MODE 'lane': every lane within sqrt(2) * map_range, each of its
    polylines subsampled, moved to the centred frame, clipped to |x|, |y| < map_range, cut into chunks of ``max_lane_points``
    points -> vectors (x0, y0, x1, y1, line type, traffic-light status), zero padded to ``max_lane_points - 1`` vectors."""
    line_type = {'center': 1.0, 'left_edge': 2.0, 'right_edge': 3.0}                 # RoadLaneType, data_utils.py:24-27
    lane_dist = math.sqrt(2.0) * map_range
    vecs = []
    for lane in lanes:
        if np.min(np.linalg.norm(lane['center'] - centre_xy[None], axis=1)) > lane_dist:      # VectorMap.get_lanes_within
            continue
        tls = float(tls_of(lane['id']))
        for k in ('center', 'left_edge', 'right_edge'):                                # dict order of data_utils.py:197
            v = lane[k]
            if k not in include or v is None:
                continue
            rate = edge_rate if 'edge' in k else center_rate
            if v.shape[0] > rate:
                v = v[::rate]
            x, y = _rot(v[:, 0] - centre_xy[0], v[:, 1] - centre_xy[1], -centre_h)        # transform_coords_np(agent_from_world_tf)
            v = np.stack([x, y], axis=1)
            v = v[(np.abs(v[:, 0]) < map_range) & (np.abs(v[:, 1]) < map_range)]
            n = v.shape[0]
            if n < 2:
                continue
            if n > max_lane_points:
                chunk = list(np.arange(0, n, max_lane_points, dtype=int))
                if chunk[-1] != n:
                    chunk.append(n)
            else:
                chunk = [0, n - 1]                                                       # (sic) data_utils.py:224: drops the last point
            for i in range(len(chunk) - 1):
                c = v[chunk[i]:chunk[i + 1]]
                m = len(c) - 1
                if m < 1:
                    continue
                vec = np.zeros((max_lane_points - 1, 6))
                vec[:m, 0:2], vec[:m, 2:4] = c[:-1], c[1:]
                vec[:m, 4], vec[:m, 5] = line_type[k], tls
                vecs.append(vec)
    if not vecs:
        return torch.zeros(1, max_lane_points - 1, 6)
    return torch.tensor(np.stack(vecs, axis=0)).float()


def center_vec_init_map(full_vec, local_pos, local_range=200.0, max_points=2048):
    """prosim/dataset/format_utils.py:153-263 for one scene -> InputMaskData fields (input [1, M, 19, 11], mask, position, heading)."""
    mask = full_vec[..., 4] > 0
    cnt = mask.sum(dim=1)
    cnt[cnt == 0] = 1
    position = full_vec[..., :2].sum(dim=1) / cnt[:, None]
    full_dist = torch.norm(position - local_pos, dim=-1)
    keep = full_dist < local_range
    local_vec, local_dist = full_vec[keep], full_dist[keep]
    P = local_vec.shape[1]
    local_mask = torch.zeros(max_points, P, dtype=torch.bool)
    p_num = min(max_points, local_vec.shape[0])
    local_mask[:p_num] = local_vec[:p_num, :, 4] > 0
    if local_vec.shape[0] > max_points:
        local_vec = local_vec[torch.argsort(local_dist)[:max_points]]
    else:
        local_vec = torch.cat([local_vec, torch.zeros([max_points - local_vec.shape[0]] + list(local_vec.shape[1:]))], dim=0)
    # local_map_to_sym_coord (:184-218)
    M = local_vec.shape[0]
    cnt = (local_vec[..., 4] > 0).sum(dim=1)
    start = local_vec[:, 0, :2]
    end = local_vec[torch.arange(M)[:, None], cnt[:, None] - 1, 2:4].squeeze(1)
    heading = torch.atan2(end[..., 1] - start[..., 1], end[..., 0] - start[..., 0])[:, None]
    pos = ((start + end) / 2)[:, None]

    def rot(x, y, a):
        c, s = torch.cos(a), torch.sin(a)
        return torch.stack([x * c - y * s, x * s + y * c], dim=-1)

    local_vec = local_vec.clone()
    local_vec[..., :2] -= pos
    local_vec[..., :2] = rot(local_vec[..., 0], local_vec[..., 1], -heading)
    local_vec[..., 2:4] -= pos
    local_vec[..., 2:4] = rot(local_vec[..., 2], local_vec[..., 3], -heading)
    types = local_vec[..., 4]
    one_hot = torch.zeros_like(local_vec[..., :3])
    for t in (1, 2, 3):
        one_hot[..., t - 1] = (types == t)
    diff = local_vec[..., 2:4] - local_vec[..., :2]
    direction = diff / torch.clip(torch.norm(diff, dim=-1, keepdim=True), min=1e-6)
    all_map = torch.cat([local_vec, one_hot, direction], dim=-1)
    return all_map[None], local_mask[None], pos[None], heading[None]


# ----------------------------------------------------------------------------------------------- observations
def center_obs(hist, extent, agent_type, names, keep_rows):
    """prosim/dataset/format_utils.py:357-447 (get_center_obs) for one scene.  hist [n, 11, 8] states in the centred frame
    ending at the step the observation is taken at (NaN = absent), extent [n, 2], agent_type [n].  Agents whose state at that
    step is NaN are dropped unless listed in keep_rows (the target agents)."""
    n = hist.shape[0]
    origin = hist[:, -1]
    sel = [i for i in range(n) if i in keep_rows or not np.isnan(origin[i, [0, 1, 7, 3, 4, 5, 6]]).any()]
    N = len(sel)
    inp = torch.full((1, N, HIST, 24), float('nan'))
    pos, head = torch.zeros(1, N, 2), torch.zeros(1, N)
    ids = []
    for k, i in enumerate(sel):
        rel = _to_frame(hist[i], origin[i, :2], origin[i, 7])
        obs = np.stack([rel[:, 0], rel[:, 1], np.sin(rel[:, 7]), np.cos(rel[:, 7]), rel[:, 3], rel[:, 4], rel[:, 5], rel[:, 6]], axis=1)
        inp[0, k, :, :8] = torch.from_numpy(obs).float()
        inp[0, k, :, 8:10] = torch.from_numpy(extent[i]).float()[None]
        inp[0, k, :, 10:13] = 0.0
        if 1 <= agent_type[i] <= 3:
            inp[0, k, :, 9 + int(agent_type[i])] = 1.0
        inp[0, k, :, 13:24] = torch.eye(HIST)
        pos[0, k] = torch.from_numpy(origin[i, :2]).float()
        head[0, k] = float(origin[i, 7])
        ids.append(names[i])
    mask = ~inp.isnan()
    return InputMaskData(inp, mask, pos, head, [ids])


def load_demo_scene(cache_root, scene='scene_11', ts=10, steps=20, max_agents=16, split='waymo_train'):
    """One cached scene -> SceneBatch with extras {init_obs, init_map, prompt, condition, all_t_indices, fut_obs} (configs[0]:
    16 agents, 20 steps -> ticks [0, 10]).  ``ts`` is the current step (11 history steps 0..10 at the default)."""
    base = os.path.join(cache_root, 'trajdata_cache', split)
    names, types, states, extents, info = read_scene(os.path.join(base, scene))
    if 'ego' not in names:
        raise ValueError('the scene has no ego agent to centre on')
    ego = names.index('ego')
    c_xy, c_h = states[ego, ts, :2].copy(), float(states[ego, ts, 7])
    present = [i for i in range(len(names)) if not np.isnan(states[i, ts, 0]) and types[i] in (1, 2, 3)]
    present.sort(key=lambda i: (i != ego, float(np.linalg.norm(states[i, ts, :2] - c_xy))))
    present = present[:max_agents]
    cen = _to_frame(states[present], c_xy, c_h)                                   # [n, T, 8] in the centred frame
    ext_all = extents[present]
    ext = np.nan_to_num(np.where(np.isnan(ext_all), -1.0, ext_all).max(axis=1), nan=0.0)     # format_utils.py:320-324
    a_type = types[present]
    a_names = [names[i] for i in present]
    T = states.shape[1]
    fut_len = np.array([int((~np.isnan(cen[i, ts + 1:, 0])).sum()) for i in range(len(present))])
    # format_utils.py:765-769: targets = agents with a future, ranked by future length (stable)
    tgt = sorted([i for i in range(len(present)) if fut_len[i] > 0], key=lambda i: -fut_len[i])

    def window(end):                    # the 11 steps ending at step `end` (inclusive), NaN padded
        out = np.full((len(present), HIST, 8), np.nan)
        lo = end - HIST + 1
        for k in range(HIST):
            if 0 <= lo + k < T:
                out[:, k] = cen[:, lo + k]
        return out

    init_obs = center_obs(window(ts), ext, a_type, a_names, set(tgt))
    # format_utils.py:699-713: ROLLOUT split, TAIL_PADDING: arange(steps)[::10]
    all_t = np.arange(steps)[::10]
    fut = {}
    for t in all_t[1:]:
        # format_utils.py:667-687 with FUTURE_OBS_TYPE 'latest': the ground-truth window ending at ts + t; the rollout overwrites the
        # motion columns of the agents it controls (traj_sam.py:266-270) and keeps the static ones
        f = center_obs(window(ts + int(t)), ext, a_type, a_names, set(tgt))
        # the rollout needs a slot for every controlled agent at every tick
        assert all(n in f.agent_ids[0] for n in [a_names[i] for i in tgt])
        fut[int(t)] = f
    # prompt: prosim/dataset/prompt_utils.py:24-150 (agent status: local velocity, extent, type one-hot)
    n = len(tgt)
    last = cen[tgt, ts]
    vx, vy = _rot(last[:, 3], last[:, 4], -last[:, 7])
    one_hot = np.zeros((n, 3))
    for k, i in enumerate(tgt):
        if 1 <= a_type[i] <= 3:
            one_hot[k, a_type[i] - 1] = 1.0
    prompt = {'prompt': torch.from_numpy(np.concatenate([np.stack([vx, vy], 1), ext[tgt], one_hot], axis=1)).float()[None],
              'prompt_mask': torch.ones(1, n, dtype=torch.bool),
              'position': torch.from_numpy(last[:, :2]).float()[None], 'heading': torch.from_numpy(last[:, 7:8]).float()[None],
              'agent_type': torch.from_numpy(a_type[tgt]).long()[None], 'agent_ids': [[a_names[i] for i in tgt]]}
    # map
    tls = info['tls']
    status = {}
    if tls is not None and len(tls):
        cur = tls[tls['scene_ts'] == ts]
        status = {str(k): float(v) for k, v in zip(cur['lane_id'], cur['status'])}
    lanes = read_vector_map(os.path.join(base, 'maps', f'{info["meta"].location}.pb'))
    full_vec = vectorize_lanes(lanes, c_xy, c_h, lambda lane_id: status.get(lane_id, -1.0))
    m_in, m_mask, m_pos, m_head = center_vec_init_map(full_vec, torch.zeros(2))   # the centre agent sits at the origin of its own frame
    keep = int(m_mask.any(-1).sum())
    keep = max(keep, 1)
    init_map = InputMaskData(m_in[:, :keep].contiguous(), m_mask[:, :keep].contiguous(), m_pos[:, :keep].contiguous(),
                             m_head[:, :keep].contiguous())
    extras = {'init_obs': init_obs, 'init_map': init_map, 'prompt': BatchPrompt({'motion_pred': prompt}),
              'condition': BatchCondition({}), 'all_t_indices': torch.tensor(all_t), 'fut_obs': BatchDataDict(fut)}
    batch = SceneBatch([f'{split}_{scene}'], extras)
    tf = np.eye(3)
    tf[0, 0], tf[0, 1], tf[1, 0], tf[1, 1] = math.cos(c_h), -math.sin(c_h), math.sin(c_h), math.cos(c_h)
    tf[:2, 2] = c_xy
    batch.centered_world_from_agent_tf = torch.from_numpy(tf).float()[None]
    return batch


# ----------------------------------------------------------------------------------------------- fixtures
def batch_to_arrays(batch):
    """Flat {name: numpy array} form of a loaded scene (committed as a test fixture: the GPU box has no demo_dataset)."""
    ex = batch.extras
    out = {'all_t_indices': ex['all_t_indices'].numpy(), 'tf': batch.centered_world_from_agent_tf.numpy(),
           'scene_ids': np.array(batch.scene_ids)}
    for key, d in [('init_obs', ex['init_obs']), ('init_map', ex['init_map'])] + [(f'fut_obs_{t}', ex['fut_obs'][t]) for t in ex['fut_obs'].keys()]:
        for k in ('input', 'mask', 'position', 'heading'):
            out[f'{key}.{k}'] = d[k].numpy()
        if d.agent_ids is not None:
            out[f'{key}.agent_ids'] = np.array(d.agent_ids[0])
    p = ex['prompt']['motion_pred']
    for k in ('prompt', 'prompt_mask', 'position', 'heading', 'agent_type'):
        out[f'prompt.{k}'] = p[k].numpy()
    out['prompt.agent_ids'] = np.array(p['agent_ids'][0])
    return out


def batch_from_arrays(a):
    def imd(key):
        ids = [a[f'{key}.agent_ids'].tolist()] if f'{key}.agent_ids' in a else None
        return InputMaskData(torch.from_numpy(a[f'{key}.input']), torch.from_numpy(a[f'{key}.mask']), torch.from_numpy(a[f'{key}.position']),
                             torch.from_numpy(a[f'{key}.heading']), ids)
    ticks = [int(t) for t in a['all_t_indices']]
    prompt = {k: torch.from_numpy(a[f'prompt.{k}']) for k in ('prompt', 'prompt_mask', 'position', 'heading', 'agent_type')}
    prompt['agent_ids'] = [a['prompt.agent_ids'].tolist()]
    extras = {'init_obs': imd('init_obs'), 'init_map': imd('init_map'), 'prompt': BatchPrompt({'motion_pred': prompt}),
              'condition': BatchCondition({}), 'all_t_indices': torch.from_numpy(a['all_t_indices']),
              'fut_obs': BatchDataDict({t: imd(f'fut_obs_{t}') for t in ticks[1:]})}
    batch = SceneBatch(a['scene_ids'].tolist(), extras)
    batch.centered_world_from_agent_tf = torch.from_numpy(a['tf'])
    return batch
