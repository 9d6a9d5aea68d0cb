"""Parameter table of the rollout path, under the reference's own ``state_dict`` names.

The names/shapes below are what ``ProSim(cfg).state_dict()`` holds for the released model
shape (verified against the live reference in tests/test_oracle_vs_reference.py and pinned
in tests/golden/state_dict_keys.json), so a reference checkpoint loads by key:
  scene_encoder.*            prosim/models/scene_encoder/{pointnet_encoder,attn_fusion}.py
  prompt_encoder.motion_pred prosim/models/prompt_encoder/base.py:30
  decoder.*                  prosim/models/decoder/sym_coord.py:28-35
  policy.act_decoder.*       prosim/models/policy/act_decoder.py:50-76,177-196
  condition_transformers.policy_decoder.*   condition_transformer/{condition_encoders,condition_attns}.py
No checkpoint ships with the reference (README.md:56-58), so parity work uses
``random_state_dict``: every tensor is filled from its own generator keyed by
(seed, crc32(name)), independent of module construction order.
"""
import zlib
from collections import OrderedDict

import torch

D = 128
HEADS = 8
HEAD_DIM = 16


def _mlp(prefix, dims, ret_before_act=False, without_norm=False):
    """layers/mlp.py:475-494 -- Linear [+LN] + ReLU stack; index math of nn.Sequential."""
    out = []
    idx = 0
    n = len(dims) - 1
    for i in range(n):
        out.append((f'{prefix}.mlp.{idx}.weight', (dims[i + 1], dims[i]), 'linear_w'))
        out.append((f'{prefix}.mlp.{idx}.bias', (dims[i + 1],), 'linear_b'))
        idx += 1
        if i < n - 1:
            if not without_norm:
                out.append((f'{prefix}.mlp.{idx}.weight', (dims[i + 1],), 'ln_w'))
                out.append((f'{prefix}.mlp.{idx}.bias', (dims[i + 1],), 'ln_b'))
                idx += 1
            idx += 1  # ReLU
    return out


def _attn_layer(prefix):
    """layers/attention_layer.py:28-54 (has_pos_emb=True; both prenorm names exist even when shared)."""
    p = prefix
    out = [
        (f'{p}.to_q.weight', (D, D), 'linear_w'), (f'{p}.to_q.bias', (D,), 'linear_b'),
        (f'{p}.to_k.weight', (D, D), 'linear_w'),
        (f'{p}.to_v.weight', (D, D), 'linear_w'), (f'{p}.to_v.bias', (D,), 'linear_b'),
        (f'{p}.to_k_r.weight', (D, D), 'linear_w'),
        (f'{p}.to_v_r.weight', (D, D), 'linear_w'), (f'{p}.to_v_r.bias', (D,), 'linear_b'),
        (f'{p}.to_s.weight', (D, D), 'linear_w'), (f'{p}.to_s.bias', (D,), 'linear_b'),
        (f'{p}.to_g.weight', (D, 2 * D), 'linear_w'), (f'{p}.to_g.bias', (D,), 'linear_b'),
        (f'{p}.to_out.weight', (D, D), 'linear_w'), (f'{p}.to_out.bias', (D,), 'linear_b'),
        (f'{p}.ff_mlp.0.weight', (4 * D, D), 'linear_w'), (f'{p}.ff_mlp.0.bias', (4 * D,), 'linear_b'),
        (f'{p}.ff_mlp.3.weight', (D, 4 * D), 'linear_w'), (f'{p}.ff_mlp.3.bias', (D,), 'linear_b'),
    ]
    for ln in ('attn_prenorm_x_src', 'attn_prenorm_x_dst', 'attn_prenorm_r', 'attn_postnorm',
               'ff_prenorm', 'ff_postnorm'):
        out += [(f'{p}.{ln}.weight', (D,), 'ln_w'), (f'{p}.{ln}.bias', (D,), 'ln_b')]
    return out


def _pointnet(prefix, in_dim, num_layers, num_pre):
    """scene_encoder/pointnet_encoder.py:13-22."""
    return (_mlp(f'{prefix}.pre_mlps', [in_dim] + [D] * num_pre)
            + _mlp(f'{prefix}.mlps', [2 * D] + [D] * (num_layers - num_pre))
            + _mlp(f'{prefix}.out_mlps', [D] * 3, ret_before_act=True, without_norm=True))


SHARED_LN_STACKS = ('scene_encoder.a2a_attn_layers', 'scene_encoder.s2s_attn_layers',
                    'decoder.p2p_attn_layers',
                    'condition_transformers.policy_decoder.condition_attn.attn_layers')


V_ACTION_TAGS = ('Accelerate', 'Decelerate', 'KeepSpeed', 'Stopping', 'LeftLaneChange', 'RightLaneChange', 'KeepLane',
                 'LeftTurn', 'RightTurn', 'Straight', 'Parked')      # PROMPT.CONDITION.MOTION_TAG.USED_TAGS order
V_ACTION_TAG_ID = {'Stopping': 0, 'Accelerate': 1, 'Decelerate': 2, 'KeepSpeed': 3, 'LeftLaneChange': 4,
                   'RightLaneChange': 5, 'KeepLane': 6, 'LeftTurn': 7, 'RightTurn': 8, 'Straight': 9,
                   'Parked': 10}                                      # dataset/motion_tag_utils.py:4-15


def cond_types(x):
    """False / True / a sequence of PROMPT.CONDITION.TYPES -> tuple of types (True = the goal-only demo config)."""
    if x is True:
        return ('goal',)
    if not x:
        return ()
    return tuple(x)


def param_specs(goal_condition=False, num_layers=6, cond_layers=3, obs_fusion='replace'):
    types = cond_types(goal_condition)
    specs = []
    specs += _pointnet('scene_encoder.map_encoder', 11, 5, 3)
    specs += _pointnet('scene_encoder.obs_encoder', 24, 3, 1)
    for stack in ('a2a', 's2s'):
        for i in range(num_layers):
            specs += _attn_layer(f'scene_encoder.{stack}_attn_layers.{i}')
    if obs_fusion == 'mlp':     # MODEL.OBS_UPDATE.FUSION = 'mlp' (scene_encoder/attn_fusion.py:18-19)
        specs += _mlp('scene_encoder.obs_update_mlp', [2 * D, D, D], ret_before_act=True)
    specs += _mlp('prompt_encoder.motion_pred.state_encoder', [7, D, D], ret_before_act=True)
    for stack in ('p2p', 's2p'):
        for i in range(num_layers):
            specs += _attn_layer(f'decoder.{stack}_attn_layers.{i}')
    if types:
        # condition_transformer/base.py:22-35: one encoder per condition type in TYPES order, then the GNN attention
        ct = 'condition_transformers.policy_decoder'
        for t in types:
            if t == 'goal':
                specs += _mlp(f'{ct}.condition_encoders.goal.goal_encoder', [2, D, D], ret_before_act=True,
                              without_norm=True)
            elif t == 'v_action_tag':   # condition_encoders.py:71-74: one learned vector per used tag
                specs += [(f'{ct}.condition_encoders.v_action_tag.tag_encoder.{tag}', (D,), 'emb') for tag in V_ACTION_TAGS]
            elif t == 'drag_point':     # condition_encoders.py:159-160, config DRAG_POINTS: 1 pre layer, 3 mlp layers
                specs += _pointnet(f'{ct}.condition_encoders.drag_point.pointnet_encoder', 2, 3, 1)
            else:
                raise NotImplementedError(f'condition type {t}')
        for i in range(cond_layers):
            specs += _attn_layer(f'{ct}.condition_attn.attn_layers.{i}')
    pa = 'policy.act_decoder'
    for stack in ('a2p', 'm2p'):
        for i in range(num_layers):
            specs += _attn_layer(f'{pa}.{stack}_attn_layers.{i}')
    specs += _mlp(f'{pa}.motion_head', [D, D, D // 2, 50], ret_before_act=True)
    for i in range(3):
        specs += [(f'{pa}.CG_decode.CGs.{i}.MLP.0.weight', (D, D), 'linear_w'),
                  (f'{pa}.CG_decode.CGs.{i}.MLP.0.bias', (D,), 'linear_b'),
                  (f'{pa}.CG_decode.CGs.{i}.MLP.1.weight', (D,), 'ln_w'),
                  (f'{pa}.CG_decode.CGs.{i}.MLP.1.bias', (D,), 'ln_b')]
    specs += [(f'{pa}.motion_anchors.weight', (3, D), 'emb')]
    specs += _mlp(f'{pa}.pred_mlp', [D, D, D // 2, 2], ret_before_act=True)
    return specs


def random_state_dict(seed=0, goal_condition=False, dtype=torch.float32, obs_fusion='replace'):
    """Seeded random weights, torch-default-like scales; LayerNorm affine is NOT identity on purpose
    (so the gamma/beta folding in the kernels is exercised).  Non-bipartite layers share one LayerNorm
    under two names (attention_layer.py:48-49): the dst copy is tied to the src one.
    goal_condition: False / True (= ('goal',)) / a tuple of PROMPT.CONDITION.TYPES."""
    goal_condition = cond_types(goal_condition)
    sd = OrderedDict()
    shapes = {n: sh for n, sh, _ in param_specs(goal_condition, obs_fusion=obs_fusion)}
    for name, shape, kind in param_specs(goal_condition, obs_fusion=obs_fusion):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        if kind == 'linear_w':
            bound = 1.0 / (shape[1] ** 0.5)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        elif kind == 'linear_b':
            fan_in = shapes[name[:-len('bias')] + 'weight'][1]
            bound = 1.0 / (fan_in ** 0.5)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        elif kind == 'ln_w':
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == 'ln_b':
            t = 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:  # emb
            t = torch.randn(shape, generator=g, dtype=torch.float64)
        sd[name] = t.to(torch.float32).to(dtype)
    for name in list(sd.keys()):
        if '.attn_prenorm_x_dst.' in name and name.startswith(SHARED_LN_STACKS):
            sd[name] = sd[name.replace('attn_prenorm_x_dst', 'attn_prenorm_x_src')].clone()
    return sd


_SHAPES = {}


def param_shapes_cache(goal_condition):
    goal_condition = cond_types(goal_condition)
    if goal_condition not in _SHAPES:
        _SHAPES[goal_condition] = [(n, s) for n, s, _ in param_specs(goal_condition)]
    return _SHAPES[goal_condition]


# ----------------------------------------------------------------------------- packing (mirror of csrc/weights_layout.h)
ATTN_FP32_FLOATS = 512 + 3 * (16384 + 128) + 2 * 16384 + 12288 + (16384 + 128) + (2 * 16384 + 128) + (16384 + 128) + 512 \
    + (65536 + 512) + (65536 + 128) + 256
ATTN_TC_FLOATS = 8 * 3072 + 8 * 4096 + 2 * 4 * 8192 + 32 * 8192 + 3 * 4 * 8192 + 8 * 4096 + 8 * 8192 + 32 * 8192
ATTN_LAYER_FLOATS = ATTN_FP32_FLOATS + ATTN_TC_FLOATS
POINTNET_FLOATS = (24 * 128 + 384) + (16384 + 384) + (16384 + 128) + (2 * 16384 + 384) + (16384 + 128) + 2 * (16384 + 128)
_MLP3_FLOATS = 2 * (16384 + 384) + (64 * 128 + 128)
HEAD_FLOATS = 3 * 128 + 3 * (16384 + 384) + 2 * _MLP3_FLOATS
MLP2_FLOATS = 8 * 128 + 384 + 16384 + 128
OBS_FUSE_FLOATS = 256 * 128 + 384 + 16384 + 128


def _f64(t):
    return t.detach().double().cpu()


def _fold96(wvrgt):
    """[128 d][128 c] -> [96 d][128 c]: the PE's feature groups 64..95 and 96..127 are identical, so their weight
    rows are summed once here instead of per edge."""
    out = wvrgt[:96].clone()
    out[64:96] += wvrgt[96:128]
    return out


def _tf32_rna(x):
    """cvt.rna.tf32.f32: round the fp32 mantissa to 10 bits, ties away from zero (the low 13 bits end up zero)."""
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & -8192).view(torch.float32)


def _umma_chunk(b):
    """[N][K] fp32 B operand -> flat [hi | lo], each in the tcgen05 K-major no-swizzle core-matrix order
    idx(n, k) = (n // 8) * 8K + (k // 4) * 32 + (n % 8) * 4 + k % 4   (weights_layout.h, aw::TC_*)."""
    b = b.contiguous().float()
    n, k = b.shape
    hi = _tf32_rna(b)
    lo = _tf32_rna(b - hi)
    lay = lambda m: m.reshape(n // 8, 8, k // 4, 4).permute(0, 2, 1, 3).contiguous().reshape(-1)
    return torch.cat([lay(hi), lay(lo)])


def _pack_attn_tc(wqt, wkrg, wvrgt, wvrg96t, wst, wgat, wgxt, wot, w1t, w2t, wkt, wvt):
    """Tensor-core operand block of one layer from the fp32-rounded K-major weights the FFMA kernels use."""
    f = lambda x: x.float()
    wqt, wkrg, wvrgt, wvrg96t, wst, wgat, wgxt, wot, w1t, w2t, wkt, wvt = map(f, (wqt, wkrg, wvrgt, wvrg96t, wst, wgat, wgxt,
                                                                                 wot, w1t, w2t, wkt, wvt))
    out = []
    out += [_umma_chunk(wvrg96t[:, h * 16:(h + 1) * 16].t()) for h in range(HEADS)]
    out += [_umma_chunk(wvrgt[:, h * 16:(h + 1) * 16].t()) for h in range(HEADS)]
    kchunks = lambda wt: [_umma_chunk(wt.t()[:, 32 * c:32 * c + 32]) for c in range(4)]
    out += kchunks(wgat) + kchunks(wot)
    w1, w2 = w1t.t(), w2t.t()                         # [512][128], [128][512]
    up = lambda j: _umma_chunk(w1[32 * j:32 * j + 32, :])
    down = lambda j: _umma_chunk(w2[:, 32 * j:32 * j + 32])
    out += [up(0), up(1)]                             # consumption order of the MMA warp (tc_post.cuh)
    for j in range(14):
        out += [up(j + 2), down(j)]
    out += [down(14), down(15)]
    out += kchunks(wqt) + kchunks(wst) + kchunks(wgxt)
    out += [_umma_chunk(wkrg[h * 16:(h + 1) * 16, :].t()) for h in range(HEADS)]
    for c in range(4):                                # kv_tc.cuh: to_k and to_v chunks interleaved
        out += [_umma_chunk(wkt.t()[:, 32 * c:32 * c + 32]), _umma_chunk(wvt.t()[:, 32 * c:32 * c + 32])]
    # post_sw.cuh: the FFN with the weights as the M-side operand, hidden tiles of 128 features, 4 k-chunks each
    up2 = lambda t: [_umma_chunk(w1[128 * t:128 * t + 128, 32 * c:32 * c + 32]) for c in range(4)]
    down2 = lambda t: [_umma_chunk(w2[:, 128 * t + 32 * c:128 * t + 32 * c + 32]) for c in range(4)]
    out += up2(0) + up2(1) + down2(0) + up2(2) + down2(1) + up2(3) + down2(2) + down2(3)
    out = torch.cat(out)
    assert out.numel() == ATTN_TC_FLOATS
    return out


def pack_attn_layer(sd, p):
    """One AttentionLayer -> the aw:: layout.  LayerNorm(r)'s affine and the 1/sqrt(head_dim) scale are folded
    into the weights (see weights_layout.h); the folds are evaluated in fp64 and rounded once."""
    w = lambda n: _f64(sd[f'{p}.{n}'])
    g_r, b_r = w('attn_prenorm_r.weight'), w('attn_prenorm_r.bias')
    wkr, wvr = w('to_k_r.weight'), w('to_v_r.weight')
    wg = w('to_g.weight')
    parts = [
        w('attn_prenorm_x_src.weight'), w('attn_prenorm_x_src.bias'),
        w('attn_prenorm_x_dst.weight'), w('attn_prenorm_x_dst.bias'),
        (0.25 * w('to_q.weight')).t(), 0.25 * w('to_q.bias'),
        w('to_k.weight').t(), wkr @ b_r,
        w('to_v.weight').t(), w('to_v.bias') + wvr @ b_r + w('to_v_r.bias'),
        wkr * g_r[None, :],
        (wvr * g_r[None, :]).t(),
        _fold96((wvr * g_r[None, :]).t()),
        w('to_s.weight').t(), w('to_s.bias'),
        wg[:, :D].t(), wg[:, D:].t(), w('to_g.bias'),
        w('to_out.weight').t(), w('to_out.bias'),
        w('attn_postnorm.weight'), w('attn_postnorm.bias'),
        w('ff_prenorm.weight'), w('ff_prenorm.bias'),
        w('ff_mlp.0.weight').t(), w('ff_mlp.0.bias'),
        w('ff_mlp.3.weight').t(), w('ff_mlp.3.bias'),
        w('ff_postnorm.weight'), w('ff_postnorm.bias'),
    ]
    tc = _pack_attn_tc(wqt=parts[4], wkrg=parts[10], wvrgt=parts[11], wvrg96t=parts[12], wst=parts[13], wgat=parts[15],
                       wgxt=parts[16], wot=parts[18], w1t=parts[24], w2t=parts[26], wkt=parts[6], wvt=parts[8])
    out = torch.cat([x.contiguous().reshape(-1) for x in parts]).float()
    assert out.numel() == ATTN_FP32_FLOATS
    return torch.cat([out, tc])


def _pad_rows(wt, rows):
    out = torch.zeros(rows, wt.shape[1], dtype=wt.dtype)
    out[:wt.shape[0]] = wt
    return out


def _pad_cols(wt, cols=D):
    out = torch.zeros(wt.shape[0], cols, dtype=wt.dtype)
    out[:, :wt.shape[1]] = wt
    return out


def _pad_vec(v, n=D):
    out = torch.zeros(n, dtype=v.dtype)
    out[:v.shape[0]] = v
    return out


def pack_pointnet(sd, p, n_pre):
    w = lambda n: _f64(sd[f'{p}.{n}'])
    z = torch.zeros(D, dtype=torch.float64)
    zw = torch.zeros(D, D, dtype=torch.float64)
    if n_pre == 1:
        pre = [_pad_rows(w('pre_mlps.mlp.0.weight').t(), 24), w('pre_mlps.mlp.0.bias'), z, z,
               zw, z, z, z, zw, z]
    else:
        pre = [_pad_rows(w('pre_mlps.mlp.0.weight').t(), 24), w('pre_mlps.mlp.0.bias'),
               w('pre_mlps.mlp.1.weight'), w('pre_mlps.mlp.1.bias'),
               w('pre_mlps.mlp.3.weight').t(), w('pre_mlps.mlp.3.bias'),
               w('pre_mlps.mlp.4.weight'), w('pre_mlps.mlp.4.bias'),
               w('pre_mlps.mlp.6.weight').t(), w('pre_mlps.mlp.6.bias')]
    m0 = w('mlps.mlp.0.weight')
    parts = pre + [m0[:, :D].t(), m0[:, D:].t(), w('mlps.mlp.0.bias'), w('mlps.mlp.1.weight'), w('mlps.mlp.1.bias'),
                   w('mlps.mlp.3.weight').t(), w('mlps.mlp.3.bias'),
                   w('out_mlps.mlp.0.weight').t(), w('out_mlps.mlp.0.bias'),
                   w('out_mlps.mlp.2.weight').t(), w('out_mlps.mlp.2.bias')]
    out = torch.cat([x.contiguous().reshape(-1) for x in parts]).float()
    assert out.numel() == POINTNET_FLOATS
    return out


def pointnet_tc_floats(n_pre):
    return (1 + (8 if n_pre == 3 else 0) + 8 + 4 + 4 + 4) * 8192


def pack_pointnet_tc(sd, p, n_pre):
    """Tensor-core operand block of one PointNet (csrc/pointnet_tc.cuh): every weight matrix [128 out][K in] cut into
    K = 32 chunks, each pre-split into tf32 hi / lo in the tcgen05 core-matrix order, in the order the kernel consumes them:
    pre_mlps.0 (input dim zero-padded to 32) [, pre_mlps.3, pre_mlps.6], mlps.0 (K = 256: point half, pooled half), mlps.3,
    out_mlps.0, out_mlps.2."""
    w = lambda n: _f64(sd[f'{p}.{n}']).float()
    w0 = w('pre_mlps.mlp.0.weight')
    mats = [torch.cat([w0, torch.zeros(D, 32 - w0.shape[1])], dim=1)]
    if n_pre == 3:
        mats += [w('pre_mlps.mlp.3.weight'), w('pre_mlps.mlp.6.weight')]
    mats += [w('mlps.mlp.0.weight'), w('mlps.mlp.3.weight'), w('out_mlps.mlp.0.weight'), w('out_mlps.mlp.2.weight')]
    out = torch.cat([_umma_chunk(m[:, 32 * c:32 * c + 32]) for m in mats for c in range(m.shape[1] // 32)])
    assert out.numel() == pointnet_tc_floats(n_pre)
    return out


def _pack_mlp3(sd, p):
    w = lambda n: _f64(sd[f'{p}.{n}'])
    parts = [w('mlp.0.weight').t(), w('mlp.0.bias'), w('mlp.1.weight'), w('mlp.1.bias'),
             _pad_cols(w('mlp.3.weight').t()), _pad_vec(w('mlp.3.bias')), _pad_vec(w('mlp.4.weight')),
             _pad_vec(w('mlp.4.bias')),
             _pad_cols(w('mlp.6.weight').t()), _pad_vec(w('mlp.6.bias'))]
    return [x.contiguous().reshape(-1) for x in parts]


def pack_head(sd, p='policy.act_decoder'):
    w = lambda n: _f64(sd[f'{p}.{n}'])
    parts = [w('motion_anchors.weight').reshape(-1)]
    for i in range(3):
        parts += [w(f'CG_decode.CGs.{i}.MLP.0.weight').t().contiguous().reshape(-1), w(f'CG_decode.CGs.{i}.MLP.0.bias'),
                  w(f'CG_decode.CGs.{i}.MLP.1.weight'), w(f'CG_decode.CGs.{i}.MLP.1.bias')]
    parts += _pack_mlp3(sd, f'{p}.motion_head') + _pack_mlp3(sd, f'{p}.pred_mlp')
    out = torch.cat(parts).float()
    assert out.numel() == HEAD_FLOATS
    return out


def pack_mlp2(sd, p, with_ln):
    w = lambda n: _f64(sd[f'{p}.{n}'])
    z = torch.zeros(D, dtype=torch.float64)
    second = 'mlp.3' if with_ln else 'mlp.2'
    parts = [_pad_rows(w('mlp.0.weight').t(), 8), w('mlp.0.bias'),
             w('mlp.1.weight') if with_ln else z, w('mlp.1.bias') if with_ln else z,
             w(f'{second}.weight').t(), w(f'{second}.bias')]
    out = torch.cat([x.contiguous().reshape(-1) for x in parts]).float()
    assert out.numel() == MLP2_FLOATS
    return out


def pack_obs_fuse(sd, p='scene_encoder.obs_update_mlp'):
    """MLP([256, 128, 128], ret_before_act): Linear -> LN -> ReLU -> Linear, first weight K-major [256 k][128 n] with the
    OLD token's 128 inputs first (torch.cat([old, new]), attn_fusion.py:196) -- csrc/weights_layout.h fw::."""
    w = lambda n: _f64(sd[f'{p}.{n}'])
    parts = [w('mlp.0.weight').t(), w('mlp.0.bias'), w('mlp.1.weight'), w('mlp.1.bias'), w('mlp.3.weight').t(), w('mlp.3.bias')]
    out = torch.cat([x.contiguous().reshape(-1) for x in parts]).float()
    assert out.numel() == OBS_FUSE_FLOATS
    return out


def pack_model(sd, num_layers=6, cond_layers=3):
    """state_dict -> (arena fp32 [n], {section: float offset}).  Sections are 256-byte aligned."""
    goal = any(k.startswith('condition_transformers.') for k in sd)
    sections = OrderedDict()
    sections['map_enc'] = pack_pointnet(sd, 'scene_encoder.map_encoder', 3)
    sections['obs_enc'] = pack_pointnet(sd, 'scene_encoder.obs_encoder', 1)
    sections['map_enc_tc'] = pack_pointnet_tc(sd, 'scene_encoder.map_encoder', 3)
    sections['obs_enc_tc'] = pack_pointnet_tc(sd, 'scene_encoder.obs_encoder', 1)
    stacks = [('enc_a2a', 'scene_encoder.a2a_attn_layers', num_layers),
              ('enc_s2s', 'scene_encoder.s2s_attn_layers', num_layers),
              ('dec_p2p', 'decoder.p2p_attn_layers', num_layers),
              ('dec_s2p', 'decoder.s2p_attn_layers', num_layers),
              ('pol_a2p', 'policy.act_decoder.a2p_attn_layers', num_layers),
              ('pol_m2p', 'policy.act_decoder.m2p_attn_layers', num_layers)]
    if goal:
        stacks.append(('cond_attn', 'condition_transformers.policy_decoder.condition_attn.attn_layers', cond_layers))
    for name, prefix, n in stacks:
        sections[name] = torch.cat([pack_attn_layer(sd, f'{prefix}.{i}') for i in range(n)])
    sections['prompt_mlp'] = pack_mlp2(sd, 'prompt_encoder.motion_pred.state_encoder', True)
    if 'scene_encoder.obs_update_mlp.mlp.0.weight' in sd:
        sections['obs_fuse'] = pack_obs_fuse(sd)
    ce = 'condition_transformers.policy_decoder.condition_encoders'
    if f'{ce}.goal.goal_encoder.mlp.0.weight' in sd:
        sections['goal_mlp'] = pack_mlp2(sd, f'{ce}.goal.goal_encoder', False)
    if f'{ce}.v_action_tag.tag_encoder.{V_ACTION_TAGS[0]}' in sd:
        # row = V_Action_MotionTag value (the id found in the condition input), 11 rows padded to 16
        table = torch.zeros(16, D)
        for tag, tid in V_ACTION_TAG_ID.items():
            table[tid] = sd[f'{ce}.v_action_tag.tag_encoder.{tag}'].float()
        sections['tag_vec'] = table.reshape(-1)
        dt = torch.arange(64, dtype=torch.float32)               # FourierEmbeddingFix(num_pos_feats=64): fourier_embedding.py:66-67
        sections['dim_t64'] = (10000 ** (2 * (dt // 2) / 64)).contiguous()
    if f'{ce}.drag_point.pointnet_encoder.pre_mlps.mlp.0.weight' in sd:
        sections['drag_enc'] = pack_pointnet(sd, f'{ce}.drag_point.pointnet_encoder', 1)
        sections['drag_enc_tc'] = pack_pointnet_tc(sd, f'{ce}.drag_point.pointnet_encoder', 1)
    sections['head'] = pack_head(sd)
    # FourierEmbeddingFix denominators, evaluated by torch exactly as the reference does (fourier_embedding.py:66-67)
    dt = torch.arange(128 / 4, dtype=torch.float32)
    sections['dim_t16'] = (10000 ** (2 * (dt // 2) / (128 / 4)))[0::2].contiguous()
    dt = torch.arange(128, dtype=torch.float32)
    sections['dim_t128'] = (10000 ** (2 * (dt // 2) / 128)).contiguous()
    offsets, total = OrderedDict(), 0
    for k, v in sections.items():
        offsets[k] = total
        total += (v.numel() + 63) // 64 * 64
    arena = torch.zeros(total, dtype=torch.float32)
    for k, v in sections.items():
        arena[offsets[k]:offsets[k] + v.numel()] = v
    return arena, offsets
