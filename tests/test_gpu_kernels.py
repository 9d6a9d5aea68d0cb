"""GPU (-m gpu): every C-ABI entry point against the matching piece of the CPU oracle, on seeded inputs.
Tolerances are written per test; integer/index results must be bit exact."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import graph as ograph
from oracle.prosim_oracle import ProSimOracle, fourier_fix, rel_pe_input
from prosim_b200 import weights
from tests.helpers import ALL_COND, edge_set

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from prosim_b200 import lib, ops
    lib.load()
    sd = weights.random_state_dict(0, ALL_COND)
    arena, off = weights.pack_model(sd)
    return dict(ops=ops, sd=sd, arena=arena.cuda(), off=off, oracle=ProSimOracle(sd, ALL_COND))


def _scene_points(g, counts, spread):
    pos, batch = [], []
    for b, n in enumerate(counts):
        pos.append((torch.rand(n, 2, generator=g) - 0.5) * spread)
        batch.append(torch.full((n,), b, dtype=torch.long))
    return torch.cat(pos), torch.cat(batch)


def _seg(counts, extra_counts=None, extra_base=0):
    rows, off, off2 = [], 0, extra_base
    for i, n in enumerate(counts):
        e = extra_counts[i] if extra_counts else 0
        rows.append([off, n, off2 if extra_counts else 0, e])
        off += n
        off2 += e
    return torch.tensor(rows, dtype=torch.int32)


# ------------------------------------------------------------------------------------ pointnet
@pytest.mark.parametrize('tensor_core', [False, True])
@pytest.mark.parametrize('kind', ['obs', 'map'])
def test_pointnet_matches_oracle(ctx, kind, tensor_core):
    ops, orc = ctx['ops'], ctx['oracle']
    g = torch.Generator().manual_seed(3)
    B, N = 3, 37
    if kind == 'obs':
        x = torch.randn(B, N, 11, 24, generator=g)
        pmask = torch.rand(B, N, 11, generator=g) > 0.25
        pmask[0, 0] = False            # a polyline with no valid point
        pmask[1, 3] = True
        mask = pmask[..., None].expand(B, N, 11, 24).clone()
        x[~mask] = float('nan')        # the reference leaves NaN under a False mask
        ref = orc.pointnet('scene_encoder.obs_encoder', torch.nan_to_num(x), pmask, 1, 2)
        key, k = 'obs_enc', 0
    else:
        x = torch.randn(B, N, 19, 11, generator=g)
        pmask = torch.rand(B, N, 19, generator=g) > 0.25
        pmask[2, 5] = False
        mask = pmask
        ref = orc.pointnet('scene_encoder.map_encoder', x, pmask, 3, 2)
        key, k = 'map_enc', 1
    valid = pmask.any(-1).reshape(-1)
    rows = torch.nonzero(valid).reshape(-1).to(torch.int32)
    out = ops.pointnet(k, x.cuda(), mask.cuda(), rows.cuda(), ctx['arena'], ctx['off'][key],
                       tc_off=ctx['off'][key + '_tc'] if tensor_core else None).cpu()
    ref = ref.reshape(B * N, 128)[valid]
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max()
    print(kind, 'tensor_core' if tensor_core else 'ffma', 'max err', float(err), 'max |ref|', float(ref.abs().max()))
    assert err < 2e-5      # fp32 FFMA: different summation order only; tensor cores: 3xTF32 (2^-21 per product)


# ------------------------------------------------------------------------------------ neighbour search
def test_radius_edges_bit_exact(ctx):
    ops = ctx['ops']
    g = torch.Generator().manual_seed(5)
    q_counts, s_counts = [40, 7, 0, 25], [90, 33, 12, 64]
    qpos, qb = _scene_points(g, q_counts, 120.0)
    spos, sb = _scene_points(g, s_counts, 120.0)
    for r, cap in ((50.0, 768), (30.0, 5), (300.0, 16)):
        ref = ograph.radius(spos, qpos, r, sb, qb, max_num_neighbors=cap)       # [2,E]: row0 = query, row1 = source
        stride = min(cap, max(s_counts))
        e = ops.radius_edges(qpos.cuda(), qb.int().cuda(), spos.cuda(), _seg(s_counts).cuda(), r, cap, stride)
        mine = e.to_edge_index()                                                # row0 = source, row1 = dst(query)
        assert torch.equal(mine[1], ref[0]) and torch.equal(mine[0], ref[1]), (r, cap)


def test_radius_graph_drop_self_bit_exact(ctx):
    ops = ctx['ops']
    g = torch.Generator().manual_seed(6)
    counts = [33, 1, 20]
    pos, b = _scene_points(g, counts, 60.0)
    for r, cap in ((300.0, 512), (20.0, 4)):
        ref = ograph.radius_graph(pos, r, b, loop=False, max_num_neighbors=cap)  # row0 = source, row1 = target
        e = ops.radius_edges(pos.cuda(), b.int().cuda(), pos.cuda(), _seg(counts).cuda(), r, cap,
                             min(cap + 1, max(counts)), drop_self=True)
        assert edge_set(e.to_edge_index()) == edge_set(ref)
        assert e.to_edge_index().shape[1] == ref.shape[1]


def test_knn_edges_bit_exact(ctx):
    ops = ctx['ops']
    g = torch.Generator().manual_seed(7)
    m_counts, a_counts = [50, 9, 70], [20, 3, 40]
    mpos, mb = _scene_points(g, m_counts, 200.0)
    apos, ab = _scene_points(g, a_counts, 100.0)
    pos, b = torch.cat([mpos, apos]), torch.cat([mb, ab])
    NM = mpos.shape[0]
    # scene graph: two source segments per scene (map rows, then agent rows), self loops kept
    ref = ograph.knn_graph(pos, 32, b, loop=True)
    seg = _seg(m_counts, a_counts, extra_base=NM)
    nmax = max(m + a for m, a in zip(m_counts, a_counts))
    e = ops.knn_edges(pos.cuda(), b.int().cuda(), pos.cuda(), seg.cuda(), 32, nmax, min(32, nmax))
    assert edge_set(e.to_edge_index()) == edge_set(ref)
    assert e.to_edge_index().shape[1] == ref.shape[1]
    # agent graph with k larger than some scenes
    ref = ograph.knn_graph(apos, 16, ab, loop=True)
    e = ops.knn_edges(apos.cuda(), ab.int().cuda(), apos.cuda(), _seg(a_counts).cuda(), 16, max(a_counts), 16)
    assert edge_set(e.to_edge_index()) == edge_set(ref)


# ------------------------------------------------------------------------------------ relative PE
def _random_graph(g, n_dst, n_src, stride, empty_row=True):
    deg = torch.randint(1, min(stride, n_src) + 1, (n_dst,), generator=g)
    if empty_row:
        deg[1] = 0
    nbr = torch.zeros(n_dst, stride, dtype=torch.long)
    for i in range(n_dst):
        nbr[i, :deg[i]] = torch.sort(torch.randperm(n_src, generator=g)[:deg[i]])[0]
    j = torch.arange(stride)[None, :] < deg[:, None]
    edge_index = torch.stack([nbr[j], torch.arange(n_dst)[:, None].expand(n_dst, stride)[j]])
    return nbr, deg, j, edge_index


def test_edge_pe_matches_oracle(ctx):
    ops = ctx['ops']
    g = torch.Generator().manual_seed(9)
    n_dst, n_src, stride = 50, 80, 24
    nbr, deg, j, ei = _random_graph(g, n_dst, n_src, stride)
    dpos, spos = (torch.rand(n_dst, 2, generator=g) - 0.5) * 300, (torch.rand(n_src, 2, generator=g) - 0.5) * 300
    dori, sori = (torch.rand(n_dst, generator=g) - 0.5) * 2 * math.pi, (torch.rand(n_src, generator=g) - 0.5) * 2 * math.pi
    spos[nbr[0, 0]] = dpos[0]        # a zero-offset edge: atan2(0, 0) = 0 must be reproduced
    pe = fourier_fix(rel_pe_input(ei, dori[:, None], dpos, sori[:, None], spos), 128 / 4)
    ref = F.layer_norm(pe, (128,))
    e = ops.EdgeList(nbr.reshape(-1).int().cuda(), deg.int().cuda(), stride, stride)
    dim_t = ctx['arena'][ctx['off']['dim_t16']:ctx['off']['dim_t16'] + 16]
    ops.edge_pe(e, dpos.cuda(), dori.cuda(), spos.cuda(), sori.cuda(), dim_t)
    assert e.zd == 96                 # features 96..127 repeat 64..95 and are not stored
    z = e.z.view(n_dst, stride, 96).cpu()[j]
    ref = ref[:, :96]
    # sin/cos arguments reach ~2000 rad where one fp32 ulp of the argument is 1.2e-4: the kernel rounds |dp| exactly
    # like torch.norm does, so only atan2 / sin / cos implementation differences (<= 2 ulp) remain
    assert (z - ref).abs().max() < 2e-5


# ------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize('bipartite', [True, False])
@pytest.mark.parametrize('n_dst', [23, 300, 1100, 7700])
def test_attention_layer_matches_oracle(ctx, bipartite, n_dst):
    ops, orc = ctx['ops'], ctx['oracle']
    g = torch.Generator().manual_seed(11 + n_dst)
    n_src = n_dst if not bipartite else 2 * n_dst + 5
    stride = 40
    nbr, deg, j, ei = _random_graph(g, n_dst, n_src, stride)
    x_src = torch.randn(n_src, 128, generator=g)
    x_dst = x_src if not bipartite else torch.randn(n_dst, 128, generator=g)
    r = torch.randn(ei.shape[1], 128, generator=g)
    prefix = 'policy.act_decoder.m2p_attn_layers.2' if bipartite else 'decoder.p2p_attn_layers.4'
    key = ('pol_m2p', 2) if bipartite else ('dec_p2p', 4)
    ref = orc.attention_layer(prefix, x_src, x_dst, r, ei, bipartite)
    z = torch.zeros(n_dst, stride, 128)
    z[j] = F.layer_norm(r, (128,))
    e = ops.EdgeList(nbr.reshape(-1).int().cuda(), deg.int().cuda(), stride, stride, z.reshape(-1, 128).cuda())
    xs = x_src.cuda()
    xd = xs if not bipartite else x_dst.cuda()
    out = ops.attn_layer(xs, xd, e, ctx['arena'], ctx['off'][key[0]] + key[1] * weights.ATTN_LAYER_FLOATS).cpu()
    assert (out - ref).abs().max() < 2e-5


@pytest.mark.parametrize('stride,n_dst', [(768, 70), (37, 1300), (203, 1100)])
def test_attention_layer_edge_tiling(ctx, stride, n_dst):
    """Edge kernel tiling corners: the largest supported neighbour list (768 = 24 z tiles per row), strides that are not
    a multiple of 8 (partial TMA boxes run past the row's list / the end of z), rows without edges; small launches
    take the FFMA node kernel, >= 1024 rows the tensor-core one."""
    ops, orc = ctx['ops'], ctx['oracle']
    g = torch.Generator().manual_seed(100 + stride)
    n_src = stride + 150
    nbr, deg, j, ei = _random_graph(g, n_dst, n_src, stride)
    deg_full = min(stride, n_src)
    nbr[0, :deg_full] = torch.sort(torch.randperm(n_src, generator=g)[:deg_full])[0]   # one row with a full list
    deg[0] = deg_full
    j = torch.arange(stride)[None, :] < deg[:, None]
    ei = torch.stack([nbr[j], torch.arange(n_dst)[:, None].expand(n_dst, stride)[j]])
    x_src, x_dst = torch.randn(n_src, 128, generator=g), torch.randn(n_dst, 128, generator=g)
    r = torch.randn(ei.shape[1], 128, generator=g)
    r[:, 96:] = r[:, 64:96]                       # 96-wide z path: features 96..127 duplicate 64..95
    ref = orc.attention_layer('policy.act_decoder.m2p_attn_layers.1', x_src, x_dst, r, ei, True)
    z = torch.zeros(n_dst, stride, 96)
    z[j] = F.layer_norm(r, (128,))[:, :96]
    e = ops.EdgeList(nbr.reshape(-1).int().cuda(), deg.int().cuda(), stride, stride, z.reshape(-1, 96).cuda(), zd=96)
    out = ops.attn_layer(x_src.cuda(), x_dst.cuda(), e, ctx['arena'],
                         ctx['off']['pol_m2p'] + 1 * weights.ATTN_LAYER_FLOATS).cpu()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() < 2e-5


@pytest.mark.parametrize('tensor_core', [True, 7, False])
def test_attention_stack_matches_oracle(ctx, tensor_core):
    """6 x (a2p, m2p) with fixed sources -- the policy tick's core -- and 3 x self-source layers; with the tcgen05 node
    kernel (3xTF32, activations in TMEM) and with the FFMA node kernels."""
    from prosim_b200 import lib
    lib.set_tensor_core(tensor_core)
    try:
        # True: the 32-row swapped tcgen05 node kernel (post_sw.cuh); 7: the 128-row one (tc_post.cuh); False: FFMA
        _attention_stack_case(ctx, 1e-4 if tensor_core else 5e-5, 4e-5 if tensor_core else 2e-5)
    finally:
        lib.set_tensor_core(True)


def _attention_stack_case(ctx, tol12, tol3):
    ops, orc = ctx['ops'], ctx['oracle']
    g = torch.Generator().manual_seed(13)
    P, NA, NM, sa, sm = 1100, 1064, 8150, 30, 50    # >= 1024 rows: the gemm_tile (v2) kernels run
    nbr_a, deg_a, ja, ei_a = _random_graph(g, P, NA, sa)
    nbr_m, deg_m, jm, ei_m = _random_graph(g, P, NM, sm, empty_row=False)
    x_p, x_a, x_m = torch.randn(P, 128, generator=g), torch.randn(NA, 128, generator=g), torch.randn(NM, 128, generator=g)
    r_a, r_m = torch.randn(ei_a.shape[1], 128, generator=g), torch.randn(ei_m.shape[1], 128, generator=g)
    ref = x_p
    for i in range(6):
        ref = orc.attention_layer(f'policy.act_decoder.a2p_attn_layers.{i}', x_a, ref, r_a, ei_a, True)
        ref = orc.attention_layer(f'policy.act_decoder.m2p_attn_layers.{i}', x_m, ref, r_m, ei_m, True)

    def edges(nbr, deg, j, r, stride):
        z = torch.zeros(P, stride, 128)
        z[j] = F.layer_norm(r, (128,))
        return ops.EdgeList(nbr.reshape(-1).int().cuda(), deg.int().cuda(), stride, stride, z.reshape(-1, 128).cuda())

    ar, off, lf = ctx['arena'], ctx['off'], weights.ATTN_LAYER_FLOATS
    e_a, e_m = edges(nbr_a, deg_a, ja, r_a, sa), edges(nbr_m, deg_m, jm, r_m, sm)
    kv_a = ops.attn_kv(x_a.cuda(), ar, off['pol_a2p'], 6, lf)
    kv_m = ops.attn_kv(x_m.cuda(), ar, off['pol_m2p'], 6, lf)
    out = ops.attn_stack(x_p.cuda(), 6, ops.stack_side(ar, off['pol_a2p'], e_a, kv_a),
                         ops.stack_side(ar, off['pol_m2p'], e_m, kv_m)).cpu()
    err = float((out - ref).abs().max())
    print('12-layer stack max err', err, 'of max |ref|', float(ref.abs().max()))
    assert err < tol12       # 12 layers deep (measured on B200: 5.6e-5 with 3xTF32 tensor-core GEMMs, below 5e-5 with FFMA)

    nbr_s, deg_s, js, ei_s = _random_graph(g, P, P, sa)
    r_s = torch.randn(ei_s.shape[1], 128, generator=g)
    ref = x_p
    ct = 'condition_transformers.policy_decoder.condition_attn.attn_layers'
    for i in range(3):
        ref = orc.attention_layer(f'{ct}.{i}', ref, ref, r_s, ei_s, False)
    out = ops.attn_stack(x_p.cuda(), 3, ops.stack_side(ar, off['cond_attn'], edges(nbr_s, deg_s, js, r_s, sa))).cpu()
    assert (out - ref).abs().max() < tol3


def test_tensor_core_chunk_pipeline_is_race_free(ctx):
    """The tcgen05 PointNet / K'|V' kernels reuse shared memory between generic-proxy tiles and async-proxy (bulk copy,
    MMA) operands.  A first version of attn_kv_tc_kernel issued a weight copy over rows other threads were still loading:
    ~1 % of the launches returned a few corrupted rows.  Repeat both kernels a few hundred times: every result must be
    bit-identical to the first."""
    ops, ar, off = ctx['ops'], ctx['arena'], ctx['off']
    g = torch.Generator().manual_seed(29)
    x_m = torch.randn(8150, 128, generator=g).cuda()
    ref = None
    for _ in range(400):
        kv = ops.attn_kv(x_m, ar, off['pol_m2p'], 6, weights.ATTN_LAYER_FLOATS)
        if ref is None:
            ref = kv.clone()
        else:
            assert torch.equal(kv, ref)
    x = torch.randn(1200, 11, 24, generator=g).cuda()
    mask = (torch.rand(1200, 11, 1, generator=g) > 0.2).expand(1200, 11, 24).contiguous().cuda()
    rows = torch.arange(1200, dtype=torch.int32).cuda()
    ref = None
    for _ in range(200):
        out = ops.pointnet(0, x, mask, rows, ar, off['obs_enc'], tc_off=off['obs_enc_tc'])
        if ref is None:
            ref = out.clone()
        else:
            assert torch.equal(out, ref)


# ------------------------------------------------------------------------------------ heads
@pytest.mark.parametrize('P', [45, 1024, 4100])
def test_policy_head_and_reconst_match_oracle(ctx, P):
    """P = 45: policy_head_kernel (per-thread weight loads); P >= 1024: policy_head2_kernel (weights through the
    shared-memory stream), the kernel every launch of the bench workload takes."""
    ops, orc = ctx['ops'], ctx['oracle']
    g = torch.Generator().manual_seed(17)
    feat, emd = torch.randn(P, 128, generator=g), torch.randn(P, 128, generator=g)
    a_type = torch.randint(1, 4, (P,), generator=g)
    ref = orc.policy_head(feat, a_type, emd)
    mp = ops.policy_head(feat.cuda(), a_type.int().cuda(), ctx['arena'], ctx['off']['head']).cpu()
    rc = ops.reconst(emd.cuda(), ctx['arena'], ctx['off']['head']).cpu()
    assert (mp - ref['motion_pred']).abs().max() < 2e-5
    assert (rc - ref['reconst_pred']).abs().max() < 1e-5


def test_prompt_and_goal_encoders_match_oracle(ctx):
    ops, orc = ctx['ops'], ctx['oracle']
    g = torch.Generator().manual_seed(19)
    n = 31
    x = torch.randn(n, 7, generator=g)
    ref = orc.mlp('prompt_encoder.motion_pred.state_encoder', x, 2, ret_before_act=True)
    out = ops.mlp2(x.cuda(), 7, True, ctx['arena'], ctx['off']['prompt_mlp']).cpu()
    assert (out - ref).abs().max() < 1e-5
    gi = torch.cat([torch.randn(n, 2, generator=g) * 40, torch.full((n, 1), 80.0), torch.zeros(n, 1)], dim=1)
    ct = 'condition_transformers.policy_decoder.condition_encoders.goal.goal_encoder'
    ref = orc.mlp(ct, gi[:, :2], 2, ret_before_act=True, without_norm=True) + fourier_fix(gi[:, 2:3], 128)
    dim_t128 = ctx['arena'][ctx['off']['dim_t128']:ctx['off']['dim_t128'] + 128]
    out = ops.mlp2(gi.cuda(), 2, False, ctx['arena'], ctx['off']['goal_mlp'], tpe_col=2, dim_t128=dim_t128).cpu()
    assert (out - ref).abs().max() < 2e-5


@pytest.mark.parametrize('T', [16, 8])
def test_condition_encoders_and_pooling_match_oracle(ctx, T):
    """Action-tag / drag-point encoders and the per-agent mean pooling (condition_encoders.py:76-191,
    condition_attns.py:114-189) against the oracle's encode_conditions / edge matrix."""
    ops, orc, ar, off = ctx['ops'], ctx['oracle'], ctx['arena'], ctx['off']
    g = torch.Generator().manual_seed(11)
    B, N, C = 3, 9, 14
    tags = torch.stack([torch.randint(-1, 11, (B, C), generator=g), torch.randint(0, 80, (B, C), generator=g),
                        torch.randint(0, 200, (B, C), generator=g)], dim=-1)
    tags[tags[..., 0] < 0] = -1
    tmask = (tags != -1).all(-1)
    # every (agent, tag name) pair at most once per scene, like the dataset
    tidx = torch.zeros(B, C, 1, dtype=torch.long)
    for b in range(B):
        seen = set()
        for c in range(C):
            n = int(torch.randint(0, N, (1,), generator=g))
            while (n, int(tags[b, c, 0])) in seen:
                n = (n + 1) % N
            seen.add((n, int(tags[b, c, 0])))
            tidx[b, c, 0] = n if tmask[b, c] else -1
    drag = torch.randn(B, N, T, 2, generator=g) * 20
    drag[torch.rand(B, N, T, generator=g) < 0.4] = float('nan')
    drag[0, 1] = float('nan')                                        # an agent without any drag point
    dmask = ~drag.isnan().all(-1).all(-1)
    didx = torch.arange(N).view(1, N, 1).expand(B, N, 1).clone()
    didx[~dmask] = -1
    goal = torch.cat([torch.randn(B, N, 2, generator=g) * 50, torch.full((B, N, 1), 80.0)], -1)
    gmask = torch.rand(B, N, generator=g) < 0.6
    gidx = torch.arange(N).view(1, N, 1).expand(B, N, 1).clone()
    gidx[~gmask] = -1
    pm = torch.ones(B, N, dtype=torch.bool)
    cond = {'goal': dict(input=goal, mask=gmask, prompt_idx=gidx, prompt_mask=pm),
            'v_action_tag': dict(input=tags, mask=tmask, prompt_idx=tidx, prompt_mask=pm),
            'drag_point': dict(input=drag, mask=dmask, prompt_idx=didx, prompt_mask=pm)}
    emds = orc.encode_conditions(cond)
    # --- encoders, entry by entry
    t_gpu = ops.tag_embed(tags.reshape(-1, 3).cuda(), ar[off['tag_vec']:off['tag_vec'] + 16 * 128],
                          ar[off['dim_t64']:off['dim_t64'] + 64]).cpu().view(B, C, 128)
    assert (t_gpu[tags[..., 0] < 0] == 0).all()
    for tag, tid in weights.V_ACTION_TAG_ID.items():
        sel = tags[..., 0] == tid
        if tag in emds:
            ref = emds[tag]['emd'][torch.arange(emds[tag]['emd'].shape[1])[None, :] < sel.sum(1)[:, None]]
            assert (t_gpu[sel] - ref).abs().max() < 2e-5      # sin/cos of arguments up to 2 pi 200
    d_gpu = ops.pointnet(2 if T == 16 else 3, drag.reshape(B * N, T, 2).cuda(), None,
                         torch.arange(B * N, dtype=torch.int32).cuda(), ar, off['drag_enc']).cpu().view(B, N, 128)
    assert torch.isfinite(d_gpu).all()
    assert (d_gpu - emds['drag_point']['emd']).abs().max() < 2e-5
    # --- pooling: build the slot table the way model._conditions does, compare with the oracle's pooled edge attribute
    emd = torch.zeros(B, N, 128)
    x_ref = orc.condition_attn(cond, emd, pm, torch.zeros(B, N, 2), torch.zeros(B, N, 1))
    attr, has = orc._dbg_cond['attr'].view(B * N, 128), orc._dbg_cond['has'].view(B * N)
    from prosim_b200.config import get_config
    from prosim_b200.model import ProSimB200
    from types import SimpleNamespace
    model = ProSimB200(get_config(opts=['PROMPT.CONDITION.TYPES', list(ALL_COND)]), ctx['sd'], device='cuda')
    P = B * N
    pl = SimpleNamespace(P=P, B=B, N=N, i={'prow_lut': torch.arange(P, dtype=torch.int32).cuda()})
    captured = {}
    real_pool = ops.cond_pool

    def spy(e, s):
        captured['out'] = real_pool(e, s)
        return captured['out']
    ops.cond_pool = spy
    try:
        cc = {k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in cond.items()}
        p_pos, p_ori = torch.zeros(P, 2).cuda(), torch.zeros(P).cuda()
        ws = ops.attn_workspace(P, P, 'cuda', 1)
        x_gpu = model._conditions(cc, emd.view(P, 128).cuda(), p_pos, p_ori, pl, ws).cpu()
    finally:
        ops.cond_pool = real_pool
    extra, has_gpu = captured['out']
    assert torch.equal(has_gpu.cpu().bool(), has)
    assert (extra.cpu() - attr).abs().max() < 2e-5
    assert (x_gpu - x_ref.view(P, 128)).abs().max() < 5e-5


def test_bad_arguments_raise(ctx):
    from prosim_b200 import lib
    ops = ctx['ops']
    with pytest.raises(lib.ProSimLibError):
        ops.pointnet(0, torch.zeros(1, 11, 24), torch.ones(1, 11, 24, dtype=torch.bool), torch.zeros(1, dtype=torch.int32),
                     ctx['arena'], 0)        # CPU tensors: no fallback
    with pytest.raises(lib.ProSimLibError):
        lib.call('prosim_step_agent_traj', None, None, 4, 91, 90, None, None, None)   # tidx + 10 > T


@pytest.mark.parametrize('split3', [0, 1])
def test_tcgen05_gemm_block(ctx, split3):
    """tcgen05 / TMEM building block: one TF32 pass is ~1e-3 accurate, the 3xTF32 scheme is fp32-class."""
    from prosim_b200 import lib
    g = torch.Generator().manual_seed(23)
    m = 300
    a, w = torch.randn(m, 128, generator=g), torch.randn(128, 128, generator=g) / 11.3
    c = torch.full((m, 128), float('nan'), device='cuda')
    a_d, w_d = a.cuda(), w.cuda()
    lib.call('prosim_tc_gemm_test', lib.ptr(a_d), lib.ptr(w_d), lib.ptr(c), m, split3,
             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = (a.double() @ w.double().t()).float()
    err = (c.cpu() - ref).abs().max().item()
    print('tcgen05 gemm split3 =', split3, 'max err', err)
    assert err < (2e-5 if split3 else 2e-2)
