"""CPU: host-side logic that needs no GPU -- the C-ABI library loads and exports every declared symbol,
the weight packer matches the library's layout, config / registry mirrors behave like the reference's."""
import ctypes
import os
import re

import pytest
import torch

from prosim_b200 import config, lib, synthetic, weights
from prosim_b200.registry import registry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'prosim_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(prosim_[a-z0-9_]+)\s*\(', header)))
    assert declared, 'no declarations found'
    l = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(l, s)]
    assert not missing, missing
    assert sorted(lib.SYMBOLS) == declared


def test_library_layout_matches_packer():
    l = lib.load()
    assert l.prosim_attn_layer_floats() == weights.ATTN_LAYER_FLOATS
    assert l.prosim_pointnet_floats() == weights.POINTNET_FLOATS
    assert l.prosim_head_floats() == weights.HEAD_FLOATS
    assert l.prosim_mlp2_floats() == weights.MLP2_FLOATS
    assert l.prosim_attn_workspace_floats(10, 0, 32) >= 10 * (2 * (128 + 1024 + 256) + 1024 + 128 + 256 + 16 * 32)


def test_tick_workspace_layout():
    """prosim_workspace_bytes(cfg): monotone in every size, 256-byte granular, rejects nonsense (no compute, no GPU)."""
    from prosim_b200 import ops
    base = dict(P=4096, n_agent=4096, n_map=16384, max_a=128, max_m=512, max_neigh=768, n_layers=6)
    b0 = ops.tick_workspace_bytes(ops.tick_cfg(**base))
    # dominated by z: P x (128 + 512) x 96 floats
    assert b0 % 256 == 0 and b0 > 4096 * (128 + 512) * 96 * 4
    for k in base:
        bigger = dict(base, **{k: base[k] * 2})
        assert ops.tick_workspace_bytes(ops.tick_cfg(**bigger)) >= b0
    assert ops.tick_workspace_bytes(ops.tick_cfg(**dict(base, max_neigh=0))) == 0
    # strides are capped by MAX_NUM_NEIGH
    capped = ops.tick_workspace_bytes(ops.tick_cfg(**dict(base, max_m=100000)))
    assert capped == ops.tick_workspace_bytes(ops.tick_cfg(**dict(base, max_m=768)))
    import prosim_b200.model  # noqa: F401  (registers the sub-modules)
    assert registry.get_scene_encoder('attn_fusion_relpe_b200') is not None and registry.get_policy('rel_pe_temporal_b200') is not None


def test_pack_model_sections():
    for goal in (False, True):
        arena, off = weights.pack_model(weights.random_state_dict(0, goal))
        assert ('cond_attn' in off) == goal and ('goal_mlp' in off) == goal
        assert all(o % 64 == 0 for o in off.values())
        assert torch.isfinite(arena).all()
        d16 = arena[off['dim_t16']:off['dim_t16'] + 16]
        assert d16[0] == 1.0 and abs(float(d16[15]) - 10000 ** (15 / 16)) < 1.0


def test_attn_fold_is_algebraically_exact():
    """The packed (folded) weights reproduce k_e / v_e of the reference formulation in fp64."""
    sd = {k: v.double() for k, v in weights.random_state_dict(1).items()}
    p = 'policy.act_decoder.a2p_attn_layers.3'
    pk = weights.pack_attn_layer(sd, p).double()
    g = torch.Generator().manual_seed(0)
    z = torch.randn(5, 128, generator=g, dtype=torch.float64)
    xs = torch.randn(5, 128, generator=g, dtype=torch.float64)
    r_hat = z * sd[f'{p}.attn_prenorm_r.weight'] + sd[f'{p}.attn_prenorm_r.bias']
    k_e = xs @ sd[f'{p}.to_k.weight'].t() + r_hat @ sd[f'{p}.to_k_r.weight'].t()
    v_e = xs @ sd[f'{p}.to_v.weight'].t() + sd[f'{p}.to_v.bias'] + r_hat @ sd[f'{p}.to_v_r.weight'].t() + sd[f'{p}.to_v_r.bias']
    o = 512
    wqt, o = pk[o:o + 16384].view(128, 128), o + 16384 + 128
    wkt, kb = pk[o:o + 16384].view(128, 128), pk[o + 16384:o + 16384 + 128]
    o += 16384 + 128
    wvt, vb = pk[o:o + 16384].view(128, 128), pk[o + 16384:o + 16384 + 128]
    o += 16384 + 128
    wkrg = pk[o:o + 16384].view(128, 128)
    wvrgt = pk[o + 16384:o + 2 * 16384].view(128, 128)
    assert torch.allclose(xs @ wkt + kb + z @ wkrg.t(), k_e, atol=1e-6)
    assert torch.allclose(xs @ wvt + vb + z @ wvrgt, v_e, atol=1e-6)
    assert torch.allclose(wqt, 0.25 * sd[f'{p}.to_q.weight'].t().float().double(), atol=1e-7)


def test_config_and_registry_mirror():
    cfg = config.get_config(opts=['PROMPT.CONDITION.TYPES', ['goal']])
    assert cfg.DATASET.FORMAT.TARGET.ELEMENTS == 'x,y,h,xd,yd'
    assert cfg.MODEL.POLICY.ACT_DECODER.ATTN.MAX_NUM_NEIGH == 768 and cfg.MODEL.DECODER.ATTN.SCENE_RADIUS == 300
    config.check_supported(cfg)
    bad = config.get_config(opts=['MODEL.REL_POS_EDGE_FUNC', 'knn'])
    with pytest.raises(NotImplementedError):
        config.check_supported(bad)
    from prosim_b200.model import ProSimB200
    assert registry.get_model('prosim_b200') is ProSimB200
    assert registry.get_model('nope') is None


def test_synthetic_batch_layout():
    b = synthetic.make_batch(agents_per_scene=[5, 3], map_per_scene=[7, 9], steps=30, goal=True, permute_obs=True)
    ex = b.extras
    assert ex['init_obs']['input'].shape == (2, 5, 11, 24) and ex['init_map']['input'].shape == (2, 9, 19, 11)
    assert ex['init_obs']['mask'][1, 3:].sum() == 0 and torch.isnan(ex['init_obs']['input'][1, 3:]).all()
    assert sorted(ex['fut_obs'].keys()) == [10, 20] and ex['all_t_indices'].tolist() == [0, 10, 20]
    assert ex['condition']['goal']['input'].shape == (2, 5, 3)
    assert sorted(ex['init_obs']['agent_ids'][0]) == sorted(ex['prompt']['motion_pred']['agent_ids'][0])
    again = synthetic.make_batch(agents_per_scene=[5, 3], map_per_scene=[7, 9], steps=30, goal=True, permute_obs=True)
    assert torch.equal(torch.nan_to_num(again.extras['init_obs']['input']), torch.nan_to_num(ex['init_obs']['input']))


def test_synthetic_inputs_reproduce_the_checksums_recorded_with_the_goldens():
    """The golden-compared GPU tests build their inputs through helpers.make_case_batch, which insists on the checksum recorded
    when the golden was generated (tests/golden/input_checksums.json).  Here: the generator reproduces every one of them on this
    host, twice in a row, and the guarded map geometry equals a plain evaluation."""
    import torch
    from prosim_b200 import synthetic
    from tests.helpers import BENCH_CASES, CASES, batch_sha1, make_case_batch
    for name in list(CASES) + list(BENCH_CASES):
        a = make_case_batch(name)
        assert batch_sha1(a) == batch_sha1(make_case_batch(name)), name
    g = torch.Generator().manual_seed(7)
    centre = torch.rand(40, 2, generator=g) * 300.0 - 150.0
    mhead = torch.rand(40, generator=g) * 6.0 - 3.0
    kappa = (torch.rand(40, generator=g) - 0.5) * 0.08
    for x, y in zip(synthetic._map_geometry(centre, mhead, kappa), synthetic._map_geometry_checked(centre, mhead, kappa)):
        assert torch.equal(x, y)
