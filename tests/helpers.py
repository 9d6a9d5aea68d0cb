"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

ALL_COND = ('goal', 'v_action_tag', 'drag_point')      # PROMPT.CONDITION.TYPES of the released prosim_demo/cfg/no_text.yaml

# name -> (synthetic.make_batch kwargs, condition types of the model: False / True (= goal only) / tuple of types)
CASES = {
    'cfg1_a16_m256_s20': (dict(n_scenes=1, n_agents=16, n_map=256, steps=20), False),
    'cfg2_a64_m256_s40': (dict(n_scenes=1, n_agents=64, n_map=256, steps=40), False),
    'cfg3_a128_m512_s80': (dict(n_scenes=1, n_agents=128, n_map=512, steps=80), False),
    'cfg4_goal_a128_m512_s80': (dict(n_scenes=1, n_agents=128, n_map=512, steps=80, goal=True), True),
    'ragged_b3_s30': (dict(agents_per_scene=[24, 9, 17], map_per_scene=[40, 64, 33], steps=30,
                           permute_obs=True), False),
    'ragged_goal_b2_s20': (dict(agents_per_scene=[12, 20], map_per_scene=[48, 30], steps=20, goal=True,
                                permute_obs=True), True),
    'mixed_cond_a64_m256_s40': (dict(n_scenes=1, n_agents=64, n_map=256, steps=40, goal=True, tags=True, drag=True),
                                ALL_COND),
    'ragged_mixed_b2_s20': (dict(agents_per_scene=[12, 20], map_per_scene=[48, 30], steps=20, goal=True, tags=True,
                                 drag=True, permute_obs=True), ALL_COND),
}


# The benchmarked shape (bench.py: scenes x 128 agents x 512 polylines x 80 steps): 8 scenes = 1024 policy rows, so every
# launch takes the kernels the bench workload runs (tcgen05 node kernels, policy_head2_kernel).  The golden holds the rows of
# KEEP_SCENES of the reference's own B = 8 run (fp32 and fp64): name -> (make_batch kwargs, condition types, kept scenes)
BENCH_CASES = {
    'bench_b8_a128_m512_s80': (dict(n_scenes=8, n_agents=128, n_map=512, steps=80), False, (0, 3, 7)),
}


# BASELINE configs[0], literal: a scene of the reference's demo_dataset (prosim_b200/demo_loader.py), 16 agents, 20 steps.
# The fixture holds the loaded batch itself (the GPU box has no demo_dataset) next to the reference's rollout of it.
DEMO_CASES = {'cfg1_demo_scene6_a16_s20': dict(scene='scene_6', ts=10, steps=20, max_agents=16)}


def demo_batch(name):
    from prosim_b200 import demo_loader
    g = load_golden(name)
    return demo_loader.batch_from_arrays({k[len('batch.'):]: v for k, v in g.items() if k.startswith('batch.')})


def cond_suffix(cond):
    """File-name suffix of the per-model fixtures."""
    return '' if not cond else '_goal' if cond is True else '_' + '_'.join(cond)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def stack_rollout(out):
    """rollout_trajs dict -> (names, traj [P,steps,4], vel [P,steps,2]) as float64 numpy on the host."""
    rt = out['rollout_trajs']
    names = list(rt.keys())
    traj = torch.stack([rt[n]['traj'] for n in names]).detach().cpu().double().numpy()
    vel = torch.stack([rt[n]['vel'] for n in names]).detach().cpu().double().numpy()
    return names, traj, vel


def per_tick_max(a, b, step=10):
    """max |a-b| per 10-step tick over agents and coordinates."""
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    P, S = d.shape[:2]
    return d.reshape(P, S // step, step, -1).max(axis=(0, 2, 3))


def edge_set(edge_index):
    e = edge_index.detach().cpu().long()
    return set(zip(e[1].tolist(), e[0].tolist()))


def to_double(batch):
    """Cast every float tensor of a synthetic batch to float64 (fp64 arbiter runs)."""
    ex = batch.extras
    for key in ('init_obs', 'init_map'):
        for k in ('input', 'position', 'heading'):
            ex[key][k] = ex[key][k].double()
    for t in ex['fut_obs'].keys():
        for k in ('input', 'position', 'heading'):
            ex['fut_obs'][t][k] = ex['fut_obs'][t][k].double()
    p = ex['prompt']['motion_pred']
    for k in ('prompt', 'position', 'heading'):
        p[k] = p[k].double()
    for c in ex['condition'].keys():
        if ex['condition'][c]['input'].is_floating_point():       # action tags stay int64
            ex['condition'][c]['input'] = ex['condition'][c]['input'].double()
    return batch


def batch_sha1(batch):
    """Checksum of every tensor of a (CPU) synthetic batch, in a fixed traversal order."""
    import hashlib
    h = hashlib.sha1()

    def walk(o):
        if torch.is_tensor(o):
            h.update(o.detach().contiguous().cpu().numpy().tobytes())
        elif isinstance(o, dict):
            for k in sorted(o, key=str):
                walk(o[k])
        elif hasattr(o, '_data'):
            walk(o._data)
        elif hasattr(o, 'input') and hasattr(o, 'mask'):
            walk(o.input), walk(o.mask), walk(getattr(o, 'position', None)), walk(getattr(o, 'heading', None))
        elif isinstance(o, (list, tuple)):
            for x in o:
                walk(x)
    walk(batch.extras)
    return h.hexdigest()


def make_case_batch(name):
    """The synthetic input batch of a golden case, verified against the checksum recorded when the golden was generated
    (tests/golden/input_checksums.json): a parity failure must never be a test-input difference in disguise
    (prosim_b200/synthetic.py::_map_geometry_checked has the story)."""
    import json
    from prosim_b200 import synthetic
    kw = (CASES.get(name) or BENCH_CASES[name])[0]
    want = json.load(open(os.path.join(GOLDEN, 'input_checksums.json')))[name]
    for _ in range(3):
        batch = synthetic.make_batch(**kw)
        if batch_sha1(batch) == want:
            return batch
    raise AssertionError(f'{name}: synthetic.make_batch does not reproduce the inputs the golden was generated from '
                         f'(sha1 {batch_sha1(batch)} != {want})')
