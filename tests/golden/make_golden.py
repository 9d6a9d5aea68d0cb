"""Generate the committed golden vectors by running the UNMODIFIED reference (oracle/ref_shim.py).

Run once in the authoring container (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/<case>.npz (+ state_dict_keys.json).  For every case of BASELINE.json's
configs 1-4, two ragged batches and two batches with mixed goal / action-tag / drag-point conditions: the reference's fp32 rollout (traj, vel, motion_pred,
reconst_pred, init_pos/heading, pair_names) and the reference's own fp64 evaluation of the same
inputs/weights (traj64), the arbiter for closed-loop rounding disputes.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from prosim_b200 import synthetic, weights  # noqa: E402

from tests.helpers import BENCH_CASES, CASES, DEMO_CASES, cond_suffix, to_double as _to_double  # noqa: E402


def _collect(out, ids):
    names = [f'{b}-{a}' for b, row in enumerate(ids) for a in row]
    rt = out['rollout_trajs']
    assert list(rt.keys()) == names
    return dict(
        traj=torch.stack([rt[n]['traj'] for n in names]).numpy(),
        vel=torch.stack([rt[n]['vel'] for n in names]).numpy(),
        init_pos=torch.stack([rt[n]['init_pos'] for n in names]).numpy(),
        init_heading=torch.stack([rt[n]['init_heading'] for n in names]).numpy(),
        motion_pred=out['motion_pred'].numpy(), motion_prob=out['motion_prob'].numpy(),
        reconst_pred=out['reconst_pred'].numpy(), pair_names=np.array(out['pair_names']),
        agent_names=np.array(names))


def main():
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]                      # optional: regenerate just the named cases
    keep = {n: c[2] for n, c in BENCH_CASES.items()}
    cases = {n: c[:2] for n, c in list(CASES.items()) + list(BENCH_CASES.items()) if not only or n in only}
    demo = {n: kw for n, kw in DEMO_CASES.items() if not only or n in only}
    if demo:
        cases.update({n: (None, False) for n in demo})
    models = {}
    for goal in dict.fromkeys(c[1] for c in cases.values()):
        sd = weights.random_state_dict(0, goal)
        m32, _ = ref_shim.build_reference_model(weights.cond_types(goal))
        m32.load_state_dict(sd)
        m64, _ = ref_shim.build_reference_model(weights.cond_types(goal), dtype=torch.float64)
        m64.load_state_dict({k: v.double() for k, v in sd.items()})
        models[goal] = (m32, m64)
        with open(os.path.join(HERE, f'state_dict_keys{cond_suffix(goal)}.json'), 'w') as f:
            json.dump({'keys': [[k, list(v.shape)] for k, v in m32.state_dict().items()],
                       'checksum': float(sum(v.double().sum() for v in sd.values())),
                       'abs_checksum': float(sum(v.double().abs().sum() for v in sd.values()))}, f)
    for name, (kw, goal) in cases.items():
        m32, m64 = models[goal]
        if name in demo:      # a real scene of the reference's demo_dataset through the mini-loader
            from prosim_b200 import demo_loader
            make = lambda: demo_loader.load_demo_scene(os.path.join(ref_shim.REF_ROOT, 'demo_dataset'), **demo[name])
        else:
            make = lambda: synthetic.make_batch(**kw)
        b32 = make()
        ids = b32.extras['prompt']['motion_pred']['agent_ids']
        with torch.no_grad():
            out32 = m32.forward(b32, 'val')['motion_pred']
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            with torch.no_grad():
                out64 = m64.forward(_to_double(make()), 'val')['motion_pred']
        finally:
            torch.set_default_dtype(prev)
        res = _collect(out32, ids)
        r64 = _collect(out64, ids)
        res['traj64'], res['vel64'], res['motion_pred64'] = r64['traj'], r64['vel'], r64['motion_pred']
        gap = np.abs(res['traj'][..., :2].astype(np.float64) - res['traj64'][..., :2]).reshape(len(res['traj']), -1, 10, 2)
        print(name, 'fp32-vs-fp64 xy gap per tick:', ['%.1e' % g for g in gap.max(axis=(0, 2, 3))])
        if name in demo:
            from prosim_b200 import demo_loader
            res.update({'batch.' + k: v for k, v in demo_loader.batch_to_arrays(make()).items()})
        if name in keep:                     # large batch: keep the rows of a few scenes of THIS run (file size)
            scenes = [int(n.split('-')[0]) for n in res['agent_names']]
            rows = np.array([i for i, sc in enumerate(scenes) if sc in keep[name]])
            P = len(scenes)
            pair_rows = np.concatenate([rows + k * P for k in range(len(res['pair_names']) // P)])
            for k in ('traj', 'vel', 'init_pos', 'init_heading', 'agent_names', 'traj64', 'vel64'):
                res[k] = res[k][rows]
            for k in ('motion_pred', 'motion_prob', 'reconst_pred', 'pair_names', 'motion_pred64'):
                res[k] = res[k][pair_rows]
            res['rows'], res['n_rows'] = rows, np.array(P)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **res)


if __name__ == '__main__':
    main()
