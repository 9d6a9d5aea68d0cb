#!/usr/bin/env python
"""Repeat a forward N times and count the runs whose trajectories differ from the first one (bit for bit).
Usage: python tools/stress_determinism.py [demo|one|eight] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prosim_b200 import synthetic, weights
from prosim_b200.model import ProSimB200
from tests.helpers import demo_batch
what = sys.argv[1] if len(sys.argv) > 1 else 'demo'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
if what == 'demo':
    make = lambda: demo_batch('cfg1_demo_scene6_a16_s20').to(dev)
else:
    kw = dict(n_scenes=1 if what == 'one' else 8, n_agents=128, n_map=512, steps=20)
    pristine = synthetic.clone_batch(synthetic.make_batch(**kw), dev)[0]
    make = lambda: synthetic.clone_batch(pristine)[0]
ref, bad, worst = None, 0, 0.0
with torch.no_grad():
    for i in range(n):
        t = model.forward(make(), 'val')['motion_pred']['_state']['traj'].clone()
        if ref is None:
            ref = t
        elif not torch.equal(t, ref):
            bad += 1
            worst = max(worst, float((t - ref).abs().max()))
print(f'{what}: {bad} of {n - 1} repeats differ from the first run (max |diff| {worst:.3e}); PROSIM_NO_PDL={os.environ.get("PROSIM_NO_PDL")}')
