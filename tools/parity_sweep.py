#!/usr/bin/env python
"""Closed-loop error of the bench-shape case (8 scenes x 128 x 512 x 80) against the reference's fp64 golden, per
tensor-core kernel selection (prosim_set_tensor_core mask): which kernel family costs how much accuracy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from prosim_b200 import lib, synthetic, weights
from prosim_b200.model import ProSimB200
from tests.helpers import BENCH_CASES, load_golden, per_tick_max, stack_rollout

name = 'bench_b8_a128_m512_s80'
kw, goal, scenes = BENCH_CASES[name]
gold = load_golden(name)
model = ProSimB200(state_dict=weights.random_state_dict(0), device='cuda')
rows = gold['rows']
print('ref32-vs-64 ', ' '.join('%.2e' % x for x in per_tick_max(gold['traj'], gold['traj64'])))
for mask in [int(x) for x in sys.argv[1:]] or [0, 2, 4, 6, 7, 9, 15]:
    lib.set_tensor_core(mask)
    with torch.no_grad():
        out = model.forward(synthetic.make_batch(**kw).to('cuda'), 'val')['motion_pred']
    _, traj, _ = stack_rollout(out)
    print('mask %2d vs 64 ' % mask, ' '.join('%.2e' % x for x in per_tick_max(traj[rows], gold['traj64'])),
          '| vs ref32', ' '.join('%.2e' % x for x in per_tick_max(traj[rows], gold['traj'])))
lib.set_tensor_core(True)
