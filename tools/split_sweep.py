#!/usr/bin/env python
"""Forward time of the bench workload for every row-split setting of the fixed-source attention stacks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prosim_b200 import lib, synthetic, weights
from prosim_b200.model import ProSimB200
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
pristine = synthetic.clone_batch(synthetic.make_batch(n_scenes=32, n_agents=128, n_map=512, steps=80), dev)[0]
ref = None
with torch.no_grad():
    for tc in (True,):
        lib.set_tensor_core(tc)
        for parts in (1, 2, 3, 4):
            lib.set_stack_split(parts)
            ts = []
            for i in range(5):
                b = synthetic.clone_batch(pristine)[0]
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = model.forward(b, 'val'); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            traj = out['motion_pred']['_state']['traj'].clone()
            if parts == 1:
                ref = traj
            print(f'tensor_core={tc} parts={parts} forward_ms={sorted(ts)[2]:.2f} bit_identical_to_parts1={bool(torch.equal(ref, traj))}')
