#!/usr/bin/env python
"""Uninitialised-memory check: run a forward in clean memory, then fill the caching allocator's free blocks with a poison
pattern (NaN / 1e30 / int garbage) and run the same forward again in a NEW model (fresh workspaces carved from poisoned blocks).
Any difference means some kernel consumes memory it never wrote.  Usage: python tools/poison_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prosim_b200 import synthetic, weights
from prosim_b200.model import ProSimB200
from tests.helpers import demo_batch, CASES
dev = torch.device('cuda', 0)
sd = weights.random_state_dict(0)
cases = {'demo': lambda: demo_batch('cfg1_demo_scene6_a16_s20'),
         'cfg1': lambda: synthetic.make_batch(**CASES['cfg1_a16_m256_s20'][0]),
         'cfg2': lambda: synthetic.make_batch(**CASES['cfg2_a64_m256_s40'][0]),
         'ragged': lambda: synthetic.make_batch(agents_per_scene=[20, 31, 12], map_per_scene=[50, 64, 40], steps=30),
         'eight': lambda: synthetic.make_batch(n_scenes=8, n_agents=128, n_map=512, steps=20)}
def run(make):
    model = ProSimB200(state_dict=sd, device=dev)
    with torch.no_grad():
        st = model.forward(make().to(dev), 'val')['motion_pred']['_state']
    return st['traj'].clone().cpu(), st['vel'].clone().cpu()
import ctypes
from prosim_b200 import lib
def scrub(v):
    lib.call('prosim_debug_scrub', ctypes.c_float(v), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
for name, make in cases.items():
    torch.cuda.empty_cache()
    scrub(0.0)
    ref = run(make)
    for label, val in (('nan', float('nan')), ('1e30', 1e30), ('3.0', 3.0)):   # on-chip state: shared memory + tensor memory
        scrub(val)
        got = run(make)
        same = all(torch.equal(a, b) for a, b in zip(ref, got))
        d = max(float((a - b).abs().nan_to_num(nan=1e9).max()) for a, b in zip(ref, got))
        print(f'{name:7s} scrub  {label:8s}: {"identical" if same else f"DIFFERENT (max |diff| {d:.3e})"}')
    for label, val in (('nan', float('nan')), ('1e30', 1e30), ('-1 bits', None)):
        torch.cuda.empty_cache()
        big = torch.empty(6 << 30, dtype=torch.uint8, device=dev)       # 6 GB of free blocks, poisoned
        if val is None:
            big.fill_(255)
        else:
            big.view(torch.float32).fill_(val)
        del big                                                          # stays in the caching allocator
        got = run(make)
        same = all(torch.equal(a, b) for a, b in zip(ref, got))
        d = max(float((a - b).abs().nan_to_num(nan=1e9).max()) for a, b in zip(ref, got))
        print(f'{name:7s} poison {label:8s}: {"identical" if same else f"DIFFERENT (max |diff| {d:.3e})"}')
