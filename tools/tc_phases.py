#!/usr/bin/env python
"""Phase timeline (SM cycles) of CTA 0 of the last tcgen05 node-kernel launch of a forward."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prosim_b200 import lib, synthetic, weights
from prosim_b200.model import ProSimB200
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
lib.set_stack_split(1)
b = synthetic.clone_batch(synthetic.make_batch(n_scenes=32, n_agents=128, n_map=512, steps=20), dev)[0]
with torch.no_grad():
    model.forward(b, 'val')
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
lib.call('prosim_tc_debug_read', ctypes.cast(buf, ctypes.c_void_p))
t = list(buf)
t0 = t[0]
names_e = ['start', 'G1 staged', 'G1 acc', 'agg->A', 'gate acc', 'u->A', 'out acc', 'xn->A', 'ffn done', 'y acc', 'out,xd->A', 'q->A',
           'qhat done', 'stores done']
print('epilogue:', ' | '.join(f'{n} {t[i] - t0}' for i, n in enumerate(names_e)))
names_m = ['start', 'G1 issued', 'a(agg)', 'gate issued', 'a(u)', 'out issued', 'a(xn)', 'ffn issued', 'a(xd)', 'sgq issued', 'a(q)',
           'qhat issued']
print('mma     :', ' | '.join(f'{n} {t[16 + i] - t0}' for i, n in enumerate(names_m)))
print('mma waits (cycles): before FFN weights', t[28], 'epilogue', t[29], '| FFN only: weights', t[30] - t[28], 'epilogue', t[31] - t[29])
