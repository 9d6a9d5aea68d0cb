#!/usr/bin/env python
"""One line per process: checksum of the synthetic input batch, checksum of the first-tick motion_pred of the first forward,
and its error against the reference golden (the quantity test_closed_loop_matches_reference_golden gates at 1e-5)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from prosim_b200 import synthetic, weights
from prosim_b200.model import ProSimB200
from tests.helpers import CASES, load_golden
name = 'cfg1_a16_m256_s20'
b = synthetic.make_batch(**CASES[name][0])
h = hashlib.sha1()
def walk(o):
    if torch.is_tensor(o): h.update(o.contiguous().cpu().numpy().tobytes())
    elif isinstance(o, dict): [walk(o[k]) for k in sorted(o, key=str)]
    elif hasattr(o, '_data'): walk(o._data)
    elif hasattr(o, 'input') and hasattr(o, 'mask'): walk(o.input); walk(o.mask)
    elif hasattr(o, '__dict__'): walk(vars(o))
    elif isinstance(o, (list, tuple)): [walk(x) for x in o]
walk(b.extras)
sd = weights.random_state_dict(0)
hw = hashlib.sha1()
for k in sorted(sd): hw.update(sd[k].numpy().tobytes())
model = ProSimB200(state_dict=sd, device='cuda')
with torch.no_grad():
    out = model.forward(b.to('cuda'), 'val')['motion_pred']
mp = out['motion_pred'].cpu().numpy()
gold = load_golden(name)
R = len(out['rollout_trajs'])
err = float(np.abs(mp[:R] - gold['motion_pred'][:R]).max())
print(h.hexdigest()[:10], hw.hexdigest()[:10], hashlib.sha1(mp.tobytes()).hexdigest()[:10], f'{err:.4e}')
