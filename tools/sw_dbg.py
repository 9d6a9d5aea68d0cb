#!/usr/bin/env python
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prosim_b200 import lib, synthetic, weights
from prosim_b200.model import ProSimB200
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
def dbg():
    buf = (ctypes.c_longlong * 32)()
    lib.call('prosim_tc_debug_read', ctypes.cast(buf, ctypes.c_void_p))
    return list(buf)
for kw in (dict(n_scenes=11, n_agents=100, n_map=64, steps=20), dict(n_scenes=32, n_agents=128, n_map=512, steps=20)):
    b = synthetic.clone_batch(synthetic.make_batch(**kw), dev)[0]
    with torch.no_grad():
        model.forward(b, 'val')
    torch.cuda.synchronize()
    t = dbg()
    print(kw, 'timeouts', t[14], 'last id', t[15])
