#!/usr/bin/env python
"""Repeat a small forward with the L2 flushed (and a fresh model / fresh workspaces) before every run; count runs whose result
differs from the first.  Emulates the timing of the first forward of a process (weights and tokens come from HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prosim_b200 import synthetic, weights
from prosim_b200.model import ProSimB200
from tests.helpers import CASES
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device('cuda', 0)
sd = weights.random_state_dict(0)
kw = CASES['cfg1_a16_m256_s20'][0]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ref, bad, vals = None, 0, {}
for i in range(n):
    model = ProSimB200(state_dict=sd, device=dev) if i % 10 == 0 else model
    b = synthetic.make_batch(**kw).to(dev)
    flush.fill_(i & 255)
    torch.cuda.synchronize()
    with torch.no_grad():
        out = model.forward(b, 'val')['motion_pred']
    t = out['motion_pred'].clone()
    if ref is None:
        ref = t
    elif not torch.equal(t, ref):
        bad += 1
        d = float((t - ref).abs().max())
        vals[d] = vals.get(d, 0) + 1
print(f'cold stress: {bad} of {n - 1} differ; distinct max|diff| values: {vals}; PROSIM_NO_PDL={os.environ.get("PROSIM_NO_PDL")}')
