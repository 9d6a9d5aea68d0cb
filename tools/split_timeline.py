#!/usr/bin/env python
"""Kernel timeline (torch.profiler / CUPTI) of a window of the policy stack under the row-split schedule:
which kernels of the sibling chains actually overlap.  usage: split_timeline.py [parts] [first_kernel] [count]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from prosim_b200 import lib, synthetic, weights
from prosim_b200.model import ProSimB200

parts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
first = int(sys.argv[2]) if len(sys.argv) > 2 else 400
count = int(sys.argv[3]) if len(sys.argv) > 3 else 60
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
lib.set_stack_split(parts)
pristine = synthetic.clone_batch(synthetic.make_batch(n_scenes=32, n_agents=128, n_map=512, steps=80), dev)[0]
with torch.no_grad():
    for _ in range(3):
        model.forward(synthetic.clone_batch(pristine)[0], 'val')
    torch.cuda.synchronize()
    b = synthetic.clone_batch(pristine)[0]
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.forward(b, 'val')
        torch.cuda.synchronize()
prof.export_chrome_trace('/tmp/trace.json')
ev = json.load(open('/tmp/trace.json'))['traceEvents']
k = sorted([e for e in ev if e.get('cat') == 'kernel'], key=lambda e: e['ts'])
t0 = k[0]['ts']
span = k[-1]['ts'] + k[-1]['dur'] - t0
print(f'parts={parts} kernels={len(k)} span_ms={span / 1e3:.2f} busy_sum_ms={sum(e["dur"] for e in k) / 1e3:.2f}')
for e in k[first:first + count]:
    a = e.get('args', {})
    print(f'{e["ts"] - t0:10.1f} +{e["dur"]:7.1f} us  stream {a.get("stream", "?"):>3}  grid {str(a.get("grid", "?")):>16}  {e["name"][:48]}')
