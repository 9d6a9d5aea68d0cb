#!/usr/bin/env python
"""GPU timeline of one forward via torch.profiler (CUPTI): busy time vs span, largest idle gaps and what follows them."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from prosim_b200 import synthetic, weights
from prosim_b200.model import ProSimB200

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
pristine = synthetic.clone_batch(synthetic.make_batch(n_scenes=S, n_agents=128, n_map=512, steps=80), dev)[0]
with torch.no_grad():
    for _ in range(3):
        model.forward(synthetic.clone_batch(pristine)[0], 'val')
    torch.cuda.synchronize()
    b = synthetic.clone_batch(pristine)[0]
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model.forward(b, 'val')
        torch.cuda.synchronize()
prof.export_chrome_trace('/tmp/trace.json')
ev = json.load(open('/tmp/trace.json'))['traceEvents']
k = sorted([e for e in ev if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')], key=lambda e: e['ts'])
busy = sum(e['dur'] for e in k)
span = k[-1]['ts'] + k[-1]['dur'] - k[0]['ts']
gaps = []
for a, b2 in zip(k[:-1], k[1:]):
    g = b2['ts'] - (a['ts'] + a['dur'])
    gaps.append((g, a['name'][:40], b2['name'][:40]))
print(json.dumps({'kernels': len(k), 'busy_ms': busy / 1e3, 'span_ms': span / 1e3, 'gap_total_ms': sum(max(g[0], 0) for g in gaps) / 1e3}))
import collections
by = collections.defaultdict(lambda: [0, 0.0])
for g, a, b2 in gaps:
    by[b2][0] += 1; by[b2][1] += max(g, 0)
for name, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'gap before {name:42s} n={n:4d} total_ms={t/1e3:7.2f} avg_us={t/n:7.1f}')
cpu = [e for e in ev if e.get('cat') in ('cpu_op', 'cuda_runtime', 'user_annotation')]
rt = collections.defaultdict(lambda: [0, 0.0])
for e in cpu:
    rt[e['name'][:40]][0] += 1; rt[e['name'][:40]][1] += e.get('dur', 0)
for name, (n, t) in sorted(rt.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'cpu {name:42s} n={n:4d} total_ms={t/1e3:7.2f}')
