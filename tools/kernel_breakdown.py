#!/usr/bin/env python
"""Per-kernel-class device time of one forward, measured with the library's CUDA-event hooks
(prosim_profile_enable): one forward per class, so the numbers are warm-cache and un-serialised
(unlike an ncu launch list).  Usage: python tools/kernel_breakdown.py [--scenes 32] [--agents 128] [--map 512]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prosim_b200 import lib, synthetic, weights  # noqa: E402
from prosim_b200.model import ProSimB200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--scenes', type=int, default=32)
ap.add_argument('--agents', type=int, default=128)
ap.add_argument('--map', type=int, default=512)
ap.add_argument('--steps', type=int, default=80)
a = ap.parse_args()
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
pristine = synthetic.clone_batch(synthetic.make_batch(n_scenes=a.scenes, n_agents=a.agents, n_map=a.map, steps=a.steps), dev)[0]
with torch.no_grad():
    for _ in range(2):
        model.forward(synthetic.clone_batch(pristine)[0], 'val')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b = synthetic.clone_batch(pristine)[0]
    e0.record(); model.forward(b, 'val'); e1.record(); e1.synchronize()
    total = e0.elapsed_time(e1)
    out = {'forward_ms': total}
    for name in lib.KERNEL_CLASSES:
        b = synthetic.clone_batch(pristine)[0]
        lib.profile_enable(name)
        model.forward(b, 'val')
        ms, n = lib.profile_read()
        out[name] = {'ms': round(ms, 3), 'launches': n, 'avg_us': round(1e3 * ms / max(n, 1), 1)}
    lib.profile_enable(None)
out['sum_ms'] = round(sum(v['ms'] for k, v in out.items() if isinstance(v, dict)), 3)
print(json.dumps(out))

# host-side cost of the per-batch bookkeeping (index maps, one mask D2H, one index H2D)
import time  # noqa: E402
ts = []
for _ in range(5):
    b = synthetic.clone_batch(pristine)[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model._plan(b)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({'plan_ms': [round(t, 2) for t in ts]}))
with torch.no_grad():
    b = synthetic.clone_batch(pristine)[0]
    model._plan(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.forward(b, 'val')
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(json.dumps({'forward_host_enqueue_ms': round((t1 - t0) * 1e3, 2), 'forward_total_ms': round((t2 - t0) * 1e3, 2)}))
