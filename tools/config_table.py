#!/usr/bin/env python
"""one row per BASELINE config on 1 GPU -- agent-steps/s of the GPU path (eager forward, device
resident, median of 10 after 3 warm-ups; CUDA-graph replay incl. H2D / D2H for the single-scene configs), of the reference's
CPU path on this box's host cores, and the closed-loop error against the reference-generated golden (fp64 run)."""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import CpuReference, timed_forward
from prosim_b200 import synthetic, weights
from prosim_b200.config import get_config
from prosim_b200.graph_runner import GraphedForward
from prosim_b200.model import ProSimB200
from tests.helpers import CASES, DEMO_CASES, demo_batch, load_golden, per_tick_max, stack_rollout

rows = [('1 demo_dataset scene_6, 16 agents, 20 steps', 'cfg1_demo_scene6_a16_s20', None, False),
        ('2 synthetic 64 x 256, 40 steps', 'cfg2_a64_m256_s40', CASES['cfg2_a64_m256_s40'][0], False),
        ('3 synthetic 128 x 512, 80 steps', 'cfg3_a128_m512_s80', CASES['cfg3_a128_m512_s80'][0], False),
        ('4 goal-prompted 128 x 512, 80 steps', 'cfg4_goal_a128_m512_s80', CASES['cfg4_goal_a128_m512_s80'][0], True)]
dev = torch.device('cuda', 0)
out = []
for label, name, kw, goal in rows:
    cfg = get_config(opts=['PROMPT.CONDITION.TYPES', ['goal']] if goal else None)
    model = ProSimB200(cfg, weights.random_state_dict(0, goal), device=dev)
    make = (lambda: demo_batch(name)) if kw is None else (lambda: synthetic.make_batch(**kw))
    pristine = synthetic.clone_batch(make(), dev)[0]
    with torch.no_grad():
        lat = [timed_forward(model, synthetic.clone_batch(pristine)[0])[0] for _ in range(13)]
        res = model.forward(synthetic.clone_batch(pristine)[0], 'val')['motion_pred']
        runner = GraphedForward(model)
        host = make()
        for k in ('init_obs', 'init_map'):
            pass
        glat = []
        for _ in range(13):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = runner(host, 'val')['motion_pred']['_state']
            st['traj'].to('cpu', non_blocking=True)
            torch.cuda.synchronize()
            glat.append((time.perf_counter() - t0) * 1e3)
    ms, gms = statistics.median(lat[3:]), statistics.median(glat[3:])
    names, traj, _ = stack_rollout(res)
    P, steps = traj.shape[0], traj.shape[1]
    gold = load_golden(name)
    gap_gpu, gap_ref = per_tick_max(traj, gold['traj64']), per_tick_max(gold['traj'], gold['traj64'])
    ref = CpuReference() if not goal else None
    if ref is not None and kw is not None:
        t = ref.time(kw.get('n_scenes', 1), kw['n_agents'], kw['n_map'], kw['steps'], repeats=3, warmup=1)
        cpu_ms = 1e3 * statistics.median(t)
        cpu = f'{P * steps / (cpu_ms * 1e-3):,.0f} ({ref.kind}, {ref.cores} threads)'
    else:
        cpu = 'n/a'
    out.append(f'| {label} | {ms:.2f} ms = {P * steps / (ms * 1e-3):,.0f} | {gms:.2f} ms = {P * steps / (gms * 1e-3):,.0f} | {cpu} | '
               f'{gap_gpu[-1]:.1e} (reference fp32: {gap_ref[-1]:.1e}) |')
print('| config | GPU eager forward (device resident): agent-steps/s | CUDA-graph replay incl. H2D / D2H | reference CPU path: agent-steps/s | '
      'max abs xy error vs the reference fp64 run at the last tick |')
print('|---|---|---|---|---|')
print('\n'.join(out))
