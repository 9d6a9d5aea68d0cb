#!/usr/bin/env python
"""One profiled forward of the bench workload for ncu (run under `ncu --profile-from-start off`): two warm-up
forwards outside the profiled range, then exactly one forward between cudaProfilerStart/Stop.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_forward.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_edge3 -c 3 \
      -o gpurun_out/edge3 python tools/profile_forward.py --ticks 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prosim_b200 import synthetic, weights  # noqa: E402
from prosim_b200.model import ProSimB200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--scenes', type=int, default=32)
ap.add_argument('--agents', type=int, default=128)
ap.add_argument('--map', type=int, default=512)
ap.add_argument('--ticks', type=int, default=8)
a = ap.parse_args()
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
pristine = synthetic.clone_batch(synthetic.make_batch(n_scenes=a.scenes, n_agents=a.agents, n_map=a.map,
                                                      steps=10 * a.ticks), dev)[0]
with torch.no_grad():
    for _ in range(2):
        model.forward(synthetic.clone_batch(pristine)[0], 'val')
    b = synthetic.clone_batch(pristine)[0]
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.forward(b, 'val')
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('profiled one forward')
