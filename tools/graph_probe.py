import sys, time, statistics, torch
sys.path.insert(0, '/root/repo')
from prosim_b200 import synthetic, weights
from prosim_b200.model import ProSimB200
from prosim_b200.graph_runner import GraphedForward
dev = torch.device('cuda', 0)
model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
runner = GraphedForward(model)
for S in (1, 32):
    hosts = [synthetic.make_batch(n_scenes=S, n_agents=128, n_map=512, steps=80, first_scene=v * S, pin_memory=True) for v in range(3)]
    with torch.no_grad():
        for i in range(3):
            runner(hosts[i % 3], 'val')
        torch.cuda.synchronize()
        ts = []
        for i in range(10):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = runner(hosts[i % 3], 'val')['motion_pred']
            st = out['_state']; tr = st['traj'].to('cpu', non_blocking=True); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        eager = []
        for i in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            b = synthetic.clone_batch(hosts[i % 3], dev, non_blocking=True)[0]
            out = model.forward(b, 'val')['motion_pred']
            tr = out['_state']['traj'].to('cpu', non_blocking=True); torch.cuda.synchronize()
            eager.append((time.perf_counter() - t0) * 1e3)
    print(S, 'graphed e2e ms', round(statistics.median(ts), 2), 'eager e2e ms', round(statistics.median(eager[2:]), 2))
