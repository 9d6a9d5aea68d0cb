"""Worker of tests/test_gpu_multigpu.py, launched by torchrun with one process per GPU: every rank rolls out its block of
scenes, the trajectories are gathered on rank 0 over NCCL (prosim_b200.sharding.gather_rollouts -- the path's only
exchange), and rank 0 compares them bit for bit with the same scenes rolled out in ONE batch on its own GPU."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from prosim_b200 import synthetic, weights  # noqa: E402
from prosim_b200.model import HIST, ProSimB200  # noqa: E402
from prosim_b200.sharding import gather_rollouts, shard_scenes  # noqa: E402


def rollout(model, first, n, kw):
    batch = synthetic.make_batch(n_scenes=n, first_scene=first, **kw).to(model.device)
    with torch.no_grad():
        st = model.forward(batch, 'val')['motion_pred']['_state']
    return st['traj'][:, :, HIST:].contiguous(), st['vel'][:, :, HIST:].contiguous()


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    model = ProSimB200(state_dict=weights.random_state_dict(0), device=dev)
    # (scenes, kwargs): a small case whose launches stay below 1024 rows on every GPU count (fp32 FFMA node kernels) and a
    # large one whose launches are >= 1024 rows everywhere (tcgen05 kernels); uneven blocks (scenes % world != 0)
    cases = [(2 * world + 1, dict(n_agents=40, n_map=48, steps=20)),
             (12 * world + 1, dict(n_agents=128, n_map=64, steps=20))]
    for n_scenes, kw in cases:
        mine = shard_scenes(n_scenes, world, rank)
        traj, vel = rollout(model, mine[0], len(mine), kw)
        got = gather_rollouts(traj, vel, mine)
        dist.barrier()
        if rank == 0:
            g_traj, g_vel, g_ids = got
            assert g_ids == list(range(n_scenes)), g_ids
            ref_traj, ref_vel = rollout(model, 0, n_scenes, kw)
            assert torch.equal(g_traj, ref_traj) and torch.equal(g_vel, ref_vel), \
                f'{world}-GPU result differs from the 1-GPU result: {float((g_traj - ref_traj).abs().max())}'
            print(f'nccl gather ok: {n_scenes} scenes over {world} GPUs == 1 GPU, bit for bit ({kw})', flush=True)
        else:
            assert got is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
