"""CPU, world_size 2 over gloo: scene sharding is a partition and the result gather restores scene order."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prosim_b200 import sharding


def test_shard_scenes_is_a_partition():
    for n, w in ((256, 8), (10, 4), (3, 8), (0, 2)):
        for mode in ('block', 'round_robin'):
            parts = [sharding.shard_scenes(n, w, r, mode) for r in range(w)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sharding.shard_scenes(256, 8, 3) == list(range(96, 128))
    assert sharding.shard_scenes(10, 4, 1, 'round_robin') == [1, 5, 9]
    with pytest.raises(ValueError):
        sharding.shard_scenes(4, 2, 2)


def _fake_rollout(scene):
    g = torch.Generator().manual_seed(scene)
    return torch.rand(3, 5, 4, generator=g), torch.rand(3, 5, 2, generator=g)


def _worker(rank, world, port, n_scenes, mode, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = sharding.shard_scenes(n_scenes, world, rank, mode)
    res = [_fake_rollout(s) for s in mine]
    traj = torch.stack([r[0] for r in res]) if res else torch.zeros(0, 3, 5, 4)
    vel = torch.stack([r[1] for r in res]) if res else torch.zeros(0, 3, 5, 2)
    out = sharding.gather_rollouts(traj, vel, mine)
    dist.barrier()
    if rank == 0:
        q.put((out[0], out[1], out[2]))
    else:
        assert out is None
    dist.destroy_process_group()


@pytest.mark.parametrize('n_scenes,mode', [(5, 'block'), (4, 'round_robin')])
def test_gather_rollouts_world2_gloo(n_scenes, mode):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + n_scenes
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_scenes, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    traj, vel, ids = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == list(range(n_scenes))
    for s in range(n_scenes):
        t, v = _fake_rollout(s)
        assert torch.equal(traj[s], t) and torch.equal(vel[s], v)


def test_gather_without_process_group_sorts_locally():
    traj, vel = torch.arange(3.)[:, None].repeat(1, 2), torch.arange(3.)[:, None]
    t, v, ids = sharding.gather_rollouts(traj, vel, [7, 2, 5])
    assert ids == [2, 5, 7] and t[:, 0].tolist() == [1.0, 2.0, 0.0]
