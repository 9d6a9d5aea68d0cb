"""Scene sharding across the GPUs of one box.

Scenes are independent -- every graph op of the path is restricted to same-scene pairs
(attn_fusion.py:107-109, sym_coord.py:86,94, act_decoder.py:250,259) -- so the path shards with NO data-path
collective: one process per GPU (torchrun), rank r rolls out its own scenes, weights are replicated (45 MB).
The reference shards the same way (rollout/callbacks.py:76,247: ``i % device_cnt == device_id``) and only
calls ``dist.barrier()`` afterwards (callbacks.py:104-105).  The one optional exchange is the final gather of
the result trajectories to rank 0 (8 MB per GPU at 32 scenes x 128 agents x 80 steps) over NCCL/NVLink.
Because the kernels are batch invariant (fixed-order reductions), an N-GPU run equals the 1-GPU run bit for bit.
"""
import torch
import torch.distributed as dist


def shard_scenes(n_scenes, world_size, rank, mode='block'):
    """Scene indices of ``rank``.  'block': contiguous blocks (sizes differ by at most one);
    'round_robin': the reference's ``i % world_size == rank`` split (rollout/callbacks.py:76)."""
    if not 0 <= rank < world_size:
        raise ValueError('rank out of range')
    if mode == 'round_robin':
        return list(range(rank, n_scenes, world_size))
    if mode != 'block':
        raise ValueError(mode)
    base, extra = divmod(n_scenes, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def gather_rollouts(traj, vel, scene_ids, group=None, dst=0):
    """Gather per-rank results on ``dst``: traj [B_local, A, steps, 4], vel [B_local, A, steps, 2] and the global
    scene index of every local scene.  Returns (traj, vel, scene_ids) ordered by scene index on ``dst``, None elsewhere.
    Ranks may hold different scene counts (padded to the maximum for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        order = sorted(range(len(scene_ids)), key=lambda i: scene_ids[i])
        return traj[order], vel[order], [scene_ids[i] for i in order]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = traj.device
    counts = [torch.zeros(1, dtype=torch.long, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(scene_ids)], dtype=torch.long, device=dev), group=group)
    counts = [int(c) for c in counts]
    bmax = max(counts)

    def pad(t):
        out = torch.zeros((bmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        out[:t.shape[0]] = t
        return out

    ids = pad(torch.tensor(scene_ids, dtype=torch.long, device=dev))
    packed = [pad(traj.contiguous()), pad(vel.contiguous()), ids]
    outs = []
    for t in packed:
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst, group=group)
        outs.append(bufs)
    if rank != dst:
        return None
    tr = torch.cat([b[:c] for b, c in zip(outs[0], counts)])
    ve = torch.cat([b[:c] for b, c in zip(outs[1], counts)])
    sid = torch.cat([b[:c] for b, c in zip(outs[2], counts)])
    order = torch.argsort(sid)
    return tr[order], ve[order], sid[order].tolist()
