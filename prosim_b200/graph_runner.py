"""CUDA-graph replay of the whole rollout for a fixed batch shape.

One ``forward`` is ~600 kernel launches; for a single 128-agent scene the kernels are short and the step is bound by
launch latency, not by the GPU.  ``GraphedForward`` captures ``ProSimB200.forward`` once per batch *shape* (scene /
agent / polyline / tick counts and the per-scene maxima that size the neighbour lists) into a CUDA graph whose inputs
are model-owned static tensors, and replays it: per call only the inputs (and the integer index maps of the new batch)
are copied into the static buffers -- straight from pinned host memory if the batch lives on the host.

The returned tensors are the graph's static outputs: they are overwritten by the next call with the same shape.
"""
import torch

from . import synthetic
from .model import HIST, STEP, _RolloutTrajs


class _Entry:
    pass


def _tensors(batch):
    """Every input tensor of the rollout path, in a fixed order."""
    ex = batch.extras
    out = []
    for key in ('init_obs', 'init_map'):
        out += [ex[key]['input'], ex[key]['mask'], ex[key]['position'], ex[key]['heading']]
    for t in sorted(ex['fut_obs'].keys()):
        f = ex['fut_obs'][t]
        out += [f['input'], f['mask'], f['position'], f['heading']]
    for task in sorted(ex['prompt'].keys()):
        p = ex['prompt'][task]
        out += [p[k] for k in ('prompt', 'prompt_mask', 'position', 'heading', 'agent_type')]
    for c in sorted(ex['condition'].keys()):
        d = ex['condition'][c]
        out += [d[k] for k in ('input', 'mask', 'prompt_idx', 'prompt_mask') if k in d]
    return out


class GraphedForward:
    def __init__(self, model):
        self.model = model
        self._cache = {}

    def __call__(self, batch, mode='val', write_back=False):
        model = self.model
        pl = model._plan(batch)            # host bookkeeping of THIS batch (names, index maps); works for host batches too
        cond = tuple(sorted((c, tuple(batch.extras['condition'][c]['input'].shape)) for c in batch.extras['condition'].keys()))
        key = (pl.key, cond, mode)
        ent = self._cache.get(key)
        if ent is None:
            ent = self._capture(batch, mode)
            self._cache[key] = ent
        with torch.no_grad():
            for dst, src in zip(ent.inputs, _tensors(batch)):
                dst.copy_(src, non_blocking=True)
            ent.plan.int_dev.copy_(pl.int_host, non_blocking=True)
            ent.graph.replay()
            if write_back:
                for t in batch.extras['fut_obs'].keys():
                    for k in ('input', 'mask', 'position', 'heading'):
                        batch.extras['fut_obs'][t][k].copy_(ent.batch.extras['fut_obs'][t][k], non_blocking=True)
        res = dict(ent.out)
        names, agent_names = [], []
        for b, ids in enumerate(pl.policy_ids):
            agent_names += [f'{b}-{a}' for a in ids]
        for t in pl.all_t:
            names += [f'{n}-{t}' for n in agent_names]
        res['pair_names'] = names
        res['rollout_trajs'] = _RolloutTrajs(agent_names, pl.p_b * pl.N + pl.p_n, ent.out['_state'])
        return {model.tasks[0]: res}

    def _capture(self, batch, mode):
        model = self.model
        ent = _Entry()
        ent.batch = synthetic.clone_batch(batch, model.device)[0]
        # The graph bakes in device addresses.  Everything it touches therefore belongs to this entry: a PRIVATE index plan
        # (never shared through the model's plan cache -- every replay overwrites its device maps in place) and a PRIVATE
        # set of model work buffers (the model replaces a buffer when a larger batch arrives, which would free memory an
        # earlier graph still reads and writes).
        ent.plan = model._plan(ent.batch, private=True)
        ent.inputs = _tensors(ent.batch)
        pristine = [t.clone() for t in ent.inputs]
        saved_bufs, ent.bufs = model._bufs, {}
        model._bufs = ent.bufs
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):              # warm-up: allocates every model buffer, sets kernel attributes
                    model.forward(ent.batch, mode)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            ent.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(ent.graph):
                ent.out = model.forward(ent.batch, mode)[model.tasks[0]]
        finally:
            model._bufs = saved_bufs
        for dst, src in zip(ent.inputs, pristine):
            dst.copy_(src)
        return ent
