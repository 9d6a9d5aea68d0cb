"""Diagnostic, run explicitly (pytest tests/diag_first_forward.py -m gpu -s): checksums after every kernel call of the scene
encoder in the first forward of a fresh process (and of the second forward for comparison)."""
import hashlib

import pytest
import torch

from prosim_b200 import ops, synthetic
from tests import test_gpu_rollout as T
from tests.helpers import CASES

pytestmark = pytest.mark.gpu


def _h(t):
    return hashlib.sha1(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:4]


def test_first_forward_probe(monkeypatch):
    name = 'cfg1_a16_m256_s20'
    kw, goal = CASES[name]
    model = T._model(goal)
    log = []

    def wrap(fn_name, pick):
        orig = getattr(ops, fn_name)

        def f(*a, **k):
            r = orig(*a, **k)
            log.append((fn_name[:4], [t.clone() for t in pick(a, k, r)]))     # device-side copies in stream order: no sync here
            return r
        monkeypatch.setattr(ops, fn_name, f)
    wrap('pointnet', lambda a, k, r: [k['out']])
    wrap('gather_pose', lambda a, k, r: [a[3], a[4]])
    wrap('knn_edges', lambda a, k, r: [r.nbr, r.deg])
    wrap('edge_pe', lambda a, k, r: [k['z']])
    wrap('attn_layer', lambda a, k, r: [k['out']])
    seqs = []
    for _ in range(2):
        log.clear()
        batch = synthetic.make_batch(**kw).to('cuda')
        with torch.no_grad(), torch.cuda.device(model._device):
            model.mode = 'val'
            ex = batch.extras
            log.append(('in', [ex['init_map'][k].clone() for k in ('input', 'mask', 'position', 'heading')]))
            pl = model._plan(batch)
            log.append(('rows', [pl.i['map_rows'].clone(), pl.i['agent_rows0'].clone(), pl.int_dev.clone()]))
            model.encode_scene(batch)
            log.append(('in2', [ex['init_map'][k].clone() for k in ('input', 'mask', 'position', 'heading')]))
            log.append(('rows2', [pl.i['map_rows'].clone(), pl.int_dev.clone()]))
        torch.cuda.synchronize()
        seqs.append(' '.join(n + ':' + ''.join(_h(t) for t in ts) for n, ts in log))
    print('PROBE', 'same' if seqs[0] == seqs[1] else 'DIFF', '|', seqs[0], '|', seqs[1] if seqs[0] != seqs[1] else '')
