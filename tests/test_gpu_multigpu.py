"""GPU (-m gpu), needs >= 2 GPUs (skipped otherwise): the N-GPU scene-sharded rollout gathered over NCCL equals the 1-GPU
rollout of the same scenes bit for bit (reference split: rollout/callbacks.py:76,104-105)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_nccl_gather_equals_single_gpu_bits():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    n = 2 if n < 4 else 4
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(HERE, 'nccl_gather_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0
    assert res.stdout.count('nccl gather ok') == 2
