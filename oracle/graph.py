"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Restatement of the neighbour-search boundary the reference reaches through the
un-vendored wheel ``torch-cluster`` (unpinned in install_local_env.sh:4, 1.6.3 at
the reference's release date).  Call sites in the reference:
  prosim/models/scene_encoder/attn_fusion.py:107,109   knn_graph(loop=True)
  prosim/models/decoder/sym_coord.py:86,94             radius_graph / radius
  prosim/models/policy/act_decoder.py:250,259          radius

Published semantics restated here (brute force, deterministic):
  radius(x, y, r, batch_x, batch_y, max_num_neighbors) -> [2, E]
      row0 = query (y) index, row1 = x index; only pairs with equal batch id;
      strict ``|x - y|^2 < r^2``; at most ``max_num_neighbors`` per query, the
      FIRST found in ascending x index (the torch-cluster CUDA kernel's order).
  radius_graph(x, r, batch, loop, max_num_neighbors) = radius(x, x, cap(+1 if not
      loop)) with rows swapped to (source = neighbour, target = centre), self
      pairs dropped unless loop.
  knn(x, y, k, batch_x, batch_y) -> [2, E]: the k x-points nearest to every y
      (fewer when the scene is smaller), ties broken to the lower x index.
  knn_graph(x, k, batch, loop) = knn(x, x, k (+1 if not loop)) rows swapped,
      self pairs dropped unless loop.

Squared distance is ``dx*dx + dy*dy`` with every operation rounded separately
(no FMA contraction); the CUDA builders use __fmul_rn/__fadd_rn to match bit
for bit.  Edges come out grouped by query in ascending query index; inside a
query, ascending x index for radius and ascending (distance, index) for knn.
"""
import torch


def _sqdist(y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    d = y[:, None, :] - x[None, :, :]
    d = d * d
    return d[..., 0] + d[..., 1]


def _batch_or_zeros(batch, n, device):
    if batch is None:
        return torch.zeros(n, dtype=torch.long, device=device)
    return batch


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    x = x.view(-1, 1) if x.dim() == 1 else x
    y = y.view(-1, 1) if y.dim() == 1 else y
    batch_x = _batch_or_zeros(batch_x, x.shape[0], x.device)
    batch_y = _batch_or_zeros(batch_y, y.shape[0], y.device)
    if x.shape[0] == 0 or y.shape[0] == 0:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    d2 = _sqdist(y, x)
    r2 = torch.tensor(float(r), dtype=x.dtype) * torch.tensor(float(r), dtype=x.dtype)
    ok = (d2 < r2) & (batch_y[:, None] == batch_x[None, :])
    rank = ok.long().cumsum(dim=1)
    ok = ok & (rank <= max_num_neighbors)
    yi, xi = ok.nonzero(as_tuple=True)
    return torch.stack([yi, xi], dim=0)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow='source_to_target'):
    assert flow == 'source_to_target'
    ei = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    row, col = ei[1], ei[0]
    if not loop:
        keep = row != col
        row, col = row[keep], col[keep]
    return torch.stack([row, col], dim=0)


def knn(x, y, k, batch_x=None, batch_y=None):
    x = x.view(-1, 1) if x.dim() == 1 else x
    y = y.view(-1, 1) if y.dim() == 1 else y
    batch_x = _batch_or_zeros(batch_x, x.shape[0], x.device)
    batch_y = _batch_or_zeros(batch_y, y.shape[0], y.device)
    if x.shape[0] == 0 or y.shape[0] == 0:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    d2 = _sqdist(y, x)
    same = batch_y[:, None] == batch_x[None, :]
    d2 = torch.where(same, d2, torch.full_like(d2, float('inf')))
    kk = min(int(k), x.shape[0])
    vals, idx = torch.sort(d2, dim=1, stable=True)
    vals, idx = vals[:, :kk], idx[:, :kk]
    ok = torch.isfinite(vals)
    yi = torch.arange(y.shape[0], device=x.device)[:, None].expand_as(idx)
    return torch.stack([yi[ok], idx[ok]], dim=0)


def knn_graph(x, k, batch=None, loop=False, flow='source_to_target'):
    assert flow == 'source_to_target'
    ei = knn(x, x, k if loop else k + 1, batch, batch)
    row, col = ei[1], ei[0]
    if not loop:
        keep = row != col
        row, col = row[keep], col[keep]
    return torch.stack([row, col], dim=0)
