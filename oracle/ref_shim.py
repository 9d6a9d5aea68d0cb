"""TEST INFRASTRUCTURE (oracle tier A) -- never imported by the product path.

Runs the UNMODIFIED reference (``/root/reference/prosim``) on CPU behind
``sys.modules`` shims.  Nothing from the reference is copied: its own Python
files are imported from where they lie.  Only this container has the tree, so
this module is used (a) by ``tests/golden/make_golden.py`` to generate the
committed golden vectors and (b) by the ``ref_tree`` tests that pin the
self-contained restatement (``oracle/prosim_oracle.py``) to the reference.

What is shimmed:
  * no-arithmetic stubs: pytorch_lightning, torchmetrics, trajdata, matplotlib,
    seaborn, shapely, imageio, intervaltree, peft, zarr, tensorflow,
    waymo_open_dataset, scipy.ndimage.filters, wandb;
  * yacs.config.CfgNode: a small attribute-dict with the merge methods
    prosim/config/default.py:690-733 calls;
  * the arithmetic boundary in absent wheels: ``torch_cluster`` -> oracle/graph.py,
    ``torch_geometric`` MessagePassing.propagate (aggr='add', node_dim=0, flow
    source->target) and utils.softmax (segment softmax, +1e-16), used at
    prosim/models/layers/attention_layer.py:22,91,117.
"""
import ast
import copy
import importlib.abc
import importlib.machinery
import inspect
import os
import sys
import types
from unittest import mock

import torch
import torch.nn as nn
import yaml

def _find_reference_root():
    """PROSIM_REFERENCE_ROOT, else the mounted tree (authoring container), else the copy staged by baseline/install_ref.py
    (git-ignored; travels to the GPU box)."""
    env = os.environ.get('PROSIM_REFERENCE_ROOT')
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')
    for root in ([env] if env else []) + ['/root/reference', staged]:
        if os.path.isdir(os.path.join(root, 'prosim')):
            return root
    return env or '/root/reference'


REF_ROOT = _find_reference_root()

_MOCKED_TOPLEVEL = (
    'pytorch_lightning', 'torchmetrics', 'trajdata', 'matplotlib', 'seaborn', 'shapely',
    'imageio', 'intervaltree', 'peft', 'zarr', 'tensorflow', 'waymo_open_dataset', 'wandb',
    'cv2', 'PIL', 'moviepy', 'nuscenes', 'l5kit', 'nuplan', 'bokeh', 'kornia', 'dill',
)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'prosim'))


# --------------------------------------------------------------------------- stubs
class _AutoMockModule(types.ModuleType):
    """Module whose every missing attribute is a MagicMock (no arithmetic)."""

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        val = mock.MagicMock(name=f'{self.__name__}.{name}')
        setattr(self, name, val)
        return val


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _MOCKED_TOPLEVEL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _AutoMockModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _LightningModule(nn.Module):
    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device('cpu')

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    def save_hyperparameters(self, *a, **k):
        pass


class _Callback:
    pass


class _Metric(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def add_state(self, name, default, dist_reduce_fx=None):
        setattr(self, name, default)


class _Dataset(torch.utils.data.Dataset):
    def __init__(self, *a, **k):
        pass


class _Plain:
    def __init__(self, *a, **k):
        pass


# --------------------------------------------------------------------------- yacs
class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        if init_dict:
            for k, v in init_dict.items():
                self[k] = CfgNode(v, new_allowed=True) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        pass

    def defrost(self):
        pass

    def register_renamed_key(self, *a, **k):
        pass

    def dump(self, **k):
        return yaml.safe_dump(_to_plain(self))

    @staticmethod
    def _coerce(v):
        if isinstance(v, str):
            try:
                return ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        return v

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], dict):
                    self[k] = type(self)() if type(self) is not CfgNode else CfgNode(new_allowed=True)
                if not isinstance(self[k], CfgNode):
                    self[k] = CfgNode(self[k], new_allowed=True)
                CfgNode._merge(self[k], v)
            else:
                self[k] = copy.deepcopy(v)

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split('.')
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = self._coerce(val)


def _to_plain(n):
    if isinstance(n, dict):
        return {k: _to_plain(v) for k, v in n.items()}
    return n


# --------------------------------------------------------------------------- PyG
def _segment_softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """torch_geometric.utils.softmax: exp(src - segmax) / (segsum + 1e-16)."""
    n = int(index.max()) + 1 if num_nodes is None and index.numel() > 0 else (num_nodes or 0)
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    seg_max = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
    seg_max = seg_max.scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    out = (src - seg_max.gather(0, idx)).exp()
    seg_sum = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(0, index, out)
    return out / (seg_sum.gather(0, idx) + 1e-16)


class MessagePassing(nn.Module):
    """aggr='add', node_dim=0, flow source_to_target: ``_j`` = edge_index[0], ``_i`` = edge_index[1]."""

    def __init__(self, aggr='add', node_dim=0, **kwargs):
        super().__init__()
        assert aggr == 'add' and node_dim == 0

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        msg_params = [p for p in inspect.signature(self.message).parameters]
        n_dst = None
        args = {}
        for p in msg_params:
            if p.endswith('_i'):
                base = kwargs[p[:-2]]
                n_dst = base.shape[0]
                args[p] = base.index_select(0, dst)
            elif p.endswith('_j'):
                args[p] = kwargs[p[:-2]].index_select(0, src)
            elif p == 'index':
                args[p] = dst
            elif p == 'ptr':
                args[p] = None
            else:
                args[p] = kwargs[p]
        assert n_dst is not None
        self._n_dst = n_dst
        msg = self.message(**args)
        out = torch.zeros((n_dst,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        out.index_add_(0, dst, msg)
        upd_params = [p for p in inspect.signature(self.update).parameters][1:]
        return self.update(out, **{p: kwargs[p] for p in upd_params})


def _pyg_softmax(src, index, ptr=None, num_nodes=None, dim=0):
    return _segment_softmax(src, index, ptr, num_nodes, dim)


_INSTALLED = False


def install_shims():
    """Idempotently install every shim, then put the reference tree on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from oracle import graph

    sys.meta_path.insert(0, _MockFinder())
    for name in _MOCKED_TOPLEVEL:
        sys.modules.pop(name, None)

    import pytorch_lightning as pl
    pl.LightningModule = _LightningModule
    pl.Callback = _Callback
    import pytorch_lightning.callbacks as plc
    plc.Callback = _Callback
    import torchmetrics
    torchmetrics.Metric = _Metric
    torchmetrics.MeanMetric = _Metric
    torchmetrics.Accuracy = _Metric
    import trajdata
    import trajdata.dataset
    trajdata.UnifiedDataset = _Dataset
    trajdata.dataset.UnifiedDataset = _Dataset
    import trajdata.augmentation
    trajdata.augmentation.BatchAugmentation = _Plain
    import trajdata.simulation.sim_metrics
    trajdata.simulation.sim_metrics.SimMetric = _Plain
    import trajdata.simulation
    trajdata.simulation.SimulationScene = _Plain

    import scipy.ndimage
    filt = types.ModuleType('scipy.ndimage.filters')
    filt.gaussian_filter = scipy.ndimage.gaussian_filter
    sys.modules['scipy.ndimage.filters'] = filt

    yacs = types.ModuleType('yacs')
    yacs_config = types.ModuleType('yacs.config')
    yacs_config.CfgNode = CfgNode
    yacs.config = yacs_config
    sys.modules['yacs'] = yacs
    sys.modules['yacs.config'] = yacs_config

    tc = types.ModuleType('torch_cluster')
    tc.radius, tc.radius_graph, tc.knn, tc.knn_graph = graph.radius, graph.radius_graph, graph.knn, graph.knn_graph
    sys.modules['torch_cluster'] = tc

    tg = types.ModuleType('torch_geometric')
    tg_nn = types.ModuleType('torch_geometric.nn')
    tg_conv = types.ModuleType('torch_geometric.nn.conv')
    tg_utils = types.ModuleType('torch_geometric.utils')
    tg_conv.MessagePassing = MessagePassing
    tg_utils.softmax = _pyg_softmax
    tg.nn, tg.utils, tg_nn.conv = tg_nn, tg_utils, tg_conv
    sys.modules.update({'torch_geometric': tg, 'torch_geometric.nn': tg_nn,
                        'torch_geometric.nn.conv': tg_conv, 'torch_geometric.utils': tg_utils})

    sys.path.insert(0, REF_ROOT)
    _INSTALLED = True


def reference_config(cond_types=(), opts=()):
    """The released model shape (prosim_demo/cfg/no_text.yaml) with PROMPT.CONDITION.TYPES (and further yacs
    [key, value, ...] options) overridden."""
    install_shims()
    from prosim.config.default import get_config
    return get_config(os.path.join(REF_ROOT, 'prosim_demo/cfg/no_text.yaml'),
                      ['PROMPT.CONDITION.TYPES', repr(list(cond_types))] + list(opts), 'local')


def build_reference_model(cond_types=(), dtype=torch.float32, opts=()):
    """registry.get_model(cfg.MODEL.TYPE)(cfg).eval() -- the reference's own class (traj_sam.py:13-14)."""
    install_shims()
    import prosim  # noqa: F401  (registers everything)
    from prosim.core.registry import registry
    cfg = reference_config(cond_types, opts)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        torch.manual_seed(0)
        model = registry.get_model(cfg.MODEL.TYPE)(cfg).eval()
    finally:
        torch.set_default_dtype(prev)
    return model, cfg


def reference_containers():
    """The reference's own batch containers (dataset/format_utils.py:31-145)."""
    install_shims()
    from prosim.dataset.format_utils import InputMaskData, BatchDataDict
    return InputMaskData, BatchDataDict
