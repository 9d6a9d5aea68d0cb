"""ctypes binding of libprosim_b200.so (include/prosim_b200.h).  No pybind / torch extension layer:
tensors cross the boundary as raw device pointers (``tensor.data_ptr()``) plus sizes and the CUDA
stream handle.  There is NO fallback: a missing library, a missing symbol or a non-zero return code
raises immediately."""
import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int32, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libprosim_b200.so')
ABI_VERSION = 9

SYMBOLS = (
    'prosim_abi_version', 'prosim_attn_layer_floats', 'prosim_pointnet_floats', 'prosim_head_floats',
    'prosim_mlp2_floats', 'prosim_attn_workspace_floats', 'prosim_pointnet_fwd', 'prosim_build_radius_edges',
    'prosim_build_knn_edges', 'prosim_edge_pe', 'prosim_attn_kv', 'prosim_attn_layer_fwd', 'prosim_attn_stack_fwd',
    'prosim_policy_head_fwd', 'prosim_reconst_fwd', 'prosim_mlp2_fwd', 'prosim_init_traj', 'prosim_step_env',
    'prosim_gather_pose', 'prosim_step_agent_traj', 'prosim_launch_count', 'prosim_profile_enable', 'prosim_profile_read',
    'prosim_rollout_to_world', 'prosim_tc_gemm_test', 'prosim_set_tensor_core', 'prosim_tc_debug_read', 'prosim_debug_scrub', 'prosim_set_stack_split', 'prosim_tag_embed_fwd', 'prosim_cond_pool_fwd', 'prosim_workspace_bytes', 'prosim_policy_tick', 'prosim_obs_fuse_floats', 'prosim_obs_fuse_fwd',
)

KERNEL_CLASSES = {'pointnet': 0, 'radius': 1, 'knn': 2, 'edge_pe': 3, 'attn_kv': 4, 'attn_dstpre': 5, 'attn_edge': 6,
                  'attn_post': 7, 'head': 8, 'mlp2': 9, 'state': 10, 'edge_qk': 11, 'edge_av': 12, 'attn_post_sw': 13}


class Graph(Structure):
    _fields_ = [('z', c_void_p), ('nbr', c_void_p), ('deg', c_void_p), ('stride', c_int32), ('max_deg', c_int32),
                ('zd', c_int32), ('warps_per_row', c_int32)]


class StackSide(Structure):
    _fields_ = [('w', c_void_p), ('kv', c_void_p), ('kv_layer_stride', c_size_t), ('graph', Graph)]


class Cfg(Structure):
    _fields_ = [(n, c_int32) for n in ('n_policy_rows', 'n_agent_tokens', 'n_map_tokens', 'max_agents_per_scene',
                                       'max_map_per_scene', 'max_neigh', 'n_layers')]


class Tick(Structure):
    _fields_ = [('cfg', Cfg)] + [(n, c_void_p) for n in (
        'emd', 'agent_type', 'p_scene', 'p_pos', 'p_ori', 'x_agent', 'agent_pos', 'agent_ori', 'seg_agent', 'map_pos', 'map_ori',
        'seg_map', 'kv_map', 'w_a2p', 'w_m2p', 'w_head', 'dim_t16')] + [('agent_radius', c_float), ('map_radius', c_float),
        ('noise', c_void_p), ('noise_std', c_float), ('fuse', c_void_p), ('motion_pred', c_void_p), ('p_row', c_void_p),
        ('T', c_int32), ('tidx', c_int32), ('traj', c_void_p), ('vel', c_void_p)]


class ProSimLibError(RuntimeError):
    pass


_P = c_void_p
_SIGS = {
    'prosim_pointnet_fwd': [c_int, _P, _P, _P, c_int, _P, _P, _P, _P],
    'prosim_build_radius_edges': [_P, _P, c_int, _P, _P, c_float, c_int, c_int, _P, _P, c_int, _P],
    'prosim_build_knn_edges': [_P, _P, c_int, _P, _P, c_int, c_int, _P, _P, c_int, _P],
    'prosim_edge_pe': [_P, _P, c_int, _P, _P, _P, _P, c_int, _P, _P, c_int, _P, _P],
    'prosim_attn_kv': [_P, c_int, _P, c_size_t, c_int, _P, c_size_t, _P],
    'prosim_attn_layer_fwd': [_P, c_int, _P, c_int, POINTER(Graph), _P, _P, c_size_t, _P, _P],
    'prosim_attn_stack_fwd': [_P, c_int, c_int, POINTER(StackSide), POINTER(StackSide), _P, c_size_t, _P, _P],
    'prosim_policy_head_fwd': [_P, _P, c_int, _P, _P, c_float, _P, _P],
    'prosim_reconst_fwd': [_P, c_int, _P, _P, _P],
    'prosim_mlp2_fwd': [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P],
    'prosim_tag_embed_fwd': [_P, c_int, c_int, _P, _P, _P, _P],
    'prosim_cond_pool_fwd': [_P, _P, c_int, c_int, _P, _P, _P],
    'prosim_init_traj': [_P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P],
    'prosim_step_env': [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P],
    'prosim_gather_pose': [_P, _P, _P, c_int, _P, _P, _P],
    'prosim_step_agent_traj': [_P, _P, c_int, c_int, c_int, _P, _P, _P],
    'prosim_rollout_to_world': [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P],
    'prosim_tc_gemm_test': [_P, _P, _P, c_int, c_int, _P],
    'prosim_set_tensor_core': [c_int],
    'prosim_tc_debug_read': [_P],
    'prosim_debug_scrub': [c_float, _P],
    'prosim_set_stack_split': [c_int],
    'prosim_policy_tick': [POINTER(Tick), _P, c_size_t, _P],
    'prosim_obs_fuse_fwd': [_P, _P, _P, _P, c_int, _P, _P],
}

_lib = None


def load():
    """Load the shared library once; verify every declared symbol and the packed-weight sizes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ProSimLibError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                             '(there is no CPU or PyTorch fallback for the rollout path)')
    lib = ctypes.CDLL(LIB_PATH)
    missing = [s for s in SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise ProSimLibError(f'libprosim_b200.so lacks symbols {missing}')
    for name in ('prosim_abi_version', 'prosim_attn_layer_floats', 'prosim_pointnet_floats', 'prosim_head_floats',
                 'prosim_mlp2_floats', 'prosim_obs_fuse_floats'):
        getattr(lib, name).restype = c_int
        getattr(lib, name).argtypes = []
    lib.prosim_launch_count.restype = ctypes.c_longlong
    lib.prosim_launch_count.argtypes = [c_int]
    lib.prosim_profile_enable.restype = c_int
    lib.prosim_profile_enable.argtypes = [c_int]
    lib.prosim_profile_read.restype = c_int
    lib.prosim_profile_read.argtypes = [POINTER(ctypes.c_double), POINTER(c_int)]
    lib.prosim_attn_workspace_floats.restype = c_size_t
    lib.prosim_attn_workspace_floats.argtypes = [c_int, c_int, c_int]
    lib.prosim_workspace_bytes.restype = c_size_t
    lib.prosim_workspace_bytes.argtypes = [POINTER(Cfg)]
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = sig
    if lib.prosim_abi_version() != ABI_VERSION:
        raise ProSimLibError('libprosim_b200.so ABI version mismatch: rebuild')
    from . import weights
    sizes = (lib.prosim_attn_layer_floats(), lib.prosim_pointnet_floats(), lib.prosim_head_floats(),
             lib.prosim_mlp2_floats(), lib.prosim_obs_fuse_floats())
    want = (weights.ATTN_LAYER_FLOATS, weights.POINTNET_FLOATS, weights.HEAD_FLOATS, weights.MLP2_FLOATS, weights.OBS_FUSE_FLOATS)
    if sizes != want:
        raise ProSimLibError(f'packed weight layout mismatch: library {sizes} vs packer {want}')
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        kind = 'bad argument' if code == -1 else 'workspace too small' if code == -2 else f'cudaError {code}'
        raise ProSimLibError(f'{what} failed: {kind}')


def call(name, *args):
    check(getattr(load(), name)(*args), name)


def ptr(t, offset_elems=0):
    """Device pointer of a tensor (plus an element offset); None -> NULL."""
    if t is None:
        return None
    return t.data_ptr() + offset_elems * t.element_size()


def launch_count(kernel_class=-1):
    return int(load().prosim_launch_count(kernel_class))


def profile_enable(name):
    """Record a CUDA-event pair around every launch of one kernel class (None disables)."""
    check(load().prosim_profile_enable(-1 if name is None else KERNEL_CLASSES[name]), 'prosim_profile_enable')


def profile_read():
    """(total milliseconds, launches) recorded since the last enable/read."""
    ms, n = ctypes.c_double(0.0), c_int(0)
    check(load().prosim_profile_read(ctypes.byref(ms), ctypes.byref(n)), 'prosim_profile_read')
    return ms.value, n.value


def set_tensor_core(on):
    """Tensor-core (tcgen05 / TMEM, 3xTF32) kernels: True = all of them (the default), False = fp32 FFMA kernels, an int
    mask selects by bit -- 1 node kernels, 2 K'|V', 4 PointNet, 8 the 32-row "swapped" node kernel (post_sw.cuh; 7 = the 128-row
    kernel of tc_post.cuh for launches >= 1024 rows, FFMA below), 16 the fused single-launch edge phase of small launches.  PROCESS-WIDE switch."""
    call('prosim_set_tensor_core', int(on) if not isinstance(on, bool) else (1 if on else 0))


def set_stack_split(parts):
    """Row-split chains of the fixed-source attention stacks (1 = single stream, the default; bit-identical results)."""
    call('prosim_set_stack_split', int(parts))

