"""ctypes binding of libunires_b200.so (include/unires_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing the
import fails loudly, and every operator refuses non-CUDA tensors.
"""
import ctypes as C
import os
import re
import weakref

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libunires_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(HERE), 'include', 'unires_b200.h')

UR_MAX_TAPS = 32
UR_MAX_OBS = 8
UR_MAX_CHANNELS = 16
UR_CG_MAX_ITER = 256

UR_OK, UR_ERR_ARG, UR_ERR_CUDA, UR_ERR_UNSUPPORTED = 0, 1, 2, 3
UR_SUPERRES, UR_DENOISE = 0, 1
UR_OP_A, UR_OP_AT, UR_OP_ATA = 0, 1, 2
UR_STOP_NONE, UR_STOP_RESIDUAL, UR_STOP_ENERGY = 0, 1, 2

OPS = {'A': UR_OP_A, 'At': UR_OP_AT, 'AtA': UR_OP_ATA}
METHODS = {'super-resolution': UR_SUPERRES, 'denoising': UR_DENOISE}


class ur_proj(C.Structure):
    _fields_ = [('method', C.c_int32),
                ('dim_y', C.c_int32 * 3),
                ('dim_x', C.c_int32 * 3),
                ('dim_yx', C.c_int32 * 3),
                ('ratio', C.c_int32 * 3),
                ('ksize', C.c_int32 * 3),
                ('ker', (C.c_float * UR_MAX_TAPS) * 3),
                ('mat', C.c_float * 12),
                ('scl', C.c_float),
                ('dim_thick', C.c_int32)]


class ur_lhs(C.Structure):
    _fields_ = [('dim_y', C.c_int32 * 3),
                ('vx', C.c_float * 3),
                ('rho_lam2', C.c_float),
                ('do_proj', C.c_int32),
                ('n_obs', C.c_int32),
                ('tau', C.c_float * UR_MAX_OBS),
                ('obs', ur_proj * UR_MAX_OBS)]


class ur_cg_opts(C.Structure):
    _fields_ = [('max_iter', C.c_int32),
                ('stop_rule', C.c_int32),
                ('tolerance', C.c_double),
                ('variant', C.c_int32)]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            'unires_b200: %s is missing -- the CUDA extension has not been built. '
            'Run `python -m unires_b200.build` (needs nvcc); there is no CPU fallback.'
            % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()

_p = C.c_void_p
_i3 = C.POINTER(C.c_int32)
_f3 = C.POINTER(C.c_float)
_sz = C.c_size_t

_SIGNATURES = {
    'ur_last_error': (C.c_char_p, []),
    'ur_version': (C.c_int, []),
    'ur_device_info': (C.c_int, [C.POINTER(C.c_int)] * 3),
    'ur_launch_count': (C.c_uint64, []),
    'ur_profile_matvec': (C.c_int, [C.c_int]),
    'ur_profile_matvec_read': (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_double)]),
    'ur_tune': (C.c_int, [C.c_char_p, C.c_int]),
    'ur_last_lhs_path': (C.c_int, []),
    'ur_scaling_sums': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int,
                                  C.c_void_p, C.c_void_p]),
    'ur_scale_slices': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_float,
                                  C.c_float, C.c_int, C.c_void_p]),
    'ur_affine_grad': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                 C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_void_p]),
    'ur_rigid_sums': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    'ur_im_gradient': (C.c_int, [_p, _p, _i3, _f3, _p]),
    'ur_im_divergence': (C.c_int, [_p, _p, _i3, _f3, _p]),
    'ur_dtd': (C.c_int, [_p, _p, _i3, _f3, _p]),
    'ur_grid_pull': (C.c_int, [_p, _i3, _p, _p, _i3, C.c_int, C.c_int, _p]),
    'ur_grid_push': (C.c_int, [_p, _i3, _p, _p, _i3, C.c_int, C.c_int, C.c_float, _p]),
    'ur_affine_pull': (C.c_int, [_p, _i3, _f3, _p, _i3, C.c_int, C.c_int, _p]),
    'ur_affine_push': (C.c_int, [_p, _i3, _f3, _p, _i3, C.c_int, C.c_int, C.c_float, _p]),
    'ur_affine_grid': (C.c_int, [_f3, _p, _i3, _p]),
    'ur_conv_axis': (C.c_int, [_p, _i3, _p, C.c_int, _f3, C.c_int, C.c_int, C.c_int, _p]),
    'ur_apply_scaling': (C.c_int, [_p, _p, _i3, C.c_float, C.c_int, _p]),
    'ur_proj_is_lattice': (C.c_int, [C.POINTER(ur_proj)]),
    'ur_proj_workspace_bytes': (_sz, [C.POINTER(ur_proj)]),
    'ur_rot_cell_colours': (C.c_int, [C.POINTER(C.c_float)]),
    'ur_proj_apply': (C.c_int, [C.c_int, C.POINTER(ur_proj), _p, _p, _p, _sz, _p]),
    'ur_proj_accumulate': (C.c_int, [C.c_int, C.POINTER(ur_proj), _p, _p, C.c_float, _p, _sz, _p]),
    'ur_lhs_workspace_bytes': (_sz, [C.POINTER(ur_lhs)]),
    'ur_lhs_apply': (C.c_int, [C.POINTER(ur_lhs), _p, _p, _p, _p, _sz, _p]),
    'ur_cg_workspace_bytes': (_sz, [C.POINTER(ur_lhs)]),
    'ur_cg_solve': (C.c_int, [C.POINTER(ur_lhs), _p, _p, _p, _sz, C.POINTER(ur_cg_opts), _p]),
    'ur_cg_fetch': (C.c_int, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int32, _p]),
    'ur_dot': (C.c_int, [_p, _p, _sz, _p, _p]),
    'ur_cg_update_xr': (C.c_int, [_p, _p, _p, _p, _sz, _p, _p, _p]),
    'ur_cg_update_p': (C.c_int, [_p, _p, _sz, _p, _p]),
    'ur_admm_rhs': (C.c_int, [_p, _p, _p, _i3, _f3, C.c_float, C.c_float, _p]),
    'ur_backproject': (C.c_int, [C.POINTER(ur_lhs), C.POINTER(_p), _p, _p, _p]),
    'ur_admm_rhs_fused': (C.c_int, [C.POINTER(ur_lhs), C.POINTER(_p), _p, _p, _p, C.c_float,
                                    C.c_float, _p]),
    'ur_axpy': (C.c_int, [_p, _p, C.c_float, _sz, _p]),
    'ur_jtv_prox': (C.c_int, [C.POINTER(_p), _p, _p, _p, C.c_int, _f3, _i3, _f3, C.c_float,
                              C.c_float, _p]),
    'ur_jtv_norm2': (C.c_int, [C.POINTER(_p), _p, _p, _p, C.c_int, _f3, _i3, _f3, C.c_float,
                               C.c_float, C.c_int, _p]),
    'ur_jtv_apply': (C.c_int, [C.POINTER(_p), _p, _p, _p, _p, C.c_int, _f3, _i3, _f3, C.c_float,
                               C.c_float, _p]),
    'ur_nll_data': (C.c_int, [_p, _p, _sz, C.c_float, _p, C.c_int, _p]),
    'ur_nll_data_proj': (C.c_int, [C.POINTER(ur_proj), _p, _p, C.c_float, _p, C.c_int, _p, _sz, _p]),
    'ur_nll_prior_energy': (C.c_int, [C.POINTER(_p), _p, C.c_int, _f3, _i3, _f3, C.c_int, _p]),
    'ur_sqrt_sum': (C.c_int, [_p, _sz, _p, _p]),
    'ur_intensity_range': (C.c_int, [_p, _sz, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_float),
                                     C.POINTER(C.c_int32), _p]),
    'ur_histc': (C.c_int, [_p, _sz, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_int,
                           _p, _p]),
}

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def header_symbols():
    """Every function name declared in include/unires_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ur_[a-z0-9_]+)\s*\(', text)))


def check(rc):
    """Map a C return code to the exception the reference would raise."""
    if rc == UR_OK:
        return
    msg = lib.ur_last_error().decode('utf-8', 'replace')
    if rc == UR_ERR_ARG:
        raise ValueError(msg)
    if rc == UR_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError('unires_b200 CUDA error: ' + msg)


# ---------------------------------------------------------------------------
# tensor plumbing
# ---------------------------------------------------------------------------
def stream(device=None):
    """Current CUDA stream of `device` (default: the current device) as a void*."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr())


def i3(v):
    return (C.c_int32 * 3)(*[int(a) for a in v])


def f3(v):
    return (C.c_float * 3)(*[float(a) for a in v])


def farr(v):
    return (C.c_float * len(v))(*[float(a) for a in v])


def require_cuda_f32(t, name='tensor'):
    if not isinstance(t, torch.Tensor):
        raise TypeError('%s must be a torch.Tensor' % name)
    if not t.is_cuda:
        raise RuntimeError('unires_b200: %s is on %s; the operators are CUDA-only '
                           '(there is no CPU fallback)' % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError('unires_b200: %s must be float32, got %s' % (name, t.dtype))
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's stream: pointers of another device would
        # be dereferenced on the wrong GPU
        raise RuntimeError('unires_b200: %s is on %s but the current CUDA device is cuda:%d; '
                           'wrap the call in torch.cuda.device(...)'
                           % (name, t.device, torch.cuda.current_device()))
    return t if t.is_contiguous() else t.contiguous()


_host_cache = {}


def host_values(v):
    """Python floats of a scalar / small tensor WITHOUT a device sync on repeat calls.

    UniRes keeps tau, lam, rho, scl and the affine matrices as (often CUDA) tensors; reading
    them with float() / .tolist() synchronises the stream, which would serialise host and
    device in the ADMM loop.  Values are cached per tensor object and in-place version."""
    if isinstance(v, torch.Tensor):
        if not v.is_cuda:
            return v.detach().reshape(-1).to(torch.float64).tolist()
        # keyed by object id, validated by a weak reference (ids and device addresses are
        # recycled after garbage collection) and by the in-place version counter
        hit = _host_cache.get(id(v))
        if hit is not None and hit[0]() is v and hit[1] == v._version:
            return hit[2]
        if len(_host_cache) > 4096:
            _host_cache.clear()
        vals = v.detach().reshape(-1).to('cpu', torch.float64).tolist()
        _host_cache[id(v)] = (weakref.ref(v), v._version, vals)
        return vals
    if isinstance(v, (int, float)):
        return [float(v)]
    return [float(a) for a in v]


def host_scalar(v):
    return host_values(v)[0]


_ws_cache = {}


def workspace(nbytes, device, tag='ws'):
    """Grow-only byte scratch per (device, stream, tag)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def copy_bag(obj):
    """Shallow copy of an attribute bag (struct._proj_op and friends)."""
    import copy
    return copy.copy(obj)
