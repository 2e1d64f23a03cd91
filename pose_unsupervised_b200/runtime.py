"""Device plumbing shared by the reference-facing modules.

PyTorch is used for device memory, streams and (in parallel.py) NCCL only; every
arithmetic step of the lifting path runs in libposeb200.so.  Inputs may be numpy
arrays (the reference's convention: results come back as numpy) or CUDA tensors
(results stay on the device, nothing synchronises).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import F32, F64, Pb200Error

_checked = False
_workspaces = {}


def require_device():
    """Raise unless a CUDA device of compute capability 10.x is usable."""
    global _checked
    if _checked:
        return
    if not torch.cuda.is_available():
        raise Pb200Error('pose_unsupervised_b200 needs a CUDA device (sm_100a); there is no CPU path')
    _lib.check(_lib.load().pb200_device_check())
    _checked = True


def device():
    return torch.device('cuda', torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def is_device_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def to_device(x, dtype=None):
    """numpy array / CPU tensor / CUDA tensor -> contiguous CUDA tensor (async when pinned)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if not a.flags['C_CONTIGUOUS'] or not a.flags['WRITEABLE']:
            a = np.array(a, order='C')
        if not a.dtype.isnative:
            a = a.astype(a.dtype.newbyteorder('='))
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.to(device(), non_blocking=True)
    return t.contiguous()


def float_dtype_tag(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise Pb200Error('expected a float32 or float64 array, got %s' % t.dtype)


def to_device_float(x):
    """Keep float32/float64 as they are (the kernels honour the dtype); promote the rest to float64."""
    if isinstance(x, torch.Tensor):
        t = x if x.dtype in (torch.float32, torch.float64) else x.to(torch.float64)
    else:
        a = np.asarray(x)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = a
    return to_device(t)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype):
    return torch.zeros(shape, dtype=dtype, device=device())


def workspace(key, nbytes):
    """Zero-initialised, cached device scratch (kernels leave it clean)."""
    k = (torch.cuda.current_device(), key)
    ws = _workspaces.get(k)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=device())
        _workspaces[k] = ws
    return ws


def reset_workspaces():
    """Drop cached scratch (use after a failed launch: counters may be dirty)."""
    _workspaces.clear()


def to_host(t, like=None):
    """CUDA tensor -> numpy (synchronises)."""
    a = t.cpu().numpy()
    if like is not None and a.dtype != like:
        a = a.astype(like)
    return a


def set_decode_schedule(dynamic):
    """Process-wide: let the decode kernel's warps claim the tail of their maps dynamically (default: it
    evens out the end of the kernel and absorbs blocks that start late under an overlapped collective, see
    parallel.PoseExchange) or use pure static striding.  Results are identical."""
    _lib.call('pb200_set_tuning', _lib.TUNE_DECODE_SCHEDULE,
              _lib.DECODE_DYNAMIC if dynamic else _lib.DECODE_STATIC)
