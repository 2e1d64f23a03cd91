"""ctypes binding of libposeb200.so (the C ABI declared in include/poseb200.h).

There is no fallback: if the shared library is missing, or a call fails (for
instance because no CUDA device is present), an exception is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_size_t, c_void_p

LIB_PATH = os.environ.get('PB200_LIB',   # override used by tuning sweeps only
                          os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libposeb200.so'))

F32, F64 = 0, 1
MAX_VIEWS = 8
CAM_STRIDE = 24
RPSM_MAX_JOINTS = 32
TUNE_DECODE_SCHEDULE, DECODE_STATIC, DECODE_DYNAMIC = 2, 0, 1


class Pb200Error(RuntimeError):
    pass


_PROTOS = {
    'pb200_version': (c_int, []),
    'pb200_last_error': (c_char_p, []),
    'pb200_device_check': (c_int, []),
    'pb200_sm_count': (c_int, []),
    'pb200_debug_enabled': (c_int, []),
    'pb200_debug_violations': (c_int, [c_void_p, c_int]),
    'pb200_debug_selftest': (c_int, []),
    'pb200_set_tuning': (c_int, [c_int, c_int]),
    'pb200_crop_affine': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_double, c_double, c_int,
                                  c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'pb200_decode': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    'pb200_decode_flip': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                  c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pb200_transform_preds': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'pb200_project': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'pb200_frame_change': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'pb200_triangulate': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                  c_int, c_void_p, c_void_p]),
    'pb200_reproject': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pb200_ransac': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                             c_int, c_double, c_int, c_void_p, c_void_p]),
    'pb200_epipolar': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                               c_int, c_void_p, c_void_p, c_void_p]),
    'pb200_fundamental': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    'pb200_limb_break': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p]),
    'pb200_softargmax_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    'pb200_softargmax_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                     c_void_p, c_void_p]),
    'pb200_epipolar_grad': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                    c_double, c_void_p, c_void_p]),
    'pb200_mpjpe_stats': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'pb200_lift_fused': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                 c_void_p, c_void_p, c_int, c_int, c_float,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    'pb200_lift_decoded': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pb200_rpsm_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'pb200_rpsm': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                           c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                           c_int, c_int, c_int, c_double, c_double, c_void_p, c_size_t,
                           c_void_p, c_void_p, c_void_p]),
    'pb200_pairwise_level0': (c_int, [c_void_p, c_int, c_int, c_double, c_void_p, c_void_p]),
    'pb200_pairwise_lut_check': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(sorted(_PROTOS))

_lib = None


def load():
    """Load the shared library (once) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Pb200Error(
            'libposeb200.so is not built (%s). Run `python -m pose_unsupervised_b200.build` '
            '(needs nvcc); there is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)           # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pb200_last_error()
        raise Pb200Error('libposeb200 call failed (%d): %s' % (rc, msg.decode() if msg else '?'))


def call(name, *args):
    check(getattr(load(), name)(*args))
