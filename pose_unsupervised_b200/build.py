"""Build libposeb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m pose_unsupervised_b200.build

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.environ.get('PB200_LIB', os.path.join(PKG_DIR, 'libposeb200.so'))   # PB200_LIB: tuning sweeps only
SOURCES = ['api.cu', 'decode.cu', 'geometry.cu', 'lift_fused.cu', 'rpsm.cu', 'softargmax.cu']
HEADERS = ['pb_common.cuh', 'lift_math.cuh', 'decode.cuh', 'lift.cuh',
           os.path.join('..', '..', 'include', 'poseb200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              # numpy's element-wise float64 arithmetic is unfused; fma() is explicit where wanted
              '-fmad=false', '-std=c++17', '-shared', '-Xcompiler', '-fPIC']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libposeb200.so cannot be built')


def is_stale(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        return True
    built = os.path.getmtime(path)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


DEBUG_LIB_PATH = os.path.join(PKG_DIR, 'libposeb200_debug.so')


def build(force=False, verbose=False, debug=False):
    """Compile every CUDA source into pose_unsupervised_b200/libposeb200.so (debug=True: the debug-assert
    build libposeb200_debug.so, -DPB200_DEBUG_CHECKS=1, used by tests/test_gpu_debug_build.py)."""
    out = DEBUG_LIB_PATH if debug else LIB_PATH
    if not force and not is_stale(out):
        return out
    extra = os.environ.get('PB200_NVCC_EXTRA', '').split()           # e.g. -DPB_STAGES=4 (tuning sweeps)
    if debug:
        extra = extra + ['-DPB200_DEBUG_CHECKS=1']
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + \
        ['-o', out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed (exit %d): %s' % (res.returncode, ' '.join(cmd)))
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, debug='--debug' in sys.argv))
