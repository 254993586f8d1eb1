"""Build libunires_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m unires_b200.build [--force] [--verbose]

The library has no torch / Python dependency: plain CUDA runtime only.  The
.so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
BUILD = os.path.join(HERE, 'csrc', 'build')
LIB = os.path.join(HERE, 'libunires_b200.so')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
         '-I', INCLUDE, '-I', CSRC]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    paths += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    paths.append(os.path.abspath(__file__))
    return max(os.path.getmtime(p) for p in paths)


def _compile(src, verbose):
    obj = os.path.join(BUILD, src[:-3] + '.o')
    cmd = [NVCC] + ARCH + FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
        ['-c', os.path.join(CSRC, src), '-o', obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, res.stdout, res.stderr))
    return obj, res.stderr


def build_library(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libunires_b200.so."""
    os.makedirs(BUILD, exist_ok=True)
    srcs = sources()
    dep_t = _deps_mtime()
    todo, objs = [], []
    for s in srcs:
        obj = os.path.join(BUILD, s[:-3] + '.o')
        objs.append(obj)
        src_t = max(os.path.getmtime(os.path.join(CSRC, s)), dep_t)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_t:
            todo.append(s)
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for obj, log in ex.map(lambda s: _compile(s, verbose), todo):
                if verbose and log:
                    print(log)
    if todo or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (res.stdout, res.stderr))
    return LIB


if __name__ == '__main__':
    lib = build_library(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(lib)
