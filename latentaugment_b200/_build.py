"""In-tree build of the CUDA library (``liblatentaugment_b200.so``) with nvcc for sm_100a.

``python -m latentaugment_b200._build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU.  The library links the static CUDA runtime and exposes only
the C ABI of ``include/latentaugment_b200.h``.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'liblatentaugment_b200.so')
SOURCES = ['tapgemm.cu', 'kernels.cu', 'distance.cu', 'engine.cu', 'disc.cu', 'filtered_lrelu.cu']
HEADERS = ['tapgemm.cuh', 'kernels.cuh', 'sm100.cuh', 'plan.cuh', 'disc.cuh', os.path.join(ROOT, 'include', 'latentaugment_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(PKG, 'build', src.replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
