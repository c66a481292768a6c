"""In-tree build of libnnb.so (hand-written sm_100a CUDA + C ABI).

    python -m nnest_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the tree to the
GPU box; it is rebuilt only when a source is newer.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(LIBDIR, 'obj')
LIB = os.path.join(LIBDIR, 'libnnb.so')
UNITS = ['nnb_api.cu', 'nnb_h16.cu', 'nnb_h32.cu', 'nnb_h64.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ARCH + ['-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libnnb.so cannot be built (there is no CPU fallback)')
    return exe


def _sources():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(HERE, '..', 'include', 'nnb.h'))
    return out


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources())


def _compile(unit):
    obj = os.path.join(OBJDIR, unit.replace('.cu', '.o'))
    cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, unit), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return unit, obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        results = list(ex.map(_compile, UNITS))
    log = []
    for unit, obj, rc, out in results:
        log.append('== %s ==\n%s' % (unit, out))
        if rc != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (unit, out))
    with open(os.path.join(LIBDIR, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    cmd = [_nvcc()] + ARCH + ['-shared', '-o', LIB] + [r[1] for r in results]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
