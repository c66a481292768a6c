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
UNITS = ['nnb_api.cu', 'nnb_tc.cu', 'nnb_train.cu', 'nnb_h16.cu', 'nnb_h32.cu', 'nnb_h64.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ARCH + ['-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-Xptxas', '-v'] + \
    os.environ.get('NNB_EXTRA_NVCC_FLAGS', '').split()


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libnnb.so cannot be built (there is no CPU fallback)')
    return exe


def _deps(path, seen=None):
    """The file plus every local header it (transitively) includes."""
    import re
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for inc in re.findall(r'#include\s+"([^"]+)"', open(path).read()):
        _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _obj(unit):
    return os.path.join(OBJDIR, unit.replace('.cu', '.o'))


def _unit_stale(unit):
    obj = _obj(unit)
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in _deps(os.path.join(CSRC, unit)))


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return not any(_unit_stale(u) for u in UNITS) and all(os.path.getmtime(_obj(u)) <= t for u in UNITS)


def _compile(unit):
    obj = _obj(unit)
    cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, unit), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return unit, obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [u for u in UNITS if force or _unit_stale(u)]
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(todo))) as ex:
        results = list(ex.map(_compile, todo))
    for unit, obj, rc, out in results:
        with open(os.path.join(LIBDIR, 'ptxas_%s.log' % unit.replace('.cu', '')), 'w') as f:
            f.write(out)
        if rc != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (unit, out))
        if verbose:
            print('== %s ==\n%s' % (unit, out))
    cmd = [_nvcc()] + ARCH + ['-shared', '-o', LIB] + [_obj(u) for u in UNITS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
