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
# NNB_LIB_DIR: a directory (relative to the package) for a variant build, e.g. NNB_EXTRA_NVCC_FLAGS=-DNNB_TC_TIMING
# NNB_LIB_DIR=lib_timing -- development A/B runs; the product library is lib/libnnb.so
LIBDIR = os.path.join(HERE, os.environ.get('NNB_LIB_DIR', 'lib'))
OBJDIR = os.path.join(LIBDIR, 'obj')
LIB = os.path.join(LIBDIR, 'libnnb.so')
UNITS = ['nnb_api.cu', 'nnb_tc.cu', 'nnb_tc_d2.cu', 'nnb_tc_d10.cu', 'nnb_tc_d30.cu', 'nnb_tc_d50.cu', 'nnb_warp.cu', 'nnb_spline.cu', 'nnb_train.cu', 'nnb_stats.cu', 'nnb_h16.cu', 'nnb_h32.cu', 'nnb_h64.cu']
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


def _digest(unit):
    """sha1 over the unit, every local header it includes and the compiler flags (mtimes do not survive a copy of the
    tree to another machine; a rebuild of all units takes minutes)."""
    import hashlib
    h = hashlib.sha1(' '.join(NVCC_FLAGS).encode())
    for path in sorted(_deps(os.path.join(CSRC, unit))):
        h.update(path[len(CSRC):].encode())
        h.update(open(path, 'rb').read())
    return h.hexdigest()


def _manifest():
    import json
    try:
        return json.load(open(os.path.join(LIBDIR, 'build_manifest.json')))
    except Exception:
        return {}


def _unit_stale(unit, manifest=None):
    manifest = _manifest() if manifest is None else manifest
    return not os.path.exists(_obj(unit)) or manifest.get(unit) != _digest(unit)


def up_to_date():
    """libnnb.so matches the sources (content digests; the object files are not needed once the library is linked, so
    they do not have to travel with the tree)."""
    if not os.path.exists(LIB):
        return False
    manifest = _manifest()
    return manifest.get('__lib__') == ' '.join(_digest(u) for u in UNITS)


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
    import json
    manifest = {u: _digest(u) for u in UNITS}
    manifest['__lib__'] = ' '.join(manifest[u] for u in UNITS)
    with open(os.path.join(LIBDIR, 'build_manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=1)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
