"""Recipe: install the UNMODIFIED reference (adammoss/nnest) into oracle/_ref/ so that the CPU baseline of bench.py can be
the reference itself (`cpu_baseline.kind == "reference"`) on the GPU box, where /root/reference does not exist.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the history) but not
gpurun-ignored, so it travels with the tree like the built libnnb.so.  The reference's setup.py writes into its source
tree, which is read-only, so the install runs from a copy under /tmp:

    pip install --no-index --no-build-isolation --no-deps --target oracle/_ref /tmp/<copy of /root/reference>

Called by __graft_entry__.build() when /root/reference is present; a no-op when oracle/_ref/nnest already exists.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
SRC = os.environ.get('NNEST_REFERENCE_ROOT', '/root/reference')


def installed():
    return os.path.isfile(os.path.join(REF_DIR, 'nnest', 'sampler.py'))


def build(force=False):
    if installed() and not force:
        return REF_DIR
    if not os.path.isdir(os.path.join(SRC, 'nnest')):
        return None          # GPU box / no reference tree: use what travelled with the snapshot, if anything
    tmp = tempfile.mkdtemp(prefix='nnest_ref_src_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(SRC, src)
        if os.path.isdir(REF_DIR):
            shutil.rmtree(REF_DIR)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
               '--find-links', '/opt/wheelhouse', '--target', REF_DIR, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not installed():
            raise RuntimeError('pip install of the reference failed:\n' + r.stdout[-2000:] + r.stderr[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return REF_DIR


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
