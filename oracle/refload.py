"""Import the real reference package (adammoss/nnest) from /root/reference.

TEST INFRASTRUCTURE ONLY.  /root/reference exists in the build container but NOT on the
GPU box, so nothing under tests/ -m gpu, smoke() or bench.py may call this at run time; it
is used by tests/golden/make_golden.py (which writes the committed fixtures) and by
CPU-side tests that are skipped when the reference tree is absent.

Three shims are needed: matplotlib and getdist are not installed
(stub packages in oracle/_shims), and SummaryWriter.add_figure needs a real figure.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("NNEST_REFERENCE_ROOT", "/root/reference")
# the reference installed by oracle/build_ref.py (pip --target; git-ignored, travels to the GPU box): used by bench.py's
# cpu_baseline / --impl reference legs when the source tree is absent
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nnest"))


def installed_reference_available():
    return os.path.isfile(os.path.join(INSTALLED_ROOT, "nnest", "sampler.py"))


def load_reference(installed=False):
    """Returns the imported reference `nnest` module (never the product package).  installed=True imports the copy
    installed under oracle/_ref instead of the source tree."""
    root = INSTALLED_ROOT if installed else REFERENCE_ROOT
    if not (installed_reference_available() if installed else reference_available()):
        raise RuntimeError("reference not present at %s" % root)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        if _SHIMS not in sys.path:
            sys.path.insert(0, _SHIMS)
    else:
        try:
            import getdist  # noqa: F401
        except ImportError:
            # only getdist missing: expose just that stub
            import importlib.util
            for name, rel in (("getdist", "getdist/__init__.py"),
                              ("getdist.mcsamples", "getdist/mcsamples.py")):
                spec = importlib.util.spec_from_file_location(name, os.path.join(_SHIMS, rel))
                mod = importlib.util.module_from_spec(spec)
                sys.modules[name] = mod
                spec.loader.exec_module(mod)
    if root not in sys.path:
        sys.path.insert(0, root)
    from torch.utils.tensorboard import SummaryWriter
    SummaryWriter.add_figure = lambda *a, **k: None
    import nnest
    assert os.path.abspath(nnest.__file__).startswith(os.path.abspath(root)), nnest.__file__
    return nnest
