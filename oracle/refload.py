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
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nnest"))


def load_reference():
    """Returns the imported reference `nnest` module (never the product package)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
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
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from torch.utils.tensorboard import SummaryWriter
    SummaryWriter.add_figure = lambda *a, **k: None
    import nnest
    assert os.path.abspath(nnest.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), nnest.__file__
    return nnest
