"""TEST INFRASTRUCTURE (not product): stage the UNMODIFIED reference package for the CPU arm of bench.py.

The reference (thoglu/jammy_flows) is pure Python -- there is nothing to compile.  `stage()` installs its package
directory from /root/reference/jammy_flows into oracle/_ref/jammy_flows (byte-for-byte, like `pip install --target`):
oracle/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so the staged package
travels to the GPU box, where /root/reference does not exist.  `import_staged()` imports it from there; matplotlib /
pylab (plotting only: layers/bisection_n_newton.py:3, layers/euclidean/gaussianization_flow.py:21,
helper_fns/contours.py:4-6) are absent from this image and are stubbed before the import.

Only __graft_entry__.build() calls stage(); only bench.py's CPU arm (`--impl reference` / `cpu_baseline`) and tests/
call import_staged().  The product package never imports anything from oracle/.
"""
import os
import shutil
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/jammy_flows"
REF_DST_ROOT = os.path.join(HERE, "_ref")
REF_DST = os.path.join(REF_DST_ROOT, "jammy_flows")


def stage(force=False):
    """Copy the reference package into oracle/_ref/ (no-op where /root/reference is absent). Returns the path or None."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if os.path.isdir(REF_DST) else None
    if os.path.isdir(REF_DST):
        if not force:
            return REF_DST
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST_ROOT, exist_ok=True)
    shutil.copytree(REF_SRC, REF_DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return REF_DST


def available():
    return os.path.isfile(os.path.join(REF_DST, "__init__.py"))


def import_staged():
    """The staged reference package (raises ImportError if it was never staged)."""
    if not available():
        raise ImportError("oracle/_ref/jammy_flows is not staged (run __graft_entry__.build() where /root/reference exists)")
    for name in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.gridspec",
                 "matplotlib.projections", "matplotlib.transforms", "matplotlib._api", "matplotlib.patches",
                 "matplotlib.path", "matplotlib.ticker", "matplotlib.collections", "pylab"):
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    if REF_DST_ROOT not in sys.path:
        sys.path.insert(0, REF_DST_ROOT)
    import jammy_flows
    assert os.path.abspath(jammy_flows.__file__).startswith(REF_DST_ROOT), jammy_flows.__file__
    return jammy_flows


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
