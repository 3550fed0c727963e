"""Import the UNMODIFIED reference (thoglu/jammy_flows, /root/reference) in the build container.

Only used by `make_golden.py` (and ad-hoc validation) in the container that has /root/reference.
Nothing on the GPU box imports this module: /root/reference does not exist there.

The reference imports matplotlib/pylab at module top level purely for plotting
(reference: jammy_flows/layers/bisection_n_newton.py:3, layers/euclidean/gaussianization_flow.py:21,
helper_fns/contours.py:4-6, helper_fns/plotting/*.py); those modules are absent in this image, so they are
stubbed with MagicMock before the import.  No reference source is modified or copied.
"""
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"


def import_reference():
    for name in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.gridspec",
                 "matplotlib.projections", "matplotlib.transforms", "matplotlib._api", "matplotlib.patches",
                 "matplotlib.path", "matplotlib.ticker", "matplotlib.collections", "pylab"):
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import jammy_flows  # noqa: E402
    return jammy_flows
