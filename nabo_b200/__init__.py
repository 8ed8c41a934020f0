"""nabo_b200 - B200-native (sm_100a) implementation of nabo's cell-projection
hot path behind the reference's Dataset / Mapping / Graph call surface.

Compute goes through the C-ABI shared library ``libnabo_b200.so`` (hand-written
CUDA, see ``csrc/`` and ``include/nabo_b200.h``).  There is no CPU fallback: any
compute call raises if the library is missing or no B200 is visible.
"""
__version__ = "0.1.0"

from . import store, synth  # noqa: F401  (pure-host helpers)


def __getattr__(name):
    # Lazy so that host-only helpers import without touching the CUDA library.
    if name in ("Dataset",):
        from .dataset import Dataset
        return Dataset
    if name in ("Mapping",):
        from .mapping import Mapping
        return Mapping
    if name in ("Graph",):
        from .graph import Graph
        return Graph
    if name in ("core", "parallel", "dataset", "mapping", "graph", "_lib"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
