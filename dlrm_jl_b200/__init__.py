"""dlrm_jl_b200: B200-native (sm_100a) drop-in for DLRM.jl's embedding + dot-interaction hot path.

The compute lives in ``lib/libdlrm_b200.so`` (C ABI: ``include/dlrm_b200.h``; CUDA sources:
``csrc/``).  This package is the host-side mirror of the reference's entry points for that
path.  Importing it does not require a GPU; calling any op does, and there is no fallback.
"""
from ._lib import DLRMB200Error, LIB_PATH, launch_count, load  # noqa: F401

__all__ = ["DLRMB200Error", "LIB_PATH", "launch_count", "load"]


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import dlrm_jl_b200` stays light
    import importlib
    for mod in ("embedding", "interact", "model", "train", "validation", "sharded"):
        try:
            m = importlib.import_module(f"{__name__}.{mod}")
        except ModuleNotFoundError:
            continue
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError(name)
