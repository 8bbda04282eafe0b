"""dlrm_jl_b200: B200-native (sm_100a) drop-in for DLRM.jl's embedding + dot-interaction hot path.

The compute lives in ``lib/libdlrm_b200.so`` (C ABI: ``include/dlrm_b200.h``; CUDA sources:
``csrc/``).  This package is the host-side mirror of the reference's entry points for that
path.  Importing it does not require a GPU; calling any op does, and there is no fallback.

Submodules: ``embedding`` (tables, maplookup, sparse update), ``interact`` (DotInteraction),
``model`` (DLRMModel, dlrm, kaggle_dlrm), ``train`` (bce_loss, train_step, train),
``sharded`` (table-wise sharding over torch.distributed), ``hdf5_min`` (golden-file reader).
"""
from ._lib import DLRMB200Error, LIB_PATH, launch_count, load  # noqa: F401

__all__ = ["DLRMB200Error", "LIB_PATH", "launch_count", "load"]

