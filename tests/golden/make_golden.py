"""Regenerate tests/golden/*.npz from the reference's own PyTorch golden files.

Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [/root/reference]

Source fixtures: ref/pytorch_reference_single.hdf5 and ref/pytorch_reference_multi.hdf5,
the files `test/integration.jl:4-41` and `src/validation.jl:1-44` of the reference check
against.  Every dataset is kept verbatim (same names, HDF5/C order, same dtypes); only the
container changes (npz instead of HDF5) so the tests need no HDF5 library at run time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from dlrm_jl_b200.hdf5_min import read_hdf5  # noqa: E402


def main(ref_root: str) -> None:
    for name in ("single", "multi"):
        src = os.path.join(ref_root, "ref", f"pytorch_reference_{name}.hdf5")
        data = read_hdf5(src)
        assert len(data) == 58, len(data)
        dst = os.path.join(HERE, f"pytorch_reference_{name}.npz")
        np.savez_compressed(dst, **data)
        print(f"{src} -> {dst}: {len(data)} datasets, {os.path.getsize(dst)} bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
