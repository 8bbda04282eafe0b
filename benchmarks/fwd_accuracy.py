#!/usr/bin/env python
"""Accuracy of the interaction forward variants against a float64 Gram matrix:
tensor-core 3xTF32 (default) and the general tiled FP32 kernels, on normal and on wide-dynamic-range
inputs.  Prints one JSON line per (variant, input)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlrm_jl_b200 import _lib  # noqa: E402
from dlrm_jl_b200.interact import interaction_fwd  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(7)
    B, F, d = 2048, 27, 128
    inputs = {
        "normal": rng.standard_normal((B, F, d)),
        "lognormal_x_sign": np.exp(3.0 * rng.standard_normal((B, F, d))) * np.sign(rng.standard_normal((B, F, d))),
        "embedding_like": rng.uniform(-1, 1, (B, F, d)) / np.sqrt(1e6),
    }
    jj, ii = np.tril_indices(F, -1)
    for name, T64 in inputs.items():
        T32 = T64.astype(np.float32)
        G = np.einsum("bik,bjk->bij", T32.astype(np.float64), T32.astype(np.float64))[:, jj, ii]
        scale = np.einsum("bik,bjk->bij", np.abs(T32).astype(np.float64), np.abs(T32).astype(np.float64))[:, jj, ii]
        Td = torch.from_numpy(T32).to(dev)
        for variant in ("mma_3xtf32", "tiled_fp32"):
            _lib.set_option("interact_general", 0 if variant == "mma_3xtf32" else 1)
            out = interaction_fwd(Td).cpu().numpy()[:, d:].astype(np.float64)
            err = np.abs(out - G)
            print(json.dumps({"input": name, "variant": variant,
                              "rel_l2": float(np.linalg.norm(out - G) / np.linalg.norm(G)),
                              "max_err_over_sum_abs_products": float(np.max(err / scale)),
                              "mean_signed_err_over_sum_abs_products": float(np.mean((out - G) / scale))}), flush=True)
    _lib.set_option("interact_general", 0)


if __name__ == "__main__":
    main()
