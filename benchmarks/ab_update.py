#!/usr/bin/env python
"""A/B of the sparse update's switches in ONE process (tables built once): tile length (`update_tile`), one- or
two-launch fix-up (`update_two_launches`).  (Round 2, visit U also ran it with two switches that were measured
and then removed from the library: L2 prefetch hints for a tile's rows and a larger same-launch fix-up limit --
results in profiles/r02_update_ab.txt.)

    python benchmarks/ab_update.py [--workload terabyte|kaggle] [--B 2048] [--nb 16]

Prints one JSON object per line: {"B", "opts", "sort_plus_update_us", "chain_us", "update_only_us", ...}.
`update_only` = (sort + update) - sort, cold gradient rows and cold table rows (nb different batches per
graph, gradient buffers rotated so a launch does not find the previous launch's rows in L2).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dlrm_jl_b200 import _lib, _prof  # noqa: E402
from dlrm_jl_b200.embedding import EmbeddingTables  # noqa: E402
from dlrm_jl_b200.model import KAGGLE_EMBEDDING_SIZES, TERABYTE_EMBEDDING_SIZES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="terabyte", choices=["terabyte", "kaggle"])
    ap.add_argument("--B", type=int, nargs="*", default=[2048])
    ap.add_argument("--nb", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cases", default="tile,two")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    if a.workload == "kaggle":
        rows, D = list(KAGGLE_EMBEDDING_SIZES), 64
    else:
        rows, D = [min(r, 40_000_000) for r in TERABYTE_EMBEDDING_SIZES], 128
    ntab = len(rows)
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    t = EmbeddingTables(rows, D, max(a.B), dev)
    t.init_uniform(1)
    rng = np.random.default_rng(1234)
    for B in a.B:
        idx_np = [np.stack([rng.integers(0, r, size=(B, 1)) for r in rows]).astype(np.int32) for _ in range(a.nb)]
        idx = [torch.from_numpy(x).to(dev) for x in idx_np]
        uniq = float(np.mean([sum(len(np.unique(x[k])) for k in range(ntab)) for x in idx_np]))
        F = ntab + 1
        # rotate gradient buffers so the gradient rows are as cold as the table rows (nb * B * F * D * 4 bytes)
        ngrad = min(a.nb, max(1, int(400e6 // (B * F * D * 4))))
        dTs = [torch.randn((B, F, D), device=dev) * 0.01 for _ in range(ngrad)]
        T = torch.empty((B, F, D), device=dev)
        update_bytes = ntab * (B * D * 4 + B * 4) + 2 * uniq * D * 4
        lookup_bytes = ntab * (B * D * 4 + B * D * 4 + B * 4)

        def measure(opts):
            for k, v in opts.items():
                _lib.set_option(k, v)
            us_sort = _prof.time_launches(lambda i: t.sort(idx[i]), a.nb, a.iters, True)

            def upd(i):
                t.sort(idx[i])
                t.update_sorted(dTs[i % ngrad], 1, 0.01)

            def chain(i):
                t.lookup(idx[i], T, 1, sort=True)
                t.update_sorted(dTs[i % ngrad], 1, 0.01)
            us_both = _prof.time_launches(upd, a.nb, a.iters, True)
            us_chain = _prof.time_launches(chain, a.nb, a.iters, True)
            for k in opts:
                _lib.set_option(k, 0)
            uo = max(us_both - us_sort, 1e-3)
            print(json.dumps({"B": B, "opts": opts, "sort_us": round(us_sort, 2), "sort_plus_update_us": round(us_both, 2),
                              "update_only_us": round(uo, 2), "update_frac_hbm": round(update_bytes / uo / 1e3 / peak, 3),
                              "chain_us": round(us_chain, 2),
                              "chain_frac_hbm": round((update_bytes + lookup_bytes) / us_chain / 1e3 / peak, 3)}), flush=True)

        cases = a.cases.split(",")
        measure({})
        if "tile" in cases:
            for tile in (8, 12, 20, 24, 32):
                measure({"update_tile": tile})
        if "two" in cases:
            measure({"update_two_launches": 1})
        measure({})
        del idx, dTs, T
        torch.cuda.empty_cache()
    t.close()


if __name__ == "__main__":
    main()
