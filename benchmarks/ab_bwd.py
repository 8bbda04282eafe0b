#!/usr/bin/env python
"""A/B of the warp-per-sample interaction backward's variants (`bwd_variant`) in one process: cold inputs
(nb different T / dOut per CUDA graph, together larger than L2), us per launch and fraction of the HBM peak.

    python benchmarks/ab_bwd.py [--variants 1 2 3 0] [--B 2048 16384] [--F 27] [--D 128]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dlrm_jl_b200 import _lib, _prof  # noqa: E402
from dlrm_jl_b200.interact import interaction_bwd, interaction_fwd, interaction_width  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, nargs="*", default=[1, 2, 3, 0])
    ap.add_argument("--B", type=int, nargs="*", default=[2048, 16384])
    ap.add_argument("--F", type=int, default=27)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--fwd", action="store_true", help="also time the forward")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    F, D = a.F, a.D
    w = interaction_width(F, D)
    for B in a.B:
        nb = max(4, min(16, int(600e6 // (B * F * D * 4))))
        Ts = [torch.randn((B, F, D), device=dev) for _ in range(nb)]
        gs = [torch.randn((B, w), device=dev) for _ in range(nb)]
        nbytes = B * (w + 2 * F * D + D) * 4
        ref = None
        for v in a.variants:
            _lib.set_option("bwd_variant", v)
            dx, dT = interaction_bwd(gs[0], Ts[0])
            if ref is None:
                ref = (dx.clone(), dT.clone())
            same = bool(torch.equal(dx, ref[0]) and torch.equal(dT, ref[1]))
            us = _prof.time_launches(lambda i: interaction_bwd(gs[i], Ts[i]), nb, a.iters, True)
            print(json.dumps({"kernel": "interaction_bwd", "B": B, "F": F, "D": D, "variant": v, "us": round(us, 2),
                              "frac_hbm": round(nbytes / us / 1e3 / peak, 3), "same_bits_as_first": same, "nb": nb}), flush=True)
        _lib.set_option("bwd_variant", 0)
        if a.fwd:
            us = _prof.time_launches(lambda i: interaction_fwd(Ts[i]), nb, a.iters, True)
            print(json.dumps({"kernel": "interaction_fwd", "B": B, "us": round(us, 2),
                              "frac_hbm": round(B * (F * D + w) * 4 / us / 1e3 / peak, 3)}), flush=True)
        del Ts, gs


if __name__ == "__main__":
    main()
