#!/usr/bin/env python
"""Hot-path-only microbenchmark: embedding lookup, index sort, scatter-add/SGD update and the
dot interaction, each timed alone on the device (BASELINE.json config 5 and the per-kernel
numbers in DESIGN.md).

    python benchmarks/hotpath.py --workload kaggle            # 26 Kaggle tables, D 64, B 2048
    python benchmarks/hotpath.py --workload terabyte          # 26 tables <= 40M rows, D 128
    python benchmarks/hotpath.py --rows 10000000 --D 64 --B 1048576 --P 1 [--zipf 1.05]
    python benchmarks/hotpath.py --sweep                      # rows x D x P grid, one table

Each kernel is captured into a CUDA graph of `--nb` launches over `--nb` different pre-generated
index batches (fresh rows every launch, so L2 does not stand in for HBM) and the graph replay is
timed with CUDA events: per-launch time = replay time / nb, free of host launch gaps.
`--no-graph --iters 1` gives a plain launch sequence for ncu.
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
from dlrm_jl_b200.interact import interaction_bwd, interaction_fwd, interaction_width  # noqa: E402
from dlrm_jl_b200.model import KAGGLE_EMBEDDING_SIZES, TERABYTE_EMBEDDING_SIZES  # noqa: E402


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def zipf_indices(rng, rows: int, n: int, alpha: float) -> np.ndarray:
    """inverse-CDF Zipf rank, then a fixed random permutation of row ids."""
    u = rng.random(n)
    N = float(rows)
    r = np.floor(((N ** (1.0 - alpha) - 1.0) * u + 1.0) ** (1.0 / (1.0 - alpha))).astype(np.int64) - 1
    r = np.clip(r, 0, rows - 1)
    # multiplicative hash as the "fixed random permutation" (bijective mod rows when coprime)
    mult = 2654435761 % rows
    while np.gcd(mult, rows) != 1:
        mult += 1
    return (r * mult + 12345) % rows


def make_indices(rng, rows, B, P, alpha):
    out = []
    for r in rows:
        if alpha and alpha > 0:
            out.append(zipf_indices(rng, r, B * P, alpha).reshape(B, P))
        else:
            out.append(rng.integers(0, r, size=(B, P), dtype=np.int64))
    return np.stack(out).astype(np.int32)


def time_graph(fn, nb: int, use_graph: bool, iters: int) -> float:
    """fn(i) launches batch i.  Returns microseconds per launch."""
    return _prof.time_launches(fn, nb, iters, use_graph)


def run_case(rows, D, B, P, alpha, nb, use_graph, iters, interaction=True, label="", only=None, dtype="f32"):
    dev = torch.device("cuda", 0)
    peak, peak_src = hbm_peak()
    ntab = len(rows)
    L = B * P
    eb = 4 if dtype == "f32" else 2
    t = EmbeddingTables(rows, D, L, dev, dtype=torch.float32 if dtype == "f32" else torch.bfloat16)
    t.init_uniform(1)
    rng = np.random.default_rng(1234)
    idx_np = [make_indices(rng, rows, B, P, alpha) for _ in range(nb)]
    idx = [torch.from_numpy(a).to(dev) for a in idx_np]
    uniq = float(np.mean([sum(len(np.unique(a[k])) for k in range(ntab)) for a in idx_np]))
    F = ntab + 1
    T = torch.empty((B, F, D), device=dev)
    dT = torch.randn((B, F, D), device=dev) * 0.01
    res = {"label": label, "tables": ntab, "rows_total": int(sum(rows)), "rows_max": int(max(rows)), "D": D, "B": B,
           "P": P, "zipf": alpha, "table_dtype": dtype, "lookups_per_launch": ntab * L, "distinct_rows_per_launch": uniq,
           "hbm_peak_gbs": peak, "peak_source": peak_src}

    def rec(name, us, nbytes):
        res[name] = {"us": us, "algorithmic_bytes": int(nbytes), "gbs": nbytes / us / 1e3, "frac_hbm": nbytes / us / 1e3 / peak}

    def interaction_inputs():
        # a different T / dOut per launch, together larger than the 126 MB L2 when nb * B * F * D * 4
        # allows it, so a launch cannot find the previous launch's tile in L2
        w = interaction_width(F, D)
        Ts, gs = [], []
        for i in range(nb):
            Ti = torch.empty((B, F, D), device=dev)
            t.lookup(idx[i], Ti, 1)
            Ti[:, 0] = torch.randn((B, D), device=dev)
            Ts.append(Ti)
            gs.append(torch.randn((B, w), device=dev))
        res["interaction_inputs_mb"] = nb * B * F * D * 4 / 1e6
        return w, Ts, gs

    if only and only.startswith("interaction"):
        w, Ts, gs = interaction_inputs()
        if only == "interaction_fwd":
            rec("interaction_fwd", time_graph(lambda i: interaction_fwd(Ts[i]), nb, use_graph, iters), B * (F * D + w) * 4)
        else:
            rec("interaction_bwd", time_graph(lambda i: interaction_bwd(gs[i], Ts[i]), nb, use_graph, iters), B * (w + 2 * F * D + D) * 4)
        t.close()
        return res
    us = time_graph(lambda i: t.lookup(idx[i], T, 1), nb, use_graph, iters)
    lookup_bytes = ntab * (L * D * eb + B * D * 4 + L * 4)
    rec("lookup", us, lookup_bytes)
    us_sort = time_graph(lambda i: t.sort(idx[i]), nb, use_graph, iters)
    res["sort"] = {"us": us_sort}
    # training-step form: lookup + sort in one launch (the sort rides in extra CTAs when B*P <= 4096)
    rec("lookup_sort", time_graph(lambda i: t.lookup(idx[i], T, 1, sort=True), nb, use_graph, iters), lookup_bytes)

    def chain(i):
        t.lookup(idx[i], T, 1, sort=True)
        t.update_sorted(dT, 1, 0.01)

    def upd(i):
        t.sort(idx[i])
        t.update_sorted(dT, 1, 0.01)
    us_both = time_graph(upd, nb, use_graph, iters)
    update_bytes = ntab * (B * D * 4 + L * 4) + 2 * uniq * D * eb
    rec("sort_plus_update", us_both, update_bytes)
    rec("update_only", max(us_both - us_sort, 1e-3), update_bytes)
    rec("embedding_lookup_plus_update", res["lookup"]["us"] + us_both, lookup_bytes + update_bytes)
    # the BASELINE metric as the step runs it: (lookup + sort) launch, then the update launch, back to back
    rec("embedding_chain", time_graph(chain, nb, use_graph, iters), lookup_bytes + update_bytes)
    if interaction and D % 4 == 0 and F <= 64:
        w, Ts, gs = interaction_inputs()
        us = time_graph(lambda i: interaction_fwd(Ts[i]), nb, use_graph, iters)
        rec("interaction_fwd", us, B * (F * D + w) * 4)
        res["interaction_fwd"]["gflops_useful"] = 2.0 * B * (F * (F - 1) // 2) * D / us / 1e3
        us = time_graph(lambda i: interaction_bwd(gs[i], Ts[i]), nb, use_graph, iters)
        rec("interaction_bwd", us, B * (w + 2 * F * D + D) * 4)
        res["interaction_bwd"]["gflops"] = 2.0 * B * F * F * D / us / 1e3
    res["options"] = {k: _lib.get_option(k) for k in ("interact_general", "update_two_launches", "update_tile")}
    t.close()
    del T, dT, idx
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=["kaggle", "terabyte"])
    ap.add_argument("--rows", type=int, nargs="*")
    ap.add_argument("--D", type=int, default=64)
    ap.add_argument("--B", type=int, default=2048)
    ap.add_argument("--P", type=int, default=1)
    ap.add_argument("--zipf", type=float, default=0.0)
    ap.add_argument("--nb", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--full", action="store_true", help="with --sweep: all of P in {1,4,16,64} and Zipf in {0,1.05,1.2}")
    ap.add_argument("--out", default=None)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"], help="table storage type")
    ap.add_argument("--only", default=None, help="time just this kernel (e.g. interaction_fwd)")
    ap.add_argument("--no-interaction", action="store_true", help="embedding kernels only")
    ap.add_argument("--opt", action="append", default=[], help="library switch name=value (dlrmb_set_option)")
    ap.add_argument("--small-tables", action="store_true",
                    help="cap every table at 1000 rows (the interaction kernels do not depend on table size; keeps ncu replays cheap)")
    a = ap.parse_args()
    for kv in a.opt:
        name, value = kv.split("=")
        _lib.set_option(name, int(value))
    results = []
    if a.sweep:
        for rows in (100_000, 1_000_000, 10_000_000, 100_000_000):
            for D in (16, 32, 64, 128, 256):
                if rows * D * 4 > 120e9:
                    continue
                for P in ((1, 4, 16, 64) if a.full else (1, 16)):
                    B = (1 << 20) // P
                    for alpha in ((0.0, 1.05, 1.2) if a.full else (0.0, 1.05)):
                        r = run_case([rows], D, B, P, alpha, 3, not a.no_graph, 3, interaction=False,
                                     label=f"sweep rows={rows} D={D} P={P} zipf={alpha}")
                        results.append(r)
                        print(json.dumps(r), flush=True)
    else:
        if a.workload == "kaggle":
            rows, D = list(KAGGLE_EMBEDDING_SIZES), 64
        elif a.workload == "terabyte":
            rows, D = [min(r, 40_000_000) for r in TERABYTE_EMBEDDING_SIZES], 128
        else:
            rows, D = a.rows or [1_000_000], a.D
        if a.workload and a.D != 64:
            D = a.D
        if a.small_tables:
            rows = [min(r, 1000) for r in rows]
        r = run_case(rows, D, a.B, a.P, a.zipf, a.nb, not a.no_graph, a.iters, interaction=not a.no_interaction,
                     label=a.workload or "custom", only=a.only, dtype=a.dtype)
        results.append(r)
        print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            for r in results:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
