#!/usr/bin/env python
"""Where the time of a one-wave kernel goes: per-CTA entry / exit stamps (%globaltimer, dlrmb_clock_enable) of the
hot-path kernels on the Terabyte-shaped batch, each launched alone on cold inputs.

    python benchmarks/cta_timeline.py [--B 2048] [--reps 8]

Prints, per kernel, microseconds from the first CTA's entry: percentiles of the CTA entry times (launch ramp), of
the exit times, of the CTA lifetimes, and for the sparse update the exit of every table's last CTA (the one that
runs the same-launch fix-up).
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

from dlrm_jl_b200 import _lib  # noqa: E402
from dlrm_jl_b200.embedding import EmbeddingTables  # noqa: E402
from dlrm_jl_b200.interact import interaction_bwd, interaction_fwd, interaction_width  # noqa: E402
from dlrm_jl_b200.model import TERABYTE_EMBEDDING_SIZES  # noqa: E402

NAMES = ["lookup", "sort", "update", "update_fixup", "interaction_fwd", "interaction_bwd", "bce"]


def pct(a, qs=(0, 10, 50, 90, 99, 100)):
    return [round(float(np.percentile(a, q)), 2) for q in qs]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--bwd-variants", type=int, nargs="*", default=None, help="run the backward once per listed bwd_variant")
    a = ap.parse_args()
    for kv in a.opt:
        k, v = kv.split("=")
        _lib.set_option(k, int(v))
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    rows, D, B = [min(r, 40_000_000) for r in TERABYTE_EMBEDDING_SIZES], 128, a.B
    ntab, F = len(rows), len(rows) + 1
    t = EmbeddingTables(rows, D, B, dev)
    t.init_uniform(1)
    rng = np.random.default_rng(7)
    nb = a.reps
    idx = [torch.from_numpy(np.stack([rng.integers(0, r, size=(B, 1)) for r in rows]).astype(np.int32)).to(dev) for _ in range(nb)]
    w = interaction_width(F, D)
    Ts = [torch.randn((B, F, D), device=dev) for _ in range(nb)]
    gs = [torch.randn((B, w), device=dev) for _ in range(nb)]
    dTs = [torch.randn((B, F, D), device=dev) * 0.01 for _ in range(nb)]
    Tout = torch.empty((B, F, D), device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB: evicts L2 between launches

    nk = int(lib.dlrmb_clock_kernels())
    n = int(lib.dlrmb_clock_buffer_bytes()) // 8
    buf = torch.empty(n, dtype=torch.int64, device=dev)
    view = buf.view(nk, -1, 2)

    def stamped(fn, which, prep=None):
        """Run fn(i) for every batch with stamps on; returns list of (entry[], exit[]) in us from the first entry."""
        out = []
        for i in range(nb):
            if prep is not None:
                prep(i)
            flush.zero_()
            torch.cuda.synchronize()
            view[:, :, 0] = 1 << 62
            view[:, :, 1] = 0
            _lib.check(lib.dlrmb_clock_enable(buf.data_ptr()))
            fn(i)
            _lib.check(lib.dlrmb_clock_enable(None))
            torch.cuda.synchronize()
            e = view[which, :, 0].cpu().numpy()
            x = view[which, :, 1].cpu().numpy()
            live = x > 0
            e, x = e[live].astype(np.float64), x[live].astype(np.float64)
            o = e.min()
            out.append(((e - o) * 1e-3, (x - o) * 1e-3, np.nonzero(live)[0]))
        return out

    def report(name, runs, extra=None):
        ent = np.mean([pct(r[0]) for r in runs], axis=0)
        ext = np.mean([pct(r[1]) for r in runs], axis=0)
        life = np.mean([pct(r[1] - r[0]) for r in runs], axis=0)
        rec = {"kernel": name, "ctas": int(len(runs[0][0])), "percentiles": [0, 10, 50, 90, 99, 100],
               "entry_us": [round(v, 2) for v in ent], "exit_us": [round(v, 2) for v in ext],
               "lifetime_us": [round(v, 2) for v in life]}
        if extra:
            rec.update(extra(runs))
        print(json.dumps(rec), flush=True)

    # warm-up of every kernel (attribute calls, module load)
    t.lookup(idx[0], Tout, 1, sort=True)
    t.update_sorted(dTs[0], 1, 0.0)
    interaction_fwd(Ts[0])
    interaction_bwd(gs[0], Ts[0])
    torch.cuda.synchronize()

    def lookup_extra(runs):
        # slots 0..ntab-1 are the sort CTAs of the fused launch
        s_exit = np.mean([r[1][r[2] < ntab].max() for r in runs])
        g_exit = np.mean([r[1][r[2] >= ntab].max() for r in runs])
        return {"sort_ctas_last_exit_us": round(float(s_exit), 2), "gather_ctas_last_exit_us": round(float(g_exit), 2)}

    report("lookup_sort", stamped(lambda i: t.lookup(idx[i], Tout, 1, sort=True), 0), lookup_extra)

    def upd_prep(i):
        t.sort(idx[i])

    def upd_extra(runs):
        per_table_last, per_table_rest = [], []
        for e, x, slot in runs:
            chunks = (slot.max() + 1) // ntab
            k = slot // chunks
            last = np.array([x[k == kk].max() for kk in range(ntab)])
            second = np.array([np.sort(x[k == kk])[-2] if (k == kk).sum() > 1 else 0.0 for kk in range(ntab)])
            per_table_last.append(last)
            per_table_rest.append(second)
        return {"per_table_last_exit_us": [round(float(v), 1) for v in np.mean(per_table_last, axis=0)],
                "per_table_second_last_exit_us": [round(float(v), 1) for v in np.mean(per_table_rest, axis=0)],
                "table_rows": rows}

    report("update", stamped(lambda i: t.update_sorted(dTs[i], 1, 0.0), 2, upd_prep), upd_extra)
    report("interaction_fwd", stamped(lambda i: interaction_fwd(Ts[i]), 4))
    if a.bwd_variants:
        for v in a.bwd_variants:
            _lib.set_option("bwd_variant", v)
            interaction_bwd(gs[0], Ts[0])
            report(f"interaction_bwd(variant {v})", stamped(lambda i: interaction_bwd(gs[i], Ts[i]), 5))
        _lib.set_option("bwd_variant", 0)
    else:
        report("interaction_bwd", stamped(lambda i: interaction_bwd(gs[i], Ts[i]), 5))
    t.close()


if __name__ == "__main__":
    main()
