"""Shared loaders for the golden fixtures (tests only)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str):
    """name in {"single", "multi"} -> dict of arrays (see tests/golden/make_golden.py)."""
    with np.load(os.path.join(GOLDEN, f"pytorch_reference_{name}.npz")) as z:
        return {k: z[k] for k in z.files}


def golden_model(g):
    """Split a golden dict into (bot, top, tables, dense, idx, labels) in oracle layout."""
    bot = [(g[f"bot_l.{i}.weight"], g[f"bot_l.{i}.bias"]) for i in range(4)]
    top = [(g[f"top_l.{i}.weight"], g[f"top_l.{i}.bias"]) for i in range(3)]
    tables = [g[f"emb_{k}"].copy() for k in range(7)]
    B = g["labels"].shape[0]
    idx = [g[f"input_emb_{k}"].reshape(B, -1) for k in range(7)]
    return bot, top, tables, g["input_bot"], idx, g["labels"].reshape(-1)


def golden_updates(g):
    ubot = [(g[f"update_bot_{2 * i}.weight"], g[f"update_bot_{2 * i}.bias"]) for i in range(4)]
    utop = [(g[f"update_top_{2 * i}.weight"], g[f"update_top_{2 * i}.bias"]) for i in range(3)]
    uemb = [g[f"update_emb_{k}"] for k in range(7)]
    return ubot, utop, uemb


def load_known_answer():
    with open(os.path.join(GOLDEN, "known_answer_small.json")) as fh:
        d = json.load(fh)
    return {k: (np.asarray(v, dtype=np.float32) if k != "py_sparse_input" and k != "py_embedding_outputs" else v)
            for k, v in d.items()}
