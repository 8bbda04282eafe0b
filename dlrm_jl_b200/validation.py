"""Model import from the PyTorch-reference HDF5 files and the reference's `validate`.

Mirrors `load_hdf5` / `load_inputs` (src/data/criteo.jl:464-560) and `validate`,
`validate_mlp`, `validate_embeddings` (src/validation.jl:1-146).  Adds what upstream lacks: table
export / import (`save_tables`, `load_tables`) so the HBM-resident tables can be checkpointed
(SURVEY.md section 8(f) row 4).
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .embedding import Descent, EmbeddingTables, PreallocationStrategy
from .hdf5_min import read_hdf5
from .interact import DotInteraction
from .model import DLRMModel
from .train import bce_loss, train_step, wrap_loss

Arrays = Dict[str, np.ndarray]


def _natural(name: str):
    """NaturalSort.natural (src/data/criteo.jl:484): digit runs compare as numbers."""
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


def _open(source: Union[str, Arrays]) -> Arrays:
    if isinstance(source, dict):
        return source
    if str(source).endswith(".npz"):
        with np.load(source) as z:
            return {k: z[k] for k in z.files}
    return read_hdf5(str(source))


def _isapprox(a: np.ndarray, b: np.ndarray) -> bool:
    """Julia `isapprox` default for arrays: norm(a-b) <= sqrt(eps(Float32)) * max(norm a, norm b)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return bool(np.linalg.norm(a - b) <= float(np.sqrt(np.finfo(np.float32).eps)) * max(np.linalg.norm(a), np.linalg.norm(b)))


def load_mlp(data: Arrays, prefix_filter: str, device) -> nn.Sequential:
    """src/data/criteo.jl:494-534: layers in natural-sort order of their name prefix; relu on
    every layer except the last layer of the top MLP, which is followed by a sigmoid."""
    names = sorted((k for k in data if k.startswith(prefix_filter)), key=_natural)
    prefixes: List[str] = []
    for n in names:
        p = n.rsplit(".", 1)[0]
        if p not in prefixes:
            prefixes.append(p)
    mods: List[nn.Module] = []
    for p in prefixes:
        W, b = data[f"{p}.weight"], data[f"{p}.bias"]
        lin = nn.Linear(W.shape[1], W.shape[0], device=device)
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(np.ascontiguousarray(W)))
            lin.bias.copy_(torch.from_numpy(np.ascontiguousarray(b)))
        isrelu = p != prefixes[-1] or prefix_filter == "bot_"
        mods += [lin, nn.ReLU() if isrelu else nn.Sigmoid()]
    return nn.Sequential(*mods)


def load_hdf5(source: Union[str, Arrays], device=0, max_lookups: int = 0) -> DLRMModel:
    """``load_hdf5(path)`` (src/data/criteo.jl:464-482): embeddings `emb_*`, MLPs `bot_*` / `top_*`."""
    data = _open(source)
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    emb = [data[k] for k in sorted((k for k in data if k.startswith("emb")), key=_natural)]
    if max_lookups <= 0:
        max_lookups = max(int(data[k].size) for k in data if k.startswith("input_emb"))
    tables = EmbeddingTables.from_arrays(emb, max_lookups, dev)
    return DLRMModel(load_mlp(data, "bot_", dev), tables, DotInteraction(), load_mlp(data, "top_", dev))


def load_inputs(source: Union[str, Arrays]) -> Tuple[np.ndarray, np.ndarray, List[np.ndarray]]:
    """``load_inputs(file)`` (src/data/criteo.jl:536-560) -> (labels, dense, sparse).  Indices stay
    0-based (the files are PyTorch's); a vector longer than the batch is reshaped to [B][P]."""
    data = _open(source)
    labels = data["labels"].reshape(-1)
    dense = data["input_bot"]
    B = labels.shape[0]
    names = sorted((k for k in data if k.startswith("input_emb")), key=_natural)
    sparse = [data[n].reshape(B, -1) for n in names]
    return labels, dense, sparse


def validate(source: Union[str, Arrays], strategy=None, learning_rate: float = 10.0, device=0) -> bool:
    """``validate(path, strategy)`` (src/validation.jl:1-44): inference loss, then ONE SGD step at
    lr = 10.0 and a comparison of every MLP parameter and embedding table with the stored
    post-step values.  Raises RuntimeError with the reference's messages on mismatch."""
    data = _open(source)
    model = load_hdf5(data, device)
    dev = model.embeddings.device
    labels, dense, sparse = load_inputs(data)
    D = model.embeddings.D
    strategy = strategy or PreallocationStrategy(D)
    loss_fn = wrap_loss(bce_loss, strategy=strategy)
    dense_d = torch.from_numpy(np.ascontiguousarray(dense)).to(dev)
    labels_d = torch.from_numpy(np.ascontiguousarray(labels)).to(dev)
    with torch.no_grad():
        l0, _ = loss_fn(model, labels_d, dense_d, sparse)
    loss_ref = float(data["loss"])
    if not np.isclose(loss_ref, float(l0), rtol=float(np.sqrt(np.finfo(np.float32).eps))):
        raise RuntimeError(f"Loss mismatch between this build and PyTorch inference.\nPytorch: {loss_ref}\nHere: {float(l0)}")
    learning_rate = 10.0  # the reference overwrites the keyword (src/validation.jl:23)
    originals = {n: p.detach().cpu().numpy().copy() for n, p in _named_dense(model)}
    train_step(loss_fn, model, Descent(learning_rate), labels_d, dense_d, sparse)
    for key, mlp_name in (("top", "top_mlp"), ("bot", "bottom_mlp")):
        upd = sorted({k.rsplit(".", 1)[0] for k in data if k.startswith(f"update_{key}")}, key=_natural)
        lins = [m for m in getattr(model, mlp_name) if isinstance(m, nn.Linear)]
        if len(upd) != len(lins):
            raise RuntimeError(f"{key} MLP: {len(lins)} layers here, {len(upd)} in the file")
        for i, (lin, u) in enumerate(zip(lins, upd)):
            for what, tensor in (("weight", lin.weight), ("bias", lin.bias)):
                ref = data[f"{u}.{what}"]
                if _isapprox(originals[f"{mlp_name}.{i}.{what}"], ref):
                    raise RuntimeError("Pytorch original and updated weights match!")
                if not _isapprox(ref, tensor.detach().cpu().numpy()):
                    raise RuntimeError(f"{key} MLP layer {i} {what}: updated values differ from PyTorch's")
    emb_names = sorted((k for k in data if k.startswith("emb_")), key=_natural)
    for k, name in enumerate(emb_names):
        ref = data[f"update_{name}"]
        if not _isapprox(ref, model.embeddings.download(k)):
            raise RuntimeError("Updated embeddings don't match PyTorch's!")
        if _isapprox(data[name], ref):
            raise RuntimeError("Pytorch original and updated embeddings match!")
    return True


def _named_dense(model: DLRMModel):
    for mlp_name in ("bottom_mlp", "top_mlp"):
        i = 0
        for m in getattr(model, mlp_name):
            if isinstance(m, nn.Linear):
                yield f"{mlp_name}.{i}.weight", m.weight
                yield f"{mlp_name}.{i}.bias", m.bias
                i += 1


def save_tables(tables: EmbeddingTables, path: str) -> None:
    """Export every table (D2H) to an .npz checkpoint.  No upstream equivalent."""
    np.savez(path, D=np.int64(tables.D), **{f"emb_{k}": tables.download(k) for k in range(tables.ntab)})


def load_tables(path: str, max_lookups: int, device=0) -> EmbeddingTables:
    with np.load(path) as z:
        arrays = [z[k] for k in sorted((k for k in z.files if k.startswith("emb_")), key=_natural)]
    return EmbeddingTables.from_arrays(arrays, max_lookups, device)
