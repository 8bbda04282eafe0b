"""Model assembly: the reference's DLRMModel functor and `dlrm` builder (src/model/model.jl).

Only the two hot-path call sites are ours -- `maplookup` (:161) and the interaction (:163);
the dense MLPs stay library GEMMs (torch.nn.Linear -> cuBLAS, fp32, TF32 off), standing in for
the reference's OneDNN.Dense layers.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn

from .embedding import DefaultStrategy, EmbeddingTables, PreallocationStrategy, maplookup
from .interact import DotInteraction, POST_INTERACTION_PAD_TO_MUL, dot_interaction, up_to_mul_of

# src/data/criteo.jl:350-377 and :379-406
KAGGLE_EMBEDDING_SIZES = [
    1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
    5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572,
]
TERABYTE_EMBEDDING_SIZES = [
    227605432, 39060, 17295, 7424, 20265, 3, 7122, 1543, 63, 130229467, 3067956, 405282, 10, 2209,
    11938, 155, 4, 976, 14, 292775614, 40790948, 187188510, 590152, 12973, 108, 36,
]


def donothing(*_a, **_k):
    """src/utils/utils.jl `donothing`."""


def callback(cb: Callable, sym: str, f: Callable, *args, **kw):
    """src/model/model.jl:131-149: run f, fire cb(sym); the backward symbol `<sym>_back` is fired
    by a gradient hook on the result."""
    y = f(*args, **kw)
    if cb is not donothing:
        cb(sym)
        t = y if isinstance(y, torch.Tensor) else None
        if t is not None and t.requires_grad:
            t.register_hook(lambda g, _s=sym: (cb(f"{_s}_back"), g)[1])
    return y


def create_mlp(sizes: Sequence[int], sigmoid_index: int, device, generator: Optional[torch.Generator] = None) -> nn.Sequential:
    """src/model/model.jl:72-93: relu after every layer, except layer `sigmoid_index - 1`
    (1-based) which gets a sigmoid.  GlorotNormal weights (:57-58), zero bias."""
    layers: List[nn.Module] = []
    for i in range(1, len(sizes)):
        fan_in, fan_out = sizes[i - 1], sizes[i]
        lin = nn.Linear(fan_in, fan_out, bias=True, device=device, dtype=torch.float32)
        with torch.no_grad():
            std = math.sqrt(2.0 / (fan_in + fan_out))
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=generator, dtype=torch.float32).to(device) * std)
            lin.bias.zero_()
        layers.append(lin)
        layers.append(nn.Sigmoid() if i == sigmoid_index - 1 else nn.ReLU())
    return nn.Sequential(*layers)


class DLRMModel:
    """``DLRMModel(bottom_mlp, embeddings, interaction, top_mlp)`` (src/model/model.jl:117-122)."""

    def __init__(self, bottom_mlp: nn.Module, embeddings: EmbeddingTables, interaction, top_mlp: nn.Module):
        self.bottom_mlp = bottom_mlp
        self.embeddings = embeddings
        self.interaction = interaction
        self.top_mlp = top_mlp

    def dense_parameters(self) -> List[torch.nn.Parameter]:
        return list(self.bottom_mlp.parameters()) + list(self.top_mlp.parameters())

    def __call__(self, dense: torch.Tensor, sparse, strategy=None, cb=donothing, idx_base: int = 0,
                 training: bool = False, check_indices: bool = False):
        """The functor, src/model/model.jl:152-166.  Returns (out [B], T) where T is the lookup
        buffer (its ``.grad`` after backward is the sparse gradient source).  ``idx_base`` is 1 for
        reference-format data (1-based ids, src/data/criteo.jl:249-253), 0 for PyTorch-style ids;
        ``check_indices`` validates the batch against the table sizes before the unchecked kernels run."""
        D = self.embeddings.D
        if strategy is None:
            strategy = PreallocationStrategy(D)
        y = callback(cb, "lookup", maplookup, strategy, self.embeddings, sparse, idx_base, training, check_indices)
        x = callback(cb, "bottom_mlp", self.bottom_mlp, dense)
        if isinstance(strategy, DefaultStrategy):
            z = callback(cb, "interaction", dot_interaction, x, y)
        else:
            z = callback(cb, "interaction", self.interaction, x, y)
        out = callback(cb, "top_mlp", self.top_mlp, z)
        return out.reshape(-1), y


def dlrm(bottom_mlp_sizes: Sequence[int], top_mlp_sizes: Sequence[int], sparse_feature_size: int,
         embedding_sizes: Sequence[int], *, max_lookups: int, device=0, interaction=None,
         seed: int = 51234, init_tables: bool = True, fused_dense: bool = False) -> DLRMModel:
    """``dlrm(bottom, top, feature_size, embedding_sizes; ...)`` (src/model/model.jl:173-233)."""
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    gen = torch.Generator().manual_seed(seed)  # Random.seed!(51234), :193
    bottom = create_mlp(bottom_mlp_sizes, 0, dev, gen)
    tables = EmbeddingTables(embedding_sizes, sparse_feature_size, max_lookups, dev)
    if init_tables:
        tables.init_uniform(seed)
    num_features = len(embedding_sizes)
    bottom_out = bottom_mlp_sizes[-1]
    assert (sparse_feature_size * num_features) % bottom_out == 0  # :220
    pre_triangle = (sparse_feature_size * num_features) // bottom_out + 1
    top_in = up_to_mul_of((pre_triangle * pre_triangle - pre_triangle) // 2 + bottom_out, POST_INTERACTION_PAD_TO_MUL)
    sizes = [top_in, *top_mlp_sizes]
    top = create_mlp(sizes, len(sizes), dev, gen)  # sigmoid on the last layer, :230
    if fused_dense:   # OneDNN.Dense-style fused layers (dlrm_jl_b200.dense), same parameters
        from .dense import FusedMLP
        bottom, top = FusedMLP(bottom), FusedMLP(top)
    return DLRMModel(bottom, tables, interaction or DotInteraction(), top)


def kaggle_dlrm(*, feature_size: int = 16, max_lookups: int = 2048, device=0,
                embedding_sizes: Sequence[int] = KAGGLE_EMBEDDING_SIZES) -> DLRMModel:
    """``kaggle_dlrm`` (src/data/criteo.jl:408-433)."""
    return dlrm([13, 512, 256, feature_size], [1024, 1024, 512, 256, 1], feature_size, embedding_sizes,
                max_lookups=max_lookups, device=device, interaction=DotInteraction())
