"""CPU restatement (numpy) of DLRM.jl's embedding + dot-interaction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dlrm_jl_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs do,
and there only as the checker.

Parity status
-------------
* Interaction, BCE loss, the training-step ordering: restated from in-tree reference
  source (file:line given per function, relative to the DLRM.jl checkout) and pinned by the
  reference's own known-answer vectors (3x3 triangle example, the hand-typed 4-sample PyTorch
  case, both ``ref/pytorch_reference_*.hdf5`` files) in ``tests/test_oracle_golden.py``.
* Embedding lookup / sparse gradient / sparse SGD: the arithmetic lives in
  EmbeddingTables.jl v0.1.0, pinned in the reference's ``Manifest.toml:246-250`` as
  ``path = "../EmbeddingTables"`` (no SHA, not vendored, absent here).  Restated from the
  reference's call sites and tests; VALUES are pinned by the goldens
  (``concatenated_result`` for lookup+pool, ``update_emb_*`` for the post-SGD tables,
  ``uncompress`` semantics from ``test/train/backprop.jl:148-158``).
* The dedup ORDER (sorted unique ids, segment offsets, stable permutation) is defined by
  this build, not by the reference (which dedups through a dictionary, so no reference
  output observes an order): **parity unpinned** for those three integer arrays beyond
  "a correct stable sort"; they are checked bit-exactly against numpy's stable argsort.

Layout: every array is C-order.  A Julia ``D x N`` column-major matrix is the C array
``[N][D]``.  Indices are 0-based here (``index_base`` is handled at the C-ABI).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

F32 = np.float32
EPS32 = np.float32(np.finfo(np.float32).eps)  # Julia eps(Float32), src/train/train.jl:54


# --------------------------------------------------------------------------------------
# Triangle index math -- line-by-line loop restatements (small cases only)
# --------------------------------------------------------------------------------------
def triangular_slice_kernel(x: np.ndarray) -> np.ndarray:
    """src/model/interact.jl:64-75.  ``x`` is the Julia matrix given as ``x[i][j]``
    (i = Julia row, j = Julia column); returns the strict upper triangle walked column by
    column: for j = 2..sz, i = 1..j-1."""
    sz = x.shape[1]
    y = np.zeros(sz * (sz - 1) // 2, dtype=x.dtype)
    yindex = 0
    for jcol in range(1, sz):          # Julia `i in 1:sz-1` selects column i+1
        for irow in range(jcol):       # Julia `j in 1:i` selects row j
            y[yindex + irow] = x[irow, jcol]
        yindex += jcol
    return y


def triangular_slice_back_kernel(y: np.ndarray, sz: int) -> np.ndarray:
    """src/model/interact.jl:103-120: scatter ``y`` to the strict upper triangle, zero the
    rest.  1-based ``m = i + ((j-2)(j-1))>>1``."""
    x = np.zeros((sz, sz), dtype=y.dtype)
    for j in range(1, sz + 1):
        for i in range(1, sz + 1):
            if i < j:
                m = i + (((j - 2) * (j - 1)) >> 1)
                x[i - 1, j - 1] = y[m - 1]
    return x


def triangular_slice_back_fuse_add_transpose_kernel(y: np.ndarray, sz: int) -> np.ndarray:
    """src/model/interact.jl:154-173: symmetric matrix with zero diagonal."""
    x = np.zeros((sz, sz), dtype=y.dtype)
    for j in range(1, sz + 1):
        for i in range(1, sz + 1):
            if i == j:
                continue
            if i > j:
                m = j + (((i - 2) * (i - 1)) >> 1)
            else:
                m = i + (((j - 2) * (j - 1)) >> 1)
            x[i - 1, j - 1] = y[m - 1]
    return x


def num_pairs(F: int) -> int:
    return F * (F - 1) // 2


def cdiv(x: int, y: int) -> int:
    """src/model/model.jl:33"""
    return 1 + (x - 1) // y


def up_to_mul_of(x: int, y: int) -> int:
    """src/model/model.jl:34"""
    return y * cdiv(x, y)


def interaction_out_width(F: int, d: int, pad_to_mul: int = 1) -> Tuple[int, int]:
    """src/model/interact.jl:453-455 -> (padded width, padding)."""
    unpadded = num_pairs(F) + d
    padded = up_to_mul_of(unpadded, pad_to_mul)
    return padded, padded - unpadded


# --------------------------------------------------------------------------------------
# Dot interaction
# --------------------------------------------------------------------------------------
def interaction_fwd(T: np.ndarray, pad_to_mul: int = 1) -> np.ndarray:
    """DotInteraction forward, src/model/interact.jl:394-411,449-467,338-362.

    T: [B][F][d] with slot 0 = bottom-MLP output x (the reference copies x there with
    ``fast_vcat``, :271-281).  Returns [B][d + F(F-1)/2 + pad]:
    ``out[b] = [x_b ; <T_b[i], T_b[j]> for j=1..F-1, i=0..j-1 ; 0-pad]``.
    """
    T = np.ascontiguousarray(T, dtype=F32)
    B, F, d = T.shape
    width, _pad = interaction_out_width(F, d, pad_to_mul)
    out = np.zeros((B, width), dtype=F32)
    out[:, :d] = T[:, 0, :]
    G = np.einsum("bik,bjk->bij", T, T, dtype=F32)       # gemmavx!(scratch, T', T), :358
    jj, ii = np.tril_indices(F, -1)                      # (1,0),(2,0),(2,1),(3,0)...
    out[:, d:d + num_pairs(F)] = G[:, jj, ii]
    return out


def interaction_fwd_loops(T: np.ndarray, pad_to_mul: int = 1) -> np.ndarray:
    """Same, as explicit loops with k ascending (the `gemmavx!` triple loop, :318-326)."""
    B, F, d = T.shape
    width, _pad = interaction_out_width(F, d, pad_to_mul)
    out = np.zeros((B, width), dtype=F32)
    for b in range(B):
        out[b, :d] = T[b, 0]
        for j in range(1, F):
            for i in range(j):
                acc = F32(0)
                for k in range(d):
                    acc = F32(acc + F32(T[b, i, k] * T[b, j, k]))
                out[b, d + j * (j - 1) // 2 + i] = acc
    return out


def interaction_bwd(dOut: np.ndarray, T: np.ndarray, pad: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """DotInteraction pullback, src/model/interact.jl:424-436,469-489,154-173,329-336.

    dOut: [B][d + pairs + pad], T: [B][F][d] saved from forward.
    Returns (dx [B][d], dT [B][F][d]); dT keeps slot 0 (the reference returns the whole
    (d*F) x B matrix as dy) and dx = dOut[:, :d] + dT[:, 0].
    """
    dOut = np.ascontiguousarray(dOut, dtype=F32)
    T = np.ascontiguousarray(T, dtype=F32)
    B, F, d = T.shape
    npair = num_pairs(F)
    assert dOut.shape == (B, d + npair + pad)
    S = np.zeros((B, F, F), dtype=F32)
    jj, ii = np.tril_indices(F, -1)
    tri = dOut[:, d:d + npair]
    S[:, jj, ii] = tri
    S[:, ii, jj] = tri
    dT = np.einsum("bjf,bjk->bfk", S, T, dtype=F32)      # gemmavx!(dT_b, T_b, S), :486
    dx = (dOut[:, :d] + dT[:, 0, :]).astype(F32)         # sumavx, :434
    return dx, dT


# --------------------------------------------------------------------------------------
# Embedding lookup / sparse gradient / sparse SGD  (EmbeddingTables.jl contract)
# --------------------------------------------------------------------------------------
def lookup(tables: Sequence[np.ndarray], idx: Sequence[np.ndarray], slot0: int = 0) -> np.ndarray:
    """maplookup(PreallocationStrategy, tables, idx): call site src/model/model.jl:161,
    semantics pinned by test/model/model.jl:265-271 and the `concatenated_result` golden.

    tables[k]: [rows_k][D]; idx[k]: [B] or [B][P] 0-based (sample b owns P consecutive
    entries, src/data/criteo.jl:551-557).  Returns [B][slot0 + ntab][D]; slots < slot0 are
    left zero (reserved for x).  Pool = sum over p in ascending p.
    """
    ntab = len(tables)
    D = tables[0].shape[1]
    B = idx[0].shape[0]
    out = np.zeros((B, slot0 + ntab, D), dtype=F32)
    for k in range(ntab):
        ik = np.asarray(idx[k]).reshape(B, -1)
        acc = tables[k][ik[:, 0]].astype(F32)
        for p in range(1, ik.shape[1]):
            acc = (acc + tables[k][ik[:, p]]).astype(F32)
        out[:, slot0 + k, :] = acc
    return out


def sort_dedup(idx_flat: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Integer oracle for the build-defined dedup: (uniq ascending, seg_offsets = exclusive
    prefix sum of multiplicities with a trailing total, perm = stable argsort by id)."""
    idx_flat = np.asarray(idx_flat).reshape(-1).astype(np.int64)
    perm = np.argsort(idx_flat, kind="stable").astype(np.int32)
    uniq, counts = np.unique(idx_flat, return_counts=True)
    seg = np.zeros(len(uniq) + 1, dtype=np.int32)
    np.cumsum(counts, out=seg[1:])
    return uniq.astype(np.int64), seg, perm


def uncompress(delta: np.ndarray, idx: np.ndarray, nrows: int) -> np.ndarray:
    """EmbeddingTables.uncompress(SparseEmbeddingUpdate, nrows), test/train/backprop.jl:156:
    dense gradient ``grad[r] = sum_{(b,p): idx[b,p]=r} delta[b]``, accumulated in ascending
    flat position (b major, p minor)."""
    B, D = delta.shape
    ik = np.asarray(idx).reshape(B, -1)
    grad = np.zeros((nrows, D), dtype=F32)
    for b in range(B):
        for p in range(ik.shape[1]):
            grad[ik[b, p]] = grad[ik[b, p]] + delta[b]
    return grad


def sparse_sgd_update(table: np.ndarray, idx: np.ndarray, delta: np.ndarray, lr: float) -> None:
    """EmbeddingTables.update!(Flux.Descent(lr), table, SparseEmbeddingUpdate(delta, idx)),
    call site src/train/train.jl:283-290; Flux Descent = ``D .*= lr; x .-= D``.
    In place.  Duplicates accumulate in ascending flat position before the single RMW."""
    B, D = delta.shape
    ik = np.asarray(idx).reshape(B, -1)
    P = ik.shape[1]
    uniq, seg, perm = sort_dedup(ik)
    lr32 = F32(lr)
    for u in range(len(uniq)):
        acc = np.zeros(D, dtype=F32)
        for s in range(seg[u], seg[u + 1]):
            acc = (acc + delta[perm[s] // P]).astype(F32)
        table[uniq[u]] = (table[uniq[u]] - (lr32 * acc).astype(F32)).astype(F32)


def sparse_sgd_update_fast(table: np.ndarray, idx: np.ndarray, delta: np.ndarray, lr: float) -> None:
    """Vectorised variant of ``sparse_sgd_update`` for mid-size cases (np.add.at adds in
    ascending flat position too, so the association is identical)."""
    B, D = delta.shape
    ik = np.asarray(idx).reshape(B, -1)
    P = ik.shape[1]
    uniq, inv = np.unique(ik.reshape(-1), return_inverse=True)
    acc = np.zeros((len(uniq), D), dtype=F32)
    np.add.at(acc, inv, np.repeat(delta, P, axis=0) if P > 1 else delta)
    table[uniq] = (table[uniq] - (F32(lr) * acc).astype(F32)).astype(F32)


# --------------------------------------------------------------------------------------
# Loss + MLPs (out of the CUDA scope; needed to reproduce `validate` end to end)
# --------------------------------------------------------------------------------------
def bce_loss(x: np.ndarray, y: np.ndarray) -> np.float32:
    """src/train/train.jl:33-41: mean BCE with log clamped at -100."""
    x = x.astype(F32).reshape(-1)
    y = y.astype(F32).reshape(-1)
    s = -y * np.maximum(np.log(x), F32(-100)) + (y - F32(1)) * np.maximum(np.log(F32(1) - x), F32(-100))
    return F32(s.astype(F32).sum(dtype=F32) / F32(x.size))


def bce_loss_back(x: np.ndarray, y: np.ndarray, delta: float = 1.0) -> np.ndarray:
    """src/train/train.jl:45-71 (dx only)."""
    x = x.astype(F32).reshape(-1)
    y = y.astype(F32).reshape(-1)
    dl = F32(delta) / F32(x.size)
    c = F32(1) - x + EPS32
    dd = x + EPS32
    return (dl * ((F32(1) - y) / c - y / dd)).astype(F32)


def sigmoid_bce(logits: np.ndarray, y: np.ndarray):
    """Last top-MLP activation (sigmoid, src/model/model.jl:87-90) + bce_loss + its pullback
    chained through the sigmoid: returns (loss, dloss/dlogits)."""
    z = logits.astype(F32).reshape(-1)
    x = (F32(1) / (F32(1) + np.exp(-z))).astype(F32)
    dx = bce_loss_back(x, y)
    return bce_loss(x, y), (dx * x * (F32(1) - x)).astype(F32)


def mlp_forward(layers: List[Tuple[np.ndarray, np.ndarray]], x: np.ndarray, sigmoid_last: bool):
    """Dense chain, weights PyTorch-oriented [out][in] (src/data/criteo.jl:494-534): relu on
    every layer; for the top MLP the last layer is sigmoid instead."""
    acts = [x.astype(F32)]
    pre = []
    n = len(layers)
    for i, (W, b) in enumerate(layers):
        z = (acts[-1] @ W.T.astype(F32) + b.astype(F32)).astype(F32)
        pre.append(z)
        if sigmoid_last and i == n - 1:
            a = (F32(1) / (F32(1) + np.exp(-z))).astype(F32)
        else:
            a = np.maximum(z, F32(0))
        acts.append(a)
    return acts, pre


def mlp_backward(layers, acts, pre, dout: np.ndarray, sigmoid_last: bool):
    grads = [None] * len(layers)
    g = dout.astype(F32)
    n = len(layers)
    for i in range(n - 1, -1, -1):
        W, _b = layers[i]
        if sigmoid_last and i == n - 1:
            a = acts[i + 1]
            g = (g * a * (F32(1) - a)).astype(F32)
        else:
            g = (g * (pre[i] > 0)).astype(F32)
        grads[i] = ((g.T @ acts[i]).astype(F32), g.sum(axis=0, dtype=F32))
        g = (g @ W.astype(F32)).astype(F32)
    return grads, g


def dlrm_forward(bot, top, tables, dense, idx):
    """DLRMModel functor, src/model/model.jl:152-166 with PreallocationStrategy(d)."""
    bacts, bpre = mlp_forward(bot, dense, sigmoid_last=False)
    x = bacts[-1]
    T = lookup(tables, idx, slot0=1)          # :161
    T[:, 0, :] = x                            # fast_vcat, interact.jl:271-281
    z = interaction_fwd(T)                    # :163
    tacts, tpre = mlp_forward(top, z, sigmoid_last=True)
    out = tacts[-1].reshape(-1)               # :165
    return dict(x=x, T=T, z=z, out=out, bacts=bacts, bpre=bpre, tacts=tacts, tpre=tpre)


def dlrm_train_step(bot, top, tables, dense, idx, labels, lr):
    """One iteration of train!, src/train/train.jl:215-237 (fwd+bwd, dense SGD, sparse SGD).
    Mutates ``tables`` in place; returns (loss, new_bot, new_top, fwd intermediates, grads)."""
    fwd = dlrm_forward(bot, top, tables, dense, idx)
    loss = bce_loss(fwd["out"], labels)
    dout = bce_loss_back(fwd["out"], labels).reshape(-1, 1)
    tgrads, dz = mlp_backward(top, fwd["tacts"], fwd["tpre"], dout, sigmoid_last=True)
    dx, dT = interaction_bwd(dz, fwd["T"])
    bgrads, _ = mlp_backward(bot, fwd["bacts"], fwd["bpre"], dx, sigmoid_last=False)
    lr32 = F32(lr)
    new_bot = [((W - lr32 * gW).astype(F32), (b - lr32 * gb).astype(F32)) for (W, b), (gW, gb) in zip(bot, bgrads)]
    new_top = [((W - lr32 * gW).astype(F32), (b - lr32 * gb).astype(F32)) for (W, b), (gW, gb) in zip(top, tgrads)]
    for k, table in enumerate(tables):
        sparse_sgd_update_fast(table, idx[k], np.ascontiguousarray(dT[:, 1 + k, :]), lr)
    return loss, new_bot, new_top, fwd, dict(dT=dT, dx=dx, dz=dz, bgrads=bgrads, tgrads=tgrads)


def isapprox(a: np.ndarray, b: np.ndarray, rtol: float | None = None) -> bool:
    """Julia isapprox default for arrays: norm(a-b) <= sqrt(eps(f32)) * max(norm a, norm b)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    if rtol is None:
        rtol = float(np.sqrt(np.finfo(np.float32).eps))
    return bool(np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2 in float64 (the measure the parity tests gate on)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a - b))


# --------------------------------------------------------------------------------------
# Batch marshalling
# --------------------------------------------------------------------------------------
def dac_unpack(records: np.ndarray):
    """`load!(labels, dense, sparse, records)`, src/data/criteo.jl:284-310: per record copy the
    label, the 13 continuous values into dense[13 x B] (C [B][13]) and the 26 categorical values
    into sparse[B x 26] (C [26][B])."""
    labels = records["label"].astype(F32)
    dense = np.ascontiguousarray(records["continuous"], dtype=F32)
    sparse = np.ascontiguousarray(records["categorical"].T)
    return labels, dense, sparse


# --------------------------------------------------------------------------------------
# BF16 table storage
# --------------------------------------------------------------------------------------
def to_bf16(x: np.ndarray) -> np.ndarray:
    """Round float32 to bfloat16 (nearest even) and return it widened back to float32: what a
    table stored as BFloat16 (the reference's `embedding_eltype`, src/model/model.jl:187) holds."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(F32).reshape(np.shape(x))


def sparse_sgd_update_bf16(table: np.ndarray, idx: np.ndarray, delta: np.ndarray, lr: float) -> None:
    """sparse_sgd_update_fast for a bf16-stored table: fp32 arithmetic, the written rows rounded."""
    B, D = delta.shape
    ik = np.asarray(idx).reshape(B, -1)
    P = ik.shape[1]
    uniq, inv = np.unique(ik.reshape(-1), return_inverse=True)
    acc = np.zeros((len(uniq), D), dtype=F32)
    np.add.at(acc, inv, np.repeat(delta, P, axis=0) if P > 1 else delta)
    table[uniq] = to_bf16((table[uniq] - (F32(lr) * acc).astype(F32)).astype(F32))
