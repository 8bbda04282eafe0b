"""Training-step glue: the embedding branch of `custom_update!` and the step ordering of
`train!` (src/train/train.jl:189-293), plus `bce_loss` and its rrule (:33-71).
"""
from __future__ import annotations

import time
from typing import Callable, Iterable, List, Optional

import torch

from .embedding import Descent, sparse_updates, update_
from .model import DLRMModel, callback, donothing

_EPS32 = float(torch.finfo(torch.float32).eps)


class _BCELoss(torch.autograd.Function):
    """bce_loss + its hand-written pullback (src/train/train.jl:33-41 and :45-71):
    forward clamps the logs at -100, backward uses the eps-regularised quotient."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, y: torch.Tensor):
        ctx.save_for_backward(x, y)
        lx = torch.clamp_min(torch.log(x), -100.0)
        l1x = torch.clamp_min(torch.log(1.0 - x), -100.0)
        return (-y * lx + (y - 1.0) * l1x).sum() / x.numel()

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        d = g / x.numel()
        c = 1.0 - x + _EPS32
        dd = x + _EPS32
        return d * ((1.0 - y) / c - y / dd), None


def bce_loss(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return _BCELoss.apply(x, y)


class _SigmoidBCE(torch.autograd.Function):
    """Final sigmoid + bce_loss + its pullback as ONE kernel launch (csrc/loss.cu)."""

    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor, scratch: torch.Tensor):
        from . import _lib, _prof
        z = logits.reshape(-1).contiguous()
        y = labels.reshape(-1).contiguous()
        if not z.is_cuda:
            raise _lib.DLRMB200Error(_lib.EINVAL, "sigmoid_bce_loss runs on the GPU only (no CPU fallback)")
        dz = torch.empty_like(z)
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        with _prof.range("bce"):
            _lib.check(_lib.load().dlrmb_bce_sigmoid_fwd_bwd(
                z.device.index or 0, z.data_ptr(), y.data_ptr(), z.numel(), None, dz.data_ptr(), loss.data_ptr(),
                scratch.data_ptr(), int(torch.cuda.current_stream(z.device).cuda_stream)))
        ctx.save_for_backward(dz)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return (dz * g).reshape(ctx.shape), None, None


class SigmoidBCELoss:
    """Callable replacing `sigmoid` (last top-MLP layer) + `bce_loss`: takes the top MLP's
    pre-sigmoid logits.  Owns the 65-float scratch the kernel's ordered block reduction needs."""

    def __init__(self, device):
        self.scratch = torch.zeros(80, dtype=torch.float32, device=device)

    def __call__(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return _SigmoidBCE.apply(logits, labels, self.scratch)


class LossWrapper:
    """``wrap_loss(loss_fn; kw...)`` (src/train/train.jl:74-92)."""

    def __init__(self, f: Callable, **kw):
        self.f = f
        self.kw = kw

    def __call__(self, model: DLRMModel, labels, dense, sparse, training: bool = False):
        cb = self.kw.get("cb", donothing)
        out, T = model(dense, sparse, training=training, **self.kw)
        loss = callback(cb, "loss", self.f, out, labels)
        return loss, T


def wrap_loss(loss_fn: Callable, **kw) -> LossWrapper:
    return LossWrapper(loss_fn, **kw)


def custom_update_(opt: Descent, model: DLRMModel, T: torch.Tensor, telemetry: Callable = donothing,
                   presorted: bool = False) -> None:
    """``custom_update!`` (src/train/train.jl:242-293): dense SGD, then the sparse embedding
    update from the lookup buffer's gradient."""
    params = [p for p in model.dense_parameters() if p.grad is not None]
    with torch.no_grad():
        torch._foreach_add_(params, [p.grad for p in params], alpha=-opt.eta)  # Flux.update!: x .-= eta*grad
    telemetry("weight_update_done")
    grads = sparse_updates(T.grad, T.indices, T.slot0, T.idx_base)
    update_(opt, model.embeddings, grads, presorted=presorted)
    telemetry("embedding_update_done")


def train_step(loss: LossWrapper, model: DLRMModel, opt: Descent, labels, dense, sparse,
               overlap_sort: bool = True) -> torch.Tensor:
    """One iteration of the loop body of ``train!`` (src/train/train.jl:215-237).
    Returns the loss tensor (device); no host synchronisation happens here.  Index base and range
    checking travel in the loss wrapper's keywords (``wrap_loss(bce_loss, idx_base=1, check_indices=True)``)."""
    telemetry = loss.kw.get("cb", donothing)
    telemetry("start")
    for p in model.dense_parameters():
        p.grad = None
    l, T = loss(model, labels, dense, sparse, training=True)
    presorted = bool(getattr(T, "presorted", False))     # the lookup launch already sorted the indices
    if overlap_sort and not presorted:
        # the dedup sort needs the indices only: run it beside the backward pass
        model.embeddings.sort(T.indices, T.idx_base, side_stream=True)
        presorted = True
    l.backward()
    for mlp in (model.bottom_mlp, model.top_mlp):      # fused dense layers keep their gradients in buffers
        if hasattr(mlp, "bind_param_grads"):
            mlp.bind_param_grads()
    telemetry("grads_done")
    custom_update_(opt, model, T, telemetry, presorted=presorted)
    telemetry("update_done")
    return l.detach()


def train(loss: LossWrapper, model: DLRMModel, data: Iterable, opt: Descent, cb: Callable = lambda: None,
          maxiters: Optional[int] = None):
    """``train!(loss, model, data, opt; cb, maxiters)`` (src/train/train.jl:189-240).
    ``data`` yields (labels, dense, sparse).  Returns dict(iteration_times [ns], losses).

    If ``data`` carries an ``idx_base`` attribute (a :class:`~dlrm_jl_b200.loader.DACLoader` over a
    reference-preprocessed file is 1-based) and the loss wrapper does not set one, it is adopted; the
    first batch is always range-checked against the table sizes (the kernels are unchecked, like the
    reference's ``@inbounds`` loops), so a base mismatch fails loudly instead of training on shifted rows."""
    losses: List[float] = []
    iteration_times: List[int] = []
    count = 0
    if "idx_base" not in loss.kw and hasattr(data, "idx_base"):
        loss.kw["idx_base"] = int(data.idx_base)
    for labels, dense, sparse in data:
        if count == 0:
            from .embedding import _as_index_tensor
            model.embeddings.check_indices(
                _as_index_tensor(sparse, model.embeddings.ntab, model.embeddings.device), loss.kw.get("idx_base", 0))
        start = time.perf_counter_ns()
        l = train_step(loss, model, opt, labels, dense, sparse)
        losses.append(float(l))  # the reference pushes the loss every iteration (:231); this syncs
        iteration_times.append(time.perf_counter_ns() - start)
        count += 1
        if maxiters is not None and count == maxiters:
            break
        cb()
    cb()
    return {"iteration_times": iteration_times, "losses": losses}
