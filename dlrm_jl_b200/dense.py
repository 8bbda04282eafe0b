"""Fused dense layers: the host-side mirror of ``OneDNN.Dense(Flux.Dense(in, out, relu))``, the layer
type DLRM.jl's ``create_mlp`` stacks (src/model/model.jl:72-93).

The GEMMs are library calls (torch / cuBLAS, fp32, TF32 off) exactly as the reference leaves them to
oneDNN.  What is fused here is the glue a framework otherwise spends a dozen launches per layer on:

* forward: bias add and relu are one in-place pass over the GEMM output (``dlrmb_dense_fwd_bias_act``;
  cuBLASLt's fp32 SIMT GEMMs run their own bias / relu epilogue as a separate, slower kernel);
* backward: relu mask + bias gradient in ONE launch of this repo's ``dlrmb_dense_bwd_act_bias`` kernel
  (csrc/dense.cu), the weight gradient GEMM writes straight into the caller's gradient buffer (for the
  data-parallel step: a view of the flat all-reduce bucket), so there is no gradient accumulation
  pass and no per-parameter ``.grad`` tensor.

``FusedMLP`` wraps an ``nn.Sequential`` built by :func:`dlrm_jl_b200.model.create_mlp` and shares its
parameters; results equal the unfused module up to the GEMM epilogue's rounding.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib


def _stream(t: torch.Tensor) -> int:
    return int(torch.cuda.current_stream(t.device).cuda_stream)


class _DenseFn(torch.autograd.Function):
    """y = act(x W^T + b); backward writes dW into `gw`, db into `gb` and returns only dx."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, relu: bool,
                gw: torch.Tensor, gb: torch.Tensor, scratch: torch.Tensor):
        if not x.is_cuda:
            raise _lib.DLRMB200Error(_lib.EINVAL, "fused dense layers run on the GPU only (no CPU fallback)")
        x = x.contiguous()
        y = torch.mm(x, weight.t())                            # plain library GEMM
        _lib.check(_lib.load().dlrmb_dense_fwd_bias_act(       # + bias, relu: one in-place pass
            y.device.index or 0, y.data_ptr(), bias.data_ptr(), y.shape[0], y.shape[1], 1 if relu else 0, _stream(y)))
        ctx.relu = relu
        ctx.gw, ctx.gb, ctx.scratch = gw, gb, scratch
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        x, weight, y = ctx.saved_tensors
        dy = dy.contiguous()
        B, N = dy.shape
        dz = torch.empty_like(dy) if ctx.relu else dy
        _lib.check(_lib.load().dlrmb_dense_bwd_act_bias(
            dy.device.index or 0, dy.data_ptr(), y.data_ptr() if ctx.relu else None, B, N, dz.data_ptr(),
            ctx.gb.data_ptr(), ctx.scratch.data_ptr(), _stream(dy)))
        torch.mm(dz.t(), x, out=ctx.gw)                       # dW [out][in], written in place
        dx = dz @ weight if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None, None, None


class FusedDense(nn.Module):
    """One ``Dense(in, out, act)``; `linear` supplies (and keeps owning) weight and bias."""

    def __init__(self, linear: nn.Linear, relu: bool, grad_weight: Optional[torch.Tensor] = None,
                 grad_bias: Optional[torch.Tensor] = None):
        super().__init__()
        self.linear = linear
        self.relu = relu
        dev = linear.weight.device
        self.grad_weight = grad_weight if grad_weight is not None else torch.zeros_like(linear.weight)
        self.grad_bias = grad_bias if grad_bias is not None else torch.zeros_like(linear.bias)
        assert self.grad_weight.is_contiguous() and self.grad_weight.shape == linear.weight.shape
        assert self.grad_bias.is_contiguous() and self.grad_bias.shape == linear.bias.shape
        n = int(_lib.load().dlrmb_dense_bwd_scratch_floats(linear.out_features))
        self.scratch = torch.zeros(n, dtype=torch.float32, device=dev)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _DenseFn.apply(x, self.linear.weight, self.linear.bias, self.relu, self.grad_weight,
                              self.grad_bias, self.scratch)


class FusedMLP(nn.Module):
    """``create_mlp`` stack with fused layers.  `seq` = [Linear, act, Linear, act, ...] where act is
    ReLU (fused), Sigmoid (applied unfused after the layer) or absent for a trailing Linear.
    `grads`: optional gradient destinations [gw0, gb0, gw1, gb1, ...] in parameter order (e.g.
    FlatGrads.views); the gradients are OVERWRITTEN there every backward pass, parameters' ``.grad``
    stays untouched."""

    def __init__(self, seq: Sequence[nn.Module], grads: Optional[Sequence[torch.Tensor]] = None):
        super().__init__()
        mods = list(seq)
        layers: List[nn.Module] = []
        gi = 0
        i = 0
        while i < len(mods):
            lin = mods[i]
            assert isinstance(lin, nn.Linear), "expected Linear (+ activation) pairs"
            act = mods[i + 1] if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear) else None
            gw = grads[gi] if grads is not None else None
            gb = grads[gi + 1] if grads is not None else None
            gi += 2
            layers.append(FusedDense(lin, isinstance(act, nn.ReLU), gw, gb))
            if act is not None and not isinstance(act, nn.ReLU):
                layers.append(act)
            i += 1 if act is None else 2
        self.layers = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.layers(x)

    def bind_param_grads(self) -> None:
        """Point every parameter's ``.grad`` at its gradient buffer (call after backward when an
        optimiser reads ``.grad``, as ``custom_update_`` does)."""
        for m in self.layers:
            if isinstance(m, FusedDense):
                m.linear.weight.grad = m.grad_weight
                m.linear.bias.grad = m.grad_bias

    def grad_buffers(self) -> List[torch.Tensor]:
        out = []
        for m in self.layers:
            if isinstance(m, FusedDense):
                out += [m.grad_weight, m.grad_bias]
        return out


__all__ = ["FusedDense", "FusedMLP"]
