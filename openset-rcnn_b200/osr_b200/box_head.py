"""Box head: drop-in for detectron2's ``FastRCNNConvFCHead`` in its shipped form - flatten,
``fc1``, ReLU, ``fc2``, ReLU (``configs/Base-RCNN-FPN.yaml``: ``ROI_BOX_HEAD.NUM_FC 2``, ``FC_DIM 1024``; built at
``osrcnn_roi_heads.py:119-121``, called at ``:308``).  Same parameter names (``fc1.weight`` ... ``fc2.bias``) and the same
``c2_xavier_fill`` initialisation, so detectron2 checkpoints load.

Forward runs on the 5th-generation tensor cores (``osr_linear_bf16_fwd``: tcgen05 + TMEM + TMA, bf16 operands, fp32
accumulate, bias + ReLU fused).  Feed it the bf16 pooled tensor from ``ROIPooler.forward(..., out_dtype=torch.bfloat16)``
and the ROIAlign output makes one trip through memory in bf16 instead of the reference's fp32 write + read (411 MB each
at cfg 2).  The backward GEMMs (grad_x = g W, grad_W = g^T x) are plain library GEMMs in bf16, like any ``nn.Linear``
backward under bf16 autocast - they are not part of the hot path this package replaces.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import _lib


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy through ``osr_cast_bf16`` (numel % 4 == 0)."""
    _lib.require_cuda(x)
    xc = x.detach().contiguous().float()
    out = torch.empty(xc.shape, dtype=torch.bfloat16, device=xc.device)
    if xc.numel():
        _lib.check(_lib.lib().osr_cast_bf16(xc.data_ptr(), out.data_ptr(), xc.numel(), _lib.stream_ptr(xc.device)), "osr_cast_bf16")
    return out


def linear_bf16(a: torch.Tensor, w_bf16: torch.Tensor, bias: Optional[torch.Tensor], relu: bool, out_dtype=torch.bfloat16):
    """``act(a @ w^T + bias)`` on tcgen05: ``a`` (R, K) bf16, ``w_bf16`` (N, K) bf16, fp32 accumulate; K % 64 == 0, N % 256 == 0."""
    _lib.require_cuda(a, w_bf16)
    assert a.dtype == torch.bfloat16 and w_bf16.dtype == torch.bfloat16 and a.dim() == 2 and w_bf16.dim() == 2
    a = a.contiguous()
    w_bf16 = w_bf16.contiguous()
    R, K = a.shape
    N = w_bf16.shape[0]
    assert w_bf16.shape[1] == K
    assert out_dtype in (torch.bfloat16, torch.float32)
    out = torch.empty((R, N), dtype=out_dtype, device=a.device)
    b = None if bias is None else bias.detach().contiguous().float()
    rc = _lib.lib().osr_linear_bf16_fwd(a.data_ptr(), w_bf16.data_ptr(), _lib.ptr(b), R, K, N, 1 if relu else 0, out.data_ptr(),
                                        1 if out_dtype == torch.bfloat16 else 0, _lib.stream_ptr(a.device))
    _lib.check(rc, "osr_linear_bf16_fwd")
    return out


class _LinearReluTc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, w_bf16, out_dtype):
        y = linear_bf16(x, w_bf16, bias, True, out_dtype)
        ctx.save_for_backward(x, w_bf16, y)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, w_bf16, y = ctx.saved_tensors
        g = (g * (y > 0)).to(torch.bfloat16)          # ReLU mask, then the two library GEMMs
        gx = (g @ w_bf16) if ctx.needs_input_grad[0] else None
        gw = (g.t() @ x).float() if ctx.needs_input_grad[1] else None
        gb = g.float().sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None, None


class FastRCNNConvFCHead(nn.Module):
    """``FastRCNNConvFCHead(input_shape, conv_dims=[], fc_dims=[1024, 1024])`` (detectron2 ``box_head.py``), fc layers only."""

    def __init__(self, input_shape: Sequence[int], *, conv_dims: Sequence[int] = (), fc_dims: Sequence[int] = (1024, 1024),
                 device="cuda"):
        super().__init__()
        if len(conv_dims):
            raise NotImplementedError("the shipped Openset R-CNN configs have no conv layers in the box head (NUM_CONV 0)")
        assert len(fc_dims) > 0
        self._output_size = int(np.prod(input_shape))
        self.fcs: List[nn.Linear] = []
        for k, fc_dim in enumerate(fc_dims):
            fc = nn.Linear(self._output_size, fc_dim, device=device)
            nn.init.kaiming_uniform_(fc.weight, a=1)   # fvcore c2_xavier_fill
            nn.init.constant_(fc.bias, 0)
            self.add_module("fc{}".format(k + 1), fc)
            self.fcs.append(fc)
            self._output_size = fc_dim
        self._w_bf16 = [None] * len(self.fcs)
        self._w_version = [-1] * len(self.fcs)

    @property
    def output_size(self) -> int:
        return self._output_size

    def _weight_bf16(self, k: int) -> torch.Tensor:
        w = self.fcs[k].weight
        if self._w_bf16[k] is None or self._w_version[k] != w._version or self._w_bf16[k].device != w.device:
            self._w_bf16[k] = cast_bf16(w)            # once per optimizer step (the version counter moves on in-place updates)
            self._w_version[k] = w._version
        return self._w_bf16[k]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``x``: pooled features (M, C, P, P) or (M, C*P*P), bf16 (from ``ROIPooler(..., out_dtype=torch.bfloat16)``) or fp32
        (cast here: one extra pass).  Returns (M, fc_dim) fp32 like the reference's head."""
        x = torch.flatten(x, 1)
        if x.dtype != torch.bfloat16:
            x = cast_bf16(x) if not x.requires_grad else x.to(torch.bfloat16)
        for k, fc in enumerate(self.fcs):
            last = k == len(self.fcs) - 1
            x = _LinearReluTc.apply(x, fc.weight, fc.bias, self._weight_bf16(k), torch.float32 if last else torch.bfloat16)
        return x
