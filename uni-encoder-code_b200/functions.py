"""Autograd wrapper with the reference's interface:
``MSDeformAttnFunction.apply(value, spatial_shapes, level_start_index, sampling_locations,
attention_weights, im2col_step)`` (ops/functions/ms_deform_attn_func.py:35-52).

The reference's own ``MSDeformAttnFunction`` also runs unchanged on top of
``dropin/MultiScaleDeformableAttention.py``; this mirror exists so that code on a box
without the reference checkout (the GPU box, bench.py) has the same entry point.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = ops.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                            sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, w = ctx.saved_tensors
        grad_value, grad_loc, grad_w = ops.ms_deform_attn_backward(
            value, shapes, lsi, loc, w, grad_output, ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_w, None


class MSDeformAttnFusedFunction(Function):
    """``apply(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
    attn_logits)`` -- the op with its caller's softmax and location arithmetic
    (ops/modules/ms_deform_attn.py:105-112) inside the kernels.  Gradients flow to value,
    sampling_offsets and attn_logits; reference_points are treated as constants (they are in the
    encoder, msdeformattn.py:152-166)."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                attn_logits):
        output = ops.ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index,
                                                  reference_points, sampling_offsets, attn_logits)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points,
                              sampling_offsets, attn_logits)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, ref, off, logits = ctx.saved_tensors
        grad_value, grad_off, grad_logits = ops.ms_deform_attn_fused_backward(
            value, shapes, lsi, ref, off, logits, grad_output)
        return grad_value, None, None, None, grad_off, grad_logits


class LinearTF32x3Function(Function):
    """``apply(x, weight, bias)`` = ``F.linear`` with all three GEMMs on the tensor cores: forward and
    input gradient (``grad_y @ weight``) through ops.linear_tf32x3, weight and bias gradient
    (``grad_y^T @ x``, column sums of ``grad_y``) through ops.linear_wgrad, which reduces over the rows
    without transposing the operands in memory (0.15 ms against 0.53 + 0.12 ms for torch's fp32 GEMM and
    column sum at 172 032 x 256 x 256).  Needs in_features and out_features divisible by 32.
    ``apply(x, weight, bias, True)`` = ``F.relu(F.linear(...))`` with the ReLU in the GEMM's epilogue (no separate
    pass over the output); its backward masks ``grad_y`` where the saved output is zero, as ``F.relu`` does.
    ``apply(x, weight, bias, True, p)`` = ``F.dropout(F.relu(F.linear(...)), p, training=True)`` (the FFN's hidden
    activation, msdeformattn.py:126-130): the mask is drawn by torch's own dropout kernel (same generator stream
    as ``F.dropout``); a kept element is positive exactly where the ReLU was active, so the backward needs ONE pass
    over the hidden gradient (``threshold_backward`` against the saved output) instead of two, and the 1/(1-p)
    scale moves onto the small tensors (the transposed weight, grad_weight, grad_bias)."""

    @staticmethod
    def supported(x, weight) -> bool:
        return (ops.linear_tf32x3_supported(x, weight) and x.is_contiguous()
                and weight.size(0) % 32 == 0 and weight.size(1) % 32 == 0)

    @staticmethod
    def forward(ctx, x, weight, bias, relu=False, dropout_p=0.0):
        if dropout_p and not (relu and 0.0 < dropout_p < 1.0):
            raise ValueError("dropout_p needs relu=True and 0 < p < 1")
        y = ops.linear_tf32x3(x, weight.contiguous(), bias, relu=relu)
        ctx.has_bias = bias is not None
        ctx.relu = bool(relu)
        ctx.scale = 1.0
        if dropout_p:
            y, _ = torch.native_dropout(y, float(dropout_p), True)       # y * mask / (1 - p)
            ctx.scale = 1.0 / (1.0 - float(dropout_p))
        if relu:
            ctx.save_for_backward(x, weight, y)
        else:
            ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        if ctx.relu:
            x, weight, y = ctx.saved_tensors
            # F.relu's backward; with dropout, y > 0 exactly where the element was kept AND the ReLU was active,
            # and the scale 1/(1-p) is applied to the results below instead of to this (large) tensor
            grad_y = torch.ops.aten.threshold_backward(grad_y, y, 0)
        else:
            x, weight = ctx.saved_tensors
        grad_y = grad_y.contiguous()
        scale = ctx.scale
        grad_x = grad_w = grad_b = None
        if ctx.needs_input_grad[0]:
            wt = weight.t().contiguous() if scale == 1.0 else (weight.t() * scale).contiguous()
            grad_x = ops.linear_tf32x3(grad_y, wt, None)
        g2 = grad_y.reshape(-1, grad_y.size(-1))
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            x2 = x.reshape(-1, x.size(-1))
            if ops.linear_wgrad_supported(g2, x2):
                grad_w, grad_b = ops.linear_wgrad(g2, x2, with_bias=want_b)
            else:
                grad_w = g2.t() @ x2
        if want_b and grad_b is None:
            grad_b = g2.sum(0)
        if scale != 1.0:
            grad_w = grad_w * scale if grad_w is not None else None
            grad_b = grad_b * scale if grad_b is not None else None
        return grad_x, grad_w, grad_b, None, None


class AddLayerNormFunction(Function):
    """``apply(x, residual, weight, bias, eps)`` = ``F.layer_norm(x + residual, (C,), weight, bias, eps)`` with one
    fused kernel each way (ops.add_layernorm / add_layernorm_backward); C in {128, 256}."""

    @staticmethod
    def supported(x, residual, weight) -> bool:
        return ops.add_layernorm_supported(x, residual, weight) and x.size(-1) in (128, 256) and residual is not None

    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps):
        ctx.save_for_backward(x, residual, weight)
        ctx.eps = eps
        return ops.add_layernorm(x, residual, weight, bias, eps)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        x, residual, weight = ctx.saved_tensors
        gv, gg, gb = ops.add_layernorm_backward(grad_y.contiguous(), x, residual, weight, ctx.eps)
        return gv, gv, gg, gb, None
