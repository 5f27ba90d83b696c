"""Host side of the drop-in boundary: the two functions the reference's pybind11 module
``MultiScaleDeformableAttention`` exports (ops/src/vision.cpp:18-21), re-implemented as
thin Python over the C ABI (include/msda_b200.h).

Argument meaning, order, return types and error behaviour follow the reference's
dispatch + host wrappers (ops/src/ms_deform_attn.h:25-66,
ops/src/cuda/ms_deform_attn_cuda.cu:25-85 and :88-158):

* a non-CUDA ``value`` raises ``RuntimeError("Not implemented on the CPU")``
  (ms_deform_attn.h:43,65) -- there is no CPU path here either;
* every tensor must be contiguous and on a CUDA device (cu:33-43, :98-110);
* ``im2col_step_ = min(batch, im2col_step)`` must divide ``batch`` (cu:55-57, :122-124).
  It is validated for error parity and otherwise ignored: one launch covers the batch;
* dtype float32 or float64 (AT_DISPATCH_FLOATING_TYPES, cu:69,139); the shape tensors
  are int64 on the device (``.data<int64_t>()``, cu:72-73);
* outputs are allocated here with torch's caching allocator on the current stream
  (the reference allocates with ``value.options()``, cu:59, :126-128); the C side never
  allocates, frees or synchronises.
"""
from __future__ import annotations

import torch

from . import _lib

# Optional per-call CUDA-event timing of the two C-ABI launches (bench tools only; off by
# default, no effect on results).  When enabled, every call appends (kind, start, end).
_TIMING = None


def enable_timing(on: bool = True):
    """Start (or stop) collecting CUDA events around each kernel launch; returns the list."""
    global _TIMING
    _TIMING = [] if on else None
    return _TIMING


class _timed:
    def __init__(self, kind):
        self.kind = kind

    def __enter__(self):
        if _TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _TIMING is not None:
            self.e1.record()
            _TIMING.append((self.kind, self.e0, self.e1))


_FWD = {torch.float32: _lib.lib.msda_b200_forward_f32, torch.float64: _lib.lib.msda_b200_forward_f64}
_BWD = {torch.float32: _lib.lib.msda_b200_backward_f32, torch.float64: _lib.lib.msda_b200_backward_f64}


def _check_common(named, im2col_step, opname):
    value = named[0][1]
    if not isinstance(value, torch.Tensor) or not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")                     # ms_deform_attn.h:43,65
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")        # cu:33-37
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")              # cu:39-43
        if t.device != value.device:
            raise RuntimeError(f"{name} is on {t.device}, value is on {value.device}")
    if value.dtype not in _FWD:
        raise RuntimeError(f'"{opname}" not implemented for \'{value.dtype}\'')   # cu:69,139
    _, shapes, lsi, loc, w = (t for _, t in named[:5])
    for name, t in named[1:3]:
        if t.dtype != torch.int64:
            raise RuntimeError(f"expected scalar type Long but found {t.dtype} for {name}")
    for name, t in named[3:]:
        if t.dtype != value.dtype:
            raise RuntimeError(f"expected scalar type {value.dtype} but found {t.dtype} for {name}")
    if value.dim() != 4 or loc.dim() != 6 or shapes.dim() != 2 or shapes.size(1) != 2:
        raise RuntimeError("value must be [N,S,M,D], sampling_loc [N,Lq,M,L,P,2], spatial_shapes [L,2]")
    N, S, M, D = value.shape
    L = shapes.size(0)
    Lq, P = loc.size(1), loc.size(4)
    if tuple(loc.shape) != (N, Lq, M, L, P, 2) or w.numel() != N * Lq * M * L * P or lsi.numel() != L:
        raise RuntimeError(
            f"inconsistent shapes: value {tuple(value.shape)}, spatial_shapes {tuple(shapes.shape)}, "
            f"level_start_index {tuple(lsi.shape)}, sampling_loc {tuple(loc.shape)}, "
            f"attn_weight {tuple(w.shape)}")
    step = min(N, int(im2col_step))
    if N == 0:
        return N, S, M, D, L, Lq, P
    if step <= 0 or N % step != 0:
        raise RuntimeError(f"batch({N}) must divide im2col_step({step})")     # cu:57, :124
    return N, S, M, D, L, Lq, P


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           im2col_step):
    """-> output [N, Lq, M*D]  (vision.cpp:19; ms_deform_attn_cuda.cu:25-85)."""
    named = [("value", value), ("spatial_shapes", spatial_shapes),
             ("level_start_index", level_start_index), ("sampling_loc", sampling_loc),
             ("attn_weight", attn_weight)]
    N, S, M, D, L, Lq, P = _check_common(named, im2col_step, "ms_deform_attn_forward_cuda")
    if N * Lq * M * D == 0 or S == 0 or L * P == 0:
        # nothing to sample: the reference returns its zero-initialised output (cu:59)
        return torch.zeros((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        output = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        with _timed("forward"):
            rc = _FWD[value.dtype](
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P,
                output.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_forward")
    return output


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                            grad_output, im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight]  (vision.cpp:20; cu:88-158)."""
    named = [("value", value), ("spatial_shapes", spatial_shapes),
             ("level_start_index", level_start_index), ("sampling_loc", sampling_loc),
             ("attn_weight", attn_weight), ("grad_output", grad_output)]
    N, S, M, D, L, Lq, P = _check_common(named, im2col_step, "ms_deform_attn_backward_cuda")
    if grad_output.numel() != N * Lq * M * D:
        raise RuntimeError(f"grad_output has {grad_output.numel()} elements, expected {N * Lq * M * D}")
    if N * Lq * M * D == 0 or S == 0 or L * P == 0:
        # nothing to sample: all three gradients stay at the reference's zero fill (cu:126-128)
        return [torch.zeros_like(value), torch.zeros_like(sampling_loc), torch.zeros_like(attn_weight)]
    with torch.cuda.device(value.device):
        grad_value = torch.zeros_like(value)            # accumulated by reductions (cu:126)
        grad_loc = torch.empty_like(sampling_loc)       # fully overwritten (cf. cu:127)
        grad_w = torch.empty_like(attn_weight)          # fully overwritten (cf. cu:128)
        with _timed("backward"):
            rc = _BWD[value.dtype](
                grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(),
                level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                N, S, M, D, L, Lq, P, grad_value.data_ptr(), grad_loc.data_ptr(), grad_w.data_ptr(),
                torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_backward")
    return [grad_value, grad_loc, grad_w]


def debug_indices(value_shape, spatial_shapes, level_start_index, sampling_loc):
    """Integer known-answer hook (msda_b200_debug_indices_f32): (idx int32 [...,4], off int64 [...,4])."""
    N, S, M, D = (int(x) for x in value_shape)
    if not sampling_loc.is_cuda or sampling_loc.dtype != torch.float32 or not sampling_loc.is_contiguous():
        raise RuntimeError("sampling_loc must be a contiguous float32 CUDA tensor")
    L = spatial_shapes.size(0)
    Lq, P = sampling_loc.size(1), sampling_loc.size(4)
    with torch.cuda.device(sampling_loc.device):
        idx = torch.empty(tuple(sampling_loc.shape[:-1]) + (4,), dtype=torch.int32, device=sampling_loc.device)
        off = torch.empty(tuple(sampling_loc.shape[:-1]) + (4,), dtype=torch.int64, device=sampling_loc.device)
        rc = _lib.lib.msda_b200_debug_indices_f32(
            spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            N, S, M, D, L, Lq, P, idx.data_ptr(), off.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "debug_indices")
    return idx, off


# ----------------------------------------------------------------------------------------------
# Fused producers: softmax over the L*P logits and `ref + offset / (W, H)` inside
# the kernels (ops/modules/ms_deform_attn.py:105-112).  No counterpart in the reference's extension.
# ----------------------------------------------------------------------------------------------
def fused_supported(value, reference_points, sampling_offsets, attn_logits):
    """The fused kernels cover the pixel-decoder shape only: fp32, 32 channels per head,
    L*P in {4, 8, 12, 16}, 2-d reference points shared by the batch or per image."""
    if not (isinstance(value, torch.Tensor) and value.is_cuda and value.dtype == torch.float32):
        return False
    if value.dim() != 4 or value.size(3) != 32 or sampling_offsets.dim() != 6:
        return False
    L, P = sampling_offsets.size(3), sampling_offsets.size(4)
    if L * P not in (4, 8, 12, 16) or value.size(1) * value.size(2) * 8 >= 0x7fffffff:
        return False
    return (reference_points.dim() == 4 and reference_points.size(-1) == 2
            and reference_points.size(0) in (1, value.size(0)))


def _check_fused(value, shapes, lsi, ref, off, logits, extra=()):
    named = [("value", value), ("spatial_shapes", shapes), ("level_start_index", lsi),
             ("reference_points", ref), ("sampling_offsets", off), ("attn_logits", logits)] + list(extra)
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda or t.device != value.device:
            raise RuntimeError(f"{name} must be a CUDA tensor on {value.device}")
    if not fused_supported(value, ref, off, logits):
        raise RuntimeError("fused MSDA covers fp32, 32 channels per head, L*P in {4,8,12,16}, 2-d "
                           "reference points only; compose the unfused op for other shapes")
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = off.shape
    if tuple(off.shape) != (N, Lq, M, L, P, 2) or logits.numel() != N * Lq * M * L * P \
            or tuple(ref.shape[1:]) != (Lq, L, 2) or shapes.size(0) != L:
        raise RuntimeError("inconsistent shapes for fused MSDA")
    for t in (ref, off, logits) + tuple(t for _, t in extra):
        if t.dtype != torch.float32:
            raise RuntimeError("fused MSDA is float32 only")
    ref_stride = 0 if ref.size(0) == 1 and N > 1 else Lq * L * 2
    return N, S, M, D, L, Lq, P, ref_stride


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, reference_points,
                                 sampling_offsets, attn_logits):
    """output [N, Lq, M*D] from raw offsets / logits (msda_b200_fused_forward_f32)."""
    N, S, M, D, L, Lq, P, rs = _check_fused(value, spatial_shapes, level_start_index, reference_points,
                                            sampling_offsets, attn_logits)
    with torch.cuda.device(value.device):
        output = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        with _timed("forward"):
            rc = _lib.lib.msda_b200_fused_forward_f32(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                reference_points.data_ptr(), rs, sampling_offsets.data_ptr(), attn_logits.data_ptr(),
                N, S, M, D, L, Lq, P, output.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_fused_forward")
    return output


def ms_deform_attn_fused_forward_packed(value, spatial_shapes, level_start_index, reference_points,
                                        projections, n_levels, n_points, query_table=None):
    """The fused forward reading the raw offsets and logits in place from ONE projection output
    ``projections`` [N, Lq, 3*M*L*P] = [sampling_offsets (2*M*L*P) | attn_logits (M*L*P)] per query -- the
    stacked ``sampling_offsets`` / ``attention_weights`` Linear (ops/modules/ms_deform_attn.py:105-106) run
    as one GEMM (msda_b200_fused_forward_strided_f32).  ``query_table`` [Lq, 3*M*L*P] (optional) is added to
    the rows of every image inside the kernel: the part of the projection that depends on the query's
    position only (``pos @ W^T + b`` when the projection input is ``src + pos``).  Inference only."""
    N, S, M, D = value.shape
    L, P = int(n_levels), int(n_points)
    width = 3 * M * L * P
    if projections.dim() != 3 or projections.size(0) != N or projections.size(2) != width:
        raise RuntimeError(f"projections must be [N, Lq, {width}]")
    Lq = projections.size(1)
    off = projections[..., :2 * M * L * P]
    logits = projections[..., 2 * M * L * P:]
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("reference_points", reference_points), ("projections", projections)]
    if query_table is not None:
        named.append(("query_table", query_table))
        if tuple(query_table.shape) != (Lq, width) or query_table.dtype != torch.float32:
            raise RuntimeError(f"query_table must be a float32 [Lq, {width}] tensor")
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda or t.device != value.device:
            raise RuntimeError(f"{name} must be a CUDA tensor on {value.device}")
    if not fused_supported(value, reference_points, off.view(N, Lq, M, L, P, 2), logits) \
            or tuple(reference_points.shape[1:]) != (Lq, L, 2) or spatial_shapes.size(0) != L \
            or projections.dtype != torch.float32 or reference_points.dtype != torch.float32:
        raise RuntimeError("fused MSDA covers fp32, 32 channels per head, L*P in {4,8,12,16}, 2-d "
                           "reference points only; compose the unfused op for other shapes")
    rs = 0 if reference_points.size(0) == 1 and N > 1 else Lq * L * 2
    with torch.cuda.device(value.device):
        output = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        with _timed("forward"):
            rc = _lib.lib.msda_b200_fused_forward_strided_f32(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                reference_points.data_ptr(), rs, off.data_ptr(), width, logits.data_ptr(), width,
                query_table.data_ptr() if query_table is not None else None,
                query_table[:, 2 * M * L * P:].data_ptr() if query_table is not None else None,
                N, S, M, D, L, Lq, P, output.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_fused_forward_packed")
    return output


def ms_deform_attn_fused_backward(value, spatial_shapes, level_start_index, reference_points,
                                  sampling_offsets, attn_logits, grad_output):
    """[grad_value, grad_sampling_offsets, grad_attn_logits] (msda_b200_fused_backward_f32)."""
    N, S, M, D, L, Lq, P, rs = _check_fused(value, spatial_shapes, level_start_index, reference_points,
                                            sampling_offsets, attn_logits, [("grad_output", grad_output)])
    with torch.cuda.device(value.device):
        grad_value = torch.zeros_like(value)
        grad_off = torch.empty_like(sampling_offsets)
        grad_logits = torch.empty_like(attn_logits)
        with _timed("backward"):
            rc = _lib.lib.msda_b200_fused_backward_f32(
                grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(),
                level_start_index.data_ptr(), reference_points.data_ptr(), rs,
                sampling_offsets.data_ptr(), attn_logits.data_ptr(), N, S, M, D, L, Lq, P,
                grad_value.data_ptr(), grad_off.data_ptr(), grad_logits.data_ptr(),
                torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_fused_backward")
    return [grad_value, grad_off, grad_logits]


# ----------------------------------------------------------------------------------------------
# fp32 Linear on the tensor cores: the nn.Linear layers around the op
# (ops/modules/ms_deform_attn.py:62-65, msdeformattn.py:126-130) as an error-compensated 3xTF32 GEMM.
# Inference only (no backward); torch.nn.functional.linear semantics.
# ----------------------------------------------------------------------------------------------
def linear_tf32x3_supported(x, weight) -> bool:
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32
            and weight.dtype == torch.float32 and weight.dim() == 2 and x.size(-1) == weight.size(1)
            and weight.is_contiguous() and weight.device == x.device
            and weight.size(1) % 32 == 0 and weight.size(0) % 4 == 0 and x.numel() > 0)


def split_weight(weight):
    """The split form of a Linear weight that ``linear_tf32x3(..., presplit=...)`` consumes: ``tf32(W)`` followed by
    ``tf32(W - tf32(W))`` as one flat fp32 tensor of 2 * out * in elements (msda_b200_split_weight_f32).  For weights
    that do not change between calls (inference): one launch per Linear instead of two."""
    if not (isinstance(weight, torch.Tensor) and weight.is_cuda and weight.dtype == torch.float32 and weight.dim() == 2
            and weight.is_contiguous() and weight.numel() % 4 == 0 and weight.numel() > 0):
        raise RuntimeError("split_weight needs a contiguous 2-d fp32 CUDA weight")
    out = torch.empty(2 * weight.numel(), dtype=torch.float32, device=weight.device)
    with torch.cuda.device(weight.device):
        rc = _lib.lib.msda_b200_split_weight_f32(weight.data_ptr(), out.data_ptr(), weight.size(0), weight.size(1),
                                                 torch.cuda.current_stream(weight.device).cuda_stream)
    _lib.check(rc, "split_weight")
    return out


def linear_tf32x3(x, weight, bias=None, relu=False, split_weight_in_kernel=False, presplit=None):
    """``F.linear(x, weight, bias)`` (then ``relu`` if set) for fp32 CUDA tensors; ``x`` [..., in],
    ``presplit``: ``split_weight(weight)`` computed earlier -- the GEMM then runs alone (same bits);
    ``weight`` [out, in] with in % 32 == 0 and out % 4 == 0.  Non-finite inputs give non-finite outputs in
    the affected rows, but an infinite input may come out as NaN where F.linear returns +-inf (the
    error-compensated split forms inf - inf and inf * w_lo of either sign).  ``split_weight_in_kernel``: no pre-pass over
    (and no workspace for) the weight -- for a one-shot "weight" such as transposed activations."""
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for name, t in (("x", x), ("weight", weight)) + ((("bias", bias),) if bias is not None else ()):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda or t.device != x.device:
            raise RuntimeError(f"{name} must be a CUDA tensor on {x.device}")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name} must be float32")
    if not linear_tf32x3_supported(x, weight) or (bias is not None and bias.shape != (weight.size(0),)):
        raise RuntimeError("linear_tf32x3 needs x [..., in] and weight [out, in] with in % 32 == 0, out % 4 == 0")
    out_f, in_f = weight.shape
    rows = x.numel() // in_f
    y = torch.empty(x.shape[:-1] + (out_f,), dtype=x.dtype, device=x.device)
    if presplit is not None:
        if not (presplit.is_cuda and presplit.device == x.device and presplit.dtype == torch.float32
                and presplit.is_contiguous() and presplit.numel() == 2 * out_f * in_f):
            raise RuntimeError("presplit must be split_weight(weight) on x's device")
        with torch.cuda.device(x.device):
            rc = _lib.lib.msda_b200_linear_presplit_f32(x.data_ptr(), presplit.data_ptr(),
                                                        bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                                        rows, out_f, in_f, 1 if relu else 0,
                                                        torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(rc, "linear_tf32x3")
        return y
    workspace = None if split_weight_in_kernel else \
        torch.empty(2 * out_f * in_f, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib.msda_b200_linear_f32(x.data_ptr(), weight.data_ptr(),
                                           bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                           rows, out_f, in_f, 1 if relu else 0,
                                           workspace.data_ptr() if workspace is not None else None,
                                           torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "linear_tf32x3")
    return y


def add_layernorm_supported(x, residual, weight, bias=None, need_bias=False) -> bool:
    """Whether the fused residual + LayerNorm kernels cover these tensors; callers pass the norm's bias
    with need_bias=True so that an nn.LayerNorm(bias=False), or affine parameters that are not
    contiguous fp32 tensors on x's device, take torch's own path."""
    def affine_ok(t):
        return (isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.is_contiguous()
                and t.device == x.device and t.numel() == x.size(-1))
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
            and x.size(-1) in (128, 256, 384, 512) and x.numel() > 0 and affine_ok(weight)
            and (affine_ok(bias) if (need_bias or bias is not None) else True)
            and (residual is None or (residual.shape == x.shape and residual.is_contiguous()
                                      and residual.dtype == torch.float32 and residual.device == x.device)))


def add_layernorm(x, residual, weight, bias, eps=1e-5):
    """``F.layer_norm(x + residual, (C,), weight, bias, eps)`` in one pass (fp32 CUDA, C in {128,...,512});
    ``residual`` may be None.  Inference only (no backward)."""
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if not add_layernorm_supported(x, residual, weight) or bias is None or bias.shape != weight.shape:
        raise RuntimeError("add_layernorm needs contiguous fp32 CUDA tensors with a last dimension of "
                           "128, 256, 384 or 512 and affine parameters")
    y = torch.empty_like(x)
    cols = x.size(-1)
    with torch.cuda.device(x.device):
        rc = _lib.lib.msda_b200_add_layernorm_f32(
            x.data_ptr(), residual.data_ptr() if residual is not None else None, weight.data_ptr(),
            bias.data_ptr(), y.data_ptr(), x.numel() // cols, cols, float(eps),
            torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "add_layernorm")
    return y


def group_norm_supported(x, norm, up=None) -> bool:
    """Whether the fused GroupNorm kernel covers `norm(x)` (x: NCHW fp32 CUDA, contiguous)."""
    ok = (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
          and x.is_contiguous() and x.numel() > 0 and isinstance(norm, torch.nn.GroupNorm) and norm.affine
          and norm.weight is not None and norm.bias is not None and norm.weight.device == x.device
          and norm.weight.dtype == torch.float32 and norm.weight.is_contiguous() and norm.bias.is_contiguous()
          and x.size(1) == norm.num_channels and (x.size(2) * x.size(3)) % 4 == 0
          and x.data_ptr() % 16 == 0)
    if ok and up is not None:
        ok = (up.is_cuda and up.dtype == torch.float32 and up.dim() == 4 and up.is_contiguous()
              and up.device == x.device and up.shape[:2] == x.shape[:2] and x.size(3) % 4 == 0)
    return ok


def group_norm(x, norm, relu=False, up=None, channel_bias=None):
    """``norm(x)`` (``norm(x + channel_bias[None, :, None, None])`` when the producing convolution ran without
    its bias) for an ``nn.GroupNorm`` on an NCHW fp32 CUDA map, optionally followed by ReLU and by
    ``+ F.interpolate(up, size=x.shape[-2:], mode="bilinear", align_corners=False)`` -- the epilogues of the
    pixel decoder's input projections and FPN level (msdeformattn.py:233-248, :369-379) -- in two passes
    over x (statistics, apply) instead of torch's four to six kernels.  Inference only (no backward)."""
    if not group_norm_supported(x, norm, up):
        raise RuntimeError("group_norm needs a contiguous NCHW fp32 CUDA map with H*W (and, with `up`, W) a multiple of 4, "
                           "an affine nn.GroupNorm on the same device and, if given, a contiguous `up` map "
                           "with the same batch and channels")
    N, C, H, W = x.shape
    if channel_bias is not None and not (channel_bias.is_cuda and channel_bias.device == x.device
                                         and channel_bias.dtype == torch.float32 and channel_bias.is_contiguous()
                                         and channel_bias.shape == (C,)):
        raise RuntimeError("channel_bias must be a contiguous fp32 [C] tensor on x's device")
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ws = torch.empty(int(_lib.lib.msda_b200_group_norm_workspace_bytes(N, norm.num_groups)), dtype=torch.uint8,
                         device=x.device)
        rc = _lib.lib.msda_b200_group_norm_nchw_f32(
            x.data_ptr(), channel_bias.data_ptr() if channel_bias is not None else None,
            norm.weight.data_ptr(), norm.bias.data_ptr(), y.data_ptr(), N, C, H, W, norm.num_groups,
            float(norm.eps), int(bool(relu)), up.data_ptr() if up is not None else None,
            up.size(2) if up is not None else 0, up.size(3) if up is not None else 0, ws.data_ptr(),
            torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "group_norm")
    return y


def group_norm_rows_supported(x, norm, out_rows) -> bool:
    """Whether `group_norm_rows` covers these tensors: the map as for `group_norm`, channels a multiple of 32, and
    `out_rows` an fp32 [N, H*W, C] view with unit channel stride (a slice of the concatenated [N, S, C] tensor)."""
    return (group_norm_supported(x, norm) and x.size(1) % 32 == 0 and isinstance(out_rows, torch.Tensor)
            and out_rows.is_cuda and out_rows.device == x.device and out_rows.dtype == torch.float32
            and out_rows.dim() == 3 and tuple(out_rows.shape) == (x.size(0), x.size(2) * x.size(3), x.size(1))
            and out_rows.stride(2) == 1 and out_rows.stride(1) >= x.size(1) and out_rows.stride(1) % 4 == 0
            and out_rows.stride(0) % 4 == 0 and out_rows.data_ptr() % 16 == 0)


def group_norm_rows(x, norm, out_rows, channel_bias=None, relu=False):
    """``out_rows[n, p, c] = norm(x + channel_bias)[n, c, p]`` (p = flattened pixel): GroupNorm of an NCHW map written
    straight into the ``[N, pixels, C]`` layout of the deformable encoder (``src.flatten(2).transpose(1, 2)``),
    e.g. into the level's rows of the concatenated tensor.  Inference only (no backward).  Returns out_rows."""
    if not group_norm_rows_supported(x, norm, out_rows):
        raise RuntimeError("group_norm_rows needs what group_norm needs, channels % 32 == 0 and an fp32 [N, H*W, C] "
                           "destination view with unit channel stride on the same device")
    N, C, H, W = x.shape
    if channel_bias is not None and not (channel_bias.is_cuda and channel_bias.device == x.device
                                         and channel_bias.dtype == torch.float32 and channel_bias.is_contiguous()
                                         and channel_bias.shape == (C,)):
        raise RuntimeError("channel_bias must be a contiguous fp32 [C] tensor on x's device")
    with torch.cuda.device(x.device):
        ws = torch.empty(int(_lib.lib.msda_b200_group_norm_workspace_bytes(N, norm.num_groups)), dtype=torch.uint8,
                         device=x.device)
        rc = _lib.lib.msda_b200_group_norm_nchw_to_rows_f32(
            x.data_ptr(), channel_bias.data_ptr() if channel_bias is not None else None, norm.weight.data_ptr(),
            norm.bias.data_ptr(), out_rows.data_ptr(), out_rows.stride(1), out_rows.stride(0), N, C, H * W,
            norm.num_groups, float(norm.eps), int(bool(relu)), ws.data_ptr(),
            torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "group_norm_rows")
    return out_rows


def channel_bias_supported(x, bias) -> bool:
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
            and x.is_contiguous() and x.numel() > 0 and (x.size(2) * x.size(3)) % 4 == 0 and x.data_ptr() % 16 == 0
            and isinstance(bias, torch.Tensor) and bias.device == x.device and bias.dtype == torch.float32
            and bias.is_contiguous() and bias.shape == (x.size(1),))


def add_channel_bias_(x, bias):
    """``x += bias[None, :, None, None]`` in place on a contiguous NCHW fp32 CUDA map (the bias of a convolution
    that ran without one: torch's own broadcast add runs at less than half of the memory roofline)."""
    if not channel_bias_supported(x, bias):
        raise RuntimeError("add_channel_bias_ needs a contiguous NCHW fp32 CUDA map with H*W % 4 == 0 and a [C] bias")
    with torch.cuda.device(x.device):
        rc = _lib.lib.msda_b200_add_channel_bias_nchw_f32(x.data_ptr(), bias.data_ptr(), x.size(0), x.size(1),
                                                          x.size(2) * x.size(3),
                                                          torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "add_channel_bias_")
    return x


def transpose2d(x, out=None):
    """``x.t().contiguous()`` for a contiguous fp32 CUDA matrix (operand preparation for the weight-gradient GEMM;
    the [pixels, C] -> [C, pixels] turn of the encoder memory); ``out``: a contiguous [cols, rows] destination."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous() and x.numel() > 0):
        raise RuntimeError("transpose2d needs a contiguous 2-d fp32 CUDA tensor")
    if out is not None and not (out.is_cuda and out.device == x.device and out.dtype == torch.float32
                                and out.is_contiguous() and out.numel() == x.numel()):
        raise RuntimeError("transpose2d: out must be a contiguous fp32 tensor of x's size on x's device")
    y = torch.empty(x.size(1), x.size(0), dtype=x.dtype, device=x.device) if out is None else out
    with torch.cuda.device(x.device):
        rc = _lib.lib.msda_b200_transpose_f32(x.data_ptr(), y.data_ptr(), x.size(0), x.size(1),
                                              torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "transpose2d")
    return y


def linear_wgrad_supported(grad_y, x) -> bool:
    return (isinstance(grad_y, torch.Tensor) and grad_y.is_cuda and grad_y.dtype == torch.float32 and x.dtype == torch.float32
            and grad_y.dim() == 2 and x.dim() == 2 and grad_y.size(0) == x.size(0) and grad_y.numel() > 0 and x.numel() > 0
            and grad_y.is_contiguous() and x.is_contiguous() and grad_y.size(1) % 4 == 0 and x.size(1) % 4 == 0
            and x.device == grad_y.device)


def linear_wgrad(grad_y, x, with_bias=True):
    """(grad_weight [out, in], grad_bias [out] or None) of ``F.linear`` from ``grad_y`` [rows, out] and ``x``
    [rows, in] (contiguous fp32 CUDA): grad_y^T @ x and the column sums of grad_y in one kernel."""
    if not grad_y.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if not linear_wgrad_supported(grad_y, x):
        raise RuntimeError("linear_wgrad needs contiguous fp32 CUDA matrices grad_y [rows, out], x [rows, in] with "
                           "out % 4 == 0 and in % 4 == 0")
    gw = torch.empty(grad_y.size(1), x.size(1), dtype=torch.float32, device=x.device)
    gb = torch.empty(grad_y.size(1), dtype=torch.float32, device=x.device) if with_bias else None
    with torch.cuda.device(x.device):
        rc = _lib.lib.msda_b200_linear_wgrad_f32(grad_y.data_ptr(), x.data_ptr(), gw.data_ptr(),
                                                 gb.data_ptr() if gb is not None else None, grad_y.size(0),
                                                 grad_y.size(1), x.size(1),
                                                 torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "linear_wgrad")
    return gw, gb


def add_layernorm_backward(grad_y, x, residual, weight, eps=1e-5):
    """(grad_v, grad_weight, grad_bias) of ``layer_norm(x + residual)``; grad_v is the gradient of both x and
    residual.  Last dimension 128 or 256."""
    if not (add_layernorm_supported(x, residual, weight) and x.size(-1) in (128, 256)
            and grad_y.shape == x.shape and grad_y.is_contiguous() and grad_y.dtype == torch.float32):
        raise RuntimeError("add_layernorm_backward needs contiguous fp32 CUDA tensors with a last dimension of 128 or 256")
    gv = torch.empty_like(x)
    gg, gb = torch.empty_like(weight), torch.empty_like(weight)
    cols = x.size(-1)
    with torch.cuda.device(x.device):
        rc = _lib.lib.msda_b200_add_layernorm_backward_f32(
            grad_y.data_ptr(), x.data_ptr(), residual.data_ptr() if residual is not None else None,
            weight.data_ptr(), gv.data_ptr(), gg.data_ptr(), gb.data_ptr(), x.numel() // cols, cols, float(eps),
            torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "add_layernorm_backward")
    return gv, gg, gb
