"""CPU oracle for multi-scale deformable attention (MSDA).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does;
it fails loudly when its CUDA library is missing instead of falling back to this.

Two independent restatements of the reference live here:

* ``forward`` / ``backward`` / ``indices`` -- ctypes bindings of ``msda_oracle.c``, the
  plain-C restatement of the reference CUDA arithmetic
  (``ops/src/cuda/ms_deform_im2col_cuda.cuh:38-89, 92-164, 242-304``).  Used for the
  bit-exact integer known-answer tests and, in fp64, as the float golden.
* ``core_grid_sample`` -- a torch restatement of the reference's own debug path
  ``ms_deform_attn_core_pytorch`` (``ops/functions/ms_deform_attn_func.py:55-75``):
  per level, ``grid_sample(bilinear, zeros, align_corners=False)`` on ``2*loc-1`` and a
  weighted sum over the L*P samples.  It is differentiable, so autograd through it
  gives the three reference gradients.  It is also what the CPU baseline times
  (``cpu_baseline.kind == "port"``): the reference's CPU path *is* this function.

Parity pin: both are checked in ``tests/test_oracle.py`` against
``tests/golden/*.npz``, produced by importing the reference's
``ms_deform_attn_core_pytorch`` from ``/root/reference`` in the build container
(``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile msda_oracle.c with the system gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _np(t, dtype):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(value, shapes, loc):
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    assert loc.shape == (N, Lq, M, L, P, 2), loc.shape
    return [ctypes.c_int(int(x)) for x in (N, S, M, D, L, Lq, P)]


def _suffix(dtype, geometry=None):
    base = {np.float32: "f32", np.float64: "f64"}[np.dtype(dtype).type]
    if geometry is not None and np.dtype(geometry) == np.float32 and base == "f64":
        return "f64g32"     # fp32 sampling-point geometry, fp64 accumulation
    return base


def indices(shapes, level_start, loc, value_shape, dtype=np.float32):
    """Integer decomposition of every sampling point.

    Returns ``idx [N,Lq,M,L,P,4] int32`` = (valid, h_low, w_low, corner mask) and
    ``off [N,Lq,M,L,P,4] int64`` = flat element offset of each contributing corner
    (channel 0) in the whole value tensor, -1 where the corner does not contribute.
    """
    lib = _load()
    shapes = _np(shapes, np.int64)
    level_start = _np(level_start, np.int64)
    loc = _np(loc, dtype)
    N, S, M, D = value_shape
    L, Lq, P = shapes.shape[0], loc.shape[1], loc.shape[4]
    idx = np.empty(loc.shape[:-1] + (4,), np.int32)
    off = np.empty(loc.shape[:-1] + (4,), np.int64)
    getattr(lib, "msda_oracle_indices_" + _suffix(dtype))(
        _ptr(shapes), _ptr(level_start), _ptr(loc),
        *[ctypes.c_int(int(x)) for x in (N, S, M, D, L, Lq, P)], _ptr(idx), _ptr(off))
    return idx, off


def forward(value, shapes, level_start, loc, weight, dtype=np.float64, geometry=None):
    """out [N, Lq, M*D]; arithmetic in ``dtype`` (fp64 on fp32 inputs is the golden).

    ``geometry=np.float32`` with ``dtype=np.float64`` computes the sampling-point geometry
    (pixel coordinate, floor, fractions) in fp32 exactly like the reference kernel
    (cuh:290-291, 43-50) and everything after it in fp64."""
    lib = _load()
    value, loc, weight = _np(value, dtype), _np(loc, dtype), _np(weight, dtype)
    shapes, level_start = _np(shapes, np.int64), _np(level_start, np.int64)
    dims = _dims(value, shapes, loc)
    N, S, M, D = value.shape
    out = np.empty((N, loc.shape[1], M * D), dtype)
    getattr(lib, "msda_oracle_forward_" + _suffix(dtype, geometry))(
        _ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(weight), *dims, _ptr(out))
    return out


def backward(grad_out, value, shapes, level_start, loc, weight, dtype=np.float64, geometry=None):
    """(grad_value, grad_sampling_loc, grad_attn_weight), shapes as their primals."""
    lib = _load()
    value, loc, weight = _np(value, dtype), _np(loc, dtype), _np(weight, dtype)
    grad_out = _np(grad_out, dtype)
    shapes, level_start = _np(shapes, np.int64), _np(level_start, np.int64)
    dims = _dims(value, shapes, loc)
    gv = np.zeros_like(value)
    gl = np.empty_like(loc)
    gw = np.empty_like(weight)
    getattr(lib, "msda_oracle_backward_" + _suffix(dtype, geometry))(
        _ptr(grad_out), _ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(weight),
        *dims, _ptr(gv), _ptr(gl), _ptr(gw))
    return gv, gl, gw


def core_grid_sample(value, shapes, loc, weight):
    """Torch restatement of the reference CPU path (func.py:55-75); differentiable.

    value [N,S,M,D], shapes iterable of (H,W), loc [N,Lq,M,L,P,2], weight [N,Lq,M,L,P]
    -> [N, Lq, M*D].
    """
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = [(int(h), int(w)) for h, w in (shapes.tolist() if hasattr(shapes, "tolist") else shapes)]
    grid = loc * 2 - 1                                   # func.py:61
    per_head_w = weight.permute(0, 2, 1, 3, 4).reshape(N * M, Lq, L * P)
    acc = None
    start = 0
    for lvl, (H, W) in enumerate(hw):
        img = value[:, start:start + H * W]              # func.py:60 (split by level)
        start += H * W
        img = img.permute(0, 2, 3, 1).reshape(N * M, D, H, W)          # func.py:65
        g = grid[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)  # func.py:67
        smp = F.grid_sample(img, g, mode="bilinear", padding_mode="zeros",
                            align_corners=False)         # func.py:69-70 -> [N*M, D, Lq, P]
        w_l = per_head_w[:, None, :, lvl * P:(lvl + 1) * P]
        part = (smp * w_l).sum(-1)                       # func.py:73-74
        acc = part if acc is None else acc + part
    return acc.view(N, M * D, Lq).transpose(1, 2).contiguous()         # func.py:75


def core_grid_sample_grads(value, shapes, loc, weight, grad_out):
    """Forward + the three gradients through ``core_grid_sample`` (autograd)."""
    v = value.detach().clone().requires_grad_(True)
    l = loc.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    out = core_grid_sample(v, shapes, l, w)
    out.backward(grad_out)
    return out.detach(), v.grad, l.grad, w.grad
