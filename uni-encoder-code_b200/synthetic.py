"""Deterministic synthetic inputs for the MSDA hot path (no datasets, no checkpoints).

Shapes follow the reference's pixel decoder: levels are fed low-res -> high-res
(msdeformattn.py:341 reverses res3..res5), 8 heads x 32 channels, 4 points
(configs/cityscapes/oneformer_R50_bs16_90k.yaml:9-18, msdeformattn.py:31), and in the
encoder the queries are the value pixels themselves (Lq == S) with reference points at
the pixel centres of the query's own level (msdeformattn.py:152-166) and
``loc = ref + offset / (W_l, H_l)`` (ops/modules/ms_deform_attn.py:109-112).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import torch


def pyramid(height: int, width: int, strides=(32, 16, 8)) -> List[Tuple[int, int]]:
    """(H_l, W_l) per level, low-res first, for an image of height x width."""
    return [(-(-height // s), -(-width // s)) for s in strides]


@dataclass(frozen=True)
class Workload:
    name: str
    levels: Tuple[Tuple[int, int], ...]
    batch: int
    heads: int = 8
    channels: int = 32
    points: int = 4

    @property
    def spatial_size(self) -> int:
        return sum(h * w for h, w in self.levels)

    @property
    def queries(self) -> int:          # per op call (encoder: Lq == S)
        return self.batch * self.spatial_size


# BASELINE.json configs (op-level shapes; see BASELINE.md / SURVEY.md section 8d)
WORKLOADS = {
    "cityscapes_1024x2048_b1": Workload("cityscapes_1024x2048_b1", tuple(pyramid(1024, 2048)), 1),
    "cityscapes_512x1024_b8": Workload("cityscapes_512x1024_b8", tuple(pyramid(512, 1024)), 8),
    "cityscapes_1024x2048_b8": Workload("cityscapes_1024x2048_b8", tuple(pyramid(1024, 2048)), 8),
    "kitti_384x1248_b16": Workload("kitti_384x1248_b16", tuple(pyramid(384, 1248)), 16),
    "cityscapes_512x1024_b16": Workload("cityscapes_512x1024_b16", tuple(pyramid(512, 1024)), 16),
}

FWD_BYTES_PER_QUERY = 3200      # value 1024 + loc 768 + weights 384 + out 1024
BWD_BYTES_PER_QUERY = 5376      # grad_out 1024 + value 1024 + loc 768 + w 384 | gv 1024 + gloc 768 + gw 384


def level_tensors(levels, device="cpu"):
    shapes = torch.tensor(list(levels), dtype=torch.int64, device=device)
    lsi = torch.cat((shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]))
    return shapes, lsi


def reference_points(levels, dtype=torch.float32):
    """[S, 2] (x, y) pixel centres of every query's own level (msdeformattn.py:152-166, valid_ratio 1)."""
    pts = []
    for H, W in levels:
        ys = (torch.arange(H, dtype=dtype) + 0.5) / H
        xs = (torch.arange(W, dtype=dtype) + 0.5) / W
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    return torch.cat(pts, 0)


def make_inputs(levels, batch, heads=8, channels=32, points=4, num_query=None, mode="model",
                seed=0, dtype=torch.float32, device="cpu", with_grad_output=True):
    """Seeded tensors for one op call.

    mode "model":   encoder-like, spatially coherent: offsets ~ N(0, 2^2) px clipped to +-8 px
                    plus U(-0.5, 0.5) px jitter (keeps coordinates off the integer lattice,
                    where floor() makes grad_sampling_loc discontinuous). Requires Lq == S.
    mode "uniform": loc ~ U(-0.1, 1.1): no locality, ~27-30 % of points out of bounds.
    Returns dict(value, spatial_shapes, level_start_index, sampling_locations,
                 attention_weights[, grad_output]).
    """
    gen = torch.Generator().manual_seed(int(seed))
    L = len(levels)
    S = sum(h * w for h, w in levels)
    Lq = S if num_query is None else int(num_query)
    value = torch.randn(batch, S, heads, channels, generator=gen, dtype=dtype)
    logits = torch.randn(batch, Lq, heads, L * points, generator=gen, dtype=dtype)
    weights = torch.softmax(logits, -1).view(batch, Lq, heads, L, points)
    if mode == "model":
        if Lq != S:
            raise ValueError('mode "model" needs num_query == spatial size')
        ref = reference_points(levels, dtype)                              # [S, 2]
        off = torch.randn(batch, Lq, heads, L, points, 2, generator=gen, dtype=dtype) * 2.0
        off = off.clamp_(-8.0, 8.0)
        off += torch.rand(batch, Lq, heads, L, points, 2, generator=gen, dtype=dtype) - 0.5
        norm = torch.tensor([[w, h] for h, w in levels], dtype=dtype)      # (W_l, H_l)
        loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    elif mode == "uniform":
        loc = torch.rand(batch, Lq, heads, L, points, 2, generator=gen, dtype=dtype) * 1.2 - 0.1
    else:
        raise ValueError(mode)
    shapes, lsi = level_tensors(levels)
    out = dict(value=value, spatial_shapes=shapes, level_start_index=lsi,
               sampling_locations=loc.contiguous(), attention_weights=weights.contiguous())
    if with_grad_output:
        out["grad_output"] = torch.randn(batch, Lq, heads * channels, generator=gen, dtype=dtype)
    if str(device) != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


def make_workload_inputs(name, mode="model", seed=0, device="cpu", batch=None):
    w = WORKLOADS[name]
    return make_inputs(w.levels, w.batch if batch is None else batch, w.heads, w.channels, w.points,
                       mode=mode, seed=seed, device=device)
