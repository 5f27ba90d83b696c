"""Host-side mirror of the reference's ``MSDeformAttnPixelDecoder``
(model/modeling/pixel_decoder/msdeformattn.py:178-386) -- the caller that owns the MSDA hot path:
1x1 input projections + GroupNorm, sine position embedding, the 6-layer deformable encoder, the
per-level split and one FPN level on the stride-4 feature.

Same submodule / parameter names as the reference (``input_proj.{i}.{0,1}``, ``transformer.*``,
``mask_features``, ``adapter_1``, ``layer_1``), so its checkpoints load with ``load_state_dict``.
Constructor arguments are the explicit ones of the reference's ``__init__`` (no detectron2 config
object); ``input_shape`` maps a feature name to ``(channels, stride)``.

Differences, all from SURVEY.md section 8f.4 (input-independent work and host syncs):
* the position embedding of a level depends only on (H, W): computed once per shape / device and
  cached (the reference recomputes it every forward, position_encoding.py:32-55);
* the level split uses the host-side shapes instead of indexing the device-side
  ``level_start_index`` / ``spatial_shapes`` tensors (msdeformattn.py:351-365 -> ``.item()`` syncs).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .modules import CoreFn, MSDeformAttnTransformerEncoderOnly, ShapeCache, _capturing


class PositionEmbeddingSine(nn.Module):
    """transformer_decoder/position_encoding.py:15-55 for the mask-free case, cached per shape."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and not normalize:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats, self.temperature, self.normalize = num_pos_feats, temperature, normalize
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = ShapeCache()        # per instance, most recent shapes only

    def _build(self, H, W, device):
        ones = torch.ones((1, H, W), dtype=torch.float32, device=device)
        y_embed, x_embed = ones.cumsum(1), ones.cumsum(2)
        if self.normalize:
            eps = 1e-6
            y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
            x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_pos_feats)
        px, py = x_embed[:, :, :, None] / dim_t, y_embed[:, :, :, None] / dim_t
        px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
        py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
        return torch.cat((py, px), dim=3).permute(0, 3, 1, 2).contiguous()      # [1, 2*npf, H, W]

    def forward(self, x, mask=None):
        if mask is not None:
            raise NotImplementedError("the pixel decoder never passes a mask (msdeformattn.py:345)")
        key = (x.shape[2], x.shape[3], str(x.device))
        pos = self._cache.get(key)
        if pos is None:
            pos = self._build(x.shape[2], x.shape[3], x.device)
            if not _capturing(x):                      # built during a graph capture: not materialised yet, not kept
                self._cache.put(key, pos)
        return pos.expand(x.shape[0], -1, -1, -1)


class _ConvNormAct(nn.Conv2d):
    """Conv2d followed by an optional norm and activation (what detectron2.layers.Conv2d does for
    the reference's lateral / output convolutions, msdeformattn.py:286-300)."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm, self.activation = norm, activation

    def forward(self, x, up=None, fused_norm=False):
        """conv -> norm -> activation (-> + bilinear up-sampling of `up`, the FPN top-down add of
        msdeformattn.py:372-376).  `fused_norm` (inference, GroupNorm, fp32 CUDA): the norm, the ReLU and
        the up-sampled add run as one statistics + one apply kernel (ops.group_norm)."""
        x = super().forward(x)
        if fused_norm and not torch.is_grad_enabled() and self.activation in (None, F.relu) \
                and ops.group_norm_supported(x, self.norm, up):
            return ops.group_norm(x, self.norm, relu=self.activation is F.relu, up=up)
        if self.norm is not None:
            x = self.norm(x)
        x = x if self.activation is None else self.activation(x)
        if up is not None:
            x = x + F.interpolate(up, size=x.shape[-2:], mode="bilinear", align_corners=False)
        return x


def _norm(kind, channels):
    if kind in (None, ""):
        return None
    if kind == "GN":
        return nn.GroupNorm(32, channels)
    raise NotImplementedError(f"norm {kind!r}")


def _c2_xavier_fill(m):
    nn.init.kaiming_uniform_(m.weight, a=1)
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


class MSDeformAttnPixelDecoder(nn.Module):
    def __init__(self, input_shape: Dict[str, Tuple[int, int]], *, transformer_dropout: float,
                 transformer_nheads: int, transformer_dim_feedforward: int, transformer_enc_layers: int,
                 conv_dim: int, mask_dim: int, norm: Optional[str] = None,
                 transformer_in_features: Sequence[str], common_stride: int,
                 core: Optional[CoreFn] = None, fused: bool = False, linear: str = "torch"):
        super().__init__()
        by_stride = sorted(input_shape.items(), key=lambda kv: kv[1][1])
        self.in_features = [k for k, _ in by_stride]
        self.feature_channels = [v[0] for _, v in by_stride]
        tr = [(k, v) for k, v in by_stride if k in transformer_in_features]
        self.transformer_in_features = [k for k, _ in tr]
        self.transformer_feature_strides = [v[1] for _, v in tr]
        self.transformer_num_feature_levels = len(tr)
        chans = [v[0] for _, v in tr]
        chans = chans[::-1] if len(tr) > 1 else chans[-1:]          # low -> high resolution
        self.input_proj = nn.ModuleList(
            nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim)) for c in chans)
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
            dim_feedforward=transformer_dim_feedforward, num_encoder_layers=transformer_enc_layers,
            num_feature_levels=self.transformer_num_feature_levels, core=core, fused=fused, linear=linear)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, normalize=True)
        self.linear = linear
        self.mask_dim = mask_dim
        self.mask_features = nn.Conv2d(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        _c2_xavier_fill(self.mask_features)
        self.oneformer_num_feature_levels = 3
        self.common_stride = common_stride
        self.num_fpn_levels = int(math.log2(min(self.transformer_feature_strides)) - math.log2(common_stride))
        lateral, output = [], []
        use_bias = norm == ""
        for idx, c in enumerate(self.feature_channels[:self.num_fpn_levels]):
            lat = _ConvNormAct(c, conv_dim, kernel_size=1, bias=use_bias, norm=_norm(norm, conv_dim))
            out = _ConvNormAct(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=use_bias,
                               norm=_norm(norm, conv_dim), activation=F.relu)
            _c2_xavier_fill(lat)
            _c2_xavier_fill(out)
            self.add_module(f"adapter_{idx + 1}", lat)
            self.add_module(f"layer_{idx + 1}", out)
            lateral.append(lat)
            output.append(out)
        self.lateral_convs, self.output_convs = lateral[::-1], output[::-1]

    def forward_features(self, features: Dict[str, torch.Tensor]):
        """-> (mask_features [N, mask_dim, H/4, W/4], lowest-resolution map, 3 multi-scale maps),
        as msdeformattn.py:336-386 (autocast disabled, inputs promoted to fp32)."""
        with torch.autocast(device_type=next(iter(features.values())).device.type, enabled=False):
            srcs, pos = [], []
            fused_norm = self.linear == "tf32x3"        # the inference kernels of SURVEY 8f.3 / 8f.4
            names = self.transformer_in_features[::-1]
            levels = [(int(features[k].shape[2]), int(features[k].shape[3])) for k in names]
            # inference kernels: every input projection's GroupNorm writes its rows of the concatenated
            # [N, S, C] tensor directly (no transposing `cat` afterwards) when all levels qualify
            src_flat, starts = None, [0]
            for h, w in levels:
                starts.append(starts[-1] + h * w)
            rows_ok = fused_norm and not torch.is_grad_enabled()
            for idx, name in enumerate(names):
                x = features[name].float()
                proj = self.input_proj[idx]
                if fused_norm and not torch.is_grad_enabled():
                    # the convolution runs without its bias (torch adds it in a separate pass over the map);
                    # the GroupNorm kernels add it on the fly
                    y = F.conv2d(x, proj[0].weight, None, proj[0].stride, proj[0].padding, proj[0].dilation,
                                 proj[0].groups)
                    if rows_ok and src_flat is None:
                        src_flat = y.new_empty((y.shape[0], starts[-1], y.shape[1]))
                    rows = src_flat[:, starts[idx]:starts[idx + 1]] if rows_ok else None
                    if rows_ok and ops.group_norm_rows_supported(y, proj[1], rows):
                        ops.group_norm_rows(y, proj[1], rows, channel_bias=proj[0].bias)
                        srcs.append(None)
                    else:
                        if ops.group_norm_supported(y, proj[1]):
                            z = ops.group_norm(y, proj[1], channel_bias=proj[0].bias)
                        else:
                            z = proj[1](y if proj[0].bias is None else y + proj[0].bias.view(1, -1, 1, 1))
                        srcs.append(z)
                        if rows_ok:                                   # keep the rows complete: this level through torch
                            rows.copy_(z.flatten(2).transpose(1, 2))
                else:
                    srcs.append(proj[1](proj[0](x)))
                pos.append(self.pe_layer(x))
            if rows_ok:
                memory, _, _, _ = self.transformer(None, pos, src_flat=src_flat, levels=levels)
            else:
                memory, _, _, _ = self.transformer(srcs, pos)
            n = memory.shape[0]
            out: List[torch.Tensor] = [z.transpose(1, 2).reshape(n, -1, h, w) for z, (h, w) in
                                       zip(memory.split([h * w for h, w in levels], dim=1), levels)]
            for idx, name in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
                y = self.lateral_convs[idx](features[name].float(), up=self._nchw(out[-1], memory, fused_norm),
                                            fused_norm=fused_norm)
                out.append(self.output_convs[idx](y, fused_norm=fused_norm))
            return self._mask_features(out[-1], fused_norm), out[0], out[:self.oneformer_num_feature_levels]

    @staticmethod
    def _nchw(view, memory, fused_norm):
        """Contiguous NCHW copy of a level of the encoder memory (`view` = [N, C, H, W] strided over the
        [N, S, C] memory): one [pixels, C] -> [C, pixels] transpose kernel per image in inference."""
        if view.is_contiguous():
            return view
        n, c, h, w = view.shape
        if not (fused_norm and not torch.is_grad_enabled() and view.is_cuda and view.dtype == torch.float32
                and view.stride(1) == 1 and view.stride(3) == c and view.stride(2) == w * c):
            return view.contiguous()
        dst = torch.empty((n, c, h, w), dtype=view.dtype, device=view.device)
        for i in range(n):
            # view[i] is the transpose of the contiguous [h*w, c] block of rows of image i
            rows = torch.as_strided(view, (h * w, c), (c, 1), view[i].storage_offset())
            ops.transpose2d(rows, out=dst[i])
        return dst

    def _mask_features(self, x, fused_norm):
        conv = self.mask_features
        if fused_norm and not torch.is_grad_enabled() and conv.bias is not None:
            y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
            if ops.channel_bias_supported(y, conv.bias):
                return ops.add_channel_bias_(y, conv.bias)
            return y + conv.bias.view(1, -1, 1, 1)
        return conv(x)
