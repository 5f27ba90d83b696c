"""Multi-GPU partitioning of the MSDA path.  One process per GPU,
``torch.distributed`` for the plumbing; collectives only where the path has a real exchange.

* **Batch sharding** (training and batched inference): every image is independent in forward and
  backward (the reference already chunks the batch, ms_deform_attn_cuda.cu:66-80), so ranks own
  disjoint images and the op needs no collective.  Training adds the usual DDP gradient all-reduce
  (the reference's only strategy, tools/trainers/trainer.py:110).
* **Query-range sharding** (single-image inference): a rank owns a contiguous range of the
  ``Lq == S`` query/pixel rows.  Per-row work (projections, softmax, LayerNorm, FFN) stays local;
  the only exchange is one all-gather of the projected ``value`` rows per encoder layer, because a
  query may sample any pixel.  The op is then called with the full ``value`` and the rank's slice
  of ``sampling_locations`` / ``attention_weights``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .modules import MSDeformAttnTransformerEncoderOnly, _add_norm, _linear, reference_points_for


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of `total` items for `rank` (first ranks get the remainder)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensors: Dict[str, torch.Tensor], rank: int, world: int,
                replicated=("spatial_shapes", "level_start_index")) -> Dict[str, torch.Tensor]:
    """Slice dim 0 of every per-image tensor; the level tables are replicated."""
    n = next(v for k, v in tensors.items() if k not in replicated).shape[0]
    a, b = shard_range(n, rank, world)
    return {k: (v if k in replicated else v[a:b].contiguous()) for k, v in tensors.items()}


class RowGather:
    """All-gather of per-rank row blocks ``[N, rows_r, C]`` along dim 1 (ragged sizes allowed), started
    asynchronously: ``RowGather(local, sizes, group)`` issues the collective (NCCL runs it on its own
    stream), ``.result()`` makes the current stream wait for it and returns ``[N, sum(rows), C]``.
    Work that only needs the local rows goes between the two calls."""

    def __init__(self, local: torch.Tensor, sizes: List[int], group=None):
        world = dist.get_world_size(group)
        n = local.shape[0]
        self.sizes, self.n, self.world = sizes, n, world
        self.even = len(set(sizes)) == 1
        if self.even:
            send = local.contiguous()
        else:
            # ragged: pad every block to the largest one (one collective; the backends' list-based
            # all_gather does not accept uneven sizes everywhere), the padding is dropped in result()
            send = local.new_zeros((n, max(sizes)) + tuple(local.shape[2:]))
            send[:, :local.shape[1]] = local
        self.block = tuple(send.shape)
        self.out = local.new_empty((world * n,) + self.block[1:])        # rank-major along dim 0
        self.work = dist.all_gather_into_tensor(self.out, send, group=group, async_op=True)

    def result(self) -> torch.Tensor:
        self.work.wait()
        if self.even and self.n == 1:
            # one image: rank-major blocks of rows ARE the concatenation along dim 1 -- no copy
            return self.out.view((1, self.world * self.block[1]) + self.block[2:])
        blocks = self.out.view((self.world,) + self.block)
        if self.even:
            return torch.cat(list(blocks.unbind(0)), 1)
        return torch.cat([blocks[r, :, :sz] for r, sz in enumerate(self.sizes)], 1)


def all_gather_rows(local: torch.Tensor, sizes: List[int], group=None) -> torch.Tensor:
    """Blocking form of :class:`RowGather`."""
    return RowGather(local, sizes, group).result()


class QueryShardedEncoder:
    """Inference-only query-range sharding of an ``MSDeformAttnTransformerEncoderOnly``.

    Every rank holds the same weights and the same flattened inputs; rank r keeps rows
    ``shard_range(S, r, world)`` of the running ``src``.  Per layer: local ``value_proj`` on the
    rank's rows, all-gather of ``value`` (S*256*4 bytes in total) started asynchronously and
    overlapped with the sampling-offset / attention-weight projections and softmax of the local rows
    (they do not depend on it), the MSDA op on the rank's queries against the full ``value``, then
    the local output projection / LayerNorm / FFN.
    A final all-gather returns the full memory on every rank.
    """

    def __init__(self, encoder_only: MSDeformAttnTransformerEncoderOnly, group=None):
        self.model = encoder_only
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    @torch.no_grad()
    def forward(self, srcs, pos_embeds):
        m = self.model
        src, pos, shapes, lsi, levels = m.flatten_inputs(srcs, pos_embeds)
        S = src.shape[1]
        sizes = [b - a for a, b in (shard_range(S, r, self.world) for r in range(self.world))]
        a, b = shard_range(S, self.rank, self.world)
        ref = reference_points_for(levels, src.device)[:, a:b].expand(src.shape[0], -1, -1, -1)
        x, p = src[:, a:b], pos[:, a:b]
        x = x.contiguous()
        for layer in m.encoder.layers:
            attn = layer.self_attn
            gather = RowGather(_linear(attn.value_proj, x, attn.linear), sizes, self.group)      # in flight ...

            def full_value(gather=gather, attn=attn):
                v = gather.result()
                return v.view(v.shape[0], S, attn.n_heads, attn.d_model // attn.n_heads)

            if attn.can_fold_pos(x, p, ref):
                # ... while the stacked query projection of the local rows runs; the same kernels, the same
                # per-row arithmetic as the unsharded encoder: results are bit-identical
                out = attn.forward_shared_pos(x, p, ref, shapes, lsi, value=full_value)
            else:
                loc, w = attn.sampling_inputs(x + p, ref, shapes)          # ... while the local rows' producers run
                out = _linear(attn.output_proj,
                              attn._core(full_value(), shapes, lsi, loc.contiguous(), w.contiguous(), attn.im2col_step),
                              attn.linear)
            x = layer.forward_ffn(_add_norm(layer.norm1, x, layer.dropout1(out), layer.linear))
        return all_gather_rows(x, sizes, self.group), shapes, lsi

    __call__ = forward


def ddp_wrap(module: torch.nn.Module, device: Optional[torch.device] = None):
    """DistributedDataParallel with the reference's settings (broadcast_buffers=False,
    tools/trainers/trainer.py:110); gradients are all-reduced over NCCL in 25 MB buckets."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    if device is not None and device.type == "cuda":
        return DDP(module, device_ids=[device.index], broadcast_buffers=False)
    return DDP(module, broadcast_buffers=False)
