"""torchrun worker (one rank per GPU, NCCL): the multi-GPU modes of SURVEY.md section 8e on real devices.

  1. batch sharding of the op: per-rank images, outputs / gradients equal the unsharded call;
  2. query-range sharded single-image encoder inference == unsharded encoder;
  3. DDP training step (NCCL gradient all-reduce) == the full batch on one GPU.

Prints one JSON line on rank 0; any mismatch raises (non-zero exit)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = load_package()
    res = {"world": world}

    # ---- 1. batch sharding of the op (no collective on the data path)
    levels = [(16, 32), (32, 64), (64, 128)]
    inp = pkg.synthetic.make_inputs(levels, batch=2 * world, mode="model", seed=21)
    full = {k: v.to(dev) for k, v in inp.items()}
    mine = pkg.sharding.shard_batch(full, rank, world)
    a = lambda d: (d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_locations"], d["attention_weights"])
    out = pkg.ms_deform_attn_forward(*a(mine), 128)
    gv, gl, gw = pkg.ms_deform_attn_backward(*a(mine), mine["grad_output"], 128)
    ref_out = pkg.ms_deform_attn_forward(*a(full), 128)
    rgv, rgl, rgw = pkg.ms_deform_attn_backward(*a(full), full["grad_output"], 128)
    lo, hi = pkg.sharding.shard_range(2 * world, rank, world)
    assert torch.equal(out, ref_out[lo:hi]) and torch.equal(gl, rgl[lo:hi]) and torch.equal(gw, rgw[lo:hi])
    err = ((gv - rgv[lo:hi]).abs().max() / rgv.abs().max()).item()
    assert err <= 1e-5, err                       # grad_value: reduction order only
    res["batch_sharding_grad_value_rel"] = err

    # ---- 2. query-range sharding, single image
    torch.manual_seed(7)
    kw = dict(d_model=256, nhead=8, num_encoder_layers=3, dim_feedforward=512, dropout=0.0,
              num_feature_levels=3, enc_n_points=4)
    enc = pkg.modules.MSDeformAttnTransformerEncoderOnly(**kw).to(dev).eval()
    with torch.no_grad():
        for layer in enc.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(std=0.02)
            layer.self_attn.attention_weights.weight.normal_(std=0.05)
    for p in enc.parameters():                    # same weights everywhere
        dist.broadcast(p.data, 0)
    gen = torch.Generator().manual_seed(3)
    srcs = [torch.randn(1, 256, h, w, generator=gen).to(dev) for h, w in levels]
    pos = [(torch.randn(1, 256, h, w, generator=gen) * 0.1).to(dev) for h, w in levels]
    with torch.no_grad():
        whole = enc(srcs, pos)[0]
        sharded = pkg.sharding.QueryShardedEncoder(enc)(srcs, pos)[0]
    res["query_sharding_max_abs_diff_vs_unsharded"] = (whole - sharded).abs().max().item()
    assert res["query_sharding_max_abs_diff_vs_unsharded"] <= 2e-5       # row-wise GEMMs on fewer rows: at most rounding

    # ---- 3. DDP step vs the full batch on one GPU
    torch.manual_seed(11)
    tr = pkg.modules.MSDeformAttnTransformerEncoderOnly(**kw).to(dev)
    with torch.no_grad():
        for layer in tr.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(std=0.02)
            layer.self_attn.attention_weights.weight.normal_(std=0.05)
    for p in tr.parameters():
        dist.broadcast(p.data, 0)
    import copy
    single = copy.deepcopy(tr)
    ddp = pkg.sharding.ddp_wrap(tr, dev)
    B = 2 * world
    gen = torch.Generator().manual_seed(5)
    bs = [torch.randn(B, 256, h, w, generator=gen).to(dev) for h, w in levels]
    bp = [(torch.randn(B, 256, h, w, generator=gen) * 0.1).to(dev) for h, w in levels]
    lo, hi = pkg.sharding.shard_range(B, rank, world)
    ddp([s[lo:hi] for s in bs], [p[lo:hi] for p in bp])[0].square().mean().backward()
    single(bs, bp)[0].square().mean().backward()
    worst = 0.0
    for (k, p), q in zip(tr.named_parameters(), single.parameters()):
        worst = max(worst, ((p.grad - q.grad).abs().max() / q.grad.abs().max().clamp_min(1e-20)).item())
    res["ddp_grad_rel_vs_single_gpu"] = worst
    assert worst <= 5e-4, worst
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
