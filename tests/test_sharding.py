"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in uni-encoder-code_b200/sharding.py.
The CUDA op cannot run here, so the core op is injected from the oracle (test infrastructure); what
is under test is the partitioning, the per-layer all-gather, and the DDP gradient all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEVELS = [(2, 3), (4, 6), (7, 9)]       # S = 93: odd, so the two query shards are ragged


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load_oracle, load_package
    pkg, oracle = load_package(), load_oracle()

    def core(value, shapes, lsi, loc, w, im2col_step):
        return oracle.core_grid_sample(value, shapes, loc, w)
    return pkg, oracle, core


def _encoder(pkg, core, seed=5):
    torch.manual_seed(seed)
    m = pkg.modules.MSDeformAttnTransformerEncoderOnly(
        d_model=64, nhead=2, num_encoder_layers=2, dim_feedforward=96, dropout=0.0,
        num_feature_levels=3, enc_n_points=4, core=core).double()
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():       # move offsets off the integer lattice of the default init
        for layer in m.encoder.layers:
            a = layer.self_attn
            a.sampling_offsets.weight.copy_(torch.randn(a.sampling_offsets.weight.shape, generator=gen,
                                                        dtype=torch.float64) * 0.05)
            a.attention_weights.weight.copy_(torch.randn(a.attention_weights.weight.shape, generator=gen,
                                                         dtype=torch.float64) * 0.1)
    return m


def _inputs(batch, seed=11):
    gen = torch.Generator().manual_seed(seed)
    srcs = [torch.randn(batch, 64, h, w, generator=gen, dtype=torch.float64) for h, w in LEVELS]
    pos = [torch.randn(batch, 64, h, w, generator=gen, dtype=torch.float64) * 0.1 for h, w in LEVELS]
    return srcs, pos


def _worker_query_sharding(rank, world, port, ret):
    pkg, oracle, core = _setup(rank, world, port)
    enc = _encoder(pkg, core).eval()
    srcs, pos = _inputs(1)
    with torch.no_grad():
        full = enc(srcs, pos)[0]
    sharded, shapes, lsi = pkg.sharding.QueryShardedEncoder(enc)(srcs, pos)
    err = (sharded - full).abs().max().item()
    a, b = pkg.sharding.shard_range(full.shape[1], rank, world)
    ret[rank] = (err, a, b, tuple(sharded.shape))
    dist.destroy_process_group()


def _worker_batch_and_ddp(rank, world, port, ret):
    pkg, oracle, core = _setup(rank, world, port)
    # --- batch sharding of the op: concatenated per-rank outputs == unsharded output
    inp = pkg.synthetic.make_inputs(LEVELS, batch=4, heads=2, channels=32, points=4, mode="model",
                                    seed=3, dtype=torch.float64)
    mine = pkg.sharding.shard_batch(inp, rank, world)
    assert mine["value"].shape[0] == 2 and mine["spatial_shapes"] is inp["spatial_shapes"]
    out_local = oracle.core_grid_sample(mine["value"], mine["spatial_shapes"], mine["sampling_locations"],
                                        mine["attention_weights"])
    parts = [torch.empty_like(out_local) for _ in range(world)]
    dist.all_gather(parts, out_local)
    full = oracle.core_grid_sample(inp["value"], inp["spatial_shapes"], inp["sampling_locations"],
                                   inp["attention_weights"])
    err_batch = (torch.cat(parts, 0) - full).abs().max().item()

    # --- DDP step: all-reduced gradients == gradients of the full batch on one process
    enc = _encoder(pkg, core)
    ddp = pkg.sharding.ddp_wrap(enc)
    srcs, pos = _inputs(4)
    a, b = pkg.sharding.shard_range(4, rank, world)
    mem = ddp([s[a:b] for s in srcs], [p[a:b] for p in pos])[0]
    mem.square().mean().backward()
    ref = _encoder(pkg, core)
    ref(srcs, pos)[0].square().mean().backward()
    err_ddp = max((p.grad - q.grad).abs().max().item() / max(q.grad.abs().max().item(), 1e-30)
                  for p, q in zip(enc.parameters(), ref.parameters()))
    ret[rank] = (err_batch, err_ddp)
    dist.destroy_process_group()


def _spawn(fn):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(fn, args=(2, port, ret), nprocs=2, join=True)
    return dict(ret)


def test_shard_range_is_a_balanced_partition(pkg):
    sr = pkg.sharding.shard_range
    for total in (0, 1, 7, 93, 10752, 43008):
        for world in (1, 2, 3, 4, 8):
            spans = [sr(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sr(10, 2, 2)


@pytest.mark.timeout(300)
def test_query_range_sharding_equals_unsharded_encoder_gloo_world2():
    ret = _spawn(_worker_query_sharding)
    assert set(ret) == {0, 1}
    for rank, (err, a, b, shape) in ret.items():
        assert err <= 1e-12, (rank, err)
        assert shape == (1, 93, 64)
    assert ret[0][1:3] == (0, 47) and ret[1][1:3] == (47, 93)      # ragged shards


@pytest.mark.timeout(300)
def test_batch_sharding_and_ddp_allreduce_gloo_world2():
    ret = _spawn(_worker_batch_and_ddp)
    for rank, (err_batch, err_ddp) in ret.items():
        assert err_batch <= 1e-12, (rank, err_batch)
        assert err_ddp <= 1e-10, (rank, err_ddp)
