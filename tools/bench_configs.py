#!/usr/bin/env python
"""All five BASELINE.json configs on 1..N B200s (one process per GPU, torchrun for N > 1).

    python tools/bench_configs.py [--configs 1,2,3,4,5,q] [--steps 10] [--out gpurun_out/configs.jsonl]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... tools/bench_configs.py

 1  single MSDA forward, 1024x2048 pyramid, batch 1                      (op level)
 2  MSDA forward+backward, 512x1024 crop, batch 8 per GPU                (op level; bench.py headline)
 3  MSDeformAttnPixelDecoder.forward_features (input proj + 6-layer encoder + FPN), 1024x2048,
    batch 8 per GPU (batch-sharded), Swin-T feature shapes
 4  KITTI 384x1248 two-frame input (8 pairs = batch 16 per GPU), same pixel-decoder forward
 5  encoder training step (fwd+bwd+AdamW) at 512x1024, batch 16 per GPU, DDP all-reduce over NCCL
 q  query-range sharded single-image encoder inference, 1024x2048, batch 1 over all ranks

For 3-5 and q the line reports (i) MSDA-only time summed over the 6 layers (CUDA events around the
library launches), (ii) whole step time, (iii) the MSDA roofline fraction from (i) -- the rest of the
encoder is torch Linear / LayerNorm (cuBLAS), reported for context.
Timing: CUDA events, max over ranks, 3 warm-ups.  The encoder is uni-encoder-code_b200/modules.py, the
host-side mirror of the reference's MSDeformAttnTransformerEncoderOnly (random init, synthetic features).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from bench import measured_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--fused", action="store_true", help="use the fused-producer kernels inside the encoder mirror")
    ap.add_argument("--linear", default="torch", choices=["torch", "tf32x3"],
                    help="nn.Linear implementation inside the encoder mirror in inference (tf32x3: tensor-core GEMM)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = load_package()
    syn = pkg.synthetic
    peak, _ = measured_peak()

    def reduce_max(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(rec):
        rec.update(n_gpus=world, fused_producers=bool(args.fused), linear=args.linear)
        if rank == 0:
            os.makedirs(os.path.dirname(args.out), exist_ok=True)
            with open(args.out, "a") as f:
                f.write(json.dumps(rec) + "\n")
            print(json.dumps(rec), flush=True)

    def timed(fn, steps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return reduce_max(e0.elapsed_time(e1) / steps)

    def msda_ms(events, steps):
        torch.cuda.synchronize()
        f = sum(a.elapsed_time(b) for k, a, b in events if k == "forward") / steps
        b = sum(a.elapsed_time(b_) for k, a, b_ in events if k == "backward") / steps
        return reduce_max(f), reduce_max(b)

    def features(levels, batch, seed):
        gen = torch.Generator().manual_seed(seed + rank)
        srcs = [torch.randn(batch, 256, h, w, generator=gen).to(dev) for h, w in levels]
        pos = [(torch.randn(batch, 256, h, w, generator=gen) * 0.1).to(dev) for h, w in levels]
        return srcs, pos

    def encoder(seed=0):
        torch.manual_seed(seed)
        m = pkg.modules.MSDeformAttnTransformerEncoderOnly(
            d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1,
            num_feature_levels=3, enc_n_points=4, fused=args.fused, linear=args.linear).to(dev)
        # trained-model-like sampling: small learned offsets instead of the integer-lattice init
        gen = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for layer in m.encoder.layers:
                a = layer.self_attn
                a.sampling_offsets.weight.copy_((torch.randn(a.sampling_offsets.weight.shape, generator=gen) * 0.02).to(dev))
                a.attention_weights.weight.copy_((torch.randn(a.attention_weights.weight.shape, generator=gen) * 0.05).to(dev))
        return m

    for cfg in args.configs.split(","):
        if cfg == "1":
            d = syn.make_workload_inputs("cityscapes_1024x2048_b1", seed=1 + rank, device=dev)
            a = (d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_locations"], d["attention_weights"])
            ms = timed(lambda: pkg.ms_deform_attn_forward(*a, 128), args.steps * 5)
            q = world * 43008
            emit({"config": 1, "what": "MSDA forward, 1024x2048 pyramid, batch 1 per GPU (replicas)", "ms": ms,
                  "queries_per_s": q / ms * 1e3, "algorithmic_GBs_per_gpu": 43008 * 3200 / ms / 1e6,
                  "frac_of_measured_hbm": 43008 * 3200 / ms / 1e6 / peak})
        elif cfg == "2":
            d = syn.make_workload_inputs("cityscapes_512x1024_b8", seed=1 + rank, device=dev)
            a = (d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_locations"], d["attention_weights"])

            def step():
                pkg.ms_deform_attn_forward(*a, 128)
                pkg.ms_deform_attn_backward(*a, d["grad_output"], 128)
            ms = timed(step, args.steps * 3)
            q = 86016
            emit({"config": 2, "what": "MSDA forward+backward via the plugin functions, 512x1024, batch 8 per GPU", "ms": ms,
                  "queries_per_s": world * q / ms * 1e3, "algorithmic_GBs_per_gpu": q * 8576 / ms / 1e6,
                  "frac_of_measured_hbm": q * 8576 / ms / 1e6 / peak})
        elif cfg in ("3", "4"):
            (Himg, Wimg), batch, name = (((1024, 2048), 8, "pixel-decoder forward_features 1024x2048, batch 8 per GPU") if cfg == "3"
                                         else ((384, 1248), 16, "KITTI 384x1248 two-frame (8 pairs = batch 16 per GPU) pixel-decoder forward_features"))
            # Swin-T feature pyramid: (channels, stride) per backbone output (model/modeling/backbone/swin.py:730-741)
            shapes = {"res2": (96, 4), "res3": (192, 8), "res4": (384, 16), "res5": (768, 32)}
            torch.manual_seed(0)
            dec = pkg.pixel_decoder.MSDeformAttnPixelDecoder(
                shapes, transformer_dropout=0.1, transformer_nheads=8, transformer_dim_feedforward=1024,
                transformer_enc_layers=6, conv_dim=256, mask_dim=256, norm="GN",
                transformer_in_features=["res3", "res4", "res5"], common_stride=4, fused=args.fused,
                linear=args.linear).to(dev).eval()
            gen = torch.Generator().manual_seed(10 + rank)
            feats = {k: torch.randn(batch, c, -(-Himg // st), -(-Wimg // st), generator=gen).to(dev)
                     for k, (c, st) in shapes.items()}
            with torch.no_grad():
                ms = timed(lambda: dec.forward_features(feats), args.steps)
                # the same call replayed from one CUDA graph (the mirror's forward has no host sync, so it
                # records; tests/test_pixel_decoder.py::test_forward_features_is_cuda_graph_capturable)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    dec.forward_features(feats)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    dec.forward_features(feats)
                ms_graph = timed(graph.replay, args.steps)
                del graph
                ev = pkg.ops.enable_timing(True)
                for _ in range(args.steps):
                    dec.forward_features(feats)
                f_ms, _ = msda_ms(ev, args.steps)
                pkg.ops.enable_timing(False)
            S = sum(h * w for h, w in syn.pyramid(Himg, Wimg))
            q = batch * S
            emit({"config": int(cfg), "what": name, "ms": ms, "ms_cuda_graph": ms_graph, "msda_ms_6_layers": f_ms, "msda_share": f_ms / ms,
                  "queries_per_layer_per_gpu": q, "images_per_s": world * batch / ms * 1e3,
                  "msda_algorithmic_GBs": 6 * q * 3200 / f_ms / 1e6, "msda_frac_of_measured_hbm": 6 * q * 3200 / f_ms / 1e6 / peak})
            del dec, feats
        elif cfg == "5":
            levels, batch = syn.pyramid(512, 1024), 16
            m = encoder().train()
            model = pkg.sharding.ddp_wrap(m, dev) if world > 1 else m
            opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
            srcs, pos = features(levels, batch, 20)

            def step():
                opt.zero_grad(set_to_none=True)
                model(srcs, pos)[0].square().mean().backward()
                opt.step()
            ms = timed(step, args.steps)
            ev = pkg.ops.enable_timing(True)
            for _ in range(args.steps):
                step()
            f_ms, b_ms = msda_ms(ev, args.steps)
            pkg.ops.enable_timing(False)
            q = batch * sum(h * w for h, w in levels)
            nparam = sum(p.numel() for p in m.parameters())
            emit({"config": 5, "what": "encoder training step (fwd+bwd+AdamW), 512x1024, batch 16 per GPU, DDP all-reduce (NCCL)",
                  "ms": ms, "msda_fwd_ms_6_layers": f_ms, "msda_bwd_ms_6_layers": b_ms, "msda_share": (f_ms + b_ms) / ms,
                  "queries_per_layer_per_gpu": q, "images_per_s": world * batch / ms * 1e3,
                  "allreduce_bytes": nparam * 4, "msda_algorithmic_GBs": 6 * q * 8576 / (f_ms + b_ms) / 1e6,
                  "msda_frac_of_measured_hbm": 6 * q * 8576 / (f_ms + b_ms) / 1e6 / peak})
            del m, model, opt, srcs, pos
        elif cfg == "f":
            # op level, configs[1] shape: what the module does around the op (softmax + location
            # arithmetic in torch, then the op) against the fused kernels, forward + backward
            w = syn.WORKLOADS["cityscapes_512x1024_b8"]
            gen = torch.Generator().manual_seed(40 + rank)
            N, S, M, L, P = w.batch, w.spatial_size, 8, 3, 4
            value = torch.randn(N, S, M, 32, generator=gen).to(dev).requires_grad_(True)
            off = (torch.randn(N, S, M, L, P, 2, generator=gen) * 2 + torch.rand(N, S, M, L, P, 2, generator=gen) - 0.5).to(dev).requires_grad_(True)
            logits = torch.randn(N, S, M, L * P, generator=gen).to(dev).requires_grad_(True)
            go = torch.randn(N, S, M * 32, generator=gen).to(dev)
            shapes, lsi = (t.to(dev) for t in syn.level_tensors(w.levels))
            ref = pkg.modules.reference_points_for(w.levels, dev)
            wh = shapes.flip(-1).float()

            def unfused():
                wts = torch.softmax(logits, -1).view(N, S, M, L, P)
                loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
                out = pkg.MSDeformAttnFunction.apply(value, shapes, lsi, loc, wts, 128)
                torch.autograd.grad(out, (value, off, logits), go)

            def fused():
                out = pkg.MSDeformAttnFusedFunction.apply(value, shapes, lsi, ref, off, logits)
                torch.autograd.grad(out, (value, off, logits), go)
            def unfused_fwd():
                with torch.no_grad():
                    wts = torch.softmax(logits, -1).view(N, S, M, L, P)
                    loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
                    pkg.ms_deform_attn_forward(value, shapes, lsi, loc, wts, 128)

            def fused_fwd():
                with torch.no_grad():
                    pkg.ms_deform_attn_fused_forward(value, shapes, lsi, ref, off, logits)
            ms_u = timed(unfused, args.steps * 3)
            ms_f = timed(fused, args.steps * 3)
            ms_uf = timed(unfused_fwd, args.steps * 3)
            ms_ff = timed(fused_fwd, args.steps * 3)
            emit({"config": "f", "what": "op + its producers (softmax, ref + off/(W,H)) forward+backward at configs[1] shape",
                  "unfused_ms": ms_u, "fused_ms": ms_f, "speedup": ms_u / ms_f,
                  "unfused_fwd_ms": ms_uf, "fused_fwd_ms": ms_ff})
        elif cfg == "q":
            levels = syn.pyramid(1024, 2048)
            m = encoder().eval()
            gen = torch.Generator().manual_seed(30)      # same image on every rank
            srcs = [torch.randn(1, 256, h, w, generator=gen).to(dev) for h, w in levels]
            pos = [(torch.randn(1, 256, h, w, generator=gen) * 0.1).to(dev) for h, w in levels]
            with torch.no_grad():
                if world > 1:
                    sharded = pkg.sharding.QueryShardedEncoder(m)
                    ms = timed(lambda: sharded(srcs, pos), args.steps)
                    err = (sharded(srcs, pos)[0] - m(srcs, pos)[0]).abs().max().item()
                else:
                    ms = timed(lambda: m(srcs, pos), args.steps)
                    err = 0.0
            emit({"config": "q", "what": "single-image 1024x2048 encoder inference, query-range sharded over all ranks",
                  "ms": ms, "images_per_s": 1e3 / ms, "max_abs_diff_vs_unsharded": err,
                  "all_gather_bytes_per_layer": 43008 * 256 * 4})
            del m, srcs, pos
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
