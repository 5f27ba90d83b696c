#!/usr/bin/env python
"""Single-image latency of the pixel-decoder forward (the per-frame case of the KITTI / Cityscapes demos):
eager launches against one CUDA graph of the whole forward_features call (no host reads of device data
remain in the mirror, so it captures), for the reference-equivalent settings and with the round-1 kernels."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); dev = torch.device("cuda", 0)
shapes = {"res2": (96, 4), "res3": (192, 8), "res4": (384, 16), "res5": (768, 32)}
for (H, W), name in (((384, 1248), "KITTI 384x1248"), ((512, 1024), "Cityscapes crop 512x1024"), ((1024, 2048), "Cityscapes 1024x2048")):
    for linear, fused in (("torch", False), ("tf32x3", True)):
        torch.manual_seed(0)
        dec = pkg.pixel_decoder.MSDeformAttnPixelDecoder(
            shapes, transformer_dropout=0.1, transformer_nheads=8, transformer_dim_feedforward=1024,
            transformer_enc_layers=6, conv_dim=256, mask_dim=256, norm="GN",
            transformer_in_features=["res3", "res4", "res5"], common_stride=4, fused=fused, linear=linear).to(dev).eval()
        gen = torch.Generator().manual_seed(3)
        feats = {k: torch.randn(1, c, -(-H // st), -(-W // st), generator=gen).to(dev) for k, (c, st) in shapes.items()}
        with torch.no_grad():
            for _ in range(3):
                ref = dec.forward_features(feats)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                dec.forward_features(feats)
            torch.cuda.synchronize()
            eager_ms = (time.perf_counter() - t0) / 20 * 1e3
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                dec.forward_features(feats)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(graph):
                out = dec.forward_features(feats)
            graph.replay(); torch.cuda.synchronize()
            same = bool(torch.equal(out[0], ref[0]))
            t0 = time.perf_counter()
            for _ in range(20):
                graph.replay()
            torch.cuda.synchronize()
            graph_ms = (time.perf_counter() - t0) / 20 * 1e3
        print(json.dumps({"image": name, "linear": linear, "fused": fused, "eager_ms": round(eager_ms, 3),
                          "cuda_graph_ms": round(graph_ms, 3), "graph_output_identical": same}), flush=True)
        del dec, graph, out, ref
