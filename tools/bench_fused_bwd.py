#!/usr/bin/env python
"""Fused-producer backward at the configs[1] shape: probe-gated default vs merging vs per-row kernel."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
dev = "cuda:0"
levels = [(16, 32), (32, 64), (64, 128)]
N, M, L, P = 8, 8, 3, 4
S = sum(h * w for h, w in levels)
gen = torch.Generator().manual_seed(0)
value = torch.randn(N, S, M, 32, generator=gen).to(dev)
ref = pkg.modules.reference_points_for(levels, dev)
off = (torch.randn(N, S, M, L, P, 2, generator=gen) * 2).clamp_(-8, 8).to(dev)
logits = torch.randn(N, S, M, L * P, generator=gen).to(dev)
go = torch.randn(N, S, M * 32, generator=gen).to(dev)
shapes, lsi = pkg.synthetic.level_tensors(levels, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for var in (0, 20, 2):
    pkg.set_option("bwd_variant", var)
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pkg.ms_deform_attn_fused_backward(value, shapes, lsi, ref, off, logits, go)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    print(json.dumps({"bwd_variant": var, "fused_backward_ms_incl_zero_fill": sum(ts) / len(ts)}))
pkg.set_option("bwd_variant", 0)
