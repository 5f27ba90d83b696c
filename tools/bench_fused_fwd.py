#!/usr/bin/env python
"""Fused-producer forward at the configs[2] shape (1024x2048, batch 8): separate offsets / logits tensors,
both read in place from one [N, S, 288] projection, and the same with the per-query table added in the kernel."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
dev = "cuda:0"
levels = [(32, 64), (64, 128), (128, 256)]
N, M, L, P = 8, 8, 3, 4
S = sum(h * w for h, w in levels)
gen = torch.Generator().manual_seed(0)
value = torch.randn(N, S, M, 32, generator=gen).to(dev)
ref = pkg.modules.reference_points_for(levels, dev)
off = (torch.randn(N, S, M, L, P, 2, generator=gen) * 2).clamp_(-8, 8).to(dev)
logits = torch.randn(N, S, M, L * P, generator=gen).to(dev)
proj = torch.cat((off.reshape(N, S, -1), logits.reshape(N, S, -1)), -1).contiguous()
table = (torch.randn(S, proj.shape[-1], generator=gen) * 0.1).to(dev)
shapes, lsi = pkg.synthetic.level_tensors(levels, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
cases = {
    "separate": lambda: pkg.ms_deform_attn_fused_forward(value, shapes, lsi, ref, off, logits),
    "packed": lambda: pkg.ops.ms_deform_attn_fused_forward_packed(value, shapes, lsi, ref, proj, L, P),
    "packed+table": lambda: pkg.ops.ms_deform_attn_fused_forward_packed(value, shapes, lsi, ref, proj, L, P,
                                                                        query_table=table),
}
for name, fn in cases.items():
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    print(json.dumps({"case": name, "fused_forward_ms": sum(ts) / len(ts)}))
