#!/usr/bin/env python
"""HBM-bound helper kernels against the measured copy bandwidth: fused residual + LayerNorm (forward, backward),
transpose; torch's own kernels beside them.  Rows = BASELINE configs[2] (344 064 x 256)."""
import json, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
from bench import measured_peak
pkg = load_package(); dev = "cuda:0"
peak, _ = measured_peak()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
rows, cols = 344064, 256
x = torch.randn(rows, cols, device=dev); r = torch.randn(rows, cols, device=dev); gy = torch.randn(rows, cols, device=dev)
w = torch.randn(cols, device=dev); b = torch.randn(cols, device=dev)
MB = rows * cols * 4 / 1e6
def line(name, ms, passes, torch_ms):
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(passes * MB, 1),
                      "GBs": round(passes * MB / ms, 1), "frac_of_measured_hbm": round(passes * MB / ms / peak, 3),
                      "torch_ms": round(torch_ms, 4)}), flush=True)
line("add_layernorm forward (read x, residual; write y)", t(lambda: pkg.add_layernorm(x, r, w, b, 1e-5)), 3,
     t(lambda: F.layer_norm(x + r, (cols,), w, b, 1e-5)))
xg = x.clone().requires_grad_(True); rg = r.clone().requires_grad_(True); wg = w.clone().requires_grad_(True); bg = b.clone().requires_grad_(True)
y = F.layer_norm(xg + rg, (cols,), wg, bg, 1e-5)
line("add_layernorm backward (read grad_y, x, residual; write grad_v)", t(lambda: pkg.ops.add_layernorm_backward(gy, x, r, w, 1e-5)), 4,
     t(lambda: torch.autograd.grad(y, (xg, rg, wg, bg), gy, retain_graph=True)))
line("transpose (read x; write x^T)", t(lambda: pkg.ops.transpose2d(x)), 2, t(lambda: x.t().contiguous()))
