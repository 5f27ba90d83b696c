#!/usr/bin/env python
"""Timing experiments on the 3xTF32 GEMM with both operands in shared memory (linear_variant 3 / the
four-accumulator kernel for in = 1024); results WRONG on purpose for the knob runs."""
# needs the profiling build: make -C uni-encoder-code_b200/csrc profile && MSDA_B200_LIB=uni-encoder-code_b200/lib/libmsda_b200_profile.so
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); lib = pkg._lib.lib; dev = "cuda:0"
pkg.set_option("linear_variant", 3)
M = 344064
for N, K in ((256, 256), (256, 1024)):
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for knob, what in ((0, "as shipped"), (4, "no output stores"), (2, "no split work"), (1, "no MMAs"),
                       (8, "no W loads"), (16, "no X loads"), (24, "no loads"), (1 + 2 + 4, "loads only"),
                       (2 + 4 + 8 + 16, "MMAs only")):
        pkg.set_option("whatif_linear", knob)
        def f(): lib.msda_b200_linear_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, 0, ws.data_ptr(), st)
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tiles_per_sm = (M / 128) * (N / 128) / 148
        print(json.dumps({"N": N, "K": K, "what": what, "ms": round(ms, 4),
                          "kcycles_per_tile": round(ms * 1e-3 * 1.92e9 / tiles_per_sm / 1e3, 2),
                          "cycles_per_kblock": round(ms * 1e-3 * 1.92e9 / tiles_per_sm / (K / 32))}), flush=True)
pkg.set_option("whatif_linear", 0)
pkg.set_option("linear_variant", 0)
