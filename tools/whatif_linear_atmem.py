#!/usr/bin/env python
"""Timing experiments on the shipped 3xTF32 GEMM (A operand in tensor memory); results WRONG on purpose for the
knob runs.  Needs the profiling build:
make -C uni-encoder-code_b200/csrc profile && MSDA_B200_LIB=uni-encoder-code_b200/lib/libmsda_b200_profile.so"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); lib = pkg._lib.lib; dev = "cuda:0"
M = 344064
for N, K in ((256, 256),):
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for knob, what in ((0, "as shipped"), (512, "no tcgen05.st (MMAs read stale A)"), (512 + 1024, "no tcgen05.st, no shared-memory loads in the split warps"),
                       (512 + 1024 + 2, "split warps: wait and arrive only"), (4, "no output stores"), (2, "no split arithmetic"), (1, "no x_lo MMA"),
                       (8, "no x_hi MMA"), (16, "x_hi MMA single width"), (1 + 16, "one single-width MMA per k-step"),
                       (32, "x_lo MMA into main (no shared accumulator half)"), (1 + 8, "no MMAs"),
                       (1 + 8 + 4, "no MMAs, no stores"), (2 + 4, "no split arithmetic, no stores"),
                       (256, "no X loads"), (128 + 256, "no loads (barrier traffic only)"),
                       (1 + 8 + 4 + 128 + 256, "no loads, no MMAs, no stores"),
                       (1 + 8 + 4 + 512, "no MMAs, no stores, no tcgen05.st"),
                       (1 + 8 + 4 + 512 + 1024, "no MMAs, no stores, no tcgen05.st, no LDS"),
                       (1 + 8 + 4 + 128 + 256 + 512 + 1024, "barriers only"),
                       (1 + 8 + 4 + 128 + 256 + 512 + 1024 + 4096, "barriers only, no epilogue transposes"),
                       (1 + 8 + 4 + 128 + 256 + 512 + 1024 + 2048, "barriers only, plain arrives instead of tcgen05.commit"),
                       (1 + 8 + 4 + 128 + 256 + 512 + 1024 + 2048 + 4096, "barriers only, plain arrives, no epilogue transposes"),
                       (4096, "no epilogue transposes / stores"),
                       (8192, "mbarrier.test_wait polling instead of try_wait"),
                       (8192 + 1 + 8 + 4 + 128 + 256 + 512 + 1024, "barriers only, polling"),
                       (64, "no W_lo loads"), (128, "no W loads"), (128 + 4, "no W loads, no stores"),
                       (128 + 1 + 8 + 4, "no W loads, no MMAs, no stores")):
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
                          "kcycles_per_tile": round(ms * 1e-3 * 1.965e9 / tiles_per_sm / 1e3, 2),
                          "cycles_per_kblock": round(ms * 1e-3 * 1.965e9 / tiles_per_sm / (K / 32))}), flush=True)
pkg.set_option("whatif_linear", 0)
