#!/usr/bin/env python
"""Quick GPU check of msda_b200_linear_f32 (3xTF32 tcgen05 GEMM) against fp64 and torch fp32."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
lib = pkg._lib.lib
dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
def run(M, N, K, relu=0, bias=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    y = torch.full((M, N), float("nan"), device=dev)
    ws = torch.empty(2 * N * K, device=dev)
    rc = lib.msda_b200_linear_f32(x.data_ptr(), w.data_ptr(), b.data_ptr() if bias else None, y.data_ptr(),
                                  M, N, K, relu, ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref64 = torch.nn.functional.linear(x.double(), w.double(), b.double() if bias else None)
    ref32 = torch.nn.functional.linear(x, w, b)
    if relu:
        ref64, ref32 = ref64.relu(), ref32.relu()
    err = (y.double() - ref64).abs().max().item()
    err32 = (ref32.double() - ref64).abs().max().item()
    print(json.dumps({"M": M, "N": N, "K": K, "relu": relu, "rc": rc, "max_err_vs_fp64": err,
                      "torch_fp32_err_vs_fp64": err32, "nan": bool(torch.isnan(y).any().item())}), flush=True)
for shape in [(128, 256, 32), (128, 256, 256), (300, 256, 256), (1000, 192, 256), (777, 96, 256), (512, 1024, 256),
              (640, 256, 1024), (130, 128, 64)]:
    run(*shape)
run(1000, 256, 256, relu=1)
pkg.set_option("linear_variant", 2)
run(1000, 256, 256); run(640, 256, 1024); run(777, 96, 256)
pkg.set_option("linear_variant", 4)
print("variant 4 (A in TMEM):")
run(128, 256, 32); run(1000, 256, 256); run(640, 256, 1024); run(777, 96, 256); run(513, 1024, 256)
pkg.set_option("linear_variant", 0)
run(1000, 256, 256, bias=False)
# timing at the pixel-decoder shape
M = 344064
import itertools
for variant, (N, K) in itertools.product((0, 2), ((256, 256), (192, 256), (96, 256), (1024, 256), (256, 1024))):
    pkg.set_option("linear_variant", variant)
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev)
    ws = torch.empty(2 * N * K, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def mine(): lib.msda_b200_linear_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, 0, ws.data_ptr(), st)
    def ref(): torch.nn.functional.linear(x, w, b)
    res = {}
    for name, fn in (("tf32x3", mine),) + ((("torch_fp32", ref),) if variant == 0 else ()):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        res[name + "_ms"] = e0.elapsed_time(e1) / 10
    res.update(variant=variant, M=M, N=N, K=K, tflops_fp32_equiv=2 * M * N * K / res["tf32x3_ms"] / 1e9)
    print(json.dumps(res), flush=True)
pkg.set_option("linear_variant", 0)
