#!/usr/bin/env python
"""One shape of msda_b200_linear_f32 a few times (for an ncu capture): python tools/prof_linear_one.py N K"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); lib = pkg._lib.lib; dev = "cuda:0"
N, K = int(sys.argv[1]), int(sys.argv[2])
M = 344064
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
y = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    lib.msda_b200_linear_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, 0, ws.data_ptr(), st)
torch.cuda.synchronize()
