#!/usr/bin/env python
"""Where the warp roles of the 3xTF32 GEMM (A operand in tensor memory) spend their cycles: per k-block / per tile
averages from clock64 counters.  Needs the profiling build:
make -C uni-encoder-code_b200/csrc profile && MSDA_B200_LIB=uni-encoder-code_b200/lib/libmsda_b200_profile.so"""
import ctypes, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); lib = pkg._lib.lib; dev = "cuda:0"
raw = ctypes.CDLL(pkg._lib.LIB_PATH)
raw.msda_b200_debug_linear_counters.argtypes = [ctypes.c_void_p]
M = 344064
knob = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # whatif_linear bits (tools/whatif_linear_atmem.py)
pkg.set_option("whatif_linear", knob)
for N, K in ((256, 256), (1024, 256)):
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def f(): lib.msda_b200_linear_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, 0, ws.data_ptr(), st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    raw.msda_b200_debug_linear_counters(buf)          # clear
    f(); torch.cuda.synchronize()
    raw.msda_b200_debug_linear_counters(buf)
    c = list(buf)
    kblocks = c[1]; tiles = kblocks / (K / 32); ctas = 148
    per_kb = lambda v: round(v / kblocks)
    per_tile = lambda v: round(v / tiles)
    print(json.dumps({"N": N, "K": K, "cycles_per_cta": round(c[11] / ctas), "per_kblock_wall": round(c[11] / ctas / (kblocks / ctas)),
                      "producer_wait_empty_per_kb": per_kb(c[0]), "commit_to_producer_seen_per_kb": per_kb(c[12]),
                      "tma_issue_to_full_seen_per_kb": per_kb(c[4]), "split_wait_full_per_kb": per_kb(c[2]),
                      "split_work_per_kb": per_kb(c[3]), "mma_wait_ready_per_kb": per_kb(c[5]),
                      "mma_issue_per_kb": per_kb(c[6]), "mma_wait_acc_empty_per_tile": per_tile(c[7]),
                      "epi_wait_acc_full_per_tile": per_tile(c[8]), "epi_drain_per_tile": per_tile(c[9]),
                      "epi_store_phase_per_tile": per_tile(c[10])}), flush=True)
