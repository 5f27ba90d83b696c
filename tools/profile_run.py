#!/usr/bin/env python
"""Small driver for ncu: runs chosen kernel variants a few times on one workload.

    python tools/profile_run.py --fwd 2,6 --bwd 2 [--workload cityscapes_512x1024_b8] [--iters 2]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cityscapes_512x1024_b8")
    ap.add_argument("--mode", default="model")
    ap.add_argument("--fwd", default="0")
    ap.add_argument("--bwd", default="0")
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    pkg = load_package()
    d = pkg.synthetic.make_workload_inputs(a.workload, mode=a.mode, seed=1, device="cuda:0")
    args = (d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_locations"],
            d["attention_weights"])
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    for it in range(a.iters):
        for v in [int(x) for x in a.fwd.split(",")]:
            pkg.set_option("fwd_variant", v)
            flush.zero_()
            pkg.ms_deform_attn_forward(*args, 128)
        for v in [int(x) for x in a.bwd.split(",")]:
            pkg.set_option("bwd_variant", v)
            flush.zero_()
            pkg.ms_deform_attn_backward(*args, d["grad_output"], 128)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
