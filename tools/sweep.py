#!/usr/bin/env python
"""Kernel-variant sweep on one B200: times the forward and backward kernels for every tile
shape / query order / CTA cap the library exposes through msda_b200_set_option.

    python tools/sweep.py [--out gpurun_out/sweep.jsonl] [--iters 10]

Each line: workload, location mode, option values, mean/min kernel ms (CUDA events on the
launching stream, 512 MiB L2 flush between iterations), algorithmic GB/s and fraction of the
measured HBM peak.  Used to pick the defaults in csrc/; results summarised under profiles/.
"""
import argparse
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from bench import measured_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--workloads", default="cityscapes_512x1024_b8,cityscapes_1024x2048_b1")
    ap.add_argument("--modes", default="model,uniform")
    ap.add_argument("--variants", default="1,2,3,4")
    ap.add_argument("--orders", default="0,1")
    ap.add_argument("--caps", default="0,1,2")
    args = ap.parse_args()
    pkg = load_package()
    lib, syn = pkg._lib.lib, pkg.synthetic
    peak, _ = measured_peak()
    dev = torch.device("cuda", 0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fout = open(args.out, "a")
    sp = torch.cuda.current_stream().cuda_stream
    for wl, mode in itertools.product(args.workloads.split(","), args.modes.split(",")):
        w = syn.WORKLOADS[wl]
        d = syn.make_workload_inputs(wl, mode=mode, seed=1, device=dev)
        N, S, M, D = d["value"].shape
        dims = (N, S, M, D, len(w.levels), S, w.points)
        out = torch.empty(N, S, M * D, device=dev)
        gv, gl, gw = (torch.empty_like(d[k]) for k in ("value", "sampling_locations", "attention_weights"))
        p = {k: v.data_ptr() for k, v in d.items()}

        def fwd():
            rc = lib.msda_b200_forward_f32(p["value"], p["spatial_shapes"], p["level_start_index"],
                                           p["sampling_locations"], p["attention_weights"], *dims,
                                           out.data_ptr(), sp)
            assert rc == 0, rc

        def bwd():
            rc = lib.msda_b200_backward_f32(p["grad_output"], p["value"], p["spatial_shapes"],
                                            p["level_start_index"], p["sampling_locations"],
                                            p["attention_weights"], *dims, gv.data_ptr(), gl.data_ptr(),
                                            gw.data_ptr(), sp)
            assert rc == 0, rc

        def time_it(fn, pre=None):
            ts = []
            for i in range(3 + args.iters):
                flush.zero_()
                if pre:
                    pre()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            return sum(ts) / len(ts), min(ts)

        q = N * S
        for var, order, cap in itertools.product(
                [int(x) for x in args.variants.split(",")], [int(x) for x in args.orders.split(",")],
                [int(x) for x in args.caps.split(",")]):
            pkg.set_option("fwd_variant", var)
            pkg.set_option("bwd_variant", var)
            pkg.set_option("tile_order", order)
            pkg.set_option("ctas_per_sm", cap)
            f_mean, f_min = time_it(fwd)
            b_mean, b_min = time_it(bwd, pre=gv.zero_)
            rec = {"workload": wl, "mode": mode, "variant": var, "tile_order": order, "ctas_per_sm": cap,
                   "fwd_ms": f_mean, "fwd_ms_min": f_min, "bwd_ms": b_mean, "bwd_ms_min": b_min,
                   "fwd_GBs": q * syn.FWD_BYTES_PER_QUERY / f_mean / 1e6,
                   "bwd_GBs": q * syn.BWD_BYTES_PER_QUERY / b_mean / 1e6,
                   "fwd_frac": q * syn.FWD_BYTES_PER_QUERY / f_mean / 1e6 / peak,
                   "bwd_frac": q * syn.BWD_BYTES_PER_QUERY / b_mean / 1e6 / peak,
                   "queries_per_s_fwd_bwd": q / ((f_mean + b_mean) * 1e-3)}
            fout.write(json.dumps(rec) + "\n")
            fout.flush()
            print(json.dumps(rec), flush=True)
        for k in ("fwd_variant", "bwd_variant", "tile_order", "ctas_per_sm"):
            pkg.set_option(k, 0)
        del d, out, gv, gl, gw


if __name__ == "__main__":
    main()
