#!/usr/bin/env python
"""Same-GPU comparison with the reference's own CUDA op (baseline/_ref/ref_msda_cuda.so, built by
baseline/build_reference_cuda.py): parity on identical inputs and kernel times, forward and backward.

    python tools/bench_vs_reference_cuda.py [--out gpurun_out/vs_reference_cuda.jsonl]
"""
import argparse
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

REF_SO = os.path.join(ROOT, "baseline", "_ref", "ref_msda_cuda.so")


def load_reference_cuda():
    if not os.path.exists(REF_SO):
        return None
    spec = importlib.util.spec_from_file_location("ref_msda_cuda", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "vs_reference_cuda.jsonl"))
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    ref = load_reference_cuda()
    if ref is None:
        print(json.dumps({"unavailable": "baseline/_ref/ref_msda_cuda.so not built"}))
        return
    pkg = load_package()
    dev = torch.device("cuda", 0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)

    def time_it(fn):
        ts = []
        for i in range(3 + args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    with open(args.out, "a") as fout:
        for wl in ("cityscapes_512x1024_b8", "cityscapes_1024x2048_b1", "kitti_384x1248_b16"):
            for mode in ("model", "uniform"):
                d = pkg.synthetic.make_workload_inputs(wl, mode=mode, seed=1, device=dev)
                a = (d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_locations"],
                     d["attention_weights"])
                o_ref = ref.ms_deform_attn_forward(*a, 128)
                o_new = pkg.ms_deform_attn_forward(*a, 128)
                g_ref = ref.ms_deform_attn_backward(*a, d["grad_output"], 128)
                g_new = pkg.ms_deform_attn_backward(*a, d["grad_output"], 128)
                rel = lambda x, y: ((x - y).abs().max() / y.abs().max()).item()
                rec = {
                    "workload": wl, "mode": mode,
                    "fwd_max_abs_diff": (o_ref - o_new).abs().max().item(),
                    "grad_value_rel": rel(g_new[0], g_ref[0]), "grad_loc_rel": rel(g_new[1], g_ref[1]),
                    "grad_weight_rel": rel(g_new[2], g_ref[2]),
                    # includes each side's own output allocation / zero-fill, as a caller sees it
                    "ref_fwd_ms": time_it(lambda: ref.ms_deform_attn_forward(*a, 128)),
                    "new_fwd_ms": time_it(lambda: pkg.ms_deform_attn_forward(*a, 128)),
                    "ref_bwd_ms": time_it(lambda: ref.ms_deform_attn_backward(*a, d["grad_output"], 128)),
                    "new_bwd_ms": time_it(lambda: pkg.ms_deform_attn_backward(*a, d["grad_output"], 128)),
                }
                rec["fwd_speedup"] = rec["ref_fwd_ms"] / rec["new_fwd_ms"]
                rec["bwd_speedup"] = rec["ref_bwd_ms"] / rec["new_bwd_ms"]
                rec["fwd_bwd_speedup"] = (rec["ref_fwd_ms"] + rec["ref_bwd_ms"]) / (rec["new_fwd_ms"] + rec["new_bwd_ms"])
                fout.write(json.dumps(rec) + "\n")
                print(json.dumps(rec), flush=True)
                del d, a, o_ref, o_new, g_ref, g_new


if __name__ == "__main__":
    main()
