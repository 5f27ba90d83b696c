#!/usr/bin/env python
"""What-if timing for profiles/: the backward kernel with the grad_value reductions of the coarse
levels dropped (results are WRONG on purpose) -- an upper bound for any scheme that merges those
contributions before they reach L2."""
# needs the profiling build: make -C uni-encoder-code_b200/csrc profile && MSDA_B200_LIB=uni-encoder-code_b200/lib/libmsda_b200_profile.so
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
d = pkg.synthetic.make_workload_inputs("cityscapes_512x1024_b8", mode="model", seed=1, device="cuda:0")
a = (d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_locations"], d["attention_weights"])
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
for drop, what in ((0, "all reductions (shipped kernel)"), (2, "level 0 dropped (1/3 of the rows)"),
                   (4, "levels 0+1 dropped (2/3)"), (6, "all dropped (gather + small gradients only)")):
    pkg.set_option("whatif_drop_reds", drop)
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.ms_deform_attn_backward(*a, d["grad_output"], 128); e1.record()
        torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    print(json.dumps({"dropped_pairs": drop, "what": what, "bwd_ms_incl_memset": sum(ts) / len(ts)}), flush=True)
pkg.set_option("whatif_drop_reds", 0)
