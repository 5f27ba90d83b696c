#!/usr/bin/env python
"""Markdown table from a tools/sweep.py jsonl: python tools/sweep_table.py sweep.jsonl"""
import json
import sys

PEAK = 6540.8
QUERIES = {"cityscapes_512x1024_b8": 86016, "cityscapes_1024x2048_b1": 43008, "cityscapes_1024x2048_b8": 344064,
           "kitti_384x1248_b16": 157248}
print("| workload | locations | bwd_variant | forward ms | backward ms | backward % of HBM |")
print("|---|---|---|---:|---:|---:|")
for line in open(sys.argv[1]):
    r = json.loads(line)
    q = QUERIES.get(r["workload"], 0)
    pct = f"{q * 5376 / r['bwd_ms'] / 1e6 / PEAK * 100:.1f}" if q else ""
    print(f"| {r['workload']} | {r['mode']} | {r['variant']} | {r['fwd_ms']:.3f} | {r['bwd_ms']:.3f} | {pct} |")
