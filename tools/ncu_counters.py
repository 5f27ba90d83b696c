#!/usr/bin/env python
"""Per-launch counters of the bench command's kernels from an ncu --set full report, as the small JSON
that bench.py reads for `roofline.traffic` and the on-chip limiter (nothing in bench.py is pasted by hand).

    python tools/ncu_counters.py gpurun_out/prof_bench.ncu-rep profiles/r2_ncu_bench_counters.json \
        [--workload cityscapes_512x1024_b8] [--mode model]

Needs `ncu` (build container); averages over the captured launches of each kernel."""
import argparse
import csv
import io
import json
import subprocess

METRICS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum": "red_sectors",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--workload", default="cityscapes_512x1024_b8")
    ap.add_argument("--mode", default="model")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for d in data:
        name = d[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1].strip()
        k = acc.setdefault(name, {"launches": 0})
        k["launches"] += 1
        for m, key in METRICS.items():
            if m in ix and d[ix[m]] not in ("", "n/a"):
                v = float(d[ix[m]].replace(",", "")) * UNIT_SCALE.get(units[ix[m]], 1.0)
                k[key] = k.get(key, 0.0) + v
    for k in acc.values():
        n = k["launches"]
        for key in list(k):
            if key != "launches":
                k[key] = k[key] / n
        if "dram_read_bytes" in k:
            k["dram_bytes"] = k["dram_read_bytes"] + k.get("dram_write_bytes", 0.0)
    doc = {"workload": a.workload, "mode": a.mode, "source": a.report,
           "how": "ncu --set full --clock-control none of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline`; "
                  "per launch, averaged over the captured launches", "kernels": acc}
    with open(a.out, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    for n, k in acc.items():
        print(n, {x: round(y, 2) for x, y in k.items()})


if __name__ == "__main__":
    main()
