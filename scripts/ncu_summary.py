#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, without a GPU) into markdown + a traffic table bench.py can quote.

    python scripts/ncu_summary.py gpurun_out/<tag>/gtconv_full.ncu-rep profiles/r01/ncu_gtconv  ->  .md and .json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, out_prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    md = [f"# ncu summary of `{rep}`", "", "Captured with `ncu --set full --clock-control none --import-source on` (replayed, cold-cache: "
          "durations are for shares, not absolutes; bench.py times the same kernels with CUDA events).", ""]
    traffic = {}
    for r in body:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("ab2::", "")
        md += [f"## `{short}`", "", "| metric | value |", "|---|---|"]
        for key, label in KEYS:
            if key in idx:
                md.append(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * TO_BYTES[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * TO_BYTES[units[idx["dram__bytes_write.sum"]]]
        md += [f"| **DRAM traffic per launch** | {(rd + wr) / 1e9:.3f} GB |", ""]
        base = short.split("<")[0]
        traffic[base] = {"dram_bytes_per_launch": rd + wr, "kernel": short}
    open(out_prefix + ".md", "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(out_prefix + ".json", "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
