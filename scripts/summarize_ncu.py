#!/usr/bin/env python3
"""Turns the captures scripts/gpu_session.sh brought back into the files kept under profiles/ (run here, no GPU):

  python scripts/summarize_ncu.py gpurun_out/r02 profiles/r02

  <in>_launches.csv  -> <out>_ncu_launch_summary.csv   per kernel: launches, total ms, share of the window
  <in>_step.ncu-rep  -> <out>_ncu_step_kernels.json    per launch: duration, DRAM bytes read / written, DRAM / L1 / FP64
                                                       pipe utilisation, registers, shared-memory wavefronts and
                                                       bank conflicts (ncu --page raw)
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict


def launch_summary(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            value = float(row["Metric Value"].replace(",", ""))
            unit = row.get("Metric Unit", "ns")
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            rows.append((row["Kernel Name"], value * scale))
    total = sum(ms for _, ms in rows)
    per = OrderedDict()
    for name, ms in rows:
        short = name.split("(")[0].split("::")[-1]
        n, t = per.get(short, (0, 0.0))
        per[short] = (n + 1, t + ms)
    with open(out, "w") as f:
        f.write(f"# total device time in window: {total:.3f} ms over {len(rows)} launches\nkernel,launches,total_ms,share\n")
        for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name},{n},{t:.3f},{t / total:.4f}\n")
    print(open(out).read())


METRICS = {
    "gpu__time_duration.sum": "ns", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
}


def step_kernels(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in raw.splitlines() if not l.startswith("==")]
    reader = csv.reader(io.StringIO("\n".join(lines)))
    header = next(reader)
    units = next(reader)
    col = {name: i for i, name in enumerate(header)}
    result = []
    for row in reader:
        entry = {"kernel": row[col["Kernel Name"]].split("(")[0].split("::")[-1]}
        for metric, key in METRICS.items():
            if metric in col:
                try:
                    entry[key] = float(row[col[metric]].replace(",", ""))
                    entry[key + "_unit"] = units[col[metric]]
                except ValueError:
                    pass
        result.append(entry)
    with open(out, "w") as f:
        json.dump(result, f, indent=1)
    print(f"{len(result)} launches -> {out}")


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    launch_summary(src + "_launches.csv", dst + "_ncu_launch_summary.csv")
    step_kernels(src + "_step.ncu-rep", dst + "_ncu_step_kernels.json")
