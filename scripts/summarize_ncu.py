#!/usr/bin/env python3
"""Turns an `ncu --set full` capture brought back from a GPU session into the summaries kept under profiles/ (run here,
no GPU needed):

  python scripts/summarize_ncu.py gpurun_out/<capture>.ncu-rep profiles/<name>.json [profiles/<name>_stalls.md]

  <name>.json        one entry per captured launch: duration, DRAM bytes read / written (bench.py's roofline.traffic
                     is read from this file), DRAM / LSU / FP64 pipe utilisation, issue-slot utilisation, resident
                     warps, registers, shared-memory wavefronts and bank conflicts (ncu --page raw)
  <name>_stalls.md   per kernel: warp-stall reasons and the instruction mix per line of the sweep (ncu --page source)
"""
import collections
import csv
import io
import json
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": ("duration", None),
    "dram__bytes_read.sum": ("dram_read", None),
    "dram__bytes_write.sum": ("dram_write", None),
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": ("dram_pct", 1.0),
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": ("lsu_wavefront_pct", 1.0),
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": ("smem_wavefronts", 1.0),
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": ("smem_bank_conflict_wavefronts", 1.0),
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": ("fp64_pipe_pct", 1.0),
    "smsp__issue_active.avg.pct_of_peak_sustained_active": ("issue_active_pct", 1.0),
    "sm__warps_active.avg.pct_of_peak_sustained_active": ("warps_active_pct", 1.0),
    "launch__registers_per_thread": ("regs", 1.0),
    "launch__grid_size": ("grid", 1.0),
    "launch__block_size": ("block", 1.0),
    "sm__cycles_elapsed.max": ("sm_cycles", 1.0),
}
SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6,
         "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}


def ncu_page(rep, page):
    raw = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO("\n".join(l for l in raw.splitlines() if not l.startswith("==")))))


def short(name):
    name = name.replace("(int)", "").replace("(bool)", "")
    return name.split("(")[0].replace("void ", "").replace("mifgpu::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")


def launches(rep, out):
    rows = ncu_page(rep, "raw")
    header, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(header)}
    result = []
    for row in rows[2:]:
        entry = {"kernel": short(row[col["Kernel Name"]])}
        for metric, (key, plain) in METRICS.items():
            if metric not in col:
                continue
            try:
                value = float(row[col[metric]].replace(",", ""))
            except ValueError:
                continue
            unit = units[col[metric]]
            if key == "duration":
                entry["duration_ms"] = round(value * SCALE.get(unit, 1e-6), 6)
            elif key in ("dram_read", "dram_write"):
                entry[key + "_GB"] = round(value * SCALE.get(unit, 1e-9), 6)
            else:
                entry[key] = round(value, 3)
        result.append(entry)
    with open(out, "w") as f:
        json.dump(result, f, indent=1)
    print(f"{len(result)} launches -> {out}")
    return result


def stalls(rep, out, lines_per_launch):
    rows = ncu_page(rep, "source")
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": short(r[1]), "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    seen, text = set(), ["# Warp stalls and instruction mix per kernel (ncu --page source of %s)\n" % rep.split("/")[-1]]
    for k in kernels:
        if k["name"] in seen or "hdr" not in k:
            continue
        seen.add(k["name"])
        h = {n: i for i, n in enumerate(k["hdr"])}
        total = sum(int(r[h["# Samples"]]) for r in k["rows"]) or 1
        agg = collections.Counter()
        for r in k["rows"]:
            for n in k["hdr"]:
                if n.startswith("stall_") and "Not Issued" not in n:
                    agg[n[6:]] += int(r[h[n]])
        execs, wf = collections.Counter(), collections.Counter()
        for r in k["rows"]:
            parts = r[h["Source"]].split()
            op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
            execs[op] += int(r[h["Instructions Executed"]])
            wf[op] += int(r[h["L1 Wavefronts Shared"]])
        text.append(f"## {k['name']}\n")
        text.append("stall samples: " + ", ".join(f"{n} {100 * v / total:.1f} %" for n, v in agg.most_common(8)) + "\n")
        if lines_per_launch:
            text.append(f"warp instructions per line ({lines_per_launch} lines per launch): " +
                        ", ".join(f"{op} {execs[op] / lines_per_launch:.0f}" for op, _ in execs.most_common(12)) + "\n")
            text.append("shared-memory wavefronts per line: " + ", ".join(f"{op} {wf[op] / lines_per_launch:.0f}" for op in wf if wf[op]) + "\n")
    with open(out, "w") as f:
        f.write("\n".join(text))
    print(f"{len(seen)} kernels -> {out}")


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:
        stalls(sys.argv[1], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 513 * 513)
