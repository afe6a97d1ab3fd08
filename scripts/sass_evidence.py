#!/usr/bin/env python3
"""TMA / mbarrier evidence and the instruction mix of the Poisson sweep kernels, straight from the built library:
   python scripts/sass_evidence.py mpi-incompressible-fluid_b200/libmifgpu.so > profiles/r02_sass_sweeps.txt"""
import collections
import re
import subprocess
import sys

HEADER = """# SASS evidence for the Poisson sweep kernels of libmifgpu.so (cuobjdump -sass, sm_100a), static instruction counts per kernel
# UTMALDG / UTMASTG = cp.async.bulk.tensor loads / stores, UBLKCP = cp.async.bulk (per-rank bulk copies of the multi-GPU path),
# SYNCS.* = mbarrier arrive.expect_tx / try_wait.parity, UTMAPF = prefetch.tensormap
"""


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("mifgpu::", "")
            cur = re.sub(r"\(.*", "", name)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    print(HEADER)
    for name, ops in kernels.items():
        if not name.startswith("tmasweep::"):
            continue
        tma = {k: v for k, v in ops.items() if k.startswith(("UTMA", "UBLKCP", "SYNCS"))}
        base = collections.Counter()
        for k, v in ops.items():
            base[k.split(".")[0]] += v
        print(name)
        print("    TMA / mbarrier: " + ", ".join(f"{k} x{tma[k]}" for k in sorted(tma)))
        print("    other: " + ", ".join(f"{k} {base[k]}" for k in ("LDG", "LDS", "STS", "SHFL", "DADD", "DMUL", "DFMA", "BAR")))


if __name__ == "__main__":
    main(sys.argv[1])
