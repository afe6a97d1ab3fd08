#!/usr/bin/env python3
"""Reads a phase trace written by the diagnostic build of libmifgpu (make -C mpi-incompressible-fluid_b200 trace;
MIFGPU_LIB=.../build/libmifgpu_trace.so MIFGPU_TRACE_FILE=<file> [MIFGPU_TRACE_MODE=0|1|2]): SM-clock stamps of the 8 warps
of the two persistent CTAs with blockIdx 0 and 148 of tma_dct512_kernel at the boundaries of the phases of every tile.
Prints the average length of every phase and how much of the time the two CTAs spend in the same kind of phase.
usage: phase_trace.py <file>"""
import sys

import numpy as np

TILES, SLOTS = 24, 16
NAMES = ["wait for tile (mbarrier)", "stage reads (shared loads)", "phase A butterflies (FP64)", "barrier 1 (+ wait for bulk stores)",
         "phase A stores (shared stores)", "barrier 2", "phase B: loads, butterflies, lane exchange", "unpack (shuffles + FP64)",
         "scale / repack / second transform (fused z only)", "barrier 3", "output-stage stores", "fence + barrier 4"]
KIND = ["wait", "lsu", "fp64", "wait", "lsu", "wait", "mixed", "mixed", "mixed", "wait", "lsu", "wait"]

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
stamps = raw[:2 * 8 * TILES * SLOTS].reshape(2, 8, TILES, SLOTS).astype(np.int64)
smid = raw[2 * 8 * TILES * SLOTS:]
print("SM of CTA 0 / CTA 148:", int(smid[0]), int(smid[1]))
valid = stamps[:, :, 2:TILES - 1, :13]  # skip the first tiles (start-up)
dur = np.diff(valid, axis=3)            # (cta, warp, tile, 12)
print("cycles per tile and CTA: %.0f" % np.mean(valid[:, :, 1:, 0] - valid[:, :, :-1, 0]))
for i, name in enumerate(NAMES):
    print("  %-52s %7.0f cycles  (%s)" % (name, dur[..., i].mean(), KIND[i]))
# timeline overlap between the two CTAs (warp 0 of each as representative), if they share an SM
if smid[0] == smid[1]:
    t0, t1 = valid[0, 0], valid[1, 0]
    lo, hi = max(t0[0, 0], t1[0, 0]), min(t0[-1, 12], t1[-1, 12])

    def kind_at(t, when):
        flat = t.reshape(-1)
        idx = np.searchsorted(flat, when, side="right") - 1
        slot = idx % 13
        return np.where(slot < 12, np.array(KIND + ["wait"])[np.minimum(slot, 12)], "wait")

    when = np.linspace(lo, hi, 20000).astype(np.int64)
    a, b = kind_at(t0, when), kind_at(t1, when)
    for ka in ("fp64", "lsu", "mixed", "wait"):
        print("  CTA 0 in %-5s: CTA 148 in " % ka + ", ".join("%s %.0f %%" % (kb, 100 * np.mean(b[a == ka] == kb)) for kb in ("fp64", "lsu", "mixed", "wait")))
