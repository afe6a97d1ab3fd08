#!/usr/bin/env python3
"""A GPU check that fits into a minute: smoke() (16^3, generic / Bluestein sweeps, stage / bc / correct kernels), then the
parity cases of the kernels written without a GPU, most important first; every line is flushed so that a cut-off run
still tells how far it got."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.time()


def say(text):
    print(f"[{time.time() - t0:6.1f}s] {text}", flush=True)


import __graft_entry__ as entry  # noqa: E402

say("imports done")
entry.smoke()
say("smoke ok")
import mif_b200 as mif  # noqa: E402
import test_gpu_vs_oracle as parity  # noqa: E402

F, T = False, True
for N, periodic in [((20, 13, 11), (F, F, F)), ((2049, 3, 4), (F, F, F)), ((5, 3, 2049), (F, F, F)), ((3, 2049, 4), (F, F, F)),
                    ((33, 6, 481), (F, F, T)), ((513, 4, 9), (F, F, F)), ((9, 513, 3), (F, F, F)), ((11, 3, 513), (F, F, F)), ((5, 257, 6), (F, F, F)),
                    ((6, 5, 257), (F, F, F)), ((9, 1025, 3), (F, F, F)), ((10, 4, 1025), (F, F, F)), ((3000, 3, 2), (F, F, F))]:
    parity.test_pressure_solve_random_velocity(mif, N, periodic)
    say(f"solve parity ok {N} {periodic}")
say("done")
