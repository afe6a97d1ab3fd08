#!/usr/bin/env python3
"""Per-kernel milliseconds of the full step without torch (numpy + ctypes only, so it starts in a second): the library's
own CUDA-event profile (mifgpu_profile_*) over `steps` steps after 3 warm-up steps, plus the wall-clock time per step
around a stream synchronise.  Used for A/B runs of the switches (one process per setting: they are read once).
usage: ab_timing.py [points per direction = 513] [steps = 5] [label]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mif_b200 as mif  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 513
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
label = sys.argv[3] if len(sys.argv) > 3 else ""
dt = 1e-3
total = 3 + 2 * steps
ctx = mif.Context(N, N, N, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, dt * total, total)
vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
bc = ctx.make_bc(mif.BC_TEST_CASE_1, 1e3)
sx, sy, sz = vel[1].shape
v0 = np.zeros((sz, sy, sx))
v0[:, :, N - 1] = 1.0
vel[1].upload(v0)
del v0
step = 0
for _ in range(3):
    ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp)
    step += 1
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp)
    step += 1
ctx.synchronize()
wall_ms = (time.perf_counter() - t0) * 1e3 / steps
ctx.profile_enable(True)
for _ in range(steps):
    ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp)
    step += 1
prof = ctx.profile_read()
kernels = {k: round(ms / steps, 4) for k, (ms, n) in prof.items() if n}
print(json.dumps({"label": label, "points": N, "steps": steps, "wall_ms_per_step": round(wall_ms, 3),
                  "profile_ms_per_step": round(sum(kernels.values()), 3), "kernels": kernels,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("MIFGPU_")}}), flush=True)
ctx.close()
