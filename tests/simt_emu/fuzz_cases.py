"""TEST INFRASTRUCTURE ONLY: random grids / periodic flags / boundary families, two time steps each, libmifgpu (through
whatever MIFGPU_LIB names -- the SIMT build on a CPU box, the real library on a GPU box) against the oracle.
usage: fuzz_cases.py [seed] [cases]      (round 1: seeds 1, 70 cases, all <= 4e-14 under the SIMT interpreter)"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ('', 'tests', 'oracle'):
    sys.path.insert(0, os.path.join(ROOT, sub))
import numpy as np
import mif_b200 as mif, mif_oracle as mo
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
def rel(a, b): return float(np.max(np.abs(a-b))) / max(float(np.max(np.abs(b))), 1e-300)
bad = 0
for case in range(n_cases):
    special = [3, 4, 5, 9, 17, 33, 65, 129, 257]
    N = [random.choice(special) if random.random() < 0.35 else random.randint(2, 40) for _ in range(3)]
    periodic = tuple(random.random() < 0.3 for _ in range(3))
    # periodic directions need >= 3 points
    N = [max(n, 4) if p else n for n, p in zip(N, periodic)]
    kind = random.choice(["ethier_steinman", "test_case_1", "test_case_2"])
    try:
        ctx = mif.Context(N[0], N[1], N[2], 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-3, 4, periodic=periodic)
        grid = mo.Grid(N[0], N[1], N[2], 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-3, 4, periodic=periodic)
        rng = np.random.default_rng(case)
        okind = {"ethier_steinman": mo.BC_ETHIER_STEINMAN, "test_case_1": mo.BC_TEST_CASE_1, "test_case_2": mo.BC_TEST_CASE_2}[kind]
        gkind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1, "test_case_2": mif.BC_TEST_CASE_2}[kind]
        h_vel = [0.3 * rng.uniform(-1, 1, grid.shape(c)) for c in range(3)]
        h_p = rng.uniform(-1, 1, grid.shape(3))
        h_buf, h_buf2, h_dp = [grid.zeros(c) for c in range(3)], [grid.zeros(c) for c in range(3)], grid.zeros(3)
        vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
        p, dp = ctx.tensor(3), ctx.tensor(3)
        for t, h in zip(vel + [p], h_vel + [h_p]): t.upload(h)
        bc = ctx.make_bc(gkind, 1e3)
        worst = 0.0
        for step in range(2):
            ctx.timestep(vel, vb, vb2, bc, step * ctx.dt, p, dp)
            grid.timestep(okind, step * ctx.dt, h_vel, h_buf, h_buf2, h_p, h_dp)
            for t, h in zip(vel + [p], h_vel + [h_p]): worst = max(worst, rel(t.download(), h))
        ctx.close()
        status = "ok" if worst <= 1e-11 else "MISMATCH"
        if status != "ok": bad += 1
        print(case, N, periodic, kind, "%.2e" % worst, status, flush=True)
    except Exception as exc:
        bad += 1
        print(case, N, periodic, kind, "EXCEPTION", repr(exc)[:300], flush=True)
print("bad:", bad)
