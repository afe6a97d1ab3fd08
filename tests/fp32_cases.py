"""TEST INFRASTRUCTURE ONLY: parity cases of the single-precision build of the library (libmifgpu_f32.so, the
reference's USE_DOUBLE=0 build, include/Real.h:9-17).  Run in a subprocess with MIFGPU_LIB pointing at the float library
(a process binds one library): tests/test_fp32_cpu_simt.py uses the SIMT interpreter's build, tests/test_gpu_zzz_fp32.py
the sm_100a one.  Prints one JSON line {case: max relative L-inf error}.

Two checkers, same cases for both libraries:
  * the reference itself compiled with USE_DOUBLE=0 (oracle/_ref/f32, fields in tests/golden/f32_*.npz made by
    `oracle/make_golden.py --f32`): two float implementations with different operation order;
  * the FP64 oracle (oracle/mif_oracle.c) on inputs that are exactly representable in float.
Tolerances (written here, asserted by the callers), relative L-infinity per step:
  TOL32_VEL = 2e-5 for u, v, w and for a stand-alone pressure solve (observed 2e-7 ... 1.3e-6: a few float ulps);
  TOL32_P   = 1e-3 for the pressure of a time step.  Float storage alone sets its floor: the rounding of u* (6e-8) is
              amplified by 1 / (dx dt_s) in rhs = div(u*) / dt_s, e.g. 5e-5 of max|p| for the Ethier-Steinman golden
              (dt = 5e-5, 16^3), which is also the distance between the float and the double build of the REFERENCE
              on that case (3.5e-5 / 5.3e-5 after one / two steps); the library is at 2e-4 there, 1e-6 at dt = 2.5e-4.
The FP64 contract (1e-11) is untouched by this build.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

TOL32_VEL = 2e-5
TOL32_P = 1e-3


def rel(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), floor, 1e-300)


def golden_timestep(mif, case):
    """Replays a dumped run of the float reference (same set-up as tests/test_gpu_golden.py)."""
    from conftest import load_golden
    meta, f = load_golden(case)
    N = meta["N"]
    ctx = mif.Context(N[0], N[1], N[2], meta["x_size"], meta["y_size"], meta["z_size"], *meta["min"], meta["Re"],
                      meta["final_time"], meta["steps"], periodic=[bool(p) for p in meta["periodic"]])
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, name in zip(vel, "uvw"):
        assert f[name + "_s0"].dtype == np.float32
        t.upload(f[name + "_s0"])
    p.upload(f["p_s0"])
    kind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1,
            "test_case_2": mif.BC_TEST_CASE_2}[meta["bc"]]
    bc = ctx.make_bc(kind, meta["Re"])
    dt = meta["final_time"] / meta["steps"]
    worst = {"vel": 0.0, "p": 0.0}
    for step in range(meta["steps"]):
        ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp)
        vmax = max(float(np.max(np.abs(f[f"{c}_s{step + 1}"]))) for c in "uvw")
        for t, name in zip(vel + [p], "uvwp"):
            ref = f[f"{name}_s{step + 1}"]
            got = t.download()
            assert got.dtype == np.float32
            # a component that is zero in exact arithmetic (w of the z-periodic lid case) is pure round-off
            key = "p" if name == "p" else "vel"
            worst[key] = max(worst[key], rel(got, ref, floor=1e-3 * vmax if name in "uvw" else 0.0))
    ctx.close()
    return worst


def oracle_pair(mif, mo, N, periodic):
    size, lo = (1.0, 1.0, 2.0), (0.0, 0.0, -1.0)
    return (mif.Context(N[0], N[1], N[2], *size, *lo, 1e3, 1e-3, 4, periodic=periodic),
            mo.Grid(N[0], N[1], N[2], *size, *lo, 1e3, 1e-3, 4, periodic=periodic))


def as_float(a):
    return a.astype(np.float32).astype(np.float64)


def oracle_solve(mif, mo, N, periodic):
    ctx, grid = oracle_pair(mif, mo, N, periodic)
    rng = np.random.default_rng(1234)
    host = [as_float(rng.uniform(-1, 1, grid.shape(c))) for c in range(3)]
    vel = ctx.velocity()
    for t, h in zip(vel, host):
        t.upload(h)
    p = ctx.tensor(mif.STAGGER_NONE)
    ctx.solve_pressure(p, vel, 0.37)
    err = rel(p.download(), grid.solve_pressure(*host, 0.37))
    ctx.close()
    return err


def oracle_timestep(mif, mo, N, periodic, kind):
    ctx, grid = oracle_pair(mif, mo, N, periodic)
    rng = np.random.default_rng(99)
    okind = {"ethier_steinman": mo.BC_ETHIER_STEINMAN, "test_case_1": mo.BC_TEST_CASE_1, "test_case_2": mo.BC_TEST_CASE_2}[kind]
    gkind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1, "test_case_2": mif.BC_TEST_CASE_2}[kind]
    h_vel = [as_float(0.3 * rng.uniform(-1, 1, grid.shape(c))) for c in range(3)]
    h_p = as_float(rng.uniform(-1, 1, grid.shape(3)))
    h_buf, h_buf2, h_dp = [grid.zeros(c) for c in range(3)], [grid.zeros(c) for c in range(3)], grid.zeros(3)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, h in zip(vel + [p], h_vel + [h_p]):
        t.upload(h)
    bc = ctx.make_bc(gkind, 1e3)
    worst = {"vel": 0.0, "p": 0.0}
    for step in range(2):
        ctx.timestep(vel, vb, vb2, bc, step * ctx.dt, p, dp)
        grid.timestep(okind, step * ctx.dt, h_vel, h_buf, h_buf2, h_p, h_dp)
        for t, h, name in zip(vel + [p], h_vel + [h_p], "uvwp"):
            key = "p" if name == "p" else "vel"
            worst[key] = max(worst[key], rel(t.download(), h))
    ctx.close()
    return worst


def velocity_only_case(mif, case):
    """mif::timestep_velocity with the manufactured forcing (src/TimestepVelocity.cpp) against the DOUBLE reference's fields
    of the velocity tests, inputs rounded to float: the forcing is evaluated in double in both builds."""
    from conftest import load_golden
    meta, f = load_golden(case)
    N = meta["N"]
    ctx = mif.Context(N[0], N[1], N[2], meta["x_size"], meta["y_size"], meta["z_size"], *meta["min"], meta["Re"],
                      meta["final_time"], meta["steps"], periodic=[bool(p) for p in meta["periodic"]])
    vel, vb, rb = ctx.velocity(), ctx.velocity(), ctx.velocity()
    for t, name in zip(vel, "uvw"):
        t.upload(f[name + "_s0"])
    bc = ctx.make_bc(mif.BC_VELOCITY_TEST, meta["Re"])
    dt = meta["final_time"] / meta["steps"]
    worst = 0.0
    for step in range(meta["steps"]):
        ctx.timestep_velocity(vel, vb, rb, bc, step * dt)
        vmax = max(float(np.max(np.abs(f[f"{c}_s{step + 1}"]))) for c in "uvw")
        for t, name in zip(vel, "uvw"):
            worst = max(worst, rel(t.download(), f[f"{name}_s{step + 1}"], floor=1e-3 * vmax))
    ctx.close()
    return worst


def nhn_solve_case(mif):
    """Pressure solve with non-homogeneous Neumann data through the host callback (float face tables) against the DOUBLE
    reference's ptest_nhn golden (test/pressure_test_nhn.cpp)."""
    from conftest import load_golden
    import test_gpu_golden as g64
    meta, f = load_golden("ptest_nhn_8x24x40")
    ctx = g64.make_ctx(mif, meta)
    vel = ctx.velocity()
    for t, name in zip(vel, ("u_in", "v_in", "w_in")):
        t.upload(f[name])
    p = ctx.tensor(mif.STAGGER_NONE)

    def cb(which, time, time_prev, comp, face, values):
        assert which == 1 and comp == 2 - face // 2 and values.dtype == np.float32
        x, y, z = g64.face_points(meta, face)
        values[...] = g64.ptest_gradient(comp, time, x, y, z)

    ctx.solve_pressure(p, vel, ctx.dt, ctx.make_bc(mif.BC_HOST_CALLBACK, 1.0, cb), meta["time"])
    err = rel(p.download(), f["p_out"])
    ctx.close()
    return err


def norms_case(mif):
    """The nine numbers `full_test 16 1 1` prints (test/full_test.cpp:36-187) from the float library, started from the
    t = 0 fields of the float reference's own run: returns the largest relative deviation of the six velocity /
    pressure norms from what the float reference prints (tests/golden/f32_norms.json).  The reference accumulates the
    norms serially in float, the library per CTA in float and across CTAs in double, hence the looser bound."""
    from conftest import GOLDEN_DIR, load_golden
    want = json.load(open(os.path.join(GOLDEN_DIR, "f32_norms.json")))["full_test 16 1 1"]
    _, f = load_golden("f32_full_16_2")
    Re, final_time = 1e3, 1e-4
    ctx = mif.Context(16, 16, 16, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, Re, final_time, 1)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, name in zip(vel + [p], "uvwp"):
        t.upload(f[name + "_s0"])
    bc = ctx.make_bc(mif.BC_ETHIER_STEINMAN, Re)
    ctx.timestep(vel, vb, vb2, bc, 0.0, p, dp)
    ctx.adjust_pressure(p, bc, final_time)
    got = list(ctx.velocity_error_norms(vel, bc, final_time)) + list(ctx.pressure_error_norms(p, bc, final_time))
    ctx.close()
    return max(abs(a - b) / abs(b) for a, b in zip(got, want[:6]))


def main():
    lib = os.environ.get("MIFGPU_LIB", "")
    assert "f32" in lib, "MIFGPU_LIB must name the float build of the library"
    import mif_b200 as mif
    import mif_oracle as mo
    assert mif.lib().mifgpu_real_bytes() == 4 and mif.real_dtype() == np.float32
    quick = "--quick" in sys.argv
    F, T = False, True
    out = {}
    for case in ("f32_full_16_2", "f32_lid1_12x10x14_2", "f32_lid2_10x12x9_2"):
        out["golden " + case] = golden_timestep(mif, case)
    grids = [((9, 7, 6), (F, F, F)), ((17, 33, 9), (F, F, F)), ((12, 10, 14), (F, F, T)), ((9, 10, 12), (T, T, T))]
    if not quick:
        grids += [((65, 9, 65), (F, F, F)), ((257, 6, 5), (F, F, F)), ((6, 5, 513), (F, F, F)), ((20, 9, 512 + 1), (F, F, T))]
    for N, periodic in grids:
        out[f"solve {N} {periodic}"] = oracle_solve(mif, mo, N, periodic)
    steps = [((9, 7, 6), (F, F, F), "ethier_steinman"), ((20, 13, 11), (F, F, F), "test_case_1"),
             ((12, 10, 14), (F, F, T), "test_case_2")]
    if not quick:
        steps += [((33, 17, 65), (F, F, F), "ethier_steinman")]
    for N, periodic, kind in steps:
        out[f"timestep {N} {periodic} {kind}"] = oracle_timestep(mif, mo, N, periodic, kind)
    out["velocity-only vtest_12_2"] = velocity_only_case(mif, "vtest_12_2")
    out["velocity-only vtest_mixed_12_2"] = velocity_only_case(mif, "vtest_mixed_12_2")
    out["solve nhn callback 8x24x40"] = nhn_solve_case(mif)
    out["norms full_test 16 1 1"] = norms_case(mif)
    print(json.dumps(out))
    return out


def check(out):
    """The assertions both callers make on main()'s output."""
    for name, err in out.items():
        if name.startswith("norms"):
            assert err <= 2e-3, (name, err)
        elif isinstance(err, dict):
            assert err["vel"] <= TOL32_VEL and err["p"] <= TOL32_P, (name, err)
        else:
            assert err <= TOL32_VEL, (name, err)


if __name__ == "__main__":
    main()
