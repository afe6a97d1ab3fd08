"""GPU parity against the golden vectors dumped from the UNMODIFIED reference (oracle/make_golden.py).

Every call goes through the C ABI (include/mifgpu.h) via the ctypes view in mif_b200.  Tolerance: the
north star's 1e-11 relative L-infinity per step for u, v, w, p, each normalised by its own maximum
(delta-p alone is already at 2e-11 between two builds of the reference).
"""
import numpy as np
import pytest

from conftest import load_golden, rel_linf

pytestmark = pytest.mark.gpu

TOL = 1e-11
# timestep_nhn feeds g(t_new) - g(t_old) of the exact pressure gradient into the rhs: the two terms agree
# to ~7 digits (the stage is 5e-5 long and the decay rate is 5e-3), so the data itself carries ~1e-9
# relative rounding noise that differs between any two evaluations of the same formula (sympy-generated
# C in the reference, numpy here).  The comparison for that variant is therefore looser.
TOL_NHN = 2e-9


def make_ctx(mif, meta):
    N = meta["N"]
    return mif.Context(N[0], N[1], N[2], meta["x_size"], meta["y_size"], meta["z_size"], *meta["min"], meta["Re"],
                       meta["final_time"], meta["steps"], periodic=[bool(p) for p in meta["periodic"]])


def ptest_gradient(comp, t, x, y, z):
    # grad of p = t cos x cos y cos z (generators/manufsol_pressure.py:36-39)
    if comp == 0:
        return -t * np.sin(x) * np.cos(y) * np.cos(z)
    if comp == 1:
        return -t * np.cos(x) * np.sin(y) * np.cos(z)
    return -t * np.cos(x) * np.cos(y) * np.sin(z)


def face_points(meta, face):
    """Unstaggered coordinates of the pressure points on one face, shaped like the callback's `values`."""
    N = meta["N"]
    h = [meta["x_size"] / (N[0] - 1), meta["y_size"] / (N[1] - 1), meta["z_size"] / (N[2] - 1)]
    axes = [meta["min"][d] + h[d] * np.arange(N[d]) for d in range(3)]
    d = 2 - face // 2
    fixed = axes[d][-1] if face & 1 else axes[d][0]
    if d == 2:
        y, x = np.meshgrid(axes[1], axes[0], indexing="ij")
        return x, y, np.full_like(x, fixed)
    if d == 1:
        z, x = np.meshgrid(axes[2], axes[0], indexing="ij")
        return x, np.full_like(x, fixed), z
    z, y = np.meshgrid(axes[2], axes[1], indexing="ij")
    return np.full_like(y, fixed), y, z


@pytest.mark.parametrize("case", ["ptest_hn_8x24x40", "ptest_mixed_8x24x40", "ptest_nhn_8x24x40", "ptest_mixed_9x17x17",
                                  "ptest_hn_17x9x33", "ptest_mixed_7x6x10", "ptest_hn_65x9x129",
                                  "ptest_mixed_9x65x17"])
def test_pressure_solve_matches_reference(mif, case):
    meta, f = load_golden(case)
    ctx = make_ctx(mif, meta)
    vel = ctx.velocity()
    for t, name in zip(vel, ("u_in", "v_in", "w_in")):
        assert t.shape == f[name].shape[::-1]
        t.upload(f[name])
    p = ctx.tensor(mif.STAGGER_NONE)
    bc = None
    if meta["ptest"] == "nhn":
        def cb(which, time, time_prev, comp, face, values):
            assert which == 1 and comp == 2 - face // 2
            x, y, z = face_points(meta, face)
            values[...] = ptest_gradient(comp, time, x, y, z)
        bc = ctx.make_bc(mif.BC_HOST_CALLBACK, 1.0, cb)
    ctx.solve_pressure(p, vel, ctx.dt, bc, meta["time"])
    got = p.download()
    err = rel_linf(got, f["p_out"])
    assert err <= TOL, err
    ctx.close()


def es_gradient_functions(Re):
    """dp/dx, dp/dy, dp/dz of the Ethier-Steinman pressure (generators/manufsol.py:58-72), via sympy."""
    import sympy as sp
    t, x, y, z = sp.symbols("t x y z")
    a, d = sp.pi / 4, sp.pi / 2
    p = (-a * a / 2 * (sp.exp(2 * a * x) + sp.exp(2 * a * y) + sp.exp(2 * a * z)
                       + 2 * sp.sin(a * x + d * y) * sp.cos(a * z + d * x) * sp.exp(a * (y + z))
                       + 2 * sp.sin(a * y + d * z) * sp.cos(a * x + d * y) * sp.exp(a * (z + x))
                       + 2 * sp.sin(a * z + d * x) * sp.cos(a * y + d * z) * sp.exp(a * (x + y)))
         * sp.exp(-2 * d * d * t / Re))
    return [sp.lambdify((t, x, y, z), sp.diff(p, v), "numpy") for v in (x, y, z)]


@pytest.mark.parametrize("case", ["full_16_2", "full_17_1", "full_12_1_nhn", "lid1_12x10x14_2", "lid2_10x12x9_2",
                                  "full_65x17x9_1", "full_6x65x9_1"])
def test_timestep_matches_reference(mif, case):
    meta, f = load_golden(case)
    ctx = make_ctx(mif, meta)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, name in zip(vel, "uvw"):
        t.upload(f[name + "_s0"])
    p.upload(f["p_s0"])
    kind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1,
            "test_case_2": mif.BC_TEST_CASE_2}[meta["bc"]]
    nhn = bool(meta.get("nhn", 0))
    cb = None
    if nhn:
        grads = es_gradient_functions(meta["Re"])

        def cb(which, time, time_prev, comp, face, values):
            assert which == 1
            x, y, z = face_points(meta, face)
            # The reference passes exact_pressure_gradient.get_difference_over_time(new_time, prev_time)
            # (src/Timestep.cpp:92), which evaluates f(prev_time) - f(new_time) (src/VectorFunction.cpp:52-60).
            values[...] = grads[comp](time_prev, x, y, z) - grads[comp](time, x, y, z)
    bc = ctx.make_bc(kind, meta["Re"], cb)
    dt = meta["final_time"] / meta["steps"]
    for step in range(meta["steps"]):
        ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp, nhn=nhn)
        ctx.synchronize()
        # A velocity component that is identically zero in exact arithmetic (w in the z-periodic lid case) holds
        # only rounding noise (max|w| ~ 1e-10, and two builds of the reference itself differ by 9e-11 of that), so
        # a component is normalised by max(its own maximum, 1e-6 of the largest velocity component).
        vmax = max(float(np.max(np.abs(f[f"{c}_s{step + 1}"]))) for c in "uvw")
        for t, name in zip(vel + [p], "uvwp"):
            ref = f[f"{name}_s{step + 1}"]
            floor = 1e-6 * vmax if name in "uvw" else 0.0
            err = float(np.max(np.abs(t.download() - ref))) / max(float(np.max(np.abs(ref))), floor)
            assert err <= (TOL_NHN if nhn else TOL), (name, step + 1, err)
    ctx.close()


@pytest.mark.parametrize("case", ["vtest_12_2", "vtest_mixed_12_2"])
def test_timestep_velocity_matches_reference(mif, case):
    """mif::timestep_velocity (src/TimestepVelocity.cpp) against the fields of the reference's velocity tests."""
    meta, f = load_golden(case)
    ctx = make_ctx(mif, meta)
    vel, vb, rb = ctx.velocity(), ctx.velocity(), ctx.velocity()
    for t, name in zip(vel, "uvw"):
        t.upload(f[name + "_s0"])
    bc = ctx.make_bc(mif.BC_VELOCITY_TEST, meta["Re"])
    dt = meta["final_time"] / meta["steps"]
    for step in range(meta["steps"]):
        ctx.timestep_velocity(vel, vb, rb, bc, step * dt)
        vmax = max(float(np.max(np.abs(f[f"{c}_s{step + 1}"]))) for c in "uvw")
        for t, name in zip(vel, "uvw"):
            ref = f[f"{name}_s{step + 1}"]
            err = float(np.max(np.abs(t.download() - ref))) / max(float(np.max(np.abs(ref))), 1e-6 * vmax)
            assert err <= TOL, (name, step + 1, err)
    ctx.close()
