"""GPU: the CUDA path at BASELINE.json's full sizes (513^3 pressure points = 512^3 cells), where the oracle is too
slow to be the checker, through the size-independent properties of tests/properties.py -- the same numpy checkers that
tests/test_properties_cpu.py validates against the oracle at small sizes.

* configs[1] (pressure_test_mixed-type Poisson solve, x / y Neumann, z periodic): the solution satisfies the discrete
  equation L p = rhs - <rhs> with <p> = 0, the solve is linear, and it commutes with a periodic shift in z.
* configs[2] (full projection time step, input.txt boundary-layer set-up): the lid-driven flow stays mirror symmetric
  in z (u, v, p even, w odd) through whole time steps; apply_bc is idempotent; host <-> device copies are exact.

MIF_FULL_SIZE overrides the number of points per direction (default 513).
(File name sorts last on purpose: added after the last GPU session of round 1.)"""
import os

import numpy as np
import pytest

import properties as prop

pytestmark = pytest.mark.gpu

FULL = int(os.environ.get("MIF_FULL_SIZE", "513"))


def np_shape(t):
    sx, sy, sz = t.shape
    return (sz, sy, sx)


def test_config2_poisson_solve_properties_at_full_size(mif):
    N, periodic = FULL, (False, False, True)
    two_pi = 2.0 * np.pi
    h = [two_pi / (N - 1)] * 3
    ctx = mif.Context(N, N, N, two_pi, two_pi, two_pi, 0.0, 0.0, 0.0, 1e3, 1.0, 1, periodic=periodic)
    vel, p = ctx.velocity(), ctx.tensor(mif.STAGGER_NONE)
    rng = np.random.default_rng(2024)
    n = N - 1  # period in z

    def solve(fields):
        for t, f in zip(vel, fields):
            t.upload(f)
        ctx.solve_pressure(p, vel, 1.0)
        return p.download()

    bases = [rng.uniform(-1, 1, (n,) + np_shape(t)[1:]) for t in vel]
    fa = [prop.periodic_z_field(b, np_shape(t)[0]) for b, t in zip(bases, vel)]
    pa = solve(fa)
    assert np.isfinite(pa).all()
    residual, gauge = prop.poisson_defects(pa, *fa, 1.0, h, periodic)
    assert residual <= 1e-11, residual   # oracle at 33^3 .. 257x129x65: 1e-15 .. 2.5e-15
    assert gauge <= 1e-13, gauge
    own = prop.owner_slices(pa.shape, periodic)

    shift = 37
    ps = solve([prop.periodic_z_field(b, np_shape(t)[0], shift) for b, t in zip(bases, vel)])
    assert prop.rel_diff(ps[own], np.roll(pa[own], -shift, axis=0)) <= 1e-11
    del ps, bases

    fb = [rng.uniform(-1, 1, np_shape(t)) for t in vel]
    pb = solve(fb)
    a, b = 0.75, -1.5
    for x, y in zip(fa, fb):  # fa <- a fa + b fb, in place
        x *= a
        x += b * y
    del fb
    pc = solve(fa)
    pa *= a
    pa += b * pb
    assert prop.rel_diff(pc[own], pa[own]) <= 1e-11
    ctx.close()


def test_config3_timestep_keeps_the_mirror_symmetry_at_full_size(mif):
    N, dt, steps = FULL, 1e-3, 2
    ctx = mif.Context(N, N, N, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, dt * steps, steps)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    bc = ctx.make_bc(mif.BC_TEST_CASE_1, 1e3)
    # velocity.set(exact(t = 0), include_border = true), src/main.cpp:144-146: v = 1 on the plane x = 1, else 0
    v0 = np.zeros(np_shape(vel[1]))
    v0[:, :, N - 1] = 1.0
    vel[1].upload(v0)
    assert np.array_equal(vel[1].download(), v0)  # host -> device -> host is exact at full size
    del v0
    for step in range(steps):
        ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp)
    u, v, w = (t.download() for t in vel)
    pr = p.download()
    assert all(np.isfinite(a).all() for a in (u, v, w, pr))
    assert float(np.max(np.abs(w))) > 1e-6 and float(np.max(np.abs(u))) > 1e-6  # the flow has developed
    defects = prop.z_mirror_defects(u, v, w, pr)
    assert max(defects[:3]) <= 1e-10 and defects[3] <= 1e-9, defects
    # one plane through the box download equals the same plane of the whole-field download
    k = N // 3
    assert np.array_equal(vel[0].download_box((0, 0, k), (vel[0].shape[0], vel[0].shape[1], k + 1))[0], u[k])
    # apply_bc is idempotent: it writes boundary values from analytic data and from interior values it leaves alone
    ctx.apply_bc(vel, bc, steps * dt)
    once = [t.download() for t in vel]
    ctx.apply_bc(vel, bc, steps * dt)
    for t, a in zip(vel, once):
        assert np.array_equal(t.download(), a)
    ctx.close()
