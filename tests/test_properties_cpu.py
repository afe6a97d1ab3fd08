"""CPU: the property checkers of tests/properties.py against the oracle at small sizes.  The same checkers run against
the CUDA path at 513^3 points in tests/test_gpu_zz_full_size.py; here they are shown to hold for the reference's
algorithm (and to fail for a deliberately broken input), so a failure on the GPU means the CUDA path, not the checker."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
import properties as prop

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mif_oracle as mo  # noqa: E402


def make_grid(N, periodic, size=(1.0, 1.0, 2.0), lo=(0.0, 0.0, -1.0), dt=1e-3, steps=4):
    grid = mo.Grid(N[0], N[1], N[2], *size, *lo, 1e3, dt * steps, steps, periodic=periodic)
    h = [size[d] / (N[d] - 1) for d in range(3)]
    return grid, h


@pytest.mark.parametrize("N,periodic", [((17, 12, 9), (False, False, False)), ((33, 17, 20), (False, False, True)),
                                        ((12, 16, 9), (True, False, False)), ((9, 10, 12), (True, True, True)),
                                        ((65, 33, 17), (False, False, False))])
def test_poisson_residual_and_gauge(N, periodic):
    grid, h = make_grid(N, periodic)
    rng = np.random.default_rng(3)
    u, v, w = (rng.uniform(-1, 1, grid.shape(c)) for c in range(3))
    p = grid.solve_pressure(u, v, w, 0.37)
    residual, gauge = prop.poisson_defects(p, u, v, w, 0.37, h, periodic)
    assert residual <= 1e-11 and gauge <= 1e-13, (residual, gauge)
    # the checker notices a solution that is wrong in a single point
    q = p.copy()
    q[N[2] // 2, N[1] // 2, N[0] // 2] *= 1.001
    assert prop.poisson_defects(q, u, v, w, 0.37, h, periodic)[0] > 1e-6


def test_poisson_linearity_and_shift():
    N, periodic = (17, 9, 21), (False, False, True)
    two_pi = 2 * np.pi
    grid, h = make_grid(N, periodic, size=(two_pi,) * 3, lo=(0.0,) * 3)
    rng = np.random.default_rng(11)
    n = N[2] - 1  # period in z
    bases = [[rng.uniform(-1, 1, (n,) + grid.shape(c)[1:]) for c in range(3)] for _ in range(2)]
    fields = [[prop.periodic_z_field(b, grid.shape(c)[0]) for c, b in enumerate(base)] for base in bases]
    pa, pb = (grid.solve_pressure(*f, 1.0) for f in fields)
    a, b = 0.75, -1.5
    pc = grid.solve_pressure(*[a * x + b * y for x, y in zip(*fields)], 1.0)
    own = prop.owner_slices(pa.shape, periodic)
    assert prop.rel_diff(pc[own], a * pa[own] + b * pb[own]) <= 1e-12
    shift = 5
    shifted = [prop.periodic_z_field(base, grid.shape(c)[0], shift) for c, base in enumerate(bases[0])]
    ps = grid.solve_pressure(*shifted, 1.0)
    assert prop.rel_diff(ps[own], np.roll(pa[own], -shift, axis=0)) <= 1e-12
    assert prop.rel_diff(ps[own], pa[own]) > 1e-3  # the shift is not a no-op


def test_lid_case_is_mirror_symmetric_in_z():
    N = (13, 11, 17)
    steps = 3
    grid, _ = make_grid(N, (False, False, False), steps=steps)
    vel = list(grid.set_velocity(mo.BC_TEST_CASE_1, 0.0))
    buf, buf2 = [grid.zeros(c) for c in range(3)], [grid.zeros(c) for c in range(3)]
    p, dp = grid.zeros(3), grid.zeros(3)
    for step in range(steps):
        grid.timestep(mo.BC_TEST_CASE_1, step * grid.dt, vel, buf, buf2, p, dp)
    assert float(np.max(np.abs(vel[2]))) > 1e-6  # w has developed: the odd symmetry is not trivial
    defects = prop.z_mirror_defects(*vel, p)
    assert max(defects) <= 1e-12, defects
    vel[2][3, 4, 5] += 1e-6
    assert prop.z_mirror_defects(*vel, p)[2] > 1e-8


@pytest.mark.parametrize("size", [17, 33])
def test_full_size_gpu_test_bodies_hold_for_the_oracle(size, monkeypatch):
    """The bodies of tests/test_gpu_zz_full_size.py, with the oracle standing in for libmifgpu at a small size."""
    import oracle_backend
    import test_gpu_zz_full_size as full
    monkeypatch.setattr(full, "FULL", size)
    full.test_config2_poisson_solve_properties_at_full_size(oracle_backend)
    full.test_config3_timestep_keeps_the_mirror_symmetry_at_full_size(oracle_backend)
