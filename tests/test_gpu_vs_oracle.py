"""GPU parity proper: the CUDA path (through the C ABI) against the CPU oracle (oracle/mif_oracle.c, itself pinned
to the unmodified reference by tests/test_oracle_vs_reference.py) on the same seeded inputs, at sizes the oracle
finishes in seconds.  Covers non-cubic grids, every periodic/Neumann combination, power-of-two (register-blocked
fast path) and awkward (prime, Bluestein) line lengths, and both device-evaluated and host-callback boundary data.
Tolerance: 1e-11 relative L-infinity (north star)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mif_oracle as mo  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-11


def make_pair(mif, N, periodic, Re=1e3, final_time=1e-3, steps=4, size=(1.0, 1.0, 2.0), lo=(0.0, 0.0, -1.0)):
    ctx = mif.Context(N[0], N[1], N[2], *size, *lo, Re, final_time, steps, periodic=periodic)
    grid = mo.Grid(N[0], N[1], N[2], *size, *lo, Re, final_time, steps, periodic=periodic)
    return ctx, grid


def rel(a, b):
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300)


GRIDS = [
    ((9, 7, 6), (False, False, False)),
    ((20, 13, 11), (False, False, False)),
    ((17, 33, 9), (False, False, False)),      # DCT-I lengths 16+1, 32+1, 8+1 (power-of-two, generic kernel)
    ((65, 9, 65), (False, False, False)),      # fast path in x and z
    ((12, 129, 7), (False, False, False)),     # fast path (radix 8,8,2) in y
    ((12, 10, 14), (False, False, True)),      # periodic z: odd real-FFT length 13 (Bluestein)
    ((10, 12, 9), (False, False, True)),       # periodic z: even length 8
    ((11, 9, 8), (True, False, False)),        # periodic x
    ((8, 11, 9), (False, True, False)),        # periodic y
    ((9, 10, 12), (True, True, True)),         # fully periodic
    ((2, 5, 3), (False, False, False)),        # smallest legal extents
    ((257, 6, 5), (False, False, False)),      # warp-per-line DCT path, M = 256, x sweep
    ((5, 257, 6), (False, False, False)),      # ... y sweep
    ((6, 5, 257), (False, False, False)),      # ... fused z sweep
    ((513, 4, 9), (False, False, False)),      # M = 512
    ((9, 513, 3), (False, False, False)),
    ((11, 3, 513), (False, False, False)),
    ((1025, 3, 4), (False, False, False)),     # M = 1024 (split into two 512-point halves, one warp each)
    ((10, 4, 1025), (False, False, False)),
    ((9, 1025, 3), (False, False, False)),
    ((20, 9, 513), (False, False, True)),      # periodic z, n = 512: warp real-FFT path (M = 256), fused z sweep
    ((513, 5, 6), (True, False, False)),       # ... periodic x: R2HC / HC2R as separate x sweeps
    ((6, 513, 5), (False, True, False)),       # ... periodic y
    ((5, 4, 1025), (False, False, True)),      # n = 1024 (M = 512)
    ((1025, 4, 5), (True, False, False)),
    ((4, 1025, 5), (False, True, False)),
    ((513, 513, 5), (True, True, False)),      # two periodic directions, full tiles
]


@pytest.mark.parametrize("N,periodic", GRIDS)
def test_pressure_solve_random_velocity(mif, N, periodic):
    ctx, grid = make_pair(mif, N, periodic)
    rng = np.random.default_rng(1234)
    host = [rng.uniform(-1, 1, grid.shape(c)) for c in range(3)]
    vel = ctx.velocity()
    for t, h in zip(vel, host):
        t.upload(h)
    p = ctx.tensor(mif.STAGGER_NONE)
    ctx.solve_pressure(p, vel, 0.37)
    want = grid.solve_pressure(*host, 0.37)
    assert rel(p.download(), want) <= TOL
    ctx.close()


@pytest.mark.parametrize("N,periodic", GRIDS[:10] + GRIDS[11:17:2])
@pytest.mark.parametrize("kind", ["ethier_steinman", "test_case_1", "test_case_2"])
def test_timestep_random_state(mif, N, periodic, kind):
    ctx, grid = make_pair(mif, N, periodic)
    rng = np.random.default_rng(99)
    okind = {"ethier_steinman": mo.BC_ETHIER_STEINMAN, "test_case_1": mo.BC_TEST_CASE_1, "test_case_2": mo.BC_TEST_CASE_2}[kind]
    gkind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1, "test_case_2": mif.BC_TEST_CASE_2}[kind]
    h_vel = [0.3 * rng.uniform(-1, 1, grid.shape(c)) for c in range(3)]
    h_p = rng.uniform(-1, 1, grid.shape(3))
    h_buf = [grid.zeros(c) for c in range(3)]
    h_buf2 = [grid.zeros(c) for c in range(3)]
    h_dp = grid.zeros(3)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, h in zip(vel + [p], h_vel + [h_p]):
        t.upload(h)
    bc = ctx.make_bc(gkind, 1e3)
    for step in range(2):
        t_n = step * ctx.dt
        ctx.timestep(vel, vb, vb2, bc, t_n, p, dp)
        grid.timestep(okind, t_n, h_vel, h_buf, h_buf2, h_p, h_dp)
        for t, h, name in zip(vel + [p], h_vel + [h_p], "uvwp"):
            assert rel(t.download(), h) <= TOL, (name, step)
    # the scratch tensors hold the same contents the reference leaves in them
    for t, h, name in zip(vb + vb2 + [dp], h_buf + h_buf2 + [h_dp], ["ub", "vb", "wb", "ub2", "vb2", "wb2", "dp"]):
        # delta-p is the one field whose own noise floor is at the 1e-11 level already between two builds of the
        # reference (rhs = div(u)/dt_s amplifies round-off), so it gets a looser bound.
        assert rel(t.download(), h) <= (TOL if name != "dp" else 1e-8), name
    ctx.close()


@pytest.mark.parametrize("N,kind", [
    ((64, 64, 64), "ethier_steinman"),      # BASELINE configs[0] at its own size (full_test 64): generic sweep kernel, field level
    ((257, 257, 257), "ethier_steinman"),   # a full 3-D grid with all three directions on the fast kernels (x sweeps from
                                            # registers, TMA-staged y / fused z sweeps, persistent CTAs over many tiles)
    ((513, 1025, 4), "test_case_1"),        # planes of more than 2.2 MB: the y-chunked stage launch (n_chunks > 1), 513- and
                                            # 1025-point lines (16 x 32 transform, split 1024 transform) on one GPU
])
def test_timestep_full_grids(mif, N, kind):
    """One projection step on grids the small cases above cannot reach, u v w p and the scratch tensors at 1e-11."""
    periodic = (False, False, False)
    ctx, grid = make_pair(mif, N, periodic)
    rng = np.random.default_rng(7)
    okind = {"ethier_steinman": mo.BC_ETHIER_STEINMAN, "test_case_1": mo.BC_TEST_CASE_1}[kind]
    gkind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1}[kind]
    h_vel = [0.3 * rng.uniform(-1, 1, grid.shape(c)) for c in range(3)]
    h_p = rng.uniform(-1, 1, grid.shape(3))
    h_buf, h_buf2, h_dp = [grid.zeros(c) for c in range(3)], [grid.zeros(c) for c in range(3)], grid.zeros(3)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, h in zip(vel + [p], h_vel + [h_p]):
        t.upload(h)
    ctx.timestep(vel, vb, vb2, ctx.make_bc(gkind, 1e3), 0.0, p, dp)
    grid.timestep(okind, 0.0, h_vel, h_buf, h_buf2, h_p, h_dp)
    for t, h, name in zip(vel + [p] + vb + vb2, h_vel + [h_p] + h_buf + h_buf2, ["u", "v", "w", "p", "ub", "vb", "wb", "ub2", "vb2", "wb2"]):
        assert rel(t.download(), h) <= TOL, name
    ctx.close()


def test_timestep_nhn_matches_oracle(mif):
    N, periodic = (14, 11, 9), (False, False, False)
    ctx, grid = make_pair(mif, N, periodic, final_time=1e-4, steps=2)
    h_vel = list(grid.set_velocity(mo.BC_ETHIER_STEINMAN, 0.0))
    rng = np.random.default_rng(5)
    h_p = rng.uniform(-1, 1, grid.shape(3))
    h_buf = [grid.zeros(c) for c in range(3)]
    h_buf2 = [grid.zeros(c) for c in range(3)]
    h_dp = grid.zeros(3)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, h in zip(vel + [p], h_vel + [h_p]):
        t.upload(h)
    h = [1.0 / (N[0] - 1), 1.0 / (N[1] - 1), 2.0 / (N[2] - 1)]
    axes = [np.array([lo + h[d] * i for i in range(N[d])]) for d, lo in enumerate((0.0, 0.0, -1.0))]
    grad = np.vectorize(lambda comp, t, x, y, z: mo.lib().mo_es_pressure_gradient(int(comp), t, x, y, z, 1e3))

    def cb(which, time, time_prev, comp, face, values):
        d = 2 - face // 2
        fixed = axes[d][-1] if face & 1 else axes[d][0]
        if d == 2:
            y, x = np.meshgrid(axes[1], axes[0], indexing="ij"); z = np.full_like(x, fixed)
        elif d == 1:
            z, x = np.meshgrid(axes[2], axes[0], indexing="ij"); y = np.full_like(x, fixed)
        else:
            z, y = np.meshgrid(axes[2], axes[1], indexing="ij"); x = np.full_like(y, fixed)
        values[...] = grad(comp, time_prev, x, y, z) - grad(comp, time, x, y, z)

    bc = ctx.make_bc(mif.BC_ETHIER_STEINMAN, 1e3, cb)
    ctx.timestep(vel, vb, vb2, bc, 0.0, p, dp, nhn=True)
    grid.timestep(mo.BC_ETHIER_STEINMAN, 0.0, h_vel, h_buf, h_buf2, h_p, h_dp, nhn=True)
    for t, hh, name in zip(vel + [p], h_vel + [h_p], "uvwp"):
        assert rel(t.download(), hh) <= TOL, name
    ctx.close()


@pytest.mark.parametrize("N,periodic", [((9, 7, 6), (False, False, False)), ((10, 12, 9), (False, False, True)),
                                        ((11, 9, 8), (True, False, False))])
def test_apply_bc_device_and_host_callback_agree_with_oracle(mif, N, periodic):
    ctx, grid = make_pair(mif, N, periodic)
    rng = np.random.default_rng(3)
    host = [rng.uniform(-1, 1, grid.shape(c)) for c in range(3)]
    want = [h.copy() for h in host]
    grid.apply_bc(mo.BC_ETHIER_STEINMAN, 0.123, *want)
    vel = ctx.velocity()
    # (a) analytic family evaluated on the device
    for t, h in zip(vel, host):
        t.upload(h)
    ctx.apply_bc(vel, ctx.make_bc(mif.BC_ETHIER_STEINMAN, 1e3), 0.123)
    for t, w in zip(vel, want):
        assert rel(t.download(), w) <= 1e-13
    # (b) the generic path: faces filled on the host (here: cut out of the oracle's result)
    def cb(which, time, time_prev, comp, face, values):
        assert which == 0 and time == 0.123
        w = want[comp]
        d = 2 - face // 2
        idx = -1 if face & 1 else 0
        values[...] = w[idx, :, :] if d == 2 else (w[:, idx, :] if d == 1 else w[:, :, idx])
    for t, h in zip(vel, host):
        t.upload(h)
    ctx.apply_bc(vel, ctx.make_bc(mif.BC_HOST_CALLBACK, 1e3, cb), 0.123)
    for t, w in zip(vel, want):
        assert rel(t.download(), w) <= 1e-15
    ctx.close()


@pytest.mark.parametrize("N,periodic", [((16, 16, 16), (False, False, False)), ((70, 9, 11), (False, False, False)),
                                        ((10, 12, 9), (False, False, True)), ((11, 9, 8), (True, False, False)),
                                        ((130, 67, 5), (False, True, False))])
def test_device_norms_and_adjust_pressure_match_oracle(mif, N, periodic):
    """ErrorL1/L2/LInfNorm (src/Norms.cpp) and adjust_pressure (src/PressureEquation.cpp:288-343)
    on the device against the oracle's serial restatement; the summation order differs, hence 1e-12 and not 0."""
    ctx, grid = make_pair(mif, N, periodic)
    rng = np.random.default_rng(21)
    t = 0.37
    h_vel = [a + 1e-3 * rng.uniform(-1, 1, a.shape) for a in grid.set_velocity(mo.BC_ETHIER_STEINMAN, t)]
    h_p = rng.uniform(-1, 1, grid.shape(3))
    vel, p = ctx.velocity(), ctx.tensor(mif.STAGGER_NONE)
    for ten, h in zip(vel + [p], h_vel + [h_p]):
        ten.upload(h)
    exact = ctx.make_bc(mif.BC_ETHIER_STEINMAN, 1e3)
    for got, want in zip(ctx.velocity_error_norms(vel, exact, t), grid.velocity_error_norms(mo.BC_ETHIER_STEINMAN, t, *h_vel)):
        assert abs(got - want) <= 1e-12 * abs(want)
    for got, want in zip(ctx.pressure_error_norms(p, exact, t), grid.pressure_error_norms(mo.BC_ETHIER_STEINMAN, t, h_p)):
        assert abs(got - want) <= 1e-12 * abs(want)
    ctx.adjust_pressure(p, exact, t)
    grid.adjust_pressure(mo.BC_ETHIER_STEINMAN, t, h_p)
    assert rel(p.download(), h_p) <= 1e-13
    # the lid-driven families have p = 0 as their reference pressure and an exact velocity that is zero inside
    lid = ctx.make_bc(mif.BC_TEST_CASE_1, 1e3)
    for got, want in zip(ctx.velocity_error_norms(vel, lid, t), grid.velocity_error_norms(mo.BC_TEST_CASE_1, t, *h_vel)):
        assert abs(got - want) <= 1e-12 * abs(want)
    # arbitrary host functions cannot be evaluated on the device: refused, never silently computed elsewhere
    with pytest.raises(mif.MifGpuError):
        ctx.velocity_error_norms(vel, ctx.make_bc(mif.BC_HOST_CALLBACK, 1e3, lambda *a: None), t)
    ctx.close()


@pytest.mark.parametrize("N,periodic", [((9, 7, 6), (False, False, False)), ((20, 13, 11), (False, False, False)),
                                        ((12, 10, 14), (True, True, False)), ((70, 6, 9), (False, False, True)),
                                        ((8, 11, 9), (False, True, False))])
def test_timestep_velocity_random_state(mif, N, periodic):
    """mif::timestep_velocity with the manufactured forcing against the oracle on seeded states."""
    ctx, grid = make_pair(mif, N, periodic, Re=1e4, final_time=1e-2, steps=4)
    rng = np.random.default_rng(17)
    h_vel = [0.3 * rng.uniform(-1, 1, grid.shape(c)) for c in range(3)]
    h_buf, h_rhs = [grid.zeros(c) for c in range(3)], [grid.zeros(c) for c in range(3)]
    vel, vb, rb = ctx.velocity(), ctx.velocity(), ctx.velocity()
    for t, h in zip(vel, h_vel):
        t.upload(h)
    bc = ctx.make_bc(mif.BC_VELOCITY_TEST, 1e4)
    for step in range(2):
        ctx.timestep_velocity(vel, vb, rb, bc, 0.3 + step * ctx.dt)
        grid.timestep_velocity(mo.BC_VELOCITY_TEST, 0.3 + step * ctx.dt, h_vel, h_buf, h_rhs)
        for t, h, name in zip(vel + vb + rb, h_vel + h_buf + h_rhs, ["u", "v", "w", "ub", "vb", "wb", "ru", "rv", "rw"]):
            assert rel(t.download(), h) <= TOL, (name, step)
    ctx.close()


def test_upload_download_roundtrip_and_swap(mif):
    ctx, grid = make_pair(mif, (7, 5, 6), (False, True, False))
    rng = np.random.default_rng(0)
    a, b = ctx.tensor(mif.STAGGER_Y), ctx.tensor(mif.STAGGER_Y)
    ha, hb = rng.uniform(-1, 1, grid.shape(1)), rng.uniform(-1, 1, grid.shape(1))
    assert a.shape == grid.shape(1)[::-1]
    assert np.all(a.download() == 0.0)  # zero-initialised like std::vector
    a.upload(ha); b.upload(hb)
    mif._check(mif.lib().mifgpu_tensor_swap(a.handle, b.handle))
    assert np.array_equal(a.download(), hb) and np.array_equal(b.download(), ha)
    ctx.close()


def test_async_transfers_pipeline_matches_synchronous_calls(mif, N=(33, 20, 17)):
    """mifgpu_tensor_upload_async / _download_async (per-direction copy streams, event ordered against the compute
    stream): three independent single-step jobs rotated through two device field sets give exactly the fields of the
    same jobs run one after the other with the blocking transfers -- including the re-use of a set whose previous
    download is still in flight when the next upload is enqueued."""
    periodic = (False, False, False)
    ctx, grid = make_pair(mif, N, periodic)
    rng = np.random.default_rng(2024)
    jobs = [[0.3 * rng.uniform(-1, 1, grid.shape(c)) for c in range(3)] + [rng.uniform(-1, 1, grid.shape(3))] for _ in range(3)]
    bc = ctx.make_bc(mif.BC_ETHIER_STEINMAN, 1e3)
    vb, vb2, dp = ctx.velocity(), ctx.velocity(), ctx.tensor(mif.STAGGER_NONE)
    # blocking reference
    vel, p = ctx.velocity(), ctx.tensor(mif.STAGGER_NONE)
    want = []
    for fields in jobs:
        for t, h in zip(vel + [p], fields):
            t.upload(h)
        ctx.timestep(vel, vb, vb2, bc, 0.0, p, dp)
        want.append([t.download() for t in vel + [p]])
    # pipelined: nothing blocks the host until the final synchronize
    sets = [(ctx.velocity(), ctx.tensor(mif.STAGGER_NONE)) for _ in range(2)]
    got = [[np.empty_like(h) for h in fields] for fields in jobs]
    for i, fields in enumerate(jobs):
        v_i, p_i = sets[i % 2]
        for t, h in zip(v_i + [p_i], fields):
            t.upload_async(h)
        ctx.timestep(v_i, vb, vb2, bc, 0.0, p_i, dp)
        for t, out in zip(v_i + [p_i], got[i]):
            t.download_async(out)
    ctx.synchronize()
    for i in range(3):
        for a, b, name in zip(got[i], want[i], "uvwp"):
            assert np.array_equal(a, b), (i, name)
    with pytest.raises(ValueError):
        sets[0][1].download_async(np.empty(3))  # wrong size: refused before the library sees it
    ctx.close()
