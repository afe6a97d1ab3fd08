"""GPU parity against the oracle for the line lengths whose kernels were written or changed after the last GPU session of
round 1 (so far verified under the SIMT interpreter only): the M = 2048 instantiation of the warp-per-line kernel (the
2049-point x lines of BASELINE configs[4]) and the two-stage passes of the generic kernel (Bluestein lengths).
(File name sorts last on purpose.)"""
import pytest

import test_gpu_vs_oracle as parity

pytestmark = pytest.mark.gpu

NEW_GRIDS = [
    ((2049, 3, 4), (False, False, False)),     # M = 2048, four warps per line: x sweeps
    ((5, 3, 2049), (False, False, False)),     # ... as a fused z sweep
    ((3, 2049, 4), (False, False, False)),     # ... as y sweeps
    ((3000, 3, 2), (False, False, False)),     # beyond the warp kernels: generic kernel (Bluestein, m = 2999, P = 8192: 155 KB of shared memory)
    ((33, 6, 481), (False, False, True)),      # the reference's default 480-point period (input/input.txt): Bluestein
    ((20, 13, 11), (False, False, False)),     # generic kernel, small Bluestein lengths in all directions
]


@pytest.mark.parametrize("N,periodic", NEW_GRIDS)
def test_pressure_solve_random_velocity_new_sizes(mif, N, periodic):
    parity.test_pressure_solve_random_velocity(mif, N, periodic)


@pytest.mark.parametrize("N,periodic", NEW_GRIDS[:2])
def test_timestep_random_state_new_sizes(mif, N, periodic):
    parity.test_timestep_random_state(mif, N, periodic, "ethier_steinman")


def test_line_lengths_beyond_the_shared_memory_transform_are_refused(mif):
    """Bluestein transform lengths above 4096 (P = 16384 complex values do not fit 227 KB): a clean error at create time."""
    with pytest.raises(mif.MifGpuError, match="-3"):
        mif.Context(4100, 3, 2, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-3, 4)
    ctx = mif.Context(4097, 3, 2, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-3, 4)  # 2^12 + 1: power-of-two transform, fits
    ctx.close()
