"""TEST INFRASTRUCTURE ONLY: the subset of the `mif_b200` Python API that the full-size GPU tests use, served by the CPU
oracle.  tests/test_properties_cpu.py runs the bodies of tests/test_gpu_zz_full_size.py against it at a small size, so
the test code itself (shapes, index maps, tolerances) is exercised without a GPU.  Never imported by the product."""
import os
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mif_oracle as mo  # noqa: E402

STAGGER_X, STAGGER_Y, STAGGER_Z, STAGGER_NONE = 0, 1, 2, 3
BC_TEST_CASE_1, BC_TEST_CASE_2, BC_ETHIER_STEINMAN = mo.BC_TEST_CASE_1, mo.BC_TEST_CASE_2, mo.BC_ETHIER_STEINMAN


class Tensor:
    def __init__(self, grid, staggering):
        self.data = grid.zeros(staggering)
        sz, sy, sx = self.data.shape
        self.shape = (sx, sy, sz)

    def upload(self, host):
        self.data[...] = np.asarray(host, dtype=np.float64).reshape(self.data.shape)

    def download(self):
        return self.data.copy()

    def download_box(self, lo, hi):
        return self.data[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]].copy()


class Bc:
    def __init__(self, kind):
        self.kind = kind


class Context:
    def __init__(self, Nx, Ny, Nz, x_size, y_size, z_size, min_x, min_y, min_z, Re, final_time, num_time_steps,
                 periodic=(False, False, False)):
        self.grid = mo.Grid(Nx, Ny, Nz, x_size, y_size, z_size, min_x, min_y, min_z, Re, final_time, num_time_steps,
                            periodic=periodic)
        self.dt = final_time / num_time_steps

    def tensor(self, staggering):
        return Tensor(self.grid, staggering)

    def velocity(self):
        return [Tensor(self.grid, s) for s in (0, 1, 2)]

    def make_bc(self, kind, Re=1.0):
        return Bc(kind)

    def solve_pressure(self, p, vel, dt):
        p.data[...] = self.grid.solve_pressure(*[t.data for t in vel], dt)

    def apply_bc(self, vel, bc, time):
        self.grid.apply_bc(bc.kind, time, *[t.data for t in vel])

    def timestep(self, vel, vb, vb2, bc, t_n, p, dp):
        self.grid.timestep(bc.kind, t_n, [t.data for t in vel], [t.data for t in vb], [t.data for t in vb2], p.data, dp.data)

    def close(self):
        pass
