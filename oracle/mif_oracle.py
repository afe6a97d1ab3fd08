"""oracle/mif_oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes view of oracle/libmif_oracle.so (mif_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
path (mif_b200 / libmifgpu.so) never does.  Arrays are numpy float64, shaped (sz, sy, sx), i.e. the
reference layout i + j*sx + k*sx*sy.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_long, c_uint, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmif_oracle.so")

BC_TEST_CASE_1, BC_TEST_CASE_2, BC_ETHIER_STEINMAN = 1, 2, 3
BC_VELOCITY_TEST = 5  # generators/manufsol_velocity.py (4 is the host-callback kind of include/mifgpu.h)
REDFT00, R2HC, HC2R = 0, 1, 2

_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = ctypes.CDLL(LIB_PATH)
        l.mo_grid_size.restype = c_int
        l.mo_grid_init.argtypes = [c_void_p, c_long, c_long, c_long] + [c_double] * 8 + [c_uint, POINTER(c_int)]
        l.mo_extents.argtypes = [c_void_p, c_int, POINTER(c_long)]
        dp = POINTER(c_double)
        l.mo_set_velocity.argtypes = [c_void_p, c_int, c_double, dp, dp, dp]
        l.mo_apply_bc.argtypes = [c_void_p, c_int, c_double, dp, dp, dp]
        l.mo_solve_pressure.argtypes = [c_void_p, dp, dp, dp, dp, c_double, POINTER(dp), c_int]
        l.mo_timestep.argtypes = [c_void_p, c_int, c_double] + [dp] * 11 + [c_int, c_int]
        l.mo_transform.argtypes = [c_int, c_int, c_int, dp, dp]
        l.mo_exact_velocity.argtypes = [c_int, c_int] + [c_double] * 5
        l.mo_exact_velocity.restype = c_double
        l.mo_es_pressure_gradient.argtypes = [c_int] + [c_double] * 5
        l.mo_es_pressure_gradient.restype = c_double
        l.mo_forcing.argtypes = [c_int] + [c_double] * 5
        l.mo_forcing.restype = c_double
        l.mo_timestep_velocity.argtypes = [c_void_p, c_int, c_double] + [dp] * 9
        l.mo_exact_pressure.argtypes = [c_int] + [c_double] * 5
        l.mo_exact_pressure.restype = c_double
        l.mo_velocity_error_norms.argtypes = [c_void_p, c_int, c_double, dp, dp, dp, dp]
        l.mo_pressure_error_norms.argtypes = [c_void_p, c_int, c_double, dp, dp]
        l.mo_adjust_pressure.argtypes = [c_void_p, c_int, c_double, dp]
        _lib = l
    return _lib


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(POINTER(c_double))


class Grid:
    """mif::Constants for one rank (src/Constants.cpp:58-101)."""

    def __init__(self, Nx, Ny, Nz, x_size, y_size, z_size, min_x, min_y, min_z, Re, final_time, num_time_steps,
                 periodic=(False, False, False)):
        l = lib()
        self._buf = ctypes.create_string_buffer(l.mo_grid_size())
        per = (c_int * 3)(*[int(b) for b in periodic])
        l.mo_grid_init(self._buf, Nx, Ny, Nz, x_size, y_size, z_size, min_x, min_y, min_z, Re, final_time,
                       num_time_steps, per)
        self.dt = final_time / num_time_steps
        self.Re = Re

    def shape(self, stag):
        """numpy shape (sz, sy, sx) of a tensor staggered in direction stag (0, 1, 2) or unstaggered (3)."""
        ext = (c_long * 3)()
        lib().mo_extents(self._buf, stag, ext)
        return int(ext[2]), int(ext[1]), int(ext[0])

    def zeros(self, stag):
        return np.zeros(self.shape(stag), dtype=np.float64)

    def set_velocity(self, kind, t):
        u, v, w = self.zeros(0), self.zeros(1), self.zeros(2)
        lib().mo_set_velocity(self._buf, kind, t, _ptr(u), _ptr(v), _ptr(w))
        return u, v, w

    def apply_bc(self, kind, t, u, v, w):
        lib().mo_apply_bc(self._buf, kind, t, _ptr(u), _ptr(v), _ptr(w))

    def solve_pressure(self, u, v, w, dt, nhn_faces=None, direct=False):
        p = self.zeros(3)
        faces = None
        if nhn_faces is not None:
            keep = [np.ascontiguousarray(f, dtype=np.float64) for f in nhn_faces]
            faces = (POINTER(c_double) * 6)(*[_ptr(f) for f in keep])
        lib().mo_solve_pressure(self._buf, _ptr(p), _ptr(u), _ptr(v), _ptr(w), dt, faces, int(direct))
        return p

    def timestep_velocity(self, kind, t_n, vel, buf, rhs_buf):
        """mif::timestep_velocity (src/TimestepVelocity.cpp:58-90), in place."""
        lib().mo_timestep_velocity(self._buf, kind, t_n, *[_ptr(a) for a in (*vel, *buf, *rhs_buf)])

    def velocity_error_norms(self, kind, t, u, v, w):
        """(L1, L2, LInf) of src/Norms.cpp:49-86."""
        out = np.zeros(3)
        lib().mo_velocity_error_norms(self._buf, kind, t, _ptr(u), _ptr(v), _ptr(w), _ptr(out))
        return tuple(out)

    def pressure_error_norms(self, kind, t, p):
        """(L1, L2, LInf) of src/Norms.cpp:103-118."""
        out = np.zeros(3)
        lib().mo_pressure_error_norms(self._buf, kind, t, _ptr(p), _ptr(out))
        return tuple(out)

    def adjust_pressure(self, kind, t, p):
        """src/PressureEquation.cpp:288-343, in place."""
        lib().mo_adjust_pressure(self._buf, kind, t, _ptr(p))

    def timestep(self, kind, t_n, vel, buf, buf2, p, dp, nhn=False, direct=False):
        args = [_ptr(a) for a in (*vel, *buf, *buf2, p, dp)]
        lib().mo_timestep(self._buf, kind, t_n, *args, int(nhn), int(direct))


def transform(kind, x, direct=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().mo_transform(kind, x.size, int(direct), _ptr(x), _ptr(out))
    return out
