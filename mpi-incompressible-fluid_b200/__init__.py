"""ctypes binding of libmifgpu (include/mifgpu.h) used by tests/, bench.py and __graft_entry__.py.

The product is the C-ABI shared library `libmifgpu.so` (CUDA, sm_100a) plus the C++ host layer in
`host/` that mirrors the reference's classes; this module is only the thinnest possible Python view of
the same entry points (numpy arrays in the reference's layout in, numpy arrays out).  It never falls
back to a CPU implementation: if the library or a CUDA device is missing, calls raise.

The directory name contains a hyphen, so import it through the `mif_b200` symlink at the repository
root (`import mif_b200`).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int32, c_uint32, c_uint64, c_void_p
from typing import Callable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("MIFGPU_LIB", "libmifgpu.so"))

STAGGER_X, STAGGER_Y, STAGGER_Z, STAGGER_NONE = 0, 1, 2, 3
BC_TEST_CASE_1, BC_TEST_CASE_2, BC_ETHIER_STEINMAN, BC_HOST_CALLBACK, BC_VELOCITY_TEST = 1, 2, 3, 4, 5

# Every symbol include/mifgpu.h declares (checked by tests/test_abi.py without a GPU).
EXPORTED_SYMBOLS = [
    "mifgpu_abi_version", "mifgpu_last_error", "mifgpu_create", "mifgpu_destroy", "mifgpu_tensor_extents",
    "mifgpu_tensor_create", "mifgpu_tensor_destroy", "mifgpu_tensor_upload", "mifgpu_tensor_download",
    "mifgpu_tensor_swap", "mifgpu_timestep", "mifgpu_apply_bc", "mifgpu_solve_pressure", "mifgpu_synchronize",
    "mifgpu_stream", "mifgpu_launch_count", "mifgpu_profile_enable", "mifgpu_profile_read", "mifgpu_comm_unique_id",
    "mifgpu_create_distributed", "mifgpu_slab_plan", "mifgpu_velocity_error_norms", "mifgpu_pressure_error_norms",
    "mifgpu_adjust_pressure", "mifgpu_timestep_velocity", "mifgpu_tensor_download_box",
    "mifgpu_allreduce", "mifgpu_gather", "mifgpu_rank_count", "mifgpu_tensor_upload_async", "mifgpu_tensor_download_async",
    "mifgpu_transpose_path", "mifgpu_real_bytes",
]


class MifGpuError(RuntimeError):
    pass


class Params(ctypes.Structure):
    """mifgpu_params: the 16 constructor arguments of mif::Constants (include/Constants.h:86-90) + device."""
    _fields_ = [
        ("Nx_global", c_uint64), ("Ny_global", c_uint64), ("Nz_global", c_uint64),
        ("x_size", c_double), ("y_size_global", c_double), ("z_size_global", c_double),
        ("min_x_global", c_double), ("min_y_global", c_double), ("min_z_global", c_double),
        ("Re", c_double), ("final_time", c_double), ("num_time_steps", c_uint32),
        ("Py", c_int32), ("Pz", c_int32), ("rank", c_int32),
        ("periodic_bc", c_int32 * 3), ("device", c_int32),
    ]


FACE_CALLBACK = ctypes.CFUNCTYPE(None, c_void_p, c_int, c_double, c_double, c_int, c_int, c_void_p)  # values: mifgpu_real *


class Bc(ctypes.Structure):
    """mifgpu_bc"""
    _fields_ = [("kind", c_int32), ("Re", c_double), ("callback", FACE_CALLBACK), ("user", c_void_p)]


def build(force: bool = False) -> str:
    """Compile libmifgpu.so in-tree with nvcc for sm_100a (no GPU needed)."""
    args = ["make", "-C", _HERE, "-j4"]
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    proc = subprocess.run(args, capture_output=True, text=True)
    if proc.returncode != 0:
        raise MifGpuError("building libmifgpu.so failed:\n" + proc.stdout + proc.stderr)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load libmifgpu.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MifGpuError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    l = ctypes.CDLL(LIB_PATH)
    l.mifgpu_abi_version.restype = c_int
    l.mifgpu_real_bytes.restype = c_int
    l.mifgpu_last_error.restype = ctypes.c_char_p
    l.mifgpu_create.argtypes = [POINTER(Params), POINTER(c_void_p)]
    l.mifgpu_destroy.argtypes = [c_void_p]
    l.mifgpu_destroy.restype = None
    l.mifgpu_comm_unique_id.argtypes = [c_void_p]
    l.mifgpu_create_distributed.argtypes = [POINTER(Params), c_void_p, POINTER(c_void_p)]
    l.mifgpu_slab_plan.argtypes = [c_uint64, c_int32, POINTER(c_int32)]
    l.mifgpu_tensor_extents.argtypes = [c_void_p, c_int, POINTER(c_uint64)]
    l.mifgpu_tensor_create.argtypes = [c_void_p, c_int, POINTER(c_void_p)]
    l.mifgpu_tensor_destroy.argtypes = [c_void_p]
    l.mifgpu_tensor_destroy.restype = None
    l.mifgpu_tensor_upload.argtypes = [c_void_p, c_void_p]
    l.mifgpu_tensor_download.argtypes = [c_void_p, c_void_p]
    l.mifgpu_tensor_swap.argtypes = [c_void_p, c_void_p]
    l.mifgpu_tensor_upload_async.argtypes = [c_void_p, c_void_p]
    l.mifgpu_tensor_download_async.argtypes = [c_void_p, c_void_p]
    l.mifgpu_transpose_path.argtypes = [c_void_p]
    l.mifgpu_tensor_download_box.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32), c_void_p]
    l.mifgpu_timestep.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(Bc),
                                  c_double, c_void_p, c_void_p, c_int]
    l.mifgpu_apply_bc.argtypes = [c_void_p, POINTER(c_void_p), POINTER(Bc), c_double]
    l.mifgpu_timestep_velocity.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(Bc), c_double]
    l.mifgpu_solve_pressure.argtypes = [c_void_p, c_void_p, POINTER(c_void_p), c_double, POINTER(Bc), c_double]
    l.mifgpu_synchronize.argtypes = [c_void_p]
    l.mifgpu_velocity_error_norms.argtypes = [c_void_p, POINTER(c_void_p), POINTER(Bc), c_double, POINTER(c_double)]
    l.mifgpu_pressure_error_norms.argtypes = [c_void_p, c_void_p, POINTER(Bc), c_double, POINTER(c_double)]
    l.mifgpu_adjust_pressure.argtypes = [c_void_p, c_void_p, POINTER(Bc), c_double]
    l.mifgpu_allreduce.argtypes = [c_void_p, POINTER(c_double), c_int32, c_int32]
    l.mifgpu_gather.argtypes = [c_void_p, POINTER(c_double), c_uint64, POINTER(c_double), POINTER(c_uint64)]
    l.mifgpu_rank_count.argtypes = [c_void_p]
    l.mifgpu_stream.argtypes = [c_void_p]
    l.mifgpu_stream.restype = c_void_p
    l.mifgpu_launch_count.argtypes = [c_void_p]
    l.mifgpu_launch_count.restype = c_uint64
    l.mifgpu_profile_enable.argtypes = [c_void_p, c_int]
    l.mifgpu_profile_read.argtypes = [c_void_p, c_int, POINTER(ctypes.c_char_p), POINTER(c_double), POINTER(c_uint64)]
    _lib = l
    return l


def real_dtype():
    """numpy dtype of mifgpu_real in the loaded library: float64 for libmifgpu.so, float32 for libmifgpu_f32.so (the
    reference's USE_DOUBLE=0 build, selected with MIFGPU_LIB=libmifgpu_f32.so)."""
    return np.float64 if lib().mifgpu_real_bytes() == 8 else np.float32


def _check(rc: int) -> None:
    if rc != 0:
        raise MifGpuError(f"libmifgpu error {rc}: {lib().mifgpu_last_error().decode()}")


UNIQUE_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """Rank 0: a fresh communicator id to hand to every rank's Context(comm_id=...) (any transport will do)."""
    buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
    _check(lib().mifgpu_comm_unique_id(buf))
    return buf.raw


def slab_plan(n_points: int, parts: int):
    """first[r]..first[r+1]: block distribution of n_points over `parts` ranks (host-only, no GPU needed)."""
    first = (c_int32 * (parts + 1))()
    _check(lib().mifgpu_slab_plan(n_points, parts, first))
    return [int(v) for v in first]


class Tensor:
    """Device-resident StaggeredTensor (src/StaggeredTensor.cpp:5-36)."""

    def __init__(self, ctx: "Context", staggering: int):
        self.ctx = ctx
        self.staggering = staggering
        self.shape = ctx.extents(staggering)  # (sx, sy, sz), x fastest
        handle = c_void_p()
        _check(lib().mifgpu_tensor_create(ctx.handle, staggering, ctypes.byref(handle)))
        self.handle = handle

    def upload(self, host: np.ndarray) -> None:
        """host: array of shape (sz, sy, sx) C-order, i.e. the reference layout i + j*sx + k*sx*sy."""
        sx, sy, sz = self.shape
        arr = np.ascontiguousarray(host, dtype=real_dtype())
        if arr.size != sx * sy * sz:
            raise ValueError(f"expected {sx * sy * sz} values, got {arr.size}")
        _check(lib().mifgpu_tensor_upload(self.handle, arr.ctypes.data_as(c_void_p)))

    def _check_host(self, arr: np.ndarray, what: str) -> None:
        sx, sy, sz = self.shape
        if not isinstance(arr, np.ndarray) or arr.dtype != real_dtype() or arr.size != sx * sy * sz or not arr.flags.c_contiguous:
            raise ValueError(f"{what} must be a C-contiguous {np.dtype(real_dtype()).name} array of {sx * sy * sz} values "
                             f"(shape ({sz}, {sy}, {sx}))")

    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        sx, sy, sz = self.shape
        if out is None:
            out = np.empty((sz, sy, sx), dtype=real_dtype())
        self._check_host(out, "out")
        _check(lib().mifgpu_tensor_download(self.handle, out.ctypes.data_as(c_void_p)))
        return out

    def upload_async(self, host: np.ndarray) -> None:
        """Enqueue the copy on the context's upload stream; `host` (page-locked) must stay alive and untouched until
        Context.synchronize()."""
        self._check_host(host, "host")
        _check(lib().mifgpu_tensor_upload_async(self.handle, host.ctypes.data_as(c_void_p)))

    def download_async(self, out: np.ndarray) -> None:
        """Enqueue the copy on the context's download stream; `out` holds the values after Context.synchronize()."""
        self._check_host(out, "out")
        _check(lib().mifgpu_tensor_download_async(self.handle, out.ctypes.data_as(c_void_p)))

    def download_box(self, lo: Sequence[int], hi: Sequence[int]) -> np.ndarray:
        """The index box lo <= (i, j, k) < hi as an array of shape (hi[2]-lo[2], hi[1]-lo[1], hi[0]-lo[0])."""
        out = np.empty((hi[2] - lo[2], hi[1] - lo[1], hi[0] - lo[0]), dtype=real_dtype())
        _check(lib().mifgpu_tensor_download_box(self.handle, (c_int32 * 3)(*lo), (c_int32 * 3)(*hi),
                                                out.ctypes.data_as(c_void_p)))
        return out

    def close(self) -> None:
        if self.handle:
            lib().mifgpu_tensor_destroy(self.handle)
            self.handle = None


class Context:
    """mif::Constants + mif::PressureSolverStructures on one GPU."""

    def __init__(self, Nx: int, Ny: int, Nz: int, x_size: float, y_size: float, z_size: float, min_x: float,
                 min_y: float, min_z: float, Re: float, final_time: float, num_time_steps: int, Py: int = 1,
                 Pz: int = 1, rank: int = 0, periodic: Sequence[bool] = (False, False, False), device: int = 0,
                 comm_id: Optional[bytes] = None):
        self.params = Params(Nx, Ny, Nz, x_size, y_size, z_size, min_x, min_y, min_z, Re, final_time,
                             num_time_steps, Py, Pz, rank, (c_int32 * 3)(*[int(b) for b in periodic]), device)
        handle = c_void_p()
        if comm_id is None:
            _check(lib().mifgpu_create(ctypes.byref(self.params), ctypes.byref(handle)))
        else:
            if len(comm_id) != UNIQUE_ID_BYTES:
                raise ValueError("comm_id must be the 128 bytes returned by comm_unique_id() on rank 0")
            _check(lib().mifgpu_create_distributed(ctypes.byref(self.params), ctypes.c_char_p(comm_id), ctypes.byref(handle)))
        self.handle = handle
        self.dt = final_time / num_time_steps
        self._keepalive = []

    def extents(self, staggering: int):
        ext = (c_uint64 * 3)()
        _check(lib().mifgpu_tensor_extents(self.handle, staggering, ext))
        return int(ext[0]), int(ext[1]), int(ext[2])

    def tensor(self, staggering: int) -> Tensor:
        return Tensor(self, staggering)

    def velocity(self):
        return [Tensor(self, s) for s in (STAGGER_X, STAGGER_Y, STAGGER_Z)]

    @staticmethod
    def _triple(tensors):
        return (c_void_p * 3)(*[t.handle for t in tensors])

    def make_bc(self, kind: int, Re: float = 1.0, callback: Optional[Callable] = None) -> Bc:
        """callback(which, time, time_prev, component, face, values: np.ndarray) fills `values` in place."""
        if callback is None:
            return Bc(kind, Re, FACE_CALLBACK(), None)
        ctx = self

        def trampoline(_user, which, time, time_prev, comp, face, ptr):
            t = 3 if which == 1 else comp
            sx, sy, sz = ctx.extents(t)
            d = 2 - face // 2
            na = sy if d == 0 else sx
            nb = sy if d == 2 else sz
            ctype = c_double if real_dtype() == np.float64 else ctypes.c_float
            values = np.ctypeslib.as_array((ctype * (na * nb)).from_address(ptr)).reshape(nb, na)
            callback(which, time, time_prev, comp, face, values)

        cb = FACE_CALLBACK(trampoline)
        self._keepalive.append(cb)
        return Bc(kind, Re, cb, None)

    def timestep(self, vel, vel_buf, vel_buf2, bc: Bc, t_n: float, pressure: Tensor, pressure_buffer: Tensor,
                 nhn: bool = False) -> None:
        _check(lib().mifgpu_timestep(self.handle, self._triple(vel), self._triple(vel_buf), self._triple(vel_buf2),
                                     ctypes.byref(bc), t_n, pressure.handle, pressure_buffer.handle, int(nhn)))

    def timestep_velocity(self, vel, vel_buf, rhs_buf, bc: Bc, t_n: float) -> None:
        """mif::timestep_velocity; the device data of vel and vel_buf are swapped inside the call, as in the reference."""
        _check(lib().mifgpu_timestep_velocity(self.handle, self._triple(vel), self._triple(vel_buf), self._triple(rhs_buf),
                                              ctypes.byref(bc), t_n))

    def apply_bc(self, vel, bc: Bc, time: float) -> None:
        _check(lib().mifgpu_apply_bc(self.handle, self._triple(vel), ctypes.byref(bc), time))

    def solve_pressure(self, pressure: Tensor, vel, dt: float, nhn_bc: Optional[Bc] = None,
                       nhn_time: float = 0.0) -> None:
        bc_ptr = ctypes.byref(nhn_bc) if nhn_bc is not None else None
        _check(lib().mifgpu_solve_pressure(self.handle, pressure.handle, self._triple(vel), dt, bc_ptr, nhn_time))

    def velocity_error_norms(self, vel, exact: Bc, time: float):
        """(L1, L2, LInf) of src/Norms.cpp:49-86 for this rank, computed on the device."""
        out = (c_double * 3)()
        _check(lib().mifgpu_velocity_error_norms(self.handle, self._triple(vel), ctypes.byref(exact), time, out))
        return float(out[0]), float(out[1]), float(out[2])

    def pressure_error_norms(self, pressure: Tensor, exact: Bc, time: float):
        """(L1, L2, LInf) of src/Norms.cpp:103-118 for this rank, computed on the device."""
        out = (c_double * 3)()
        _check(lib().mifgpu_pressure_error_norms(self.handle, pressure.handle, ctypes.byref(exact), time, out))
        return float(out[0]), float(out[1]), float(out[2])

    def adjust_pressure(self, pressure: Tensor, exact: Bc, time: float) -> None:
        _check(lib().mifgpu_adjust_pressure(self.handle, pressure.handle, ctypes.byref(exact), time))

    def synchronize(self) -> None:
        _check(lib().mifgpu_synchronize(self.handle))

    def profile_enable(self, enable: bool) -> None:
        _check(lib().mifgpu_profile_enable(self.handle, int(enable)))

    def profile_read(self):
        """{category: (milliseconds, timed kernel groups)} accumulated since the previous read."""
        cap = 32
        names = (ctypes.c_char_p * cap)()
        ms = (c_double * cap)()
        counts = (c_uint64 * cap)()
        n = lib().mifgpu_profile_read(self.handle, cap, names, ms, counts)
        if n < 0:
            _check(n)
        return {names[i].decode(): (float(ms[i]), int(counts[i])) for i in range(n)}

    @property
    def stream(self) -> int:
        return int(lib().mifgpu_stream(self.handle) or 0)

    @property
    def transpose_path(self) -> int:
        """0 none (one rank), 1 peer-memory fused, 2 NCCL all-to-all, 3 pencil box exchanges (include/mifgpu.h)."""
        return int(lib().mifgpu_transpose_path(self.handle))

    @property
    def launch_count(self) -> int:
        return int(lib().mifgpu_launch_count(self.handle))

    def close(self) -> None:
        if self.handle:
            lib().mifgpu_destroy(self.handle)
            self.handle = None
