"""CPU: the single-precision build of the library (libmifgpu_f32.so = the reference's USE_DOUBLE=0 build,
include/Real.h:9-17) -- its ABI, and its kernels executed by the SIMT interpreter against the float reference's goldens
and the FP64 oracle (tests/fp32_cases.py; the hardware run of the same cases is tests/test_gpu_zzz_fp32.py)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

from conftest import GOLDEN_DIR, ROOT

EMU = os.path.join(ROOT, "tests", "simt_emu")
PKG = os.path.join(ROOT, "mpi-incompressible-fluid_b200")


def test_float_library_exports_the_same_abi_with_four_byte_reals():
    import mif_b200 as mif
    path = os.path.join(PKG, "libmifgpu_f32.so")
    assert os.path.exists(path), "libmifgpu_f32.so missing: __graft_entry__.build() makes it"
    lib = ctypes.CDLL(path)
    for name in mif.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name
    assert lib.mifgpu_real_bytes() == 4
    assert lib.mifgpu_abi_version() == ctypes.CDLL(os.path.join(PKG, "libmifgpu.so")).mifgpu_abi_version()
    assert ctypes.CDLL(os.path.join(PKG, "libmifgpu.so")).mifgpu_real_bytes() == 8


def test_float_goldens_come_from_the_reference_float_build():
    from conftest import load_golden
    for case in ("f32_full_16_2", "f32_lid1_12x10x14_2", "f32_lid2_10x12x9_2"):
        meta, f = load_golden(case)
        assert meta["real"] == "float32" and all(f[k].dtype == np.float32 for k in f if k != "meta")
    # the float and the double build of the reference print the same norms to 3-4 digits (tests/golden/norms.json)
    f32 = json.load(open(os.path.join(GOLDEN_DIR, "f32_norms.json")))
    f64 = json.load(open(os.path.join(GOLDEN_DIR, "norms.json")))
    for key, values in f32.items():
        for a, b in zip(values, f64[key]):
            assert abs(a - b) <= 2e-3 * abs(b), (key, a, b)


def test_float_kernels_pass_the_parity_cases_under_the_simt_interpreter():
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    env = dict(os.environ, MIFGPU_LIB=os.path.join(EMU, "build", "libmifgpu_simt_f32.so"))
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fp32_cases.py"), "--quick"], env=env, capture_output=True,
                         text=True, timeout=1500)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-3000:]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fp32_cases
    out = json.loads([l for l in run.stdout.splitlines() if l.startswith("{")][-1])
    assert len(out) >= 11
    fp32_cases.check(out)


def test_float_library_on_two_ranks_matches_one_rank():
    """World size 2 under torchrun: plane halos and the pack -> all-to-all -> unpack transposes move floats
    (ncclFloat); the slabs reproduce the single-rank run of the float REFERENCE (lid-driven case, two steps)."""
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    env = dict(os.environ, MIFGPU_LIB=os.path.join(EMU, "build", "libmifgpu_simt_f32.so"),
               MIFGPU_NCCL_LIB=os.path.join(EMU, "build", "libmif_fake_nccl.so"), MIF_SIMT_IPC="1", MIF_PY="1", MIF_WORKER_TOL="2e-5")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "mp_worker.py"), "f32_lid1_12x10x14_2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["world"] == 2 and res["max_rel_err"] <= 2e-5, res


def test_float_host_layer_prints_the_float_reference_numbers():
    """host/bin_f32: the ported full_test and the reference's own test/full_test.cpp compiled UNCHANGED, both with
    -DUSE_DOUBLE=0 (Real = float, include/Real.h:9-17) against the float host layer and libmifgpu_f32.so; the nine numbers
    are those of the float build of the reference (tests/golden/f32_norms.json; the error norms are sums of float
    round-off sized terms, hence 2e-3)."""
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    want = json.load(open(os.path.join(GOLDEN_DIR, "f32_norms.json")))["full_test 16 1 1"]
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(EMU, "build", "as_libmifgpu"))
    ran = 0
    for name in ("full_test", "ref_full_test"):
        exe = os.path.join(PKG, "host", "bin_f32", name)
        if name == "ref_full_test" and not os.path.exists(exe):
            continue  # built only where the reference tree is present
        out = subprocess.run([exe, "16", "1", "1"], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
        got = [float(x) for x in out.stdout.split()]
        assert len(got) == 9
        for a, b in zip(got, want):
            assert abs(a - b) <= 2e-3 * abs(b), (name, got, want)
        ran += 1
    assert ran >= 1
