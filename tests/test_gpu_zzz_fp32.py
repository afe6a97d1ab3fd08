"""GPU: the single-precision build (libmifgpu_f32.so, the reference's USE_DOUBLE=0 build) on the device, through the C
ABI, against the float reference's goldens and the FP64 oracle -- the cases and tolerances of tests/fp32_cases.py, run in
a subprocess because a process binds one build of the library.  Named to sort last: the FP64 parity suite comes first."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_float_library_passes_the_parity_cases_on_the_device():
    env = dict(os.environ, MIFGPU_LIB="libmifgpu_f32.so")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fp32_cases.py")], env=env, capture_output=True, text=True,
                         timeout=900)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-3000:]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fp32_cases
    out = json.loads([l for l in run.stdout.splitlines() if l.startswith("{")][-1])
    assert len(out) >= 16
    fp32_cases.check(out)


def test_float_host_layer_prints_the_float_reference_numbers_on_the_device():
    """host/bin_f32/full_test and the reference's unchanged full_test.cpp (-DUSE_DOUBLE=0) on libmifgpu_f32.so."""
    from conftest import GOLDEN_DIR
    want = json.load(open(os.path.join(GOLDEN_DIR, "f32_norms.json")))
    for name in ("full_test", "ref_full_test"):
        exe = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin_f32", name)
        if name == "ref_full_test" and not os.path.exists(exe):
            continue
        for args in (("16", "1", "1"), ("32", "2", "1")):
            out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
            assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
            got = [float(x) for x in out.stdout.split()]
            assert len(got) == 9
            for a, b in zip(got, want["full_test " + " ".join(args)]):
                assert abs(a - b) <= 5e-3 * abs(b), (name, args, got)  # error norms of ~4e-4 carry float round-off of ~3e-7
