"""CPU: the kernel SOURCES of libmifgpu executed by the SIMT interpreter of tests/simt_emu (every CUDA thread a fiber;
see tests/simt_emu/README.md) on a small selection of the GPU parity cases.  This is a development aid that catches
indexing / arithmetic mistakes in a kernel on a machine without a GPU; the parity claims themselves rest on the
`-m gpu` tests on a B200, not on this."""
import os
import subprocess
import sys

from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "simt_emu")


def test_kernel_sources_pass_parity_cases_under_the_simt_interpreter():
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    env = dict(os.environ, MIFGPU_LIB=os.path.join(EMU, "build", "libmifgpu_simt.so"))
    run = subprocess.run([sys.executable, os.path.join(EMU, "run_cases.py")], env=env, capture_output=True, text=True,
                         timeout=900)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    assert "simt cases ok" in run.stdout
