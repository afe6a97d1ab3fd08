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
