"""CPU: the transforms of oracle/fft_cpu.h (used by the oracle and by the FFTW stand-in of oracle/_ref) against
independent implementations: scipy.fft.dct(type=1) for FFTW_REDFT00 and numpy.fft.rfft for R2HC / HC2R."""
import os
import sys

import numpy as np
import pytest
import scipy.fft

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mif_oracle as mo  # noqa: E402

SIZES = [2, 3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 63, 64, 65, 100, 129, 480, 512, 513]


def halfcomplex(x):
    n = x.size
    spec = np.fft.rfft(x)
    out = np.empty(n)
    out[: n // 2 + 1] = spec.real
    for k in range(1, (n + 1) // 2):
        out[n - k] = spec.imag[k]
    return out


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("direct", [False, True])
def test_redft00(n, direct):
    x = np.random.default_rng(n).uniform(-1, 1, n)
    ref = scipy.fft.dct(x, type=1)
    got = mo.transform(mo.REDFT00, x, direct)
    assert np.max(np.abs(got - ref)) <= 1e-13 * n * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("direct", [False, True])
def test_r2hc_and_hc2r(n, direct):
    x = np.random.default_rng(1000 + n).uniform(-1, 1, n)
    hc = mo.transform(mo.R2HC, x, direct)
    assert np.max(np.abs(hc - halfcomplex(x))) <= 1e-13 * n
    back = mo.transform(mo.HC2R, hc, direct)
    assert np.max(np.abs(back - n * x)) <= 1e-12 * n
