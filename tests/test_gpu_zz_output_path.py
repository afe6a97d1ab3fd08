"""GPU: the output path of SURVEY.md section 8f-3 -- planes / lines gathered on the device
(mifgpu_tensor_download_box) instead of whole-field downloads.  The box download is checked bit for bit against
slices of the whole-field download, and the host layer's writers (which prefetch boxes and read through peek) are
checked to write byte-identical files with and without the box path after device-side time steps.

(File name sorts last on purpose: these tests were added after the last GPU session of round 1.)"""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

HOST_BIN = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin")


@pytest.mark.parametrize("N,periodic", [((20, 13, 11), (False, False, False)), ((33, 9, 18), (False, False, True)),
                                        ((12, 40, 7), (True, False, False))])
def test_download_box_equals_slices_of_the_whole_field(mif, N, periodic):
    ctx = mif.Context(N[0], N[1], N[2], 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-3, 4, periodic=periodic)
    rng = np.random.default_rng(7)
    for staggering in (mif.STAGGER_X, mif.STAGGER_Y, mif.STAGGER_Z, mif.STAGGER_NONE):
        t = ctx.tensor(staggering)
        sx, sy, sz = t.shape
        full = rng.uniform(-1, 1, (sz, sy, sx))
        t.upload(full)
        assert np.array_equal(t.download(), full)
        boxes = [((0, 0, 0), (sx, sy, sz)),                    # everything
                 ((0, 0, sz // 2), (sx, sy, sz // 2 + 1)),      # a z plane
                 ((0, sy // 2, 0), (sx, sy // 2 + 2, sz)),      # two y planes
                 ((sx - 1, 0, 0), (sx, sy, sz)),                # the last x plane (stride-PX gather)
                 ((1, 2, 0), (3, 4, sz)),                       # a 2 x 2 line bundle along z
                 ((0, sy - 1, sz - 1), (sx, sy, sz)),           # one line along x
                 ((sx - 1, sy - 1, sz - 1), (sx, sy, sz))]      # one value
        for _ in range(5):
            lo = [int(rng.integers(0, n)) for n in (sx, sy, sz)]
            hi = [int(rng.integers(l + 1, n + 1)) for l, n in zip(lo, (sx, sy, sz))]
            boxes.append((tuple(lo), tuple(hi)))
        for lo, hi in boxes:
            got = t.download_box(lo, hi)
            assert np.array_equal(got, full[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]]), (staggering, lo, hi)
        for lo, hi in [((0, 0, 0), (sx + 1, sy, sz)), ((0, 0, 0), (sx, sy, sz + 1)), ((-1, 0, 0), (sx, sy, sz))]:
            with pytest.raises(mif.MifGpuError, match="-1"):
                t.download_box(lo, hi)
        t.close()
    ctx.close()


@pytest.mark.parametrize("case", [1, 2])
def test_mif_driver_writes_identical_files_with_and_without_the_box_path(case, tmp_path):
    """The ported driver (host/apps/mif.cpp) after real device-side time steps: solution.vtk and the profiles written
    from device-gathered planes / lines equal those written after whole-field downloads (MIF_EXPORT_WHOLE_FIELDS)."""
    exe = os.path.join(HOST_BIN, "mif")
    text = "Nt : 3\ndt : 1e-3\nNx : 24\nNy : 20\nNz : 28\nPy : 1\nPz : 1\ntest_case_2 : %s\n" % ("true" if case == 2 else "false")
    outputs = {}
    for label, extra in (("box", {}), ("whole", {"MIF_EXPORT_WHOLE_FIELDS": "1"})):
        work = tmp_path / label
        work.mkdir()
        (work / "input.txt").write_text(text)
        env = dict(os.environ, **extra)
        subprocess.run([exe, "input.txt"], cwd=work, env=env, check=True, timeout=300, capture_output=True)
        outputs[label] = work
    names = ["solution.vtk", "profile1.dat", "profile2.dat"] + (["profile3.dat"] if case == 2 else [])
    for name in names:
        a, b = outputs["box"] / name, outputs["whole"] / name
        assert a.exists() and b.exists(), name
        assert filecmp.cmp(a, b, shallow=False), name
