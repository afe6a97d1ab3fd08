"""GPU, >= 2 devices: the parts of the multi-GPU path that were added after the last GPU session of round 1 and have so
far only run on emulated ranks (tests/test_multi_cpu_simt.py): the Py x Pz pencil decomposition and the ported drivers
started with scripts/mifrun.  (File name sorts last on purpose.)"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def failure_report(out):
    """The ranks' own tracebacks / library errors first (torchrun's summary hides them at the end of a long stderr)."""
    own = [l for l in out.stderr.splitlines() if l.startswith("[rank") or "libmifgpu" in l or "Error" in l]
    return "\n".join(own[-60:]) + "\n--- stdout ---\n" + out.stdout[-1500:] + "\n--- stderr tail ---\n" + out.stderr[-1500:]


def device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("case", ["full_17_1", "lid1_12x10x14_2", "es:9x257x257"])
@pytest.mark.parametrize("world,py", [(2, 2), (4, 2), (8, 2), (8, 4)])
def test_pencil_decomposition_matches_single_rank_reference(case, world, py):
    """Py x Pz pencils (the reference's decomposition, src/Constants.cpp:68-101): two-phase halos and the four
    2Decomp transposes as box exchanges; every rank's block reproduces the single-rank goldens."""
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + world + py), os.path.join(ROOT, "tests", "mp_worker.py"), case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MIF_PY=str(py)))
    assert out.returncode == 0, failure_report(out)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["Py"] == py and res["max_rel_err"] <= 1e-11


@pytest.mark.parametrize("py,pz", [(1, 2), (2, 2), (2, 4)])
def test_ported_driver_on_several_gpus_writes_the_reference_files(py, pz, tmp_path):
    """`scripts/mifrun -n P mif input.txt` (the reference's `mpirun -n P ./mif`) with one process per GPU: the files hold
    the single-rank golden's points rank by rank, the profiles are the golden's."""
    import numpy as np

    from conftest import GOLDEN_DIR
    from vtk_util import check_multi_rank_solution
    if device_count() < py * pz:
        pytest.skip(f"needs {py * pz} GPUs")
    golden = os.path.join(GOLDEN_DIR, "mif_case1")
    text = open(os.path.join(golden, "input.txt")).read().replace("Py : 1", f"Py : {py}").replace("Pz : 1", f"Pz : {pz}")
    (tmp_path / "input.txt").write_text(text)
    exe = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin", "mif")
    out = subprocess.run([os.path.join(ROOT, "scripts", "mifrun"), "-n", str(py * pz), exe, "input.txt"], cwd=tmp_path,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    check_multi_rank_solution(tmp_path / "solution.vtk", os.path.join(golden, "solution.vtk"), (17, 13, 15), (0.0, 0.0, -1.0),
                              (1.0 / 16, 1.0 / 12, 2.0 / 14), py, pz)
    for name in ("profile1.dat", "profile2.dat"):
        a, b = np.loadtxt(os.path.join(golden, name)), np.loadtxt(tmp_path / name)
        assert a.shape == b.shape and np.max(np.abs(a - b)) <= 1e-12, name


@pytest.mark.parametrize("ranks,pz", [(2, 2), (4, 2), (8, 4)])
def test_ported_full_test_on_several_gpus_prints_the_reference_numbers(ranks, pz):
    from conftest import GOLDEN_DIR
    if device_count() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    want = json.load(open(os.path.join(GOLDEN_DIR, "norms.json")))["full_test 16 1 2 (4 ranks)"]  # decomposition independent
    exe = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin", "full_test")
    out = subprocess.run([os.path.join(ROOT, "scripts", "mifrun"), "-n", str(ranks), exe, "16", "1", str(pz)], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    # (NCCL may print its version banner on stdout: keep the numeric lines)
    got = [float(x) for l in out.stdout.splitlines() if l.strip() and not l.startswith("NCCL") for x in l.split()]
    assert len(got) == 9
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-5 * abs(b), (got, want)


def test_velocity_test_mixed_with_periodic_y_split_over_two_gpus():
    """`mifrun -n 2 velocity_test_mixed 16 1 1` (Py = 2: the periodic y direction is distributed): the ported driver and
    the reference's unchanged test/velocity_test_mixed.cpp print the reference's numbers."""
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import GOLDEN_DIR
    want = json.load(open(os.path.join(GOLDEN_DIR, "norms.json")))["velocity_test_mixed 16 1 1"]
    bins = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin")
    runs = [[os.path.join(bins, "velocity_test"), "16", "1", "1", "mixed"]]
    if os.path.exists(os.path.join(bins, "ref_velocity_test_mixed")):
        runs.append([os.path.join(bins, "ref_velocity_test_mixed"), "16", "1", "1"])
    for cmd in runs:
        out = subprocess.run([os.path.join(ROOT, "scripts", "mifrun"), "-n", "2"] + cmd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
        got = [float(x) for x in out.stdout.split()[-3:]]
        for a, b in zip(got, want):
            assert abs(a - b) <= 2e-5 * abs(b), (cmd, got, want)
