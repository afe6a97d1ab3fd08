"""GPU, >= 2 devices: the slab-decomposed path (NCCL halos + all-to-all transposes inside libmifgpu) reproduces the
single-rank reference goldens on every rank's slab (tests/mp_worker.py under torchrun)."""
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


@pytest.mark.parametrize("case", ["full_16_2", "full_17_1", "lid1_12x10x14_2", "full_65x17x9_1", "full_6x65x9_1",
                                  "es:9x257x257", "es:7x513x257", "es:5x1025x1025"])
@pytest.mark.parametrize("world", [2, 4])
def test_slab_decomposition_matches_single_rank_reference(case, world):
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mp_worker.py"), case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, failure_report(out)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["max_rel_err"] <= 1e-11


@pytest.mark.parametrize("no_peer", ["", "1"])
def test_peer_memory_and_nccl_transposes_agree(no_peer):
    """The same case through the fused peer-store transposes (default) and through the NCCL all-to-all (MIFGPU_NO_PEER)."""
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    if no_peer:
        env["MIFGPU_NO_PEER"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mp_worker.py"), "es:9x257x257"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, failure_report(out)


@pytest.mark.parametrize("world", [2, 4])
def test_periodic_z_distributed_over_the_ranks(world):
    """Periodic z split over the ranks (src/Constants.cpp:98-101): the solve and its halo exchange against the oracle's
    single-rank solve, ghost planes included (two ranks: both neighbours are the same peer)."""
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "tests", "mp_worker.py"), "pz:12x10x33"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, failure_report(out)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["max_rel_err"] <= 1e-11


@pytest.mark.parametrize("case,world,py,tol", [
    ("mr_lid2_10x12x9_pz2", 2, 1, 1e-11), ("mr_lid2_10x12x14_pz3", 3, 1, 1e-11), ("mr_full_p010_9x12x10_py2", 2, 2, 1e-11),
    ("mr_full_p010_9x13x10_py3", 3, 3, 1e-11), ("mr_vtest_mixed_12_py2", 2, 2, 1e-10), ("mr_full_p011_9x13x11_py2pz2", 4, 2, 1e-6),
])
def test_distributed_periodic_directions_match_the_reference_on_the_same_ranks(case, world, py, tol):
    """Time steps with a periodic z and / or y direction split over the ranks against the reference ITSELF run on the same
    number of ranks, every rank's arrays, ghosts included (tests/test_multi_cpu_simt.py explains the tolerances)."""
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MIF_PY=str(py), MIF_WORKER_TOL=str(tol))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + world + py), os.path.join(ROOT, "tests", "mp_worker.py"), "mr:" + case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, failure_report(out)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["max_rel_err"] <= tol


@pytest.mark.parametrize("case,world,py", [("mr_full_17_py2pz2", 4, 2), ("mr_full_p011_9x13x11_py2pz2", 4, 2), ("mr_vtest_mixed_12_py2", 2, 2)])
def test_reference_halo_mode_reproduces_the_reference_pencil_runs_point_for_point(case, world, py):
    """MIFGPU_REFERENCE_HALOS=1 (the reference's exchange order and extents, stale edge ghosts included): a Py x Pz run
    equals the reference's own Py x Pz run at every point of every rank's arrays."""
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MIF_PY=str(py), MIFGPU_REFERENCE_HALOS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29580 + world + py), os.path.join(ROOT, "tests", "mp_worker.py"), "mr:" + case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, failure_report(out)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["max_rel_err"] <= 1e-11
