"""CPU, world size 2: the multi-rank paths of libmifgpu -- plane halos, the grouped send/recv all-to-all transposes, and
the fused compute + transpose over peer memory with its rank barriers -- executed for real, one PROCESS per rank under
torchrun (gloo for the rendezvous, as on a GPU box), with the kernel sources run by the SIMT interpreter
(tests/simt_emu), NCCL replaced by the mailbox stand-in (tests/simt_emu/fake_nccl.cpp, bound through MIFGPU_NCCL_LIB)
and CUDA IPC by shared-memory files.  Every rank's slab must reproduce the single-rank result to 1e-11, exactly the
check tests/test_gpu_multi.py makes on real GPUs with tests/mp_worker.py.  A development aid for the host-side
decomposition logic and the kernels' peer addressing; the parity claim for N > 1 rests on the GPU run."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "simt_emu")


@pytest.fixture(scope="module")
def simt_env():
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    return dict(os.environ, MIFGPU_LIB=os.path.join(EMU, "build", "libmifgpu_simt.so"),
                MIFGPU_NCCL_LIB=os.path.join(EMU, "build", "libmif_fake_nccl.so"), MIF_SIMT_IPC="1")


def run_worker(env, case, world, port, py=1):
    env = dict(env, MIF_PY=str(py))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_two_rank_slabs_with_nccl_style_transposes(simt_env):
    # 16^3 points, two time steps: off the fast path, so the Y<->Z transposes are pack -> all-to-all -> unpack
    res = run_worker(simt_env, "full_16_2", 2, 29711)
    assert res["world"] == 2 and res["max_rel_err"] <= 1e-11


def test_two_rank_slabs_with_peer_memory_transposes(simt_env):
    # 257-point y and z lines: the sweeps store straight into the other rank's (blocked) pencil / staging buffers
    res = run_worker(simt_env, "es:3x257x257", 2, 29712)
    assert res["world"] == 2 and res["max_rel_err"] <= 1e-11


@pytest.mark.parametrize("case,world,port", [("full_16_2", 2, 29721), ("full_17_1", 3, 29722), ("pz:6x6x13", 3, 29723)])
def test_overlapped_halo_exchanges_completing_as_late_as_stream_order_allows(simt_env, case, world, port):
    """Inside the time step the plane exchanges run on a second stream while the consuming stencil kernel covers its
    interior planes (mif_api.cu, launch_around_halo).  The interpreter runs kernels synchronously, so the default run is
    the exchange finishing at once; MIF_FAKE_NCCL_LATE=1 defers every side-stream send / receive to the point where the
    compute stream waits for it -- a kernel that touched a ghost plane (or a plane still to be sent) too early would
    now compute with stale data and miss the single-rank result."""
    res = run_worker(dict(simt_env, MIF_FAKE_NCCL_LATE="1", MIF_EMU_LAZY_COPIES="all"), case, world, port)  # copies as late as allowed, too
    assert res["world"] == world and res["max_rel_err"] <= 1e-11


def test_halo_overlap_can_be_switched_off(simt_env):
    res = run_worker(dict(simt_env, MIFGPU_NO_HALO_OVERLAP="1"), "full_16_2", 2, 29724)
    assert res["world"] == 2 and res["max_rel_err"] <= 1e-11


@pytest.mark.parametrize("world", [2, 3])
def test_periodic_z_distributed_over_the_ranks(simt_env, world):
    # two ranks: prev and next neighbour are the same peer (the halo planes pair in issue order); three ranks: a ring
    res = run_worker(simt_env, "pz:6x6x13", world, 29717 + world)
    assert res["world"] == world and res["max_rel_err"] <= 1e-11


@pytest.mark.parametrize("case,world,tol", [
    ("mr_lid2_10x12x9_pz2", 2, 1e-11),          # periodic z over two slabs, lid-driven case, two projection steps
    ("mr_full_p010_9x12x10_py2", 2, 1e-11),     # periodic y over two y ranks (same-peer sheets in issue order)
    ("mr_full_p010_9x13x10_py3", 3, 1e-11),     # ... over a ring of three
    ("mr_vtest_mixed_12_py2", 2, 1e-10),        # periodic x on the rank + periodic y over two ranks, velocity-only integrator
    ("mr_full_p011_9x13x11_py2pz2", 4, 1e-6),   # periodic y AND z distributed (2 x 2)
])
def test_distributed_periodic_directions_match_the_reference_on_the_same_ranks(simt_env, case, world, tol):
    """The reference ITSELF run on `world` ranks (oracle/make_golden.py --multi-rank keeps every rank's arrays): each rank
    uploads its local arrays of that run and must reproduce the reference's local arrays after every step, ghost planes
    and rows included.  A single-rank result cannot be the target here -- with a distributed periodic direction the
    reference's own runs differ between decompositions (its one-rank ghost copy of the staggered component is shifted by
    one cell).  Two known differences, both in points / effects the reference leaves one exchange
    stale: the x-periodic corners of received y sheets (1.2e-11 on the velocity-only case) and, with Py > 1 AND Pz > 1,
    the y-ghost / z-ghost edges (2e-7 on the 2 x 2 case, where the library keeps them fresh)."""
    py = {"mr_full_p010_9x12x10_py2": 2, "mr_full_p010_9x13x10_py3": 3, "mr_vtest_mixed_12_py2": 2, "mr_full_p011_9x13x11_py2pz2": 2}.get(case, 1)
    res = run_worker(dict(simt_env, MIF_WORKER_TOL=str(tol)), "mr:" + case, world, 29740 + world + 7 * py, py=py)
    assert res["world"] == world and res["max_rel_err"] <= tol, res


@pytest.mark.parametrize("case,world,py", [("mr_full_17_py2pz2", 4, 2), ("mr_full_p011_9x13x11_py2pz2", 4, 2), ("mr_vtest_mixed_12_py2", 2, 2)])
def test_reference_halo_mode_reproduces_the_reference_pencil_runs_point_for_point(simt_env, case, world, py):
    """MIFGPU_REFERENCE_HALOS=1: z planes first, y sheets unpacked without their borders -- the reference's order and
    extents (src/StaggeredTensor.cpp:60-165), stale edge ghosts included.  A Py x Pz run then equals the reference's
    Py x Pz run at every point of every rank's arrays, with no point left out of the comparison (the default keeps the
    edges fresh and equals the ONE-rank run instead: test_four_rank_pencils; on 17^3 the two differ by 4e-8)."""
    res = run_worker(dict(simt_env, MIFGPU_REFERENCE_HALOS="1"), "mr:" + case, world, 29770 + world + py, py=py)
    assert res["world"] == world and res["max_rel_err"] <= 1e-11, res


def test_four_rank_pencils(simt_env):
    # Py x Pz = 2 x 2 on 17^3 points (uneven blocks 9 + 8): y sheets then z planes as halos, the four 2Decomp transposes
    # as box exchanges, x / y / z sweeps on the sub-domain, the y pencil and the z pencil; the staging copies of the box
    # exchanges are asynchronous 3-D copies: run with every second stream racing ahead of the lazily served others
    res = run_worker(dict(simt_env, MIF_EMU_LAZY_COPIES="odd"), "full_17_1", 4, 29714, py=2)
    assert res["world"] == 4 and res["Py"] == 2 and res["Pz"] == 2 and res["max_rel_err"] <= 1e-11


def test_ported_driver_on_four_ranks_writes_the_reference_files(simt_env, tmp_path):
    """`scripts/mifrun -n 4 mif input.txt` with Py = Pz = 2 (the reference's `mpirun -n 4 ./mif`): rank and size from the
    launcher's environment, communicator id through the rendezvous file, norms / writers through mifgpu_gather.  The
    files hold the single-rank golden's points rank by rank; the profiles (sorted on rank 0) are the golden's."""
    import shutil

    import numpy as np

    from conftest import GOLDEN_DIR
    from vtk_util import check_multi_rank_solution
    golden = os.path.join(GOLDEN_DIR, "mif_case1")
    text = open(os.path.join(golden, "input.txt")).read().replace("Py : 1", "Py : 2").replace("Pz : 1", "Pz : 2")
    (tmp_path / "input.txt").write_text(text)
    env = dict(simt_env, LD_LIBRARY_PATH=os.path.join(EMU, "build", "as_libmifgpu"))
    exe = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin", "mif")
    out = subprocess.run([os.path.join(ROOT, "scripts", "mifrun"), "-n", "4", exe, "input.txt"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    check_multi_rank_solution(tmp_path / "solution.vtk", os.path.join(golden, "solution.vtk"), (17, 13, 15), (0.0, 0.0, -1.0),
                              (1.0 / 16, 1.0 / 12, 2.0 / 14), 2, 2)
    for name in ("profile1.dat", "profile2.dat"):
        a, b = np.loadtxt(os.path.join(golden, name)), np.loadtxt(tmp_path / name)
        assert a.shape == b.shape and np.max(np.abs(a - b)) <= 1e-12, name


def test_ported_full_test_on_four_ranks_prints_the_reference_numbers(simt_env):
    """`mifrun -n 4 full_test 16 1 2` (Py = Pz = 2): per-rank norms folded on rank 0 as in src/Norms.cpp:120-162, the
    pressure gauge fixed across ranks (src/PressureEquation.cpp:288-343); the nine numbers are the ones the reference
    prints for `mpirun -n 4 full_test 16 1 2` (tests/golden/norms.json)."""
    from conftest import GOLDEN_DIR
    want = json.load(open(os.path.join(GOLDEN_DIR, "norms.json")))["full_test 16 1 2 (4 ranks)"]
    env = dict(simt_env, LD_LIBRARY_PATH=os.path.join(EMU, "build", "as_libmifgpu"))
    exe = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin", "full_test")
    out = subprocess.run([os.path.join(ROOT, "scripts", "mifrun"), "-n", "4", exe, "16", "1", "2"], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    got = [float(x) for x in out.stdout.split()]
    assert len(got) == 9
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-5 * abs(b), (got, want)


def test_ported_pressure_tests_on_several_ranks_print_the_reference_numbers(simt_env):
    """pressure_test_hn / _nhn (the latter with host-callback Neumann faces per rank) on 2 x 2 and 3 x 1 ranks."""
    from conftest import GOLDEN_DIR
    norms = json.load(open(os.path.join(GOLDEN_DIR, "norms.json")))
    env = dict(simt_env, LD_LIBRARY_PATH=os.path.join(EMU, "build", "as_libmifgpu"))
    exe = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin", "pressure_test")
    for kind, ranks, pz in (("hn", 4, 2), ("nhn", 3, 1)):
        out = subprocess.run([os.path.join(ROOT, "scripts", "mifrun"), "-n", str(ranks), exe, kind, "8", str(pz)], env=env,
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
        line = [l for l in out.stdout.splitlines() if l.startswith("Errors:")][-1]
        got, want = [float(x) for x in line.split()[1:]], norms[f"pressure_test_{kind} 8 1"]
        for a, b in zip(got, want):
            assert abs(a - b) <= 2e-5 * abs(b), (kind, got, want)


def test_bench_line_on_two_ranks_in_a_dry_run(simt_env):
    """bench.py's own arm under torchrun on two emulated ranks (tests/simt_emu/bench_dry_run.py: gloo for torch.distributed,
    inert stand-ins for torch.cuda): rank 0 prints one JSON line with the multi-GPU keys.  Values are meaningless."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29716", os.path.join(EMU, "bench_dry_run.py"), "--gpus", "2", "--size", "9", "--steps", "2", "--warmup", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(simt_env, MIF_BENCH_PARITY_CASE="5x17x17"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["config"]["points"] == [9, 9, 17]
    # the multi-rank result is checked against a single-rank context before the timed region, and the lid is where
    # the boundary data puts it
    assert line["parity_vs_single_rank"]["ok"] and line["parity_vs_single_rank"]["max_rel_linf"] <= 1e-11
    assert line["config"]["lid_on_last_x_face"] is True and line["config"]["transpose_path"] == "NCCL all-to-all"
    assert line["e2e"]["result_finite"] is True
    assert line["nvlink"]["bytes_sent_per_gpu_per_step"] == 6 * 8 * (9 * 9 * 17 / 2) * 0.5
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0 and "halo_exchange" in line["kernels"]


def test_velocity_test_mixed_with_periodic_y_split_over_two_ranks(simt_env):
    """`mifrun -n 2 velocity_test_mixed 16 1 1` (Pz = 1, so Py = 2: the periodic y direction is distributed, its neighbours
    wrap around) -- the ported driver and, where it was built, the reference's own test/velocity_test_mixed.cpp compiled
    unchanged print the numbers of the reference (which prints the same on one and on two ranks)."""
    from conftest import GOLDEN_DIR
    want = json.load(open(os.path.join(GOLDEN_DIR, "norms.json")))["velocity_test_mixed 16 1 1"]
    env = dict(simt_env, LD_LIBRARY_PATH=os.path.join(EMU, "build", "as_libmifgpu"))
    bins = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin")
    runs = [[os.path.join(bins, "velocity_test"), "16", "1", "1", "mixed"]]
    if os.path.exists(os.path.join(bins, "ref_velocity_test_mixed")):
        runs.append([os.path.join(bins, "ref_velocity_test_mixed"), "16", "1", "1"])
    for cmd in runs:
        out = subprocess.run([os.path.join(ROOT, "scripts", "mifrun"), "-n", "2"] + cmd, env=env, capture_output=True, text=True,
                             timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
        got = [float(x) for x in out.stdout.split()[-3:]]
        for a, b in zip(got, want):
            assert abs(a - b) <= 2e-5 * abs(b), (cmd, got, want)
