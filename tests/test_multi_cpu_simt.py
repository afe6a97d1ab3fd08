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


def test_four_rank_pencils(simt_env):
    # Py x Pz = 2 x 2 on 17^3 points (uneven blocks 9 + 8): y sheets then z planes as halos, the four 2Decomp transposes
    # as box exchanges, x / y / z sweeps on the sub-domain, the y pencil and the z pencil
    res = run_worker(simt_env, "full_17_1", 4, 29714, py=2)
    assert res["world"] == 4 and res["Py"] == 2 and res["Pz"] == 2 and res["max_rel_err"] <= 1e-11
