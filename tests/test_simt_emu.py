"""CPU: the kernel SOURCES of libmifgpu executed by the SIMT interpreter of tests/simt_emu (every CUDA thread a fiber;
see tests/simt_emu/README.md) on a small selection of the GPU parity cases.  This is a development aid that catches
indexing / arithmetic mistakes in a kernel on a machine without a GPU; the parity claims themselves rest on the
`-m gpu` tests on a B200, not on this."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "simt_emu")


def test_kernel_sources_pass_parity_cases_under_the_simt_interpreter():
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    env = dict(os.environ, MIFGPU_LIB=os.path.join(EMU, "build", "libmifgpu_simt.so"))
    shards = 4  # the cases are independent: spread them over a few processes (the interpreter is single-threaded)
    runs = [subprocess.Popen([sys.executable, os.path.join(EMU, "run_cases.py"), "--shard", f"{i}/{shards}"], env=env,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for i in range(shards)]
    seen = 0
    for run in runs:
        out, err = run.communicate(timeout=900)
        assert run.returncode == 0, out[-3000:] + err[-3000:]
        assert "simt cases ok" in out
        seen += out.count(" ok\n") - 1
    assert seen == 13, seen  # every case of run_cases.py ran in exactly one shard


@pytest.mark.parametrize("policy", ["all", "odd", "even"])
def test_asynchronous_transfers_with_copies_carried_out_as_late_as_allowed(policy):
    """mifgpu_tensor_upload_async / _download_async (link and re-pitching streams, two staging buffers per direction,
    per-tensor events) with the interpreter's asynchronous copies queued per stream and carried out as late as the
    programming model allows ("all"), or with every second stream racing ahead ("odd" / "even"): a missing dependency
    or a staging buffer refilled before it was consumed changes the fields.  (Removing either of the two staging
    hand-over waits fails exactly one of the policies; this model also found the one real ordering bug of the API --
    a new tensor's zero fill could be overtaken by an asynchronous upload.)"""
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    env = dict(os.environ, MIFGPU_LIB=os.path.join(EMU, "build", "libmifgpu_simt.so"), MIF_EMU_LAZY_COPIES=policy)
    run = subprocess.run([sys.executable, os.path.join(EMU, "run_async_cases.py")], env=env, capture_output=True, text=True,
                         timeout=900)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-3000:]
    assert "async cases ok" in run.stdout


def test_bench_line_carries_the_contract_keys_in_a_dry_run():
    """bench.py's own arm end to end with the kernels under the SIMT interpreter and torch.cuda's stream / event calls
    replaced by inert stand-ins (tests/simt_emu/bench_dry_run.py): the JSON line is assembled without a Python error and
    has every key of the contract.  The values are timings of an interpreter and mean nothing."""
    import json
    build = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout[-2000:] + build.stderr[-2000:]
    env = dict(os.environ, MIFGPU_LIB=os.path.join(EMU, "build", "libmifgpu_simt.so"))
    run = subprocess.run([sys.executable, os.path.join(EMU, "bench_dry_run.py"), "--size", "9", "--steps", "2", "--warmup", "3",
                          "--ref-size", "9"], env=env, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-3000:]
    lines = [l for l in run.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, run.stdout
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["unit"] == "cell-steps/s" and line["dtype"] == "f64" and line["warmup"] >= 3 and line["n_gpus"] == 1
    assert "workload" in line["config"] and "model" not in line["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["peak"] > 1000
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in line["cpu_baseline"], key
    assert line["cpu_baseline"]["kind"] == "reference"
    assert line["gpu_launches"] == 2 * 27  # 27 kernels per projection step
