"""GPU: the C++ host layer (mpi-incompressible-fluid_b200/host, the reference's own class and function names over
the C ABI) reproduces the numbers printed by the reference's own test programs (tests/golden/norms.json, written by
oracle/make_golden.py from test/full_test.cpp and test/pressure_test_*.cpp of the unmodified reference)."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR, ROOT

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host", "bin")
NORMS = json.load(open(os.path.join(GOLDEN_DIR, "norms.json")))


def run(exe, *args):
    out = subprocess.run([os.path.join(BIN, exe), *map(str, args)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    return out.stdout


def close(got, want):
    # the reference prints 6 significant digits
    return all(abs(g - w) <= 2e-5 * abs(w) for g, w in zip(got, want)) and len(got) == len(want)


@pytest.mark.parametrize("n,steps", [(16, 1), (32, 2), (64, 4)])
def test_full_test_prints_reference_numbers(n, steps):
    got = [float(x) for x in run("full_test", n, steps).split()]
    assert close(got, NORMS[f"full_test {n} {steps} 1"]), got


@pytest.mark.parametrize("kind", ["hn", "mixed", "nhn"])
@pytest.mark.parametrize("n", [8, 16, 32])
def test_pressure_tests_print_reference_numbers(kind, n):
    out = run("pressure_test", kind, n)
    line = [l for l in out.splitlines() if l.startswith("Errors:")][0]
    got = [float(x) for x in line.split()[1:]]
    assert close(got, NORMS[f"pressure_test_{kind} {n} 1"]), got


@pytest.mark.parametrize("mixed", [False, True])
@pytest.mark.parametrize("n", [16, 32])
def test_velocity_tests_print_reference_numbers(mixed, n):
    """test/velocity_test{,_mixed}.cpp: timestep_velocity + the (device-side) error norms through the host layer."""
    got = [float(x) for x in run("velocity_test", n, n // 16, *(["mixed"] if mixed else [])).split()]
    assert close(got, NORMS[f"velocity_test{'_mixed' if mixed else ''} {n} {n // 16} 1"]), got


def test_full_test_convergence_order():
    """Velocity converges with order ~2, pressure with order ~1.5 (analysis/plot_convergence.py:63-64,80-81)."""
    import math
    e16 = [float(x) for x in run("full_test", 16, 1).split()]
    e32 = [float(x) for x in run("full_test", 32, 2).split()]
    assert 1.8 <= math.log2(e16[1] / e32[1]) <= 2.7   # velocity L2
    assert 1.1 <= math.log2(e16[4] / e32[4]) <= 2.2   # pressure L2


@pytest.mark.parametrize("case", ["mif_case1", "mif_case2"])
def test_mif_driver_writes_the_reference_outputs(case, tmp_path):
    """`mif input.txt` (host/apps/mif.cpp, the port of src/main.cpp) against the files the reference driver wrote for the
    same input (tests/golden/mif_case*/, oracle/make_golden.py): solution.vtk and the profiles, compared numerically."""
    import re
    import shutil

    import numpy as np
    golden = os.path.join(GOLDEN_DIR, case)
    shutil.copy(os.path.join(golden, "input.txt"), tmp_path)
    out = subprocess.run([os.path.join(BIN, "mif"), "input.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr

    def vtk_fields(path):
        blob = open(path, "rb").read()
        n = int(re.search(rb"POINTS (\d+) double\n", blob).group(1))
        fields, off = {}, re.search(rb"POINTS \d+ double\n", blob).end()
        fields["points"] = np.frombuffer(blob[off:off + 24 * n], dtype=">f8")
        off += 24 * n
        for name in "uvwp":
            m = re.search(rb"SCALARS " + name.encode() + rb" double 1\nLOOKUP_TABLE default\n", blob[off:])
            off += m.end()
            fields[name] = np.frombuffer(blob[off:off + 8 * n], dtype=">f8")
            off += 8 * n
        return fields

    ref, got = vtk_fields(os.path.join(golden, "solution.vtk")), vtk_fields(os.path.join(tmp_path, "solution.vtk"))
    for key in ref:
        assert ref[key].shape == got[key].shape
        scale = max(float(np.max(np.abs(ref[key]))), 1e-6)
        assert float(np.max(np.abs(ref[key] - got[key]))) <= 1e-10 * scale, key
    profiles = [f for f in sorted(os.listdir(golden)) if f.startswith("profile")]
    assert profiles and sorted(f for f in os.listdir(tmp_path) if f.startswith("profile")) == profiles
    for name in profiles:
        a, b = np.loadtxt(os.path.join(golden, name)), np.loadtxt(os.path.join(tmp_path, name))
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=1e-7, atol=1e-12), name


# ---- the reference's own drivers, compiled unchanged against the host layer (host/Makefile: bin/ref_*) --------------------
def ref_binary(name):
    path = os.path.join(BIN, "ref_" + name)
    if not os.path.exists(path):
        pytest.skip("bin/ref_%s is built only where /root/reference is present (host/Makefile)" % name)
    return "ref_" + name


@pytest.mark.parametrize("n,steps", [(16, 1), (32, 2), (64, 4)])
def test_unchanged_reference_full_test(n, steps, tmp_path):
    """test/full_test.cpp of the reference, zero edits: host-side element access (pressure-gradient error through
    VELOCITY_TENSOR_SET_FOR_ALL_POINTS), adjust_pressure, the nine norms, writeVTK / writeVTKFullMesh / writeDat."""
    out = subprocess.run([os.path.join(BIN, ref_binary("full_test")), str(n), str(steps), "1"], cwd=tmp_path, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    got = [float(x) for l in out.stdout.splitlines() if l.strip() and not l.startswith("NCCL") for x in l.split()]
    assert close(got, NORMS[f"full_test {n} {steps} 1"]), got
    assert os.path.getsize(tmp_path / "solution.vtk") > 0 and os.path.getsize(tmp_path / "line1.dat") > 0


@pytest.mark.parametrize("kind", ["hn", "mixed", "nhn"])
@pytest.mark.parametrize("n", [8, 16, 32])
def test_unchanged_reference_pressure_tests(kind, n):
    out = run(ref_binary("pressure_test_" + kind), n, 1)
    line = [l for l in out.splitlines() if l.startswith("Errors:")][0]
    got = [float(x) for x in line.split()[1:]]
    assert close(got, NORMS[f"pressure_test_{kind} {n} 1"]), got


@pytest.mark.parametrize("mixed", [False, True])
def test_unchanged_reference_velocity_tests(mixed):
    name = "velocity_test_mixed" if mixed else "velocity_test"
    got = [float(x) for x in run(ref_binary(name), 16, 1, 1).split()]
    assert close(got, NORMS[f"{name} 16 1 1"]), got


@pytest.mark.parametrize("case", ["mif_case1", "mif_case2"])
def test_unchanged_reference_main_writes_the_reference_outputs(case, tmp_path):
    """src/main.cpp of the reference, zero edits, on the reference's input-file format: profiles against the files the CPU
    build of the same source wrote (tests/golden/mif_case*/)."""
    import shutil

    import numpy as np
    golden = os.path.join(GOLDEN_DIR, case)
    shutil.copy(os.path.join(golden, "input.txt"), tmp_path)
    out = subprocess.run([os.path.join(BIN, ref_binary("mif")), "input.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    profiles = [f for f in sorted(os.listdir(golden)) if f.startswith("profile")]
    assert profiles and sorted(f for f in os.listdir(tmp_path) if f.startswith("profile")) == profiles
    for name in profiles:
        a, b = np.loadtxt(os.path.join(golden, name)), np.loadtxt(os.path.join(tmp_path, name))
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=1e-7, atol=1e-12), name
    assert os.path.getsize(tmp_path / "solution.vtk") == os.path.getsize(os.path.join(golden, "solution.vtk"))
