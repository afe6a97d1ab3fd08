"""CPU: host-side pieces of the C++ layer that need no GPU -- the output writers against files written by the
reference's own writers (tests/golden/export_*/, produced by oracle/ref_dump.cpp `export` from src/VTKDatExport.cpp on
the same analytically set fields) and the input parser against the reference's file format and error behaviour
(src/InputParser.cpp)."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT

HOST = os.path.join(ROOT, "mpi-incompressible-fluid_b200", "host")
EXE = os.path.join(HOST, "bin", "export_test")


@pytest.fixture(scope="module", autouse=True)
def built():
    proc = subprocess.run(["make", "-C", HOST, "bin/export_test"], capture_output=True, text=True)
    if proc.returncode != 0 and not os.path.exists(EXE):
        pytest.skip("host layer not buildable here (libmifgpu.so missing?): " + proc.stderr[-300:])


def parse_binary_vtk(path):
    blob = open(path, "rb").read()
    m = re.search(rb"# vtk DataFile Version 2.0\nvtk output\nBINARY\nDATASET UNSTRUCTURED_GRID \nPOINTS (\d+) double\n", blob)
    assert m and m.start() == 0
    n, off = int(m.group(1)), m.end()
    fields = {"points": np.frombuffer(blob[off:off + 24 * n], dtype=">f8")}
    off += 24 * n
    headers = [b"\nPOINT_DATA %d\nSCALARS u double 1\nLOOKUP_TABLE default\n" % n] + \
              [b"\nSCALARS %s double 1\nLOOKUP_TABLE default\n" % c for c in (b"v", b"w", b"p")]
    for name, header in zip("uvwp", headers):
        assert blob[off:off + len(header)] == header, name
        off += len(header)
        fields[name] = np.frombuffer(blob[off:off + 8 * n], dtype=">f8")
        off += 8 * n
    assert off == len(blob)
    return n, fields


@pytest.mark.parametrize("periodic_z", [0, 1])
def test_writers_match_reference_files(tmp_path, periodic_z):
    subprocess.run([EXE, "8", str(periodic_z), str(tmp_path)], check=True, timeout=60)
    golden = os.path.join(GOLDEN_DIR, f"export_{periodic_z}")
    n_ref, ref = parse_binary_vtk(os.path.join(golden, "solution.vtk"))
    n_got, got = parse_binary_vtk(os.path.join(tmp_path, "solution.vtk"))
    assert n_ref == n_got
    for key in ref:
        assert np.max(np.abs(ref[key] - got[key])) <= 1e-14, key
    for name in ("profile_x.dat", "profile_y.dat", "profile_z.dat"):
        a = np.loadtxt(os.path.join(golden, name))
        b = np.loadtxt(os.path.join(tmp_path, name))
        assert a.shape == b.shape and a.shape[1] == 7
        assert np.max(np.abs(a - b)) <= 1e-12, name
        # identical text layout: "%.8f %.8f %.8f %.8e %.8e %.8e %.8e"
        first = open(os.path.join(tmp_path, name)).readline().split()
        assert all(re.fullmatch(r"-?\d+\.\d{8}", t) for t in first[:3])
        assert all(re.fullmatch(r"-?\d\.\d{8}e[+-]\d{2}", t) for t in first[3:])
    ref_tokens = open(os.path.join(golden, "full.vtk")).read().split()
    got_tokens = open(os.path.join(tmp_path, "full.vtk")).read().split()
    assert len(ref_tokens) == len(got_tokens)
    for a, b in zip(ref_tokens, got_tokens):
        if a != b:
            assert abs(float(a) - float(b)) <= 1e-14


def parse(tmp_path, text):
    path = os.path.join(tmp_path, "input.txt")
    with open(path, "w") as f:
        f.write(text)
    return subprocess.run([EXE, "parse", path], capture_output=True, text=True, timeout=30).stdout.strip()


REFERENCE_INPUT = "Nt : 1000\ndt : 1e-3\nNx : 480\nNy : 480\nNz : 480\nPy : 8\nPz : 8\ntest_case_2 : false"


def test_input_parser_reads_the_reference_format(tmp_path):
    # /root/reference/input/input.txt verbatim
    assert parse(tmp_path, REFERENCE_INPUT) == "480 480 480 0.001 1000 8 8 0"
    assert parse(tmp_path, "Nx:7\nNy :8\nNz: 9\ndt :  0.5\nNt:3\nPy:1\nPz:1\ntest_case_2:true\nnot a key line\n") == "7 8 9 0.5 3 1 1 1"


def test_input_parser_error_behaviour(tmp_path):
    assert parse(tmp_path, REFERENCE_INPUT + "\nfoo : 1") == "error: Unknown or duplicate key: foo"
    assert parse(tmp_path, REFERENCE_INPUT + "\nNx : 3") == "error: Unknown or duplicate key: Nx"
    missing = parse(tmp_path, REFERENCE_INPUT.replace("Pz : 8\n", ""))
    assert missing == "error: Missing one or more inputs. Required inputs: Nx, Ny, Nz, dt, Nt, Py, Pz, test_case_2."
    out = subprocess.run([EXE, "parse", "/nonexistent/input.txt"], capture_output=True, text=True).stdout.strip()
    assert out == "error: Error opening input file: /nonexistent/input.txt"
