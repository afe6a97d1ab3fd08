#!/usr/bin/env python3
"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference (oracle/_ref/,
built by oracle/Makefile from /root/reference) in THIS container.  The GPU box has no /root/reference,
so the vectors are committed; re-run this script (`python oracle/make_golden.py`) to regenerate them.

  tests/golden/<case>.npz   raw FP64 fields dumped by oracle/_ref/ref_dump (reference layout,
                            arrays shaped (sz, sy, sx)), plus the case parameters
  tests/golden/norms.json   the numbers printed by the reference's own test mains
                            (test/full_test.cpp:171-175, test/pressure_test_*.cpp:76-79, velocity tests)
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
GOLDEN = os.path.join(HERE, "..", "tests", "golden")


def run(cmd, np_ranks=1, cwd=None):
    env = dict(os.environ, MIF_SHIM_NP=str(np_ranks), MIF_SHIM_ABORT_ON_STALL="1")
    out = subprocess.run(cmd, env=env, cwd=cwd, check=True, capture_output=True, text=True, timeout=600)
    return out.stdout


def load_dump(outdir, rank=0, dtype=np.float64):
    fields = {}
    with open(os.path.join(outdir, f"manifest_r{rank}.txt")) as f:
        for line in f:
            name, r, sx, sy, sz = line.split()
            data = np.fromfile(os.path.join(outdir, f"{name}_r{r}.f64"), dtype=dtype)  # the dumper writes sizeof(Real)
            fields[name] = data.reshape(int(sz), int(sy), int(sx))
    return fields


def dump_case(name, args, meta, np_ranks=1, f32=False):
    tmp = tempfile.mkdtemp(prefix="mifgolden_")
    try:
        cmd = [os.path.join(REF, "f32" if f32 else "", "ref_dump")] + [str(a).replace("{out}", tmp) for a in args]
        stdout = run(cmd, np_ranks)
        fields = load_dump(tmp, dtype=np.float32 if f32 else np.float64)
        meta = dict(meta, command="ref_dump " + " ".join(str(a) for a in args), stdout=stdout.strip())
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **fields)
        print(f"{name}: {len(fields)} arrays, stdout={stdout.strip()!r}")
    finally:
        shutil.rmtree(tmp)


def velocity_cases():
    """timestep_velocity (src/TimestepVelocity.cpp:58-90) as run by test/velocity_test{,_mixed}.cpp: manufactured
    velocity + forcing of generators/manufsol_velocity.py, Re = 1e4, final time 1e-4."""
    for mixed in (0, 1):
        ln = 2 * np.pi if mixed else 1.0
        meta = dict(kind="vtest", N=[12, 12, 12], x_size=ln, y_size=ln, z_size=1.0, min=[0.0, 0.0, 0.0], Re=1e4,
                    final_time=1e-4, steps=2, periodic=[mixed, mixed, 0], bc="velocity_test")
        args = ["vtest", 12, 2, 1, "{out}"] + (["mixed"] if mixed else [])
        dump_case("vtest_mixed_12_2" if mixed else "vtest_12_2", args, meta)


def multi_rank_cases():
    """The reference ITSELF on several ranks (threads of the MPI stand-in), every rank's local arrays kept: the pins of
    the distributed periodic directions (neighbours wrap around, src/Constants.cpp:98-101; halos of the staggered
    component, src/StaggeredTensor.cpp:60-165), which cannot be checked by cutting a single-rank result -- the
    reference's one-rank periodic ghost copy has its own one-cell quirk.  Keys: <field>_s<step>_r<rank>."""
    def dump(name, args, meta, ranks):
        tmp = tempfile.mkdtemp(prefix="mifgolden_")
        try:
            run([os.path.join(REF, "ref_dump")] + [str(a).replace("{out}", tmp) for a in args], ranks)
            fields = {}
            for r in range(ranks):
                for key, arr in load_dump(tmp, rank=r).items():
                    fields[f"{key}_r{r}"] = arr
            meta = dict(meta, ranks=ranks, command=f"MIF_SHIM_NP={ranks} ref_dump " + " ".join(str(a) for a in args))
            np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **fields)
            print(f"{name}: {len(fields)} arrays on {ranks} ranks")
        finally:
            shutil.rmtree(tmp)

    lid2 = dict(kind="lid", x_size=1.0, y_size=1.0, z_size=1.0, min=[-0.5, -0.5, -0.5], Re=1e3, periodic=[0, 0, 1],
                bc="test_case_2", steps=2, dt=1e-3, final_time=2e-3)
    dump("mr_lid2_10x12x9_pz2", ["lid", 10, 12, 9, 1e-3, 2, 1, 2, "{out}"], dict(lid2, N=[10, 12, 9], Py=1, Pz=2), 2)
    dump("mr_lid2_10x12x14_pz3", ["lid", 10, 12, 14, 1e-3, 2, 1, 3, "{out}"], dict(lid2, N=[10, 12, 14], Py=1, Pz=3), 3)
    full = dict(kind="full", x_size=1.0, y_size=1.0, z_size=2.0, min=[0.0, 0.0, -1.0], Re=1e3, final_time=1e-4, steps=1,
                bc="ethier_steinman")
    dump("mr_full_p010_9x12x10_py2", ["full", 9, 1, 1, "{out}", "hn", 12, 10, "010"],
         dict(full, N=[9, 12, 10], periodic=[0, 1, 0], Py=2, Pz=1), 2)
    dump("mr_full_p011_9x13x11_py2pz2", ["full", 9, 1, 2, "{out}", "hn", 13, 11, "011"],
         dict(full, N=[9, 13, 11], periodic=[0, 1, 1], Py=2, Pz=2), 4)
    dump("mr_full_17_py2pz2", ["full", 17, 1, 2, "{out}"], dict(full, N=[17, 17, 17], periodic=[0, 0, 0], Py=2, Pz=2), 4)
    dump("mr_full_p010_9x13x10_py3", ["full", 9, 1, 1, "{out}", "hn", 13, 10, "010"],
         dict(full, N=[9, 13, 10], periodic=[0, 1, 0], Py=3, Pz=1), 3)
    ln = 2 * np.pi
    vt = dict(kind="vtest", N=[12, 12, 12], x_size=ln, y_size=ln, z_size=1.0, min=[0.0, 0.0, 0.0], Re=1e4, final_time=1e-4,
              steps=2, periodic=[1, 1, 0], bc="velocity_test")
    dump("mr_vtest_mixed_12_py2", ["vtest", 12, 2, 1, "{out}", "mixed"], dict(vt, Py=2, Pz=1), 2)


def f32_cases():
    """The reference's USE_DOUBLE=0 build (oracle/_ref/f32, `make -C oracle f32`): float32 fields of three of the cases
    above plus the numbers its full_test prints -- the pins of libmifgpu_f32.so (tests/fp32_cases.py)."""
    full = dict(kind="full", x_size=1.0, y_size=1.0, z_size=2.0, min=[0.0, 0.0, -1.0], Re=1e3, final_time=1e-4,
                periodic=[0, 0, 0], bc="ethier_steinman", real="float32")
    dump_case("f32_full_16_2", ["full", 16, 2, 1, "{out}"], dict(full, N=[16, 16, 16], steps=2, nhn=0), f32=True)
    lid1 = dict(kind="lid", x_size=1.0, y_size=1.0, z_size=2.0, min=[0.0, 0.0, -1.0], Re=1e3, periodic=[0, 0, 0],
                bc="test_case_1", real="float32")
    dump_case("f32_lid1_12x10x14_2", ["lid", 12, 10, 14, 1e-3, 2, 0, 1, "{out}"],
              dict(lid1, N=[12, 10, 14], steps=2, dt=1e-3, final_time=2e-3), f32=True)
    lid2 = dict(kind="lid", x_size=1.0, y_size=1.0, z_size=1.0, min=[-0.5, -0.5, -0.5], Re=1e3, periodic=[0, 0, 1],
                bc="test_case_2", real="float32")
    dump_case("f32_lid2_10x12x9_2", ["lid", 10, 12, 9, 1e-3, 2, 1, 1, "{out}"],
              dict(lid2, N=[10, 12, 9], steps=2, dt=1e-3, final_time=2e-3), f32=True)
    norms = {}
    tmp = tempfile.mkdtemp(prefix="mifgolden_")
    try:
        for n, steps in ((16, 1), (32, 2)):
            out = run([os.path.join(REF, "f32", "full_test"), str(n), str(steps), "1"], cwd=tmp)
            norms[f"full_test {n} {steps} 1"] = [float(x) for x in out.split()]
    finally:
        shutil.rmtree(tmp)
    with open(os.path.join(GOLDEN, "f32_norms.json"), "w") as f:
        json.dump(norms, f, indent=1)
    print("f32 norms", norms)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if not os.path.exists(os.path.join(REF, "ref_dump")):
        sys.exit("oracle/_ref/ref_dump missing: run `make -C oracle ref` first")
    if len(sys.argv) > 1 and sys.argv[1] == "--multi-rank":  # add the several-rank cases without touching the others
        multi_rank_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--f32":  # add the single-precision cases without touching the others
        if not os.path.exists(os.path.join(REF, "f32", "ref_dump")):
            sys.exit("oracle/_ref/f32/ref_dump missing: run `make -C oracle f32` first")
        f32_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--only-velocity":  # add the velocity cases without touching the others
        velocity_cases()
        return
    velocity_cases()

    # --- raw-field cases --------------------------------------------------------------------------
    full = dict(kind="full", x_size=1.0, y_size=1.0, z_size=2.0, min=[0.0, 0.0, -1.0], Re=1e3, final_time=1e-4,
                periodic=[0, 0, 0], bc="ethier_steinman")
    dump_case("full_16_2", ["full", 16, 2, 1, "{out}"], dict(full, N=[16, 16, 16], steps=2, nhn=0))
    dump_case("full_17_1", ["full", 17, 1, 1, "{out}"], dict(full, N=[17, 17, 17], steps=1, nhn=0))
    dump_case("full_12_1_nhn", ["full", 12, 1, 1, "{out}", "nhn"], dict(full, N=[12, 12, 12], steps=1, nhn=1))
    # power-of-two cell counts (N - 1 = 64, 128): these exercise the register-blocked DCT-I fast path
    dump_case("full_65x17x9_1", ["full", 65, 1, 1, "{out}", "hn", 17, 9], dict(full, N=[65, 17, 9], steps=1, nhn=0))
    dump_case("full_6x65x9_1", ["full", 6, 1, 1, "{out}", "hn", 65, 9], dict(full, N=[6, 65, 9], steps=1, nhn=0))
    lid1 = dict(kind="lid", x_size=1.0, y_size=1.0, z_size=2.0, min=[0.0, 0.0, -1.0], Re=1e3, periodic=[0, 0, 0],
                bc="test_case_1")
    dump_case("lid1_12x10x14_2", ["lid", 12, 10, 14, 1e-3, 2, 0, 1, "{out}"],
              dict(lid1, N=[12, 10, 14], steps=2, dt=1e-3, final_time=2e-3))
    lid2 = dict(kind="lid", x_size=1.0, y_size=1.0, z_size=1.0, min=[-0.5, -0.5, -0.5], Re=1e3, periodic=[0, 0, 1],
                bc="test_case_2")
    dump_case("lid2_10x12x9_2", ["lid", 10, 12, 9, 1e-3, 2, 1, 1, "{out}"],
              dict(lid2, N=[10, 12, 9], steps=2, dt=1e-3, final_time=2e-3))
    for kind, dims in (("hn", (8, 24, 40)), ("mixed", (8, 24, 40)), ("nhn", (8, 24, 40)), ("mixed", (9, 17, 17)),
                       ("hn", (17, 9, 33)), ("mixed", (7, 6, 10)), ("hn", (65, 9, 129)), ("mixed", (9, 65, 17))):
        lo = -np.pi / 2 if kind == "nhn" else 0.0
        ln = np.pi / 2 if kind == "nhn" else 2 * np.pi
        meta = dict(kind="ptest", ptest=kind, N=list(dims), x_size=ln, y_size=ln, z_size=ln, min=[lo, lo, lo], Re=1.0,
                    final_time=1.0, steps=1, periodic=[0, 0, int(kind == "mixed")], time=1.0)
        dump_case(f"ptest_{kind}_{dims[0]}x{dims[1]}x{dims[2]}", ["ptest", kind, *dims, 1, "{out}"], meta)

    # --- output files of the reference's writers (src/VTKDatExport.cpp) on analytically set fields ------------------
    for periodic_z in (0, 1):
        outdir = os.path.join(GOLDEN, f"export_{periodic_z}")
        os.makedirs(outdir, exist_ok=True)
        run([os.path.join(REF, "ref_dump"), "export", "8", str(periodic_z), outdir])
        print("export", periodic_z, sorted(os.listdir(outdir)))

    # --- the reference driver end to end (src/main.cpp): input file -> solution.vtk + profile*.dat --------------------
    for name, tc2 in (("mif_case1", "false"), ("mif_case2", "true")):
        outdir = os.path.join(GOLDEN, name)
        os.makedirs(outdir, exist_ok=True)
        with open(os.path.join(outdir, "input.txt"), "w") as f:
            f.write(f"Nt : 3\ndt : 1e-3\nNx : 17\nNy : 13\nNz : 15\nPy : 1\nPz : 1\ntest_case_2 : {tc2}")
        run([os.path.join(REF, "mif"), "input.txt"], cwd=outdir)
        print(name, sorted(os.listdir(outdir)))

    # --- numbers printed by the reference's own tests ----------------------------------------------
    norms = {}
    tmp = tempfile.mkdtemp(prefix="mifgolden_")
    try:
        for n, steps in ((16, 1), (32, 2), (64, 4)):
            out = run([os.path.join(REF, "full_test"), str(n), str(steps), "1"], cwd=tmp)
            norms[f"full_test {n} {steps} 1"] = [float(x) for x in out.split()]
        out = run([os.path.join(REF, "full_test"), "16", "1", "2"], np_ranks=4, cwd=tmp)
        norms["full_test 16 1 2 (4 ranks)"] = [float(x) for x in out.split()]
        for test in ("pressure_test_hn", "pressure_test_mixed", "pressure_test_nhn"):
            for n in (8, 16, 32):
                out = run([os.path.join(REF, test), str(n), "1"], cwd=tmp)
                line = [l for l in out.splitlines() if l.startswith("Errors:")][0]
                norms[f"{test} {n} 1"] = [float(x) for x in line.split()[1:]]
        for test in ("velocity_test", "velocity_test_mixed"):
            for n in (16, 32):
                out = run([os.path.join(REF, test), str(n), str(n // 16), "1"], cwd=tmp)
                norms[f"{test} {n} {n // 16} 1"] = [float(x) for x in out.split()[:3]]
    finally:
        shutil.rmtree(tmp)
    with open(os.path.join(GOLDEN, "norms.json"), "w") as f:
        json.dump(norms, f, indent=1, sort_keys=True)
    print("norms.json:", len(norms), "entries")


if __name__ == "__main__":
    main()
