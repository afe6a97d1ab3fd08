#!/usr/bin/env python3
"""bench.py -- FP64 cell-steps/s of the full projection time step (3 RK stages incl. 3 Poisson solves).

Workload (BASELINE.json configs[2], SURVEY.md section 8d "config 3"): the reference's input/input.txt
boundary-layer set-up (src/main.cpp:121-156, test case 1: domain [0,1]x[0,1]x[-1,1], Re = 1e3, v = 1 on
the face x = 1, dt = 1e-3) scaled to 512^3 cells = 513^3 pressure points, one mif::timestep per "step".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size POINTS]      our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...                                     the reference's own CPU code

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel: algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json
  cpu_baseline  oracle/_ref (the unmodified reference compiled against the MPI/FFT stand-ins) timed on
                this box's host cores on a bounded sample
  kernels       per-kernel-group milliseconds per step from the library's CUDA-event profile
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only (NCCL logs to stdout otherwise)
sys.path.insert(0, ROOT)

METRIC = "FP64 cell-steps/s, full RK timestep incl. Poisson solve"
UNIT = "cell-steps/s"
# Canonical algorithmic traffic of one cell-step: 111 FP64 words.
BYTES_PER_CELL_STEP = 888.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class GpuLocalCpus:
    """Context manager: while pinned host buffers are allocated, run the calling thread on the CPUs of the NUMA node the
    GPU hangs off (sysfs `local_cpulist` of its PCI function), so that the pages land on that node and the copy engines
    do not pull them across the socket interconnect; the previous affinity is restored on exit.  Best effort: without
    the sysfs entry, or in a cpuset that excludes those CPUs, nothing changes.  `info` goes into the JSON line."""

    def __init__(self, device_index):
        self.info = {"bound": False}
        self.cpus = None
        try:
            import torch
            prop = torch.cuda.get_device_properties(device_index)
            address = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
            base = "/sys/bus/pci/devices/" + address
            with open(base + "/local_cpulist") as f:
                text = f.read().strip()
            cpus = set()
            for part in text.split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            self.previous = os.sched_getaffinity(0)
            usable = cpus & self.previous
            node = None
            try:
                with open(base + "/numa_node") as f:
                    node = int(f.read())
            except (OSError, ValueError):
                pass
            self.info = {"bound": False, "pci": address, "numa_node": node, "local_cpus": text}
            if usable and usable != self.previous:
                self.cpus = usable
        except Exception as err:  # no sysfs, no such property, ...: leave the affinity alone
            self.info = {"bound": False, "why": type(err).__name__}

    def __enter__(self):
        if self.cpus:
            try:
                os.sched_setaffinity(0, self.cpus)
                self.info["bound"] = True
            except OSError:
                self.cpus = None
        return self

    def __exit__(self, *exc):
        if self.cpus:
            try:
                os.sched_setaffinity(0, self.previous)
            except OSError:
                pass
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler is started before the
    warm-up (nvidia-smi needs up to a second to produce its first row, longer on an 8-GPU box) and only the rows
    whose own timestamp falls inside [mark_begin(), mark_end()] are used."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device = device_index
        self.rows = []
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    @staticmethod
    def _stamp(cell):
        import datetime
        try:
            return datetime.datetime.strptime(cell, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(inside_only):
            sm, sm_max, reasons = [], None, set()
            for arrival, row in self.rows:
                if len(row) < 10:
                    continue
                stamp = self._stamp(row[0]) or arrival
                if inside_only and self.t_begin is not None and not (self.t_begin <= stamp <= (self.t_end or 1e300)):
                    continue
                try:
                    sm.append(float(row[2]))
                    sm_max = float(row[3])
                except ValueError:
                    continue
                for name, cell in zip(names, row[6:10]):
                    if cell.lower().startswith("active"):
                        reasons.add(name)
            return sm, sm_max, reasons

        sm, sm_max, reasons = collect(True)
        window = "timed region"
        if not sm:  # the timed region was shorter than one sampling period: fall back to the whole run under load
            sm, sm_max, reasons = collect(False)
            window = "whole run (no sample fell inside the timed region)"
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": sm_max, "samples": len(sm),
                "reasons": sorted(reasons), "window": window}


def host_core_count():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def default_reference_size(size):
    """The CPU arm runs the same grid as the GPU arm when this host can hold it (the reference keeps ~15 fields plus the
    2Decomp work arrays: ~20 GB at 513^3) and has the cores to finish a step in seconds; 257^3 otherwise."""
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable"))
    except (OSError, StopIteration, ValueError):
        avail_kb = 0
    need_kb = 20 * 8 * size ** 3 / 1024
    return size if (avail_kb > 2 * need_kb and host_core_count() >= 8) else min(size, 257)


def rank_grid(cores):
    ranks = 1
    while ranks * 2 <= min(cores, 64):
        ranks *= 2
    py = 1
    while py * py * 4 <= ranks:
        py *= 2
    return ranks, py, ranks // py


def run_reference_sample(points, iters, warmup, cores=None):
    """Times the unmodified reference (oracle/_ref/ref_bench) on the host cores; ranks are threads."""
    ranks, py, pz = rank_grid(cores or host_core_count())
    env = dict(os.environ, MIF_SHIM_NP=str(ranks))
    # oracle/_ref/ref_bench is built with the reference's -march=native on the build machine; if this host's CPU
    # lacks one of its instructions (SIGILL) the -march=x86-64-v3 build of the same sources is used instead.
    for build, exe in (("native", os.path.join(ROOT, "oracle", "_ref", "ref_bench")),
                       ("x86-64-v3", os.path.join(ROOT, "oracle", "_ref", "portable", "ref_bench"))):
        if not os.path.exists(exe):
            continue
        try:
            out = subprocess.run([exe, "step", str(points), str(points), str(points), str(iters), str(warmup), str(pz)],
                                 env=env, capture_output=True, text=True, timeout=1800)
            if out.returncode != 0:
                continue
            res = json.loads(out.stdout.strip().splitlines()[-1])
        except (OSError, ValueError, IndexError, subprocess.TimeoutExpired):
            continue
        res["cores"] = ranks
        res["build"] = build
        return res
    return None


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    points = args.ref_size
    t0 = time.time()
    res = run_reference_sample(points, args.steps, args.warmup)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_bench is not built (make -C oracle ref)"}))
        return 0
    value = res["cell_iters_per_s"]
    sample = (f"{points}^3 pressure points ({points - 1}^3 cells) of the same test-case-1 set-up, {args.steps} timed "
              f"steps + {args.warmup} warm-up, {res['ranks']} ranks (Py={res['Py']}, Pz={res['Pz']}) as threads of one "
              f"process, -march={res['build']}; reference stencils/transposes + in-repo FFT and MPI stand-ins (no FFTW/MPI "
              "in the image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "full projection timestep, input.txt boundary-layer set-up (test case 1)",
                   "points": [points] * 3, "dt": 1e-3, "Re": 1e3, "timed_on": "host CPU",
                   "wall_s": round(time.time() - t0, 1)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def multi_rank_parity(mif, rank, world, local_rank, py):
    """Driver-visible correctness of the N-rank path: one projection step of an Ethier-Steinman case (the manufactured
    solution of test/full_test.cpp:16-187, generators/manufsol.py:31-72) on a 9 x 257 x 257 grid, once on a single-rank
    context on rank 0's GPU and once on the N ranks (same decomposition, halo exchanges and transposes as the timed
    run); every rank compares its block, ghosts included, with the single-rank fields.  Returns the largest relative
    L-infinity difference over u, v, w, p and all ranks (the parity bar is 1e-11).  No CPU code is involved."""
    import numpy as np
    import torch
    import torch.distributed as dist

    N = [int(v) for v in os.environ.get("MIF_BENCH_PARITY_CASE", "9x257x257").split("x")]
    size, lo, Re, dt = (1.0, 1.0, 2.0), (0.0, 0.0, -1.0), 1e3, 1e-4
    h = [size[d] / (N[d] - 1) for d in range(3)]
    a, d_ = np.pi / 4.0, np.pi / 2.0

    def axis(n, direction, half):
        return lo[direction] + h[direction] * (np.arange(n) + (0.5 if half else 0.0))

    def field(component):
        ext = [N[0] + (component == 0), N[1] + (component == 1), N[2] + (component == 2)]
        x = axis(ext[0], 0, component == 0)[None, None, :]
        y = axis(ext[1], 1, component == 1)[None, :, None]
        z = axis(ext[2], 2, component == 2)[:, None, None]
        if component == 0:
            return -a * (np.exp(a * x) * np.sin(a * y + d_ * z) + np.exp(a * z) * np.cos(a * x + d_ * y))
        if component == 1:
            return -a * (np.exp(a * y) * np.sin(a * z + d_ * x) + np.exp(a * x) * np.cos(a * y + d_ * z))
        if component == 2:
            return -a * (np.exp(a * z) * np.sin(a * x + d_ * y) + np.exp(a * y) * np.cos(a * z + d_ * x))
        return np.cos(3 * x) * np.cos(2 * y) * np.cos(z) + 0.0 * (x + y + z)

    start = [np.ascontiguousarray(field(c)) for c in range(4)]
    single = [torch.zeros(tuple(f.shape), dtype=torch.float64, device="cuda") for f in start]
    if rank == 0:
        ctx1 = mif.Context(N[0], N[1], N[2], *size, *lo, Re, dt, 1, device=local_rank)
        vel, vb, vb2 = ctx1.velocity(), ctx1.velocity(), ctx1.velocity()
        p, dp = ctx1.tensor(mif.STAGGER_NONE), ctx1.tensor(mif.STAGGER_NONE)
        for t, f in zip(vel + [p], start):
            t.upload(f)
        ctx1.timestep(vel, vb, vb2, ctx1.make_bc(mif.BC_ETHIER_STEINMAN, Re), 0.0, p, dp)
        for t, dst in zip(vel + [p], single):
            dst.copy_(torch.from_numpy(t.download()))
        ctx1.close()
    ident = torch.zeros(mif.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(mif.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, src=0)
    for t in single:
        dist.broadcast(t, src=0)
    want = [t.cpu().numpy() for t in single]

    Py, Pz = py, world // py
    y_rank, z_rank = rank // Pz, rank % Pz
    ctx = mif.Context(N[0], N[1], N[2], *size, *lo, Re, dt, 1, Py=Py, Pz=Pz, rank=rank, device=local_rank,
                      comm_id=bytes(ident.cpu().numpy().tobytes()))
    first_z, first_y = mif.slab_plan(N[2], Pz), mif.slab_plan(N[1], Py)
    klo, khi = first_z[z_rank] - (z_rank > 0), first_z[z_rank + 1] + (z_rank < Pz - 1)
    jlo, jhi = first_y[y_rank] - (y_rank > 0), first_y[y_rank + 1] + (y_rank < Py - 1)

    def cut(c, arr):  # this rank's block: owner points plus one ghost towards each neighbour (src/Constants.cpp:78-94)
        return np.ascontiguousarray(arr[klo:khi + (c == 2 and z_rank == Pz - 1), jlo:jhi + (c == 1 and y_rank == Py - 1)])

    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for c, (t, f) in enumerate(zip(vel + [p], start)):
        t.upload(cut(c, f))
    ctx.timestep(vel, vb, vb2, ctx.make_bc(mif.BC_ETHIER_STEINMAN, Re), 0.0, p, dp)
    worst = 0.0
    for c, (t, w) in enumerate(zip(vel + [p], want)):
        worst = max(worst, float(np.max(np.abs(t.download() - cut(c, w)))) / float(np.max(np.abs(w))))
    path = ctx.transpose_path
    ctx.close()
    tmax = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return {"max_rel_linf": float(tmax.item()), "tolerance": 1e-11, "ok": bool(float(tmax.item()) <= 1e-11),
            "case": "one Ethier-Steinman projection step on %dx%dx%d points: %d ranks (Py=%d, Pz=%d) vs a single-rank "
                    "context on rank 0's GPU, every rank's block incl. ghosts, u v w p" % (N[0], N[1], N[2], world, Py, Pz),
            "transpose_path": {0: "none", 1: "peer-memory fused", 2: "NCCL all-to-all", 3: "pencil box exchanges"}.get(path, "?")}


def poisson_workload(args):
    """BASELINE.json configs[1]: solve_pressure_equation_homogeneous_periodic alone (test/pressure_test_mixed.cpp:
    domain [0, 2 pi]^3, z periodic, dt = 1), here on a cubic 513^3-point grid with a seeded random velocity (the cost
    of a solve does not depend on the data).  One "step" = divergence + 5 sweep launches + ghost refresh."""
    import numpy as np
    import torch

    import mif_b200 as mif

    torch.cuda.set_device(0)
    mif.lib()
    N = args.size
    two_pi = 2.0 * 3.14159265358979323846
    ctx = mif.Context(N, N, N, two_pi, two_pi, two_pi, 0.0, 0.0, 0.0, 1e3, 1.0, 1, periodic=(False, False, True))
    vel = ctx.velocity()
    rng = np.random.default_rng(1234)
    for t in vel:
        sx, sy, sz = t.shape
        t.upload(rng.uniform(-1, 1, (sz, sy, sx)))
    p = ctx.tensor(mif.STAGGER_NONE)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    steps, warmup = args.steps, max(args.warmup, 3)
    for _ in range(warmup):
        ctx.solve_pressure(p, vel, 1.0)
    ctx.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.5)
    sampler.mark_begin()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        ctx.solve_pressure(p, vel, 1.0)
    ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    ctx.profile_enable(True)
    for _ in range(steps):
        ctx.solve_pressure(p, vel, 1.0)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    kernels = {k: round(v / steps, 4) for k, (v, n) in prof.items() if n}
    cells = float(N - 1) ** 3
    value = cells * steps / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    # divergence 3R + 1W, six sweeps x (1R + 1W) = 16 FP64 words = 128 B per cell-solve
    bytes_per_solve = 128.0
    line = {
        "metric": "FP64 cell-solves/s, spectral pressure Poisson solve alone", "value": value, "unit": "cell-solves/s",
        "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pressure_test_mixed-type Poisson solve, {N}^3 points, x/y Neumann (DCT-I), z periodic (real FFT)",
                   "points": [N, N, N], "l2": "inputs larger than L2", "finite": bool(np.isfinite(p.download()).all())},
        "roofline": {"bound": "hbm", "kernel": "whole solve (divergence + 5 sweep launches)",
                     "achieved": round(value * bytes_per_solve / 1e9, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(value * bytes_per_solve / 1e9 / peak, 4), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_cell_solve": bytes_per_solve},
        "gpu_launches": int(launches), "kernels": kernels, "clocks": clocks,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=513, help="pressure points per direction (513 = 512^3 cells)")
    ap.add_argument("--dims", type=int, nargs=3, default=None, metavar=("NX", "NY", "NZ"),
                    help="explicit pressure points per direction (kernel studies on non-cubic grids; single GPU only)")
    ap.add_argument("--ref-size", type=int, default=0,
                    help="points per direction of the CPU reference sample (default: --size, i.e. the same 513^3 grid, when the "
                         "host has the ~20 GB and >= 8 cores it takes; 257 otherwise)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): 512^3 cells per GPU, the grid doubles in z, y, x; strong: the grid of --dims / --size "
                         "is fixed and split over the GPUs (BASELINE.json configs[3]: --scaling strong --size 1025)")
    ap.add_argument("--workload", default="timestep", choices=["timestep", "poisson", "aniso"],
                    help="timestep: the north-star metric (default); poisson: BASELINE.json configs[1], the pressure "
                         "solve alone on a pressure_test_mixed-type grid (x, y Neumann / DCT-I, z periodic / real FFT); "
                         "aniso: BASELINE.json configs[4], the full time step on the anisotropic grid weak-scaled in x, "
                         "256 x 512 x 512 cells per GPU (8 GPUs: 2048 x 512 x 512 cells, 2049-point x lines)")
    ap.add_argument("--py", type=int, default=1,
                    help="Py of a Py x Pz pencil decomposition (Pz = GPUs / Py); default 1 = z slabs, the fast configuration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.ref_size <= 0:
        args.ref_size = default_reference_size(args.size)
    if args.impl == "reference":
        return reference_arm(args)
    if args.py < 1 or int(os.environ.get("WORLD_SIZE", "1")) % args.py != 0:
        raise SystemExit("--py must divide the number of GPUs")
    if args.workload == "poisson":
        return poisson_workload(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import mif_b200 as mif

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    mif.lib()  # fails loudly if libmifgpu.so is missing
    if mif.lib().mifgpu_real_bytes() != 8:
        sys.exit("bench.py measures BASELINE.json's FP64 metric: MIFGPU_LIB names the float build of the library, unset it")

    N = args.size
    steps, warmup = args.steps, max(args.warmup, 3)
    dt = 1e-3
    total_steps = warmup + 2 * steps + 16
    # Weak scaling (default) with 512^3 cells per GPU (for the default size): the grid doubles in z, then y, then x, so
    # 8 GPUs run the 1024^3-cell problem of BASELINE.json configs[3].  Strong scaling: the grid is fixed.  The domain is
    # split into z slabs (Py = 1, Pz = GPUs) inside libmifgpu: NCCL halo exchange + fused peer-memory transposes.
    cells_1d = N - 1
    mult = [1, 1, 1]
    if args.scaling == "weak":
        for level in range(max(world.bit_length() - 1, 0)):
            mult[2 - level % 3] *= 2
    dims = [cells_1d * m + 1 for m in mult]
    if args.workload == "aniso":
        # SURVEY.md section 8d "config 5": constant 256 x 512 x 512 cells per GPU, x is never split, the GPUs split z
        dims = [256 * world + 1, 513, 513]
    if args.dims is not None:
        if world != 1 and args.scaling != "strong":
            raise SystemExit("--dims on several GPUs needs --scaling strong")
        dims = list(args.dims)
    # The cells keep the size of the 512^3 case (dx = dy = 1/512, dz = 2/512), so dt = 1e-3 is as stable as there; the
    # domain grows with the grid instead and ends at x = 1, where test case 1 has its lid: [1 - Lx, 1] x [0, Ly] x
    # [-Lz/2, Lz/2].  All extents are multiples of 2^-9, so the face coordinate min_x + dx * (Nx - 1) is exactly 1.
    mult = [(d - 1) / 512.0 for d in dims]
    comm_id = None
    if world > 1:
        ident = torch.zeros(mif.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident.copy_(torch.frombuffer(bytearray(mif.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, src=0)
        comm_id = bytes(ident.cpu().numpy().tobytes())
    # src/main.cpp:121-131, test case 1.
    parity = multi_rank_parity(mif, rank, world, local_rank, args.py) if world > 1 else None
    ctx = mif.Context(dims[0], dims[1], dims[2], 1.0 * mult[0], 1.0 * mult[1], 2.0 * mult[2], 1.0 - mult[0], 0.0, -1.0 * mult[2],
                      1e3, dt * total_steps, total_steps, Py=args.py, Pz=world // args.py, rank=rank, device=local_rank,
                      comm_id=comm_id)
    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    bc = ctx.make_bc(mif.BC_TEST_CASE_1, 1e3)
    # velocity.set(exact(t=0), include_border=true) (src/main.cpp:144-146): v = 1 on the face x = 1, else 0.  The face is
    # selected by index here; the library's boundary data finds it by coordinate, like
    # include/TestCaseBoundaries.h:18-35, and both agree because the domain ends at x = 1 exactly (checked below).
    host = []
    gpu_local = GpuLocalCpus(local_rank)  # pinned host buffers on the GPU's own NUMA node (e2e moves 8.65 GB per job)
    with gpu_local:
        for t in vel + [p]:
            sx, sy, sz = t.shape
            arr = torch.zeros((sz, sy, sx), dtype=torch.float64).pin_memory()
            host.append(arr)
    host[1][:, :, dims[0] - 1] = 1.0
    for t, arr in zip(vel + [p], host):
        t.upload(arr.numpy())
    ctx.apply_bc(vel, bc, 0.0)
    lid_ok = bool((vel[1].download()[:, :, dims[0] - 1] == 1.0).all())  # the device's own boundary data put v = 1 there too

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    cells = float(dims[0] - 1) * float(dims[1] - 1) * float(dims[2] - 1) / world  # per GPU
    step_index = [0]

    def one_step():
        ctx.timestep(vel, vb, vb2, bc, step_index[0] * dt, p, dp)
        step_index[0] += 1

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        one_step()
    barrier()

    sampler.mark_begin()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        one_step()
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    value = world * cells * steps / (elapsed_ms * 1e-3)

    # Per-kernel-group times (CUDA events on the launching stream) over the same number of steps.
    ctx.profile_enable(True)
    for _ in range(steps):
        one_step()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    kernels = {k: round(ms / steps, 4) for k, (ms, n) in prof.items() if n}
    peak, peak_src = measured_peaks()
    points = float(dims[0]) * float(dims[1]) * float(dims[2]) / world  # pressure points per GPU
    # Dominant kernel = the Poisson sweep kernel (5 launches per solve, 15 per step).  Algorithmic bytes
    # per launch: one read + one write of every pressure point = 16 B * N^3 (SURVEY.md section 8d:
    # "Poisson 6 sweeps x (1R+1W)"; the fused z launch does the work of two sweeps with the same 16 B).
    sweep_ms = sum(ms for k, (ms, n) in prof.items() if k.startswith("sweep_"))
    sweep_launches = sum(n for k, (ms, n) in prof.items() if k.startswith("sweep_"))
    per_launch_ms = sweep_ms / max(sweep_launches, 1)
    achieved = 16.0 * points / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    step_profile_ms = sum(ms for ms, n in prof.values()) / steps
    # DRAM bytes per launch of the same kernels from the committed ncu --set full capture (513^3, one GPU only).
    traffic, traffic_src = None, None
    ncu_summary = os.path.join(ROOT, "profiles", "r02_ncu_sweep_kernels.json")
    if world == 1 and dims == [513, 513, 513] and os.path.exists(ncu_summary):
        with open(ncu_summary) as f:
            rows = [r for r in json.load(f) if "dct" in r["kernel"]]
        if rows:
            traffic = round(sum(r["dram_read_GB"] + r["dram_write_GB"] for r in rows) / len(rows) * 1e9)
            traffic_src = ("profiles/r02_ncu_sweep_kernels.json (dram__bytes_read.sum + dram__bytes_write.sum, mean of the %d "
                           "sweep launches of one solve)" % len(rows))
    sweep_kernel_names = ("tma_dct512_kernel<0|1|2> (y forward / y inverse / fused z, TMA-staged) and x_dct512_kernel<0|1> "
                          "(x forward / inverse)") if dims == [513, 513, 513] and world == 1 else \
        "Poisson sweep kernels of this grid (csrc/mif_poisson.cu: launch_sweep picks them by line length)"
    roofline = {
        "bound": "hbm", "kernel": sweep_kernel_names + ", 15 launches per step",
        "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 16.0 * points,
        "peak_source": peak_src, "avg_launch_ms": round(per_launch_ms, 4),
        "share_of_step": round(sweep_ms / steps / step_profile_ms, 4) if step_profile_ms else None,
        "whole_step": {"algorithmic_bytes_per_cell_step": BYTES_PER_CELL_STEP,
                       "achieved_GBs": round(value / world * BYTES_PER_CELL_STEP / 1e9, 1),
                       "frac": round(value / world * BYTES_PER_CELL_STEP / 1e9 / peak, 4)},
    }

    # NVLink traffic of the Y<->Z pencil transposes: per all-to-all every GPU sends
    # 8 B * points/P * (P-1)/P; slab decomposition: 2 per solve, 6 per step.  With the peer-memory path these bytes are
    # the remote stores of the forward y sweep and of the fused z sweep, so their launch times bound the link rate.
    nvlink = None
    if world > 1:
        sent = 6.0 * 8.0 * points * (world - 1) / world
        carrier_ms = sum(kernels.get(k, 0.0) for k in ("sweep_y_fwd", "sweep_z_fused", "transpose_alltoall"))
        nvlink = {"bytes_sent_per_gpu_per_step": sent, "peak_GBs_per_direction": 900.0,
                  "GBs_over_carrier_kernels": round(sent / (carrier_ms * 1e-3) / 1e9, 1) if carrier_ms > 0 else None,
                  "carrier_kernels_ms_per_step": round(carrier_ms, 4),
                  "GBs_over_whole_step": round(sent / (elapsed_ms / steps * 1e-3) / 1e9, 1),
                  "what": "6 pencil transposes per step, 8 B * points/P * (P-1)/P each way per transpose; carried by the "
                          "forward y sweep and the fused z sweep (peer stores) or by the NCCL all-to-all (fallback)"}

    # End to end through the C ABI with HOST buffers: a stream of independent single-step jobs.  Every job uploads its
    # u, v, w, p from pinned host memory (mifgpu_tensor_upload_async), runs mifgpu_timestep and downloads u, v, w, p
    # (mifgpu_tensor_download_async); all host <-> device copies are inside the timed region.  Three device field sets
    # rotate, so the copies of job n-1 / n+1 run on their own streams while job n computes and PCIe is busy in both
    # directions.  Skipped (null) when a field is too large to double-buffer in pinned host memory.
    e2e = None
    field_bytes = sum(int(np.prod(t.shape)) * 8 for t in vel + [p])
    if not args.no_e2e and field_bytes <= 6e9:
        e2e_steps = max(6, min(steps, 9))
        # three sets: while set A computes, B uploads and C downloads (with two, the upload into a set would have to wait
        # for that set's own download and the two PCIe directions would take turns)
        n_sets = max(1, int(os.environ.get("MIF_BENCH_E2E_SETS", "3")))
        sets = [(vel, p)] + [(ctx.velocity(), ctx.tensor(mif.STAGGER_NONE)) for _ in range(n_sets - 1)]
        with gpu_local:
            out = [torch.zeros(tuple(a.shape), dtype=torch.float64).pin_memory() for a in host]
        for v2, p2 in sets[1:]:
            for t, arr in zip(v2 + [p2], host):
                t.upload(arr.numpy())

        def one_job(i):
            v_i, p_i = sets[i % n_sets]
            for t, arr in zip(v_i + [p_i], host):
                t.upload_async(arr.numpy())
            ctx.timestep(v_i, vb, vb2, bc, step_index[0] * dt, p_i, dp)
            step_index[0] += 1
            for t, arr in zip(v_i + [p_i], out):
                t.download_async(arr.numpy())

        for i in range(n_sets):
            one_job(i)  # warm-up: copy streams, staging buffers
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            one_job(i)
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            tmax = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e2e_s = float(tmax.item())
        e2e = {"value": world * cells * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": field_bytes,
               "d2h_bytes_per_step": field_bytes, "steps": e2e_steps, "ms_per_step": round(1e3 * e2e_s / e2e_steps, 3),
               "device_field_sets": n_sets, "host_buffers": gpu_local.info,
               "result_finite": bool(np.isfinite(out[1].numpy()).all()),
               "what": "per step (one job): async upload of u,v,w,p from pinned host memory, mifgpu_timestep, async download "
                       "of u,v,w,p; consecutive jobs are independent and rotate through three device field sets, so "
                       "their copies overlap each other and the kernels (separate H2D / D2H streams, event ordered)"}
    elif not args.no_e2e:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": field_bytes, "d2h_bytes_per_step": field_bytes,
               "what": "skipped: %.1f GB of fields per rank is more than this bench double-buffers in pinned host memory" % (field_bytes / 1e9)}

    finite = bool(np.isfinite(vel[1].download()).all())
    transpose_names = {0: "none", 1: "peer-memory fused", 2: "NCCL all-to-all", 3: "pencil box exchanges"}
    transpose_path = transpose_names.get(ctx.transpose_path, "?")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res = run_reference_sample(args.ref_size, 2, 1)
        except Exception as exc:  # the optional CPU leg must never cost the GPU line
            print(f"bench.py: cpu_baseline leg failed: {exc!r}", file=sys.stderr)
            res = None
        if res is not None:
            cpu_baseline = {
                "value": res["cell_iters_per_s"], "unit": UNIT, "cores": res["cores"], "kind": "reference",
                "sample": (f"{args.ref_size}^3 points of the same set-up, 2 timed steps + 1 warm-up, {res['ranks']} ranks "
                           f"(Py={res['Py']}, Pz={res['Pz']}) as threads, -march={res['build']}; unmodified reference "
                           "sources + in-repo FFT/MPI stand-ins (oracle/_ref)")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("full projection timestep, input.txt boundary-layer set-up (test case 1) at "
                                    f"{dims[0] - 1}x{dims[1] - 1}x{dims[2] - 1} cells (" +
                                    ("%dx%dx%d cells in total, split over %d GPU(s))" % (dims[0] - 1, dims[1] - 1, dims[2] - 1, world)
                                     if args.scaling == "strong" else
                                     (f"{cells_1d}^3" if args.workload != "aniso" else "256x512x512") + " cells per GPU)")),
                       "points": dims, "dt": dt, "Re": 1e3,
                       "parallelism": "single GPU" if world == 1 else
                       (f"pencils Py={args.py} Pz={world // args.py}: NCCL halos (y sheets, z planes), 2Decomp transposes as "
                        "grouped send/recv box exchanges" if args.py > 1 else
                        f"z slabs Py=1 Pz={world}: NCCL plane halos; Y<->Z pencil transposes: " +
                        ("fused into the y/z sweeps as NVLink peer-memory stores (no separate all-to-all)"
                         if ctx.transpose_path == 1 else "grouped NCCL send/recv all-to-all")),
                       "transpose_path": transpose_path, "lid_on_last_x_face": lid_ok,
                       "l2": "inputs larger than L2 (each field %.2f GB)" % (points * 8 / 1e9), "finite": finite},
            "roofline": roofline, "nvlink": nvlink, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "parity_vs_single_rank": parity,
            "kernels": kernels, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
