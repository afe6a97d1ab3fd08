"""Worker of the multi-GPU tests: run under torchrun (one rank per GPU).  Each rank builds its slab of a golden
case (tests/golden, dumped from the single-rank reference; all-Neumann results are decomposition independent), runs
the time steps through the C ABI with Pz = world_size and compares its slab of the result with the golden fields.
For the z-periodic case the goldens come from the reference run with the same number of ranks (rank-specific
dumps), because periodic time stepping is decomposition dependent in the reference."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mif_b200 as mif  # noqa: E402


def local_slab(arr, klo, khi):
    return np.ascontiguousarray(arr[klo:khi])


def main():
    case = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    if torch.cuda.is_available():  # (the CPU run of this worker under the SIMT interpreter has no device to select)
        torch.cuda.set_device(local_rank)
    ids = [mif.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)

    if case.startswith("pz:"):
        # Periodic z distributed over the ranks (neighbours wrap around, src/Constants.cpp:98-101): the pressure solve and the
        # halo exchange that follows it (src/PressureTensor.cpp:30-33) against the oracle's single-rank solve, GHOST PLANES
        # INCLUDED -- on two ranks both neighbours are the same peer and the two planes must not be swapped.
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mif_oracle as mo
        N = [int(v) for v in case[3:].split("x")]
        periodic = (False, False, True)
        grid = mo.Grid(N[0], N[1], N[2], 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-3, 4, periodic=periodic)
        rng = np.random.default_rng(77)
        host = [rng.uniform(-1, 1, grid.shape(c)) for c in range(3)]
        want = grid.solve_pressure(*host, 0.37)  # single rank: planes 0 and -1 are the periodic ghosts
        ctx = mif.Context(N[0], N[1], N[2], 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-3, 4, Py=1, Pz=world, rank=rank, periodic=periodic,
                          device=local_rank, comm_id=ids[0])
        first = mif.slab_plan(N[2] - 1, world)  # owner planes of the N_z - 1 periodic points
        klo, khi = first[rank], first[rank + 1] + 2  # one ghost plane on each side, in the single-rank array's numbering
        vel = ctx.velocity()
        for c, (t, h) in enumerate(zip(vel, host)):
            extra = 1 if (c == 2 and rank == world - 1) else 0  # the z-staggered component has one more plane on the last rank
            assert t.shape == h[klo:khi + extra].shape[::-1], (c, t.shape, h[klo:khi].shape)
            t.upload(np.ascontiguousarray(h[klo:khi + extra]))
        p = ctx.tensor(mif.STAGGER_NONE)
        ctx.solve_pressure(p, vel, 0.37)
        got = p.download()
        worst = float(np.max(np.abs(got - want[klo:khi]))) / float(np.max(np.abs(want)))
        errs = [None] * world
        dist.all_gather_object(errs, worst)
        ctx.close()
        if rank == 0:
            print(json.dumps({"case": case, "world": world, "Py": 1, "Pz": world, "max_rel_err": max(errs)}))
        dist.destroy_process_group()
        return 0 if max(errs) <= 1e-11 else 1
    if case.startswith("mr:"):
        # A run of the reference itself on this many ranks (oracle/make_golden.py --multi-rank): every rank uploads ITS local
        # arrays of the reference's run and must reproduce the reference's local arrays after every step, ghosts included.
        data = np.load(os.path.join(ROOT, "tests", "golden", case[3:] + ".npz"))
        meta = json.loads(str(data["meta"]))
        assert meta["ranks"] == world, (meta["ranks"], world)
        N, Py, Pz = meta["N"], meta["Py"], meta["Pz"]
        periodic = [bool(p) for p in meta["periodic"]]
        ctx = mif.Context(N[0], N[1], N[2], meta["x_size"], meta["y_size"], meta["z_size"], *meta["min"], meta["Re"],
                          meta["final_time"], meta["steps"], Py=Py, Pz=Pz, rank=rank, periodic=periodic, device=local_rank,
                          comm_id=ids[0])
        velocity_only = meta["kind"] == "vtest"
        names = "uvw" if velocity_only else "uvwp"
        y_rank = rank // Pz
        ctx_prev_y = Py > 1 and (y_rank > 0 or periodic[1])      # this rank has a y neighbour below / above
        ctx_next_y = Py > 1 and (y_rank < Py - 1 or periodic[1])
        vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
        p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
        tensors = vel if velocity_only else vel + [p]
        for t, name in zip(tensors, names):
            host = data[f"{name}_s0_r{rank}"]
            assert t.shape == host.shape[::-1], (name, t.shape, host.shape)
            t.upload(host)
        kind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1,
                "test_case_2": mif.BC_TEST_CASE_2, "velocity_test": mif.BC_VELOCITY_TEST}[meta["bc"]]
        bc = ctx.make_bc(kind, meta["Re"])
        dt = meta["final_time"] / meta["steps"]
        worst = 0.0
        for step in range(meta["steps"]):
            if velocity_only:
                ctx.timestep_velocity(vel, vb, vb2, bc, step * dt)
            else:
                ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp)
            ctx.synchronize()
            # scales: the largest value of the field over ALL ranks; a velocity component that is zero in exact arithmetic
            # is normalised by the largest component (see tests/test_gpu_golden.py)
            def field_max(name):
                return max(float(np.max(np.abs(data[f"{name}_s{step + 1}_r{r}"]))) for r in range(world))
            vmax = max(field_max(c) for c in "uvw")
            for t, name in zip(tensors, names):
                ref = data[f"{name}_s{step + 1}_r{rank}"]
                scale = max(field_max(name), 1e-6 * vmax if name in "uvw" else 0.0)
                err = np.abs(t.download() - ref)
                if Py > 1 and "MIFGPU_REFERENCE_HALOS" not in os.environ:
                    # The reference unpacks a received y sheet into the interior of the ghost row only, i and k borders
                    # keep what they held (src/StaggeredTensor.cpp:145-149,158-162); the library refreshes whole rows
                    # (equality with the single-rank result).  No stencil reads those points.
                    for j in ([0] if ctx_prev_y else []) + ([ref.shape[1] - 1] if ctx_next_y else []):
                        err[0, j, :] = err[-1, j, :] = 0.0
                        err[:, j, 0] = err[:, j, -1] = 0.0
                worst = max(worst, float(np.max(err)) / scale)
        errs = [None] * world
        dist.all_gather_object(errs, worst)
        ctx.close()
        if rank == 0:
            print(json.dumps({"case": case, "world": world, "Py": Py, "Pz": Pz, "max_rel_err": max(errs)}))
        dist.destroy_process_group()
        return 0 if max(errs) <= float(os.environ.get("MIF_WORKER_TOL", "1e-11")) else 1
    if case.startswith("es:"):
        # No golden file: the oracle (oracle/mif_oracle.c, pinned to the reference) computes the single-rank result
        # of one Ethier-Steinman step on the given grid; used for grids large enough to take the peer-memory path.
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mif_oracle as mo
        N = [int(v) for v in case[3:].split("x")]
        meta = dict(N=N, x_size=1.0, y_size=1.0, z_size=2.0, min=[0.0, 0.0, -1.0], Re=1e3, final_time=1e-4, steps=1,
                    periodic=[0, 0, 0], bc="ethier_steinman")
        grid = mo.Grid(N[0], N[1], N[2], 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, 1e3, 1e-4, 1)
        vel0 = list(grid.set_velocity(mo.BC_ETHIER_STEINMAN, 0.0))
        rng = np.random.default_rng(11)
        p0 = rng.uniform(-1, 1, grid.shape(3))
        data = {c + "_s0": a.copy() for c, a in zip("uvw", vel0)}
        data["p_s0"] = p0.copy()
        buf, buf2, dp0 = [grid.zeros(c) for c in range(3)], [grid.zeros(c) for c in range(3)], grid.zeros(3)
        grid.timestep(mo.BC_ETHIER_STEINMAN, 0.0, vel0, buf, buf2, p0, dp0)
        data.update({c + "_s1": a for c, a in zip("uvwp", vel0 + [p0])})
    else:
        data = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
        meta = json.loads(str(data["meta"]))
    N = meta["N"]
    periodic = [bool(p) for p in meta["periodic"]]
    assert not periodic[2], "slab tests use the decomposition-independent (non-periodic z) goldens"
    # MIF_PY > 1 selects a Py x Pz pencil decomposition (rank = y_rank * Pz + z_rank, src/Constants.cpp:68) instead of
    # z slabs.
    Py = int(os.environ.get("MIF_PY", "1"))
    assert world % Py == 0
    Pz = world // Py
    y_rank, z_rank = rank // Pz, rank % Pz
    ctx = mif.Context(N[0], N[1], N[2], meta["x_size"], meta["y_size"], meta["z_size"], *meta["min"], meta["Re"],
                      meta["final_time"], meta["steps"], Py=Py, Pz=Pz, rank=rank, periodic=periodic,
                      device=local_rank, comm_id=ids[0])
    # y / z range of this rank's local arrays inside the global (single-rank) arrays: owner points plus one ghost
    # towards each neighbour (src/Constants.cpp:78-94); the staggered component has one more on the last rank.
    first = mif.slab_plan(N[2], Pz)
    klo = first[z_rank] - (1 if z_rank > 0 else 0)
    khi = first[z_rank + 1] + (1 if z_rank < Pz - 1 else 0)
    first_y = mif.slab_plan(N[1], Py)
    jlo = first_y[y_rank] - (1 if y_rank > 0 else 0)
    jhi = first_y[y_rank + 1] + (1 if y_rank < Py - 1 else 0)

    def cut(name, arr):
        extra_k = 1 if (name == "w" and z_rank == Pz - 1) else 0
        extra_j = 1 if (name == "v" and y_rank == Py - 1) else 0
        return np.ascontiguousarray(arr[klo:khi + extra_k, jlo:jhi + extra_j])

    vel, vb, vb2 = ctx.velocity(), ctx.velocity(), ctx.velocity()
    p, dp = ctx.tensor(mif.STAGGER_NONE), ctx.tensor(mif.STAGGER_NONE)
    for t, name in zip(vel + [p], "uvwp"):
        host = cut(name, data[name + "_s0"])
        assert t.shape == host.shape[::-1], (name, t.shape, host.shape)
        t.upload(host)
    kind = {"ethier_steinman": mif.BC_ETHIER_STEINMAN, "test_case_1": mif.BC_TEST_CASE_1}[meta["bc"]]
    bc = ctx.make_bc(kind, meta["Re"])
    dt = meta["final_time"] / meta["steps"]
    worst = 0.0
    for step in range(meta["steps"]):
        ctx.timestep(vel, vb, vb2, bc, step * dt, p, dp)
        ctx.synchronize()
        vmax = max(float(np.max(np.abs(data[f"{c}_s{step + 1}"]))) for c in "uvw")
        for t, name in zip(vel + [p], "uvwp"):
            ref_full = data[f"{name}_s{step + 1}"]
            ref = cut(name, ref_full)
            scale = max(float(np.max(np.abs(ref_full))), 1e-6 * vmax if name in "uvw" else 0.0)
            err = float(np.max(np.abs(t.download() - ref))) / scale
            worst = max(worst, err)
    errs = [None] * world
    dist.all_gather_object(errs, worst)
    ctx.close()
    if rank == 0:
        print(json.dumps({"case": case, "world": world, "Py": Py, "Pz": Pz, "max_rel_err": max(errs)}))
    dist.destroy_process_group()
    return 0 if max(errs) <= float(os.environ.get("MIF_WORKER_TOL", "1e-11")) else 1  # the float build's cases pass their own bound


if __name__ == "__main__":
    sys.exit(main())
