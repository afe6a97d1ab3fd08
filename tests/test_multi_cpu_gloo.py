"""CPU, world_size 2 over gloo: the host-side logic of the slab decomposition -- the block distribution the library
exports (mifgpu_slab_plan, host only), the slab <-> z-pencil all-to-all with the library's buffer ordering
([dest][z_local][y in dest's range][x] on the slab side, [z][y_local][x] on the pencil side) and the plane halo
rule (plane 1 -> previous rank's last plane, plane sz-2 -> next rank's plane 0) -- checked against a global array."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _all_to_all(recv, send, rank, world):
    """Grouped send/recv all-to-all, like the library's ncclGroupStart .. ncclSend/ncclRecv .. ncclGroupEnd."""
    recv[rank].copy_(send[rank])
    ops = []
    for r in range(world):
        if r != rank:
            ops.append(dist.P2POp(dist.isend, send[r], r))
            ops.append(dist.P2POp(dist.irecv, recv[r], r))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def _worker(rank, world, port, nx, ny, nz, results):
    sys.path.insert(0, ROOT)
    import mif_b200 as mif
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ylo, zlo = mif.slab_plan(ny, world), mif.slab_plan(nz, world)
        rng = np.random.default_rng(42)
        glob = rng.uniform(-1, 1, (nz, ny, nx))  # same on every rank
        slab = glob[zlo[rank]:zlo[rank + 1]]     # this rank's owner planes, all y
        nz_me, ny_me = zlo[rank + 1] - zlo[rank], ylo[rank + 1] - ylo[rank]
        # pack per destination: [dest][z_local][y in dest's range][x]
        send = [torch.from_numpy(np.ascontiguousarray(slab[:, ylo[r]:ylo[r + 1], :]).reshape(-1)) for r in range(world)]
        recv = [torch.empty((zlo[r + 1] - zlo[r]) * ny_me * nx, dtype=torch.float64) for r in range(world)]
        _all_to_all(recv, send, rank, world)
        # the block of source r lands at planes zlo[r]..zlo[r+1] of the pencil [z][y_local][x]
        pencil = np.concatenate([recv[r].numpy().reshape(zlo[r + 1] - zlo[r], ny_me, nx) for r in range(world)], axis=0)
        ok_fwd = np.array_equal(pencil, glob[:, ylo[rank]:ylo[rank + 1], :])
        # and back: send plane ranges straight out of the pencil, unpack per source y range
        send = [torch.from_numpy(np.ascontiguousarray(pencil[zlo[r]:zlo[r + 1]]).reshape(-1)) for r in range(world)]
        recv = [torch.empty(nz_me * (ylo[r + 1] - ylo[r]) * nx, dtype=torch.float64) for r in range(world)]
        _all_to_all(recv, send, rank, world)
        back = np.empty_like(slab)
        for r in range(world):
            back[:, ylo[r]:ylo[r + 1], :] = recv[r].numpy().reshape(nz_me, ylo[r + 1] - ylo[r], nx)
        ok_bwd = np.array_equal(back, slab)
        # halo rule on a ghosted local array (one ghost plane towards each neighbour)
        klo = zlo[rank] - (1 if rank > 0 else 0)
        khi = zlo[rank + 1] + (1 if rank < world - 1 else 0)
        local = glob[klo:khi].copy()
        want = local.copy()
        if rank > 0:
            local[0] = np.nan
        if rank < world - 1:
            local[-1] = np.nan
        ops = []
        if rank > 0:
            ops += [dist.P2POp(dist.isend, torch.from_numpy(local[1].copy()), rank - 1),
                    dist.P2POp(dist.irecv, torch.from_numpy(local[0]), rank - 1)]
        if rank < world - 1:
            ops += [dist.P2POp(dist.isend, torch.from_numpy(local[-2].copy()), rank + 1),
                    dist.P2POp(dist.irecv, torch.from_numpy(local[-1]), rank + 1)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        ok_halo = np.array_equal(local, want)
        results[rank] = (ok_fwd, ok_bwd, ok_halo, ylo, zlo)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny,nz", [(5, 7, 9), (6, 13, 10), (4, 8, 8)])
def test_slab_pencil_exchange_two_ranks(nx, ny, nz):
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    port = 29611 + nx + ny
    mp.spawn(_worker, args=(world, port, nx, ny, nz, results), nprocs=world, join=True)
    for rank in range(world):
        ok_fwd, ok_bwd, ok_halo, ylo, zlo = results[rank]
        assert ok_fwd and ok_bwd and ok_halo, (rank, ok_fwd, ok_bwd, ok_halo)
        assert ylo[-1] == ny and zlo[-1] == nz


def test_slab_plan_matches_reference_distribution(mif):
    # src/Constants.cpp:78-79: owner = n / P + (rank < n % P); 513 points over 8 ranks = 65, 64, ..., 64
    first = mif.slab_plan(513, 8)
    assert [b - a for a, b in zip(first, first[1:])] == [65] + [64] * 7
    assert mif.slab_plan(10, 3) == [0, 4, 7, 10]
    assert mif.slab_plan(7, 1) == [0, 7]
