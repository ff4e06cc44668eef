"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 4) `gloo` process groups, no GPU.

Checks, across real processes: (i) the brick decomposition partitions a particle set exactly (all_reduce of counts and
id checksums), (ii) every rank's locally computed SEND list to a peer equals the peer's RECV list from it — the property
that lets the device exchange run without a handshake (csrc/comm.cuh), (iii) ghost images produced through that plan
reproduce the single-domain periodic ghost layer, (iv) the weak-scaling brick workload has globally unique ids."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exadem_b200 import parallel, workloads


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, periodic, ret, balanced=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pg = parallel.process_grid(world)
        side = [10.0 * pg[d] for d in range(3)]
        wl = workloads.spheres_box(tuple(int(s) for s in side), jitter=0.03, seed=7, periodic=periodic)
        glo, ghi = wl.lo, wl.hi
        cuts = nc_axis = None
        if balanced:   # a heap on the low side of x and y: bricks cut where the 3 N + 1 cost balances, different cell grids per brick
            r = wl.particles.vec("r")
            keep = ((r[:, 0] < 0.5 * side[0]) & (r[:, 1] < 0.6 * side[1])) | (np.random.default_rng(3).random(wl.particles.n) < 0.3)
            wl.particles = parallel.select(wl.particles, keep)
            cuts, gcells, cell_cuts = parallel.balanced_decomposition(wl.particles.vec("r"), glo, ghi, 2.5, pg, min_cells=1)
            assert any(len(set(np.diff(c))) > 1 for c in cell_cuts), cell_cuts
            nc_axis = parallel.cells_per_brick(cuts, 2.5, pg)
        pos = wl.particles.vec("r")
        mine = parallel.owner_rank(pos, glo, ghi, pg, cuts) == rank
        lo, hi = parallel.brick_bounds(glo, ghi, pg, rank, cuts)
        assert all(lo[d] <= pos[mine][:, d].min() and pos[mine][:, d].max() < hi[d] + 1e-12 for d in range(3))
        # (i) exact partition
        t = torch.tensor([int(mine.sum()), int(wl.particles.id[mine].sum())], dtype=torch.int64)
        dist.all_reduce(t)
        assert int(t[0]) == wl.particles.n and int(t[1]) == int(wl.particles.id.sum())
        # (ii) send list == the peer's recv list
        nc, cs = parallel.grid_dims(lo, hi, 2.5)
        if nc_axis is not None:
            assert nc == [nc_axis[d][parallel.rank_coords(rank, pg)[d]] for d in range(3)]
        sp = parallel.send_plan(glo, ghi, periodic, pg, rank, nc, nc_axis)
        rp = parallel.recv_plan(glo, ghi, periodic, pg, rank, nc, nc_axis)
        plans = [None] * world
        dist.all_gather_object(plans, {"send": sp, "recv": rp})
        for other in range(world):
            if other == rank:
                continue
            assert plans[other]["recv"].get(rank, []) == sp.get(other, []), f"rank {rank}->{other}: send list != recv list"
            assert plans[other]["send"].get(rank, []) == rp.get(other, [])
        # (iii) ghosts through the plan == single-domain periodic images that fall in my ghost shell
        cell = np.floor((pos[mine] - np.array(lo)) / np.array(cs)).astype(int) + 1
        local = {}
        for k, c in enumerate(map(tuple, np.clip(cell, 1, np.array(nc)))):
            local.setdefault(c, []).append(pos[mine][k])
        cells_all = [None] * world
        dist.all_gather_object(cells_all, local)
        got = []
        for (g, src_rank, s, sh) in parallel.ghost_plan(glo, ghi, periodic, pg, rank, nc, nc_axis):
            for r in cells_all[src_rank].get(tuple(s), []):
                got.append(r + np.array(sh))
        got = np.array(got) if got else np.zeros((0, 3))
        # reference: all periodic images of all particles inside my ghost shell (one cell wide)
        L = np.array(ghi) - np.array(glo)
        ref = []
        for sx in (-1, 0, 1):
            for sy in (-1, 0, 1):
                for sz in (-1, 0, 1):
                    sh = np.array([sx, sy, sz]) * L
                    if any(s_ != 0 and not periodic[d] for d, s_ in enumerate((sx, sy, sz))):
                        continue
                    q = pos + sh
                    inside_outer = np.all((q >= np.array(lo) - np.array(cs)) & (q < np.array(hi) + np.array(cs)), axis=1)
                    inside_brick = np.all((q >= np.array(lo)) & (q < np.array(hi)), axis=1)
                    ref.append(q[inside_outer & ~inside_brick])
        ref = np.concatenate(ref)
        key = lambda a: np.array(sorted(map(tuple, np.round(a, 9))))  # noqa: E731
        assert len(got) == len(ref), (len(got), len(ref))
        assert np.allclose(key(got), key(ref), atol=1e-9)
        # (iv) weak-scaling bricks: globally unique ids, bricks tile the global box
        bw = parallel.brick_workload(5, rank, world)
        ids = [None] * world
        dist.all_gather_object(ids, bw.particles.id)
        allid = np.concatenate(ids)
        assert len(np.unique(allid)) == len(allid)
        assert all(bw.lo[d] <= bw.particles.vec("r")[:, d].min() and bw.particles.vec("r")[:, d].max() < bw.hi[d] for d in range(3))
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic,balanced", [(2, (1, 1, 1), False), (2, (1, 0, 1), False), (4, (1, 1, 0), False),
                                                     (2, (1, 1, 1), True), (4, (1, 1, 0), True)])
def test_brick_plan_gloo(world, periodic, balanced):
    """balanced: the bricks are cut at the cost-balanced positions (parallel.balanced_cuts), so neighbouring bricks hold different
    cell grids; the ghost plan must still reproduce the single-domain periodic images."""
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic, ret, balanced)) for r in range(world)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(timeout=300)
    assert all(p_.exitcode == 0 for p_ in procs), [p_.exitcode for p_ in procs]
    assert len(ret) == world


def test_process_grid():
    assert parallel.process_grid(8) == (2, 2, 2)
    assert parallel.process_grid(4) == (2, 2, 1)
    assert parallel.process_grid(2) == (2, 1, 1)
    assert parallel.process_grid(1) == (1, 1, 1)
    for n in (1, 2, 4, 8):
        pg = parallel.process_grid(n)
        assert sorted(parallel.rank_coords(r, pg) for r in range(n)) == sorted(
            (x, y, z) for z in range(pg[2]) for y in range(pg[1]) for x in range(pg[0]))


def _drum_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = parallel.drum_slab_workload(3000, rank, world)
        assert wl.pgrid == (1, world, 1) and wl.periodic == (0, 1, 0)
        L = wl.hi[1] - wl.lo[1]
        # slabs tile the long drum along its axis; ids are globally unique; every slab holds the N=1 particle count
        r = wl.particles.vec("r")
        assert wl.lo[1] <= r[:, 1].min() and r[:, 1].max() < wl.hi[1]
        info = [None] * world
        dist.all_gather_object(info, {"lo": wl.lo, "hi": wl.hi, "ghi": wl.ghi, "n": wl.particles.n, "id": wl.particles.id})
        assert all(abs(info[k]["lo"][1] - k * L) < 1e-12 and abs(info[k]["hi"][1] - (k + 1) * L) < 1e-12 for k in range(world))
        assert all(i["ghi"] == wl.ghi and abs(wl.ghi[1] - world * L) < 1e-12 for i in info)
        assert len({i["n"] for i in info}) == 1
        allid = np.concatenate([i["id"] for i in info])
        assert len(np.unique(allid)) == len(allid)
        # ghost plan of a slab: only the two axis neighbours (one and the same peer when world == 2), periodic wrap at the ends
        nc, _ = parallel.grid_dims(wl.lo, wl.hi, wl.cell_size)
        rp = parallel.recv_plan(wl.glo, wl.ghi, wl.periodic, wl.pgrid, rank, nc)
        sp = parallel.send_plan(wl.glo, wl.ghi, wl.periodic, wl.pgrid, rank, nc)
        assert set(rp) == {(rank - 1) % world, (rank + 1) % world} == set(sp)
        plans = [None] * world
        dist.all_gather_object(plans, {"send": sp, "recv": rp})
        for other in rp:
            assert plans[other]["send"].get(rank, []) == rp[other]
        shifts = {sh[1] for ent in rp.values() for (_, _, sh) in ent}
        assert shifts <= {0.0, wl.ghi[1], -wl.ghi[1]} and (rank not in (0, world - 1) or len(shifts) == 2 or world == 1)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_drum_slab_decomposition_gloo(world):
    """The N>1 bench workload (BASELINE configs[1] cut into N slabs along the drum axis): tiling, ids, ghost/send plans."""
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_drum_worker, args=(r, world, port, ret)) for r in range(world)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(timeout=300)
    assert all(p_.exitcode == 0 for p_ in procs), [p_.exitcode for p_ in procs]
    assert len(ret) == world
