"""Host-side plumbing of the multi-GPU path: brick decomposition of the global box over the GPUs of one node (uniform, or cut
at cost-balanced positions: balanced_cuts / Engine.set_brick_cuts) and the NCCL bootstrap (one process per GPU; torch.distributed only carries the 128-byte NCCL unique id and the
scalar reductions of bench.py — the per-step ghost exchange itself runs inside libexadem_b200 over NCCL/NVLink).

Replaces, for this path, the reference's MPI spatial decomposition (exaNBody `load_balance_rcb` + `UpdateGhostsNode`,
SURVEY.md 2d/2e): rectilinear bricks instead of RCB — uniform for the benchmark boxes, cost-balanced per axis from the
dem_cost_model field otherwise —, same 1-cell ghost layer.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .capi import ParticleArrays


def process_grid(nranks: int) -> tuple[int, int, int]:
    """Most cubic factorisation px*py*pz = nranks (8 -> 2x2x2, 4 -> 2x2x1, 2 -> 2x1x1)."""
    best = (nranks, 1, 1)
    score = None
    for px in range(1, nranks + 1):
        if nranks % px:
            continue
        for py in range(1, nranks // px + 1):
            if (nranks // px) % py:
                continue
            pz = nranks // (px * py)
            dims = sorted((px, py, pz), reverse=True)
            s = (dims[0] - dims[2], dims[0])
            if score is None or s < score:
                score, best = s, tuple(dims)
    return best  # px >= py >= pz


def rank_coords(rank: int, pgrid) -> tuple[int, int, int]:
    return rank % pgrid[0], (rank // pgrid[0]) % pgrid[1], rank // (pgrid[0] * pgrid[1])


def balanced_cuts(cost_zyx: np.ndarray, pgrid, min_cells: int = 1):
    """Rectilinear load balance from the dem_cost_model field (Engine.dem_cost_model(), shape (nz, ny, nx) over the GLOBAL real-cell
    grid): per axis, the cell indices at which the box is cut into pgrid[d] slabs so that the largest slab cost (marginal sum of the
    axis) is as small as possible, every slab keeping at least `min_cells` cells.  Returns [cuts_x, cuts_y, cuts_z], each a list of
    pgrid[d] + 1 increasing cell indices from 0 to n_d.

    Stands in for exaNBody's load_balance_rcb on this path ("RCB by dem_cost_model weights"): tensor-product cuts keep the
    26-neighbour brick topology the NVLink ghost path is built on, where RCB would give every brick an arbitrary set of neighbours.
    The minimisation is exact (dynamic programme over the prefix sums, n_d <= a few hundred cells, <= 8 slabs)."""
    cost = np.asarray(cost_zyx, dtype=np.float64)
    out = []
    for d in range(3):
        axes = tuple(a for a in range(3) if a != 2 - d)          # array axis of x is 2, y 1, z 0
        m = cost.sum(axis=axes)
        n, parts = len(m), int(pgrid[d])
        if parts * min_cells > n:
            raise ValueError(f"balanced_cuts: axis {d} has {n} cells, cannot hold {parts} slabs of >= {min_cells} cells")
        pre = np.concatenate([[0.0], np.cumsum(m)])
        inf = float("inf")
        best = np.full((parts + 1, n + 1), inf)
        arg = np.zeros((parts + 1, n + 1), dtype=np.int64)
        best[0, 0] = 0.0
        for k in range(1, parts + 1):
            for e in range(k * min_cells, n - (parts - k) * min_cells + 1):
                b = np.arange((k - 1) * min_cells, e - min_cells + 1)
                cand = np.maximum(best[k - 1, b], pre[e] - pre[b])
                j = int(np.argmin(cand))                          # first minimum: ties go to the earliest cut (deterministic)
                best[k, e], arg[k, e] = cand[j], b[j]
        cuts, e = [n], n
        for k in range(parts, 0, -1):
            e = int(arg[k, e])
            cuts.append(e)
        out.append(cuts[::-1])
    return out


def cuts_to_positions(cuts, glo, ghi, global_cells):
    """Cell-index cuts -> coordinates (the global cell grid is uniform: cell d has size (ghi - glo) / global_cells[d])."""
    return [[glo[d] + (ghi[d] - glo[d]) * c / global_cells[d] for c in cuts[d]] for d in range(3)]


def brick_bounds(glo, ghi, pgrid, rank, cuts=None):
    """Bounds of the brick of `rank`: the global box cut into pgrid equal parts per axis, or at the coordinates `cuts`
    (cuts[d] = pgrid[d] + 1 increasing positions from glo[d] to ghi[d], see balanced_cuts / cuts_to_positions)."""
    c = rank_coords(rank, pgrid)
    if cuts is not None:
        return tuple(float(cuts[d][c[d]]) for d in range(3)), tuple(float(cuts[d][c[d] + 1]) for d in range(3))
    lo, hi = [], []
    for d in range(3):
        ln = (ghi[d] - glo[d]) / pgrid[d]
        lo.append(glo[d] + c[d] * ln)
        hi.append(glo[d] + (c[d] + 1) * ln if c[d] + 1 < pgrid[d] else ghi[d])
    return tuple(lo), tuple(hi)


def owner_rank(pos: np.ndarray, glo, ghi, pgrid, cuts=None) -> np.ndarray:
    """Rank owning each position (N,3) — same rule as the device-side migration classifier (brick_of, kernels_migrate.cuh)."""
    r = np.zeros(len(pos), dtype=np.int64)
    mult = 1
    for d in range(3):
        if cuts is not None:
            c = np.clip(np.searchsorted(np.asarray(cuts[d], dtype=np.float64), pos[:, d], side="right") - 1, 0, pgrid[d] - 1)
        else:
            ln = (ghi[d] - glo[d]) / pgrid[d]
            c = np.clip(np.floor((pos[:, d] - glo[d]) / ln).astype(np.int64), 0, pgrid[d] - 1)
        r += mult * c
        mult *= pgrid[d]
    return r


def select(pa: ParticleArrays, mask: np.ndarray) -> ParticleArrays:
    out = ParticleArrays(int(mask.sum()))
    for f in capi.F64_FIELDS + ["id", "type", "group"]:
        setattr(out, f, np.ascontiguousarray(getattr(pa, f)[mask]))
    return out


def brick_workload(side, rank: int, nranks: int, seed=0xD3A5, jitter=2e-5):
    """Periodic-box workload cut into uniform bricks (S5 topology): every rank owns a brick of `side` = (sx, sy, sz) cells of pitch 1.0
    (an int means a cube) holding sx*sy*sz spheres on the jittered SC lattice, inside a fully periodic global box of pgrid*side.
    Each rank generates ONLY its own brick (seeded per brick).  Weak scaling keeps `side` fixed; strong scaling divides a fixed global
    box by the process grid (`strong_brick_dims`)."""
    from . import workloads
    dims = (side, side, side) if np.isscalar(side) else tuple(int(x) for x in side)
    pg = process_grid(nranks)
    c = rank_coords(rank, pg)
    wl = workloads.spheres_box(dims, seed=seed + 7919 * rank, jitter=jitter,
                               name=f"periodic box, {pg[0]}x{pg[1]}x{pg[2]} bricks of {dims[0]}x{dims[1]}x{dims[2]}")
    off = [c[d] * dims[d] for d in range(3)]
    wl.particles.rx += off[0]
    wl.particles.ry += off[1]
    wl.particles.rz += off[2]
    n = wl.particles.n
    wl.particles.id[:] = (np.uint64(rank) * np.uint64(n) + wl.particles.id).astype(np.uint64)
    wl.lo = tuple(float(o) for o in off)
    wl.hi = tuple(float(off[d] + dims[d]) for d in range(3))
    wl.glo = (0.0, 0.0, 0.0)
    wl.ghi = tuple(float(pg[d] * dims[d]) for d in range(3))
    wl.pgrid = pg
    return wl


def strong_brick_dims(global_side, nranks: int):
    """Brick of a FIXED global cube of `global_side`^3 cells cut by process_grid(nranks) (strong scaling).  The global side must be
    divisible by every pgrid factor into bricks holding whole 2.5-wide cells."""
    pg = process_grid(nranks)
    dims = []
    for d in range(3):
        if global_side % pg[d]:
            raise ValueError(f"global side {global_side} not divisible by process grid {pg}")
        dims.append(global_side // pg[d])
    return tuple(dims)


def global_workload(make_rank_workload, nranks: int):
    """The whole N-rank workload as ONE domain (what a single process — the CPU reference arm, a parity test — runs): the per-rank
    generators are seeded per brick, so concatenating them reproduces exactly the particles the N GPU ranks hold together."""
    parts = [make_rank_workload(r, nranks) for r in range(nranks)]
    wl = parts[0]
    if nranks == 1:
        return wl
    pa = ParticleArrays(sum(p.particles.n for p in parts))
    for f in capi.F64_FIELDS + ["id", "type", "group"]:
        setattr(pa, f, np.ascontiguousarray(np.concatenate([getattr(p.particles, f) for p in parts])))
    wl.particles = pa
    wl.lo, wl.hi = tuple(wl.glo), tuple(wl.ghi)
    return wl


def drum_slab_workload(n_per_gpu: int, rank: int, nranks: int, seed=0xD3A2):
    """Weak scaling of BASELINE configs[1]: the rotating drum made `nranks` times longer along its (periodic) axis and cut into
    `nranks` slabs, one per GPU — per-GPU work is exactly the N=1 drum (same particle count, same cylinder, same number of
    ghost cells: the two periodic end faces become faces shared with the neighbouring slabs).  pgrid = (1, nranks, 1)."""
    from . import workloads
    wl = workloads.rotating_drum(n_per_gpu, seed=seed + 7919 * rank, name=f"drum, {nranks} slabs along the axis")
    L = wl.hi[1] - wl.lo[1]
    n = wl.particles.n
    wl.particles.ry += rank * L
    wl.particles.id[:] = (np.uint64(rank) * np.uint64(n) + wl.particles.id).astype(np.uint64)
    wl.glo = wl.lo
    wl.ghi = (wl.hi[0], float(nranks * L), wl.hi[2])
    wl.lo = (wl.lo[0], float(rank * L), wl.lo[2])
    wl.hi = (wl.hi[0], float((rank + 1) * L), wl.hi[2])
    wl.pgrid = (1, nranks, 1)
    return wl


def brick_config(wl, rank: int, nranks: int, device: int = 0):
    pg = getattr(wl, "pgrid", process_grid(nranks))
    glo = getattr(wl, "glo", wl.lo)
    ghi = getattr(wl, "ghi", wl.hi)
    cs = wl.cell_size
    for d in range(3):  # >= 3 cells per periodic axis of a single-brick axis; bricks must hold an integer number of cells
        ln = wl.hi[d] - wl.lo[d]
        if wl.periodic[d] and pg[d] == 1:
            cs = min(cs, ln / 3.0)
    rmax = float(wl.particles.radius.max()) if wl.particles.n else 0.0
    if cs < 2.0 * rmax + wl.rcut_inc:
        raise ValueError(f"brick_config: a periodic axis shorter than 3 cells of 2 r_max + rcut_inc = {2.0 * rmax + wl.rcut_inc:.4g} "
                         f"cannot be run with a one-cell stencil (cell size would be {cs:.4g})")
    return capi.make_config(wl.lo, wl.hi, cs, wl.periodic, wl.mode, device, glo=glo, ghi=ghi, rank=rank, nranks=nranks, pgrid=pg)


def balanced_decomposition(positions, glo, ghi, cell_size, pgrid, cost_zyx=None, min_cells=1):
    """Cut positions of a cost-balanced decomposition.  `cost_zyx` is the dem_cost_model field over the global real-cell grid
    (nz, ny, nx); when None the cost of a cell is 3 N + 1 from the particle positions — what dem_cost_model returns before any
    interaction list exists (dem_cost_model.cpp:74-77).  Returns (cuts as coordinates per axis, global cell counts)."""
    gcells, _ = grid_dims(glo, ghi, cell_size)
    if cost_zyx is None:
        pos = np.asarray(positions)
        idx = [np.clip(np.floor((pos[:, d] - glo[d]) / ((ghi[d] - glo[d]) / gcells[d])).astype(np.int64), 0, gcells[d] - 1) for d in range(3)]
        cnt = np.zeros((gcells[2], gcells[1], gcells[0]))
        np.add.at(cnt, (idx[2], idx[1], idx[0]), 1.0)
        cost_zyx = 3.0 * cnt + 1.0
    cuts = balanced_cuts(cost_zyx, pgrid, min_cells=min_cells)
    return cuts_to_positions(cuts, glo, ghi, gcells), gcells, cuts


def init_comm(engine, lib, rank: int):
    """Broadcast rank 0's NCCL unique id through torch.distributed and create the library's communicator."""
    import torch
    import torch.distributed as dist
    if rank == 0:
        uid = capi.comm_unique_id(lib)
        t = torch.tensor(list(uid), dtype=torch.uint8)
    else:
        t = torch.zeros(128, dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    engine.comm_init(bytes(t.cpu().tolist()))


# ---------------------------------------------------------------------------------------------------------------------
# Reference model of the ghost plan on the host (numpy).  Used by the world_size-2 gloo tests to check that the brick
# decomposition + ghost exchange reproduces the single-domain periodic ghost layer; mirrors csrc/comm.cuh.
# ---------------------------------------------------------------------------------------------------------------------
def grid_dims(lo, hi, cell_size):
    nc = [max(1, int(np.floor((hi[d] - lo[d]) / cell_size + 1e-9))) for d in range(3)]
    cs = [(hi[d] - lo[d]) / nc[d] for d in range(3)]
    return nc, cs


def cells_per_brick(cuts, cell_size, pgrid):
    """nc_axis[d][coord] = real cells along axis d of the bricks at that coordinate (same rule as the library's setup_grid)."""
    return [[max(1, int(np.floor((cuts[d][k + 1] - cuts[d][k]) / cell_size + 1e-9))) for k in range(pgrid[d])] for d in range(3)]


def ghost_plan(glo, ghi, periodic, pgrid, rank, nc, nc_axis=None):
    """List of (ghost cell ijk, source rank, source cell ijk, shift) for the brick of `rank` (cells are 1-based inside
    the ghost-shelled local grid, like the device code).  `nc_axis` (cells_per_brick) describes bricks cut at non-uniform
    positions: the grid of `rank` and of the source bricks then come from it and `nc` is ignored."""
    me = rank_coords(rank, pgrid)
    if nc_axis is not None:
        nc = [nc_axis[d][me[d]] for d in range(3)]
    g = [n + 2 for n in nc]
    plan = []
    for k in range(g[2]):
        for j in range(g[1]):
            for i in range(g[0]):
                idx = (i, j, k)
                if not any(idx[d] == 0 or idx[d] == g[d] - 1 for d in range(3)):
                    continue
                rb, src, shift, ok = list(me), list(idx), [0.0, 0.0, 0.0], True
                for d in range(3):
                    if idx[d] == 0:
                        rb[d] -= 1
                        if rb[d] < 0:
                            if not periodic[d]:
                                ok = False
                                break
                            rb[d] = pgrid[d] - 1
                            shift[d] = -(ghi[d] - glo[d])
                        src[d] = nc[d] if nc_axis is None else nc_axis[d][rb[d]]   # last real cell of the SOURCE brick
                    elif idx[d] == g[d] - 1:
                        rb[d] += 1
                        src[d] = 1
                        if rb[d] >= pgrid[d]:
                            if not periodic[d]:
                                ok = False
                                break
                            rb[d] = 0
                            shift[d] = ghi[d] - glo[d]
                if ok:
                    plan.append((idx, rb[0] + pgrid[0] * (rb[1] + pgrid[1] * rb[2]), tuple(src), tuple(shift)))
    return plan


def send_plan(glo, ghi, periodic, pgrid, rank, nc, nc_axis=None):
    """What `rank` must SEND: for every other rank, the entries of THAT rank's ghost plan whose source is `rank`, in that
    rank's enumeration order (every rank can compute every brick's cell grid, so no handshake is needed) — mirrors
    Comm::plan_send_lists in csrc/comm.cuh."""
    nranks = pgrid[0] * pgrid[1] * pgrid[2]
    out = {}
    for other in range(nranks):
        if other == rank:
            continue
        ent = [(g, s, sh) for (g, src_rank, s, sh) in ghost_plan(glo, ghi, periodic, pgrid, other, nc, nc_axis) if src_rank == rank]
        if ent:
            out[other] = ent
    return out


def recv_plan(glo, ghi, periodic, pgrid, rank, nc, nc_axis=None):
    """What `rank` RECEIVES, per source rank, in its own enumeration order (local periodic images excluded)."""
    out = {}
    for (g, src_rank, s, sh) in ghost_plan(glo, ghi, periodic, pgrid, rank, nc, nc_axis):
        if src_rank != rank:
            out.setdefault(src_rank, []).append((g, s, sh))
    return out
