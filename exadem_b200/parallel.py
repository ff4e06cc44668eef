"""Host-side plumbing of the multi-GPU path: uniform brick decomposition of the global box over the GPUs of one node
and the NCCL bootstrap (one process per GPU; torch.distributed only carries the 128-byte NCCL unique id and the
scalar reductions of bench.py — the per-step ghost exchange itself runs inside libexadem_b200 over NCCL/NVLink).

Replaces, for this path, the reference's MPI spatial decomposition (exaNBody `load_balance_rcb` + `UpdateGhostsNode`,
SURVEY.md 2d/2e): uniform bricks instead of RCB (the benchmark boxes are uniform), same 1-cell ghost layer.
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


def brick_bounds(glo, ghi, pgrid, rank):
    """Bounds of the brick of `rank`: the global box cut into pgrid equal parts per axis."""
    c = rank_coords(rank, pgrid)
    lo, hi = [], []
    for d in range(3):
        ln = (ghi[d] - glo[d]) / pgrid[d]
        lo.append(glo[d] + c[d] * ln)
        hi.append(glo[d] + (c[d] + 1) * ln if c[d] + 1 < pgrid[d] else ghi[d])
    return tuple(lo), tuple(hi)


def owner_rank(pos: np.ndarray, glo, ghi, pgrid) -> np.ndarray:
    """Rank owning each position (N,3) — same rule as the device-side migration classifier."""
    r = np.zeros(len(pos), dtype=np.int64)
    mult = 1
    for d in range(3):
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


def brick_workload(side: int, rank: int, nranks: int, seed=0xD3A5):
    """Weak-scaling workload: every rank owns a side^3 brick (side^3 spheres on the jittered SC lattice) of a fully
    periodic global box of pgrid*side cells.  Each rank generates ONLY its own brick (seeded per brick)."""
    from . import workloads
    pg = process_grid(nranks)
    c = rank_coords(rank, pg)
    wl = workloads.spheres_box((side, side, side), seed=seed + 7919 * rank, name=f"periodic box, {pg[0]}x{pg[1]}x{pg[2]} bricks of {side}^3")
    off = [c[d] * side for d in range(3)]
    wl.particles.rx += off[0]
    wl.particles.ry += off[1]
    wl.particles.rz += off[2]
    n = wl.particles.n
    wl.particles.id[:] = (np.uint64(rank) * np.uint64(n) + wl.particles.id).astype(np.uint64)
    wl.lo = tuple(float(o) for o in off)
    wl.hi = tuple(float(o + side) for o in off)
    wl.glo = (0.0, 0.0, 0.0)
    wl.ghi = tuple(float(pg[d] * side) for d in range(3))
    wl.pgrid = pg
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
    return capi.make_config(wl.lo, wl.hi, cs, wl.periodic, wl.mode, device, glo=glo, ghi=ghi, rank=rank, nranks=nranks, pgrid=pg)


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


def ghost_plan(glo, ghi, periodic, pgrid, rank, nc):
    """List of (ghost cell ijk, source rank, source cell ijk, shift) for the brick of `rank` (cells are 1-based inside
    the ghost-shelled local grid, like the device code)."""
    me = rank_coords(rank, pgrid)
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
                        src[d] = nc[d]
                        if rb[d] < 0:
                            if not periodic[d]:
                                ok = False
                                break
                            rb[d] = pgrid[d] - 1
                            shift[d] = -(ghi[d] - glo[d])
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


def send_plan(glo, ghi, periodic, pgrid, rank, nc):
    """What `rank` must SEND: for every other rank, the entries of THAT rank's ghost plan whose source is `rank`, in that
    rank's enumeration order (every brick has the same cell grid, so no handshake is needed) — mirrors
    Comm::plan_send_lists in csrc/comm.cuh."""
    nranks = pgrid[0] * pgrid[1] * pgrid[2]
    out = {}
    for other in range(nranks):
        if other == rank:
            continue
        ent = [(g, s, sh) for (g, src_rank, s, sh) in ghost_plan(glo, ghi, periodic, pgrid, other, nc) if src_rank == rank]
        if ent:
            out[other] = ent
    return out


def recv_plan(glo, ghi, periodic, pgrid, rank, nc):
    """What `rank` RECEIVES, per source rank, in its own enumeration order (local periodic images excluded)."""
    out = {}
    for (g, src_rank, s, sh) in ghost_plan(glo, ghi, periodic, pgrid, rank, nc):
        if src_rank != rank:
            out.setdefault(src_rank, []).append((g, s, sh))
    return out
