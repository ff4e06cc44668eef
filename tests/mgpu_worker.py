"""Multi-GPU parity worker — launched with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_worker.py [--mode sphere|poly] [--steps K] [--rebuild-every E]

Every rank owns one brick of a global periodic box (exadem_b200.parallel), runs K Verlet steps with NCCL ghost exchange
and rebuild-time migration inside libexadem_b200, then rank 0 gathers everything and checks it against the CPU oracle
run on the SAME particles as ONE domain: particle state within 1e-10 (matched by id), the canonical interaction set
bit-exact with its friction/moment history within 1e-10, same number of rebuilds.  Exit code 0 = parity."""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def global_workload(mode, world):
    from exadem_b200 import parallel, workloads
    pg = parallel.process_grid(world)
    if mode == "drum":
        # the bench's N>1 workload in small: a rotating drum (cylinder driver, periodic along its axis) cut into slabs along the axis
        pg = (1, world, 1)
        wl = workloads.rotating_drum(12000 * world, jitter=0.02, length=10.0 * world, seed=0xD3B4 + world)
        wl.particles.vy[:] *= 40.0                      # particles cross the slab faces (and the periodic end) within the run
        wl.particles.ry[:] = np.mod(wl.particles.ry + 0.4995, wl.hi[1])
    elif mode == "sphere":
        side = [10 * pg[d] for d in range(3)]
        wl = workloads.spheres_box(tuple(side), jitter=0.03, omega_sigma=0.3, sigma_v=0.1, seed=0xD3B0 + world)
        # lattice planes 5e-4 below the brick faces: particles cross brick (and periodic) boundaries within ~50 steps
        wl.particles.rx[:] = np.mod(wl.particles.rx + 0.4995, side[0])
        wl.particles.ry[:] = np.mod(wl.particles.ry + 0.4995, side[1])
    else:
        n_side = [10 * pg[d] for d in range(3)]
        wl = workloads.polyhedra_box(tuple(n_side), shapes=("octahedron", "cube"), pitch=0.8, seed=0xD3B8 + world, cell_size=2.0)
        L = n_side[0] * 0.8
        wl.particles.rx[:] = np.mod(wl.particles.rx + 0.3995, L)
        wl.particles.vx[:] *= 30.0
    return wl, pg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="sphere")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--rebuild-every", type=int, default=10)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--then-ops", type=int, default=0, help="after run_verlet, this many more steps operator by operator")
    ap.add_argument("--dump", default=None, help="write the ContactState view of wave 0 right after run_verlet to <dump>.rank<r>.npz")
    ap.add_argument("--balanced", action="store_true", help="thin the particles out on one side of the box and cut the bricks where "
                    "dem_cost_model's cost field balances (edb_set_brick_cuts) instead of uniformly")
    ap.add_argument("--moving", action="store_true", help="add a stress-controlled floor (surface, LINEAR_COMPRESSIVE_MOTION): reaction all-reduce")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from exadem_b200 import Engine, Library, capi, parallel
    from parity import compare_interaction_sets, field_close

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl, pg = global_workload(args.mode, world)
    if args.moving:
        # every rank holds the same drivers; the floor is touched by particles of every brick, so its reaction is a sum over ranks
        # (force_to_accel_driver.cpp:128-139) and its motion must come out identical on all of them
        area = (wl.hi[0] - wl.lo[0]) * (wl.hi[1] - wl.lo[1])
        wl.drivers = list(wl.drivers) + [dict(kind=capi.DRIVER_SURFACE, offset=0.03, normal=(0, 0, 1), center=(0, 0, 0.03))]
        wl.motions = [dict(motion_type="STATIONARY")] * (len(wl.drivers) - 1) + [
            dict(motion_type="LINEAR_COMPRESSIVE_MOTION", motion_vector=(0, 0, -1), sigma=2000.0, damprate=0.0, surface=area)]
    cuts = cell_cuts = None
    if args.balanced:
        # a heap on the low-x side: 35 % of the particles of the upper half of the box are kept.  The cost field is dem_cost_model's
        # before any list exists (3 N + 1 per global cell); the cuts minimise the largest brick cost per axis, whole global cells
        r = wl.particles.vec("r")
        rng = np.random.default_rng(0xD3B9)
        keep = (r[:, 0] < 0.5 * (wl.lo[0] + wl.hi[0])) | (rng.random(wl.particles.n) < 0.35)
        wl.particles = parallel.select(wl.particles, keep)
        cuts, gcells, cell_cuts = parallel.balanced_decomposition(wl.particles.vec("r"), wl.lo, wl.hi, wl.effective_cell_size(), pg, min_cells=2)
        assert any(len(set(np.diff(c))) > 1 for c in cell_cuts), f"the balanced cuts came out uniform: {cell_cuts}"
    system_mass = float(wl.particles.mass.sum())
    pos = wl.particles.vec("r")
    mine = parallel.owner_rank(pos, wl.lo, wl.hi, pg, cuts) == rank
    lo, hi = parallel.brick_bounds(wl.lo, wl.hi, pg, rank, cuts)
    cfg = capi.make_config(lo, hi, wl.effective_cell_size(), wl.periodic, wl.mode, local, glo=wl.lo, ghi=wl.hi, rank=rank,
                           nranks=world, pgrid=pg)
    lib = Library()
    eng = Engine(lib, cfg)
    if cuts is not None:
        eng.set_brick_cuts(cuts)
    parallel.init_comm(eng, lib, rank)
    eng.set_law(wl.law)
    eng.set_contact_params(wl.params, wl.params_driver)
    if wl.shapes_spec:
        eng.set_shapes(wl.shapes())
    eng.set_drivers(wl.drivers)
    if wl.motions:
        eng.set_driver_motions(wl.motions)
        eng.set_system_mass(system_mass)
    eng.upload_particles(parallel.select(wl.particles, mine))
    sym = wl.symetric
    eng.update_particle_neighbors(wl.rcut_inc, sym)
    nreb = eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, sym, args.steps, args.rebuild_every)

    def ops_step(e):
        # one step of numerical_scheme_verlet operator by operator, without a rebuild (config_numerical_schemes.msp:6-30)
        e.combined_compute_prolog(wl.dt)
        e.ghost_update_rq()
        e.compute_vertices()
        e.reset_force_moment()
        e.gravity_force(wl.gravity)
        e.contact(wl.dt)
        e.force_to_accel()
        e.combined_compute_epilog(wl.dt)

    if args.dump:
        # ContactState view right after run_verlet: the contact point of a record that is not in contact is recomputed from the CURRENT
        # positions of i and j — for a j on a neighbour brick, from its ghost slot
        st = eng.get_contact_state(0)
        np.savez(f"{args.dump}.rank{rank}.npz", dn=st[0], cp=st[1], fn=st[2], ft=st[3])
    for _ in range(args.then_ops):
        ops_step(eng)
    pa = eng.download_particles()
    rec = eng.export_interactions()
    ctr = eng.counters()
    drv_state = eng.get_drivers(len(wl.drivers)) if wl.motions else []
    payload = {"pa": {f: getattr(pa, f) for f in capi.F64_FIELDS + ["id", "type", "group"]}, "rec": rec, "nreb": nreb,
               "n": pa.n, "launches": ctr.kernel_launches, "drv": drv_state}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    ok = True
    if rank == 0:
        from oracle_binding import OracleEngine, OracleLibrary
        ref = wl.setup(OracleEngine(OracleLibrary(), wl.config()))
        ref.set_threads(os.cpu_count() or 1)
        ref.set_virtual_bricks(pg)   # same both-sides enumeration of brick-straddling pairs as the N ranks (quirk Q6)
        if cell_cuts is not None:
            ref.set_virtual_cuts(cell_cuts)
        if wl.motions:
            ref.set_system_mass(system_mass)
        ref.update_particle_neighbors(wl.rcut_inc, sym)
        nreb_ref = ref.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, sym, args.steps, args.rebuild_every)
        for _ in range(args.then_ops):
            ops_step(ref)
        pr = ref.download_particles()
        rr = ref.export_interactions()
        try:
            if wl.motions:
                dref = ref.get_drivers(len(wl.drivers))
                moved_drv = False
                for g_ in gathered:
                    for a_, b_ in zip(g_["drv"], dref):
                        for k_ in ("center", "offset", "normal", "vel", "acc", "forces", "radius"):
                            field_close(np.atleast_1d(a_[k_]), np.atleast_1d(b_[k_]), 1e-9, f"driver {k_}")
                moved_drv = abs(dref[-1]["offset"] - 0.03) > 0 and np.abs(dref[-1]["forces"]).sum() > 0
                assert moved_drv, "the stress-controlled floor neither moved nor felt a reaction"
            counts = [g["n"] for g in gathered]
            assert sum(counts) == pr.n, f"particle count {counts} != {pr.n}"
            assert all(g["nreb"] == nreb_ref for g in gathered), f"rebuild counts {[g['nreb'] for g in gathered]} vs {nreb_ref}"
            ids = np.concatenate([g["pa"]["id"] for g in gathered])
            order = np.argsort(ids)
            oref = np.argsort(pr.id)
            assert np.array_equal(ids[order], pr.id[oref]), "particle id sets differ"
            errs = {}
            for f in capi.F64_FIELDS:
                a = np.concatenate([g["pa"][f] for g in gathered])[order]
                b = getattr(pr, f)[oref]
                if f in ("rx", "ry", "rz"):   # positions: compare modulo the period (a particle may sit on either side of a wrap)
                    d = "xyz".index(f[1])
                    L = wl.hi[d] - wl.lo[d]
                    a = a - L * np.round((a - b) / L)
                try:
                    errs[f] = field_close(a, b, args.rtol, f, elementwise=False)   # bricks vs one domain sum in different orders over 60 steps: field-scaled bar only (tests/parity.py)
                except AssertionError as e:
                    if not args.verbose:
                        raise
                    bad = np.nonzero(np.abs(a - b) > args.rtol * max(np.max(np.abs(b)), 1e-300))[0]
                    print(f"  {e}; {len(bad)} particles off; first: " + "; ".join(
                        f"id={pr.id[oref][k]} r=({pr.rx[oref][k]:.4f},{pr.ry[oref][k]:.4f},{pr.rz[oref][k]:.4f}) got={a[k]:.6g} ref={b[k]:.6g}" for k in bad[:6]), flush=True)
                    errs[f] = 1.0
            all_rec = np.concatenate([g["rec"] for g in gathered])
            sphere = wl.mode == capi.MODE_SPHERE
            for r_ in (all_rec, rr):
                r_["cell_i"] = r_["cell_j"] = 0
                r_["p_i"] = r_["p_j"] = 0
            n = compare_interaction_sets(all_rec, rr, payload_rtol=args.rtol, elem_floor=1.0)
            cb = rr
            active = int((np.abs(cb["friction"]).sum(axis=1) > 0).sum())
            moved = sum(abs(c - counts[0]) for c in counts)
            if cell_cuts is not None:
                uni = np.bincount(parallel.owner_rank(pos, wl.lo, wl.hi, pg), minlength=world)
                print(f"balanced cuts (cells) {cell_cuts}: particles per rank {counts} (uniform bricks would hold {uni.tolist()})", flush=True)
                assert max(counts) < max(uni), "the balanced decomposition is not better than the uniform one"
            print(f"mgpu parity OK: {world} GPUs pgrid={pg} mode={args.mode} particles={pr.n} per-rank={counts} rebuilds={nreb_ref} "
                  f"interactions={n} (active {active}) max field err={max(errs.values()):.2e}", flush=True)
        except AssertionError as e:
            print(f"mgpu parity FAILED: {e}", flush=True)
            ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
