"""CPU tests of the oracle as a simulator (no GPU): physical invariants, history semantics, and agreement between the
canonical serial accumulation order and the OpenMP/atomic order the reference uses (classifier_for_all.hpp:243-255)."""
import numpy as np
import pytest

from exadem_b200 import capi, workloads
from oracle_binding import OracleEngine
from parity import compare_interaction_sets, compare_particles, sort_interactions


def mk(wl, lib, threads=1):
    e = wl.setup(OracleEngine(lib, wl.config()))
    e.set_threads(threads)
    return e


def test_serial_vs_openmp_order(oracle_lib):
    for wl in (workloads.rotating_drum(3000, jitter=0.02), workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.78)):
        a, b = mk(wl, oracle_lib, 1), mk(wl, oracle_lib, 4)
        for e in (a, b):
            e.update_particle_neighbors(wl.rcut_inc, True)
            e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 10, 4)
        compare_interaction_sets(a.export_interactions(), b.export_interactions(), payload_rtol=1e-10)
        compare_particles(a.download_particles(), b.download_particles(), rtol=1e-10)


def test_momentum_conservation_periodic_box(oracle_lib):
    wl = workloads.spheres_box((6, 6, 6), jitter=0.03, omega_sigma=0.2)
    e = mk(wl, oracle_lib)
    p0 = (wl.particles.mass[:, None] * wl.particles.vec("v")).sum(axis=0)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, (0, 0, 0), wl.rcut_inc, True, 40, 0)
    pa = e.download_particles()
    p1 = (pa.mass[:, None] * pa.vec("v")).sum(axis=0)
    scale = (pa.mass[:, None] * np.abs(pa.vec("v"))).sum()
    assert np.max(np.abs(p1 - p0)) < 1e-12 * scale  # pair forces are exactly antisymmetric (mirror records, quirk Q6)


def test_periodic_pairs_are_mirrored(oracle_lib):
    """Quirk Q6: a pair that straddles the periodic boundary is enumerated from both sides (ghost partner), once per side."""
    wl = workloads.spheres_box((6, 6, 6), jitter=0.01)
    e = mk(wl, oracle_lib)
    e.update_particle_neighbors(wl.rcut_inc, True)
    rec = e.export_interactions()
    keys = set(zip(rec["id_i"].tolist(), rec["id_j"].tolist()))
    mirrored = [k for k in keys if (k[1], k[0]) in keys]
    assert len(mirrored) > 0 and len(mirrored) % 2 == 0
    assert len(keys) == len(rec)  # no duplicate key
    # interior pairs obey id_i < id_j
    assert all(k[0] < k[1] for k in keys if (k[1], k[0]) not in keys)


def test_history_reset_and_carry_over(oracle_lib):
    wl = workloads.spheres_box((6, 6, 6), jitter=0.03)
    e = mk(wl, oracle_lib)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 5, -1)
    before = sort_interactions(e.export_interactions())
    dn, cp, fn, ft = e.get_contact_state(0)
    # quirk Q5: records without contact carry zero friction/moment and dn = 0
    inactive = dn == 0
    assert inactive.any() and (~inactive).any()
    e.update_particle_neighbors(wl.rcut_inc, True)
    after = sort_interactions(e.export_interactions())

    def as_dict(rec):
        return {tuple(int(r[f]) for f in ("type", "id_i", "id_j", "sub_i", "sub_j", "swap")): (r["friction"].tobytes(), r["moment"].tobytes())
                for r in rec}
    db, da = as_dict(before), as_dict(after)
    common = [k for k in da if k in db]
    assert len(common) > 500
    assert all(da[k] == db[k] for k in common)            # payload carried over bit for bit
    zero = np.zeros(3).tobytes()
    assert all(da[k] == (zero, zero) for k in da if k not in db)  # new candidates start without history
    assert sum(1 for k in common if da[k][0] != zero) > 50
    # restart path: export -> import into a fresh engine -> rebuild restores the same history
    e2 = mk(wl, oracle_lib)
    e2.upload_particles(e.download_particles())
    e2.import_interactions(after)
    e2.update_particle_neighbors(wl.rcut_inc, True)
    again = sort_interactions(e2.export_interactions())
    assert np.array_equal(again["friction"], after["friction"]) and np.array_equal(again["id_j"], after["id_j"])


def test_rolling_moment_quirk_q1(oracle_lib):
    """contact_force.hpp:223-234: with vrot_i == 0 the NaN threshold keeps mom = kr*(w_j - w_i)*dt; otherwise mur = 0 zeroes it."""
    wl = workloads.spheres_box((2, 1, 1), periodic=(0, 0, 0), cell_size=2.0, jitter=0.0, sigma_v=0.0)
    wl.lo, wl.hi = (-3.0, -3.0, -3.0), (5.0, 3.0, 3.0)
    pa = wl.particles
    pa.rx[:] = [0.0, 0.95]
    pa.ry[:] = 0.0
    pa.rz[:] = 0.0
    pa.id[:] = [0, 1]
    pa.vrotz[:] = [0.0, 2.0]
    e = mk(wl, oracle_lib)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.reset_force_moment()
    e.contact_sphere(wl.dt)
    rec = e.export_interactions()
    assert len(rec) == 1 and rec["moment"][0][2] == pytest.approx(wl.params["kr"] * 2.0 * wl.dt)
    pa.vrotz[:] = [1.0, 2.0]
    e = mk(wl, oracle_lib)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.reset_force_moment()
    e.contact_sphere(wl.dt)
    assert np.all(e.export_interactions()["moment"][0] == 0.0)


def test_polyhedra_candidates_cover_all_contacts(oracle_lib):
    """Every record in contact after a few steps was a candidate (the Verlet skin holds): structural invariant of
    check_consistency_grid_classifier.cpp:59-116 — counts per type stay equal between ges and ic."""
    wl = workloads.polyhedra_box((5, 5, 5), pitch=0.78)
    e = mk(wl, oracle_lib)
    e.update_particle_neighbors(wl.rcut_inc, True)
    c0, _ = e.interaction_counts()
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 8, -1)
    c1, a1 = e.interaction_counts()
    assert np.array_equal(c0, c1) and a1[:4].sum() > 0
    assert c0[1] > c0[0]  # octahedra: more vertex-edge than vertex-vertex candidates


def test_update_interaction_ghost_twins(oracle_lib):
    """update_interaction_ghost.cpp:58-285: every OwnerGhost record of a polyhedron pair that straddles the periodic boundary
    has one PartnerGhost twin (same ids, subs, swap, payload) on the side where the partner is real, so that the pair force
    is applied to both particles: total force and linear momentum are conserved in a periodic box."""
    wl = workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.78)
    wl.periodic = (1, 1, 1)
    wl.drivers = []
    e = mk(wl, oracle_lib)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, (0, 0, 0), wl.rcut_inc, True, 6, 3)
    rec = e.export_interactions()
    og, pg = rec[rec["ghost"] == 1], rec[rec["ghost"] == 2]
    assert len(og) > 50 and len(og) == len(pg)
    key = lambda r: sorted(zip(r["type"].tolist(), r["id_i"].tolist(), r["id_j"].tolist(), r["sub_i"].tolist(), r["sub_j"].tolist(), r["swap"].tolist()))
    assert key(og) == key(pg)
    e.ghost_update_rq()   # the epilog changed v and vrot since the last refresh
    e.compute_vertices()
    e.reset_force_moment()
    e.contact_polyhedron(wl.dt)
    pa = e.download_particles()
    f = pa.vec("f")
    assert np.abs(f).sum() > 0
    assert np.max(np.abs(f.sum(axis=0))) < 1e-11 * np.abs(f).sum()
    # the twin is not a history source, and a rebuild re-ships the payload with the record
    e.update_particle_neighbors(wl.rcut_inc, True)
    rec2 = e.export_interactions()
    og2, pg2 = sort_interactions(rec2[rec2["ghost"] == 1]), sort_interactions(rec2[rec2["ghost"] == 2])
    assert np.array_equal(og2["friction"], pg2["friction"]) and np.abs(og2["friction"]).sum() > 0


def test_asymmetric_spheres_get_twins_without_force(oracle_lib):
    """nbh_sphere.cu:266-294 (symetric:false): boundary pairs are OwnerGhost and get a PartnerGhost twin record, but the
    non-symmetric kernel applies the force to particle i only (contact_sphere.hpp:190-196), i.e. to the ghost copy: dropped."""
    wl = workloads.spheres_box((6, 6, 6), jitter=0.03)
    e = mk(wl, oracle_lib)
    e.update_particle_neighbors(wl.rcut_inc, False)
    rec = e.export_interactions()
    assert (rec["ghost"] == 1).sum() > 0 and (rec["ghost"] == 1).sum() == (rec["ghost"] == 2).sum()
