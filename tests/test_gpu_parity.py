"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars: interaction/candidate set bit-exact as a sorted set; forces, torques, positions, velocities, quaternions within
1e-10 (relative to the field scale) in fp64 — north_star.  Sizes are chosen so that the oracle finishes in seconds."""
import numpy as np
import pytest

from exadem_b200 import Engine, capi, workloads
from oracle_binding import OracleEngine
from parity import bit_identical, compare_interaction_sets, compare_particles, field_close

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def pair(wl, product_lib, oracle_lib):
    gpu = wl.setup(Engine(product_lib, wl.config()))
    cpu = wl.setup(OracleEngine(oracle_lib, wl.config()))
    return gpu, cpu


def both(engines, fn):
    for e in engines:
        fn(e)


def check(gpu, cpu, rtol=RTOL, payload=True):
    n = compare_interaction_sets(gpu.export_interactions(), cpu.export_interactions(), payload_rtol=rtol if payload else None)
    errs = compare_particles(gpu.download_particles(), cpu.download_particles(), rtol=rtol)
    return n, errs


def test_sphere_operators_one_by_one(product_lib, oracle_lib):
    """The reference's operator sequence of one Verlet step (config_numerical_schemes.msp:6-30), operator by operator."""
    wl = workloads.spheres_box((9, 8, 10), jitter=0.03, omega_sigma=0.5, polydisperse=0.2)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    es = (gpu, cpu)
    both(es, lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    cg, _ = gpu.interaction_counts()
    cc, _ = cpu.interaction_counts()
    assert np.array_equal(cg, cc) and cg[0] > 1000
    for step in range(3):
        for op in (lambda e: e.combined_compute_prolog(wl.dt), lambda e: e.ghost_update_rq(), lambda e: e.reset_force_moment(),
                   lambda e: e.gravity_force(wl.gravity), lambda e: e.contact_sphere(wl.dt), lambda e: e.force_to_accel(),
                   lambda e: e.combined_compute_epilog(wl.dt)):
            both(es, op)
            check(gpu, cpu)
    _, active = gpu.interaction_counts()
    assert active[0] > 20  # friction history is being exercised


def test_unfused_push_operators(product_lib, oracle_lib):
    wl = workloads.spheres_box((6, 6, 6), jitter=0.03, omega_sigma=1.0)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    es = (gpu, cpu)
    both(es, lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    for op in (lambda e: e.reset_force_moment(), lambda e: e.gravity_force((0.3, -0.2, -9.807)), lambda e: e.contact_sphere(wl.dt),
               lambda e: e.push_to_angular_acceleration(), lambda e: e.push_to_angular_velocity(wl.dt),
               lambda e: e.push_to_quaternion(wl.dt), lambda e: e.force_to_accel(), lambda e: e.freeze_particles()):
        both(es, op)
        check(gpu, cpu)


@pytest.mark.parametrize("rebuild_every", [-1, 4, 0])
def test_drum_fused_run(product_lib, oracle_lib, rebuild_every):
    """S2-like drum (cylinder driver, periodic y) through the fused numerical_scheme_verlet path."""
    wl = workloads.rotating_drum(6000, jitter=0.02)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 24, rebuild_every) for e in (gpu, cpu)]
    assert nb[0] == nb[1]
    n, errs = check(gpu, cpu)
    cg, ag = gpu.interaction_counts()
    assert cg[4] > 0 and ag[0] > 0
    # the fused path equals the operator-by-operator path bit for bit
    gpu2 = wl.setup(Engine(product_lib, wl.config()))
    gpu2.update_particle_neighbors(wl.rcut_inc, True)
    for it in range(24):
        gpu2.combined_compute_prolog(wl.dt)
        if rebuild_every > 0 and (it + 1) % rebuild_every == 0 or rebuild_every == 0 and gpu2.displ_over(0.5 * wl.rcut_inc) > 0:
            gpu2.update_particle_neighbors(wl.rcut_inc, True)
        else:
            gpu2.ghost_update_rq()
        gpu2.reset_force_moment(); gpu2.gravity_force(wl.gravity); gpu2.contact_sphere(wl.dt); gpu2.force_to_accel()
        gpu2.combined_compute_epilog(wl.dt)
    assert all(bit_identical(gpu.download_particles(), gpu2.download_particles()).values())


@pytest.mark.parametrize("switch", ["EDB_NBH_WARP", "EDB_NBH_TWO_PASS", "EDB_NBH_NO_CULL"])
@pytest.mark.parametrize("kind", ["drum", "box"])
def test_neighbour_build_variants_are_bit_identical(product_lib, monkeypatch, switch, kind):
    """The sphere neighbour build has a default form (thread per particle, candidates recorded in the count pass and copied in the
    fill pass, cell-level reach cull) and three switches read at context creation: a warp per particle for the count pass, the
    two-pass walk of round 1, and the walk without the cell cull.  All four must produce the same records in the same order — the
    exported interaction arrays and the state after rebuilding steps are compared bit for bit.  The drum has a cylinder driver
    (the driver waves run beside the particle-particle wave), the box none."""
    wl = workloads.rotating_drum(6000, jitter=0.02) if kind == "drum" else workloads.spheres_box((9, 8, 7), jitter=0.03, sigma_v=200.0)
    base = wl.setup(Engine(product_lib, wl.config()))
    monkeypatch.setenv(switch, "1")
    var = wl.setup(Engine(product_lib, wl.config()))
    monkeypatch.delenv(switch)
    exports = []
    for e in (base, var):
        e.update_particle_neighbors(wl.rcut_inc, True)
        exports.append(e.export_interactions())
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 12, 3)
    assert len(exports[0]) == len(exports[1]) > 0
    for f in exports[0].dtype.names:                                 # same records in the same order
        assert np.array_equal(exports[0][f], exports[1][f]), (switch, f)
    assert all(bit_identical(base.download_particles(), var.download_particles()).values())
    assert all(np.array_equal(a, b) for a, b in zip(base.interaction_counts(), var.interaction_counts()))


def test_displacement_trigger_rebuilds_inside_run(product_lib, oracle_lib):
    """The speculative step loop of edb_run_verlet (kernels gated on the device-side trigger, host one step behind): fast
    spheres force a neighbour rebuild every ~10 steps; the rebuild count, the interaction set and the state must match the
    oracle, which takes the decision synchronously (vertex/particle_displ_over -> update_particles_full)."""
    wl = workloads.spheres_box((8, 8, 8), jitter=0.02, sigma_v=300.0, omega_sigma=0.3)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 45, 0) for e in (gpu, cpu)]
    assert nb[0] == nb[1] and nb[0] >= 3, nb
    check(gpu, cpu)
    # again in two calls (the loop state does not leak between calls)
    nb2 = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 7, 0) + e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 1, 0) for e in (gpu, cpu)]
    assert nb2[0] == nb2[1]
    check(gpu, cpu)


def test_history_survives_rebuild(product_lib, oracle_lib):
    wl = workloads.spheres_box((8, 8, 8), jitter=0.03, omega_sigma=0.3)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    both((gpu, cpu), lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 6, -1))
    before = gpu.export_interactions()
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    after = gpu.export_interactions()
    compare_interaction_sets(after, cpu.export_interactions(), payload_rtol=RTOL)
    nz_b = int((np.abs(before["friction"]).sum(axis=1) > 0).sum())
    nz_a = int((np.abs(after["friction"]).sum(axis=1) > 0).sum())
    assert nz_b > 50 and nz_a == nz_b  # no particle moved out of range in 6 steps: every active record is carried over


@pytest.mark.parametrize("law", [capi.LAW_COHESIVE, capi.LAW_DMT])
def test_cohesive_and_dmt_laws(product_lib, oracle_lib, law):
    wl = workloads.deposit((8, 8, 6))
    wl.law = law
    wl.params = dict(wl.params, fc=10.0, dncut=0.1, gamma=10.0)      # rotating-drum-periodic-cohesion.msp:30-34, add_xyz.msp:59
    wl.params_driver = dict(wl.params)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    both((gpu, cpu), lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 12, 5))
    check(gpu, cpu)
    dn_g, cp_g, fn_g, ft_g = gpu.get_contact_state(0)
    dn_c, cp_c, fn_c, ft_c = cpu.get_contact_state(0)
    assert len(dn_g) == len(dn_c) > 0


def test_surface_ball_drivers_and_contact_state(product_lib, oracle_lib):
    wl = workloads.deposit((9, 9, 5))
    wl.drivers.append(dict(kind=capi.DRIVER_BALL, radius=5.0, center=(4.5, 4.5, 3.0), vel=(0.1, 0, 0), vrot=(0, 0, 0.2)))
    wl.drivers.append(dict(kind=capi.DRIVER_CYLINDER, radius=6.2, center=(4.5, 4.5, 4.5), axis=(1, 1, 0), vrot=(0, 0, 0.5)))
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    both((gpu, cpu), lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 10, 0))
    check(gpu, cpu)
    cg, _ = gpu.interaction_counts()
    assert cg[5] > 0 and cg[6] > 0 and cg[4] > 0
    # ContactState buffers (dn, cp, fn, ft) per wave, in wave order on both sides (owner-sorted enumeration order)
    for t in (0, 4, 5, 6):
        sg, sc = gpu.get_contact_state(t), cpu.get_contact_state(t)
        for a, b, name in zip(sg, sc, ("dn", "cp", "fn", "ft")):
            field_close(a, b, RTOL, f"contact state {name} wave {t}")
    fg, mg = gpu.compute_driver_force_moment(3)
    fc, mc = cpu.compute_driver_force_moment(3)
    field_close(fg, fc, 1e-9, "driver forces")
    field_close(mg, mc, 1e-9, "driver moments")


def test_multimat_and_asymmetric(product_lib, oracle_lib):
    wl = workloads.spheres_box((7, 7, 7), jitter=0.03, omega_sigma=0.2, periodic=(1, 0, 1))
    wl.particles.group[:] = (np.arange(wl.particles.n) % 3).astype(np.uint32)
    wl.drivers = [dict(kind=capi.DRIVER_SURFACE, offset=-0.1, normal=(0, 1, 0))]
    base = wl.params
    pp = [dict(base, kn=base["kn"] * (1 + 0.1 * (a + b)), mu=0.1 * (1 + a * b)) for a in range(3) for b in range(3)]
    pd = [dict(base, kt=5e4 * (1 + a)) for a in range(3)]
    es = []
    for mk in (lambda: Engine(product_lib, wl.config()), lambda: OracleEngine(oracle_lib, wl.config())):
        e = mk()
        e.set_law(wl.law)
        e.set_drivers(wl.drivers)
        e.set_contact_params_multimat(3, pp, 1, pd)
        e.upload_particles(wl.particles)
        es.append(e)
    gpu, cpu = es
    for sym in (True, False):
        both(es, lambda e: e.update_particle_neighbors(wl.rcut_inc, sym))
        both(es, lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, sym, 5, -1))
        check(gpu, cpu)


def test_edge_cases(product_lib, oracle_lib):
    # a single particle, no interactions; then two far-apart particles
    for n_side in ((1, 1, 1), (2, 1, 1)):
        wl = workloads.spheres_box(n_side, periodic=(0, 0, 0), cell_size=2.0)
        wl.lo, wl.hi = (-5.0, -5.0, -5.0), (5.0, 5.0, 5.0)
        wl.particles.rx[:] *= 3.0
        gpu, cpu = pair(wl, product_lib, oracle_lib)
        both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
        both((gpu, cpu), lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 3, 0))
        n, _ = check(gpu, cpu)
        assert n == 0


@pytest.mark.parametrize("shapes", [("octahedron",), ("octahedron", "cube")])
def test_polyhedra(product_lib, oracle_lib, shapes):
    wl = workloads.polyhedra_box((6, 6, 6), shapes=shapes, pitch=0.78)
    wl.drivers = [dict(kind=capi.DRIVER_SURFACE, offset=0.15, normal=(0, 0, 1)),
                  dict(kind=capi.DRIVER_BALL, radius=3.0, center=(2.3, 2.3, 2.3))]
    wl.periodic = (1, 1, 0)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    cg, _ = gpu.interaction_counts()
    cc, _ = cpu.interaction_counts()
    assert np.array_equal(cg, cc)
    assert all(cg[t] > 0 for t in (0, 1, 2, 3, 5, 6)), cg
    check(gpu, cpu)
    # operator sequence once, then the fused path with rebuilds
    for op in (lambda e: e.combined_compute_prolog(wl.dt), lambda e: e.ghost_update_rq(), lambda e: e.compute_vertices(),
               lambda e: e.reset_force_moment(), lambda e: e.gravity_force(wl.gravity), lambda e: e.contact_polyhedron(wl.dt),
               lambda e: e.force_to_accel(), lambda e: e.combined_compute_epilog(wl.dt)):
        both((gpu, cpu), op)
        check(gpu, cpu)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 12, 4) for e in (gpu, cpu)]
    assert nb[0] == nb[1] == 3
    check(gpu, cpu)
    _, ag = gpu.interaction_counts()
    assert ag[:4].sum() > 20
    for t in (0, 1, 2, 3):
        sg, sc = gpu.get_contact_state(t), cpu.get_contact_state(t)
        for a, b, name in zip(sg, sc, ("dn", "cp", "fn", "ft")):
            field_close(a, b, RTOL, f"contact state {name} wave {t}")
    assert gpu.displ_over(1e-7) == cpu.displ_over(1e-7)


def test_bounded_drift_1000_steps(product_lib, oracle_lib):
    """north_star: bounded drift over 1,000 steps (trajectories are chaotic, so the bar is on conserved-ish aggregates
    and on the divergence staying small, not on bit equality)."""
    wl = workloads.spheres_box((6, 6, 6), jitter=2e-5, sigma_v=0.05)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    g0 = (0.0, 0.0, 0.0)
    both((gpu, cpu), lambda e: e.run_verlet(wl.dt, g0, wl.rcut_inc, True, 1000, 0))
    pg, pc = gpu.download_particles(), cpu.download_particles()
    assert np.isfinite(pg.vec("r")).all()
    drift = np.max(np.abs(pg.vec("r") - pc.vec("r")))
    assert drift < 1e-6, drift
    ek_g = 0.5 * np.sum(pg.mass[:, None] * pg.vec("v") ** 2)
    ek_c = 0.5 * np.sum(pc.mass[:, None] * pc.vec("v") ** 2)
    assert abs(ek_g - ek_c) <= 1e-6 * max(ek_c, 1e-30)


def _boxed(wl, lo, hi, mink, **drv):
    """Encloses a workload in a closed box RShape driver (mesh walls) and makes the domain non-periodic."""
    spec, centre = workloads.box_mesh(lo, hi, mink)
    wl.mesh_spec = [spec]
    wl.drivers = list(wl.drivers) + [dict(kind=capi.DRIVER_RSHAPE, shape=0, center=centre, **drv)]
    wl.periodic = (0, 0, 0)
    return wl


def test_rshape_mesh_spheres(product_lib, oracle_lib):
    """Waves 7,8,9: spheres against an STL-like box mesh (vertex / edge / face of the mesh), config 4 of BASELINE.json."""
    wl = workloads.spheres_box((8, 8, 6), jitter=0.03, omega_sigma=0.3, polydisperse=0.1, cell_size=2.0)
    wl.lo, wl.hi = (-1.0, -1.0, -1.0), (9.0, 9.0, 7.0)
    c45 = np.sqrt(0.5)
    _boxed(wl, (0.03, 0.04, 0.05), (7.96, 7.97, 5.98), 0.02, vel=(0.05, 0.0, 0.0), vrot=(0.0, 0.0, 0.1))
    # a second, rotated mesh: exercises the orientation path (rshape.hpp:381-384)
    spec2, _ = workloads.box_mesh((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), 0.05)
    wl.mesh_spec.append(spec2)
    wl.drivers.append(dict(kind=capi.DRIVER_RSHAPE, shape=1, center=(4.0, 4.0, 3.0), quat=(c45, 0.0, c45 * 0.6, c45 * 0.8)))
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    cg, _ = gpu.interaction_counts()
    cc, _ = cpu.interaction_counts()
    assert np.array_equal(cg, cc) and all(cg[t] > 0 for t in (7, 8, 9)), cg
    check(gpu, cpu)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 12, 5) for e in (gpu, cpu)]
    assert nb[0] == nb[1]
    check(gpu, cpu)
    _, ag = gpu.interaction_counts()
    assert ag[9] > 0
    for t in (7, 8, 9):
        sg, sc = gpu.get_contact_state(t), cpu.get_contact_state(t)
        for a, b, name in zip(sg, sc, ("dn", "cp", "fn", "ft")):
            field_close(a, b, RTOL, f"contact state {name} wave {t}")
    fg, mg = gpu.compute_driver_force_moment(2)
    fc, mc = cpu.compute_driver_force_moment(2)
    field_close(fg, fc, 1e-9, "mesh driver forces")
    field_close(mg, mc, 1e-9, "mesh driver moments")


def test_rshape_mesh_polyhedra(product_lib, oracle_lib):
    """Waves 7..12: polyhedra against a box mesh, including the 'j to i' waves 11/12 whose stored friction flips sign."""
    wl = workloads.polyhedra_box((5, 5, 5), shapes=("octahedron", "cube"), pitch=0.8, cell_size=2.0)
    wl.lo, wl.hi = (-2.0, -2.0, -2.0), (6.0, 6.0, 6.0)
    _boxed(wl, (0.1, 0.12, 0.11), (3.9, 3.88, 3.91), 0.03, vrot=(0.0, 0.2, 0.0))
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    cg, _ = gpu.interaction_counts()
    cc, _ = cpu.interaction_counts()
    assert np.array_equal(cg, cc) and all(cg[t] > 0 for t in (7, 8, 9, 10, 11, 12)), cg
    check(gpu, cpu)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 12, 5) for e in (gpu, cpu)]
    assert nb[0] == nb[1]
    check(gpu, cpu)
    _, ag = gpu.interaction_counts()
    assert ag[7:13].sum() > 0
    for t in range(7, 13):
        sg, sc = gpu.get_contact_state(t), cpu.get_contact_state(t)
        for a, b, name in zip(sg, sc, ("dn", "cp", "fn", "ft")):
            field_close(a, b, RTOL, f"contact state {name} wave {t}")
    fg, mg = gpu.compute_driver_force_moment(1)
    fc, mc = cpu.compute_driver_force_moment(1)
    field_close(fg, fc, 1e-9, "mesh driver forces")


def test_mixed_spheres_polyhedra_cohesion_mesh(product_lib, oracle_lib):
    """BASELINE configs[3] in small: spheres (one-vertex shapes), octahedra and cubes in one polyhedron context, cohesive law,
    closed STL-like mesh walls (waves 0-3 and 7-12 together)."""
    wl = workloads.polyhedra_box((6, 6, 5), shapes=("octahedron", "cube", "sphere"), pitch=0.8, cell_size=2.0, seed=0xD3C4)
    wl.lo, wl.hi = (-2.0, -2.0, -2.0), (7.0, 7.0, 6.0)
    wl.law = capi.LAW_COHESIVE
    wl.params = dict(wl.params, fc=0.05, dncut=0.02)
    wl.params_driver = dict(wl.params_driver, fc=0.02, dncut=0.01)
    _boxed(wl, (0.1, 0.12, 0.11), (4.7, 4.68, 3.91), 0.03, vrot=(0.0, 0.2, 0.0))
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    cg, _ = gpu.interaction_counts()
    cc, _ = cpu.interaction_counts()
    assert np.array_equal(cg, cc) and all(cg[t] > 0 for t in (0, 1, 2, 3, 7, 8, 9)), cg
    check(gpu, cpu)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 15, 5) for e in (gpu, cpu)]
    assert nb[0] == nb[1]
    check(gpu, cpu)
    _, ag = gpu.interaction_counts()
    assert ag[:4].sum() > 0


@pytest.mark.parametrize("kind", ["drum", "poly"])
def test_stress_tensor(product_lib, oracle_lib, kind):
    """stress_tensor: per-particle sum of (fn + ft) (x) (cp - r) over the contacts of the step, on the device from the
    resident contact state (no host copy of the waves), against the oracle's restatement of stress_tensor.cu:41-81."""
    if kind == "drum":
        wl = workloads.rotating_drum(6000, jitter=0.02)
    else:
        wl = workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.78, cell_size=2.0)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    both((gpu, cpu), lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 8, -1))
    sg, sc = gpu.stress_tensor(), cpu.stress_tensor()
    assert sg.shape == sc.shape and np.abs(sc).max() > 0
    field_close(sg.ravel(), sc.ravel(), RTOL, "stress tensor")
    # total stress is symmetric in exchange of i and j up to the contact-point lever arms: trace of the sum equals sum f.(cp-r)
    assert np.isfinite(sg).all()


@pytest.mark.parametrize("kind", ["drum", "poly"])
def test_checkpoint_restart(product_lib, oracle_lib, kind):
    """the interaction history leaves and re-enters the device in the reference's 104-byte `Interaction`
    record layout (the payload of write_dump_particle_interaction): run -> export + download -> FRESH context -> upload + import ->
    rebuild -> continue must equal the uninterrupted run bit for bit, and the oracle taken through the same restart."""
    if kind == "drum":
        wl = workloads.rotating_drum(5000, jitter=0.02)
    else:
        wl = workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.78, cell_size=2.0)

    def run(lib, cls, restart):
        e = wl.setup(cls(lib, wl.config()))
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 10, -1)
        e.update_particle_neighbors(wl.rcut_inc, True)     # a dump is written right after a rebuild (config_iteration_dump.msp:68-86)
        if restart:
            pa, rec = e.download_particles(), e.export_interactions()
            assert (np.abs(rec["friction"]).sum(axis=1) > 0).any()
            e.close()
            e = wl.setup(cls(lib, wl.config()))
            e.upload_particles(pa)
            e.import_interactions(rec)
            e.update_particle_neighbors(wl.rcut_inc, True)
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 10, -1)
        return e

    g0, g1 = run(product_lib, Engine, False), run(product_lib, Engine, True)
    c1 = run(oracle_lib, OracleEngine, True)
    p0, p1 = g0.download_particles(), g1.download_particles()
    o0, o1 = np.argsort(p0.id), np.argsort(p1.id)
    for f in capi.F64_FIELDS:
        assert np.array_equal(getattr(p0, f)[o0], getattr(p1, f)[o1]), f
    compare_interaction_sets(g0.export_interactions(), g1.export_interactions(), payload_rtol=0.0)
    check(g1, c1)


@pytest.mark.parametrize("kind", ["drum", "poly"])
def test_simulation_state(product_lib, oracle_lib, kind):
    """simulation_state: the log line's reductions on the device — energies, mass, interaction totals (PartnerGhost
    skipped, particle-particle waves x2 when symetric), active count and min dn — against the oracle's restatement."""
    if kind == "drum":
        wl = workloads.rotating_drum(6000, jitter=0.02)
    else:
        wl = workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.78, cell_size=2.0)
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    both((gpu, cpu), lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 8, -1))
    a, b = gpu.simulation_state(True), cpu.simulation_state(True)
    assert a.n_particles == b.n_particles == wl.particles.n
    assert a.n_total_interactions == b.n_total_interactions > 0
    assert a.n_active_interactions == b.n_active_interactions > 0
    assert b.min_dn < 0 and abs(a.min_dn - b.min_dn) <= 1e-10 * abs(b.min_dn)
    for x, y in zip(list(a.kinetic_energy) + list(a.rotation_energy) + [a.mass], list(b.kinetic_energy) + list(b.rotation_energy) + [b.mass]):
        assert abs(x - y) <= 1e-12 * max(abs(y), 1e-300)


def test_polyhedra_vertex_trigger_inside_run(product_lib, oracle_lib):
    """vertex_displ_over evaluated inside the fused gather (k_gather<2,true>) and the speculative step loop, for polyhedra: fast,
    spinning particles force rebuilds; rebuild count, interaction set and state must match the oracle's synchronous decision."""
    wl = workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.8, cell_size=2.0, seed=0xD3C7)
    wl.particles.vx[:] *= 1500.0
    wl.particles.vy[:] *= 1500.0
    wl.particles.vrotz[:] += 60.0
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    both((gpu, cpu), lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 40, 0) for e in (gpu, cpu)]
    assert nb[0] == nb[1] and nb[0] >= 2, nb
    check(gpu, cpu)


@pytest.mark.parametrize("law", [capi.LAW_NONE, capi.LAW_DMT])
def test_polyhedra_homothety_multimat_cylinder(product_lib, oracle_lib, law):
    """Polyhedra whose size varies (homothety 0.8..1.2 scales vertices, Minkowski radius and OBB), three material groups with a
    full multi-material parameter table, a cylinder driver around the packing (wave 4), Hooke and Hooke+DMT."""
    wl = workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.95, cell_size=2.5, seed=0xD3C9)
    rng = np.random.default_rng(5)
    n = wl.particles.n
    wl.particles.homothety[:] = rng.uniform(0.8, 1.2, n)
    wl.particles.radius[:] *= wl.particles.homothety      # bounding radius follows the homothety (radius_from_shape)
    wl.particles.group[:] = (np.arange(n) % 3).astype(np.uint32)
    L = 6 * 0.95
    wl.periodic = (0, 1, 0)
    wl.lo, wl.hi = (-2.5, 0.0, -2.5), (L + 2.5, L, L + 2.5)
    c = 0.5 * L - 0.475
    wl.drivers = [dict(kind=capi.DRIVER_CYLINDER, radius=0.5 * L + 0.2, center=(c, c, c), axis=(1, 0, 1), vrot=(0, 0.5, 0))]
    wl.law = law
    base = dict(wl.params, gamma=0.02 if law == capi.LAW_DMT else 0.0)
    pp = [dict(base, kn=base["kn"] * (1 + 0.1 * (a + b)), mu=0.1 * (1 + a * b)) for a in range(3) for b in range(3)]
    pd = [dict(base, kt=5e3 * (1 + a)) for a in range(3)]
    es = []
    for mk in (lambda: Engine(product_lib, wl.config()), lambda: OracleEngine(oracle_lib, wl.config())):
        e = mk()
        e.set_law(wl.law)
        e.set_shapes(wl.shapes())
        e.set_drivers(wl.drivers)
        e.set_contact_params_multimat(3, pp, 1, pd)
        e.upload_particles(wl.particles)
        es.append(e)
    gpu, cpu = es
    both(es, lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    cg, _ = gpu.interaction_counts()
    cc, _ = cpu.interaction_counts()
    assert np.array_equal(cg, cc) and cg[:4].sum() > 0 and cg[4] > 0, cg
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 12, 5) for e in es]
    assert nb[0] == nb[1]
    check(gpu, cpu)
    _, ag = gpu.interaction_counts()
    assert ag[:4].sum() > 0


def test_full_size_drum_properties(product_lib):
    """BASELINE configs[1] at full size (1 M spheres, ~3 M interactions), checked through size-independent properties: (i) two
    runs in fresh contexts are BIT-identical (fixed-order force sums; the dynamic chunk queue, the side stream and the speculative
    step loop must not leak scheduling into the results), (ii) Newton's third law: the particle-particle forces cancel, so
    sum(m a) = M g - (reaction on the cylinder), (iii) every record in contact was a candidate with history keyed as exported."""
    wl = workloads.rotating_drum(1_000_000)

    def run():
        e = wl.setup(Engine(product_lib, wl.config()))
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 20, 0)
        pa = e.download_particles()
        fr = e.export_interactions()
        fd, _ = e.compute_driver_force_moment(1)
        st = e.simulation_state(True)
        e.close()
        return pa, fr, fd, st

    pa, fr, fd, st = run()
    pb, fr2, _, _ = run()
    for f in capi.F64_FIELDS:
        assert np.array_equal(getattr(pa, f), getattr(pb, f)), f"run-to-run difference in {f}"
    assert np.array_equal(fr["friction"], fr2["friction"]) and np.array_equal(fr["id_j"], fr2["id_j"])
    assert st.n_particles == pa.n and st.n_active_interactions > 1_000_000
    # (ii) fx,fy,fz hold the acceleration after force_to_accel
    ma = pa.vec("f") * pa.mass[:, None]
    total = ma.sum(axis=0)
    expect = pa.mass.sum() * np.array(wl.gravity) - fd[0]
    scale = np.abs(ma).sum()
    assert np.abs(total - expect).max() <= 1e-9 * scale, (total, expect, scale)
    # (iii) sortedness / uniqueness of the exported keys
    key = np.stack([fr["type"].astype(np.uint64), fr["id_i"], fr["id_j"]], axis=1)
    assert len(np.unique(key, axis=0)) == len(key)


def test_full_size_polyhedra_properties(product_lib):
    """BASELINE configs[2] at full size (0.49 M octahedra, 3 M records in four waves): run-to-run bit identity (the work list of
    the detection pass is filled in a scheduling-dependent order; the results must not depend on it) and sum(m a) = M g."""
    wl = workloads.polyhedra_box((79, 79, 79), pitch=0.78)

    def run():
        e = wl.setup(Engine(product_lib, wl.config()))
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 12, 5)
        pa = e.download_particles()
        c, a = e.interaction_counts()
        e.close()
        return pa, c, a

    pa, c1, a1 = run()
    pb, c2, a2 = run()
    assert np.array_equal(c1, c2) and np.array_equal(a1, a2) and a1[:4].sum() > 10_000
    for f in capi.F64_FIELDS:
        assert np.array_equal(getattr(pa, f), getattr(pb, f)), f"run-to-run difference in {f}"
    ma = pa.vec("f") * pa.mass[:, None]
    expect = pa.mass.sum() * np.array(wl.gravity)
    assert np.abs(ma.sum(axis=0) - expect).max() <= 1e-9 * np.abs(ma).sum()


def _full_size_vs_oracle(wl, product_lib, oracle_lib, more_steps):
    """CUDA path against the oracle AT THE BENCH SIZE: neighbour build + 1 step, then `more_steps` further steps, the oracle in its
    canonical serial accumulation order throughout (the CUDA gather sums in that order).  Its OpenMP order would not do here: one
    ulp of a coordinate of 70 is 1.4e-14, times kn = 1e5 that is 1.4e-9 of a friction threshold mu kn |dn| of order 1 — positions
    that differ by an ulp after 25 steps put the friction history outside 1e-10 (measured: 1.2e-9).
    Sets bit-exact, every field within 1e-10 of the field scale and per element."""
    gpu, cpu = pair(wl, product_lib, oracle_lib)
    es = (gpu, cpu)
    both(es, lambda e: e.update_particle_neighbors(wl.rcut_inc, True))
    cg, _ = gpu.interaction_counts()
    cc, _ = cpu.interaction_counts()
    assert np.array_equal(cg, cc), (cg, cc)
    both(es, lambda e: e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 1, 0))
    n1, _ = check(gpu, cpu)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, more_steps, 0) for e in es]
    assert nb[0] == nb[1]
    n2, errs = check(gpu, cpu)
    _, ag = gpu.interaction_counts()
    _, ac = cpu.interaction_counts()
    assert np.array_equal(ag, ac)
    return n1, n2, ag, errs


def test_full_size_drum_vs_oracle(product_lib, oracle_lib):
    """BASELINE configs[1] as bench.py runs it (1 M spheres in the rotating drum, ~3 M records): 1 step and 25 steps against the oracle"""
    wl = workloads.rotating_drum(1_000_000)
    n1, n2, active, errs = _full_size_vs_oracle(wl, product_lib, oracle_lib, 24)
    assert n1 > 2_900_000 and active[0] > 1_000_000, (n1, active)


def test_full_size_octahedra_vs_oracle(product_lib, oracle_lib):
    """BASELINE configs[2] as bench.py runs it (0.49 M octahedra, ~3 M records in the four particle-particle waves): 1 + 24 steps"""
    wl = workloads.polyhedra_box((79, 79, 79), pitch=0.78)
    n1, n2, active, errs = _full_size_vs_oracle(wl, product_lib, oracle_lib, 24)
    assert n1 > 2_500_000 and active[:4].sum() > 10_000, (n1, active)


def test_async_host_copies_equal_blocking(product_lib):
    """edb_upload_dynamic_async / edb_download_dynamic_async (own copy streams, per-field events) against the blocking pair: the
    host-owned-grid loop of bench.py's e2e leg, 6 steps, bit-identical host arrays; and a masked download leaves the other fields alone."""
    import torch
    wl = workloads.rotating_drum(60000, jitter=0.02)

    def pinned(pa):
        out = pa.copy()
        keep = []
        for f in capi.F64_FIELDS:
            t = torch.empty(pa.n, dtype=torch.float64, pin_memory=True)
            a = t.numpy()
            a[:] = getattr(pa, f)
            setattr(out, f, a)
            keep.append(t)
        out._keep = keep
        return out

    hosts = []
    for use_async in (False, True):
        e = wl.setup(Engine(product_lib, wl.config()))
        e.update_particle_neighbors(wl.rcut_inc, True)
        host = pinned(e.download_particles())
        for _ in range(6):
            if use_async:
                e.upload_dynamic_async(host)
                e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 1, 0)
                e.download_dynamic_async(host)
            else:
                e.upload_dynamic(host)
                e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 1, 0)
                e.download_dynamic(host)
        e.sync()
        hosts.append(host)
        if use_async:
            before = host.fx.copy()
            host.rx[:] = -1.0
            e.download_dynamic_async(host, capi.DYN_STATE)
            e.sync()
            assert np.array_equal(host.fx, before) and (host.rx != -1.0).all()
            host.rx[:] = hosts[0].rx
        e.close()
    same = bit_identical(hosts[0], hosts[1])
    assert all(same.values()), same
