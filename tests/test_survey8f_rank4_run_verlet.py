"""GPU tests of edb_run_verlet with inner bonds (alone and with a moving driver), the traction deck through the C ABI, and the golden-file
form of the moving-driver runs.  All green on a B200 (profiles/r2_gpu_tests_first.log)."""
import numpy as np
import pytest

import test_sim_golden as SG
import test_survey8f_rank4_inner_bonds as IB

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [c for c in SG.cases() if c[0] not in SG.GPU_VERIFIED], ids=lambda c: c[0])
def test_cuda_path_matches_the_moving_driver_golden_runs(product_lib, case):
    """golden-file form of test_gpu_run_verlet_with_moving_drivers[piston | moving_mesh] (green live on a B200)"""
    from exadem_b200 import Engine
    name, wl, steps, reb = case
    SG.compare(name, SG.run(Engine, product_lib, wl, steps, reb), 1e-10)


def test_gpu_traction_deck(product_lib):
    """the same known answer through the C ABI: bonds created by edb_stick_polyhedra, the interface breaks at dn = sqrt(4 A Gc / kn)"""
    from exadem_b200 import Engine
    gc = 2e-2
    wl, ibp = IB.traction_pair(gc)
    e = wl.setup(Engine(product_lib, wl.config()))
    e.set_inner_bond_params(1, ibp)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.stick_polyhedra_ids(1.5e-4, wl.particles.id, np.zeros(wl.particles.n))
    b = e.export_inner_bonds()
    assert len(b) == 4 and np.allclose(b["weight"], 0.25) and np.allclose(b["criteria_1"], 2 * 0.98 ** 2 * gc, rtol=1e-12)
    e.update_particle_neighbors(wl.rcut_inc, True)
    assert e.update_interfaces() == 1
    dn_fail, step_dn = np.sqrt(4 * 0.98 ** 2 * gc / 1e3), 10.0 * wl.dt
    broke_at = None
    for it in range(int(dn_fail / step_dn) + 10):
        if IB.step_block(e, wl):
            broke_at = it
            break
    assert broke_at is not None and abs(broke_at * step_dn - dn_fail) <= step_dn, (broke_at, dn_fail / step_dn)
    assert len(e.export_inner_bonds()) == 0      # the broken bonds left wave 13 at the rebuild


def test_gpu_run_verlet_with_inner_bonds(product_lib, oracle_lib):
    """edb_run_verlet runs the fragmentation macro (criterion -> friction cap -> rebuild on a break, inner_bond_polyhedron in
    compute_force) when bonds exist: against the oracle, and against the same steps driven operator by operator."""
    from exadem_b200 import Engine
    from oracle_binding import OracleEngine
    from parity import compare_interaction_sets, compare_particles
    wl, ibp = IB.glued_block()
    engines = [wl.setup(Engine(product_lib, wl.config())), wl.setup(OracleEngine(oracle_lib, wl.config())),
               wl.setup(Engine(product_lib, wl.config()))]
    IB.setup_block(engines, wl, ibp)
    gpu, cpu, ops = engines
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 20, -1) for e in (gpu, cpu)]
    broken = sum(IB.step_block(ops, wl) for _ in range(20))
    assert nb[0] == nb[1] == 1 and broken == 16      # the top layer tears off once; that break is the only rebuild
    IB.compare_bonds(gpu, cpu)
    compare_particles(gpu.download_particles(), cpu.download_particles(), rtol=1e-10)
    compare_interaction_sets(gpu.export_interactions(), cpu.export_interactions(), payload_rtol=1e-10)
    same = __import__("parity").bit_identical(gpu.download_particles(), ops.download_particles())
    assert all(same.values()), same


def test_gpu_run_verlet_with_inner_bonds_and_a_moving_driver(product_lib, oracle_lib):
    """moving drivers and inner bonds in one edb_run_verlet call (operator loop: driver pushes, fracture criterion, rebuild on a break or
    on a displacement, bond law + contacts, driver reaction) against the oracle"""
    from exadem_b200 import Engine
    from oracle_binding import OracleEngine
    from parity import compare_interaction_sets, compare_particles, field_close
    wl, ibp = IB.block_hit_by_a_ball()
    engines = [wl.setup(Engine(product_lib, wl.config())), wl.setup(OracleEngine(oracle_lib, wl.config()))]
    IB.setup_block(engines, wl, ibp)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 40, 0) for e in engines]
    assert nb[0] == nb[1] >= 3
    gpu, cpu = engines
    rec = IB.compare_bonds(gpu, cpu)
    assert 0 < len(rec) < 416
    compare_particles(gpu.download_particles(), cpu.download_particles(), rtol=1e-10)
    compare_interaction_sets(gpu.export_interactions(), cpu.export_interactions(), payload_rtol=1e-10)
    nd = len(wl.drivers)
    for a, b in zip(gpu.get_drivers(nd), cpu.get_drivers(nd)):
        for k in ("center", "vel", "forces"):
            field_close(np.atleast_1d(a[k]), np.atleast_1d(b[k]), 1e-10, f"driver {k}")
