"""S1 of SURVEY 8(d) = BASELINE configs[0]: the reference's own CPU-runnable deck example/spheres/deposit/add_xyz.msp — the 736 sphere
centres of input_deposit_xyz.xyz (committed as tests/golden/deposit_xyz.npz by make_deposit_fixture.py) tiled 4 x 4 to 11 776 spheres,
deck parameters, velocities from the counter-based splitmix64 generator.  The deck starts as a dilute cloud 1.3 above the surface
driver, so the oracle first advances it to the instant the layer lands (pre-settle); parity is then checked from that state:
1 step and 300 steps (several neighbour rebuilds), plain Hooke (configs[0]) and the deck's own hooke_dmt variant."""
from pathlib import Path

import numpy as np
import pytest

from exadem_b200 import workloads
from oracle_binding import OracleEngine
from parity import compare_interaction_sets, compare_particles

XYZ = np.load(Path(__file__).resolve().parent / "golden" / "deposit_xyz.npz")["xyz"]
PRESETTLE = 4200


def test_counter_based_generator_is_order_independent():
    ids = np.arange(1000, dtype=np.uint64)
    a = workloads.counter_normal(0xD301, ids)
    perm = np.random.default_rng(1).permutation(1000)
    b = workloads.counter_normal(0xD301, ids[perm])
    assert np.array_equal(a[perm], b)
    # splitmix64 known answers (reference implementation of Vigna's splitmix64.c seeded with 0: first three outputs)
    got = workloads.splitmix64(0, np.arange(3, dtype=np.uint64))
    assert [int(x) for x in got] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    u = workloads.counter_uniform(7, np.arange(100000, dtype=np.uint64))
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    assert abs(a.mean()) < 0.1 and abs(a.std() - 1.0) < 0.1


def test_deck_fixture_and_tiling():
    assert XYZ.shape == (736, 3)
    wl = workloads.deposit_xyz(XYZ, (4, 4))
    assert wl.particles.n == 11776 and wl.cell_size == 1.0 and wl.periodic == (1, 1, 0)
    assert np.allclose(wl.particles.radius, 0.4) and wl.drivers[0]["offset"] == -1.0
    assert np.array_equal(wl.particles.rx[:736] + 10.0, wl.particles.rx[736:1472])


def settled(oracle_lib, variant, tiles):
    """the deck advanced by the oracle (all host threads) to the landing of the first layer; returned as a new initial condition"""
    import os
    wl = workloads.deposit_xyz(XYZ, tiles, variant)
    e = wl.setup(OracleEngine(oracle_lib, wl.config()))
    e.set_threads(os.cpu_count() or 1)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, PRESETTLE, 0)
    wl.particles = e.download_particles()
    e.close()
    return wl


def test_deck_lands_on_the_surface_driver(oracle_lib):
    wl = settled(oracle_lib, "hooke", (2, 2))
    e = wl.setup(OracleEngine(oracle_lib, wl.config()))
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 300, 0)
    counts, active = e.interaction_counts()
    assert counts[0] > 100 and counts[5] > 20 and active[5] > 0, (counts, active)     # wave 5 = sphere / surface driver


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["hooke", "dmt"])
def test_gpu_s1_deck_vs_oracle(product_lib, oracle_lib, variant):
    from exadem_b200 import Engine
    wl = settled(oracle_lib, variant, (4, 4))
    gpu = wl.setup(Engine(product_lib, wl.config()))
    cpu = wl.setup(OracleEngine(oracle_lib, wl.config()))
    for e in (gpu, cpu):
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 1, 0)
    compare_interaction_sets(gpu.export_interactions(), cpu.export_interactions(), payload_rtol=1e-10)
    compare_particles(gpu.download_particles(), cpu.download_particles(), rtol=1e-10)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 300, 0) for e in (gpu, cpu)]
    assert nb[0] == nb[1] and nb[0] >= 1
    n = compare_interaction_sets(gpu.export_interactions(), cpu.export_interactions(), payload_rtol=1e-10)
    compare_particles(gpu.download_particles(), cpu.download_particles(), rtol=1e-10)
    _, active = gpu.interaction_counts()
    assert n > 1000 and active[5] > 0
