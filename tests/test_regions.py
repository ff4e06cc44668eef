"""The `region` slot of reset_force_moment / freeze_particles (reset_force_moment.cu:36-57, freeze_particles.cu:47-66): boxes, id ranges,
quadrics and CSG expressions over them.  Known answers on the oracle, then the CUDA path against the oracle."""
import numpy as np
import pytest

from exadem_b200 import capi, workloads
from oracle_binding import OracleEngine


def setup(engine_cls, lib):
    wl = workloads.spheres_box((8, 8, 8), jitter=0.02, omega_sigma=0.5)
    e = wl.setup(engine_cls(lib, wl.config()))
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 2, 0)
    sphere = np.diag([1.0, 1.0, 1.0, 0.0])
    c = np.array([4.0, 4.0, 4.0])
    sphere[:3, 3] = -c
    sphere[3, :3] = -c
    sphere[3, 3] = c @ c - 2.5 ** 2                      # |r - c|^2 - R^2 <= 0
    e.set_particle_regions([dict(bounds=((0, 0, 0), (8, 8, 3.0))), dict(quadric=sphere), dict(ids=(100, 200))])
    return wl, e


def masks(pa):
    r = pa.vec("r")
    box = r[:, 2] < 3.0
    ball = ((r - 4.0) ** 2).sum(axis=1) <= 2.5 ** 2
    ids = (pa.id >= 100) & (pa.id < 200)
    return box, ball, ids


def test_regions_known_answers(oracle_lib):
    wl, e = setup(OracleEngine, oracle_lib)
    before = e.download_particles()
    box, ball, ids = masks(before)
    assert box.sum() > 50 and ball.sum() > 30 and ids.sum() == 100
    # region: BOX and not BALL  ->  [0, 1, NOT, AND]
    e.freeze_particles_region([0, 1, capi.CSG_NOT, capi.CSG_AND])
    after = e.download_particles()
    hit = box & ~ball
    assert (after.vec("v")[hit] == 0).all() and (after.vec("vrot")[hit] == 0).all() and (after.vec("f")[hit] == 0).all()
    assert np.array_equal(after.vec("v")[~hit], before.vec("v")[~hit]) and np.array_equal(after.vec("f")[~hit], before.vec("f")[~hit])
    # region: BALL or IDS -> reset forces only
    e.reset_force_moment_region([1, 2, capi.CSG_OR])
    last = e.download_particles()
    hit2 = ball | ids
    assert (last.vec("f")[hit2] == 0).all() and (last.vec("mom")[hit2] == 0).all()
    assert np.array_equal(last.vec("v"), after.vec("v"))
    # no expression = every particle (the operator without a region)
    e.reset_force_moment_region([])
    assert (e.download_particles().vec("f") == 0).all()


@pytest.mark.gpu
def test_gpu_regions_vs_oracle(product_lib, oracle_lib):
    from exadem_b200 import Engine
    from parity import bit_identical
    _, g = setup(Engine, product_lib)
    _, c = setup(OracleEngine, oracle_lib)
    for expr, fn in (([0, 1, capi.CSG_NOT, capi.CSG_AND], "freeze_particles_region"), ([1, 2, capi.CSG_OR], "reset_force_moment_region"),
                     ([2, capi.CSG_NOT], "freeze_particles_region")):
        for e in (g, c):
            getattr(e, fn)(expr)
        pg, pc = g.download_particles(), c.download_particles()
        # the set of particles touched is identical; everything else within the usual bar
        for f in ("v", "vrot", "f", "mom"):
            assert np.array_equal(pg.vec(f) == 0, pc.vec(f) == 0), (fn, f)
    with pytest.raises(capi.EdbError):
        g.freeze_particles_region([0, capi.CSG_AND])
    with pytest.raises(capi.EdbError):
        g.freeze_particles_region([7])
