"""dem_cost_model (src/interaction/dem_cost_model.cpp:63-92): cost of every inner cell = 3 N + 1 + the type-weighted number of
interaction records stored with the cell's particles.  Known answers on the oracle (the sum rule against the interaction export),
then the CUDA path against the oracle — integer-valued, so the bar is equality — and the cuts the balancer derives from the costs."""
import numpy as np
import pytest

from exadem_b200 import parallel, workloads
from oracle_binding import OracleEngine

WEIGHT = np.array([1, 3, 5, 4, 1, 1, 1, 1, 3, 5, 4, 3, 5, 1], dtype=np.float64)


def spheres():
    return workloads.rotating_drum(6000, jitter=0.02)


def polyhedra():
    return workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.78)


def build(engine_cls, lib, wl, steps=3):
    e = wl.setup(engine_cls(lib, wl.config()))
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, steps, 0)
    return e


@pytest.mark.parametrize("make", [spheres, polyhedra])
def test_oracle_cost_sum_rule(oracle_lib, make):
    wl = make()
    e = build(OracleEngine, oracle_lib, wl)
    cost = e.dem_cost_model()
    nreal = wl.particles.n
    rec = e.export_interactions()
    stored = rec[rec["ghost"] != 2]                      # PartnerGhost copies sit in ghost cells: not counted
    expect_total = 3.0 * nreal + cost.size + WEIGHT[stored["type"]].sum()
    assert cost.sum() == expect_total
    assert (cost >= 1.0).all() and (cost == np.round(cost)).all()
    # an empty cell costs exactly 1, a cell never costs less than 3 N + 1
    counts, _ = e.interaction_counts()
    assert counts.sum() > 0 and cost.max() > 1.0


def test_oracle_cost_without_lists_is_3n_plus_1(oracle_lib):
    wl = workloads.spheres_box((6, 6, 6), jitter=0.0)
    wl.rcut_inc = 0.0                                     # pitch 1.0, r = 0.5: no pair closer than 2 r + 0 -> no record at all
    e = wl.setup(OracleEngine(oracle_lib, wl.config()))
    e.update_particle_neighbors(0.0, True)
    counts, _ = e.interaction_counts()
    cost = e.dem_cost_model()
    nx = cost.shape[2]
    assert cost.sum() == 3.0 * wl.particles.n + cost.size + WEIGHT[0] * counts[0]
    assert cost.shape == (nx, nx, nx)


@pytest.mark.gpu
@pytest.mark.parametrize("make", [spheres, polyhedra])
def test_gpu_cost_model_equals_oracle(product_lib, oracle_lib, make):
    from exadem_b200 import Engine
    wl = make()
    g = build(Engine, product_lib, wl)
    c = build(OracleEngine, oracle_lib, wl)
    cg, cc = g.dem_cost_model(), c.dem_cost_model()
    assert cg.shape == cc.shape
    assert np.array_equal(cg, cc)


def test_balanced_cuts_equalise_the_marginal_cost():
    rng = np.random.default_rng(5)
    cost = np.ones((12, 10, 16))
    cost[:, :, :4] += 40.0 * rng.random((12, 10, 4))      # a heap against the low-x wall
    cuts = parallel.balanced_cuts(cost, (4, 2, 1), min_cells=1)
    assert [c[0] for c in cuts] == [0, 0, 0] and [c[-1] for c in cuts] == [16, 10, 12] and [len(c) for c in cuts] == [5, 3, 2]
    assert all(b > a for c in cuts for a, b in zip(c[:-1], c[1:]))
    mx = cost.sum(axis=(0, 1))
    slabs = [mx[a:b].sum() for a, b in zip(cuts[0][:-1], cuts[0][1:])]
    uniform = [mx[a:b].sum() for a, b in zip(range(0, 16, 4), range(4, 20, 4))]
    assert max(slabs) < 0.6 * max(uniform)                # the heap is shared out instead of sitting in one brick
    # every brick keeps at least min_cells cells, also when the cost is concentrated in one plane
    spike = np.zeros((4, 4, 9))
    spike[:, :, 4] = 100.0
    c2 = parallel.balanced_cuts(spike + 1e-9, (3, 1, 1), min_cells=2)
    assert all(b - a >= 2 for a, b in zip(c2[0][:-1], c2[0][1:])) and c2[0][-1] == 9
