"""Cost-balanced decomposition (edb_set_brick_cuts): the product's ghost-layer plan and migration classifier — host code of
exadem_b200/csrc/comm.cuh and kernels_migrate.cuh, compiled for the CPU by tests/host_check_plan.cu — against the numpy mirror in
exadem_b200.parallel (which tests/test_parallel_gloo.py checks against the single-domain periodic images), for uniform bricks and for
bricks cut at non-uniform positions, on 2-, 4- and 8-rank process grids."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from exadem_b200 import capi, parallel

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def plan_lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    so = tmp_path_factory.mktemp("plan") / "libhost_check_plan.so"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                    "-I", str(ROOT / "include"), str(ROOT / "tests" / "host_check_plan.cu"), "-o", str(so)], check=True, capture_output=True)
    return C.CDLL(str(so))


def _arr(c):
    return None if c is None else (C.c_double * len(c))(*[float(x) for x in c])


CASES = {
    # world, periodic, global cells per axis, cell-index cuts per axis (None = uniform)
    "uniform-2": (2, (1, 1, 1), (8, 4, 4), None),
    "cut-2": (2, (1, 1, 1), (8, 4, 4), [[0, 3, 8], [0, 4], [0, 4]]),
    "cut-2-wall": (2, (0, 1, 1), (8, 4, 4), [[0, 6, 8], [0, 4], [0, 4]]),
    "cut-4": (4, (1, 1, 0), (8, 8, 4), [[0, 3, 8], [0, 5, 8], [0, 4]]),
    "cut-8": (8, (1, 1, 1), (8, 6, 6), [[0, 2, 8], [0, 4, 6], [0, 1, 6]]),
    "cut-x-only-4": (4, (1, 1, 1), (8, 8, 4), [[0, 5, 8], None, None]),
}


@pytest.mark.parametrize("name", list(CASES))
def test_cxx_plan_equals_numpy_mirror(plan_lib, name):
    world, periodic, gcells, cell_cuts = CASES[name]
    pg = parallel.process_grid(world)
    cs = 2.5
    glo, ghi = (0.0, 0.0, 0.0), tuple(cs * n for n in gcells)
    if cell_cuts is None:
        cuts = None
        full = [[ghi[d] * k / pg[d] for k in range(pg[d] + 1)] for d in range(3)]
    else:
        full_cells = [cell_cuts[d] if cell_cuts[d] is not None else [gcells[d] * k // pg[d] for k in range(pg[d] + 1)] for d in range(3)]
        full = parallel.cuts_to_positions(full_cells, glo, ghi, gcells)
        cuts = [full[d] if cell_cuts[d] is not None else None for d in range(3)]
    nc_axis = parallel.cells_per_brick(full, cs, pg)
    for rank in range(world):
        lo, hi = parallel.brick_bounds(glo, ghi, pg, rank, full)
        cfg = capi.make_config(lo, hi, cs, periodic, capi.MODE_SPHERE, 0, glo=glo, ghi=ghi, rank=rank, nranks=world, pgrid=pg)
        recv = np.zeros((20000, 10))
        send = np.zeros((20000, 6))
        nr, ns = C.c_int(), C.c_int()
        a = [_arr(c) for c in (cuts or [None, None, None])]
        rc = plan_lib.chk_plan(C.byref(cfg), a[0], a[1], a[2], recv.ctypes.data_as(C.POINTER(C.c_double)), len(recv), C.byref(nr),
                               send.ctypes.data_as(C.POINTER(C.c_double)), len(send), C.byref(ns))
        assert rc == 0 and nr.value <= len(recv) and ns.value <= len(send)
        # (i) the ghost table of this rank, entry by entry in enumeration order
        mirror = parallel.ghost_plan(glo, ghi, periodic, pg, rank, None, nc_axis)
        got = [((int(r[0]), int(r[1]), int(r[2])), int(r[3]), (int(r[4]), int(r[5]), int(r[6])), (r[7], r[8], r[9])) for r in recv[:nr.value]]
        assert len(got) == len(mirror)
        for g_, m_ in zip(got, mirror):
            assert g_[:3] == (tuple(m_[0]), m_[1], tuple(m_[2])) and np.allclose(g_[3], m_[3], atol=1e-12), (rank, g_, m_)
        # (ii) the send lists: per destination, cells in the RECEIVER's enumeration order, linearised in the receiver's / our grid
        me = parallel.rank_coords(rank, pg)
        g_me = [nc_axis[d][me[d]] + 2 for d in range(3)]
        sp = parallel.send_plan(glo, ghi, periodic, pg, rank, None, nc_axis)
        expect = []
        for other in sorted(sp):
            oc = parallel.rank_coords(other, pg)
            g_o = [nc_axis[d][oc[d]] + 2 for d in range(3)]
            for (g, s, sh) in sp[other]:
                expect.append((other, g[0] + g_o[0] * (g[1] + g_o[1] * g[2]), s[0] + g_me[0] * (s[1] + g_me[1] * s[2]), tuple(sh)))
        gots = [(int(r[0]), int(r[1]), int(r[2]), (r[3], r[4], r[5])) for r in send[:ns.value]]
        assert len(gots) == len(expect)
        for g_, e_ in zip(gots, expect):
            assert g_[:3] == e_[:3] and np.allclose(g_[3], e_[3], atol=1e-12), (rank, g_, e_)
    # (iii) the migration classifier agrees with owner_rank, also exactly on the cut planes and outside the box
    rng = np.random.default_rng(11)
    pos = rng.random((4000, 3)) * np.array(ghi) * 1.2 - 0.1 * np.array(ghi)
    for d in range(3):
        pos[: 3 * len(full[d]): 3, d] = np.repeat(full[d], 1)[: len(pos[: 3 * len(full[d]): 3, d])]
    out = np.zeros(len(pos), dtype=np.int32)
    cfg = capi.make_config(glo, ghi, cs, periodic, capi.MODE_SPHERE, 0, glo=glo, ghi=ghi, rank=0, nranks=world, pgrid=pg)
    a = [_arr(c) for c in (cuts or [None, None, None])]
    plan_lib.chk_brick_of(C.byref(cfg), a[0], a[1], a[2], np.ascontiguousarray(pos).ctypes.data_as(C.POINTER(C.c_double)), len(pos),
                          out.ctypes.data_as(C.POINTER(C.c_int)))
    assert np.array_equal(out, parallel.owner_rank(pos, glo, ghi, pg, cuts if cuts is not None and all(c is not None for c in cuts) else
                                                   (None if cuts is None else full)))


def test_set_brick_cuts_is_exported(product_lib):
    assert hasattr(product_lib.dll, "edb_set_brick_cuts") and hasattr(product_lib.dll, "edb_dem_cost_model")
