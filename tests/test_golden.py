"""The oracle's arithmetic layer against the committed golden vectors (tests/golden/refmath_golden.npz), which were
produced by the reference's own headers compiled in place (tests/golden/make_golden.py).  Bit-exact.  Unlike
tests/test_oracle_vs_ref.py this needs neither /root/reference nor oracle/_ref, so it pins the oracle everywhere."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle_binding import OracleLibrary

G = np.load(Path(__file__).resolve().parent / "golden" / "refmath_golden.npz")
D = C.POINTER(C.c_double)
I32 = C.POINTER(C.c_int32)
OCTA_FACES = [[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5], [0, 2, 1, 3]]


def p(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(D)


@pytest.fixture(scope="module")
def orc():
    return OracleLibrary().dll


def same(a, b):
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def test_vertex_vertex(orc):
    hits = 0
    for x, y in zip(G["vv_in"], G["vv_out"]):
        o = np.zeros(7)
        hit = orc.orc_detection_vertex_vertex_core(p(x[0:3]), C.c_double(x[6]), p(x[3:6]), C.c_double(x[7]), p(o))
        flt = orc.orc_filter_vertex_vertex(C.c_double(0.2), p(x[0:3]), C.c_double(x[6]), p(x[3:6]), C.c_double(x[7]))
        assert hit == int(y[0]) and flt == int(y[1]) and same(o, y[2:])
        hits += hit
    assert hits > 50


def test_vertex_edge_and_edge_edge(orc):
    hits = 0
    for x, y in zip(G["ve_ee_in"], G["ve_ee_out"]):
        vi, vf, vs, wf, ws = x[:15].reshape(5, 3)
        ri, rj = C.c_double(x[15]), C.c_double(x[16])
        a, b = np.zeros(7), np.zeros(7)
        h1 = orc.orc_detection_vertex_edge_core(p(vi), ri, p(vf), p(vs), rj, p(a))
        f1 = orc.orc_filter_vertex_edge_core(C.c_double(0.1), p(vi), ri, p(vf), p(vs), rj)
        h2 = orc.orc_detection_edge_edge_core(p(vf), p(vs), ri, p(wf), p(ws), rj, p(b))
        f2 = orc.orc_filter_edge_edge_core(C.c_double(0.1), p(vf), p(vs), ri, p(wf), p(ws), rj)
        assert (h1, f1, h2, f2) == (int(y[0]), int(y[1]), int(y[9]), int(y[10]))
        assert same(a, y[2:9]) and same(b, y[11:18])
        hits += h1 + h2
    assert hits > 30


def test_vertex_face(orc):
    hits = 0
    for x, y in zip(G["vf_in"], G["vf_out"]):
        vi, va = x[0:3], x[3:21].reshape(6, 3)
        idx = np.asarray(OCTA_FACES[int(x[23])], dtype=np.int32)
        o = np.zeros(7)
        hit = orc.orc_detection_vertex_face_core(p(vi), C.c_double(x[21]), 6, p(va), C.c_double(x[22]), idx.ctypes.data_as(I32), len(idx), p(o))
        flt = orc.orc_filter_vertex_face(C.c_double(0.1), p(vi), C.c_double(x[21]), 6, p(va), C.c_double(x[22]), idx.ctypes.data_as(I32), len(idx))
        assert hit == int(y[0]) and flt == int(y[1]) and same(o, y[2:])
        hits += hit
    assert hits > 10


def test_obb_and_vertices(orc):
    n_hit = 0
    for x, y in zip(G["obb_in"], G["obb_out"]):
        A, B, q, pos, h = x[0:15], x[15:30], x[30:34], x[34:37], float(x[37])
        assert orc.orc_obb_intersect(p(A), p(B)) == int(y[0])
        oa, gv = np.zeros(15), np.zeros(3)
        orc.orc_compute_obb(p(A), p(pos), p(q), C.c_double(h), C.c_double(0.05), p(oa))
        orc.orc_get_vertex(p(B[:3]), p(pos), C.c_double(h), p(q), p(gv))
        assert same(oa, y[1:16]) and same(gv, y[16:19])
        n_hit += int(y[0])
    assert 50 < n_hit < len(G["obb_in"]) - 50


@pytest.mark.parametrize("law", [0, 1, 2])
def test_contact_force_core(orc, law):
    for x, y in zip(G["law_in"][law], G["law_out"][law]):
        prm, n, dn = x[0:8], x[8:11], float(x[11])
        pi, pj, vi, vj, wi, wj, cp = x[12:33].reshape(7, 3)
        ft, mom, f = x[33:36].copy(), x[36:39].copy(), np.zeros(3)
        orc.orc_contact_force_core(law, C.c_double(dn), p(n), C.c_double(1e-4), p(prm), C.c_double(x[39]), C.c_double(x[40]), p(ft), p(cp),
                                   p(pi), p(vi), p(f), p(mom), p(wi), p(pj), p(vj), p(wj))
        assert same(np.concatenate([ft, mom, f]), y)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_drivers(orc, kind):
    hits = 0
    for x, y in zip(G["drv_in"][kind], G["drv_out"][kind]):
        prm = x[0:7] if kind == 0 else x[0:4]
        o = np.zeros(7)
        hit = orc.orc_driver_detector(kind, p(prm), C.c_double(x[10]), p(x[7:10]), p(o))
        flt = orc.orc_driver_filter(kind, p(prm), C.c_double(x[10]), p(x[7:10]))
        assert hit == int(y[0]) and flt == int(y[1]) and same(o, y[2:])
        hits += hit
    assert hits > 20
