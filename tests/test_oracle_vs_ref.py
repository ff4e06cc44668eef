"""Pins the oracle's arithmetic layer (oracle/dem_math.hpp) BIT-FOR-BIT against the reference's own header-only
functions compiled in place (oracle/_ref/libexadem_refmath.so, built by oracle/Makefile from /root/reference).

The GPU box has no /root/reference: it uses the prebuilt oracle/_ref/*.so that travels with the repo snapshot.
If neither exists the tests are skipped (and DESIGN.md's "parity unpinned" caveat applies)."""
import ctypes as C

import numpy as np
import pytest

from oracle_binding import OracleLibrary, ref_math

REF = ref_math()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (no /root/reference here)")

D = C.POINTER(C.c_double)
I32 = C.POINTER(C.c_int32)


def p(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(D)


@pytest.fixture(scope="module")
def orc():
    return OracleLibrary().dll


def _out7():
    return np.zeros(7)


def rng_points(rng, n, scale=1.0):
    return rng.normal(size=(n, 3)) * scale


def test_vertex_vertex(orc):
    rng = np.random.default_rng(1)
    for k in range(4000):
        pi, pj = rng_points(rng, 2, 0.6)
        ri, rj = rng.uniform(0.1, 0.6, 2)
        if k % 7 == 0:  # near-threshold: place pj at distance ~ ri+rj(+rV)
            d = pj - pi
            pj = pi + d / np.linalg.norm(d) * (ri + rj + (0.2 if k % 2 else 0.0)) * (1 + rng.integers(-3, 4) * 1e-16)
        a, b = _out7(), _out7()
        ha = REF.ref_detection_vertex_vertex_core(p(pi), C.c_double(ri), p(pj), C.c_double(rj), p(a))
        hb = orc.orc_detection_vertex_vertex_core(p(pi), C.c_double(ri), p(pj), C.c_double(rj), p(b))
        assert ha == hb and a.tobytes() == b.tobytes()
        fa = REF.ref_filter_vertex_vertex(C.c_double(0.2), p(pi), C.c_double(ri), p(pj), C.c_double(rj))
        fb = orc.orc_filter_vertex_vertex(C.c_double(0.2), p(pi), C.c_double(ri), p(pj), C.c_double(rj))
        assert fa == fb


def test_vertex_edge_and_edge_edge(orc):
    rng = np.random.default_rng(2)
    hits = 0
    for k in range(6000):
        vi, vf, vs, wf, ws = rng_points(rng, 5, 0.5)
        ri, rj = rng.uniform(0.05, 0.4, 2)
        a, b = _out7(), _out7()
        ha = REF.ref_detection_vertex_edge_core(p(vi), C.c_double(ri), p(vf), p(vs), C.c_double(rj), p(a))
        hb = orc.orc_detection_vertex_edge_core(p(vi), C.c_double(ri), p(vf), p(vs), C.c_double(rj), p(b))
        assert ha == hb and a.tobytes() == b.tobytes()
        hits += ha
        assert REF.ref_filter_vertex_edge_core(C.c_double(0.1), p(vi), C.c_double(ri), p(vf), p(vs), C.c_double(rj)) == \
            orc.orc_filter_vertex_edge_core(C.c_double(0.1), p(vi), C.c_double(ri), p(vf), p(vs), C.c_double(rj))
        a, b = _out7(), _out7()
        ha = REF.ref_detection_edge_edge_core(p(vf), p(vs), C.c_double(ri), p(wf), p(ws), C.c_double(rj), p(a))
        hb = orc.orc_detection_edge_edge_core(p(vf), p(vs), C.c_double(ri), p(wf), p(ws), C.c_double(rj), p(b))
        assert ha == hb and a.tobytes() == b.tobytes()
        hits += ha
        assert REF.ref_filter_edge_edge_core(C.c_double(0.1), p(vf), p(vs), C.c_double(ri), p(wf), p(ws), C.c_double(rj)) == \
            orc.orc_filter_edge_edge_core(C.c_double(0.1), p(vf), p(vs), C.c_double(ri), p(wf), p(ws), C.c_double(rj))
    assert hits > 200  # the sample actually exercises the contact branches


def _octa_faces():
    # octahedron of example/polyhedra/rotating_drum/octahedron.shp topology (8 triangles) + one quad for the polygon loop
    return [[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5], [0, 2, 1, 3]]


def test_vertex_face(orc):
    rng = np.random.default_rng(3)
    base = np.array([[0.5, 0, 0], [-0.5, 0, 0], [0, 0.5, 0], [0, -0.5, 0], [0, 0, 0.5], [0, 0, -0.5]])
    faces = _octa_faces()
    stream = []
    for f in faces:
        stream += [len(f)] + f
    stream = np.asarray(stream, dtype=np.int32)
    hits = 0
    for k in range(3000):
        va = base + rng.normal(size=(6, 3)) * 0.01 + rng.normal(size=3) * 0.2
        vi = rng.normal(size=3) * 0.45
        ri, rj = rng.uniform(0.02, 0.2, 2)
        j = int(rng.integers(0, len(faces)))
        idx = np.asarray(faces[j], dtype=np.int32)
        a, b = _out7(), _out7()
        ha = REF.ref_detection_vertex_face_core(p(vi), C.c_double(ri), 6, p(va), C.c_double(1.0), j, len(faces),
                                                stream.ctypes.data_as(I32), len(stream), C.c_double(rj), p(a))
        hb = orc.orc_detection_vertex_face_core(p(vi), C.c_double(ri), 6, p(va), C.c_double(rj),
                                                idx.ctypes.data_as(I32), len(idx), p(b))
        assert ha == hb and a.tobytes() == b.tobytes()
        hits += ha
        fa = REF.ref_filter_vertex_face(C.c_double(0.1), p(vi), C.c_double(ri), 6, p(va), C.c_double(1.0), j, len(faces),
                                        stream.ctypes.data_as(I32), len(stream), C.c_double(rj))
        fb = orc.orc_filter_vertex_face(C.c_double(0.1), p(vi), C.c_double(ri), 6, p(va), C.c_double(rj),
                                        idx.ctypes.data_as(I32), len(idx))
        assert fa == fb
    assert hits > 50


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def test_get_vertex_and_quat_rotation(orc):
    rng = np.random.default_rng(4)
    for _ in range(2000):
        q, lv, pos = rand_quat(rng), rng.normal(size=3), rng.normal(size=3) * 10
        h = float(rng.uniform(0.5, 2.0))
        a, b, c = np.zeros(3), np.zeros(3), np.zeros(3)
        REF.ref_get_vertex(p(lv), p(pos), C.c_double(h), p(q), p(a))
        orc.orc_get_vertex(p(lv), p(pos), C.c_double(h), p(q), p(b))
        assert a.tobytes() == b.tobytes()
        # the shim's Quaternion*Vec3d is the vendored Rockable formula (type/quat.hpp:79-90)
        REF.ref_rockable_quat_rotate(p(q), p(h * lv), p(c))
        assert np.array_equal(pos + c, a)


def test_obb(orc):
    rng = np.random.default_rng(5)
    n_hit = 0
    for k in range(5000):
        def rand_obb(scale):
            q = rand_quat(rng)
            w, x, y, z = q
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            return np.concatenate([rng.normal(size=3) * scale, R[:, 0], R[:, 1], R[:, 2], rng.uniform(0.1, 0.5, 3)])
        A, B = rand_obb(0.5), rand_obb(0.5)
        ra, rb = REF.ref_obb_intersect(p(A), p(B)), orc.orc_obb_intersect(p(A), p(B))
        assert ra == rb
        n_hit += ra
        q, pos, h = rand_quat(rng), rng.normal(size=3), float(rng.choice([1.0, 0.7, 1.3]))
        oa, ob = np.zeros(15), np.zeros(15)
        REF.ref_compute_obb(p(A), p(pos), p(q), C.c_double(h), C.c_double(0.05), p(oa))
        orc.orc_compute_obb(p(A), p(pos), p(q), C.c_double(h), C.c_double(0.05), p(ob))
        assert oa.tobytes() == ob.tobytes()
    assert 500 < n_hit < 4500


@pytest.mark.parametrize("law", [0, 1, 2])
def test_contact_force_core(orc, law):
    rng = np.random.default_rng(10 + law)
    for k in range(3000):
        prm = np.array([rng.uniform(1e3, 1e5), rng.uniform(1e3, 1e5), rng.uniform(0, 1), rng.uniform(0, 1),
                        rng.uniform(0, 1), rng.uniform(0, 10), rng.uniform(0.01, 0.2), rng.uniform(0, 5)])
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        dn = float(rng.uniform(-0.05, 0.02 if law == 1 else 0.0))
        pi, pj, vi, vj, wi, wj, cp = rng.normal(size=(7, 3))
        if k % 5 == 0:
            wi = np.zeros(3)  # quirk Q1: NaN threshold keeps the rolling moment
        ft0, mom0 = rng.normal(size=(2, 3)) * (0.0 if k % 3 == 0 else 5.0)
        meff, reff, dt = float(rng.uniform(0.1, 2)), float(rng.uniform(0.1, 1)), 1e-4
        res = []
        for fn in (REF.ref_contact_force_core, orc.orc_contact_force_core):
            ft, mom, f = ft0.copy(), mom0.copy(), np.zeros(3)
            fn(law, C.c_double(dn), p(n), C.c_double(dt), p(prm), C.c_double(meff), C.c_double(reff), p(ft), p(cp),
               p(pi), p(vi), p(f), p(mom), p(wi), p(pj), p(vj), p(wj))
            res.append(np.concatenate([ft, mom, f]).tobytes())
        assert res[0] == res[1]
        a, b = np.zeros(3), np.zeros(3)
        REF.ref_compute_moments(p(cp), p(pi), p(vi), p(wj), p(a))
        orc.orc_compute_moments(p(cp), p(pi), p(vi), p(wj), p(b))
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_drivers(orc, kind):
    rng = np.random.default_rng(20 + kind)
    hits = 0
    for k in range(3000):
        if kind == 0:
            axis = np.array([1.0, 0.0, 1.0])
            prm = np.concatenate([[rng.uniform(2, 5)], axis, axis * rng.normal(size=3)])
        elif kind == 1:
            nrm = rng.normal(size=3)
            nrm /= np.linalg.norm(nrm)
            prm = np.concatenate([[rng.normal()], nrm])
        else:
            prm = np.concatenate([[rng.uniform(2, 5)], rng.normal(size=3)])
        pt = rng.normal(size=3) * 3
        rcut = float(rng.uniform(0.1, 0.8))
        a, b = _out7(), _out7()
        ha = REF.ref_driver_detector(kind, p(prm), C.c_double(rcut), p(pt), p(a))
        hb = orc.orc_driver_detector(kind, p(prm), C.c_double(rcut), p(pt), p(b))
        assert ha == hb and a.tobytes() == b.tobytes()
        hits += ha
        assert REF.ref_driver_filter(kind, p(prm), C.c_double(rcut), p(pt)) == \
            orc.orc_driver_filter(kind, p(prm), C.c_double(rcut), p(pt))
    assert hits > 100


def test_interaction_pair_order_and_equality():
    """The oracle's sort / history-match key (owner id, partner id, type, swap, owner sub, partner sub) against the reference's own
    InteractionPair::operator< and operator== (interaction_pair.hpp:186-249) — ties in every field exercised."""
    import ctypes as C
    import numpy as np
    from oracle_binding import OracleLibrary, ref_math
    ref = ref_math()
    if ref is None or not hasattr(ref, "ref_pair_less"):
        import pytest
        pytest.skip("oracle/_ref not built (reference tree absent)")
    orc = OracleLibrary().dll
    rng = np.random.default_rng(3)
    n_less = n_eq = 0
    for _ in range(20000):
        a = np.array([rng.integers(0, 3), rng.integers(0, 3), rng.integers(0, 2), rng.integers(0, 2), rng.integers(0, 4), rng.integers(0, 2)], dtype=np.uint64)
        b = np.array([rng.integers(0, 3), rng.integers(0, 3), rng.integers(0, 2), rng.integers(0, 2), rng.integers(0, 4), rng.integers(0, 2)], dtype=np.uint64)
        pa, pb = a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)
        l_ref, l_orc = ref.ref_pair_less(pa, pb), orc.orc_pair_less(pa, pb)
        e_ref, e_orc = ref.ref_pair_equal(pa, pb), orc.orc_pair_equal(pa, pb)
        assert l_ref == l_orc and e_ref == e_orc, (a, b, l_ref, l_orc, e_ref, e_orc)
        n_less += l_ref
        n_eq += e_ref
    assert n_less > 5000 and n_eq > 20


def test_face_areas():
    """shape::compute_face_areas / get_face_area(i, h) (shape.hpp:673-697, :359-362) — the area that enters the rupture criterion of the
    inner bonds (stick_polyhedra.cpp:290-305) — against the oracle's fan triangulation, on the octahedron, the cube and random
    distortions of them."""
    import ctypes as C
    import numpy as np
    from exadem_b200 import capi, workloads
    from oracle_binding import OracleLibrary, ref_math
    ref = ref_math()
    if ref is None or not hasattr(ref, "ref_face_areas"):
        import pytest
        pytest.skip("oracle/_ref not built (reference tree absent)")
    orc = OracleLibrary().dll
    rng = np.random.default_rng(17)
    for spec in (workloads.OCTAHEDRON, workloads.cube_shape(), workloads.cube_shape(0.49, 0.01)):
        for rep in range(20):
            v = np.asarray(spec["vertices"], dtype=np.float64) * (1.0 + (0.2 * rng.normal(size=(len(spec["vertices"]), 3)) if rep else 0.0))
            s = capi.make_shape(v, spec["edges"], spec["faces"], spec["obb"], spec["radius"])
            h = float(rng.uniform(0.5, 2.0))
            a, b = np.zeros(s.nf), np.zeros(s.nf)
            for fn, out in ((ref.ref_face_areas, a), (orc.orc_face_areas, b)):
                fn.restype = None
                fn(C.c_int(s.nv), s.vertices, C.c_int(s.ne), s.edges, C.c_int(s.nf), s.faces, C.c_int(s.faces_len), C.c_double(h),
                   out.ctypes.data_as(C.POINTER(C.c_double)))
            assert np.array_equal(a, b) and (a > 0).all()
