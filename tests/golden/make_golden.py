"""Generates tests/golden/refmath_golden.npz from the REFERENCE's own header-only arithmetic compiled in place
(oracle/_ref/libexadem_refmath.so <- oracle/ref_math.cpp + /root/reference headers, built by oracle/Makefile).

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The .npz holds seeded inputs and the reference's outputs for every arithmetic entry point of the hot path; it is what
tests/test_golden.py checks the oracle against where neither /root/reference nor oracle/_ref exists."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle_binding import ref_math  # noqa: E402

D = C.POINTER(C.c_double)
I32 = C.POINTER(C.c_int32)


def p(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(D)


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


OCTA_FACES = [[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5], [0, 2, 1, 3]]
OCTA_BASE = np.array([[0.5, 0, 0], [-0.5, 0, 0], [0, 0.5, 0], [0, -0.5, 0], [0, 0, 0.5], [0, 0, -0.5]])


def main():
    REF = ref_math()
    if REF is None:
        raise SystemExit("oracle/_ref/libexadem_refmath.so is not built (needs /root/reference): make -C oracle ref")
    rng = np.random.default_rng(0xD3A0)
    G = {}
    # vertex-vertex (shape_detection_core.hpp:43-81)
    n = 600
    inp = np.zeros((n, 8)); out = np.zeros((n, 9))
    for k in range(n):
        pi, pj = rng.normal(size=(2, 3)) * 0.6
        ri, rj = rng.uniform(0.1, 0.6, 2)
        if k % 5 == 0:
            d = pj - pi
            pj = pi + d / np.linalg.norm(d) * (ri + rj + (0.2 if k % 2 else 0.0)) * (1 + rng.integers(-3, 4) * 1e-16)
        o = np.zeros(7)
        hit = REF.ref_detection_vertex_vertex_core(p(pi), C.c_double(ri), p(pj), C.c_double(rj), p(o))
        flt = REF.ref_filter_vertex_vertex(C.c_double(0.2), p(pi), C.c_double(ri), p(pj), C.c_double(rj))
        inp[k] = np.concatenate([pi, pj, [ri, rj]]); out[k] = np.concatenate([[hit, flt], o])
    G["vv_in"], G["vv_out"] = inp, out
    # vertex-edge + edge-edge (:83-148, :468-566)
    n = 1000
    inp = np.zeros((n, 17)); out = np.zeros((n, 18))
    for k in range(n):
        pts = rng.normal(size=(5, 3)) * 0.5
        vi, vf, vs, wf, ws = pts
        ri, rj = rng.uniform(0.05, 0.4, 2)
        a, b = np.zeros(7), np.zeros(7)
        h1 = REF.ref_detection_vertex_edge_core(p(vi), C.c_double(ri), p(vf), p(vs), C.c_double(rj), p(a))
        f1 = REF.ref_filter_vertex_edge_core(C.c_double(0.1), p(vi), C.c_double(ri), p(vf), p(vs), C.c_double(rj))
        h2 = REF.ref_detection_edge_edge_core(p(vf), p(vs), C.c_double(ri), p(wf), p(ws), C.c_double(rj), p(b))
        f2 = REF.ref_filter_edge_edge_core(C.c_double(0.1), p(vf), p(vs), C.c_double(ri), p(wf), p(ws), C.c_double(rj))
        inp[k] = np.concatenate([pts.ravel(), [ri, rj]]); out[k] = np.concatenate([[h1, f1], a, [h2, f2], b])
    G["ve_ee_in"], G["ve_ee_out"] = inp, out
    # vertex-face (:245-309, :398-466)
    stream = []
    for f in OCTA_FACES:
        stream += [len(f)] + f
    stream = np.asarray(stream, dtype=np.int32)
    n = 600
    inp = np.zeros((n, 3 + 18 + 3)); out = np.zeros((n, 9))
    for k in range(n):
        va = OCTA_BASE + rng.normal(size=(6, 3)) * 0.01 + rng.normal(size=3) * 0.2
        vi = rng.normal(size=3) * 0.45
        ri, rj = rng.uniform(0.02, 0.2, 2)
        j = int(rng.integers(0, len(OCTA_FACES)))
        o = np.zeros(7)
        hit = REF.ref_detection_vertex_face_core(p(vi), C.c_double(ri), 6, p(va), C.c_double(1.0), j, len(OCTA_FACES),
                                                 stream.ctypes.data_as(I32), len(stream), C.c_double(rj), p(o))
        flt = REF.ref_filter_vertex_face(C.c_double(0.1), p(vi), C.c_double(ri), 6, p(va), C.c_double(1.0), j, len(OCTA_FACES),
                                         stream.ctypes.data_as(I32), len(stream), C.c_double(rj))
        inp[k] = np.concatenate([vi, va.ravel(), [ri, rj, j]]); out[k] = np.concatenate([[hit, flt], o])
    G["vf_in"], G["vf_out"] = inp, out
    # OBB: 15-axis SAT + compute_obb (type/OBB.hpp:90-222, basic_types.hpp:37-49), get_vertex (shape.hpp:259-262)
    n = 800
    inp = np.zeros((n, 15 + 15 + 4 + 3 + 1)); out = np.zeros((n, 1 + 15 + 3))
    for k in range(n):
        def rand_obb(scale):
            w, x, y, z = rand_quat(rng)
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            return np.concatenate([rng.normal(size=3) * scale, R[:, 0], R[:, 1], R[:, 2], rng.uniform(0.1, 0.5, 3)])
        A, B = rand_obb(0.5), rand_obb(0.5)
        q, pos, h = rand_quat(rng), rng.normal(size=3), float(rng.choice([1.0, 0.7, 1.3]))
        hit = REF.ref_obb_intersect(p(A), p(B))
        oa, gv = np.zeros(15), np.zeros(3)
        REF.ref_compute_obb(p(A), p(pos), p(q), C.c_double(h), C.c_double(0.05), p(oa))
        REF.ref_get_vertex(p(B[:3]), p(pos), C.c_double(h), p(q), p(gv))
        inp[k] = np.concatenate([A, B, q, pos, [h]]); out[k] = np.concatenate([[hit], oa, gv])
    G["obb_in"], G["obb_out"] = inp, out
    # contact_force_core for the three laws (contact_force.hpp:146-235, common_kernels.hpp:27-145)
    n = 600
    inp = np.zeros((3, n, 8 + 3 + 1 + 21 + 6 + 2)); out = np.zeros((3, n, 9))
    for law in range(3):
        for k in range(n):
            prm = np.array([rng.uniform(1e3, 1e5), rng.uniform(1e3, 1e5), rng.uniform(0, 1), rng.uniform(0, 1),
                            rng.uniform(0, 1), rng.uniform(0, 10), rng.uniform(0.01, 0.2), rng.uniform(0, 5)])
            nn = rng.normal(size=3); nn /= np.linalg.norm(nn)
            dn = float(rng.uniform(-0.05, 0.02 if law == 1 else 0.0))
            vecs = rng.normal(size=(7, 3))   # pi pj vi vj wi wj cp
            if k % 5 == 0:
                vecs[4] = 0.0  # quirk Q1
            ft0, mom0 = rng.normal(size=(2, 3)) * (0.0 if k % 3 == 0 else 5.0)
            meff, reff = float(rng.uniform(0.1, 2)), float(rng.uniform(0.1, 1))
            pi, pj, vi, vj, wi, wj, cp = vecs
            ft, mom, f = ft0.copy(), mom0.copy(), np.zeros(3)
            REF.ref_contact_force_core(law, C.c_double(dn), p(nn), C.c_double(1e-4), p(prm), C.c_double(meff), C.c_double(reff),
                                       p(ft), p(cp), p(pi), p(vi), p(f), p(mom), p(wi), p(pj), p(vj), p(wj))
            inp[law, k] = np.concatenate([prm, nn, [dn], vecs.ravel(), ft0, mom0, [meff, reff]])
            out[law, k] = np.concatenate([ft, mom, f])
    G["law_in"], G["law_out"] = inp, out
    # drivers: cylinder / surface / ball filter + detector (cylinder.hpp:225-279, surface.hpp:320-350, ball.hpp:292-327)
    n = 600
    inp = np.zeros((3, n, 7 + 3 + 1)); out = np.zeros((3, n, 9))
    for kind in range(3):
        for k in range(n):
            if kind == 0:
                axis = np.array([1.0, 0.0, 1.0])
                prm = np.concatenate([[rng.uniform(2, 5)], axis, axis * rng.normal(size=3)])
            elif kind == 1:
                nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
                prm = np.concatenate([[rng.normal()], nrm, [0, 0, 0]])
            else:
                prm = np.concatenate([[rng.uniform(2, 5)], rng.normal(size=3), [0, 0, 0]])
            pt = rng.normal(size=3) * 3
            rcut = float(rng.uniform(0.1, 0.8))
            o = np.zeros(7)
            hit = REF.ref_driver_detector(kind, p(prm), C.c_double(rcut), p(pt), p(o))
            flt = REF.ref_driver_filter(kind, p(prm), C.c_double(rcut), p(pt))
            inp[kind, k] = np.concatenate([prm, pt, [rcut]]); out[kind, k] = np.concatenate([[hit, flt], o])
    G["drv_in"], G["drv_out"] = inp, out
    dst = Path(__file__).resolve().parent / "refmath_golden.npz"
    np.savez_compressed(dst, **G)
    print("wrote", dst, {k: v.shape for k, v in G.items()})


if __name__ == "__main__":
    main()
