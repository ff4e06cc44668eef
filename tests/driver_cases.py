"""Seeded single-driver cases for the driver-motion operators (push_f_v_r / push_f_v / push_av_to_quat / force_to_accel):
shared by tests/golden/make_driver_golden.py (reference -> golden vectors) and tests/test_survey8f_rank3_driver_motion.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from exadem_b200 import capi

# motion types each driver kind accepts (cylinder.hpp:95-96, surface.hpp:146-147, ball.hpp:142-146, rshape.hpp:148-150);
# EXPRESSION needs tinyexpr on the reference side and is covered by the oracle-vs-GPU test only
VALID = {
    0: ("STATIONARY", "SHAKER"),
    1: ("STATIONARY", "LINEAR_MOTION", "LINEAR_COMPRESSIVE_MOTION", "SHAKER", "PENDULUM_MOTION"),
    2: ("STATIONARY", "LINEAR_MOTION", "COMPRESSIVE_FORCE", "TABULATED"),
    3: ("STATIONARY", "LINEAR_MOTION", "LINEAR_FORCE_MOTION", "LINEAR_COMPRESSIVE_MOTION", "PARTICLE", "TABULATED", "SHAKER"),
}
TAB_TIME = np.array([0.0, 0.5, 1.25, 2.0, 3.5])
TAB_POS = np.array([[0, 0, 0], [0.5, 0.1, 0], [1.0, -0.2, 0.3], [1.0, 0.4, 0.9], [2.0, 0.0, 1.0]], dtype=np.float64)


def unit(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def make_case(rng, kind, motion):
    """-> (driver dict, motion dict, time, dt, system_mass)"""
    n = unit(rng)
    d = {"kind": kind, "radius": float(rng.uniform(0.5, 3.0)), "axis": (1, 0, 1), "center": rng.normal(size=3),
         "offset": float(rng.normal()), "normal": n, "vel": rng.normal(size=3) * 0.3, "vrot": rng.normal(size=3) * 0.2, "shape": 0}
    q = rng.normal(size=4)
    d["quat"] = q / np.linalg.norm(q)
    mv = unit(rng)
    if kind == 1 and motion in ("LINEAR_MOTION",):
        mv = n   # surface.hpp:176-186: motion vector aligned with the normal
    m = {"motion_type": motion, "motion_vector": mv, "const_vel": float(rng.uniform(0.1, 2.0)), "const_force": float(rng.normal()),
         "sigma": float(rng.choice([0.0, rng.uniform(0.1, 5.0)])), "damprate": float(rng.uniform(0.0, 1.0)),
         "omega": float(rng.uniform(1.0, 50.0)), "amplitude": float(rng.uniform(0.01, 0.5)), "shaker_dir": unit(rng),
         "motion_start_threshold": float(rng.choice([0.0, 0.3])), "motion_end_threshold": float(rng.choice([1e300, 2.0])),
         "mass": float(rng.uniform(0.5, 50.0)), "surface": float(rng.uniform(1.0, 20.0)), "acc": rng.normal(size=3),
         "rv": float(rng.normal() * 0.1), "ra": float(rng.normal() * 0.1), "inertia": rng.uniform(0.5, 2.0, size=3),
         "forces": rng.normal(size=3) * 10, "mom": rng.normal(size=3), "drive_by_mom": int(rng.integers(0, 2)),
         "tab_time": TAB_TIME, "tab_pos": TAB_POS}
    am = rng.normal(size=3)
    m["applied_mom"] = am
    m["mom_axis"] = am / np.linalg.norm(am)
    if motion == "LINEAR_MOTION":
        d["vel"] = m["const_vel"] * np.asarray(mv)
    if motion == "PENDULUM_MOTION":
        anchor = rng.normal(size=3)
        swing = unit(rng)
        ax = np.cross(swing, unit(rng))
        ax /= np.linalg.norm(ax)
        m.update(pendulum_anchor_point=anchor, pendulum_initial_position=anchor - ax * float(rng.uniform(1.0, 3.0)), pendulum_swing_dir=swing)
        d["center"] = anchor
    if kind == 1:
        m["mom"] = d["offset"] * np.asarray(n)   # center_proj_ travels in the mom slot of the single-driver test interface
    time = float(rng.uniform(0.0, 3.0))
    dt = float(rng.choice([1e-4, 1e-3, 0.05]))
    return d, m, time, dt, float(rng.uniform(10.0, 1000.0))


def to_structs(d, m, keep):
    D = capi.Driver()
    D.kind = d["kind"]
    D.radius = d["radius"]
    D.offset = d["offset"]
    D.shape = d["shape"]
    for k in ("axis", "center", "normal", "vel", "vrot"):
        for c in range(3):
            getattr(D, k)[c] = float(d[k][c])
    for c in range(4):
        D.quat[c] = float(d["quat"][c])
    return D, capi.make_motion(m, keep)


OUT_FIELDS_D = ("radius", "center", "offset", "normal", "vel", "vrot", "quat")
OUT_FIELDS_M = ("acc", "rv", "ra", "mom", "forces", "applied_mom", "mom_axis")


def flatten(D, M):
    out = []
    for k in OUT_FIELDS_D:
        v = getattr(D, k)
        out += list(v[:]) if hasattr(v, "__len__") else [v]
    for k in OUT_FIELDS_M:
        v = getattr(M, k)
        out += list(v[:]) if hasattr(v, "__len__") else [v]
    return np.array(out, dtype=np.float64)


def all_cases(seed=2024, reps=6):
    rng = np.random.default_rng(seed)
    cases = []
    for kind, motions in VALID.items():
        for motion in motions:
            for op in (0, 1, 2, 3):
                if op == 2 and kind != 3:
                    continue
                for _ in range(reps):
                    cases.append((kind, motion, op) + make_case(rng, kind, motion))
    return cases


def run_cases(fn, cases):
    """fn(op, time, dt, system_mass, byref(D), byref(M)) -> flattened outputs, one row per case"""
    rows = []
    for kind, motion, op, d, m, time, dt, smass in cases:
        keep = []
        D, M = to_structs(d, m, keep)
        rc = fn(C.c_int(op), C.c_double(time), C.c_double(dt), C.c_double(smass), C.byref(D), C.byref(M))
        assert rc == 0
        rows.append(flatten(D, M))
    return np.array(rows)
