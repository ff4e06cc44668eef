"""Driver operators of the time loop: push_f_v_r_driver, push_f_v_driver, push_av_to_quat_driver,
force_to_accel_driver, driver_vertices, grid_rshape_driver, driver_displ_over.

CPU: the oracle's restatement is pinned bit-for-bit against the reference's own Cylinder / Surface / Ball / RShapeDriver member
functions (oracle/_ref, compiled in place) and against the committed golden vectors made from them.
GPU: the device-resident drivers go through the same operators and through edb_run_verlet with moving drivers, against the oracle."""
from pathlib import Path

import numpy as np
import pytest

import driver_cases
from oracle_binding import OracleLibrary, ref_math

GOLD = Path(__file__).parent / "golden" / "driver_motion_golden.npz"


def _oracle_rows(cases):
    lib = OracleLibrary()
    return driver_cases.run_cases(lib.dll.orc_driver_motion_op, cases)


def _assert_bitwise(a, b, cases):
    assert a.shape == b.shape
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        bad = np.argwhere(~same)[0]
        kind, motion, op = cases[bad[0]][:3]
        raise AssertionError(f"case {bad[0]} (kind {kind}, {motion}, op {op}) column {bad[1]}: {a[bad[0], bad[1]]!r} != {b[bad[0], bad[1]]!r}")


def test_oracle_driver_motion_vs_golden():
    cases = driver_cases.all_cases()
    g = np.load(GOLD)
    assert int(g["n"]) == len(cases)
    _assert_bitwise(_oracle_rows(cases), g["out"], cases)


def test_oracle_driver_motion_vs_reference():
    ref = ref_math()
    if ref is None or not hasattr(ref, "ref_driver_motion_op"):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    cases = driver_cases.all_cases(seed=77, reps=10)
    _assert_bitwise(_oracle_rows(cases), driver_cases.run_cases(ref.ref_driver_motion_op, cases), cases)
