"""Generates tests/golden/driver_motion_golden.npz: outputs of the REFERENCE's own driver integrators
(Cylinder/Surface/Ball/RShapeDriver::push_f_v_r, push_f_v, push_av_to_quat, force_to_accel, f_ra — compiled in place from
/root/reference by oracle/Makefile into oracle/_ref/libexadem_refmath.so) on the seeded cases of tests/driver_cases.py.
Run here (needs /root/reference); the vectors travel to the GPU box, the reference does not."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import driver_cases  # noqa: E402
from oracle_binding import ref_math  # noqa: E402

ref = ref_math()
assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
cases = driver_cases.all_cases()
out = driver_cases.run_cases(ref.ref_driver_motion_op, cases)
np.savez_compressed(Path(__file__).parent / "driver_motion_golden.npz", out=out, n=len(cases))
print("wrote", len(cases), "cases")
