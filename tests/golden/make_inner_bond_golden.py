"""Generates tests/golden/inner_bond_golden.npz: outputs of the REFERENCE's own inner-bond law (force_law_core,
src/forcefield/include/exaDEM/forcefield/inner_bond_force.hpp:26-70, compiled in place into oracle/_ref/libexadem_refmath.so) on the
seeded cases of tests/test_survey8f_rank4_inner_bonds.py:law_cases().  Run here (needs /root/reference)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import test_survey8f_rank4_inner_bonds as T  # noqa: E402
from oracle_binding import ref_math  # noqa: E402

ref = ref_math()
assert ref is not None and hasattr(ref, "ref_inner_bond_force_core")
out = T.run_law(ref.ref_inner_bond_force_core, T.law_cases())
np.savez_compressed(Path(__file__).parent / "inner_bond_golden.npz", out=out)
print("wrote", len(out), "cases")
