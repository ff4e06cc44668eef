"""Simulation-level golden vectors (tests/golden/sim_golden.npz, made by tests/golden/make_sim_golden.py with the oracle in its
serial canonical order): five small runs — rotating drum (BASELINE configs[0]/[1] family), deposit on a surface driver
(configs[0]), octahedra + cubes (configs[2] family), and two runs with moving drivers (a descending roof; a translating and
spinning mesh around a free-body mesh).  The reference ships no expected values for the operator loops (SURVEY.md
8c), so these pin the oracle against ITSELF over time (CPU test) and the CUDA path against the committed numbers (GPU test):
interaction keys bit-exact, state and friction history within 1e-10."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
from make_sim_golden import cases, run  # noqa: E402

G = np.load(Path(__file__).resolve().parent / "golden" / "sim_golden.npz")


def compare(name, got, rtol):
    nb, state, keys, fr = got
    assert nb == int(G[f"{name}_nb"][0])
    assert np.array_equal(keys, G[f"{name}_keys"]), f"{name}: interaction set differs from the golden set"
    ref = G[f"{name}_state"]
    assert state.shape == ref.shape
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-300)
    err = (np.abs(state - ref) / scale).max()
    assert err <= rtol, f"{name}: state differs from the golden vectors by {err:.3e}"
    fs = max(np.abs(G[f"{name}_friction"]).max(), 1e-300)
    assert np.abs(fr - G[f"{name}_friction"]).max() <= rtol * fs


@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_oracle_reproduces_its_golden_runs(oracle_lib, case):
    from oracle_binding import OracleEngine
    name, wl, steps, reb = case
    compare(name, run(OracleEngine, oracle_lib, wl, steps, reb), 1e-13)


# the two moving-driver runs have their GPU form in tests/test_survey8f_rank4_run_verlet.py
GPU_VERIFIED = ("drum", "deposit", "poly")


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in cases() if c[0] in GPU_VERIFIED], ids=lambda c: c[0])
def test_cuda_path_matches_golden_runs(product_lib, case):
    from exadem_b200 import Engine
    name, wl, steps, reb = case
    compare(name, run(Engine, product_lib, wl, steps, reb), 1e-10)
