#!/usr/bin/env python
"""Generates tests/golden/sim_golden.npz: end states of five small simulations run by the ORACLE (serial canonical order) —
a regression anchor for the operator-loop layer that the reference's own tests do not pin: any change to the
oracle's loops, to the workload generators or to the CUDA path that moves a result shows up against these vectors.

    python tests/golden/make_sim_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from exadem_b200 import capi, workloads  # noqa: E402
from oracle_binding import OracleEngine, OracleLibrary  # noqa: E402
from parity import sort_interactions  # noqa: E402


def cases():
    """name -> (workload, steps, rebuild_every)"""
    yield "drum", workloads.rotating_drum(900, jitter=0.02), 30, 0
    yield "deposit", workloads.deposit((8, 8, 6)), 30, 10
    yield "poly", workloads.polyhedra_box((6, 6, 6), shapes=("octahedron", "cube"), pitch=0.8, cell_size=1.6), 20, 5
    # moving drivers (scenarios of tests/test_survey8f_rank3_driver_motion.py without trigonometric motions, so that the device's
    # sin / cos do not enter): a roof descending at constant velocity, a translating + spinning mesh around a free-body mesh
    import test_survey8f_rank3_driver_motion as T
    yield "piston", T.scenario_piston(), 60, 0
    yield "moving_mesh", T.scenario_moving_mesh(), 30, 0


def run(engine_cls, lib, wl, steps, rebuild_every):
    e = wl.setup(engine_cls(lib, wl.config()))
    e.update_particle_neighbors(wl.rcut_inc, True)
    nb = e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, steps, rebuild_every)
    pa = e.download_particles()
    rec = sort_interactions(e.export_interactions())
    order = np.argsort(pa.id)
    state = np.stack([getattr(pa, f)[order] for f in capi.F64_FIELDS], axis=1)
    keys = np.stack([rec[f].astype(np.uint64) for f in ("type", "id_i", "id_j", "sub_i", "sub_j", "swap")], axis=1)
    return nb, state, keys, rec["friction"].copy()


if __name__ == "__main__":
    lib = OracleLibrary()
    out = {}
    for name, wl, steps, reb in cases():
        nb, state, keys, fr = run(OracleEngine, lib, wl, steps, reb)
        out[f"{name}_nb"] = np.array([nb])
        out[f"{name}_state"] = state
        out[f"{name}_keys"] = keys
        out[f"{name}_friction"] = fr
        print(name, "particles", state.shape[0], "records", len(keys), "rebuilds", nb)
    np.savez_compressed(Path(__file__).resolve().parent / "sim_golden.npz", **out)
