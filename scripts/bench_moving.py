#!/usr/bin/env python
"""Cost of a Verlet step when drivers MOVE: edb_run_verlet then runs the macro operator by operator with the
device-resident driver integrators (csrc/kernels_driver.cuh) instead of the fused speculative loop.  Prints one JSON line per case
(1 M-sphere drum = BASELINE configs[1] with a shaken cylinder; 4 M... not here).  Not the round's bench line: an explanatory number."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from exadem_b200 import Engine, Library, capi, workloads  # noqa: E402


def timed(e, wl, steps):
    e.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, 20, 0)
    e.reset_counters()
    e.sync()
    e.event_record(0)
    nb = e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, steps, 0)
    e.event_record(1)
    ms = e.event_elapsed_ms(0, 1)
    c = e.counters()
    return ms / steps, nb, c.interaction_updates / (ms * 1e-3), c.kernel_launches


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    lib = Library()
    out = []
    for case in ("stationary", "shaker_cylinder", "piston_roof"):
        if case == "piston_roof":
            side = int(round((n / 40) ** 0.5))
            wl = workloads.deposit((side, side, 40))
            top = 40 + 0.55
            wl.drivers.append(dict(kind=capi.DRIVER_SURFACE, offset=top, normal=(0, 0, 1), center=(0, 0, top), vel=(0, 0, -1.0)))
            wl.motions = [dict(motion_type="STATIONARY"), dict(motion_type="LINEAR_MOTION", motion_vector=(0, 0, 1), const_vel=-1.0)]
        else:
            wl = workloads.rotating_drum(n)
            if case == "shaker_cylinder":
                wl.motions = [dict(motion_type="SHAKER", omega=50.0, amplitude=0.01, shaker_dir=(0, 0, 1))] + \
                             [dict(motion_type="STATIONARY")] * (len(wl.drivers) - 1)
        e = wl.setup(Engine(lib, wl.config()))
        t0 = time.perf_counter()
        ms, nb, ius, launches = timed(e, wl, steps)
        out.append({"case": case, "particles": int(wl.particles.n), "steps": steps, "ms_per_step": ms, "rebuilds": int(nb),
                    "interaction_updates_per_s": ius, "kernel_launches": int(launches), "wall_s": time.perf_counter() - t0})
        print(json.dumps(out[-1]), flush=True)
        e.close()


if __name__ == "__main__":
    main()
