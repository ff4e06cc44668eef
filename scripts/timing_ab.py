import sys, time
sys.path.insert(0,'/root/repo')
from exadem_b200 import Engine, Library, workloads
lib=Library()
wl=workloads.rotating_drum(1_000_000)
for timing in (False, True, False, True):
    e=wl.setup(Engine(lib, wl.config()))
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 20, 0)
    e.sync(); e.set_timing(timing)
    e.event_record(0)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 200, 0)
    e.event_record(1)
    ms=e.event_elapsed_ms(0,1)
    print("timing", timing, "ms/step", ms/200)
    e.close()
