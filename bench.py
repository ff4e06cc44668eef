#!/usr/bin/env python
"""bench.py — measures the DEM contact hot path on N B200s of one node (driver contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload drum|box|box8m|poly]

A "step" is one full Verlet time step of the hot path (prolog -> rebuild-or-ghost-refresh -> contact waves ->
deterministic gather fused with gravity/force_to_accel/epilog) over the synthetic workload:
  N = 1 : BASELINE.json configs[1] — 1 M monodisperse spheres in a rotating drum (cylinder driver), fp64.
  N > 1 : weak scaling of the same configuration: the drum is N times longer and cut into N slabs along its axis, one 1 M-sphere
          slab per GPU, ghost exchange between neighbouring slabs (`--workload box`: S5 topology, periodic box of bricks).
metric = interaction-updates/s (sum over waves of records processed by the contact kernels per wall second, whole job);
particle-steps/s is reported beside it.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "interaction-updates/s"
UNIT = "interaction-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="drum (default; N>1: slabs) | box | box8m | poly | mixed")
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-poly", action="store_true", help="skip the polyhedra line reported alongside (N=1 only)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks sampling
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.p = None
        self.device = device

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self, window=None):
        """window = (t0, t1) wall-clock seconds of the timed region: only the samples taken inside it count (nvidia-smi needs a
        few hundred ms to start, so the sampler is launched well before the region and left running through it)."""
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if window is not None:
                ts = self._epoch(f[0])
                if ts is None or ts < window[0] - 0.005 or ts > window[1] + 0.005:
                    continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def one_shot(self):
        """A single synchronous sample (used when the timed region was shorter than the sampling period)."""
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                                 capture_output=True, text=True, timeout=10).stdout
        except Exception:
            return None
        f = [x.strip() for x in out.strip().split(",")]
        try:
            return float(f[1]), float(f[2])
        except Exception:
            return None


# ------------------------------------------------------------------------------------------------ workloads
def make_workload(name, n, rank=0, nranks=1):
    from exadem_b200 import workloads
    if name == "drum":
        if nranks > 1:
            from exadem_b200 import parallel
            return parallel.drum_slab_workload(n, rank, nranks)
        return workloads.rotating_drum(n)
    if name in ("box", "box8m"):
        side = int(round((8_000_000 if name == "box8m" else n) ** (1.0 / 3.0)))
        side = max(8, int(round(side / 5.0)) * 5)  # multiple of 2 cells of 2.5
        from exadem_b200 import parallel
        return parallel.brick_workload(side, rank, nranks)
    if name == "poly":
        side = int(round(n ** (1.0 / 3.0)))
        return workloads.polyhedra_box((side, side, side), pitch=0.78)
    if name == "mixed":
        # BASELINE configs[3]: spheres (one-vertex shapes) + octahedra + cubes in one polyhedron context, cohesive law, closed
        # STL-like mesh walls (an RShape driver) around the packing
        from exadem_b200 import capi
        side = int(round(n ** (1.0 / 3.0)))
        wl = workloads.polyhedra_box((side, side, side), shapes=("octahedron", "cube", "sphere"), pitch=0.78, name="mixed")
        L = side * 0.78
        wl.lo, wl.hi = (-2.0, -2.0, -2.0), (L + 2.0, L + 2.0, L + 2.0)
        wl.periodic = (0, 0, 0)
        wl.law = capi.LAW_COHESIVE
        wl.params = dict(wl.params, fc=0.05, dncut=0.02)
        wl.params_driver = dict(wl.params_driver, fc=0.02, dncut=0.01)
        spec, centre = workloads.box_mesh((-0.45, -0.45, -0.45), (L - 0.33, L - 0.33, L - 0.33), 0.03)
        wl.mesh_spec = [spec]
        wl.drivers = list(wl.drivers) + [dict(kind=capi.DRIVER_RSHAPE, shape=0, center=centre)]
        return wl
    raise SystemExit(f"unknown workload {name}")


def pinned_like(pa):
    """Copy of a ParticleArrays whose dynamic fields live in pinned host memory (torch allocator)."""
    import torch
    out = pa.copy()
    keep = []
    for f in ("rx", "ry", "rz", "vx", "vy", "vz", "vrotx", "vroty", "vrotz", "qw", "qx", "qy", "qz", "fx", "fy", "fz",
              "momx", "momy", "momz", "arotx", "aroty", "arotz"):
        t = torch.empty(pa.n, dtype=torch.float64, pin_memory=torch.cuda.is_available())
        a = t.numpy()
        a[:] = getattr(pa, f)
        setattr(out, f, a)
        keep.append(t)
    out._pinned = keep
    return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_run(wl, seconds, threads=None, max_steps=10_000):
    """Times the oracle (the CPU restatement of the reference path, OpenMP over all host cores) on `wl`.
    Returns (interaction-updates/s, particle-steps/s, steps, wall seconds, threads)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_binding import OracleEngine, OracleLibrary
    lib = OracleLibrary()
    nthr = threads or (os.cpu_count() or 1)
    e = wl.setup(OracleEngine(lib, wl.config()))
    e.set_threads(nthr)
    e.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, 2, 0)
    e.reset_counters()
    t0 = time.perf_counter()
    steps = 0
    chunk = 5
    while steps < max_steps:
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, chunk, 0)
        steps += chunk
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    c = e.counters()
    return c.interaction_updates / dt, c.particle_steps / dt, steps, dt, nthr


def reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path.  The real exaDEM cannot be built here
    (needs exaNBody + Onika + MPI + yaml-cpp, none present) so this is the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or "drum"
    n_sample = min(args.particles, 200_000)
    wl = make_workload(name if name != "box8m" else "box", n_sample)
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_binding import OracleEngine, OracleLibrary
    lib = OracleLibrary()
    nthr = os.cpu_count() or 1
    e = wl.setup(OracleEngine(lib, wl.config()))
    e.set_threads(nthr)
    e.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, max(args.warmup, 1), 0)
    e.reset_counters()
    t0 = time.perf_counter()
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, args.steps, 0)
    dt = time.perf_counter() - t0
    c = e.counters()
    val = c.interaction_updates / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "particle_steps_per_s": c.particle_steps / dt,
            "config": {"workload": f"{name}: bounded sample of {wl.particles.n} spheres of the {args.particles}-particle-per-GPU workload",
                       "particles": int(wl.particles.n), "rebuild": "displacement trigger rcut_inc/2"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port",
                             "sample": f"{wl.particles.n} particles x {args.steps} steps, oracle OpenMP"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def polyhedra_line(lib, peak, device, n=500_000, steps=60, warmup=10):
    """BASELINE configs[2] reported alongside (north_star): ~500 k octahedra, the vertex/edge/face contact path."""
    from exadem_b200 import Engine
    wl = make_workload("poly", n)
    eng = wl.setup(Engine(lib, wl.config(device=device)))
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, warmup, 0)
    eng.sync()
    eng.reset_counters()
    eng.set_timing(True)
    eng.event_record(0)
    nreb = eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, steps, 0)
    eng.event_record(1)
    ms = eng.event_elapsed_ms(0, 1)
    c = eng.counters()
    counts, active = eng.interaction_counts()
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    eng.sync()
    eng.reset_counters()
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    eng.sync()
    reb = eng.counters().rebuild_ms
    eng.set_timing(False)
    t = c.contact_pp_ms * 1e-3
    achieved = c.bytes_contact_algorithmic / t / 1e9 if t > 0 else 0.0
    eng.close()
    return {"workload": f"{wl.particles.n} octahedra, periodic packing (configs[2])", "value": c.interaction_updates / (ms * 1e-3),
            "unit": UNIT, "particle_steps_per_s": c.particle_steps / (ms * 1e-3), "steps": steps, "ms_per_step": ms / steps,
            "interactions": int(counts.sum()), "active_interactions": int(active.sum()), "rebuilds_in_timed_region": int(nreb),
            "rebuild_once_ms": reb,
            "roofline": {"kernel": "k_detect_poly_pp + k_law_poly_pp (4 waves)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "note": "algorithmic bytes counted as for spheres (184 B/record + 92 B/particle per wave); vertex reads not counted"},
            "per_step_ms": {"contact_all": c.contact_ms / steps, "gather_fused": c.gather_ms / steps, "ghost": c.ghost_ms / steps}}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    from exadem_b200 import Engine, Library, capi, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local)   # started now: nvidia-smi needs a few hundred ms before its first sample
    if rank == 0:
        sampler.start()
    name = args.workload or "drum"
    wl = make_workload(name, args.particles, rank, world)
    lib = Library()
    cfg = wl.config(device=local) if world == 1 else parallel.brick_config(wl, rank, world, device=local)
    eng = Engine(lib, cfg)
    if world > 1:
        parallel.init_comm(eng, lib, rank)
    wl.setup(eng)
    g = wl.gravity

    def barrier():
        eng.sync()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        eng.sync()

    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    eng.run_verlet(wl.dt, g, wl.rcut_inc, wl.symetric, max(args.warmup, 3), 0)
    barrier()
    eng.reset_counters()
    eng.set_timing(True)
    barrier()
    eng.event_record(0)
    w0 = time.time()
    t0 = time.perf_counter()
    nreb = eng.run_verlet(wl.dt, g, wl.rcut_inc, wl.symetric, args.steps, 0)
    eng.event_record(1)
    ms = eng.event_elapsed_ms(0, 1)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop((w0, time.time())) if rank == 0 else None
    if rank == 0 and not clocks.get("samples"):
        shot = sampler.one_shot()
        if shot:
            clocks.update({"sm_mhz": shot[0], "sm_max_mhz": shot[1], "samples": 1, "note": "timed region shorter than the sampling period: one sample right after it"})
    c = eng.counters()
    eng.set_timing(False)
    counts, active = eng.interaction_counts()
    stats = np.array([ms, float(c.interaction_updates), float(c.particle_steps), float(c.kernel_launches), float(nreb)], dtype=np.float64)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor(stats, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        iu, ps, launches = float(t[1]), float(t[2]), float(t[3])
    else:
        iu, ps, launches = stats[1], stats[2], stats[3]
    value = iu / (ms * 1e-3)

    # ---- e2e: the host owns the grid; every step ships the dynamic fields in and the results out (pinned memory)
    e2e = None
    if not args.no_e2e:
        host = pinned_like(eng.download_particles())
        h2d = 13 * 8 * host.n
        d2h = 22 * 8 * host.n
        ke = max(5, min(args.steps, 40))
        eng.reset_counters()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            eng.upload_dynamic(host)
            nr = eng.run_verlet(wl.dt, g, wl.rcut_inc, wl.symetric, 1, 0)
            if nr and world > 1:   # a rebuild may have migrated particles between bricks: the host mirror follows
                host = pinned_like(eng.download_particles())
            else:
                eng.download_dynamic(host)
        barrier()
        dt = time.perf_counter() - t0
        ce = eng.counters()
        iu_e = float(ce.interaction_updates)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([iu_e, dt], device="cuda", dtype=torch.float64)
            tm = t.clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            iu_e, dt = float(t[0]), float(tm[1])
        e2e = {"value": iu_e / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": ke,
               "ms_per_step": 1e3 * dt / ke}

    # ---- cost of one neighbour rebuild (outside the timed region; the drum moves too slowly to trigger one in K steps)
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)   # (the first explicit rebuild allocates the second set of wave buffers)
    barrier()
    eng.reset_counters()
    eng.set_timing(True)
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    barrier()
    rebuild_once_ms = eng.counters().rebuild_ms
    eng.set_timing(False)

    if rank != 0:
        return
    peak, peak_src = peaks()
    # roofline of the dominant kernel (wave 0 contact kernel): algorithmic bytes per launch / mean launch duration
    rec_per_launch = c.contact_pp_records / max(c.contact_pp_launches, 1)
    alg_bytes = rec_per_launch * 184.0 + wl.particles.n * 92.0
    t_launch = (c.contact_pp_ms / max(c.contact_pp_launches, 1)) * 1e-3
    achieved = alg_bytes / t_launch / 1e9 if t_launch > 0 else 0.0
    roof = {"kernel": "k_contact_sphere_pp" if wl.mode == 0 else "k_detect_poly_pp + k_law_poly_pp", "bound": "hbm", "achieved": achieved,
            "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": t_launch * 1e3,
            "share_of_step": c.contact_pp_ms / ms if ms > 0 else None,
            "per_step_ms": {"contact_all": c.contact_ms / args.steps, "gather_fused": c.gather_ms / args.steps,
                            "prolog": c.prolog_ms / args.steps, "ghost": c.ghost_ms / args.steps,
                            "rebuild_amortised": c.rebuild_ms / args.steps}}
    prof = ROOT / "profiles" / "ncu_traffic.json"
    if prof.exists():
        try:
            roof["traffic"] = json.loads(prof.read_text()).get(roof["kernel"])
        except Exception:
            pass
    roof["rebuild_once_ms"] = rebuild_once_ms
    # The kernel's OWN compulsory traffic: 9 B per record, + 128 B per record that needs the law (friction r/w, wrench),
    # + 96 B per particle (three gathered sectors).  `achieved` above uses the reference algorithm's 184 B/record, which
    # this kernel undercuts for records that are not in contact — so `frac` can exceed what the DRAM actually delivers.
    n_act = float(active[0]) if wl.mode == 0 else float(active[:4].sum())
    own_bytes = 9.0 * rec_per_launch + 128.0 * n_act + 96.0 * wl.particles.n
    roof["own_bytes_per_launch"] = own_bytes
    roof["frac_dram_model"] = (own_bytes / t_launch / 1e9 / peak) if t_launch > 0 else None
    if roof["frac"] > 1.0:
        roof["note"] = ("frac > 1: the reference algorithm's compulsory bytes (184 B/record) divided by this kernel's time exceed the peak "
                        "because records not in contact cost this kernel 9 B; see frac_dram_model / traffic for the bytes it really moves")
    # The other half of the step: k_gather<2> (deterministic force sum + gravity + force_to_accel + epilog + next prolog + trigger).
    # Algorithmic bytes: 312 B per particle (r, v, vrot, Q, inertia, r_backup in; r, v, vrot, Q out) + 5 B per incidence
    # (code + flag byte; a particle-particle record is incident to two particles) + 64 B per incident side of a record in contact.
    roof_g = None
    if c.gather_launches > 0 and c.gather_ms > 0:
        npp = 1 if wl.mode == 0 else 4
        inc = 2.0 * float(counts[:npp].sum()) + float(counts[npp if npp == 4 else 4:].sum())
        hot = 2.0 * float(active[:npp].sum()) + float(active[4:].sum())
        g_bytes = 312.0 * wl.particles.n + 5.0 * inc + 64.0 * hot
        tg = c.gather_ms / c.gather_launches * 1e-3
        roof_g = {"kernel": "k_gather", "bound": "hbm", "achieved": g_bytes / tg / 1e9, "peak": peak, "unit": "GB/s",
                  "frac": g_bytes / tg / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": g_bytes,
                  "mean_launch_ms": tg * 1e3, "share_of_step": c.gather_ms / ms if ms > 0 else None}
        if prof.exists():
            try:
                roof_g["traffic"] = json.loads(prof.read_text()).get("k_gather")
            except Exception:
                pass
    poly = None
    if world == 1 and not args.no_poly and (args.workload or "drum") == "drum":
        eng.close()
        poly = polyhedra_line(lib, peak, local)
    cpu = None
    if not args.no_cpu_baseline:
        n_sample = min(args.particles, 200_000)
        wls = make_workload(name if name != "box8m" else "box", n_sample)
        v, pss, st, dt, nthr = cpu_run(wls, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": nthr, "kind": "port", "particle_steps_per_s": pss,
               "sample": f"{wls.particles.n}-particle sample of the workload, {st} steps in {dt:.1f} s, oracle OpenMP ({nthr} threads)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "particle_steps_per_s": ps / (ms * 1e-3),
            "config": {"workload": f"{name}: {wl.particles.n * world} particles ({wl.particles.n}/GPU), {wl.name}", "particles_per_gpu": int(wl.particles.n),
                       "interactions_per_gpu": int(counts.sum()), "active_interactions_per_gpu": int(active.sum()),
                       "rebuilds_in_timed_region": int(nreb), "rebuild": "displacement trigger rcut_inc/2",
                       "l2": "working set (particles + interaction waves) > 126 MB L2, no flush",
                       "timing": "CUDA events on the library stream, max over ranks", "wall_s": wall},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    if roof_g is not None:
        line["roofline_gather"] = roof_g
    if poly is not None:
        line["polyhedra"] = poly
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass
