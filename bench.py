#!/usr/bin/env python
"""bench.py — measures the DEM contact hot path on N B200s of one node (driver contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload drum|box|poly|mixed]
                    [--scaling weak|strong] [--particles P] [--global-side S]

A "step" is one full Verlet time step of the hot path (prolog -> rebuild-or-ghost-refresh -> contact waves ->
deterministic gather fused with gravity/force_to_accel/epilog) over the synthetic workload:
  N = 1 : BASELINE.json configs[1] — 1 M monodisperse spheres in a rotating drum (cylinder driver), fp64.
  N > 1 : weak scaling of THE SAME configuration (the contract: per-GPU work fixed as N grows): the drum N times longer, cut into N slabs
          along its periodic axis, one 1 M-sphere slab per GPU.  The line also carries `box_bricks`: the BASELINE configs[4] topology —
          a fully periodic box cut into 2x1x1 / 2x2x1 / 2x2x2 bricks of 1 M spheres (7 peers per rank at N = 8) — measured in the same
          job together with one such brick alone on one GPU (`single_gpu_same_brick`, the denominator of the weak-scaling efficiency
          on that topology).  `--workload box` makes the bricks the headline; `--scaling strong --global-side 200|400` divides a FIXED
          8 M / 64 M box by N.
metric = interaction-updates/s (sum over waves of records processed by the contact kernels per wall second, whole job);
particle-steps/s is reported beside it.  One JSON line is printed by rank 0.
`--impl reference`: the CPU arm (the oracle port of the reference operators, OpenMP on all host cores) on the SAME workload.
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
L2_NOTE = "inputs larger than L2: the working set of a step (particles + interaction waves, > 0.7 GB per GPU) exceeds the 126 MB L2; no flush"
CPU_ARM_MAX_PARTICLES = 9_000_000     # the CPU arm runs the whole N-GPU workload up to this size, a whole-brick sample beyond


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="drum (default at N=1) | box (default at N>1) | poly | mixed")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU (weak scaling)")
    ap.add_argument("--global-side", type=int, default=200, help="strong scaling: side of the fixed global periodic box (200 = 8 M, 400 = 64 M)")
    ap.add_argument("--jitter", type=float, default=2e-5, help="box workload: uniform position jitter of the SC lattice")
    ap.add_argument("--cell-size", type=float, default=None, help="poly / mixed workloads: grid cell size (default: the workload's own, 2.0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-poly", action="store_true", help="skip the polyhedra line reported alongside (N=1 only)")
    ap.add_argument("--no-mixed", action="store_true", help="skip the 4 M mixed spheres + polyhedra line (configs[3]) reported alongside (N=1 only)")
    ap.add_argument("--no-secondary", action="store_true", help="N>1: skip the box-brick / single-brick lines reported alongside")
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
        # (samples go to a file: a pipe nobody reads fills up after a few thousand lines and stalls nvidia-smi during long set-ups)
        import tempfile
        try:
            self.out = tempfile.NamedTemporaryFile("w+", prefix="edb_clocks_", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                                       "-i", str(self.device)], stdout=self.out, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self, windows=None):
        """windows = [(t0, t1), ...] wall-clock seconds of the timed regions: only the samples taken inside one of them count
        (nvidia-smi needs a few hundred ms to start, so the sampler is launched well before and left running through them)."""
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        try:
            self.out.flush()
            self.out.seek(0)
            out = self.out.read()
            self.out.close()
            os.unlink(self.out.name)
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if windows is not None:
                ts = self._epoch(f[0])
                if ts is None or not any(w[0] - 0.005 <= ts <= w[1] + 0.005 for w in windows):
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


# ------------------------------------------------------------------------------------------------ workloads
def resolve_workload(args, world):
    """-> (name, make(rank, nranks) -> Workload, description shared by both arms, scaling)"""
    from exadem_b200 import parallel, workloads
    name = args.workload or "drum"
    scaling = args.scaling
    if name == "drum":
        if scaling == "strong":
            raise SystemExit("--scaling strong is defined for the box workload")
        def make(rank, nranks):
            return parallel.drum_slab_workload(args.particles, rank, nranks) if nranks > 1 else workloads.rotating_drum(args.particles)
        desc = (f"drum: ~{args.particles} monodisperse spheres per GPU in a rotating cylinder driver (BASELINE configs[1])"
                + (f", {world} slabs along the periodic axis" if world > 1 else ""))
        return name, make, desc, scaling
    if name == "box":
        if scaling == "strong":
            dims = parallel.strong_brick_dims(args.global_side, world)
            desc = (f"box: fixed {args.global_side}^3 = {args.global_side ** 3} spheres, fully periodic (BASELINE configs[4]), strong scaling: "
                    f"{'x'.join(map(str, parallel.process_grid(world)))} bricks of {dims[0]}x{dims[1]}x{dims[2]}")
        else:
            side = max(10, int(round(round(args.particles ** (1.0 / 3.0)) / 5.0)) * 5)   # whole 2.5-wide cells, even count
            dims = (side, side, side)
            desc = (f"box: {side}^3 = {side ** 3} spheres per GPU, fully periodic (BASELINE configs[4] topology), weak scaling: "
                    f"{'x'.join(map(str, parallel.process_grid(world)))} bricks")
        def make(rank, nranks, dims=dims):
            return parallel.brick_workload(dims, rank, nranks, jitter=args.jitter)
        return name, make, desc, scaling
    if world > 1:
        raise SystemExit(f"workload {name} is single-GPU")
    if name == "poly":
        side = int(round(args.particles ** (1.0 / 3.0)))
        kw = {"cell_size": args.cell_size} if args.cell_size else {}
        return name, (lambda rank, nranks: workloads.polyhedra_box((side, side, side), pitch=0.78, **kw)), \
            f"poly: {side ** 3} octahedra, periodic packing (BASELINE configs[2])", scaling
    if name == "mixed":
        return name, (lambda rank, nranks: mixed_workload(args.particles)), \
            f"mixed: ~{args.particles} spheres + octahedra + cubes, cohesive law, closed mesh walls (BASELINE configs[3])", scaling
    raise SystemExit(f"unknown workload {name}")


def mixed_workload(n):
    """BASELINE configs[3]: spheres (one-vertex shapes) + octahedra + cubes in one polyhedron context, cohesive law, closed
    STL-like mesh walls (an RShape driver) around the packing"""
    from exadem_b200 import capi, workloads
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


DYN_FIELDS = ("rx", "ry", "rz", "vx", "vy", "vz", "vrotx", "vroty", "vrotz", "qw", "qx", "qy", "qz", "fx", "fy", "fz",
              "momx", "momy", "momz", "arotx", "aroty", "arotz")


def pinned_like(pa):
    """Copy of a ParticleArrays whose dynamic fields live in pinned host memory (torch allocator)."""
    import torch
    out = pa.copy()
    keep = []
    for f in DYN_FIELDS:
        t = torch.empty(pa.n, dtype=torch.float64, pin_memory=torch.cuda.is_available())
        a = t.numpy()
        a[:] = getattr(pa, f)
        setattr(out, f, a)
        keep.append(t)
    out._pinned = keep
    return out


def link_bandwidth(nbytes=128 << 20, reps=3):
    """pinned host <-> device copy bandwidth of this box (GB/s), one direction at a time: the floor of the e2e leg"""
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for name, (dst, src) in (("h2d_gbs", (d, h)), ("d2h_gbs", (h, d))):
        best = 0.0
        for _ in range(reps + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dst.copy_(src, non_blocking=True)
            b.record()
            b.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        out[name] = best
    # both directions at once (what the e2e loop does: the upload of step n+1 runs beside the download of step n)
    h2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    best = 0.0
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s_up):
            for _k in range(2):
                d2.copy_(h2, non_blocking=True)
        with torch.cuda.stream(s_dn):
            a.record()
            h.copy_(d, non_blocking=True)
            b.record()
        b.synchronize()
        torch.cuda.synchronize()
        best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
    out["d2h_gbs_while_h2d"] = best
    return out


# ------------------------------------------------------------------------------------------------ CPU arm
def host_threads():
    """All host cores, whatever the launcher did to this process: torchrun exports OMP_NUM_THREADS=1 and a launcher may have pinned
    the rank to a few cores — the CPU arm undoes both BEFORE the oracle library (libgomp) is loaded."""
    os.environ.pop("OMP_NUM_THREADS", None)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    os.environ.pop("GOMP_CPU_AFFINITY", None)
    os.environ["OMP_PROC_BIND"] = "false"
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except Exception:
        pass
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_engine(wl, nthr):
    sys.path.insert(0, str(ROOT / "tests"))
    import ctypes
    from oracle_binding import OracleEngine, OracleLibrary
    lib = OracleLibrary()
    e = wl.setup(OracleEngine(lib, wl.config()))
    e.set_threads(nthr)
    got = nthr
    try:
        f = lib.dll.orc_threads_in_parallel
        f.restype = ctypes.c_int
        got = int(f(ctypes.c_int(nthr)))
    except Exception:
        pass
    return e, got


def cpu_warm(e, wl, min_steps, min_seconds=1.0):
    """warm-up: at least `min_steps` steps AND `min_seconds` of stepping (page faults, thread pool, caches)"""
    e.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    t0 = time.perf_counter()
    done = 0
    while done < max(min_steps, 1) or time.perf_counter() - t0 < min_seconds:
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, 2, 0)
        done += 2
    return done


def cpu_baseline_leg(wl, seconds):
    """cpu_baseline of the GPU arm: the oracle on all host cores, the N=1 workload, a bounded number of steps (~`seconds`)."""
    nthr = host_threads()
    e, got = cpu_engine(wl, nthr)
    warm = cpu_warm(e, wl, 3)
    e.reset_counters()
    t0 = time.perf_counter()
    steps = 0
    while True:
        e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, 5, 0)
        steps += 5
        if time.perf_counter() - t0 >= seconds or steps >= 10_000:
            break
    dt = time.perf_counter() - t0
    c = e.counters()
    return {"value": c.interaction_updates / dt, "unit": UNIT, "cores": got, "kind": "port",
            "particle_steps_per_s": c.particle_steps / dt,
            "sample": f"the whole workload ({wl.particles.n} particles), {steps} steps in {dt:.1f} s after {warm} warm-up steps, "
                      f"oracle OpenMP ({got} threads)"}


def reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path.  The real exaDEM cannot be built here
    (needs exaNBody + Onika + MPI + yaml-cpp, none present) so this is the oracle port, all host threads, on the SAME workload as
    the GPU arm at this N (the N bricks / slabs as one domain)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthr = host_threads()
    from exadem_b200 import parallel
    world = args.gpus
    name, make, desc, scaling = resolve_workload(args, world)
    n_run = world
    per = make(0, world).particles.n if world > 1 else None
    sample = "the whole workload"
    if world > 1 and per * world > CPU_ARM_MAX_PARTICLES:
        # bounded sample: as many whole bricks as fit the cap, run as their own periodic domain
        n_run = max(1, [k for k in (8, 4, 2, 1) if k <= world and per * k <= CPU_ARM_MAX_PARTICLES][0]) if per <= CPU_ARM_MAX_PARTICLES else 1
        sample = f"bounded sample: {n_run} of the {world} bricks as one periodic domain"
    wl = parallel.global_workload(make, n_run) if n_run > 1 else make(0, 1)
    e, got = cpu_engine(wl, nthr)
    warm = cpu_warm(e, wl, args.warmup)
    e.reset_counters()
    t0 = time.perf_counter()
    e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, args.steps, 0)
    dt = time.perf_counter() - t0
    c = e.counters()
    val = c.interaction_updates / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "particle_steps_per_s": c.particle_steps / dt,
            "config": {"workload": desc, "rebuild": "displacement trigger rcut_inc/2", "l2": L2_NOTE},
            "run": {"particles": int(wl.particles.n), "warmup_steps_done": warm, "threads_requested": nthr, "threads_in_parallel_region": got,
                    "affinity_cpus": nthr, "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": got, "kind": "port",
                             "sample": f"{sample}: {wl.particles.n} particles x {args.steps} steps after {warm} warm-up steps, oracle OpenMP ({got} threads)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ rooflines
def load_traffic():
    """ncu `--set full` DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), keyed by kernel and operating point;
    written by scripts/ncu_traffic.py from the captures summarised in profiles/ (the capture command is recorded in the file)."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return {}
    try:
        return json.loads(p.read_text())
    except Exception:
        return {}


def traffic_for(tr, kernel, workload, steps, warmup):
    """traffic of `kernel` captured at the operating point nearest to (steps, warmup) of this run; None when no capture of this
    workload exists"""
    best = None
    for pt in tr.get("points", []):
        if pt.get("workload") != workload or kernel not in pt.get("kernels", {}):
            continue
        d = abs(pt.get("step_index", 0) - (warmup + steps // 2))
        if best is None or d < best[0]:
            best = (d, pt)
    if best is None:
        return None, None
    pt = best[1]
    return pt["kernels"][kernel], {"step_index": pt.get("step_index"), "capture": pt.get("capture")}


def sphere_rooflines(c, counts, active, n_real, ms, steps, peak, peak_src, tr, wname, warmup):
    """Both per-step kernels on equal footing.  `achieved` = bytes this kernel has to move for the state that was timed (its own
    compulsory-traffic model) / mean launch duration (CUDA events on the library stream, inside the timed region);
    `frac_algorithmic` keeps the SURVEY 8(d) figure (184 B/record + 92 B/particle of the REFERENCE algorithm, which this kernel
    undercuts for records that are not in contact); `traffic` = ncu DRAM bytes of one launch at the nearest captured operating point."""
    out = {}
    n_pp = max(c.contact_pp_launches, 1)
    rec = c.contact_pp_records / n_pp
    t_pp = (c.contact_pp_ms / n_pp) * 1e-3
    hot0 = float(active[0])
    own = 9.0 * rec + 128.0 * hot0 + 96.0 * n_real      # ij+flg per record; friction r/w + wrench w per hot record; 3 sectors per particle
    alg = rec * 184.0 + n_real * 92.0
    tpp, cap = traffic_for(tr, "k_contact_sphere_pp", wname, steps, warmup)
    out["contact"] = {"kernel": "k_contact_sphere_pp", "bound": "hbm", "achieved": own / t_pp / 1e9 if t_pp > 0 else 0.0, "peak": peak,
                      "peak_source": peak_src, "unit": "GB/s", "frac": own / t_pp / 1e9 / peak if t_pp > 0 else 0.0,
                      "bytes_model": "9 B/record + 128 B/record in contact + 96 B/particle", "bytes_per_launch": own,
                      "traffic": tpp, "traffic_capture": cap, "frac_traffic": (tpp / t_pp / 1e9 / peak) if (tpp and t_pp > 0) else None,
                      "algorithmic_bytes_per_launch": alg, "frac_algorithmic": alg / t_pp / 1e9 / peak if t_pp > 0 else None,
                      "mean_launch_ms": t_pp * 1e3, "share_of_step": c.contact_pp_ms / ms if ms > 0 else None}
    if c.gather_launches > 0 and c.gather_ms > 0:
        inc = 2.0 * float(counts[0]) + float(counts[4:].sum())
        hot = 2.0 * float(active[0]) + float(active[4:].sum())
        gb = 312.0 * n_real + 5.0 * inc + 64.0 * hot
        tg = c.gather_ms / c.gather_launches * 1e-3
        tgt, cap = traffic_for(tr, "k_gather", wname, steps, warmup)
        out["gather"] = {"kernel": "k_gather<2>", "bound": "hbm", "achieved": gb / tg / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": gb / tg / 1e9 / peak,
                         "bytes_model": "312 B/particle + 5 B/incidence + 64 B per incident side of a record in contact", "bytes_per_launch": gb,
                         "traffic": tgt, "traffic_capture": cap, "frac_traffic": (tgt / tg / 1e9 / peak) if tgt else None,
                         "mean_launch_ms": tg * 1e3, "share_of_step": c.gather_ms / ms if ms > 0 else None}
    return out


DRIFT = 8.0   # uniform drift velocity of the "flowing" polyhedra lines: 8e-4 per step against a trigger of rcut_inc/2 = 0.05


def polyhedra_line(lib, peak, peak_src, device, tr, n=500_000, steps=150, warmup=10, mixed=False):
    """BASELINE configs[2] (and, mixed=True, configs[3]) reported alongside (north_star): ~500 k octahedra, the vertex/edge/face contact
    path.  The timed region is a FLOWING state: the whole packing drifts through its periodic box (the walled mixed case: falls) at
    DRIFT, so the displacement trigger fires every ~60 steps and the timed region CONTAINS neighbour rebuilds — their cost is inside
    `value`, `rebuilds_in_timed_region` says how many."""
    from exadem_b200 import Engine, workloads
    side = int(round(n ** (1.0 / 3.0)))
    wl = mixed_workload(n) if mixed else workloads.polyhedra_box((side, side, side), pitch=0.78)
    if not mixed:   # (the walled mixed packing relaxes violently on its own: its trigger fires every ~10 steps without any drift)
        wl.particles.vx += DRIFT
    eng = wl.setup(Engine(lib, wl.config(device=device)))
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    # (the second rebuild of a context allocates the second set of wave buffers — the lists are double-buffered for the history carry-
    # over — and is the first one that records / replays the count pass: both belong to the warm-up, like for the drum below.  Without
    # it the first rebuild INSIDE the timed region paid 25-85 ms of cudaMalloc: 0.80-1.21 ms per step over three runs of the same code)
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, warmup, 0)
    eng.sync()
    eng.reset_counters()
    eng.set_timing(False)                       # timed region without the per-kernel events (see Job.timed)
    eng.event_record(0)
    nreb = eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, steps, 0)
    eng.event_record(1)
    ms = eng.event_elapsed_ms(0, 1)
    cv = eng.counters()
    eng.reset_counters()
    eng.set_timing(True)                        # second pass of the same length: per-kernel breakdown
    nreb2 = eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, steps, 0)
    eng.sync()
    c = eng.counters()
    counts, active = eng.interaction_counts()
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    eng.sync()
    eng.reset_counters()
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    eng.sync()
    reb = eng.counters().rebuild_ms
    eng.set_timing(False)
    n_real = wl.particles.n
    eng.close()
    t = c.contact_pp_ms / steps * 1e-3            # the two launches that cover the four particle-particle waves, per step
    rec = float(counts[:4].sum())
    hot = float(active[:4].sum())
    nv_mean = 6.0 if not mixed else 5.0
    # the two launches' own traffic: detection reads ij + sub + flg (17 B) per record and, through L1/L2, the two vertex sets of every
    # owner-partner pair (counted once per particle: 32 B x 6 vertices + type/homothety sectors); the law pass moves 128 B of history /
    # wrench + 6 particle sectors + the 4 B list entry per record in contact
    own = 17.0 * rec + (128.0 + 192.0 + 4.0) * hot + n_real * (32.0 * 6 + 64.0)
    # 184 B/record + 8 (sub) + 24 k of vertices (k = 2 V-V, 3 V-E, 1 + 3 V-F on triangles, 4 E-E) + 92 B/particle per wave
    alg = sum(float(counts[w]) * (192.0 + 24.0 * k) for w, k in ((0, 2), (1, 3), (2, 4), (3, 4))) + 4 * 92.0 * n_real
    tp, cap = traffic_for(tr, "k_detect_poly_pp+k_law_poly_pp", "poly", steps, warmup) if not mixed else (None, None)
    what = (f"{n_real} spheres + octahedra + cubes, cohesive law, closed mesh walls (configs[3])" if mixed
            else f"{n_real} octahedra, periodic packing (configs[2])")
    return {"workload": what + (f", drifting at {DRIFT} (flowing state)" if not mixed else ", relaxing from the lattice (flowing state)"), "value": cv.interaction_updates / (ms * 1e-3),
            "unit": UNIT, "particle_steps_per_s": cv.particle_steps / (ms * 1e-3), "steps": steps, "ms_per_step": ms / steps,
            "interactions": int(counts.sum()), "active_interactions": int(active.sum()), "rebuilds_in_timed_region": int(nreb),
            "rebuild_once_ms": reb,
            "roofline": {"kernel": "k_detect_poly_pp + k_law_poly_pp (4 waves)", "bound": "hbm", "achieved": own / t / 1e9 if t > 0 else 0.0,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": own / t / 1e9 / peak if t > 0 else 0.0,
                         "bytes_model": "17 B/record + 324 B/record in contact + 256 B/particle (vertices, type, homothety)",
                         "bytes_per_launch": own, "traffic": tp, "traffic_capture": cap,
                         "frac_traffic": (tp / t / 1e9 / peak) if (tp and t > 0) else None,
                         "algorithmic_bytes_per_launch": alg, "frac_algorithmic": alg / t / 1e9 / peak if t > 0 else None,
                         "algorithmic_note": "SURVEY 8(d): 184 B/record + 8 (sub) + 24 k vertex bytes (k = 2 V-V, 3 V-E, 1+n_face V-F, 4 E-E) + 92 B/particle per wave",
                         "mean_launch_ms": t * 1e3,
                         "note": "ncu (profiles/r2_ncu_polystep.txt): the detection pass is bound by its scattered vertex gathers (fp64 pipe 10 %, "
                                 "long-scoreboard 27 per issue), ~95 % of the records are candidates that do not touch"},
            "per_step_ms": {"contact_all": c.contact_ms / steps, "gather_fused": c.gather_ms / steps, "ghost": c.ghost_ms / steps,
                            "rebuild_amortised": c.rebuild_ms / steps, "rebuilds_in_this_pass": int(nreb2),
                            "note": "second pass of the same number of steps with per-kernel CUDA events on"}}


# ------------------------------------------------------------------------------------------------ GPU arm
class Job:
    """one engine per rank on a workload, with the barrier / timing helpers of the contract"""

    def __init__(self, lib, make, rank, world, local, comm=True):
        from exadem_b200 import Engine, parallel
        self.world, self.rank = world, rank
        self.wl = make(rank, world)
        cfg = self.wl.config(device=local) if world == 1 else parallel.brick_config(self.wl, rank, world, device=local)
        self.eng = Engine(lib, cfg)
        if world > 1 and comm:
            parallel.init_comm(self.eng, lib, rank)
        self.wl.setup(self.eng)

    def barrier(self):
        self.eng.sync()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.eng.sync()

    def timed(self, steps, warmup):
        """W warm-up steps, then EXACTLY `steps` steps bracketed by barrier + synchronize; CUDA events on the library stream.
        The per-kernel CUDA events of the library (edb_set_timing) cost ~0.015 ms per step (measured: 0.219 -> 0.234 ms on the 1 M
        drum, scripts/timing_ab.py), so they stay OFF in the timed region; the per-kernel breakdown comes from a second pass of the
        same number of steps right behind it (`self.breakdown`: counters of that pass, its ms, the interaction counts after it).
        -> (ms on this rank, counters of the timed region, rebuilds, (wall t0, t1))"""
        wl, eng = self.wl, self.eng
        eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
        eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, max(warmup, 3), 0)
        self.barrier()
        eng.reset_counters()
        eng.set_timing(False)
        self.barrier()
        eng.event_record(0)
        w0 = time.time()
        nreb = eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, steps, 0)
        eng.event_record(1)
        ms = eng.event_elapsed_ms(0, 1)
        self.barrier()
        w1 = time.time()
        c = eng.counters()
        # second pass: the same number of steps with the per-kernel events on
        eng.reset_counters()
        eng.set_timing(True)
        self.barrier()
        eng.event_record(2)
        eng.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, wl.symetric, steps, 0)
        eng.event_record(3)
        ms2 = eng.event_elapsed_ms(2, 3)
        self.barrier()
        c2 = eng.counters()
        eng.set_timing(False)
        counts, active = eng.interaction_counts()
        self.breakdown = {"c": c2, "ms": ms2, "counts": counts, "active": active, "steps": steps}
        return ms, c, nreb, (w0, w1)

    def per_step_ms(self):
        b = self.breakdown
        c, k = b["c"], b["steps"]
        return {"contact_all": c.contact_ms / k, "gather_fused": c.gather_ms / k, "prolog": c.prolog_ms / k, "ghost": c.ghost_ms / k,
                "rebuild_amortised": c.rebuild_ms / k, "step_with_kernel_events": b["ms"] / k}


def reduce_stats(world, ms, c, nreb):
    """max over ranks of the time, sums of the work -> (ms, interaction updates, particle steps, launches, rebuilds max)"""
    stats = np.array([ms, float(c.interaction_updates), float(c.particle_steps), float(c.kernel_launches), float(nreb)], dtype=np.float64)
    if world == 1:
        return tuple(stats)
    import torch
    import torch.distributed as dist
    t = torch.tensor(stats, device="cuda")
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(tmax[0]), float(t[1]), float(t[2]), float(t[3]), float(tmax[4])


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    from exadem_b200 import Library

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
    name, make, desc, scaling = resolve_workload(args, world)
    lib = Library()
    job = Job(lib, make, rank, world, local)
    wl, eng = job.wl, job.eng
    g = wl.gravity
    windows = []

    ms, c, nreb, win = job.timed(args.steps, args.warmup)
    windows.append(win)
    counts, active = job.breakdown["counts"], job.breakdown["active"]
    cb, msb_pass = job.breakdown["c"], job.breakdown["ms"]     # per-kernel pass (events on), the steps right behind the timed region
    ms, iu, ps, launches, nreb_max = reduce_stats(world, ms, c, nreb)
    value = iu / (ms * 1e-3)

    # ---- e2e: the host owns the grid; every step ships the dynamic fields in and the results out (pinned memory)
    e2e = None
    if not args.no_e2e:
        host = pinned_like(eng.download_particles())
        h2d = 13 * 8 * host.n
        d2h = 22 * 8 * host.n
        ke = max(5, min(args.steps, 40))
        for _ in range(3):   # untimed: the first iterations allocate the staging buffers, copy streams and events of the async path
            eng.upload_dynamic_async(host)
            eng.run_verlet(wl.dt, g, wl.rcut_inc, wl.symetric, 1, 0)
            eng.download_dynamic_async(host)
        eng.sync()
        if world > 1:   # (a rebuild inside the warm-up may have migrated particles: start the timed loop from the device's view)
            host = pinned_like(eng.download_particles())
        eng.reset_counters()
        job.barrier()
        w0 = time.time()
        t0 = time.perf_counter()
        hv = host.view()
        t_host = 0.0
        for _ in range(ke):
            # asynchronous entry points: the download of step n (all 22 fields), the upload of step n+1 (the arrays step n filled) and
            # the kernels of step n+1 overlap on the full-duplex link; the host only waits at the end (and inside edb_run_verlet for
            # the rebuild decision of the step)
            th = time.perf_counter()
            eng.upload_dynamic_async(host, view=hv)
            t_host += time.perf_counter() - th
            nr = eng.run_verlet(wl.dt, g, wl.rcut_inc, wl.symetric, 1, 0)
            if nr and world > 1:   # a rebuild may have migrated particles between bricks: the host mirror follows
                eng.sync()
                host = pinned_like(eng.download_particles())
                hv = host.view()
            else:
                th = time.perf_counter()
                eng.download_dynamic_async(host, view=hv)
                t_host += time.perf_counter() - th
        job.barrier()
        dt = time.perf_counter() - t0
        windows.append((w0, time.time()))
        ce = eng.counters()
        iu_e = float(ce.interaction_updates)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([iu_e, dt, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
            tm = t.clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            iu_e, dt, h2d, d2h = float(t[0]), float(tm[1]), float(t[2]), float(t[3])
        link = link_bandwidth() if rank == 0 else {}
        e2e = {"value": iu_e / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": ke,
               "ms_per_step": 1e3 * dt / ke, "link": link, "host_enqueue_ms_per_step": 1e3 * t_host / ke,
               "floor_ms_per_step": (1e3 * (d2h / world) / (link["d2h_gbs_while_h2d"] * 1e9)) if link.get("d2h_gbs_while_h2d") else None,
               "floor": "d2h bytes per rank / measured device-to-host bandwidth while a host-to-device copy runs beside it (the two copies of consecutive steps overlap)",
               "what": "edb_upload_dynamic_async (r, v, vrot, orient from pinned host memory) -> edb_run_verlet(1 step) -> "
                       "edb_download_dynamic_async (r, v, vrot, orient, f, mom, arot to pinned host memory), every step, all ranks; "
                       "PCIe floor = d2h bytes / link bandwidth (the copies of consecutive steps overlap, the download is the longer leg)"}

    # ---- cost of one neighbour rebuild (outside the timed region; the drum moves too slowly to trigger one in K steps)
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)   # (the first explicit rebuild allocates the second set of wave buffers)
    job.barrier()
    eng.reset_counters()
    eng.set_timing(True)
    eng.update_particle_neighbors(wl.rcut_inc, wl.symetric)
    job.barrier()
    rebuild_once_ms = eng.counters().rebuild_ms
    eng.set_timing(False)
    n_real = wl.particles.n
    eng.close()

    # ---- N > 1, default workload: the BASELINE configs[4] topology reported alongside — a fully periodic box cut into 2x1x1 / 2x2x1 /
    # 2x2x2 bricks of 1 M spheres (7 peers per rank at N = 8) — together with one such brick alone on one GPU (periodic onto itself),
    # the denominator of the weak-scaling efficiency on that topology
    box = None
    if world > 1 and not args.no_secondary and name == "drum" and args.scaling == "weak":
        import copy
        import torch.distributed as dist
        bargs = copy.copy(args)
        bargs.workload = "box"
        _, bmake, bdesc, _ = resolve_workload(bargs, world)
        jb = Job(lib, bmake, rank, world, local)
        msb, cbx, nrb, winb = jb.timed(args.steps, args.warmup)
        windows.append(winb)
        bcounts, bactive = jb.breakdown["counts"], jb.breakdown["active"]
        msb, iub, psb, _, nrbm = reduce_stats(world, msb, cbx, nrb)
        box = {"workload": bdesc, "particles_per_gpu": int(jb.wl.particles.n), "value": iub / (msb * 1e-3), "ms_per_step": msb / args.steps,
               "rebuilds_in_timed_region": int(nrbm), "interactions_per_gpu": int(bcounts.sum()), "active_interactions_per_gpu": int(bactive.sum()),
               "per_step_ms": jb.per_step_ms()}
        jb.eng.close()
        dist.barrier()
        if rank == 0:
            j1 = Job(lib, lambda r, n: bmake(0, 1), 0, 1, local)
            ms1, c1, nreb1, win1 = j1.timed(args.steps, args.warmup)
            windows.append(win1)
            box["single_gpu_same_brick"] = {"workload": "one brick of the same size alone on one GPU (periodic onto itself)",
                                            "particles": int(j1.wl.particles.n), "value": c1.interaction_updates / (ms1 * 1e-3),
                                            "ms_per_step": ms1 / args.steps, "rebuilds_in_timed_region": int(nreb1),
                                            "per_step_ms": j1.per_step_ms()}
            j1.eng.close()
        dist.barrier()

    single = None
    if world > 1 and not args.no_secondary and name == "box" and args.scaling == "weak":
        import torch.distributed as dist
        if rank == 0:
            j1 = Job(lib, lambda r, n: make(0, 1), 0, 1, local)
            ms1, c1, nreb1, win1 = j1.timed(args.steps, args.warmup)
            windows.append(win1)
            single = {"workload": "one brick of the same size alone on one GPU (periodic onto itself)", "particles": int(j1.wl.particles.n),
                      "value": c1.interaction_updates / (ms1 * 1e-3), "ms_per_step": ms1 / args.steps, "rebuilds_in_timed_region": int(nreb1),
                      "per_step_ms": j1.per_step_ms()}
            j1.eng.close()
        dist.barrier()
    if rank != 0:
        return
    clocks = sampler.stop(windows)
    peak, peak_src = peaks()
    tr = load_traffic()
    per_step = job.per_step_ms()
    per_step["note"] = ("second pass of the same number of steps with the library's per-kernel CUDA events on (they cost ~0.015 ms per step "
                        "and are off in the timed region)")
    roof = roof_other = None
    if wl.mode == 0:
        rf = sphere_rooflines(cb, counts, active, n_real, msb_pass, args.steps, peak, peak_src, tr, name, max(args.warmup, 3) + args.steps)
        # `roofline` = the kernel with the larger share of the step; the other one rides beside it under its own key
        keys = sorted(rf, key=lambda k: -(rf[k]["share_of_step"] or 0.0))
        roof = rf[keys[0]]
        roof_other = rf[keys[1]] if len(keys) > 1 else None
    else:
        t = cb.contact_pp_ms / args.steps * 1e-3
        alg = cb.bytes_contact_algorithmic / args.steps
        roof = {"kernel": "k_detect_poly_pp + k_law_poly_pp", "bound": "hbm", "achieved": alg / t / 1e9 if t > 0 else 0.0, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": alg / t / 1e9 / peak if t > 0 else 0.0, "traffic": None,
                "bytes_model": "SURVEY 8(d) algorithmic bytes incl. vertices", "mean_launch_ms": t * 1e3,
                "share_of_step": cb.contact_pp_ms / msb_pass if msb_pass > 0 else None}
    roof["per_step_ms"] = per_step
    roof["rebuild_once_ms"] = rebuild_once_ms
    poly = mixed = None
    if world == 1 and not args.no_poly and name == "drum":
        poly = polyhedra_line(lib, peak, peak_src, local, tr)
        if not args.no_mixed:
            mixed = polyhedra_line(lib, peak, peak_src, local, tr, n=4_000_000, steps=40, warmup=5, mixed=True)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_leg(make(0, 1), args.cpu_seconds)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "particle_steps_per_s": ps / (ms * 1e-3),
            "config": {"workload": desc, "rebuild": "displacement trigger rcut_inc/2", "l2": L2_NOTE},
            "run": {"particles_per_gpu": int(n_real), "particles": int(ps / args.steps), "interactions_per_gpu": int(counts.sum()),
                    "active_interactions_per_gpu": int(active.sum()), "rebuilds_in_timed_region": int(nreb_max),
                    "timing": "CUDA events on the library stream, max over ranks", "wall_s": windows[0][1] - windows[0][0]},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    if roof_other is not None:
        line["roofline_" + ("gather" if roof_other["kernel"].startswith("k_gather") else "contact")] = roof_other
    if poly is not None:
        line["polyhedra"] = poly
    if mixed is not None:
        line["mixed_4m"] = mixed
    if box is not None:
        line["box_bricks"] = box
    if single is not None:
        line["single_gpu_same_brick"] = single
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
