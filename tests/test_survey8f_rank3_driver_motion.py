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


def test_product_integrators_compiled_for_host_vs_oracle(tmp_path):
    """The product's integrator code (csrc/kernels_driver.cuh, __host__ __device__) compiled for the host by nvcc and run on the same
    cases: bit-identical to the oracle.  Checks the code the GPU runs without a GPU (device sin/cos aside)."""
    import ctypes as C
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    so = tmp_path / "libhost_check_driver.so"
    src = Path(__file__).parent / "host_check_driver.cu"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                    "-shared", str(src), "-o", str(so)], check=True, capture_output=True)
    chk = C.CDLL(str(so))
    cases = driver_cases.all_cases(seed=5, reps=10)
    _assert_bitwise(_oracle_rows(cases), driver_cases.run_cases(chk.chk_driver_motion_op, cases), cases)


def test_product_cycle_kernel_body_compiled_for_host_vs_oracle_sequence(tmp_path):
    """The body of the fused step-boundary kernel (driver_cycle_body = what k_driver_cycle runs per driver: reset, gravity on PARTICLE
    meshes, contact reaction, force_to_accel_driver, push_f_v(dt/2) | push_f_v_r, push_f_v(dt/2), push_av_to_quat) compiled for the host
    against the same sequence composed from the oracle's operators: bit-identical for every driver kind and motion type."""
    import ctypes as C
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    so = tmp_path / "libhost_check_driver.so"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                    "-shared", str(Path(__file__).parent / "host_check_driver.cu"), "-o", str(so)], check=True, capture_output=True)
    chk, orc = C.CDLL(str(so)), OracleLibrary().dll
    rng = np.random.default_rng(99)
    n = 0
    for kind, motions in driver_cases.VALID.items():
        for motion in motions:
            for do_end, do_begin in ((1, 1), (1, 0), (0, 1)):
                for _ in range(6):
                    d, m, time, dt, smass = driver_cases.make_case(rng, kind, motion)
                    g = np.array([0.1, -0.2, -9.807])
                    sums = np.ascontiguousarray(rng.normal(size=6) * 5.0)
                    sphere = int(rng.integers(0, 2))
                    rows = []
                    for fn in (orc.orc_driver_cycle, chk.chk_driver_cycle):
                        keep = []
                        D, M = driver_cases.to_structs(d, m, keep)
                        rc = fn(C.byref(D), C.byref(M), g.ctypes.data_as(C.c_void_p), sums.ctypes.data_as(C.c_void_p), C.c_int(sphere),
                                C.c_double(smass), C.c_double(dt), C.c_double(time), C.c_int(do_end), C.c_int(do_begin))
                        assert rc == 0
                        rows.append(driver_cases.flatten(D, M))
                    same = (rows[0] == rows[1]) | (np.isnan(rows[0]) & np.isnan(rows[1]))
                    assert same.all(), (kind, motion, do_end, do_begin, np.argwhere(~same).ravel(), rows[0][~same], rows[1][~same])
                    n += 1
    assert n > 300


# ---------------------------------------------------------------------------------------------------------------------------------
# scenarios with moving drivers, run on two engines (GPU vs oracle in the gpu tests; oracle vs oracle as a CPU self-check of the
# scenario itself: drivers are touched, rebuilds are triggered by the drivers, ...)
# ---------------------------------------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402

from exadem_b200 import capi, workloads  # noqa: E402
from parity import compare_interaction_sets, compare_particles, field_close  # noqa: E402

RTOL = 1e-10
DRIVER_FIELDS = ("radius", "center", "offset", "normal", "vel", "vrot", "quat", "acc", "rv", "ra", "forces", "mom")


def compare_drivers(a, b, n, rtol=RTOL, what=""):
    da, db = a.get_drivers(n), b.get_drivers(n)
    for i in range(n):
        for k in DRIVER_FIELDS:
            field_close(np.atleast_1d(da[i][k]), np.atleast_1d(db[i][k]), rtol, f"{what} driver {i} {k}")
    return da


def scenario_piston():
    """Spheres between a fixed floor and a roof moving down at constant velocity (surface, LINEAR_MOTION): the roof alone triggers the
    rebuild (driver_displ_over) well before any particle has moved rcut_inc / 2."""
    wl = workloads.deposit((7, 7, 4))
    wl.drivers.append(dict(kind=capi.DRIVER_SURFACE, offset=4.45, normal=(0, 0, 1), center=(0, 0, 4.45), vel=(0, 0, -25.0)))
    wl.motions = [dict(motion_type="STATIONARY"),
                  dict(motion_type="LINEAR_MOTION", motion_vector=(0, 0, 1), const_vel=-25.0)]
    return wl


def scenario_walls_and_balls():
    """Shaken floor (SHAKER), stress-controlled roof (LINEAR_COMPRESSIVE_MOTION: needs the contact reaction and the system mass), a ball
    on a tabulated trajectory and a breathing ball (COMPRESSIVE_FORCE) around the bed."""
    wl = workloads.deposit((7, 7, 4))
    wl.drivers[0] = dict(kind=capi.DRIVER_SURFACE, offset=0.05, normal=(0, 0, 1), center=(0, 0, 0.05))
    wl.drivers.append(dict(kind=capi.DRIVER_SURFACE, offset=4.3, normal=(0, 0, 1), center=(0, 0, 4.3)))
    wl.drivers.append(dict(kind=capi.DRIVER_BALL, radius=1.0, center=(3.5, 3.5, 2.0)))
    wl.drivers.append(dict(kind=capi.DRIVER_BALL, radius=5.45, center=(3.5, 3.5, 2.0)))
    t = np.array([0.0, 0.002, 0.004, 0.01])
    x = np.array([[3.5, 3.5, 2.0], [3.6, 3.5, 2.1], [3.8, 3.4, 2.1], [3.8, 3.4, 2.6]])
    wl.motions = [dict(motion_type="SHAKER", omega=900.0, amplitude=0.04, shaker_dir=(0, 0, 1)),
                  dict(motion_type="LINEAR_COMPRESSIVE_MOTION", motion_vector=(0, 0, -1), sigma=2.0, damprate=0.5, surface=49.0),
                  dict(motion_type="TABULATED", tab_time=t, tab_pos=x),
                  dict(motion_type="COMPRESSIVE_FORCE", sigma=0.01, damprate=0.2, rv=-2.0)]
    return wl


def scenario_moving_mesh(free_body=True):
    """Spheres inside a box mesh that translates and spins (LINEAR_MOTION + vrot) and, inside it, a small mesh that is a free body
    (PARTICLE: driven by the contact forces and moments).  Exercises driver_vertices, grid_rshape_driver and the vertex trigger."""
    wl = workloads.spheres_box((8, 8, 6), jitter=0.03, omega_sigma=0.3, polydisperse=0.1, cell_size=2.0)
    wl.lo, wl.hi = (-1.5, -1.5, -1.5), (9.5, 9.5, 7.5)
    spec, centre = workloads.box_mesh((0.03, 0.04, 0.05), (7.96, 7.97, 5.98), 0.02)
    spec2, _ = workloads.box_mesh((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), 0.05)
    wl.mesh_spec = [spec, spec2]
    c45 = np.sqrt(0.5)
    wl.drivers = list(wl.drivers) + [
        dict(kind=capi.DRIVER_RSHAPE, shape=0, center=centre, vel=(30.0, 0.0, 0.0), vrot=(0.0, 0.0, 4.0)),
        dict(kind=capi.DRIVER_RSHAPE, shape=1, center=(4.0, 4.0, 3.0), quat=(c45, 0.0, c45 * 0.6, c45 * 0.8), vel=(0.0, 0.5, 0.0),
             vrot=(0.3, 0.0, 0.2))]
    wl.periodic = (0, 0, 0)
    nd0 = len(wl.drivers) - 2
    wl.motions = [dict(motion_type="STATIONARY")] * nd0 + [
        dict(motion_type="LINEAR_MOTION", motion_vector=(1, 0, 0), const_vel=30.0),
        dict(motion_type="PARTICLE" if free_body else "SHAKER", mass=3.0, inertia=(2.0, 2.5, 3.0), omega=700.0, amplitude=0.03,
             shaker_dir=(0, 0, 1))]
    return wl


def scenario_polyhedra_mesh_shaker():
    wl = workloads.polyhedra_box((5, 5, 5), shapes=("octahedron", "cube"), pitch=0.8, cell_size=2.0)
    wl.lo, wl.hi = (-2.0, -2.0, -2.0), (6.0, 6.0, 6.0)
    spec, centre = workloads.box_mesh((0.1, 0.12, 0.11), (3.9, 3.88, 3.91), 0.03)
    wl.mesh_spec = [spec]
    wl.drivers = list(wl.drivers) + [dict(kind=capi.DRIVER_RSHAPE, shape=0, center=centre, vrot=(0.0, 2.0, 0.0))]
    wl.periodic = (0, 0, 0)
    wl.motions = [dict(motion_type="STATIONARY")] * (len(wl.drivers) - 1) + [
        dict(motion_type="SHAKER", omega=2000.0, amplitude=0.05, shaker_dir=(0, 0, 1))]
    return wl


def scenario_pendulum():
    """A wall hinged above the bed that swings through the vertical (surface, PENDULUM_MOTION): normal and offset are recomputed from
    the pendulum every step and the normal flips when the swing changes side (driver_params.hpp:340-364)."""
    wl = workloads.deposit((7, 7, 4))
    anchor = (0.02, 3.5, 8.0)
    wl.drivers.append(dict(kind=capi.DRIVER_SURFACE, offset=0.02, normal=(1, 0, 0), center=anchor))
    wl.motions = [dict(motion_type="STATIONARY"),
                  dict(motion_type="PENDULUM_MOTION", omega=500.0, amplitude=0.3, pendulum_anchor_point=anchor,
                       pendulum_initial_position=(0.02, 3.5, 0.0), pendulum_swing_dir=(1, 0, 0))]
    return wl


def scenario_mesh_force_motion():
    """The small mesh pushed by a constant force along x plus the projection of the contact reaction (LINEAR_FORCE_MOTION)."""
    wl = scenario_moving_mesh()
    wl.name = "mesh_force_motion"
    wl.motions[-1] = dict(motion_type="LINEAR_FORCE_MOTION", motion_vector=(1, 0, 0), const_force=40.0, mass=3.0)
    return wl


def run_scenario(wl, engines, nsteps, expect_driver_waves, min_rebuilds=1):
    """rebuild + run_verlet with the displacement trigger on every engine; parity of the interaction set, the particles and the drivers."""
    a, b = engines
    nd = len(wl.drivers)
    for e in engines:
        e.update_particle_neighbors(wl.rcut_inc, True)
    nb = [e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, nsteps, 0) for e in engines]
    assert nb[0] == nb[1] and nb[0] >= min_rebuilds, nb
    compare_interaction_sets(a.export_interactions(), b.export_interactions(), payload_rtol=RTOL)
    compare_particles(a.download_particles(), b.download_particles(), rtol=RTOL)
    drv = compare_drivers(a, b, nd, what=wl.name)
    assert abs(a.get_physical_time() - nsteps * wl.dt) < 1e-12 and abs(b.get_physical_time() - a.get_physical_time()) < 1e-15
    cnt, act = a.interaction_counts()
    for t in expect_driver_waves:
        assert cnt[t] > 0, (t, cnt)
    return drv, cnt, act


def step_by_step(wl, engines, nsteps):
    """The macro of config_numerical_schemes.msp:12-29 operator by operator through the C ABI, drivers compared after every driver
    operator (what a reference deck does when it calls the operators by name)."""
    nd = len(wl.drivers)
    sphere = wl.mode == capi.MODE_SPHERE
    t = 0.0
    smass = float(wl.particles.mass.sum())
    for e in engines:
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.backup_drivers()
    for it in range(nsteps):
        for name, op in (("prolog", lambda e: e.combined_compute_prolog(wl.dt)),
                         ("push_f_v_r_driver", lambda e: e.push_f_v_r_driver(t, wl.dt)),
                         ("push_f_v_driver", lambda e: e.push_f_v_driver(0.5 * wl.dt)),
                         ("push_av_to_quat_driver", lambda e: e.push_av_to_quat_driver(t, wl.dt))):
            for e in engines:
                op(e)
            if "driver" in name:
                compare_drivers(*engines, nd, what=f"step {it} after {name}")
        over = [e.displ_over(0.5 * wl.rcut_inc) + e.driver_displ_over(0.5 * wl.rcut_inc) for e in engines]
        assert (over[0] > 0) == (over[1] > 0), over
        for e in engines:
            if over[0] > 0:
                e.backup_drivers()
                e.driver_vertices()
                e.update_particle_neighbors(wl.rcut_inc, True)
            else:
                e.ghost_update_rq()
                if not sphere:
                    e.compute_vertices()
                e.driver_vertices()
            e.reset_force_driver()
            e.reset_force_moment()
            e.gravity_force(wl.gravity)
            (e.contact_sphere if sphere else e.contact_polyhedron)(wl.dt)
            e.force_to_accel()
            e.force_to_accel_driver(wl.dt, smass)
        compare_drivers(*engines, nd, what=f"step {it} after force_to_accel_driver")
        for e in engines:
            e.push_f_v_driver(0.5 * wl.dt)
            e.combined_compute_epilog(wl.dt)
        compare_drivers(*engines, nd, what=f"step {it} end")
        t += wl.dt
    compare_interaction_sets(engines[0].export_interactions(), engines[1].export_interactions(), payload_rtol=RTOL)
    compare_particles(engines[0].download_particles(), engines[1].download_particles(), rtol=RTOL)


SCENARIOS = {
    "piston": (scenario_piston, 60, (5,)),
    "walls_and_balls": (scenario_walls_and_balls, 40, (5, 6)),
    "moving_mesh": (scenario_moving_mesh, 30, (7, 8, 9)),
    "polyhedra_mesh_shaker": (scenario_polyhedra_mesh_shaker, 24, (7, 8, 9, 10, 11, 12)),
    "pendulum": (scenario_pendulum, 80, (5,)),
    "mesh_force_motion": (scenario_mesh_force_motion, 30, (7, 8, 9)),
}


def _oracle_pair(wl):
    from oracle_binding import OracleEngine
    lib = OracleLibrary()
    engines = [wl.setup(OracleEngine(lib, wl.config())) for _ in range(2)]
    engines[1].set_threads(1)
    return engines


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_scenarios_are_meaningful_on_the_oracle(name):
    """CPU self-check of the scenarios the GPU tests run: the drivers move, touch particles and (where stated) trigger rebuilds."""
    make, nsteps, waves = SCENARIOS[name]
    wl = make()
    engines = _oracle_pair(wl)
    before = engines[0].get_drivers(len(wl.drivers))
    drv, cnt, act = run_scenario(wl, engines, nsteps, waves, min_rebuilds=0 if name in ("walls_and_balls", "pendulum") else 1)
    moved = [i for i in range(len(drv)) if not all(np.array_equal(np.atleast_1d(before[i][k]), np.atleast_1d(drv[i][k])) for k in DRIVER_FIELDS)]
    assert moved, "no driver changed state"
    assert sum(act[t] for t in waves) > 0, (act, "no active driver contact")


def test_step_by_step_on_the_oracle():
    wl = scenario_walls_and_balls()
    step_by_step(wl, _oracle_pair(wl), 6)


# ---- GPU: the device-resident drivers against the oracle --------------------------------------------------------------------------
def _gpu_pair(wl, product_lib, oracle_lib):
    from exadem_b200 import Engine
    from oracle_binding import OracleEngine
    return [wl.setup(Engine(product_lib, wl.config())), wl.setup(OracleEngine(oracle_lib, wl.config()))]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_gpu_run_verlet_with_moving_drivers(product_lib, oracle_lib, name):
    make, nsteps, waves = SCENARIOS[name]
    wl = make()
    run_scenario(wl, _gpu_pair(wl, product_lib, oracle_lib), nsteps, waves, min_rebuilds=0 if name in ("walls_and_balls", "pendulum") else 1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["walls_and_balls", "moving_mesh"])
def test_gpu_driver_operators_one_by_one(product_lib, oracle_lib, name):
    wl = SCENARIOS[name][0]()
    step_by_step(wl, _gpu_pair(wl, product_lib, oracle_lib), 8)


@pytest.mark.gpu
def test_gpu_expression_motion(product_lib, oracle_lib):
    """EXPRESSION: the binding evaluates the tinyexpr strings on the host and hands the values over (edb_set_driver_expr)."""
    wl = scenario_moving_mesh(free_body=False)
    nd = len(wl.drivers)
    wl.motions[-1] = dict(motion_type="EXPRESSION", expr_use_v=1, expr_use_vrot=1, expr_use_mom=1, drive_by_mom=1, mass=3.0,
                          inertia=(2.0, 2.5, 3.0), applied_mom=(0, 0, 1.0), mom_axis=(0, 0, 1.0))
    engines = _gpu_pair(wl, product_lib, oracle_lib)
    t = 0.0
    for it in range(5):
        v, w, m = (0.1 * np.sin(50 * t), 0.2, 0.0), (0.0, 0.0, 1.0 + t), (0.0, 0.5 * np.cos(10 * t), 2.0)
        for e in engines:
            e.set_driver_expr(nd - 1, v, w, m)
            e.push_f_v_r_driver(t, wl.dt)
            e.push_f_v_driver(0.5 * wl.dt)
            e.push_av_to_quat_driver(t, wl.dt)
        compare_drivers(*engines, nd, what=f"expression step {it}")
        t += wl.dt


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_gpu_fused_loop_equals_operator_loop(product_lib, monkeypatch, name):
    """edb_run_verlet with moving drivers: the speculative fused loop (k_driver_cycle between gather and contact, no host sync per
    step) against the operator-by-operator loop (EDB_DRIVERS_UNFUSED=1, read at context creation) — bit-identical state."""
    from exadem_b200 import Engine
    from parity import bit_identical
    make, nsteps, _ = SCENARIOS[name]
    wl = make()
    fused = wl.setup(Engine(product_lib, wl.config()))
    monkeypatch.setenv("EDB_DRIVERS_UNFUSED", "1")
    loop = wl.setup(Engine(product_lib, wl.config()))
    monkeypatch.delenv("EDB_DRIVERS_UNFUSED")
    nb = []
    for e in (fused, loop):
        e.update_particle_neighbors(wl.rcut_inc, True)
        n1 = e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, nsteps // 2, 0)
        n2 = e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, nsteps - nsteps // 2, 0)   # call boundary inside the run
        nb.append(n1 + n2)
    assert nb[0] == nb[1], nb
    assert fused.counters().kernel_launches < loop.counters().kernel_launches
    same = bit_identical(fused.download_particles(), loop.download_particles())
    assert all(same.values()), same
    da, db = fused.get_drivers(len(wl.drivers)), loop.get_drivers(len(wl.drivers))
    for a, b in zip(da, db):
        for k in DRIVER_FIELDS:
            assert np.array_equal(np.atleast_1d(a[k]), np.atleast_1d(b[k])), k
    assert fused.get_physical_time() == loop.get_physical_time()
