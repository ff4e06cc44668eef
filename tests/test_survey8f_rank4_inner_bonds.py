"""Inner bonds (fragmentation, wave 13), SURVEY 8(f) rank 4 — the per-step operators: inner_bond_polyhedron (bond law + deterministic
force accumulation), update_interfaces, apply_interface_fracture_criterion.

CPU: the oracle's law equals the reference's own force_law_core bit for bit (oracle/_ref + golden vectors); the product's law compiled
for the host equals the oracle; the glued-block scenario is meaningful.  GPU: the block through the C ABI against the oracle."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from exadem_b200 import capi, workloads
from oracle_binding import OracleEngine, OracleLibrary, ref_math
from parity import compare_particles, field_close

GOLD = Path(__file__).parent / "golden" / "inner_bond_golden.npz"
RTOL = 1e-10
_p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def law_cases(seed=11, n=400):
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(n):
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        cases.append(dict(dn=float(rng.normal() * 1e-3), n=nrm, dn0=float(rng.normal() * 1e-4), weight=float(rng.choice([0.25, 1 / 3, 0.5, 1.0])),
                          dt=float(rng.choice([1e-5, 1e-4])), prm=(float(rng.uniform(1e3, 1e6)), float(rng.uniform(1e3, 1e6)), float(rng.uniform(0, 1)),
                                                                    int(rng.integers(0, 3)), float(rng.uniform(0.1, 10)), float(rng.uniform(0.1, 10))),
                          meff=float(rng.uniform(0.01, 2.0)), tds=rng.normal(size=3) * 1e-4, vec=rng.normal(size=(7, 3))))
    return cases


def run_law(fn, cases):
    out = []
    for c in cases:
        prm = capi.InnerBondParams(kn=c["prm"][0], kt=c["prm"][1], damp_rate=c["prm"][2], mode=c["prm"][3], crit1=c["prm"][4], crit2=c["prm"][5])
        en_et, tds, ft, fi = np.zeros(2), c["tds"].copy(), np.zeros(3), np.zeros(3)
        cp, pi, vi, wi, pj, vj, wj = [np.ascontiguousarray(v) for v in c["vec"]]
        nn = np.ascontiguousarray(c["n"])
        fn.restype = None
        fn(C.c_double(c["dn"]), _p(nn), C.c_double(c["dn0"]), C.c_double(c["weight"]), C.c_double(c["dt"]), C.byref(prm), C.c_double(c["meff"]),
           _p(en_et), _p(tds), _p(ft), _p(cp), _p(pi), _p(vi), _p(fi), _p(wi), _p(pj), _p(vj), _p(wj))
        out.append(np.concatenate([en_et, tds, ft, fi]))
    return np.array(out)


def test_oracle_bond_law_vs_golden():
    cases = law_cases()
    g = np.load(GOLD)["out"]
    assert np.array_equal(run_law(OracleLibrary().dll.orc_inner_bond_force_core, cases), g)


def test_oracle_bond_law_vs_reference():
    ref = ref_math()
    if ref is None or not hasattr(ref, "ref_inner_bond_force_core"):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    cases = law_cases(seed=12, n=800)
    a, b = run_law(OracleLibrary().dll.orc_inner_bond_force_core, cases), run_law(ref.ref_inner_bond_force_core, cases)
    assert np.array_equal(a, b) and np.abs(a).sum() > 0


def test_product_bond_law_compiled_for_host_vs_oracle(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    so = tmp_path / "libhost_check_bond.so"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                    str(Path(__file__).parent / "host_check_bond.cu"), "-o", str(so)], check=True, capture_output=True)
    chk, orc = C.CDLL(str(so)), OracleLibrary().dll
    cases = law_cases(seed=13, n=600)
    assert np.array_equal(run_law(orc.orc_inner_bond_force_core, cases), run_law(chk.chk_inner_bond_force_core, cases))
    rng = np.random.default_rng(5)
    hits = 0
    for _ in range(400):
        size = int(rng.integers(3, 6))
        en, et, dn, fn = rng.uniform(0, 1, size), rng.uniform(0, 1, size), rng.normal(size=size), np.ascontiguousarray(rng.normal(size=(size, 3)))
        mode, c1, c2 = int(rng.integers(0, 4)), float(rng.uniform(0, size)), float(rng.uniform(0, 1.5))
        args = (C.c_int(mode), C.c_double(c1), C.c_double(c2), _p(en), _p(et), _p(dn), _p(fn), C.c_uint64(size))
        a, b = orc.orc_interface_breaks(*args), chk.chk_interface_breaks(*args)
        assert a == b
        hits += a
    assert 50 < hits < 350


# ---- the glued block --------------------------------------------------------------------------------------------------------------
def glued_block():
    """4 x 4 x 3 cubes (side 0.6 + Minkowski radius 0.1) on an exact lattice of pitch 0.8 so that neighbouring faces touch: one cluster,
    glued by stick_polyhedra (4 bonds of weight 1/4 per shared face).  The top layer is sheared off so that its interfaces break."""
    wl = workloads.polyhedra_box((4, 4, 3), shapes=("cube",), pitch=0.8, jitter=0.0, periodic=(0, 0, 0), cell_size=2.0)
    pa = wl.particles
    pa.qw[:], pa.qx[:], pa.qy[:], pa.qz[:] = 1.0, 0.0, 0.0, 0.0
    pa.vrotx[:], pa.vroty[:], pa.vrotz[:] = 0.0, 0.0, 0.0
    pa.vx[:], pa.vy[:], pa.vz[:] = 0.0, 0.0, 0.0
    top = pa.rz > 1.7
    pa.vz[top] = -0.3     # sheared off under slight compression: the bonds break with a non-zero (capped) friction to hand over
    pa.vx[top] = 4.0
    wl.lo, wl.hi = (-2.0, -2.0, -2.0), (6.0, 6.0, 6.0)
    wl.periodic = (0, 0, 0)
    wl.name = "glued_block"
    ibp = [dict(kn=2e5, kt=1e5, damp_rate=0.1, mode=0, crit1=0.5, crit2=0.0)]
    return wl, ibp


def setup_block(engines, wl, ibp):
    """bonds are created by the oracle's stick_polyhedra (a one-off preprocessing that stays on the host side of the binding) and imported
    by every engine"""
    from oracle_binding import OracleEngine
    orc = next(e for e in engines if isinstance(e, OracleEngine))     # chosen by type: callers order their engines freely
    for e in engines:
        e.set_inner_bond_params(1, ibp)
        e.update_particle_neighbors(wl.rcut_inc, True)
    orc.set_particle_clusters(wl.particles.id, np.zeros(wl.particles.n))
    orc.stick_polyhedra(2e-5)
    bonds = orc.export_inner_bonds()
    for e in engines:
        e.import_inner_bonds(bonds)
        e.update_particle_neighbors(wl.rcut_inc, True)   # init_update_particle_neighbors: stick_polyhedra, then nbh_polyhedron (config_fragmentation.msp)
    n_itf = [e.update_interfaces() for e in engines]
    assert len(set(n_itf)) == 1
    return bonds, n_itf[0]


def glued_pairs_in_contact_list(e):
    """regular contact records (waves 0-3) between two particles that hold an unbroken bond: must be none (InteractionManager::ignore_)"""
    b = e.export_inner_bonds()
    glued = {(int(r["id_i"]), int(r["id_j"])) for r in b[b["unbroken"] != 0]}
    rec = e.export_interactions()
    pp = rec[rec["type"] <= 3]
    lo, hi = np.minimum(pp["id_i"], pp["id_j"]), np.maximum(pp["id_i"], pp["id_j"])
    return sum((int(a), int(c)) in glued for a, c in zip(lo, hi)), len(pp)


def step_block(e, wl, on_break=None):
    """one iteration of a fragmentation deck (config_fragmentation.msp:11-41 inside check_and_update_particles): the fracture criterion
    is evaluated on the state of the previous step; a broken interface caps its friction and forces the neighbour rebuild
    (displ_or_break), where the broken bonds turn into vertex-vertex contacts"""
    e.combined_compute_prolog(wl.dt)
    nb = e.apply_interface_fracture_criterion()
    if nb:
        e.broken_interface_friction_updater()
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.update_interfaces()
        if on_break is not None:
            on_break(e)
    else:
        e.ghost_update_rq()
        e.compute_vertices()
    e.reset_force_moment()
    e.gravity_force(wl.gravity)
    e.inner_bond_polyhedron(wl.dt)
    e.contact_polyhedron(wl.dt)
    e.force_to_accel()
    e.combined_compute_epilog(wl.dt)
    return nb


def bond_key(rec):
    return np.lexsort((rec["sub_j"], rec["sub_i"], rec["id_j"], rec["id_i"]))


def compare_bonds(a, b, rtol=RTOL):
    ra, rb = a.export_inner_bonds(), b.export_inner_bonds()
    ra, rb = ra[bond_key(ra)], rb[bond_key(rb)]
    for k in ("id_i", "id_j", "sub_i", "sub_j", "unbroken", "mode"):
        assert np.array_equal(ra[k], rb[k]), k
    for k in ("friction", "tds", "en", "et", "dn0", "weight", "criteria_1"):
        field_close(ra[k], rb[k], rtol, f"bond {k}")
    return ra


def run_block(engines, nsteps=40, on_break=None):
    wl, ibp = glued_block()
    engines = [wl.setup(e(wl)) for e in engines]
    bonds, n_itf = setup_block(engines, wl, ibp)
    # 4 x 4 x 3 cubes: 3*4*3 + 4*3*3 + 4*4*2 = 104 shared faces, 4 bonds each
    assert n_itf == 104 and len(bonds) == 416 and np.allclose(bonds["weight"], 0.25) and bonds["unbroken"].all()
    broken_total = [0] * len(engines)
    for it in range(nsteps):
        nb = [step_block(e, wl, on_break) for e in engines]
        assert len(set(nb)) == 1, (it, nb)
        for k, v in enumerate(nb):
            broken_total[k] += v
    return wl, engines, broken_total, bonds


def test_glued_block_on_the_oracle():
    lib = OracleLibrary()
    handed = []
    wl, engines, broken, bonds0 = run_block([lambda wl: OracleEngine(lib, wl.config())] * 2, on_break=lambda e: handed.append(e.export_interactions()))
    # right after the rebuild that follows the break, the vertex-vertex records of the torn faces carry the friction of their bonds
    vv0 = handed[0][handed[0]["type"] == 0]
    assert int((np.abs(vv0["friction"]).sum(axis=1) > 0).sum()) >= 4 * broken[0] > 0
    rec = compare_bonds(*engines)
    # the top layer tears off (its 16 interfaces break and leave wave 13 at the rebuild that follows), the rest stays glued
    assert 0 < broken[0] < 104 and len(rec) == 4 * (104 - broken[0]) and rec["unbroken"].all()
    assert rec["en"].max() > 0 and rec["et"].max() > 0
    dn, cp, fn, ft = engines[0].get_inner_bond_state(len(rec))
    assert np.abs(fn).sum() > 0
    # glued pairs have no regular contact records; the torn-off faces do, some of them with the friction inherited from their bonds
    e = engines[0]
    n_glued, n_pp = glued_pairs_in_contact_list(e)
    assert n_glued == 0 and n_pp > 0
    inter = e.export_interactions()
    vv = inter[inter["type"] == 0]
    torn = {(int(a), int(b)) for a, b in zip(bonds0["id_i"], bonds0["id_j"])} - {(int(a), int(b)) for a, b in zip(rec["id_i"], rec["id_j"])}
    from_bonds = [r for r in vv if (int(r["id_i"]), int(r["id_j"])) in torn]
    assert len(torn) == broken[0] and len(from_bonds) > 0


@pytest.mark.gpu
def test_gpu_glued_block(product_lib, oracle_lib):
    from exadem_b200 import Engine
    handed = []
    wl, (gpu, cpu), broken, _ = run_block([lambda wl: Engine(product_lib, wl.config()), lambda wl: OracleEngine(oracle_lib, wl.config())],
                                          on_break=lambda e: handed.append(e.export_interactions()))
    # the friction handed over from the broken bonds to the vertex-vertex records, right after the rebuild: GPU vs oracle
    from parity import compare_interaction_sets as cmp_sets
    assert len(handed) == 2
    cmp_sets(handed[0], handed[1], payload_rtol=RTOL)
    vv0 = handed[0][handed[0]["type"] == 0]
    assert int((np.abs(vv0["friction"]).sum(axis=1) > 0).sum()) >= 4 * broken[0] > 0
    assert broken[0] == broken[1] and broken[0] > 0
    rec = compare_bonds(gpu, cpu)
    compare_particles(gpu.download_particles(), cpu.download_particles(), rtol=RTOL)
    n = len(rec)
    sg, sc = gpu.get_inner_bond_state(n), cpu.get_inner_bond_state(n)
    og, oc = bond_key(gpu.export_inner_bonds()), bond_key(cpu.export_inner_bonds())
    # dn is the distance of two world vertices minus a constant: its absolute accuracy is an ulp of the coordinates, whatever its size;
    # the bond forces are that distance times the bond stiffness (kn = 2e5, kt = 1e5 in this deck)
    ulp_r = 16 * np.finfo(np.float64).eps * float(np.abs(wl.particles.vec("r")).max() + 1.0)
    ibp = glued_block()[1]
    floors = {"dn": ulp_r, "cp": ulp_r, "fn": ulp_r * ibp[0]["kn"], "ft": ulp_r * ibp[0]["kt"]}
    for a, b, name in zip(sg, sc, ("dn", "cp", "fn", "ft")):
        field_close(a[og], b[oc], RTOL, f"bond state {name}", abs_floor=floors[name])
    # the contact enumeration ignores glued pairs on both sides, and re-admits the torn-off faces after the next rebuild
    from parity import compare_interaction_sets
    assert glued_pairs_in_contact_list(gpu)[0] == 0
    compare_interaction_sets(gpu.export_interactions(), cpu.export_interactions(), payload_rtol=RTOL)
    for e in (gpu, cpu):
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.update_interfaces()
    assert glued_pairs_in_contact_list(gpu) == glued_pairs_in_contact_list(cpu)
    compare_interaction_sets(gpu.export_interactions(), cpu.export_interactions(), payload_rtol=RTOL)
    compare_bonds(gpu, cpu)


@pytest.mark.gpu
def test_gpu_bonds_follow_their_particles_across_a_rebuild(product_lib, oracle_lib):
    """persistent interactions: after a neighbour rebuild re-sorts the particles, edb_update_interfaces re-resolves the bonds by id"""
    from exadem_b200 import Engine
    wl, ibp = glued_block()
    engines = [wl.setup(Engine(product_lib, wl.config())), wl.setup(OracleEngine(oracle_lib, wl.config()))]
    setup_block(engines, wl, ibp)
    for it in range(12):
        for e in engines:
            step_block(e, wl)
        if it == 5:
            for e in engines:
                e.update_particle_neighbors(wl.rcut_inc, True)
                assert e.update_interfaces() > 0
    compare_bonds(*engines)
    compare_particles(engines[0].download_particles(), engines[1].download_particles(), rtol=RTOL)


def two_clusters(wl):
    """left and right halves of the block are different clusters: no bond across x = 1.6"""
    return np.where(wl.particles.rx < 1.6, 1.0, 2.0)


def test_stick_polyhedra_respects_clusters_on_the_oracle():
    wl, ibp = glued_block()
    e = wl.setup(OracleEngine(OracleLibrary(), wl.config()))
    e.set_inner_bond_params(1, ibp)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.stick_polyhedra_ids(2e-5, wl.particles.id, two_clusters(wl))
    b = e.export_inner_bonds()
    assert len(b) == 4 * (104 - 12) and e.update_interfaces() == 92   # the 4 x 3 faces across the cut are not glued


@pytest.mark.gpu
def test_gpu_stick_polyhedra(product_lib, oracle_lib):
    """bond creation on the device (k_stick) against the oracle's restatement of stick_polyhedra.cpp: same bonds, dn0 / weight / criterion"""
    from exadem_b200 import Engine
    wl, ibp = glued_block()
    out = []
    for e in (wl.setup(Engine(product_lib, wl.config())), wl.setup(OracleEngine(oracle_lib, wl.config()))):
        e.set_inner_bond_params(1, ibp)
        e.update_particle_neighbors(wl.rcut_inc, True)
        e.stick_polyhedra_ids(2e-5, wl.particles.id, two_clusters(wl))
        assert e.update_interfaces() == 92
        rec = e.export_inner_bonds()
        out.append(rec[bond_key(rec)])
    a, b = out
    assert len(a) == len(b) == 368
    for k in ("id_i", "id_j", "sub_i", "sub_j", "unbroken", "mode"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("weight", "criteria_1", "criteria_2"):
        assert np.array_equal(a[k], b[k]), k     # same expressions on the same inputs: bit-identical
    assert np.abs(a["dn0"] - b["dn0"]).max() < 1e-12 and np.abs(b["dn0"]).max() < 2e-5


# ---- the reference's traction verification deck (example/polyhedra/fragmentation/verification_test/traction/traction.msp) -------------
def traction_pair(gc, speed=10.0, dt=1e-5):
    """Two cubes of side 0.98 + Minkowski radius 0.01 one lattice step (1.0) apart, bonded at their common face, kn = 1e3, kt = 8e2,
    damp_rate = 0.87, g = Gc (EnergyMixedMode), sticking threshold 1.5e-4 — the traction deck.  The deck freezes cube 1 and imposes the
    velocity of cube 2 with region operators; here both are made so heavy that the forces do not move them (same kinematics), and
    the pull is faster so that the break comes after tens of steps instead of 10^5."""
    wl = workloads.polyhedra_box((2, 1, 1), shapes=("cube",), pitch=1.0, jitter=0.0, periodic=(0, 0, 0), cell_size=2.0)
    wl.shapes_spec = [workloads.cube_shape(half=0.49, mink=0.01)]
    pa = wl.particles
    pa.qw[:], pa.qx[:], pa.qy[:], pa.qz[:] = 1.0, 0.0, 0.0, 0.0
    for f in ("vx", "vy", "vz", "vrotx", "vroty", "vrotz"):
        getattr(pa, f)[:] = 0.0
    pa.radius[:] = np.sqrt(3) * 0.49 + 0.01
    pa.mass[:] = 1e15
    pa.inertiax[:] = pa.inertiay[:] = pa.inertiaz[:] = 1e15
    right = pa.rx > 1.0
    pa.vx[right] = speed
    wl.lo, wl.hi = (-2.0, -2.0, -2.0), (4.0, 4.0, 4.0)
    wl.periodic = (0, 0, 0)
    wl.gravity = (0.0, 0.0, 0.0)
    wl.dt = dt
    wl.params = dict(kn=1e4, kt=8e3, kr=0.0, mu=0.1, damp_rate=0.87, fc=0.0, dncut=0.0, gamma=0.0)
    wl.params_driver = wl.params
    wl.name = "traction"
    return wl, [dict(kn=1e3, kt=8e2, damp_rate=0.87, mode=0, crit1=gc, crit2=0.0)]


@pytest.mark.parametrize("gc", [5e-3, 2e-2, 8e-2])
def test_traction_deck_breaks_at_the_analytic_displacement(gc):
    """Known answer from the source: 4 bonds of weight 1/4 give the interface the stiffness kn, En = kn dn^2 / 2, and the interface
    breaks when En exceeds 2 A Gc with A = 0.98^2 (stick_polyhedra.cpp:290-296, inner_bond_force.hpp:50-68), i.e. at
    dn = sqrt(4 A Gc / kn).  (The table in the deck's README quotes half of that, i.e. an interface stiffness of 4 kn — bond weight 1 —
    which is not what the v1.2.2 source computes; the oracle follows the source.)"""
    wl, ibp = traction_pair(gc)
    e = wl.setup(OracleEngine(OracleLibrary(), wl.config()))
    e.set_inner_bond_params(1, ibp)
    e.update_particle_neighbors(wl.rcut_inc, True)
    e.stick_polyhedra_ids(1.5e-4, wl.particles.id, np.zeros(wl.particles.n))
    b = e.export_inner_bonds()
    assert len(b) == 4 and np.allclose(b["weight"], 0.25) and np.allclose(b["criteria_1"], 2 * 0.98 ** 2 * gc, rtol=1e-12)
    e.update_particle_neighbors(wl.rcut_inc, True)
    assert e.update_interfaces() == 1
    dn_fail = np.sqrt(4 * 0.98 ** 2 * gc / 1e3)
    step_dn = 10.0 * wl.dt
    broke_at = None
    for it in range(int(dn_fail / step_dn) + 10):
        if step_block(e, wl):
            broke_at = it
            break
        dn, _, fn, _ = e.get_inner_bond_state(4)
        # the normal force of the interface grows linearly with the opening (elastic part kn * dn / 4 per bond + the viscous term)
        assert np.allclose(dn, dn[0]) and abs(dn[0] - (it + 1) * step_dn) < 1e-9
    # the criterion is evaluated at the start of step k on the energies of step k-1, where dn = k * step_dn
    assert broke_at is not None and abs(broke_at * step_dn - dn_fail) <= step_dn, (broke_at, dn_fail / step_dn)


def test_run_verlet_with_inner_bonds_on_the_oracle():
    """run_verlet runs the fragmentation macro when bonds exist: identical to the same steps driven operator by operator"""
    from parity import bit_identical
    lib = OracleLibrary()
    wl, ibp = glued_block()
    engines = [wl.setup(OracleEngine(lib, wl.config())) for _ in range(2)]
    setup_block(engines, wl, ibp)
    a, ops = engines
    nb = a.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 20, -1)
    broken = sum(step_block(ops, wl) for _ in range(20))
    assert nb == 1 and broken == 16
    assert all(bit_identical(a.download_particles(), ops.download_particles()).values())
    compare_bonds(a, ops, rtol=0.0)


def block_hit_by_a_ball():
    """the glued block at rest, a ball driver (LINEAR_MOTION) ramming its top layer from the side: moving drivers and inner bonds in one
    run — the deck family of example/polyhedra/fragmentation (a bonded sample loaded by a moving driver)"""
    wl, ibp = glued_block()
    wl.particles.vx[:] = 0.0
    wl.particles.vz[:] = 0.0
    wl.drivers = list(wl.drivers) + [dict(kind=capi.DRIVER_BALL, radius=0.6, center=(-0.6, 1.6, 2.0), vel=(30.0, 0, 0))]
    wl.motions = [dict(motion_type="STATIONARY")] * (len(wl.drivers) - 1) + [
        dict(motion_type="LINEAR_MOTION", motion_vector=(1, 0, 0), const_vel=30.0)]
    ibp[0]["crit1"] = 0.02
    wl.name = "block_hit_by_a_ball"
    return wl, ibp


def test_bonds_and_moving_driver_on_the_oracle():
    lib = OracleLibrary()
    wl, ibp = block_hit_by_a_ball()
    e = wl.setup(OracleEngine(lib, wl.config()))
    setup_block([e], wl, ibp)
    nb = e.run_verlet(wl.dt, wl.gravity, wl.rcut_inc, True, 40, 0)
    b = e.export_inner_bonds()
    cnt, act = e.interaction_counts()
    ball = e.get_drivers(len(wl.drivers))[-1]
    assert nb >= 3 and 0 < len(b) < 416 and b["unbroken"].all()          # interfaces broke one after the other, each forcing a rebuild
    assert act[6] > 0 and abs(ball["center"][0] - (-0.6 + 40 * 30.0 * wl.dt)) < 1e-12
