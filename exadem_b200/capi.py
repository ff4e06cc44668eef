"""ctypes binding of the C ABI declared in include/exadem_b200.h.

This is the Python view of the drop-in boundary (tests, bench.py and examples drive the library through it).
It holds NO compute: every method forwards to `libexadem_b200.so` (CUDA, sm_100a).  If the shared library is
missing the import of :class:`Library` raises — there is no CPU fallback on the product path.

The struct definitions and the :class:`Engine` wrapper are prefix-agnostic, so the test suite can drive its CPU
checker through the same scripting surface (the checker itself lives outside this package).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

NTYPES = 14
MODE_SPHERE, MODE_POLYHEDRON = 0, 1
LAW_NONE, LAW_COHESIVE, LAW_DMT = 0, 1, 2
DRIVER_CYLINDER, DRIVER_SURFACE, DRIVER_BALL, DRIVER_RSHAPE, DRIVER_UNDEFINED = 0, 1, 2, 3, 4

_D = C.POINTER(C.c_double)
_U64 = C.POINTER(C.c_uint64)
_U32 = C.POINTER(C.c_uint32)


DYN_ALL, DYN_STATE = 0x7F, 0x0F   # field masks of edb_download_dynamic_async


class Config(C.Structure):
    _fields_ = [("lo", C.c_double * 3), ("hi", C.c_double * 3), ("glo", C.c_double * 3), ("ghi", C.c_double * 3),
                ("cell_size", C.c_double), ("periodic", C.c_int32 * 3), ("mode", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("pgrid", C.c_int32 * 3),
                ("emit_contact_state", C.c_int32), ("reserved", C.c_int32)]


F64_FIELDS = ["rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz", "mass", "homothety", "radius",
              "qw", "qx", "qy", "qz", "momx", "momy", "momz", "vrotx", "vroty", "vrotz",
              "arotx", "aroty", "arotz", "inertiax", "inertiay", "inertiaz"]


class Particles(C.Structure):
    _fields_ = ([("n", C.c_uint64)] + [(f, _D) for f in F64_FIELDS] + [("id", _U64), ("type", _U32), ("group", _U32)])


class ContactParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("kn", "kt", "kr", "mu", "damp_rate", "fc", "dncut", "gamma")]


class Driver(C.Structure):
    _fields_ = [("kind", C.c_int32), ("need_forces", C.c_int32), ("radius", C.c_double), ("axis", C.c_double * 3),
                ("center", C.c_double * 3), ("offset", C.c_double), ("normal", C.c_double * 3),
                ("vel", C.c_double * 3), ("vrot", C.c_double * 3), ("shape", C.c_int32), ("reserved", C.c_int32),
                ("quat", C.c_double * 4)]


_V3 = C.c_double * 3
MOTION = {n: i for i, n in enumerate(("STATIONARY", "LINEAR_MOTION", "COMPRESSIVE_FORCE", "LINEAR_FORCE_MOTION", "PARTICLE",
                                      "LINEAR_COMPRESSIVE_MOTION", "TABULATED", "SHAKER", "PENDULUM_MOTION", "EXPRESSION", "UNKNOWN"))}


class DriverMotion(C.Structure):
    """edb_driver_motion (include/exadem_b200.h): Driver_params + integrator state of one driver."""
    _fields_ = [("motion_type", C.c_int32), ("n_tab", C.c_int32), ("motion_vector", _V3),
                ("motion_start_threshold", C.c_double), ("motion_end_threshold", C.c_double),
                ("const_vel", C.c_double), ("const_force", C.c_double), ("sigma", C.c_double), ("damprate", C.c_double),
                ("tab_time", _D), ("tab_pos", _D), ("omega", C.c_double), ("amplitude", C.c_double), ("shaker_dir", _V3),
                ("pendulum_anchor_point", _V3), ("pendulum_initial_position", _V3), ("pendulum_swing_dir", _V3),
                ("expr_use_v", C.c_int32), ("expr_use_vrot", C.c_int32), ("expr_use_mom", C.c_int32), ("drive_by_mom", C.c_int32),
                ("mass", C.c_double), ("surface", C.c_double), ("acc", _V3), ("rv", C.c_double), ("ra", C.c_double),
                ("inertia", _V3), ("applied_mom", _V3), ("mom_axis", _V3), ("mom", _V3), ("forces", _V3)]


def make_motion(d: dict, keep: list) -> DriverMotion:
    """dict -> DriverMotion with the reference's defaults (driver_params.hpp:33-66, *Fields structs)."""
    M = DriverMotion()
    mt = d.get("motion_type", "STATIONARY")
    M.motion_type = MOTION[mt] if isinstance(mt, str) else int(mt)
    M.motion_start_threshold = d.get("motion_start_threshold", 0.0)
    M.motion_end_threshold = d.get("motion_end_threshold", 1e300)
    for k in ("const_vel", "const_force", "sigma", "damprate", "omega", "amplitude", "rv", "ra"):
        setattr(M, k, float(d.get(k, 0.0)))
    M.mass = d.get("mass", 1.7976931348623157e308 / 4)
    M.surface = d.get("surface", -1.0)
    for k, dflt in (("motion_vector", (0, 0, 0)), ("shaker_dir", (0, 0, 1)), ("pendulum_anchor_point", (0, 0, 0)),
                    ("pendulum_initial_position", (0, 0, 0)), ("pendulum_swing_dir", (0, 0, 0)), ("acc", (0, 0, 0)),
                    ("inertia", (0, 0, 0)), ("applied_mom", (0, 0, 0)), ("mom_axis", (0, 0, 0)), ("mom", (0, 0, 0)),
                    ("forces", (0, 0, 0))):
        v = d.get(k, dflt)
        for c in range(3):
            getattr(M, k)[c] = float(v[c])
    for k in ("expr_use_v", "expr_use_vrot", "expr_use_mom", "drive_by_mom"):
        setattr(M, k, int(d.get(k, 0)))
    if "tab_time" in d:
        t = np.ascontiguousarray(d["tab_time"], dtype=np.float64)
        x = np.ascontiguousarray(d["tab_pos"], dtype=np.float64).reshape(-1, 3)
        assert len(t) == len(x)
        keep += [t, x]
        M.n_tab = len(t)
        M.tab_time = t.ctypes.data_as(_D)
        M.tab_pos = x.ctypes.data_as(_D)
    return M


def driver_to_dict(D, M=None) -> dict:
    out = {"kind": D.kind, "need_forces": D.need_forces, "radius": D.radius, "offset": D.offset, "shape": D.shape}
    for k in ("axis", "center", "normal", "vel", "vrot"):
        out[k] = np.array(getattr(D, k)[:])
    out["quat"] = np.array(D.quat[:])
    if M is not None:
        out["motion_type"] = M.motion_type
        for k in ("mass", "surface", "rv", "ra"):
            out[k] = getattr(M, k)
        for k in ("acc", "inertia", "applied_mom", "mom_axis", "mom", "forces"):
            out[k] = np.array(getattr(M, k)[:])
    return out


class InnerBondParams(C.Structure):
    _fields_ = [("kn", C.c_double), ("kt", C.c_double), ("damp_rate", C.c_double), ("mode", C.c_int32), ("pad", C.c_int32),
                ("crit1", C.c_double), ("crit2", C.c_double)]


INNER_BOND_DTYPE = np.dtype([("id_i", "<u8"), ("id_j", "<u8"), ("sub_i", "<u4"), ("sub_j", "<u4"), ("friction", "<f8", 3), ("en", "<f8"),
                             ("tds", "<f8", 3), ("et", "<f8"), ("dn0", "<f8"), ("weight", "<f8"), ("criteria_1", "<f8"),
                             ("criteria_2", "<f8"), ("mode", "<i4"), ("unbroken", "u1"), ("ghost", "u1"), ("pad", "<u2")])
assert INNER_BOND_DTYPE.itemsize == 128


class Shape(C.Structure):
    _fields_ = [("nv", C.c_int32), ("ne", C.c_int32), ("nf", C.c_int32), ("faces_len", C.c_int32),
                ("vertices", _D), ("edges", C.POINTER(C.c_int32)), ("faces", C.POINTER(C.c_int32)),
                ("obb", C.c_double * 15), ("radius", C.c_double)]


class Region(C.Structure):
    _fields_ = [("lo", C.c_double * 3), ("hi", C.c_double * 3), ("quadric", C.c_double * 16), ("id_start", C.c_uint64), ("id_end", C.c_uint64),
                ("use_bounds", C.c_int32), ("use_ids", C.c_int32), ("use_quadric", C.c_int32), ("pad", C.c_int32)]


CSG_AND, CSG_OR, CSG_NOT = -1, -2, -3


class Interaction(C.Structure):
    _fields_ = [("id_i", C.c_uint64), ("cell_i", C.c_uint32), ("p_i", C.c_uint16), ("pad_i", C.c_uint16),
                ("sub_i", C.c_uint32), ("pad_i2", C.c_uint32),
                ("id_j", C.c_uint64), ("cell_j", C.c_uint32), ("p_j", C.c_uint16), ("pad_j", C.c_uint16),
                ("sub_j", C.c_uint32), ("pad_j2", C.c_uint32),
                ("type", C.c_uint16), ("swap", C.c_uint8), ("ghost", C.c_uint8), ("pad_t", C.c_uint32),
                ("friction", C.c_double * 3), ("moment", C.c_double * 3)]


INTERACTION_DTYPE = np.dtype([("id_i", "<u8"), ("cell_i", "<u4"), ("p_i", "<u2"), ("pad_i", "<u2"), ("sub_i", "<u4"),
                              ("pad_i2", "<u4"), ("id_j", "<u8"), ("cell_j", "<u4"), ("p_j", "<u2"), ("pad_j", "<u2"),
                              ("sub_j", "<u4"), ("pad_j2", "<u4"), ("type", "<u2"), ("swap", "u1"), ("ghost", "u1"),
                              ("pad_t", "<u4"), ("friction", "<f8", 3), ("moment", "<f8", 3)])
assert INTERACTION_DTYPE.itemsize == C.sizeof(Interaction) == 104


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("interaction_updates", C.c_uint64), ("particle_steps", C.c_uint64),
                ("rebuilds", C.c_uint64), ("contact_ms", C.c_double), ("gather_ms", C.c_double),
                ("prolog_ms", C.c_double), ("epilog_ms", C.c_double), ("rebuild_ms", C.c_double),
                ("ghost_ms", C.c_double), ("contact_launches", C.c_uint64), ("bytes_contact_algorithmic", C.c_uint64),
                ("contact_pp_ms", C.c_double), ("contact_pp_launches", C.c_uint64), ("contact_pp_records", C.c_uint64),
                ("gather_launches", C.c_uint64)]


class SimState(C.Structure):
    _fields_ = [("kinetic_energy", C.c_double * 3), ("rotation_energy", C.c_double * 3), ("mass", C.c_double), ("min_dn", C.c_double),
                ("n_particles", C.c_uint64), ("n_active_interactions", C.c_uint64), ("n_total_interactions", C.c_uint64)]


class EdbError(RuntimeError):
    pass


def repo_root() -> Path:
    return Path(__file__).resolve().parent.parent


def product_library_path() -> Path:
    return Path(__file__).resolve().parent / "libexadem_b200.so"


class Library:
    """Loads libexadem_b200.so (the CUDA product).  Raises if it is not built — no fallback."""

    prefix = "edb_"

    def __init__(self, path: os.PathLike | None = None):
        p = Path(path) if path else product_library_path()
        if not p.exists():
            raise EdbError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback for the product path.")
        self.path = p
        self.dll = C.CDLL(str(p), mode=C.RTLD_GLOBAL)

    def fn(self, name):
        f = getattr(self.dll, self.prefix + name)
        f.restype = C.c_int
        return f


def _as(a, ct):
    return a.ctypes.data_as(ct)


class ParticleArrays:
    """Host SoA mirror of DEMFieldSet (include/exaDEM/fields.h:34-54): one contiguous numpy array per field."""

    def __init__(self, n: int):
        self.n = int(n)
        for f in F64_FIELDS:
            setattr(self, f, np.zeros(n, dtype=np.float64))
        self.mass[:] = 1.0
        self.homothety[:] = 1.0
        self.radius[:] = 0.5
        self.qw[:] = 1.0
        self.inertiax[:] = 1.0
        self.inertiay[:] = 1.0
        self.inertiaz[:] = 1.0
        self.id = np.arange(n, dtype=np.uint64)
        self.type = np.zeros(n, dtype=np.uint32)
        self.group = np.zeros(n, dtype=np.uint32)

    def view(self, fields=None) -> Particles:
        P = Particles()
        P.n = self.n
        for f in F64_FIELDS:
            if fields is None or f in fields:
                a = getattr(self, f)
                assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (self.n,)
                setattr(P, f, _as(a, _D))
        if fields is None or "id" in fields:
            P.id = _as(self.id, _U64)
        if fields is None or "type" in fields:
            P.type = _as(self.type, _U32)
        if fields is None or "group" in fields:
            P.group = _as(self.group, _U32)
        return P

    def copy(self) -> "ParticleArrays":
        o = ParticleArrays(self.n)
        for f in F64_FIELDS + ["id", "type", "group"]:
            setattr(o, f, getattr(self, f).copy())
        return o

    def vec(self, name) -> np.ndarray:
        comps = {"r": ("rx", "ry", "rz"), "v": ("vx", "vy", "vz"), "f": ("fx", "fy", "fz"),
                 "mom": ("momx", "momy", "momz"), "vrot": ("vrotx", "vroty", "vrotz"),
                 "arot": ("arotx", "aroty", "arotz"), "q": ("qw", "qx", "qy", "qz")}[name]
        return np.stack([getattr(self, c) for c in comps], axis=1)


def make_shape(vertices, edges, faces, obb, radius):
    """Build a :class:`Shape` (keeps the numpy buffers alive on the returned object)."""
    v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
    stream = []
    for f in faces:
        stream.append(len(f))
        stream.extend(int(x) for x in f)
    fs = np.asarray(stream, dtype=np.int32)
    s = Shape()
    s.nv, s.ne, s.nf, s.faces_len = len(v), len(e), len(faces), len(fs)
    s.vertices = _as(v, _D)
    s.edges = _as(e, C.POINTER(C.c_int32))
    s.faces = _as(fs, C.POINTER(C.c_int32))
    for k in range(15):
        s.obb[k] = float(obb[k])
    s.radius = float(radius)
    s._keep = (v, e, fs)
    return s


class Engine:
    """Thin object wrapper over one context of the C ABI.  Method names = C function names without the prefix;
    they mirror the reference operator names (nbh/contact/prolog/epilog ...)."""

    def __init__(self, lib, config: Config):
        self.lib = lib
        self.cfg = config
        self.h = C.c_void_p()
        rc = lib.fn("create")(C.byref(config), C.byref(self.h))
        if rc != 0:
            raise EdbError(f"create failed rc={rc}: {self._err(None)}")

    # -- plumbing
    def _err(self, h):
        try:
            f = getattr(self.lib.dll, self.lib.prefix + "last_error")
        except AttributeError:
            return ""
        f.restype = C.c_char_p
        s = f(h)
        return s.decode() if s else ""

    def _call(self, name, *args):
        rc = self.lib.fn(name)(self.h, *args)
        if rc != 0:
            raise EdbError(f"{self.lib.prefix}{name} failed rc={rc}: {self._err(self.h)}")

    def close(self):
        if self.h:
            self.lib.fn("destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters
    def set_contact_params(self, particle: dict, driver: dict | None = None):
        a = ContactParams(**particle)
        b = ContactParams(**(driver if driver is not None else particle))
        self._call("set_contact_params", C.byref(a), C.byref(b))

    def set_contact_params_multimat(self, n_groups, pp, n_drivers, pd):
        A = (ContactParams * (n_groups * n_groups))(*[ContactParams(**d) for d in pp])
        B = (ContactParams * max(1, n_groups * n_drivers))(*[ContactParams(**d) for d in pd])
        self._call("set_contact_params_multimat", C.c_int32(n_groups), A, C.c_int32(n_drivers), B)

    def set_law(self, law):
        self._call("set_law", C.c_int32(law))

    def set_drivers(self, drivers):
        arr = (Driver * max(1, len(drivers)))()
        for k, d in enumerate(drivers):
            D = Driver()
            D.kind = d["kind"]
            D.need_forces = d.get("need_forces", 0)
            D.radius = d.get("radius", 0.0)
            for name, dflt in (("axis", (1, 0, 1)), ("center", (0, 0, 0)), ("normal", (0, 0, 1)), ("vel", (0, 0, 0)),
                               ("vrot", (0, 0, 0))):
                val = d.get(name, dflt)
                for c in range(3):
                    getattr(D, name)[c] = float(val[c])
            D.offset = d.get("offset", 0.0)
            D.shape = d.get("shape", -1)
            for c, qv in enumerate(d.get("quat", (1.0, 0.0, 0.0, 0.0))):
                D.quat[c] = float(qv)
            arr[k] = D
        self._call("set_drivers", C.c_int32(len(drivers)), arr)

    def set_driver_motions(self, motions):
        """motions: one dict per driver (keys of edb_driver_motion; motion_type by name)."""
        self._keep_motion = []
        arr = (DriverMotion * max(1, len(motions)))(*[make_motion(m, self._keep_motion) for m in motions])
        self._call("set_driver_motions", C.c_int32(len(motions)), arr)

    def set_driver_expr(self, driver, v=None, vrot=None, mom=None):
        p = [(_V3)(*map(float, x)) if x is not None else None for x in (v, vrot, mom)]
        self._call("set_driver_expr", C.c_int32(driver), *p)

    def get_drivers(self, n):
        D, M = (Driver * max(1, n))(), (DriverMotion * max(1, n))()
        self._call("get_drivers", C.c_int32(n), D, M)
        return [driver_to_dict(D[i], M[i]) for i in range(n)]

    # -- driver operators of the time loop (config_numerical_schemes.msp:13-24)
    def push_f_v_r_driver(self, time, dt):
        self._call("push_f_v_r_driver", C.c_double(time), C.c_double(dt))

    def push_f_v_driver(self, dt):
        self._call("push_f_v_driver", C.c_double(dt))

    def push_av_to_quat_driver(self, time, dt):
        self._call("push_av_to_quat_driver", C.c_double(time), C.c_double(dt))

    def reset_force_driver(self):
        self._call("reset_force_driver")

    def force_to_accel_driver(self, dt, system_mass):
        self._call("force_to_accel_driver", C.c_double(dt), C.c_double(system_mass))

    def driver_vertices(self):
        self._call("driver_vertices")

    def grid_rshape_driver(self):
        self._call("grid_rshape_driver")

    def backup_drivers(self):
        self._call("backup_drivers")

    def driver_displ_over(self, threshold):
        r = C.c_uint64()
        self._call("driver_displ_over", C.c_double(threshold), C.byref(r))
        return r.value

    def set_physical_time(self, t):
        self._call("set_physical_time", C.c_double(t))

    def get_physical_time(self):
        t = C.c_double()
        self._call("get_physical_time", C.byref(t))
        return t.value

    def set_system_mass(self, m):
        self._call("set_system_mass", C.c_double(m))

    # -- inner bonds (wave 13)
    def set_inner_bond_params(self, n_groups, params):
        arr = (InnerBondParams * (n_groups * n_groups))(*[InnerBondParams(**p) for p in params])
        self._call("set_inner_bond_params", C.c_int32(n_groups), arr)

    def import_inner_bonds(self, rec: np.ndarray):
        rec = np.ascontiguousarray(rec, dtype=INNER_BOND_DTYPE)
        self._call("import_inner_bonds", rec.ctypes.data_as(C.c_void_p), C.c_uint64(len(rec)))

    def stick_polyhedra_ids(self, threshold, ids, cluster):
        """[stick_polyhedra] on the device; cluster given per particle id."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        cl = np.ascontiguousarray(cluster, dtype=np.float64)
        self._call("stick_polyhedra", C.c_double(threshold), C.c_uint64(len(ids)), ids.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p))

    def export_inner_bonds(self) -> np.ndarray:
        n = C.c_uint64()
        self._call("export_inner_bonds", C.c_void_p(), C.c_uint64(0), C.byref(n))
        rec = np.zeros(n.value, dtype=INNER_BOND_DTYPE)
        if n.value:
            self._call("export_inner_bonds", rec.ctypes.data_as(C.c_void_p), C.c_uint64(n.value), C.byref(n))
        return rec

    def update_interfaces(self):
        n = C.c_uint64()
        self._call("update_interfaces", C.byref(n))
        return n.value

    def inner_bond_polyhedron(self, dt):
        self._call("inner_bond_polyhedron", C.c_double(dt))

    def apply_interface_fracture_criterion(self):
        n = C.c_uint64()
        self._call("apply_interface_fracture_criterion", C.byref(n))
        return n.value

    def broken_interface_friction_updater(self):
        self._call("broken_interface_friction_updater")

    def get_inner_bond_state(self, n):
        dn, cp, fn, ft = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        p = lambda a: a.ctypes.data_as(_D)
        self._call("get_inner_bond_state", p(dn), p(cp), p(fn), p(ft))
        return dn, cp, fn, ft

    def set_shapes(self, shapes):
        arr = (Shape * len(shapes))(*shapes)
        self._keep_shapes = shapes
        self._call("set_shapes", C.c_int32(len(shapes)), arr)

    def set_driver_shapes(self, shapes):
        """Mesh shapes of the RShape drivers (call before set_drivers)."""
        arr = (Shape * max(1, len(shapes)))(*shapes)
        self._keep_driver_shapes = shapes
        self._call("set_driver_shapes", C.c_int32(len(shapes)), arr)

    # -- grid
    def upload_particles(self, pa: ParticleArrays):
        v = pa.view()
        self._call("upload_particles", C.byref(v))

    def download_particles(self, pa: ParticleArrays | None = None) -> ParticleArrays:
        nr, _ = self.num_particles()
        if pa is None:
            pa = ParticleArrays(nr)
        assert pa.n == nr
        v = pa.view()
        self._call("download_particles", C.byref(v))
        return pa

    def upload_dynamic(self, pa: ParticleArrays):
        v = pa.view()
        self._call("upload_dynamic", C.byref(v))

    def download_dynamic(self, pa: ParticleArrays):
        v = pa.view()
        self._call("download_dynamic", C.byref(v))

    def upload_dynamic_async(self, pa: ParticleArrays, view=None):
        """pinned host arrays; returns once the copies are enqueued (see include/exadem_b200.h).  `view`: a pa.view() made once by a
        caller that steps in a loop (building the 31-pointer struct costs more host time than the call)"""
        v = view if view is not None else pa.view()
        self._call("upload_dynamic_async", C.byref(v))

    def download_dynamic_async(self, pa: ParticleArrays, field_mask: int = DYN_ALL, view=None):
        v = view if view is not None else pa.view()
        self._call("download_dynamic_async", C.byref(v), C.c_uint32(field_mask))

    def num_particles(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._call("num_particles", C.byref(a), C.byref(b))
        return a.value, b.value

    # -- operators
    def update_particle_neighbors(self, rcut_inc, symetric=True):
        self._call("update_particle_neighbors", C.c_double(rcut_inc), C.c_int32(1 if symetric else 0))

    def ghost_update_rq(self):
        self._call("ghost_update_rq")

    def compute_vertices(self):
        self._call("compute_vertices")

    def reset_force_moment(self):
        self._call("reset_force_moment")

    def gravity_force(self, g=(0.0, 0.0, -9.807)):
        self._call("gravity_force", (C.c_double * 3)(*g))

    def contact_sphere(self, dt):
        self._call("contact_sphere", C.c_double(dt))

    def contact_polyhedron(self, dt):
        self._call("contact_polyhedron", C.c_double(dt))

    def contact(self, dt):
        (self.contact_polyhedron if self.cfg.mode == MODE_POLYHEDRON else self.contact_sphere)(dt)

    def force_to_accel(self):
        self._call("force_to_accel")

    def combined_compute_prolog(self, dt):
        self._call("combined_compute_prolog", C.c_double(dt))

    def combined_compute_epilog(self, dt):
        self._call("combined_compute_epilog", C.c_double(dt))

    def push_to_quaternion(self, dt):
        self._call("push_to_quaternion", C.c_double(dt))

    def push_to_angular_velocity(self, dt):
        self._call("push_to_angular_velocity", C.c_double(dt))

    def push_to_angular_acceleration(self):
        self._call("push_to_angular_acceleration")

    def freeze_particles(self):
        self._call("freeze_particles")

    def set_particle_regions(self, regions):
        """regions: list of dicts with optional keys bounds=(lo, hi), ids=(start, end), quadric=4x4 (row-major)"""
        arr = (Region * max(len(regions), 1))()
        for i, r in enumerate(regions):
            R = arr[i]
            if "bounds" in r:
                R.use_bounds = 1
                for k in range(3):
                    R.lo[k], R.hi[k] = float(r["bounds"][0][k]), float(r["bounds"][1][k])
            if "ids" in r:
                R.use_ids = 1
                R.id_start, R.id_end = int(r["ids"][0]), int(r["ids"][1])
            if "quadric" in r:
                R.use_quadric = 1
                q = np.asarray(r["quadric"], dtype=np.float64).reshape(16)
                for k in range(16):
                    R.quadric[k] = float(q[k])
        self._call("set_particle_regions", C.c_int32(len(regions)), arr)

    def _expr(self, expr):
        e = (C.c_int32 * max(len(expr), 1))(*expr)
        return C.c_int32(len(expr)), e

    def reset_force_moment_region(self, expr):
        self._call("reset_force_moment_region", *self._expr(expr))

    def freeze_particles_region(self, expr):
        self._call("freeze_particles_region", *self._expr(expr))

    def displ_over(self, threshold) -> int:
        n = C.c_uint64()
        self._call("displ_over", C.c_double(threshold), C.byref(n))
        return n.value

    def compute_driver_force_moment(self, n_drivers):
        f = np.zeros((n_drivers, 3))
        m = np.zeros((n_drivers, 3))
        self._call("compute_driver_force_moment", _as(f, _D), _as(m, _D))
        return f, m

    def simulation_state(self, symetric=True) -> SimState:
        """[simulation_state] energies, mass, interaction counts and min dn of this rank (the log line's reductions)."""
        out = SimState()
        self._call("simulation_state", C.c_int32(1 if symetric else 0), C.byref(out))
        return out

    def set_brick_cuts(self, cuts):
        """Cost-balanced decomposition: cuts[d] = pgrid[d] + 1 increasing positions (None = uniform along d).  Before comm_init / upload."""
        arr = [None if c is None else (C.c_double * len(c))(*[float(x) for x in c]) for c in cuts]
        self._call("set_brick_cuts", *arr)

    def dem_cost_model(self):
        """[dem_cost_model] cost per real cell of this rank's grid, shape (nz, ny, nx) (x fastest like the reference's CellCosts)."""
        dims = (C.c_int32 * 3)()
        self._call("dem_cost_model", dims, None, C.c_uint64(0))
        nx, ny, nz = (int(d) for d in dims)
        out = np.zeros((nz, ny, nx))
        self._call("dem_cost_model", dims, _as(out, _D), C.c_uint64(out.size))
        return out

    def stress_tensor(self):
        """[stress_tensor] per-particle 3x3 sum of f (x) (cp - r) over the contacts of the last step, upload order."""
        nr = C.c_uint64()
        self.lib.fn("num_particles")(self.h, C.byref(nr), None)
        out = np.zeros((nr.value, 3, 3))
        self._call("stress_tensor", _as(out, _D))
        return out

    def run_verlet(self, dt, g, rcut_inc, symetric, nsteps, rebuild_every=0) -> int:
        nb = C.c_uint64()
        self._call("run_verlet", C.c_double(dt), (C.c_double * 3)(*g), C.c_double(rcut_inc),
                   C.c_int32(1 if symetric else 0), C.c_int32(nsteps), C.c_int32(rebuild_every), C.byref(nb))
        return nb.value

    # -- views
    def interaction_counts(self):
        c = (C.c_uint64 * NTYPES)()
        a = (C.c_uint64 * NTYPES)()
        self._call("interaction_counts", c, a)
        return np.array(c[:], dtype=np.uint64), np.array(a[:], dtype=np.uint64)

    def export_interactions(self) -> np.ndarray:
        counts, _ = self.interaction_counts()
        n = int(counts.sum())
        out = np.zeros(max(n, 1), dtype=INTERACTION_DTYPE)
        got = C.c_uint64()
        self._call("export_interactions", out.ctypes.data_as(C.POINTER(Interaction)), C.c_uint64(n), C.byref(got))
        assert got.value == n, (got.value, n)
        return out[:n]

    def import_interactions(self, rec: np.ndarray):
        rec = np.ascontiguousarray(rec, dtype=INTERACTION_DTYPE)
        self._call("import_interactions", rec.ctypes.data_as(C.POINTER(Interaction)), C.c_uint64(len(rec)))

    def get_contact_state(self, t):
        counts, _ = self.interaction_counts()
        n = int(counts[t])
        dn = np.zeros(max(n, 1))
        cp, fn, ft = (np.zeros((max(n, 1), 3)) for _ in range(3))
        self._call("get_contact_state", C.c_int32(t), _as(dn, _D), _as(cp, _D), _as(fn, _D), _as(ft, _D))
        return dn[:n], cp[:n], fn[:n], ft[:n]

    def counters(self) -> Counters:
        c = Counters()
        self._call("get_counters", C.byref(c))
        return c

    def reset_counters(self):
        self._call("reset_counters")

    def set_timing(self, on):
        self._call("set_timing", C.c_int32(1 if on else 0))

    def sync(self):
        self._call("sync")

    def flush_l2(self):
        self._call("flush_l2")

    def event_record(self, slot):
        self._call("event_record", C.c_int32(slot))

    def event_elapsed_ms(self, a, b) -> float:
        ms = C.c_double()
        self._call("event_elapsed_ms", C.c_int32(a), C.c_int32(b), C.byref(ms))
        return ms.value

    def comm_init(self, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._call("comm_init", buf)


def comm_unique_id(lib) -> bytes:
    buf = (C.c_uint8 * 128)()
    f = getattr(lib.dll, lib.prefix + "comm_unique_id")
    f.restype = C.c_int
    rc = f(buf)
    if rc != 0:
        raise EdbError(f"comm_unique_id rc={rc}")
    return bytes(buf)


def make_config(lo, hi, cell_size, periodic=(0, 0, 0), mode=MODE_SPHERE, device=0, glo=None, ghi=None, rank=0,
                nranks=1, pgrid=(1, 1, 1), emit_contact_state=1) -> Config:
    c = Config()
    glo = lo if glo is None else glo
    ghi = hi if ghi is None else ghi
    for k in range(3):
        c.lo[k], c.hi[k], c.glo[k], c.ghi[k] = float(lo[k]), float(hi[k]), float(glo[k]), float(ghi[k])
        c.periodic[k] = int(periodic[k])
        c.pgrid[k] = int(pgrid[k])
    c.cell_size = float(cell_size)
    c.mode = mode
    c.device = device
    c.rank, c.nranks = rank, nranks
    c.emit_contact_state = emit_contact_state
    return c
