"""Seeded synthetic inputs for the BASELINE.json configs.  Host-side numpy only.

Every workload returns a :class:`Workload` with the slots the reference decks provide: domain (bounds, cell_size,
periodic), the particle field set, drivers, contact parameters, dt, rcut_inc, gravity.  Parameters follow the
reference decks cited per function (paths relative to /root/reference); the particle data are synthetic.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import ParticleArrays

PI = 4.0 * np.arctan(1.0)  # src/field/set_fields.cu:323-327 uses 4*atan(1)


@dataclass
class Workload:
    name: str
    lo: tuple
    hi: tuple
    cell_size: float
    periodic: tuple
    particles: ParticleArrays
    params: dict
    params_driver: dict | None = None
    drivers: list = field(default_factory=list)
    shapes_spec: list = field(default_factory=list)  # list of dicts(vertices, edges, faces, obb, radius)
    mesh_spec: list = field(default_factory=list)    # shapes of the RShape (mesh) drivers, same dict layout
    motions: list = field(default_factory=list)      # one dict per driver (edb_driver_motion keys); empty = all STATIONARY
    mode: int = capi.MODE_SPHERE
    law: int = capi.LAW_NONE
    dt: float = 1e-4
    rcut_inc: float = 0.4
    gravity: tuple = (0.0, 0.0, -9.807)
    symetric: bool = True

    def effective_cell_size(self):
        """cell_size, shrunk so that every periodic axis has >= 3 cells (a 1-cell ghost layer must not alias the
        images), but never below the grid constraint 2*r_max + rcut_inc (src/field/check_rcut.cpp:72,88-116)."""
        cs = self.cell_size
        for d in range(3):
            if self.periodic[d]:
                cs = min(cs, (self.hi[d] - self.lo[d]) / 3.0)
        need = 2.0 * float(self.particles.radius.max()) + self.rcut_inc
        if cs < need:
            raise ValueError(f"periodic box too small: cell {cs} < 2*rmax+rcut_inc = {need}")
        return cs

    def config(self, device=0, emit_contact_state=1, **kw):
        return capi.make_config(self.lo, self.hi, self.effective_cell_size(), self.periodic, self.mode, device,
                                emit_contact_state=emit_contact_state, **kw)

    def shapes(self):
        return [capi.make_shape(s["vertices"], s["edges"], s["faces"], s["obb"], s["radius"]) for s in self.shapes_spec]

    def setup(self, engine):
        """Push every slot into an Engine (product or oracle)."""
        engine.set_law(self.law)
        engine.set_contact_params(self.params, self.params_driver)
        if self.shapes_spec:
            engine.set_shapes(self.shapes())
        if self.mesh_spec:
            engine.set_driver_shapes([capi.make_shape(s["vertices"], s["edges"], s["faces"], s["obb"], s["radius"]) for s in self.mesh_spec])
        engine.set_drivers(self.drivers)
        if self.motions:
            engine.set_driver_motions(self.motions)
        engine.upload_particles(self.particles)
        return engine


def sphere_mass_inertia(pa: ParticleArrays, density: float):
    r = pa.radius
    pa.mass[:] = density * (4.0 / 3.0) * PI * r * r * r
    i = 0.4 * pa.mass * r * r
    pa.inertiax[:] = i
    pa.inertiay[:] = i
    pa.inertiaz[:] = i


def _lattice(lo, hi, pitch):
    ax = [np.arange(lo[d] + 0.5 * pitch, hi[d], pitch) for d in range(3)]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3)
    return g


def spheres_box(n_side=(10, 10, 10), seed=0xD3A0, jitter=2e-5, radius=0.5, sigma_v=0.1, periodic=(1, 1, 1),
                cell_size=2.5, polydisperse=0.0, omega_sigma=0.0, name="box") -> Workload:
    """S5-like: monodisperse spheres on an SC lattice (pitch 1.0) with uniform jitter in a (periodic) box;
    law parameters of example/spheres/rotating-drum/rotating-drum-mini.msp:25-29."""
    rng = np.random.default_rng(seed)
    hi = tuple(float(n) for n in n_side)
    pos = _lattice((0, 0, 0), hi, 1.0)
    n = len(pos)
    pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    pa = ParticleArrays(n)
    pa.rx[:], pa.ry[:], pa.rz[:] = pos[:, 0], pos[:, 1], pos[:, 2]
    v = rng.normal(0.0, sigma_v, size=(n, 3))
    pa.vx[:], pa.vy[:], pa.vz[:] = v[:, 0], v[:, 1], v[:, 2]
    if omega_sigma > 0:
        w = rng.normal(0.0, omega_sigma, size=(n, 3))
        pa.vrotx[:], pa.vroty[:], pa.vrotz[:] = w[:, 0], w[:, 1], w[:, 2]
    pa.radius[:] = radius * (1.0 - polydisperse * rng.uniform(0, 1, n))
    sphere_mass_inertia(pa, 0.02)
    pa.id[:] = rng.permutation(n).astype(np.uint64)  # ids uncorrelated with position (id order decides i/j roles)
    prm = dict(kn=1e5, kt=1e5, kr=0.1, mu=0.5, damp_rate=0.9, fc=0.0, dncut=0.0, gamma=0.0)
    return Workload(name, (0.0, 0.0, 0.0), hi, cell_size, periodic, pa, prm, dict(prm, kt=8e4), [], dt=1e-4,
                    rcut_inc=0.4)


def rotating_drum(n_target=1_000_000, seed=0xD3A2, jitter=2e-5, length=None, name="drum") -> Workload:
    """S2 (= BASELINE configs[1]): monodisperse spheres r=0.5 in a y-axis cylinder driver, periodic in y.
    Deck: example/spheres/rotating-drum/rotating-drum-mini.msp:10-36 (+ -common.msp:11-19)."""
    rng = np.random.default_rng(seed)
    cell = 2.5
    if length is None:
        length = max(10.0, cell * round((n_target ** (1.0 / 3.0)) / cell))
    # pi (R-0.6)^2 L ~ n_target
    R = float(np.sqrt(n_target / (PI * length)) + 0.6)
    side = cell * np.ceil((2 * R + 0.5) / cell)
    c = 0.5 * side
    pos = _lattice((0, 0, 0), (side, length, side), 1.0)
    rad = np.hypot(pos[:, 0] - c, pos[:, 2] - c)
    pos = pos[rad < R - 0.6]
    n = len(pos)
    pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    pa = ParticleArrays(n)
    pa.rx[:], pa.ry[:], pa.rz[:] = pos[:, 0], pos[:, 1], pos[:, 2]
    v = rng.normal(0.0, 0.1, size=(n, 3))
    pa.vx[:], pa.vy[:], pa.vz[:] = v[:, 0], v[:, 1], v[:, 2]
    pa.radius[:] = 0.5
    sphere_mass_inertia(pa, 0.02)
    pa.id[:] = rng.permutation(n).astype(np.uint64)
    prm = dict(kn=1e5, kt=1e5, kr=0.1, mu=0.5, damp_rate=0.9, fc=0.0, dncut=0.0, gamma=0.0)
    drv = dict(kind=capi.DRIVER_CYLINDER, radius=R, center=(c, c, c), axis=(1, 0, 1), vrot=(0, 0.5, 0))
    return Workload(name, (0.0, 0.0, 0.0), (float(side), float(length), float(side)), cell, (0, 1, 0), pa, prm,
                    dict(prm, kt=8e4), [drv], dt=1e-4, rcut_inc=0.4)


# ---- counter-based generator ("generator = counter-based (splitmix64) so CPU and GPU produce identical inputs") -----------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, counter) -> np.ndarray:
    """splitmix64 output function applied to seed + (counter+1)*golden-gamma: value k of the stream never depends on how many values
    were drawn before it or on the order of the draws (Steele, Lea, Flood 2014)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + (np.asarray(counter, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def counter_uniform(seed: int, counter) -> np.ndarray:
    """uniform doubles in [0, 1) from the top 53 bits of splitmix64"""
    return (splitmix64(seed, counter) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def counter_normal(seed: int, counter) -> np.ndarray:
    """standard normals by Box-Muller on counters (2k, 2k+1) of the stream"""
    c = np.asarray(counter, dtype=np.uint64)
    u1 = 1.0 - counter_uniform(seed, 2 * c)          # (0, 1]
    u2 = counter_uniform(seed, 2 * c + np.uint64(1))
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * PI * u2)


def deposit_xyz(base_xyz: np.ndarray, tiles=(4, 4), variant="hooke", seed=0xD300 + 1, name="deposit_xyz") -> Workload:
    """S1 of SURVEY 8(d) (= BASELINE configs[0], the reference's own CPU-runnable deck example/spheres/deposit/add_xyz.msp): the 736
    sphere centres of `input_deposit_xyz.xyz` (box 10^3; passed in as `base_xyz`, see tests/golden/make_deposit_fixture.py) tiled
    `tiles` times in x and y (4 x 4 -> 11 776 spheres), r = 0.4, density 0.02, v = (0, 0, -0.5) + N(0, 0.1^2) per component
    (add_xyz.msp:31-44), surface driver z = -1 (:46-53), kn = kt = 1e4, kr = 0.1, mu = 0.2, damp = 0.9, dt = 1e-4, rcut_inc = 0.2,
    cell_size 1 (:59-75), periodic x and y.  variant "hooke" = plain contact_sphere (BASELINE configs[0]); "dmt" = the deck's own
    contact_singlemat_sphere_hooke_dmt with gamma = 10 and its config_driver (kt 8000, mu 0.5).  Velocities come from the
    counter-based generator keyed by particle id."""
    base = np.asarray(base_xyz, dtype=np.float64).reshape(-1, 3)
    nb = len(base)
    tx, ty = tiles
    pos = np.concatenate([base + np.array([10.0 * i, 10.0 * j, 0.0]) for j in range(ty) for i in range(tx)])
    n = len(pos)
    pa = ParticleArrays(n)
    pa.rx[:], pa.ry[:], pa.rz[:] = pos[:, 0], pos[:, 1], pos[:, 2]
    ids = np.arange(n, dtype=np.uint64)
    pa.vx[:] = 0.1 * counter_normal(seed, 3 * ids)
    pa.vy[:] = 0.1 * counter_normal(seed, 3 * ids + np.uint64(1))
    pa.vz[:] = -0.5 + 0.1 * counter_normal(seed, 3 * ids + np.uint64(2))
    pa.radius[:] = 0.4
    sphere_mass_inertia(pa, 0.02)
    pa.id[:] = ids
    prm = dict(kn=1e4, kt=1e4, kr=0.1, mu=0.2, damp_rate=0.9, fc=0.0, dncut=0.0, gamma=0.0)
    prd = dict(prm)
    law = capi.LAW_NONE
    if variant == "dmt":
        prm = dict(prm, gamma=10.0)
        prd = dict(prm, kt=8e3, mu=0.5)
        law = capi.LAW_DMT
    drv = dict(kind=capi.DRIVER_SURFACE, offset=-1.0, normal=(0, 0, 1), center=(0, 0, -1.0))
    return Workload(f"{name}[{variant}] {tx}x{ty} tiles of {nb}", (0.0, 0.0, -2.0), (10.0 * tx, 10.0 * ty, 13.0), 1.0, (1, 1, 0), pa, prm, prd, [drv],
                    law=law, dt=1e-4, rcut_inc=0.2)


def deposit(n_side=(22, 22, 22), seed=0xD3A1, name="deposit") -> Workload:
    """S1 (= BASELINE configs[0], the reference's CPU-runnable case): ~10 k spheres r=0.4 falling on a surface driver.
    Deck parameters: example/spheres/deposit/add_xyz.msp:31-59 (plain Hooke variant)."""
    rng = np.random.default_rng(seed)
    hi = tuple(float(n) for n in n_side)
    pos = _lattice((0, 0, 0), hi, 1.0)
    n = len(pos)
    pos = pos + rng.uniform(-0.09, 0.09, size=pos.shape)
    pa = ParticleArrays(n)
    pa.rx[:], pa.ry[:], pa.rz[:] = pos[:, 0], pos[:, 1], pos[:, 2]
    v = rng.normal(0.0, 0.05, size=(n, 3))
    pa.vx[:], pa.vy[:], pa.vz[:] = v[:, 0], v[:, 1], v[:, 2] - 1.0
    pa.radius[:] = 0.4 + 0.1 * rng.uniform(0, 1, n)
    sphere_mass_inertia(pa, 0.02)
    pa.id[:] = rng.permutation(n).astype(np.uint64)
    prm = dict(kn=1e4, kt=1e4, kr=0.1, mu=0.2, damp_rate=0.9, fc=0.0, dncut=0.0, gamma=0.0)
    drv = dict(kind=capi.DRIVER_SURFACE, offset=0.05, normal=(0, 0, 1), center=(0, 0, 0.05))
    return Workload(name, (0.0, 0.0, -2.0), (hi[0], hi[1], hi[2] + 2.0), 2.0, (1, 1, 0), pa, prm, prm, [drv],
                    dt=1e-4, rcut_inc=0.2)


# example/polyhedra/rotating_drum/octahedron.shp (vertices, edges, faces, OBB extent incl. Minkowski radius 0.1)
_OCT_A, _OCT_B = 0.2310789034541148, 0.32679491924311227
OCTAHEDRON = dict(
    vertices=[[_OCT_A, -_OCT_A, 0.0], [_OCT_A, _OCT_A, 0.0], [0.0, 0.0, _OCT_B], [-_OCT_A, -_OCT_A, 0.0],
              [-_OCT_A, _OCT_A, 0.0], [0.0, 0.0, -_OCT_B]],
    edges=[[0, 1], [2, 1], [2, 0], [0, 3], [2, 3], [3, 4], [4, 2], [4, 1], [5, 0], [5, 1], [5, 4], [5, 3]],
    faces=[[0, 1, 2], [2, 3, 4], [1, 2, 4], [0, 2, 3], [0, 5, 1], [0, 5, 3], [3, 5, 4], [4, 5, 1]],
    obb=[0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0.33107890345411484, 0.33107890345411484, 0.4267949192431123],
    radius=0.1, volume=0.16666666666666666, inertia_on_mass=(0.04999999999999999,) * 3)


def cube_shape(half=0.3, mink=0.1):
    """Cube (8 v / 12 e / 6 quad faces), same topology family as src/shape/include/exaDEM/shape/basic_shape/cube.hpp."""
    v = [[sx * half, sy * half, sz * half] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)]
    e = [[0, 1], [1, 3], [3, 2], [2, 0], [4, 5], [5, 7], [7, 6], [6, 4], [0, 4], [1, 5], [2, 6], [3, 7]]
    f = [[0, 1, 3, 2], [4, 6, 7, 5], [0, 4, 5, 1], [2, 3, 7, 6], [0, 2, 6, 4], [1, 5, 7, 3]]
    ext = half + mink
    side = 2 * half
    return dict(vertices=v, edges=e, faces=f, obb=[0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, ext, ext, ext], radius=mink,
                volume=side ** 3, inertia_on_mass=(side * side / 6.0,) * 3)


def sphere_shape(r=0.3):
    """A sphere as a one-vertex sphero-polyhedron (how the polyhedron operators carry spheres in mixed decks):
    src/shape/include/exaDEM/shape/basic_shape/sphere.hpp:14-36 — one vertex at the origin, Minkowski radius r,
    I/m = 0.5 r^2 and volume `4 / 3 * pi * r^3` evaluated with the integer quotient 4/3 = 1 exactly as the reference writes it."""
    return dict(vertices=[[0.0, 0.0, 0.0]], edges=[], faces=[], obb=[0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, r, r, r], radius=r,
                volume=(4 // 3) * PI * r * r * r, inertia_on_mass=(0.5 * r * r,) * 3)


def _rand_quat(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def polyhedra_box(n_side=(8, 8, 8), seed=0xD3A3, shapes=("octahedron",), pitch=0.8, periodic=(1, 1, 1), jitter=0.03,
                  name="polybox", drivers=None, cell_size=2.0) -> Workload:
    """S3-like: sphero-polyhedra (octahedra and/or cubes) on a jittered SC lattice with random orientations.
    Law parameters: example/polyhedra/rotating_drum/rotating-drum-common.msp:24-29."""
    rng = np.random.default_rng(seed)
    spec = [OCTAHEDRON if s == "octahedron" else sphere_shape() if s == "sphere" else cube_shape() for s in shapes]
    hi = tuple(float(n) * pitch for n in n_side)
    pos = _lattice((0, 0, 0), hi, pitch)
    n = len(pos)
    pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    pa = ParticleArrays(n)
    pa.rx[:], pa.ry[:], pa.rz[:] = pos[:, 0], pos[:, 1], pos[:, 2]
    v = rng.normal(0.0, 0.01, size=(n, 3))
    pa.vx[:], pa.vy[:], pa.vz[:] = v[:, 0], v[:, 1], v[:, 2]
    w = rng.normal(0.0, 0.05, size=(n, 3))
    pa.vrotx[:], pa.vroty[:], pa.vrotz[:] = w[:, 0], w[:, 1], w[:, 2]
    q = _rand_quat(rng, n)
    pa.qw[:], pa.qx[:], pa.qy[:], pa.qz[:] = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    pa.type[:] = rng.integers(0, len(spec), n).astype(np.uint32)
    pa.homothety[:] = 1.0
    density = 2.6
    for t, s in enumerate(spec):
        m = pa.type == t
        vmax = max(np.linalg.norm(vv) for vv in s["vertices"]) + s["radius"]  # shape::compute_max_rcut, shape.hpp:479-489
        pa.radius[m] = vmax
        pa.mass[m] = density * s["volume"]
        pa.inertiax[m] = s["inertia_on_mass"][0] * pa.mass[m]
        pa.inertiay[m] = s["inertia_on_mass"][1] * pa.mass[m]
        pa.inertiaz[m] = s["inertia_on_mass"][2] * pa.mass[m]
    pa.id[:] = rng.permutation(n).astype(np.uint64)
    prm = dict(kn=1e4, kt=1e4, kr=0.0, mu=0.1, damp_rate=0.999, fc=0.0, dncut=0.0, gamma=0.0)
    return Workload(name, (0.0, 0.0, 0.0), hi, cell_size, periodic, pa, prm, dict(prm, mu=0.3), drivers or [],
                    shapes_spec=spec, mode=capi.MODE_POLYHEDRON, dt=1e-4, rcut_inc=0.1)


def box_mesh(lo, hi, mink=0.01):
    """Closed axis-aligned box as an RShape mesh (8 v / 12 e / 6 quad faces), the topology of
    example/polyhedra/stl_mesh/box.stl after stl_mesh_to_driver (vertex numbering is an input, quirk Q12).  Local
    coordinates are centred on the box centre; use `center` of the driver to place it."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    c, hl = 0.5 * (lo + hi), 0.5 * (hi - lo)
    v = [[sx * hl[0], sy * hl[1], sz * hl[2]] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)]
    e = [[0, 1], [1, 3], [3, 2], [2, 0], [4, 5], [5, 7], [7, 6], [6, 4], [0, 4], [1, 5], [2, 6], [3, 7]]
    f = [[0, 1, 3, 2], [4, 6, 7, 5], [0, 4, 5, 1], [2, 3, 7, 6], [0, 2, 6, 4], [1, 5, 7, 3]]
    ext = hl + mink
    spec = dict(vertices=v, edges=e, faces=f, obb=[0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, ext[0], ext[1], ext[2]], radius=mink)
    return spec, tuple(c)
