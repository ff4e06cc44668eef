// oracle/dem_math.hpp — CPU restatement of the exaDEM contact arithmetic.  TEST INFRASTRUCTURE ONLY.
//
// This header is part of the ORACLE (the checker).  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may use it; the product path (exadem_b200/) must never
// include, link or call it.
//
// Every function cites the reference file:line whose arithmetic (including operation ORDER, which matters
// for bit-exact threshold filters and 1e-10 force parity) it restates.  Paths are relative to /root/reference.
// The vector algebra of the un-vendored Onika library (onika/math/*) is restated from its call-site contract
//  — "parity unpinned" for that layer (risk R1); everything above it is pinned bit-for-bit
// against the reference's own headers through oracle/_ref (tests/test_oracle_vs_ref.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace orc {

struct v3 {
  double x, y, z;
};
struct q4 {
  double w, x, y, z;
};  // exanb::Quaternion {w,x,y,z}

// --- onika/math/basic_types_operators.h (absent; working definition, evaluation order left-to-right) ---
static inline v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline v3 operator-(v3 a) { return {-a.x, -a.y, -a.z}; }
static inline v3 operator*(v3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline v3 operator*(double s, v3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline v3 operator/(v3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
static inline v3 cmul(v3 a, v3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // Vec3d*Vec3d (cylinder.hpp:226)
static inline double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double norm(v3 a) { return std::sqrt(dot(a, a)); }
static inline v3 cross(v3 u, v3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
static inline bool is_zero(v3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }

// quaternion rotation: src/shape/include/exaDEM/type/quat.hpp:79-90 (v + 2s(q x v) + 2 q x (q x v))
static inline v3 qrot(q4 q, v3 v) {
  const v3 u{q.x, q.y, q.z};
  v3 uv = cross(u, v);
  v3 uuv = cross(u, uv);
  uv = uv * (2.0 * q.w);
  uuv = uuv * 2.0;
  return v + uv + uuv;
}
static inline q4 qconj(q4 q) { return {q.w, -q.x, -q.y, -q.z}; }
// dQ/dt = 1/2 omega (x) Q : quat.hpp:92-96
static inline q4 qdot(q4 q, v3 o) {
  return {-0.5 * (o.x * q.x + o.y * q.y + o.z * q.z), 0.5 * (o.y * q.z - o.z * q.y + o.x * q.w),
          0.5 * (o.z * q.x - o.x * q.z + o.y * q.w), 0.5 * (o.x * q.y - o.y * q.x + o.z * q.w)};
}

struct Contact {
  bool hit = false;
  double dn = 0;
  v3 n{0, 0, 0};
  v3 cp{0, 0, 0};
};

// ---- filters / detection cores: src/shape/include/exaDEM/shape_detection_core.hpp ----
// :43-53
static inline bool filter_vv(double rV, v3 vi, double ri, v3 vj, double rj) {
  const double R = ri + rj + rV;
  const v3 d = vi - vj;
  return dot(d, d) <= R * R;
}
// :57-81
static inline Contact detect_vv(v3 pi, double ri, v3 pj, double rj) {
  const double R = ri + rj;
  const v3 d = pi - pj;
  const double len = norm(d);
  const double inv = 1.0 / len;
  Contact c;
  c.dn = len - R;
  c.n = d * inv;
  c.cp = pi - c.n * (ri + 0.5 * c.dn);
  c.hit = c.dn < 0;
  return c;
}
// :83-101
static inline bool filter_ve(double rV, v3 vi, double ri, v3 vf, v3 vs, double rj) {
  const v3 fs = vs - vf, fi = vi - vf;
  const double r = dot(fs, fi) / dot(fs, fs);
  if (r < -rV || r > 1.0 + rV) return false;
  const v3 d = fi - fs * r;
  const double Rmax = ri + rj + rV;
  return dot(d, d) < Rmax * Rmax;
}
// :122-148
static inline Contact detect_ve(v3 vi, double ri, v3 vf, v3 vs, double rj) {
  const v3 fs = vs - vf, fi = vi - vf;
  const double r = dot(fs, fi) / dot(fs, fs);
  if (r <= 0 || r >= 1.0) return Contact();
  v3 n = fi - fs * r;
  const double dn = norm(n) - (ri + rj);
  if (dn > 0) return Contact();
  n = n / norm(n);  // _normalize, include/exaDEM/normalize.hpp
  Contact c;
  c.hit = true;
  c.dn = dn;
  c.n = n;
  c.cp = vi - n * (ri + 0.5 * dn);
  return c;
}

// point-in-polygon shared by filter (:245-309) and detection (:398-466); `va` = world vertices of the
// face owner, `idx[0..nf)` = the face's vertex indices.  Returns (inside?, dist, n).
template <class VA>
static inline bool face_test(v3 vi, const VA& va, const int* idx, int nf, double cut, double& dist_out, v3& n_out) {
  const v3 a = va[idx[0]], b = va[idx[1]], c = va[idx[nf - 1]];
  const v3 v = vi - a;
  v3 v1 = b - a;
  v3 v2 = c - a;
  v1 = v1 / norm(v1);
  v3 n = cross(v1, v2);
  n = n / norm(n);
  double dist = dot(n, v);
  if (dist < 0) {
    n = n * (-1);
    dist = -dist;
  }
  dist_out = dist;
  n_out = n;
  if (dist > cut) return false;  // caller distinguishes "too far" via dist_out > cut
  const v3 P = vi - n * dist;
  int odd = 0;
  v2 = cross(n, v1);
  const double o1 = dot(P, v1), o2 = dot(P, v2);
  for (int ia = 0; ia < nf; ++ia) {
    int ib = ia + 1;
    if (ib == nf) ib = 0;
    const v3 pa = va[idx[ia]], pb = va[idx[ib]];
    const double pa1 = dot(pa, v1), pb1 = dot(pb, v1), pa2 = dot(pa, v2), pb2 = dot(pb, v2);
    if ((pa2 < o2 && pb2 >= o2) || (pb2 < o2 && pa2 >= o2)) {
      if (pa1 + (o2 - pa2) / (pb2 - pa2) * (pb1 - pa1) < o1) odd = 1 - odd;
    }
  }
  return odd == 1;
}
template <class VA>
static inline bool filter_vf(double rV, v3 vi, double ri, const VA& va, double rj, const int* idx, int nf) {
  double dist;
  v3 n;
  return face_test(vi, va, idx, nf, ri + rj + rV, dist, n);
}
template <class VA>
static inline Contact detect_vf(v3 vi, double ri, const VA& va, double rj, const int* idx, int nf) {
  double dist;
  v3 n;
  const bool inside = face_test(vi, va, idx, nf, ri + rj, dist, n);
  if (dist > ri + rj) return Contact();
  Contact c;
  c.dn = dist - (ri + rj);
  c.n = n;
  c.cp = vi - n * (ri + 0.5 * c.dn);
  c.hit = inside;
  return c;
}
// :468-504 / :521-566
static inline bool ee_closest(v3 fi, v3 si, v3 fj, v3 sj, v3& pi, v3& pj) {
  const v3 Ei = si - fi, Ej = sj - fj, v = fi - fj;
  const double c = dot(Ei, Ei), d = dot(Ej, Ej), e = dot(Ei, Ej);
  double f = (c * d) - (e * e);
  if (!(std::fabs(f) > 1.0e-12)) return false;
  f = 1.0 / f;
  const double a = dot(Ei, v), b = dot(Ej, v);
  const double s = (e * b - a * d) * f;
  const double t = (c * b - e * a) * f;
  if (s <= 0.0 || s >= 1.0 || t <= 0.0 || t >= 1.0) return false;
  pi = fi + Ei * s;
  pj = fj + Ej * t;
  return true;
}
static inline bool filter_ee(double rV, v3 fi, v3 si, double ri, v3 fj, v3 sj, double rj) {
  const double R = ri + rj + rV;
  v3 pi, pj;
  if (!ee_closest(fi, si, fj, sj, pi, pj)) return false;
  const v3 n = pi - pj;
  return dot(n, n) <= R * R;
}
static inline Contact detect_ee(v3 fi, v3 si, double ri, v3 fj, v3 sj, double rj) {
  const double R = ri + rj;
  v3 pi, pj;
  if (!ee_closest(fi, si, fj, sj, pi, pj)) return Contact();
  v3 n = pi - pj;
  const double dn = norm(n) - R;
  if (dn > 0) return Contact();
  n = n / norm(n);
  Contact c;
  c.hit = dn <= 0;
  c.dn = dn;
  c.n = n;
  c.cp = pi - n * (ri + 0.5 * dn);
  return c;
}

// ---- OBB: src/shape/include/exaDEM/type/OBB.hpp ----
struct Obb {
  v3 c, e1, e2, e3, ext;
};
// basic_types.hpp:37-49 (+ OBB::rotate :83-88, enlarge :63-67)
static inline Obb obb_world(const Obb& in, v3 pos, q4 q, double h) {
  Obb o = in;
  o.e1 = qrot(q, o.e1);
  o.e2 = qrot(q, o.e2);
  o.e3 = qrot(q, o.e3);
  o.c = qrot(q, o.c);
  if (h != 1.0) {
    o.ext = o.ext * h;
    o.c = o.c * h;
  }
  o.c = o.c + pos;
  return o;
}
static inline void obb_enlarge(Obb& o, double m) {
  o.ext.x += m;
  o.ext.y += m;
  o.ext.z += m;
}
// OBB.hpp:90-222 — 15-axis SAT, tol = FLT_EPSILON added to every |R| term
static inline bool obb_intersect(const Obb& A, const Obb& B) {
  const double tol = 1.1920928955078125e-07;  // FLT_EPSILON
  double R[3][3], Q[3][3];
  const v3 ea[3] = {A.e1, A.e2, A.e3}, eb[3] = {B.e1, B.e2, B.e3};
  const double xa[3] = {A.ext.x, A.ext.y, A.ext.z}, xb[3] = {B.ext.x, B.ext.y, B.ext.z};
  const v3 tt = A.c - B.c;
  const double t[3] = {dot(tt, A.e1), dot(tt, A.e2), dot(tt, A.e3)};
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      R[i][j] = dot(ea[i], eb[j]);
      Q[i][j] = std::fabs(R[i][j]) + tol;
    }
    const double ra = xa[i];
    const double rb = xb[0] * Q[i][0] + xb[1] * Q[i][1] + xb[2] * Q[i][2];
    if (std::fabs(t[i]) > ra + rb) return false;
  }
  for (int j = 0; j < 3; j++) {
    const double ra = xa[0] * Q[0][j] + xa[1] * Q[1][j] + xa[2] * Q[2][j];
    const double rb = xb[j];
    if (std::fabs(t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j]) > ra + rb) return false;
  }
  double ra, rb;
  // A0 x B0..B2
  ra = xa[1] * Q[2][0] + xa[2] * Q[1][0]; rb = xb[1] * Q[0][2] + xb[2] * Q[0][1];
  if (std::fabs(t[2] * R[1][0] - t[1] * R[2][0]) > ra + rb) return false;
  ra = xa[1] * Q[2][1] + xa[2] * Q[1][1]; rb = xb[0] * Q[0][2] + xb[2] * Q[0][0];
  if (std::fabs(t[2] * R[1][1] - t[1] * R[2][1]) > ra + rb) return false;
  ra = xa[1] * Q[2][2] + xa[2] * Q[1][2]; rb = xb[0] * Q[0][1] + xb[1] * Q[0][0];
  if (std::fabs(t[2] * R[1][2] - t[1] * R[2][2]) > ra + rb) return false;
  // A1 x B0..B2
  ra = xa[0] * Q[2][0] + xa[2] * Q[0][0]; rb = xb[1] * Q[1][2] + xb[2] * Q[1][1];
  if (std::fabs(t[0] * R[2][0] - t[2] * R[0][0]) > ra + rb) return false;
  ra = xa[0] * Q[2][1] + xa[2] * Q[0][1]; rb = xb[0] * Q[1][2] + xb[2] * Q[1][0];
  if (std::fabs(t[0] * R[2][1] - t[2] * R[0][1]) > ra + rb) return false;
  ra = xa[0] * Q[2][2] + xa[2] * Q[0][2]; rb = xb[0] * Q[1][1] + xb[1] * Q[1][0];
  if (std::fabs(t[0] * R[2][2] - t[2] * R[0][2]) > ra + rb) return false;
  // A2 x B0..B2
  ra = xa[0] * Q[1][0] + xa[1] * Q[0][0]; rb = xb[1] * Q[2][2] + xb[2] * Q[2][1];
  if (std::fabs(t[1] * R[0][0] - t[0] * R[1][0]) > ra + rb) return false;
  ra = xa[0] * Q[1][1] + xa[1] * Q[0][1]; rb = xb[0] * Q[2][2] + xb[2] * Q[2][0];
  if (std::fabs(t[1] * R[0][1] - t[0] * R[1][1]) > ra + rb) return false;
  ra = xa[0] * Q[1][2] + xa[1] * Q[0][2]; rb = xb[0] * Q[2][1] + xb[1] * Q[2][0];
  if (std::fabs(t[1] * R[0][2] - t[0] * R[1][2]) > ra + rb) return false;
  return true;
}
// nbh_polyhedron.cpp:305-310 — axis-aligned cube of half-width `m` around a vertex
static inline Obb obb_point(v3 c, double m) { return Obb{c, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {m, m, m}}; }

// ---- contact law: src/forcefield/include/exaDEM/forcefield/{common_kernels,contact_force}.hpp ----
struct Params {
  double kn = 0, kt = 0, kr = 0, mu = 0, damp = 0, fc = 0, dncut = 0, gamma = 0;  // contact_parameters.hpp:32-41
};
enum Law { LAW_NONE = 0, LAW_COHESIVE = 1, LAW_DMT = 2 };

// contact_force.hpp:146-235.  ft, mom: per-interaction history (in/out); f: force on i (out).
static inline void contact_force(int law, double dn, v3 n, double dt, const Params& p, double meff, double reff, v3& ft,
                                 v3 cp, v3 pos_i, v3 vel_i, v3& f, v3& mom, v3 w_i, v3 pos_j, v3 vel_j, v3 w_j) {
  if (law == LAW_COHESIVE && dn >= 0) {  // :165-173
    const double fnv = (p.fc / p.dncut) * dn - p.fc;
    f = f + fnv * n;
    ft = {0, 0, 0};
    return;
  }
  const double damp = p.damp * 2.0 * std::sqrt(p.kn * meff);  // common_kernels.hpp:28-31
  const v3 ci = vel_i - cross(cp - pos_i, w_i);                 // :34-41
  const v3 cj = vel_j - cross(cp - pos_j, w_j);
  const v3 vel = cj - ci;
  const double vn = dot(vel, n);
  const double fne = -p.kn * dn, fnv = damp * vn;               // :48-53
  double fs = (fnv + fne > 0) ? fnv + fne : 0;
  v3 fn;
  if (law == LAW_COHESIVE) {                                    // :69-76
    const double a = fs - p.fc;
    const double b = a < (-p.fc) ? -p.fc : a;
    fn = b * n;
  } else {
    fn = fs * n;
  }
  const v3 vt = vel - vn * n;                                   // :80-90
  ft = ft + p.kt * (dt * vt);
  double thr;
  if (law != LAW_COHESIVE) thr = std::fabs(p.mu * (p.kn * dn));  // :108-114
  else thr = std::fabs(p.mu * (dot(fn, n) + p.fc));              // :116-120
  const double ft2 = dot(ft, ft);                               // :122-128
  if (ft2 > 0.0 && ft2 > thr * thr) ft = ft * (thr / std::sqrt(ft2));
  f = fn + ft;
  if (law == LAW_DMT && p.gamma > 0.0 && dn < 0.0) {            // contact_force.hpp:115-127
    const double Fd = 2.0 * M_PI * reff * p.gamma;
    f = f - Fd * n;
  }
  mom = mom + p.kr * (w_j - w_i) * dt;                          // :223
  v3 branch = cp - pos_i;
  const double r = dot(branch, w_i) / dot(w_i, w_i);            // NaN when w_i == 0 (quirk Q1)
  branch = branch - r * w_i;
  const double mur = 0;
  const double thm = std::fabs(mur * norm(branch) * norm(fn));
  const double m2 = dot(mom, mom);
  if (m2 > 0.0 && m2 > thm * thm) mom = mom * (thm / std::sqrt(m2));
}
// common_kernels.hpp:137-145
static inline v3 moments(v3 cp, v3 p, v3 f, v3 m) { return cross(cp - p, f) + m; }

// ---- inner bonds (fragmentation, wave 13): src/forcefield/include/exaDEM/forcefield/inner_bond_force.hpp:26-70 ----
struct BondParams {   // InnerBondParams (inner_bond_parameters.hpp:33-45)
  double kn = 0, kt = 0, damp_rate = 0;
  int mode = 3;       // RuptureMode (rupture_criterion.hpp:26): 0 EnergyMixedMode, 1 EnergySeparateMode, 2 StressEnergySeparateMode, 3 None
  double crit1 = 0, crit2 = 0;
};
static inline void inner_bond_force_core(double dn, v3 n, double dn0, double weight, double dt, const BondParams& p, double meff, double& En,
                                         v3& tds, double& Et, v3& ft, v3 cp, v3 pi, v3 vi, v3& fi, v3 wi, v3 pj, v3 vj, v3 wj) {
  const double damp = p.damp_rate * 2.0 * std::sqrt(p.kn * meff);                       // common_kernels.hpp:28-31
  const v3 vel = (vj - cross(cp - pj, wj)) - (vi - cross(cp - pi, wi));                 // :34-41
  const double vn = dot(vel, n);
  const double delta = dn - dn0;
  const double kn_w = p.kn * weight, kt_w = p.kt * weight;
  const double fne = -kn_w * delta, fnv = damp * vn;
  const double fn = fne + fnv;
  const v3 vfn = fn * n;
  const v3 ds = dt * (vel - vn * n);                                                    // compute_tangential_force, :80-85
  tds = tds + ds;
  ft = kt_w * tds;
  fi = vfn + ft;
  En = (fne <= 0 ? 1.0 : 0.0) * (0.5 * kn_w * delta * delta);
  Et = 0.5 * kt_w * dot(tds, tds);
}
// apply_interface_fracture_criterion.hpp:40-73: one interface = `size` consecutive bonds; returns true when it breaks
static inline bool interface_breaks(int mode, double c1, double c2, const double* en, const double* et, const double* dn, const v3* fn, size_t size) {
  double En = 0.0, Et = 0.0, S = 0.0;
  for (size_t j = 0; j < size; j++) {
    En += en[j];
    Et += et[j];
    if (dn[j] > 0.0) S += norm(fn[j]);
  }
  S /= size;
  if (mode == 0) return En + Et > c1;
  if (mode == 1) return En > c1 || Et > c2;
  if (mode == 2) return (En > c1) && (S > c2);
  return false;
}

// ---- drivers: src/driver/include/exaDEM/{cylinder,surface,ball}.hpp ----
enum DriverKind { DRV_CYLINDER = 0, DRV_SURFACE = 1, DRV_BALL = 2, DRV_RSHAPE = 3, DRV_UNDEFINED = 4 };
struct Driver {
  int kind = DRV_UNDEFINED;
  double radius = 0;      // cylinder, ball
  v3 axis{1, 0, 1};       // cylinder (component mask, quirk Q11)
  v3 center{0, 0, 0};     // cylinder: already multiplied by axis (cylinder.hpp:104); ball center; surface ref. center
  double offset = 0;      // surface
  v3 normal{0, 0, 1};     // surface
  v3 vel{0, 0, 0}, vrot{0, 0, 0};
  v3 forces{0, 0, 0};     // accumulated -f when need_forces
  int need_forces = 0;    // motion_type.hpp:133-141
  int shape = -1;         // RSHAPE: index into the mesh-shape table
  q4 quat{1, 0, 0, 0};    // RSHAPE: orientation
  v3 cproj{0, 0, 0};      // surface: center_proj_ = offset * normal at initialize() (surface.hpp:149), then moved by push_f_v_r
  v3 center_proj() const { return cproj; }
  // ---- motion: Driver_params (driver_params.hpp:33-66) + integrator state of the driver's fields_ ----
  int motion = 0;         // MotionType (motion_type.hpp:23-35)
  v3 motion_vector{0, 0, 0};
  double t_start = 0, t_end = 1e300, const_vel = 0, const_force = 0, sigma = 0, damprate = 0, pmass = 0 /* Driver_params::mass_ */;
  std::vector<double> tab_time;
  std::vector<v3> tab_pos;
  double omega = 0, amplitude = 0;
  v3 shaker_dir{0, 0, 1}, pend_anchor{0, 0, 0}, pend_init{0, 0, 0}, pend_dir{0, 0, 0};
  int expr_use_v = 0, expr_use_vrot = 0, expr_use_mom = 0, drive_by_mom = 0;
  v3 expr_v{0, 0, 0}, expr_vrot{0, 0, 0}, expr_mom{0, 0, 0};
  double mass = 1e300, surface = -1, rv = 0, ra = 0;
  v3 acc{0, 0, 0}, inertia{0, 0, 0}, applied_mom{0, 0, 0}, mom_axis{0, 0, 0}, mom{0, 0, 0};
};
enum Motion { M_STATIONARY, M_LINEAR_MOTION, M_COMPRESSIVE_FORCE, M_LINEAR_FORCE_MOTION, M_PARTICLE, M_LINEAR_COMPRESSIVE_MOTION, M_TABULATED,
              M_SHAKER, M_PENDULUM_MOTION, M_EXPRESSION, M_UNKNOWN };
// motion_type.hpp:84-141
static inline bool m_linear(int m) { return m == M_LINEAR_MOTION || m == M_LINEAR_FORCE_MOTION || m == M_LINEAR_COMPRESSIVE_MOTION; }
static inline bool m_compressive(int m) { return m == M_COMPRESSIVE_FORCE || m == M_LINEAR_COMPRESSIVE_MOTION; }
static inline bool m_force_motion(int m) { return m == M_PARTICLE || m == M_LINEAR_FORCE_MOTION; }
static inline bool m_need_forces(int m) { return m_compressive(m) || m == M_PARTICLE || m == M_LINEAR_FORCE_MOTION; }
// driver_params.hpp:82-84
static inline bool motion_triggered(const Driver& d, double t) { return t >= d.t_start && t <= d.t_end; }
// driver_params.hpp:277-310 (std::lower_bound = first entry >= time)
static inline size_t tab_lower(const Driver& d, double time) {
  size_t i = 0;
  while (i < d.tab_time.size() && d.tab_time[i] < time) i++;
  return i;
}
static inline v3 tab_to_position(const Driver& d, double time) {
  const size_t it = tab_lower(d, time);
  if (it == d.tab_time.size()) return d.tab_pos.back();
  if (it == 0) return d.tab_pos[0];        // time <= tab_time[0]: the reference indexes tab_time_[-1] here (:284-288, UB); defined as the first entry
  const size_t lo = it - 1, up = lo + 1;
  if (up >= d.tab_time.size()) return d.tab_pos.back();
  const double Dt = (time - d.tab_time[lo]) / (d.tab_time[up] - d.tab_time[lo]);
  return (d.tab_pos[up] - d.tab_pos[lo]) * Dt + d.tab_pos[lo];
}
static inline v3 tab_to_velocity(const Driver& d, double time) {
  const size_t it = tab_lower(d, time);
  if (it == d.tab_time.size() || it == 0) return v3{0, 0, 0};   // it == 0: same UB as above, defined as "not moving yet"
  const size_t lo = it - 1, up = lo + 1;
  if (up >= d.tab_time.size()) return v3{0, 0, 0};
  const double Dt = d.tab_time[up] - d.tab_time[lo];
  return (d.tab_pos[up] - d.tab_pos[lo]) / Dt;
}
// driver_params.hpp:313-329
static inline double shaker_signal(const Driver& d, double time) { time -= d.t_start; return d.amplitude * std::sin(d.omega * time); }
static inline v3 shaker_velocity(const Driver& d, double time) { time -= d.t_start; return d.amplitude * d.omega * std::cos(d.omega * time) * d.shaker_dir; }
// driver_params.hpp:340-371
static inline void pendulum_offset_normal(const Driver& d, double time, double& offset, v3& normal) {
  const v3 v1 = d.pend_anchor, v2 = d.pend_init;
  const v3 v3_ = d.pend_init + d.pend_dir * shaker_signal(d, time);   // pendulum_signal == shaker_signal formula (:366-370)
  if (dot(v3_, v2) < 1e-16) { offset = 0.0; normal = d.pend_dir; return; }
  v3 v1v3 = v3_ - v1;
  v1v3 = v1v3 / norm(v1v3);
  const v3 proj = dot(v1v3, v2 - v1) * v1v3 + v1;
  const v3 dir = proj - v2;
  normal = dir / norm(dir);
  offset = dot(v1, normal);
}
// driver_params.hpp:68-74
static inline void update_forces(const Driver& d, v3& forces) {
  if (d.motion == M_LINEAR_FORCE_MOTION) forces = (dot(forces, d.motion_vector) + d.const_force) * d.motion_vector;
  else if (d.motion != M_PARTICLE) forces = v3{0, 0, 0};
}
static inline bool rshape_need_moment(const Driver& d) { return d.drive_by_mom || d.motion == M_PARTICLE; }   // rshape.hpp:406-413

// push_f_v_r of every driver kind: cylinder.hpp:160-173, surface.hpp:262-299, ball.hpp:203-213, rshape.hpp:283-320
static inline void driver_push_f_v_r(Driver& d, double time, double dt) {
  const int m = d.motion;
  if (d.kind == DRV_CYLINDER) {
    if (m != M_STATIONARY && m == M_SHAKER) {
      const double displ = shaker_signal(d, time + dt) - shaker_signal(d, time);
      d.center = d.center + displ * d.shaker_dir;
      d.vel = shaker_velocity(d, time + dt);
    }
  } else if (d.kind == DRV_SURFACE) {
    if (m == M_STATIONARY) return;
    if (m == M_PENDULUM_MOTION) {
      double off; v3 nrm;
      pendulum_offset_normal(d, time + dt, off, nrm);
      d.normal = nrm; d.offset = off;
      d.vel = v3{0, 0, 0};
      d.cproj = d.normal * d.offset;
      return;
    }
    double displ = dt * dot(d.vel, d.normal) + 0.5 * dt * dt * d.acc.x;
    if (m == M_SHAKER) {
      const double sn = shaker_signal(d, time + dt), sc = shaker_signal(d, time);
      const double af = dot(d.shaker_dir, d.normal);
      displ = (sn - sc) * af;
      d.vel = shaker_velocity(d, time + dt);
    }
    d.center = d.center + displ * d.normal;
    d.offset += displ;
    d.cproj = d.cproj + displ * d.normal;
  } else if (d.kind == DRV_BALL) {
    if (m == M_TABULATED) {
      d.center = tab_to_position(d, time);
      d.vel = tab_to_velocity(d, time);
    } else if (m != M_STATIONARY) {
      if (m_compressive(m)) d.radius += dt * d.rv + 0.5 * dt * dt * d.ra;
      d.center = d.center + dt * d.vel;
    }
  } else if (d.kind == DRV_RSHAPE) {
    if (m == M_TABULATED) {
      d.center = tab_to_position(d, time);
      d.vel = tab_to_velocity(d, time);
    } else if (m != M_STATIONARY) {
      if (m == M_SHAKER) {
        d.vel = cmul(shaker_velocity(d, time + dt), d.shaker_dir);   // Vec3d * Vec3d, rshape.hpp:302
        d.acc = v3{0, 0, 0};
      }
      if (m == M_EXPRESSION && motion_triggered(d, time)) {
        if (d.expr_use_v) d.vel = d.expr_v;
        if (d.expr_use_vrot) d.vrot = d.expr_vrot;
      }
      d.center = d.center + (dt * d.vel + 0.5 * dt * dt * d.acc);
    }
  }
}
// push_f_v: cylinder.hpp:149-151 (nothing), surface.hpp:232-252, ball.hpp:219-234, rshape.hpp:262-281
static inline void driver_push_f_v(Driver& d, double dt) {
  const int m = d.motion;
  if (d.kind == DRV_SURFACE) {
    if (m == M_STATIONARY) d.vel = v3{0, 0, 0};
    else if (m_compressive(m)) { if (d.sigma != 0) d.vel = d.vel + (dt * d.acc.x) * d.motion_vector; }
    else if (m == M_LINEAR_MOTION) d.vel = d.const_vel * d.motion_vector;
  } else if (d.kind == DRV_BALL) {
    if (m_force_motion(m)) d.vel = d.acc * dt;
    if (m == M_LINEAR_MOTION) d.vel = d.motion_vector * d.const_vel;
    if (m_compressive(m)) {
      if (m == M_COMPRESSIVE_FORCE) d.vel = v3{0, 0, 0};
      if (d.sigma != 0) d.rv += dt * d.ra;
    }
  } else if (d.kind == DRV_RSHAPE) {
    if (m == M_STATIONARY) d.vel = v3{0, 0, 0};
    else {
      if (m_force_motion(m)) d.vel = d.vel + d.acc * dt;
      if (m == M_LINEAR_MOTION) d.vel = d.motion_vector * d.const_vel;
      if (m_compressive(m) && d.sigma != 0) d.vel = d.vel + 0.5 * dt * d.acc;
    }
  }
}
// push_av_to_quat (RShape only, DriverProperty::use_quaternion): rshape.hpp:337-377 + driver_numerical_scheme_kernel.hpp:26-56
static inline void driver_push_av_to_quat(Driver& d, double time, double dt) {
  if (d.kind != DRV_RSHAPE) return;
  auto push_q = [&](double dtq) {
    const q4 dq = qdot(d.quat, d.vrot);
    q4 Q{d.quat.w + dq.w * dtq, d.quat.x + dq.x * dtq, d.quat.y + dq.y * dtq, d.quat.z + dq.z * dtq};
    const double nn = std::sqrt(Q.w * Q.w + Q.x * Q.x + Q.y * Q.y + Q.z * Q.z);
    d.quat = {Q.w / nn, Q.x / nn, Q.y / nn, Q.z / nn};
  };
  if (rshape_need_moment(d)) {
    if (d.motion == M_EXPRESSION && motion_triggered(d, time) && d.expr_use_mom) {
      d.applied_mom = d.expr_mom;
      const double mn = norm(d.applied_mom);
      d.mom_axis = mn > 1e-12 ? d.applied_mom / mn : v3{0, 0, 0};
    }
    const v3 pm = d.motion == M_PARTICLE ? d.mom : dot(d.applied_mom + d.mom, d.mom_axis) * d.mom_axis;
    const q4 Qi = qconj(d.quat);
    const v3 o = qrot(Qi, d.vrot), M = qrot(Qi, pm), I = d.inertia;
    const v3 dw{(M.x - (I.z - I.y) * o.y * o.z) / I.x, (M.y - (I.x - I.z) * o.z * o.x) / I.y, (M.z - (I.y - I.x) * o.x * o.y) / I.z};
    const v3 arot = qrot(d.quat, dw);
    d.vrot = d.vrot + arot * (dt * 0.5);
    push_q(dt);
    d.vrot = d.vrot + (dt * 0.5) * arot;
    d.mom = v3{0, 0, 0};
  } else {
    push_q(dt);
  }
}
// force_to_accel_driver.cpp:30-45 (+ Ball::f_ra ball.hpp:273-283, Surface surface.hpp:208-225, RShape rshape.hpp:220-259)
static inline void driver_force_to_accel(Driver& d, double system_mass) {
  const int m = d.motion;
  if (d.kind == DRV_BALL) {
    d.pmass = system_mass;
    if (m_compressive(m) && d.pmass != 0) {
      const double pi = 4 * std::atan(1);
      const double s = 4 * pi * d.radius * d.radius;
      d.ra = (norm(d.forces) - d.sigma * s - (d.damprate * d.rv)) / (d.pmass * 0.5);
    }
    if (m_force_motion(m)) { update_forces(d, d.forces); d.acc = d.forces / d.mass; }
    else d.acc = v3{0, 0, 0};
  } else if (d.kind == DRV_SURFACE) {
    d.pmass = system_mass;
    if (m_compressive(m)) {
      if (d.pmass != 0) {
        const v3 tmp = (d.forces - d.sigma * d.surface * d.motion_vector) / (d.pmass * 0.5);
        d.acc.x = dot(tmp, d.motion_vector);
      } else d.acc.x = 0;
    }
  } else if (d.kind == DRV_RSHAPE) {
    if (m_compressive(m)) {
      if (d.mass != 0) {
        // Vec3d - double: the scalar is subtracted from every component (exanb operator, working definition)
        const double ss = d.sigma * d.surface;
        const v3 dv = d.damprate * d.vel;
        const v3 num{d.forces.x - ss - dv.x, d.forces.y - ss - dv.y, d.forces.z - ss - dv.z};
        const v3 tmp = num / (d.mass * 0.5);
        d.acc = dot(tmp, d.motion_vector) * d.motion_vector;
      }
    } else if (m_force_motion(m)) {
      update_forces(d, d.forces);
      d.acc = d.forces / d.mass;
    } else d.acc = v3{0, 0, 0};
  }
}
// cylinder.hpp:225-235, surface.hpp:320-324, ball.hpp:292-296
static inline bool driver_filter(const Driver& d, double rcut, v3 p) {
  if (d.kind == DRV_CYLINDER) {
    const v3 dir = cmul(p, d.axis) - d.center;
    const double dn = d.radius - (norm(dir) + rcut);
    return dn <= 0;
  } else if (d.kind == DRV_SURFACE) {
    const v3 proj = dot(p, d.normal) * d.normal;
    return norm(proj - d.center_proj()) <= rcut;
  } else {
    const v3 dist = d.center - p;
    const double dd = d.radius - norm(dist);
    return std::fabs(dd) <= rcut;
  }
}
// cylinder.hpp:255-279, surface.hpp:338-350, ball.hpp:309-327
static inline Contact driver_detect(const Driver& d, double rcut, v3 p) {
  Contact c;
  if (d.kind == DRV_CYLINDER) {
    const v3 dir = d.center - cmul(p, d.axis);
    const double len = norm(dir);
    c.dn = d.radius - (rcut + len);
    if (c.dn > 0) return c;
    c.n = dir / len;
  } else if (d.kind == DRV_SURFACE) {
    const v3 proj = dot(p, d.normal) * d.normal;
    const v3 s2p = -(d.center_proj() - proj);
    const double len = norm(s2p);
    c.dn = len - rcut;
    if (c.dn > 0) return c;
    c.n = s2p / len;
  } else {
    const v3 p2c = d.center - p;
    const double len = norm(p2c);
    v3 n = p2c / len;
    if (len > d.radius) {
      c.dn = len - d.radius - rcut;
      n = (-1) * n;
    } else {
      c.dn = d.radius - len - rcut;
    }
    if (c.dn > 0) return c;
    c.n = n;
  }
  c.cp = p - c.n * (rcut + 0.5 * c.dn);
  c.hit = true;
  return c;
}

}  // namespace orc
