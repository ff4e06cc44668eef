// dem_math.cuh — device-side fp64 arithmetic of the DEM contact path (sm_100a).
//
// Operation ORDER follows the reference expressions cited per function (paths relative to /root/reference) because
// (i) the candidate filters are threshold tests whose membership must be bit-exact and (ii) forces are compared at
// 1e-10.  The whole library is compiled with -fmad=false so that nvcc does not contract a*b+c (the reference's
// host build does not), which makes every result below bit-identical to an IEEE evaluation of the same expression.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edb {

struct v3 {
  double x, y, z;
};
struct q4 {
  double w, x, y, z;
};

#define EDB_HD __host__ __device__ __forceinline__

// 256-bit global accesses (sm_100a: LDG.E.ENL2.256 / STG.E.ENL2.256).  CUDA 12.9's double4 is only 16-byte aligned, so a plain
// double4 access compiles to two 128-bit instructions = twice the L1 wavefronts on scattered gathers, which is what bounds the
// contact / gather kernels (profiles/).  All double4 arrays of the library are 32-byte aligned (cudaMalloc + 32-byte stride).
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double4 ld4(const double4* p) {
  double4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}
// read-only for the lifetime of the kernel (non-coherent path)
__device__ __forceinline__ double4 ldg4(const double4* p) {
  double4 v;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st4(double4* p, const double4 v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
// L2 residency hints (createpolicy + .L2::cache_hint).  The per-step working set of a 1 M-particle brick is ~0.8 GB, the L2 126 MB:
// plain LRU keeps nothing from one kernel to the next.  The three gathered particle sectors (pr, vm, wg: 96 MB per 1 M particles)
// are what both per-step kernels touch — the contact kernel by scattered gathers, the gather by streaming read + write — so they
// are loaded and stored evict_last, and everything that is touched once per step (history, incidence codes, inertia, backup
// positions, orientations) evict_first.  With bricks too large for the L2 the hints cost nothing.
struct L2Pol {
  unsigned long long keep, stream;
};
__device__ __forceinline__ L2Pol l2_policies() {
  L2Pol p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  return p;
}
__device__ __forceinline__ double4 ld4h(const double4* p, unsigned long long pol) {
  double4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ double4 ldg4h(const double4* p, unsigned long long pol) {
  double4 v;
  asm("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st4h(double4* p, const double4 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ double ld1h(const double* p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ldu32h(const uint32_t* p, unsigned long long pol) {
  uint32_t v;
  asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
// L1 / L2 prefetch of the line holding *p: a load the thread will issue later finds it on chip; holds no register meanwhile
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
struct L2Pol {
  unsigned long long keep, stream;
};
inline L2Pol l2_policies() { return L2Pol{0, 0}; }
inline double4 ld4h(const double4* p, unsigned long long) { return *p; }
inline double4 ldg4h(const double4* p, unsigned long long) { return *p; }
inline void st4h(double4* p, const double4 v, unsigned long long) { *p = v; }
inline double ld1h(const double* p, unsigned long long) { return *p; }
inline uint32_t ldu32h(const uint32_t* p, unsigned long long) { return *p; }
inline void prefetch_l1(const void*) {}
inline void prefetch_l2(const void*) {}
inline double4 ld4(const double4* p) { return *p; }
inline double4 ldg4(const double4* p) { return *p; }
inline void st4(double4* p, const double4 v) { *p = v; }
#endif

EDB_HD v3 mk3(double x, double y, double z) { return v3{x, y, z}; }
EDB_HD v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
EDB_HD v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
EDB_HD v3 operator-(v3 a) { return {-a.x, -a.y, -a.z}; }
EDB_HD v3 operator*(v3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
EDB_HD v3 operator*(double s, v3 a) { return {s * a.x, s * a.y, s * a.z}; }
EDB_HD v3 operator/(v3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
EDB_HD v3 cmul(v3 a, v3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
EDB_HD double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
EDB_HD double norm(v3 a) { return sqrt(dot(a, a)); }
EDB_HD v3 cross(v3 u, v3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
EDB_HD bool is_zero(v3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }
EDB_HD v3 xyz(const double4& a) { return {a.x, a.y, a.z}; }
EDB_HD double4 mk4(v3 a, double w) { return make_double4(a.x, a.y, a.z, w); }

// rotation by a unit quaternion: src/shape/include/exaDEM/type/quat.hpp:79-90
EDB_HD v3 qrot(q4 q, v3 v) {
  const v3 u{q.x, q.y, q.z};
  v3 uv = cross(u, v);
  v3 uuv = cross(u, uv);
  uv = uv * (2.0 * q.w);
  uuv = uuv * 2.0;
  return v + uv + uuv;
}
EDB_HD q4 qconj(q4 q) { return {q.w, -q.x, -q.y, -q.z}; }
// quat.hpp:92-96
EDB_HD q4 qdot(q4 q, v3 o) {
  return {-0.5 * (o.x * q.x + o.y * q.y + o.z * q.z), 0.5 * (o.y * q.z - o.z * q.y + o.x * q.w),
          0.5 * (o.z * q.x - o.x * q.z + o.y * q.w), 0.5 * (o.x * q.y - o.y * q.x + o.z * q.w)};
}

struct Contact {
  bool hit;
  double dn;
  v3 n, cp;
};
EDB_HD Contact no_contact() { return Contact{false, 0.0, {0, 0, 0}, {0, 0, 0}}; }

// ---- src/shape/include/exaDEM/shape_detection_core.hpp ----
EDB_HD bool filter_vv(double rV, v3 vi, double ri, v3 vj, double rj) {  // :43-53
  const double R = ri + rj + rV;
  const v3 d = vi - vj;
  return dot(d, d) <= R * R;
}
EDB_HD Contact detect_vv(v3 pi, double ri, v3 pj, double rj) {  // :57-81
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
EDB_HD bool filter_ve(double rV, v3 vi, double ri, v3 vf, v3 vs, double rj) {  // :83-101
  const v3 fs = vs - vf, fi = vi - vf;
  const double r = dot(fs, fi) / dot(fs, fs);
  if (r < -rV || r > 1.0 + rV) return false;
  const v3 d = fi - fs * r;
  const double Rmax = ri + rj + rV;
  return dot(d, d) < Rmax * Rmax;
}
EDB_HD Contact detect_ve(v3 vi, double ri, v3 vf, v3 vs, double rj) {  // :122-148
  const v3 fs = vs - vf, fi = vi - vf;
  const double r = dot(fs, fi) / dot(fs, fs);
  if (r <= 0 || r >= 1.0) return no_contact();
  v3 n = fi - fs * r;
  const double dn = norm(n) - (ri + rj);
  if (dn > 0) return no_contact();
  n = n / norm(n);
  Contact c;
  c.hit = true;
  c.dn = dn;
  c.n = n;
  c.cp = vi - n * (ri + 0.5 * dn);
  return c;
}
// crossing-number point-in-polygon in the face frame, shared by filter (:245-309) and detection (:398-466).
// VA: any indexable returning v3 (world vertices of the face owner).
template <class VA>
EDB_HD bool face_test(v3 vi, const VA& va, const int* idx, int nf, double cut, double& dist_out, v3& n_out) {
  const v3 a = va[idx[0]], b = va[idx[1]], c = va[idx[nf - 1]];
  const v3 v = vi - a;
  v3 v1 = b - a;
  v3 v2 = c - a;
  v1 = v1 / norm(v1);
  v3 n = cross(v1, v2);
  n = n / norm(n);
  double dist = dot(n, v);
  if (dist < 0) {
    n = n * (-1.0);
    dist = -dist;
  }
  dist_out = dist;
  n_out = n;
  if (dist > cut) return false;
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
EDB_HD bool filter_vf(double rV, v3 vi, double ri, const VA& va, double rj, const int* idx, int nf) {
  double dist;
  v3 n;
  return face_test(vi, va, idx, nf, ri + rj + rV, dist, n);
}
template <class VA>
EDB_HD Contact detect_vf(v3 vi, double ri, const VA& va, double rj, const int* idx, int nf) {
  double dist;
  v3 n;
  const bool inside = face_test(vi, va, idx, nf, ri + rj, dist, n);
  if (dist > ri + rj) return no_contact();
  Contact c;
  c.dn = dist - (ri + rj);
  c.n = n;
  c.cp = vi - n * (ri + 0.5 * c.dn);
  c.hit = inside;
  return c;
}
EDB_HD bool ee_closest(v3 fi, v3 si, v3 fj, v3 sj, v3& pi, v3& pj) {  // :468-504 / :521-566 common part
  const v3 Ei = si - fi, Ej = sj - fj, v = fi - fj;
  const double c = dot(Ei, Ei), d = dot(Ej, Ej), e = dot(Ei, Ej);
  double f = (c * d) - (e * e);
  if (!(fabs(f) > 1.0e-12)) return false;
  f = 1.0 / f;
  const double a = dot(Ei, v), b = dot(Ej, v);
  const double s = (e * b - a * d) * f;
  const double t = (c * b - e * a) * f;
  if (s <= 0.0 || s >= 1.0 || t <= 0.0 || t >= 1.0) return false;
  pi = fi + Ei * s;
  pj = fj + Ej * t;
  return true;
}
EDB_HD bool filter_ee(double rV, v3 fi, v3 si, double ri, v3 fj, v3 sj, double rj) {
  const double R = ri + rj + rV;
  v3 pi, pj;
  if (!ee_closest(fi, si, fj, sj, pi, pj)) return false;
  const v3 n = pi - pj;
  return dot(n, n) <= R * R;
}
EDB_HD Contact detect_ee(v3 fi, v3 si, double ri, v3 fj, v3 sj, double rj) {
  const double R = ri + rj;
  v3 pi, pj;
  if (!ee_closest(fi, si, fj, sj, pi, pj)) return no_contact();
  v3 n = pi - pj;
  const double dn = norm(n) - R;
  if (dn > 0) return no_contact();
  n = n / norm(n);
  Contact c;
  c.hit = dn <= 0;
  c.dn = dn;
  c.n = n;
  c.cp = pi - n * (ri + 0.5 * dn);
  return c;
}

// ---- OBB: src/shape/include/exaDEM/type/OBB.hpp, basic_types.hpp:37-49 ----
struct Obb {
  v3 c, e1, e2, e3, ext;
};
EDB_HD Obb obb_world(const Obb& in, v3 pos, q4 q, double h) {
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
EDB_HD void obb_enlarge(Obb& o, double m) {
  o.ext.x += m;
  o.ext.y += m;
  o.ext.z += m;
}
EDB_HD Obb obb_point(v3 c, double m) { return Obb{c, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {m, m, m}}; }
// 15-axis SAT, OBB.hpp:90-222 (tol = FLT_EPSILON on every |R| term)
EDB_HD bool obb_intersect(const Obb& A, const Obb& B) {
  const double tol = 1.1920928955078125e-07;
  double R[3][3], Q[3][3];
  const v3 ea[3] = {A.e1, A.e2, A.e3}, eb[3] = {B.e1, B.e2, B.e3};
  const double xa[3] = {A.ext.x, A.ext.y, A.ext.z}, xb[3] = {B.ext.x, B.ext.y, B.ext.z};
  const v3 tt = A.c - B.c;
  const double t[3] = {dot(tt, A.e1), dot(tt, A.e2), dot(tt, A.e3)};
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      R[i][j] = dot(ea[i], eb[j]);
      Q[i][j] = fabs(R[i][j]) + tol;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double ra = xa[i];
    const double rb = xb[0] * Q[i][0] + xb[1] * Q[i][1] + xb[2] * Q[i][2];
    if (fabs(t[i]) > ra + rb) return false;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double ra = xa[0] * Q[0][j] + xa[1] * Q[1][j] + xa[2] * Q[2][j];
    const double rb = xb[j];
    if (fabs(t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j]) > ra + rb) return false;
  }
  double ra, rb;
  ra = xa[1] * Q[2][0] + xa[2] * Q[1][0]; rb = xb[1] * Q[0][2] + xb[2] * Q[0][1];
  if (fabs(t[2] * R[1][0] - t[1] * R[2][0]) > ra + rb) return false;
  ra = xa[1] * Q[2][1] + xa[2] * Q[1][1]; rb = xb[0] * Q[0][2] + xb[2] * Q[0][0];
  if (fabs(t[2] * R[1][1] - t[1] * R[2][1]) > ra + rb) return false;
  ra = xa[1] * Q[2][2] + xa[2] * Q[1][2]; rb = xb[0] * Q[0][1] + xb[1] * Q[0][0];
  if (fabs(t[2] * R[1][2] - t[1] * R[2][2]) > ra + rb) return false;
  ra = xa[0] * Q[2][0] + xa[2] * Q[0][0]; rb = xb[1] * Q[1][2] + xb[2] * Q[1][1];
  if (fabs(t[0] * R[2][0] - t[2] * R[0][0]) > ra + rb) return false;
  ra = xa[0] * Q[2][1] + xa[2] * Q[0][1]; rb = xb[0] * Q[1][2] + xb[2] * Q[1][0];
  if (fabs(t[0] * R[2][1] - t[2] * R[0][1]) > ra + rb) return false;
  ra = xa[0] * Q[2][2] + xa[2] * Q[0][2]; rb = xb[0] * Q[1][1] + xb[1] * Q[1][0];
  if (fabs(t[0] * R[2][2] - t[2] * R[0][2]) > ra + rb) return false;
  ra = xa[0] * Q[1][0] + xa[1] * Q[0][0]; rb = xb[1] * Q[2][2] + xb[2] * Q[2][1];
  if (fabs(t[1] * R[0][0] - t[0] * R[1][0]) > ra + rb) return false;
  ra = xa[0] * Q[1][1] + xa[1] * Q[0][1]; rb = xb[0] * Q[2][2] + xb[2] * Q[2][0];
  if (fabs(t[1] * R[0][1] - t[0] * R[1][1]) > ra + rb) return false;
  ra = xa[0] * Q[1][2] + xa[1] * Q[0][2]; rb = xb[0] * Q[2][1] + xb[1] * Q[2][0];
  if (fabs(t[1] * R[0][2] - t[0] * R[1][2]) > ra + rb) return false;
  return true;
}

// ---- contact law: src/forcefield/include/exaDEM/forcefield/{common_kernels,contact_force}.hpp ----
struct Params {
  double kn, kt, kr, mu, damp, fc, dncut, gamma;  // contact_parameters.hpp:32-41
};
enum { LAW_NONE = 0, LAW_COHESIVE = 1, LAW_DMT = 2 };

// contact_force.hpp:146-235 (LAW is a compile-time variant like the reference's template parameters)
template <int LAW>
__device__ __forceinline__ void contact_force(double dn, v3 n, double dt, const Params& p, double meff, double reff, v3& ft,
                                              v3 cp, v3 pos_i, v3 vel_i, v3& f, v3& mom, v3 w_i, v3 pos_j, v3 vel_j,
                                              v3 w_j) {
  if (LAW == LAW_COHESIVE) {
    if (dn >= 0) {  // :165-173
      const double fnv = (p.fc / p.dncut) * dn - p.fc;
      f = f + fnv * n;
      ft = {0, 0, 0};
      return;
    }
  }
  const double damp = p.damp * 2.0 * sqrt(p.kn * meff);  // common_kernels.hpp:28-31
  const v3 ci = vel_i - cross(cp - pos_i, w_i);            // :34-41
  const v3 cj = vel_j - cross(cp - pos_j, w_j);
  const v3 vel = cj - ci;
  const double vn = dot(vel, n);
  const double fne = -p.kn * dn, fnv = damp * vn;          // :48-53
  const double fs = (fnv + fne > 0) ? fnv + fne : 0;
  v3 fn;
  if (LAW == LAW_COHESIVE) {                               // :69-76
    const double a = fs - p.fc;
    const double b = a < (-p.fc) ? -p.fc : a;
    fn = b * n;
  } else {
    fn = fs * n;
  }
  const v3 vt = vel - vn * n;                              // :80-90
  ft = ft + p.kt * (dt * vt);
  double thr;
  if (LAW != LAW_COHESIVE) thr = fabs(p.mu * (p.kn * dn));  // :108-114
  else thr = fabs(p.mu * (dot(fn, n) + p.fc));              // :116-120
  const double ft2 = dot(ft, ft);                          // :122-128
  if (ft2 > 0.0 && ft2 > thr * thr) ft = ft * (thr / sqrt(ft2));
  f = fn + ft;
  if (LAW == LAW_DMT) {                                    // contact_force.hpp:115-127
    if (p.gamma > 0.0 && dn < 0.0) {
      const double Fd = 2.0 * 3.14159265358979323846 * reff * p.gamma;
      f = f - Fd * n;
    }
  }
  mom = mom + p.kr * (w_j - w_i) * dt;                     // :223
  v3 branch = cp - pos_i;
  const double r = dot(branch, w_i) / dot(w_i, w_i);       // NaN when w_i == 0 (quirk Q1): keep IEEE semantics
  branch = branch - r * w_i;
  // The reference hard-codes mur = 0 (contact_force.hpp:229-234):  thm = |0 * norm(branch) * norm(fn)|  is +0 unless one of the
  // norms is NaN or Inf (quirk Q1: w_i = 0 makes r, hence branch, NaN), in which case it is NaN and `m2 > thm * thm` is false.
  // With thm = +0 the update is mom * (0 / sqrt(m2)) = mom * (+0) for every m2 > 0.  Evaluated that way — bit-identical,
  // three square roots and one division fewer per contact.
  const double nb2 = dot(branch, branch), nf2 = dot(fn, fn);
  const bool thm_is_zero = nb2 <= 1.7976931348623157e308 && nf2 <= 1.7976931348623157e308;   // false for NaN and +Inf
  const double m2 = dot(mom, mom);
  if (thm_is_zero && m2 > 0.0) mom = mom * 0.0;
}
EDB_HD v3 moments(v3 cp, v3 p, v3 f, v3 m) { return cross(cp - p, f) + m; }  // common_kernels.hpp:137-145

// ---- drivers: src/driver/include/exaDEM/{cylinder,surface,ball}.hpp ----
struct Driver {
  int kind;         // EDB_DRIVER_*
  int need_forces;
  double radius;
  v3 axis, center;  // cylinder center already multiplied by axis (cylinder.hpp:104)
  double offset;
  v3 normal, center_proj;
  v3 vel, vrot;
  int shape;
  int pad;
};
EDB_HD bool driver_filter(const Driver& d, double rcut, v3 p) {
  if (d.kind == 0) {  // cylinder.hpp:225-235
    const v3 dir = cmul(p, d.axis) - d.center;
    const double dn = d.radius - (norm(dir) + rcut);
    return dn <= 0;
  } else if (d.kind == 1) {  // surface.hpp:320-324
    const v3 proj = dot(p, d.normal) * d.normal;
    return norm(proj - d.center_proj) <= rcut;
  } else {  // ball.hpp:292-296
    const v3 dist = d.center - p;
    const double dd = d.radius - norm(dist);
    return fabs(dd) <= rcut;
  }
}
EDB_HD Contact driver_detect(const Driver& d, double rcut, v3 p) {
  Contact c = no_contact();
  if (d.kind == 0) {  // cylinder.hpp:255-279
    const v3 dir = d.center - cmul(p, d.axis);
    const double len = norm(dir);
    c.dn = d.radius - (rcut + len);
    if (c.dn > 0) return c;
    c.n = dir / len;
  } else if (d.kind == 1) {  // surface.hpp:338-350
    const v3 proj = dot(p, d.normal) * d.normal;
    const v3 s2p = -(d.center_proj - proj);
    const double len = norm(s2p);
    c.dn = len - rcut;
    if (c.dn > 0) return c;
    c.n = s2p / len;
  } else {  // ball.hpp:309-327
    const v3 p2c = d.center - p;
    const double len = norm(p2c);
    v3 n = p2c / len;
    if (len > d.radius) {
      c.dn = len - d.radius - rcut;
      n = (-1.0) * n;
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

}  // namespace edb
