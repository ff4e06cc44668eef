// Shim for the un-vendored Onika math types (onika/math/basic_types.h). TEST INFRASTRUCTURE ONLY.
// Written from the call-site contract in the reference; the real Onika header is
// not available here, so the operator definitions below are the *working definition* (parity risk R1).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <iostream>
#include <vector>
#include <string>
#include <span>
#include <tuple>
#include <cassert>

#ifndef ONIKA_HOST_DEVICE_FUNC
#define ONIKA_HOST_DEVICE_FUNC
#endif
#ifndef ONIKA_ASSUME_ALIGNED
#define ONIKA_ASSUME_ALIGNED(x)
#endif

namespace onika { namespace math {
struct Vec3d { double x, y, z; };
struct Mat3d { double m11, m12, m13, m21, m22, m23, m31, m32, m33; };
struct Quaternion { double w, x, y, z; };
struct AABB { Vec3d bmin, bmax; };
struct IJK { ssize_t i, j, k; };

inline bool operator==(const Vec3d& a, const Vec3d& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const Vec3d& a, const Vec3d& b) { return !(a == b); }
inline Vec3d operator+(const Vec3d& a, const Vec3d& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3d operator-(const Vec3d& a, const Vec3d& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3d operator-(const Vec3d& a) { return {-a.x, -a.y, -a.z}; }
// Vec3d - double (used by RShapeDriver::force_to_accel, rshape.hpp:230): working definition = component-wise (risk R1)
inline Vec3d operator-(const Vec3d& a, double s) { return {a.x - s, a.y - s, a.z - s}; }
inline Vec3d operator*(const Vec3d& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3d operator*(double s, const Vec3d& a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3d operator*(const Vec3d& a, const Vec3d& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3d operator/(const Vec3d& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3d& operator+=(Vec3d& a, const Vec3d& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline Vec3d& operator-=(Vec3d& a, const Vec3d& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline Vec3d& operator*=(Vec3d& a, double s) { a.x *= s; a.y *= s; a.z *= s; return a; }
inline Vec3d& operator/=(Vec3d& a, double s) { a.x /= s; a.y /= s; a.z /= s; return a; }
inline double dot(const Vec3d& a, const Vec3d& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm2(const Vec3d& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double norm(const Vec3d& a) { return std::sqrt(norm2(a)); }
inline Vec3d cross(const Vec3d& u, const Vec3d& v) {
  return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
}
inline Vec3d operator*(const Mat3d& m, const Vec3d& v) {
  return {m.m11 * v.x + m.m12 * v.y + m.m13 * v.z, m.m21 * v.x + m.m22 * v.y + m.m23 * v.z,
          m.m31 * v.x + m.m32 * v.y + m.m33 * v.z};
}
inline std::ostream& operator<<(std::ostream& o, const Vec3d& v) { return o << v.x << ' ' << v.y << ' ' << v.z; }

// Quaternion rotation: same family as the vendored Rockable quat.hpp:79-90 (v + 2s(q x v) + 2 q x (q x v)).
inline Vec3d operator*(const Quaternion& q, const Vec3d& v) {
  const Vec3d qv{q.x, q.y, q.z};
  Vec3d uv = cross(qv, v);
  Vec3d uuv = cross(qv, uv);
  uv *= (2.0 * q.w);
  uuv *= 2.0;
  return v + uv + uuv;
}
inline Quaternion get_conjugated(const Quaternion& q) { return {q.w, -q.x, -q.y, -q.z}; }
// time derivative 1/2 * omega (x) q, as quat.hpp:92-96
inline Quaternion dot(const Quaternion& q, const Vec3d& o) {
  return {-0.5 * (o.x * q.x + o.y * q.y + o.z * q.z), 0.5 * (o.y * q.z - o.z * q.y + o.x * q.w),
          0.5 * (o.z * q.x - o.x * q.z + o.y * q.w), 0.5 * (o.x * q.y - o.y * q.x + o.z * q.w)};
}
inline Quaternion operator+(const Quaternion& a, const Quaternion& b) { return {a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Quaternion operator*(const Quaternion& a, double s) { return {a.w * s, a.x * s, a.y * s, a.z * s}; }
inline Quaternion normalize(const Quaternion& q) {
  const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
}}  // namespace onika::math

namespace exanb {
using onika::math::Vec3d; using onika::math::Mat3d; using onika::math::Quaternion; using onika::math::AABB; using onika::math::IJK;
using onika::math::dot; using onika::math::norm; using onika::math::norm2; using onika::math::cross;
using onika::math::get_conjugated; using onika::math::normalize;
}
