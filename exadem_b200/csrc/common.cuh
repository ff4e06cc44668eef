// common.cuh — context, device buffers and launch helpers of libexadem_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/exadem_b200.h"
#include "dem_math.cuh"

namespace edb {

constexpr int NT = EDB_NTYPES;
constexpr int BLK = 256;

struct CudaError {
  cudaError_t e;
  const char* what;
  const char* file;
  int line;
};

#define CK(call)                                                             \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) throw edb::CudaError{_e, #call, __FILE__, __LINE__}; \
  } while (0)

struct StateError {
  int code;
  std::string msg;
};

// growable device buffer (capacity only grows; contents are NOT preserved on growth unless keep=true)
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  void ensure(size_t n, bool keep = false, cudaStream_t s = 0) {
    if (n <= cap) return;
    size_t ncap = n + n / 4 + 64;
    T* np = nullptr;
    CK(cudaMalloc(&np, ncap * sizeof(T)));
    if (keep && p && cap) {
      CK(cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
      CK(cudaStreamSynchronize(s));
    }
    if (p) CK(cudaFree(p));
    p = np;
    cap = ncap;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// ---- particle field set on the device -------------------------------------------------------------------------
// Gather-friendly packing: every per-particle quantity the contact kernels gather is one aligned 32-byte sector.
//   pr = (rx, ry, rz, radius)   vm = (vx, vy, vz, mass)   wg = (vrotx, vroty, vrotz, bits{type | group<<32})
//   q  = (w, x, y, z)
// Streaming-only fields stay scalar SoA (perfectly coalesced in the per-particle kernels).
struct PArr {
  DevBuf<double4> pr, vm, wg, q;
  DevBuf<double> f[3], m[3], a[3], I[3], h;
  DevBuf<uint64_t> id;
  DevBuf<uint32_t> orig;  // device slot -> index in the host (upload) order
  void ensure(size_t n, bool keep, cudaStream_t s) {
    pr.ensure(n, keep, s); vm.ensure(n, keep, s); wg.ensure(n, keep, s); q.ensure(n, keep, s);
    for (int k = 0; k < 3; k++) { f[k].ensure(n, keep, s); m[k].ensure(n, keep, s); a[k].ensure(n, keep, s); I[k].ensure(n, keep, s); }
    h.ensure(n, keep, s); id.ensure(n, keep, s); orig.ensure(n, keep, s);
  }
  void release() {
    pr.release(); vm.release(); wg.release(); q.release();
    for (int k = 0; k < 3; k++) { f[k].release(); m[k].release(); a[k].release(); I[k].release(); }
    h.release(); id.release(); orig.release();
  }
};
// raw-pointer view passed to kernels by value
struct PView {
  double4 *pr, *vm, *wg, *q;
  double *fx, *fy, *fz, *mx, *my, *mz, *ax, *ay, *az, *ix, *iy, *iz, *h;
  uint64_t* id;
  uint32_t* orig;
};
inline PView view(PArr& a) {
  return PView{a.pr.p, a.vm.p, a.wg.p, a.q.p, a.f[0].p, a.f[1].p, a.f[2].p, a.m[0].p, a.m[1].p, a.m[2].p,
               a.a[0].p, a.a[1].p, a.a[2].p, a.I[0].p, a.I[1].p, a.I[2].p, a.h.p, a.id.p, a.orig.p};
}

// ---- one interaction wave (type) : SoA, sorted by owner slot then enumeration order ---------------------------
//   ij   = (i, j) device slots in the reference's i/j roles (j = driver index for driver waves)
//   sub  = (sub_i, sub_j)                       [polyhedra / mesh drivers only]
//   flg  = bit0 swap, bits1-2 ghost
//   ft   = (friction.xyz, bits{partner id})     history + the key used to carry it across rebuilds
//   mo   = (moment.xyz, 0)                      history
//   fc   = the wrench of the last evaluation, 64 bytes per record: fc[2k] = fd = (f.xyz, dn), fc[2k+1] = cp = (contact point.xyz, 0)
//          ContactState (fn is derived as f - ft on export).  One aligned 64-byte unit: the contact kernel writes it with two
//          adjacent 32-byte stores and the gather fetches both halves of a partner's wrench from one 128-byte line.
struct Wave {
  DevBuf<uint2> ij, sub;
  DevBuf<uint8_t> flg;
  DevBuf<double4> ft, mo, fc;
  DevBuf<uint32_t> off;  // nreal+1 : CSR by owner slot
  uint64_t n = 0;
  void ensure(size_t nn, bool need_sub) {
    ij.ensure(nn); if (need_sub) sub.ensure(nn); flg.ensure(nn); ft.ensure(nn); mo.ensure(nn); fc.ensure(2 * nn);
  }
  void ensure_keep(size_t nn, bool need_sub, cudaStream_t s) {  // grows and preserves the records (twins are appended)
    ij.ensure(nn, true, s); if (need_sub) sub.ensure(nn, true, s); flg.ensure(nn, true, s); ft.ensure(nn, true, s); mo.ensure(nn, true, s);
    fc.ensure(2 * nn, true, s);
  }
  void release() { ij.release(); sub.release(); flg.release(); ft.release(); mo.release(); fc.release(); off.release(); n = 0; }
};
struct WView {
  uint2 *ij, *sub;
  uint8_t* flg;
  double4 *ft, *mo, *fc;
  uint32_t* off;
  uint32_t n;
  __host__ __device__ __forceinline__ double4* fd(size_t k) const { return fc + 2 * k; }
  __host__ __device__ __forceinline__ double4* cp(size_t k) const { return fc + 2 * k + 1; }
};
inline WView view(Wave& w) { return WView{w.ij.p, w.sub.p, w.flg.p, w.ft.p, w.mo.p, w.fc.p, w.off.p, (uint32_t)w.n}; }

// shapes flattened for the device
struct ShapeDev {
  int nv, ne, nf, pad;
  int v0, e0, f0, fi0;  // offsets into the pooled arrays (vertices, edges(pairs), face table, face indices)
  Obb obb;
  double radius;
  double vmax;   // largest vertex norm (homothety 1): bound of the vertex displacement under a rotation
};
struct ShapesView {
  const ShapeDev* sh;
  const v3* verts;     // local vertices
  const int* edges;    // 2 ints per edge
  const int* foff;     // per face: offset into fidx (relative to shape's fi0)
  const int* fn;       // per face: vertex count
  const int* fidx;     // face vertex indices
  int nvmax;
};

// RShape (STL mesh) drivers, device view.  One MeshDev per RSHAPE driver (Driver::shape = its slot): world vertices are
// RShapeDriver::vertices_ = position + orientation * v (rshape.hpp:381-384), computed when the drivers are set.
// cell_off/cell_data: per (mesh, grid cell, kind in {vertex, edge, face}) the primitives that can reach a particle of the cell —
// the role of the reference's RShapeDriverGridCellIndexes (rshape_grid.hpp, grid_stl_mesh.cpp:150-251).
struct MeshDev {
  int nv, ne, nf, drv;  // drv = index of the owning driver
  int v0, e0, f0, fi0;  // offsets into the pooled mesh arrays
  double radius;        // Minkowski radius (homothety 1)
};
struct MeshView {
  const MeshDev* md;
  const double4* verts;   // world vertices
  const int* edges;
  const int* foff;
  const int* fn;
  const int* fidx;
  const uint32_t* cell_off;  // [(m * ncell + cell) * 3 + kind], nmesh * ncell * 3 + 1 entries
  const int* cell_data;
  int nmesh;
  uint32_t ncell;
};

struct GridDev {
  int nc[3], g[3];
  double lo[3], cs[3];
  double glo[3], ghi[3];
  int periodic[3];
  uint32_t ncell;
};

struct ParamsDev {
  Params pp, pd;          // single-material
  const Params* tpp;      // multi-material tables (nullptr when single)
  const Params* tpd;
  int n_groups, n_drivers;
};

struct EventPair {
  cudaEvent_t a, b;
  int bucket;
};

}  // namespace edb
