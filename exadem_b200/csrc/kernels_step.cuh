// kernels_step.cuh — per-time-step kernels (sm_100a, fp64, HBM-bound gather/scatter; no tensor cores by design).
//
//   k_prolog / k_epilog ...   streaming per-particle operators (src/numericalscheme/*)
//   k_contact_*               one thread per interaction record of a wave (classifier_for_all.hpp:116-156):
//                             gather the two particles (3 aligned 32-byte sectors each), detect, Hooke law,
//                             update the friction/moment history in place, emit (f, dn, cp)
//   k_gather                  one thread per particle: deterministic, fixed-order sum of the wrenches of all records
//                             incident to the particle (owner segment = contiguous; partner side = incidence list)
//                             — replaces the reference's 12 atomic adds per contact (contact_sphere.hpp:183-196)
#pragma once
#include "common.cuh"
#include "kernels_rebuild.cuh"
#include "comm.cuh"

namespace edb {

// ---- tuning switches (profiles/r2_sweep_*.log: measured on the 1 M-sphere drum, 200 steps) ---------------------------------------
// None of the latency / residency devices below pays on this workload; they stay as switches so that the measurement can be repeated:
//   EDB_G_PREFETCH   L1 prefetch of the gather's own sectors used after the incidence walk        gather 0.115 -> 0.120 ms
//   EDB_G_WIDE       all incidence codes, then all flag bytes, then the wrenches (EDB_G_CU wide)  0.115 -> 0.114 (6), 0.124 (8: spills)
//   EDB_PP_PREFETCH  L1 prefetch of the next law round's operands in the contact kernel           contact 0.104 -> 0.104 ms
//   EDB_L2_KEEP      pr / vm / wg evict_last (createpolicy + .L2::cache_hint) in both kernels     step 0.2486 -> 0.2488 ms
//   EDB_L2_STREAM    once-per-step data (history, codes, inertia, backup, orientation) evict_first  0.2486 -> 0.2489 ms
//   EDB_G_REVERSE    gather walks the blocks from last to first (what the contact kernel wrote last)  0.2486 -> 0.2491 ms
#ifndef EDB_GU
#define EDB_GU 2
#endif
#ifndef EDB_G_PREFETCH
#define EDB_G_PREFETCH 0
#endif
#ifndef EDB_G_WIDE
#define EDB_G_WIDE 0
#endif
#ifndef EDB_G_CU
#define EDB_G_CU 6
#endif
#ifndef EDB_PP_PREFETCH
#define EDB_PP_PREFETCH 0
#endif
#ifndef EDB_L2_KEEP
#define EDB_L2_KEEP 0
#endif
#ifndef EDB_L2_STREAM
#define EDB_L2_STREAM 0
#endif
#ifndef EDB_G_REVERSE
#define EDB_G_REVERSE 0
#endif

__device__ __forceinline__ uint32_t group_of(const double4& wg) {
  return (uint32_t)(((unsigned long long)__double_as_longlong(wg.w)) >> 32);
}
__device__ __forceinline__ uint32_t type_of(const double4& wg) {
  return (uint32_t)(((unsigned long long)__double_as_longlong(wg.w)) & 0xffffffffu);
}
__device__ __forceinline__ const Params& cp_pp(const ParamsDev& P, uint32_t ga, uint32_t gb) {
  return P.tpp ? P.tpp[ga * P.n_groups + gb] : P.pp;
}
__device__ __forceinline__ const Params& cp_pd(const ParamsDev& P, uint32_t ga, uint32_t d) {
  return P.tpd ? P.tpd[ga * P.n_drivers + d] : P.pd;
}

// ---------------------------------------------------------------------------------------------------------------
// streaming operators
// ---------------------------------------------------------------------------------------------------------------
// combined_compute_prolog.cu:35-47: r += v dt + a dt^2/2 ; v += a dt/2 ; Q = normalize(Q + dot(Q,w) dt) ; w += arot dt/2
// (fx,fy,fz hold the ACCELERATION here: force_to_accel ran at the end of the previous step.)
__global__ void k_prolog(PView P, uint32_t n, double dt, double dt2, double hdt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double4 r = ld4(P.pr + p), v = ld4(P.vm + p), w = ld4(P.wg + p), Q = ld4(P.q + p);
  const double ax = P.fx[p], ay = P.fy[p], az = P.fz[p];
  r.x += v.x * dt + ax * dt2; r.y += v.y * dt + ay * dt2; r.z += v.z * dt + az * dt2;
  v.x += ax * hdt; v.y += ay * hdt; v.z += az * hdt;
  const q4 q{Q.x, Q.y, Q.z, Q.w};
  const q4 dq = qdot(q, mk3(w.x, w.y, w.z));
  q4 nq{q.w + dq.w * dt, q.x + dq.x * dt, q.y + dq.y * dt, q.z + dq.z * dt};
  const double nn = sqrt(nq.w * nq.w + nq.x * nq.x + nq.y * nq.y + nq.z * nq.z);
  Q = make_double4(nq.w / nn, nq.x / nn, nq.y / nn, nq.z / nn);
  const v3 wn = mk3(w.x, w.y, w.z) + hdt * mk3(P.ax[p], P.ay[p], P.az[p]);
  w.x = wn.x; w.y = wn.y; w.z = wn.z;
  st4(P.pr + p, r); st4(P.vm + p, v); st4(P.wg + p, w); st4(P.q + p, Q);
}
// push_to_quaternion.hpp:26-38
__global__ void k_push_to_quaternion(PView P, uint32_t n, double dt, double hdt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double4 w = P.wg[p], Q = P.q[p];
  const q4 q{Q.x, Q.y, Q.z, Q.w};
  const q4 dq = qdot(q, mk3(w.x, w.y, w.z));
  q4 nq{q.w + dq.w * dt, q.x + dq.x * dt, q.y + dq.y * dt, q.z + dq.z * dt};
  const double nn = sqrt(nq.w * nq.w + nq.x * nq.x + nq.y * nq.y + nq.z * nq.z);
  P.q[p] = make_double4(nq.w / nn, nq.x / nn, nq.y / nn, nq.z / nn);
  const v3 wn = mk3(w.x, w.y, w.z) + hdt * mk3(P.ax[p], P.ay[p], P.az[p]);
  w.x = wn.x; w.y = wn.y; w.z = wn.z;
  P.wg[p] = w;
}
// angular_acceleration.hpp:27-43
__device__ __forceinline__ v3 ang_accel(const double4& Q, v3 mom, v3 w, v3 I) {
  const q4 q{Q.x, Q.y, Q.z, Q.w};
  const q4 qi = qconj(q);
  const v3 o = qrot(qi, w), M = qrot(qi, mom);
  const v3 dw = mk3((M.x - (I.z - I.y) * o.y * o.z) / I.x, (M.y - (I.x - I.z) * o.z * o.x) / I.y,
                    (M.z - (I.y - I.x) * o.x * o.y) / I.z);
  return qrot(q, dw);
}
// MODE 0: push_to_angular_acceleration; 1: push_to_angular_velocity; 2: combined_compute_epilog (.cu:35-46)
template <int MODE>
__global__ void k_epilog(PView P, uint32_t n, double hdt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double4 w = P.wg[p];
  v3 ar;
  if (MODE == 0 || MODE == 2) {
    ar = ang_accel(P.q[p], mk3(P.mx[p], P.my[p], P.mz[p]), mk3(w.x, w.y, w.z), mk3(P.ix[p], P.iy[p], P.iz[p]));
    P.ax[p] = ar.x; P.ay[p] = ar.y; P.az[p] = ar.z;
  } else {
    ar = mk3(P.ax[p], P.ay[p], P.az[p]);
  }
  if (MODE == 1 || MODE == 2) {
    const v3 wn = mk3(w.x, w.y, w.z) + ar * hdt;  // angular_velocity.hpp:27-33
    w.x = wn.x; w.y = wn.y; w.z = wn.z;
    P.wg[p] = w;
  }
  if (MODE == 2) {
    double4 v = P.vm[p];
    v.x += P.fx[p] * hdt; v.y += P.fy[p] * hdt; v.z += P.fz[p] * hdt;
    P.vm[p] = v;
  }
}
__global__ void k_reset_fm(PView P, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  P.fx[p] = 0; P.fy[p] = 0; P.fz[p] = 0; P.mx[p] = 0; P.my[p] = 0; P.mz[p] = 0;
}
__global__ void k_gravity(PView P, uint32_t n, double gx, double gy, double gz) {  // gravity_force.hpp:38-46
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double m = P.vm[p].w;
  P.fx[p] += gx * m; P.fy[p] += gy * m; P.fz[p] += gz * m;
}
// reset_force_moment + gravity_force in one pass (same expressions: 0 + g m), gated like the other kernels of a speculative step
__global__ void k_reset_gravity(PView P, uint32_t n, double gx, double gy, double gz, const unsigned long long* gate) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n || (gate && *gate)) return;
  const double m = P.vm[p].w;
  P.fx[p] = 0.0 + gx * m; P.fy[p] = 0.0 + gy * m; P.fz[p] = 0.0 + gz * m;
  P.mx[p] = 0; P.my[p] = 0; P.mz[p] = 0;
}
__global__ void k_force_to_accel(PView P, uint32_t n) {  // force_to_accel.hpp:26-33
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double m = P.vm[p].w;
  P.fx[p] /= m; P.fy[p] /= m; P.fz[p] /= m;
}
__global__ void k_freeze(PView P, uint32_t n) {  // freeze_particles.cu:33-88 (whole grid)
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  P.fx[p] = 0; P.fy[p] = 0; P.fz[p] = 0; P.mx[p] = 0; P.my[p] = 0; P.mz[p] = 0;
  double4 v = P.vm[p], w = P.wg[p];
  v.x = v.y = v.z = 0; w.x = w.y = w.z = 0;
  P.vm[p] = v; P.wg[p] = w;
}
// particle regions (exanb ParticleRegion / ParticleRegionCSG; the `region` slot of reset_force_moment.cu:36-57, freeze_particles.cu:47-66)
struct RegionDev {
  double lo[3], hi[3], Q[16];
  unsigned long long id0, id1;
  int use_bounds, use_ids, use_quadric, pad;
};
struct RegionExpr {
  const RegionDev* reg;
  int n;
  int tok[EDB_CSG_MAX];
};
__device__ __forceinline__ bool in_region(const RegionDev& R, double x, double y, double z, unsigned long long id) {
  if (R.use_bounds && !(x >= R.lo[0] && x < R.hi[0] && y >= R.lo[1] && y < R.hi[1] && z >= R.lo[2] && z < R.hi[2])) return false;
  if (R.use_ids && !(id >= R.id0 && id < R.id1)) return false;
  if (R.use_quadric) {
    const double v[4] = {x, y, z, 1.0};
    double s = 0.0;
    for (int i = 0; i < 4; i++) {
      double t = 0.0;
      for (int j = 0; j < 4; j++) t += R.Q[4 * i + j] * v[j];
      s += v[i] * t;
    }
    if (!(s <= 0.0)) return false;
  }
  return true;
}
__device__ __forceinline__ bool csg_contains(const RegionExpr& E, double x, double y, double z, unsigned long long id) {
  if (E.n == 0) return true;
  unsigned int stack = 0;   // bit k = k-th entry from the bottom
  int sp = 0;
  for (int q = 0; q < E.n; q++) {
    const int t = E.tok[q];
    if (t >= 0) { stack = (stack & ~(1u << sp)) | ((in_region(E.reg[t], x, y, z, id) ? 1u : 0u) << sp); sp++; }
    else if (t == EDB_CSG_NOT) { stack ^= 1u << (sp - 1); }
    else {
      const unsigned a = (stack >> (sp - 2)) & 1u, b = (stack >> (sp - 1)) & 1u;
      const unsigned r = t == EDB_CSG_AND ? (a & b) : (a | b);
      sp--;
      stack = (stack & ~(1u << (sp - 1))) | (r << (sp - 1));
    }
  }
  return stack & 1u;
}
template <bool FREEZE>
__global__ void k_region_set(PView P, uint32_t n, RegionExpr E) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double4 r = P.pr[p];
  if (!csg_contains(E, r.x, r.y, r.z, P.id[p])) return;
  P.fx[p] = 0; P.fy[p] = 0; P.fz[p] = 0; P.mx[p] = 0; P.my[p] = 0; P.mz[p] = 0;
  if (FREEZE) {
    double4 v = P.vm[p], w = P.wg[p];
    v.x = v.y = v.z = 0; w.x = w.y = w.z = 0;
    P.vm[p] = v; P.wg[p] = w;
  }
}
// compute_vertices.hpp:26-67: world vertex k of particle p = r + Q*(h*v_k)  (shape.hpp:259-262)
// The 4th component of every world vertex carries the particle's Minkowski radius h * shape.radius, and ptype[p] the shape index: the
// detection pass of the polyhedron contact operator then needs no particle sector at all (it fetched wg and homothety of both
// particles of every record for these two numbers: 128 B of sectors per record, profiles/r2_ncu_polystep.txt).
__global__ void k_compute_vertices(PView P, uint32_t first, uint32_t n, ShapesView S, double4* __restrict__ vert, uint16_t* __restrict__ ptype) {
  const uint32_t p = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double4 r = P.pr[p], Q = P.q[p];
  const uint32_t ty = type_of(P.wg[p]);
  const ShapeDev sh = S.sh[ty];
  const double h = P.h[p];
  const double rm = h * sh.radius;
  const q4 q{Q.x, Q.y, Q.z, Q.w};
  if (ptype) ptype[p] = (uint16_t)ty;
  for (int k = 0; k < sh.nv; k++) {
    const v3 w = mk3(r.x, r.y, r.z) + qrot(q, h * S.verts[sh.v0 + k]);
    vert[(size_t)p * S.nvmax + k] = make_double4(w.x, w.y, w.z, rm);
  }
}
// particle_displ_over / vertex_displ_over (backup_dem.hpp:66-100): count of particles (vertices) with |d|^2 >= thr^2
template <bool POLY>
__global__ void k_displ_over(PView P, const double4* __restrict__ pr_b, const double4* __restrict__ q_b, uint32_t n,
                             ShapesView S, double thr2, unsigned long long* __restrict__ count, const unsigned long long* gate) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (gate && *gate) return;
  unsigned int c = 0;
  if (p < n) {
    const double4 r = P.pr[p], rb = pr_b[p];
    if (!POLY) {
      const v3 d = mk3(r.x, r.y, r.z) - mk3(rb.x, rb.y, rb.z);
      c = dot(d, d) >= thr2 ? 1 : 0;
    } else {
      const ShapeDev sh = S.sh[type_of(P.wg[p])];
      const double h = P.h[p];
      const double4 Q = P.q[p], Qb = q_b[p];
      const q4 q{Q.x, Q.y, Q.z, Q.w}, qb{Qb.x, Qb.y, Qb.z, Qb.w};
      for (int k = 0; k < sh.nv; k++) {
        const v3 lv = h * S.verts[sh.v0 + k];
        const v3 vo = mk3(rb.x, rb.y, rb.z) + qrot(qb, lv);
        const v3 vn = mk3(r.x, r.y, r.z) + qrot(q, lv);
        const v3 d = vn - vo;
        c += dot(d, d) >= thr2 ? 1 : 0;
      }
    }
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

// ---------------------------------------------------------------------------------------------------------------
// contact waves
// ---------------------------------------------------------------------------------------------------------------
struct ContactArgs {
  PView P;
  WView w;
  ParamsDev prm;
  const Driver* drv;
  const double4* vert;
  ShapesView S;
  double dt;
  int type;
  int symetric;
  MeshView M;
  // speculative step (edb_run_verlet): non-zero *gate = "this step needs a neighbour rebuild first" -> the kernel does nothing
  const unsigned long long* gate;
  unsigned long long* zero_flag;   // spheres, speculative loop: the next trigger slot, zeroed by the wave-0 kernel (saves a memset)
  // multi-GPU, inside edb_run_verlet: the ghosts of neighbour bricks are read where the peers' gathers stored them — this rank's
  // landing zone of the step's parity (comm.cuh) — instead of being copied into the ghost slots first.  Slots >= rem_lo are remote
  // ghosts; g_pr / g_vm / g_wg are already offset by -rem_lo.  rem_lo = 0xffffffff: everything is read from the particle arrays.
  const double4 *g_pr, *g_vm, *g_wg;
  uint32_t rem_lo;
  const uint16_t* ptype;   // polyhedra: shape index per slot (k_compute_vertices)
};
__device__ __forceinline__ bool gated(const unsigned long long* gate) { return gate && *gate; }

// Record state bit FLG_HOT (flg bit 3): the record was in contact at its last evaluation, i.e. its history (ft, mo) and its
// wrench (fd, cp) are live.  Invariant: !HOT  =>  friction = moment = 0 (contact_sphere.hpp:197-200 resets them) and fd/cp are
// never read (the gather tests the bit; the ContactState export recomputes cp).  A record that is not in contact and was not
// hot therefore costs 9 bytes (ij + flg) and two position gathers — no history read, no write.
// FLG_MOM (flg bit 4): the rolling-moment history `mo` is non-zero.  With the reference's degenerate clamp (quirk Q1,
// contact_force.hpp:229-234: mur = 0) the moment is zero for every contact whose particle i rotates, so the mo sector is
// neither read nor written by the contact kernel, nor read by the gather, unless the bit says it holds something.
constexpr uint8_t FLG_SWAP = 1, FLG_HOT = 8, FLG_MOM = 16;

struct Hist {
  double4 ft, mo;
};

// finish one record.  `contact` already includes the cohesive extension.  H is valid when `have_h`.
template <int LAW, bool HINT = false>
__device__ __forceinline__ void finish_record(const ContactArgs& A, uint32_t k, uint8_t fl, bool contact, const Contact& c,
                                              const Params& cp, double meff, double reff, v3 ri, v3 vi, v3 wi, v3 pj, v3 vj, v3 wj,
                                              const Hist& H, unsigned long long pol_stream = 0) {
  uint8_t nf = fl;
  if (contact) {
    v3 ft = mk3(H.ft.x, H.ft.y, H.ft.z), mom = mk3(H.mo.x, H.mo.y, H.mo.z), f = mk3(0, 0, 0);
    contact_force<LAW>(c.dn, c.n, A.dt, cp, meff, reff, ft, c.cp, ri, vi, f, mom, wi, pj, vj, wj);
    if (HINT) st4h(A.w.ft + k, make_double4(ft.x, ft.y, ft.z, H.ft.w), pol_stream);   // the history is next read one step from now
    else st4(A.w.ft + k, make_double4(ft.x, ft.y, ft.z, H.ft.w));
    st4(A.w.fd(k), make_double4(f.x, f.y, f.z, c.dn));
    st4(A.w.cp(k), make_double4(c.cp.x, c.cp.y, c.cp.z, 0));
    const bool live = !is_zero(mom);   // NaN components count as live (x == 0.0 is false for NaN)
    if (live || (fl & FLG_MOM)) st4(A.w.mo + k, make_double4(mom.x, mom.y, mom.z, 0.0));
    nf = (uint8_t)((fl | FLG_HOT) & ~FLG_MOM) | (live ? FLG_MOM : 0);
  } else if (fl & FLG_HOT) {
    // item.reset(); dn = 0 (contact_sphere.hpp:197-200) — friction & moment history are dropped
    st4(A.w.ft + k, make_double4(0, 0, 0, H.ft.w));
    if (fl & FLG_MOM) st4(A.w.mo + k, make_double4(0, 0, 0, 0));
    nf = fl & (uint8_t)~(FLG_HOT | FLG_MOM);
  }
  if (nf != fl) A.w.flg[k] = nf;
}


// wave 0 of the sphere operator: sphere::ContactLawFunc (contact_sphere.hpp:121-203).
//
// Persistent kernel, warp-autonomous (no block barrier anywhere).  A warp takes chunks of CHUNK = 128 consecutive records from a
// dynamic queue.  Per chunk:
//   detect  lane l owns records [k0 + 4l, k0 + 4l + 4): one 32-byte ij load + one 4-byte flg load (coalesced, 9 B/record).
//           HOT records (in contact at their last evaluation) are work by definition; the cold records gather the two position
//           sectors and run detection_vertex_vertex.  (prefetch.global.L2 of the hot records' operands at this point was
//           measured 5 % SLOWER than no prefetch: profiles/r1_ncu_session4.txt.)  A cold record that stays cold costs
//           9 B + two mostly L1/L2-resident gathers and nothing else (no history read, no write).
//   compact shuffle scan of the per-lane counts; the (k, i, j, flg) entries of the records that need the law go, in record order,
//           to a 2 KB per-warp shared-memory list.
//   law     dense: 32 entries per round, all eight operands of an entry issued back to back,
//           detection + Hooke law + history / wrench stores on fully populated warps.
// The id of the next chunk is fetched (one atomic per 128 records) before the current one is processed.
constexpr int CHUNK = 128;

template <int LAW>
__device__ __forceinline__ bool detect_need(const ContactArgs& A, const uint2 ij, unsigned long long pol_keep) {
#if EDB_L2_KEEP
  const double4 pi = ldg4h(A.P.pr + ij.x, pol_keep), pj = ldg4h((ij.y >= A.rem_lo ? A.g_pr : A.P.pr) + ij.y, pol_keep);
#else
  const double4 pi = ldg4(A.P.pr + ij.x), pj = ldg4((ij.y >= A.rem_lo ? A.g_pr : A.P.pr) + ij.y);
#endif
  const Contact c = detect_vv(mk3(pi.x, pi.y, pi.z), pi.w, mk3(pj.x, pj.y, pj.z), pj.w);
  bool need = c.hit;
  if (LAW == LAW_COHESIVE) {
    const double dncut = A.prm.tpp ? cp_pp(A.prm, group_of(ldg4(A.P.wg + ij.x)), group_of(ldg4(A.P.wg + ij.y))).dncut : A.prm.pp.dncut;
    need = need || c.dn <= dncut;
  }
  return need;
}

#ifndef EDB_PP_BLK
#define EDB_PP_BLK 160
#endif
#ifndef EDB_PP_MINB
#define EDB_PP_MINB 4
#endif
constexpr int PP_BLK = EDB_PP_BLK, PP_MINB = EDB_PP_MINB;   // tuning knobs of the sweep in profiles/ (registers vs resident warps)
template <int LAW>
__global__ void __launch_bounds__(PP_BLK, PP_MINB) k_contact_sphere_pp(ContactArgs A, uint32_t* __restrict__ queue) {
  if (gated(A.gate)) return;
  if (A.zero_flag && blockIdx.x == 0 && threadIdx.x == 0) *A.zero_flag = 0;   // trigger accumulator of the gather that follows
  __shared__ uint4 s_ent[PP_BLK / 32][CHUNK];
  const L2Pol pol = l2_policies();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint4* ents = s_ent[wid];
  const uint32_t nchunk = (A.w.n + CHUNK - 1) / CHUNK;
  const uint32_t nwarp = gridDim.x * (PP_BLK / 32);
  uint32_t chunk = blockIdx.x * (PP_BLK / 32) + wid;
  while (chunk < nchunk) {
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(queue, 1u);
    const uint32_t k0 = chunk * CHUNK + lane * 4;
    uint2 ij[4];
    uint32_t f4 = 0;
    if (k0 + 4 <= A.w.n) {
      const uint4 a = reinterpret_cast<const uint4*>(A.w.ij + k0)[0], b = reinterpret_cast<const uint4*>(A.w.ij + k0)[1];
      f4 = *reinterpret_cast<const uint32_t*>(A.w.flg + k0);
      ij[0] = make_uint2(a.x, a.y); ij[1] = make_uint2(a.z, a.w); ij[2] = make_uint2(b.x, b.y); ij[3] = make_uint2(b.z, b.w);
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        ij[u] = make_uint2(0, 0);
        if (k0 + u < A.w.n) {
          ij[u] = A.w.ij[k0 + u];
          f4 |= (uint32_t)A.w.flg[k0 + u] << (8 * u);
        }
      }
    }
    bool need[4];
#pragma unroll
    for (int u = 0; u < 4; u++) need[u] = (f4 >> (8 * u)) & FLG_HOT;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (!need[u] && k0 + u < A.w.n) need[u] = detect_need<LAW>(A, ij[u], EDB_L2_KEEP ? pol.keep : 0ull);
    uint32_t c = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) c += need[u] ? 1u : 0u;
    const uint32_t inc = warp_incl_scan(c);
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t pos = inc - c;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (need[u]) ents[pos++] = make_uint4(k0 + u, ij[u].x, ij[u].y, (f4 >> (8 * u)) & 0xffu);
    __syncwarp();
    for (uint32_t i = lane; i < total; i += 32) {
      const uint4 e = ents[i];
      const uint32_t k = e.x;
      const uint8_t fl = (uint8_t)e.w;
#if EDB_PP_PREFETCH
      if (i + 32 < total) {   // operands of this lane's entry of the NEXT round: on their way while this round's law is evaluated
        const uint4 en = ents[i + 32];
        const bool rem = en.z >= A.rem_lo;
        prefetch_l1(A.P.pr + en.y); prefetch_l1((rem ? A.g_pr : A.P.pr) + en.z); prefetch_l1(A.P.wg + en.y);
        prefetch_l1((rem ? A.g_wg : A.P.wg) + en.z); prefetch_l1(A.P.vm + en.y); prefetch_l1((rem ? A.g_vm : A.P.vm) + en.z);
        prefetch_l1(A.w.ft + en.x);
      }
#endif
      const bool remj = e.z >= A.rem_lo;   // j is a neighbour brick's particle: its sectors are in the landing zone
#if EDB_L2_KEEP
      const double4 pi = ldg4h(A.P.pr + e.y, pol.keep), pj = ldg4h((remj ? A.g_pr : A.P.pr) + e.z, pol.keep);
      const double4 wi = ldg4h(A.P.wg + e.y, pol.keep), wj = ldg4h((remj ? A.g_wg : A.P.wg) + e.z, pol.keep);
      const double4 vi = ldg4h(A.P.vm + e.y, pol.keep), vj = ldg4h((remj ? A.g_vm : A.P.vm) + e.z, pol.keep);
#else
      const double4 pi = ldg4(A.P.pr + e.y), pj = ldg4((remj ? A.g_pr : A.P.pr) + e.z);
      const double4 wi = ldg4(A.P.wg + e.y), wj = ldg4((remj ? A.g_wg : A.P.wg) + e.z);
      const double4 vi = ldg4(A.P.vm + e.y), vj = ldg4((remj ? A.g_vm : A.P.vm) + e.z);
#endif
      Hist H;
#if EDB_L2_STREAM
      H.ft = ld4h(A.w.ft + k, pol.stream);
#else
      H.ft = ld4(A.w.ft + k);
#endif
      H.mo = (fl & FLG_MOM) ? ld4(A.w.mo + k) : make_double4(0, 0, 0, 0);
      const v3 ri = mk3(pi.x, pi.y, pi.z), rj = mk3(pj.x, pj.y, pj.z);
      const Contact ct = detect_vv(ri, pi.w, rj, pj.w);
      const Params& cp = cp_pp(A.prm, group_of(wi), group_of(wj));
      bool contact = ct.hit;
      if (LAW == LAW_COHESIVE) contact = contact || ct.dn <= cp.dncut;
      const double meff = (vi.w * vj.w) / (vi.w + vj.w);   // common_kernels.hpp:93-96
      const double reff = (pi.w * pj.w) / (pi.w + pj.w);   // :99-102
#if EDB_L2_STREAM
      finish_record<LAW, true>(A, k, fl, contact, ct, cp, meff, reff, ri, mk3(vi.x, vi.y, vi.z), mk3(wi.x, wi.y, wi.z), rj,
                               mk3(vj.x, vj.y, vj.z), mk3(wj.x, wj.y, wj.z), H, pol.stream);
#else
      finish_record<LAW>(A, k, fl, contact, ct, cp, meff, reff, ri, mk3(vi.x, vi.y, vi.z), mk3(wi.x, wi.y, wi.z), rj,
                         mk3(vj.x, vj.y, vj.z), mk3(wj.x, wj.y, wj.z), H);
#endif
    }
    __syncwarp();
    chunk = nwarp + __shfl_sync(0xffffffffu, next, 0);
  }
}

// Software-pipelined form of k_contact_sphere_pp (EDB_PP_PIPE: 1 = operands prefetched into L2, 2 = into L1): a warp DETECTS chunk
// c+1 (compaction into the other entry buffer) and issues prefetches for the operands of its entries BEFORE it evaluates the law on
// the entries of chunk c, so that the seven scattered sectors of a law entry have a whole detect + law phase to arrive instead of
// being waited for at full DRAM latency by the first instruction that needs them.  Same arithmetic, same record order per warp.
#ifndef EDB_PP_PIPE
#define EDB_PP_PIPE 0
#endif
template <int LAW>
__global__ void __launch_bounds__(PP_BLK, PP_MINB) k_contact_sphere_pp_pipe(ContactArgs A, uint32_t* __restrict__ queue) {
  if (gated(A.gate)) return;
  if (A.zero_flag && blockIdx.x == 0 && threadIdx.x == 0) *A.zero_flag = 0;
  __shared__ uint4 s_ent[PP_BLK / 32][2][CHUNK];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t nchunk = (A.w.n + CHUNK - 1) / CHUNK;
  const uint32_t nwarp = gridDim.x * (PP_BLK / 32);
  // detect one chunk into `ents`; returns the number of entries (warp-uniform); prefetches their operands
  auto detect = [&](uint32_t chunk, uint4* ents) -> uint32_t {
    const uint32_t k0 = chunk * CHUNK + lane * 4;
    uint2 ij[4];
    uint32_t f4 = 0;
    if (k0 + 4 <= A.w.n) {
      const uint4 a = reinterpret_cast<const uint4*>(A.w.ij + k0)[0], b = reinterpret_cast<const uint4*>(A.w.ij + k0)[1];
      f4 = *reinterpret_cast<const uint32_t*>(A.w.flg + k0);
      ij[0] = make_uint2(a.x, a.y); ij[1] = make_uint2(a.z, a.w); ij[2] = make_uint2(b.x, b.y); ij[3] = make_uint2(b.z, b.w);
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        ij[u] = make_uint2(0, 0);
        if (k0 + u < A.w.n) {
          ij[u] = A.w.ij[k0 + u];
          f4 |= (uint32_t)A.w.flg[k0 + u] << (8 * u);
        }
      }
    }
    bool need[4];
#pragma unroll
    for (int u = 0; u < 4; u++) need[u] = (f4 >> (8 * u)) & FLG_HOT;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (!need[u] && k0 + u < A.w.n) need[u] = detect_need<LAW>(A, ij[u], 0ull);
    uint32_t c = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) c += need[u] ? 1u : 0u;
    const uint32_t inc = warp_incl_scan(c);
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t pos = inc - c;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (need[u]) {
        ents[pos++] = make_uint4(k0 + u, ij[u].x, ij[u].y, (f4 >> (8 * u)) & 0xffu);
        const bool rem = ij[u].y >= A.rem_lo;
        if (EDB_PP_PIPE == 2) {
          prefetch_l1(A.P.pr + ij[u].x); prefetch_l1((rem ? A.g_pr : A.P.pr) + ij[u].y); prefetch_l1(A.P.wg + ij[u].x);
          prefetch_l1((rem ? A.g_wg : A.P.wg) + ij[u].y); prefetch_l1(A.P.vm + ij[u].x); prefetch_l1((rem ? A.g_vm : A.P.vm) + ij[u].y);
          prefetch_l1(A.w.ft + k0 + u);
        } else {
          prefetch_l2(A.P.pr + ij[u].x); prefetch_l2((rem ? A.g_pr : A.P.pr) + ij[u].y); prefetch_l2(A.P.wg + ij[u].x);
          prefetch_l2((rem ? A.g_wg : A.P.wg) + ij[u].y); prefetch_l2(A.P.vm + ij[u].x); prefetch_l2((rem ? A.g_vm : A.P.vm) + ij[u].y);
          prefetch_l2(A.w.ft + k0 + u);
        }
      }
    __syncwarp();
    return total;
  };
  auto law = [&](const uint4* ents, uint32_t total) {
    for (uint32_t i = lane; i < total; i += 32) {
      const uint4 e = ents[i];
      const uint32_t k = e.x;
      const uint8_t fl = (uint8_t)e.w;
      const bool remj = e.z >= A.rem_lo;
      const double4 pi = ldg4(A.P.pr + e.y), pj = ldg4((remj ? A.g_pr : A.P.pr) + e.z);
      const double4 wi = ldg4(A.P.wg + e.y), wj = ldg4((remj ? A.g_wg : A.P.wg) + e.z);
      const double4 vi = ldg4(A.P.vm + e.y), vj = ldg4((remj ? A.g_vm : A.P.vm) + e.z);
      Hist H;
      H.ft = ld4(A.w.ft + k);
      H.mo = (fl & FLG_MOM) ? ld4(A.w.mo + k) : make_double4(0, 0, 0, 0);
      const v3 ri = mk3(pi.x, pi.y, pi.z), rj = mk3(pj.x, pj.y, pj.z);
      const Contact ct = detect_vv(ri, pi.w, rj, pj.w);
      const Params& cp = cp_pp(A.prm, group_of(wi), group_of(wj));
      bool contact = ct.hit;
      if (LAW == LAW_COHESIVE) contact = contact || ct.dn <= cp.dncut;
      const double meff = (vi.w * vj.w) / (vi.w + vj.w);
      const double reff = (pi.w * pj.w) / (pi.w + pj.w);
      finish_record<LAW>(A, k, fl, contact, ct, cp, meff, reff, ri, mk3(vi.x, vi.y, vi.z), mk3(wi.x, wi.y, wi.z), rj,
                         mk3(vj.x, vj.y, vj.z), mk3(wj.x, wj.y, wj.z), H);
    }
    __syncwarp();
  };
  uint32_t cur = blockIdx.x * (PP_BLK / 32) + wid;
  if (cur >= nchunk) return;
  int buf = 0;
  uint32_t tot_cur = detect(cur, s_ent[wid][0]);
  for (;;) {
    uint32_t nxt = 0;
    if (lane == 0) nxt = atomicAdd(queue, 1u);
    nxt = nwarp + __shfl_sync(0xffffffffu, nxt, 0);
    uint32_t tot_nxt = 0;
    const bool have_next = nxt < nchunk;
    if (have_next) tot_nxt = detect(nxt, s_ent[wid][buf ^ 1]);
    law(s_ent[wid][buf], tot_cur);
    if (!have_next) break;
    buf ^= 1;
    tot_cur = tot_nxt;
  }
}

// detection of a driver record (waves 4,5,6): sphere::ContactLawDriverFunc (contact_sphere.hpp:232-296),
// polyhedron::ContactLawDriverFunc (polyhedron/contact.hpp:237-303)
template <bool POLY>
__device__ __forceinline__ Contact detect_driver_record(const ContactArgs& A, uint32_t k, const uint2 ij, const Driver& D, const double4& pi,
                                                        const double4& wi, double& reff) {
  if (!POLY) {
    reff = pi.w;
    return driver_detect(D, pi.w, mk3(pi.x, pi.y, pi.z));
  }
  const ShapeDev sh = A.S.sh[type_of(wi)];
  const double rm = A.P.h[ij.x] * sh.radius;
  const double4 vx = A.vert[(size_t)ij.x * A.S.nvmax + A.w.sub[k].x];
  reff = rm;
  return driver_detect(D, rm, mk3(vx.x, vx.y, vx.z));
}
template <int LAW, bool POLY>
__global__ void __launch_bounds__(BLK) k_contact_driver(ContactArgs A) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.w.n || gated(A.gate)) return;
  const uint2 ij = A.w.ij[k];
  const uint8_t fl = A.w.flg[k];
  const bool hot = fl & FLG_HOT;
  const Driver D = A.drv[ij.y];
  const double4 pi = ldg4(A.P.pr + ij.x), wi = ldg4(A.P.wg + ij.x);
  Hist H;
  double4 vi = make_double4(0, 0, 0, 0);
  if (hot) { H.ft = ld4(A.w.ft + k); vi = ldg4(A.P.vm + ij.x); }
  H.mo = (fl & FLG_MOM) ? ld4(A.w.mo + k) : make_double4(0, 0, 0, 0);
  double reff;
  const Contact c = detect_driver_record<POLY>(A, k, ij, D, pi, wi, reff);
  const Params& cp = cp_pd(A.prm, group_of(wi), ij.y);
  bool contact = c.hit;
  if (LAW == LAW_COHESIVE) contact = contact || c.dn <= cp.dncut;
  if (!contact && !hot) return;
  if (contact && !hot) { H.ft = ld4(A.w.ft + k); vi = ldg4(A.P.vm + ij.x); }
  finish_record<LAW>(A, k, fl, contact, c, cp, vi.w, reff, mk3(pi.x, pi.y, pi.z), mk3(vi.x, vi.y, vi.z), mk3(wi.x, wi.y, wi.z), D.center,
                     D.vel, D.vrot, H);
}

// detection of a polyhedron particle-particle record of wave `type` (polyhedron/contact.hpp:124-211 -> detect<type>)
__device__ __forceinline__ Contact detect_poly_record(const ContactArgs& A, int type, const uint2 ij, const uint2 sb, const ShapeDev& si,
                                                      const ShapeDev& sj, double rmi, double rmj) {
  const VA vi_{A.vert + (size_t)ij.x * A.S.nvmax}, vj_{A.vert + (size_t)ij.y * A.S.nvmax};
  if (type == 0) return detect_vv(vi_[sb.x], rmi, vj_[sb.y], rmj);
  if (type == 1) {
    const int* e = A.S.edges + 2 * (sj.e0 + sb.y);
    return detect_ve(vi_[sb.x], rmi, vj_[e[0]], vj_[e[1]], rmj);
  }
  if (type == 2) return detect_vf(vi_[sb.x], rmi, vj_, rmj, A.S.fidx + sj.fi0 + A.S.foff[sj.f0 + sb.y], A.S.fn[sj.f0 + sb.y]);
  const int* ei = A.S.edges + 2 * (si.e0 + sb.x);
  const int* ej = A.S.edges + 2 * (sj.e0 + sb.y);
  return detect_ee(vi_[ei[0]], vi_[ei[1]], rmi, vj_[ej[0]], vj_[ej[1]], rmj);
}
// waves 0..3 of the polyhedron operator: polyhedron::ContactLawFunc<type> (polyhedron/contact.hpp:124-211).
// Two launches cover the four waves together (their records form one concatenated index space):
//   k_detect_poly_pp   light (<= 64 registers, 4 CTAs/SM): detection only, for every record — in a packing ~95 % of the records
//                      are candidates that do not touch, so this pass IS the cost of the operator.  Records that need the law
//                      (contact, cohesive range, or HOT: history to reset) are appended, in record order per block, to a list.
//   k_law_poly_pp      persistent, one list entry per thread and iteration: the full record evaluation on dense warps.
struct PolyWaves {
  WView w[4];
  uint32_t start[5];   // concatenated index space: wave t owns [start[t], start[t+1])
};
struct NeedList {
  uint32_t* ent;       // (wave << 28) | record index
  uint32_t* count;
};
template <int LAW>
__device__ __forceinline__ void poly_pp_record(const ContactArgs& A, uint32_t k) {
  const uint2 ij = A.w.ij[k], sb = A.w.sub[k];
  const uint8_t fl = A.w.flg[k];
  const bool hot = fl & FLG_HOT;
  const double4 wi = ldg4(A.P.wg + ij.x), wj = ldg4(A.P.wg + ij.y);
  Hist H;
  double4 vi = make_double4(0, 0, 0, 0), vj = vi;
  if (hot) { H.ft = ld4(A.w.ft + k); vi = ldg4(A.P.vm + ij.x); vj = ldg4(A.P.vm + ij.y); }
  H.mo = (fl & FLG_MOM) ? ld4(A.w.mo + k) : make_double4(0, 0, 0, 0);
  const ShapeDev si = A.S.sh[type_of(wi)], sj = A.S.sh[type_of(wj)];
  const double rmi = A.P.h[ij.x] * si.radius, rmj = A.P.h[ij.y] * sj.radius;
  const Contact c = detect_poly_record(A, A.type, ij, sb, si, sj, rmi, rmj);
  const Params& cp = cp_pp(A.prm, group_of(wi), group_of(wj));
  bool contact = c.hit;
  if (LAW == LAW_COHESIVE) contact = contact || c.dn <= cp.dncut;
  if (!contact && !hot) return;
  if (contact && !hot) { H.ft = ld4(A.w.ft + k); vi = ldg4(A.P.vm + ij.x); vj = ldg4(A.P.vm + ij.y); }
  const double4 pi = ldg4(A.P.pr + ij.x), pj = ldg4(A.P.pr + ij.y);
  double meff = 0, reff = 0;
  if (contact) {
    meff = (vi.w * vj.w) / (vi.w + vj.w);
    reff = (rmi * rmj) / (rmi + rmj);  // contact.hpp:183 (effective-mass formula on the Minkowski radii)
  }
  finish_record<LAW>(A, k, fl, contact, c, cp, meff, reff, mk3(pi.x, pi.y, pi.z), mk3(vi.x, vi.y, vi.z), mk3(wi.x, wi.y, wi.z),
                     mk3(pj.x, pj.y, pj.z), mk3(vj.x, vj.y, vj.z), mk3(wj.x, wj.y, wj.z), H);
}
constexpr int PDPT = 2;   // records per thread of the detection pass
template <int LAW>
__global__ void __launch_bounds__(BLK, 4) k_detect_poly_pp(ContactArgs A, PolyWaves PW, NeedList L) {
  if (gated(A.gate)) return;
  const uint32_t g0 = (blockIdx.x * BLK + threadIdx.x) * PDPT;
  uint32_t ent[PDPT];
  bool need[PDPT];
#pragma unroll
  for (int u = 0; u < PDPT; u++) {
    need[u] = false;
    const uint32_t g = g0 + u;
    if (g >= PW.start[4]) continue;
    const int t = g >= PW.start[2] ? (g >= PW.start[3] ? 3 : 2) : (g >= PW.start[1] ? 1 : 0);
    const WView& w = PW.w[t];
    const uint32_t k = g - PW.start[t];
    ent[u] = ((uint32_t)t << 28) | k;
    const uint8_t fl = w.flg[k];
    if (fl & FLG_HOT) { need[u] = true; continue; }
    const uint2 ij = w.ij[k], sb = w.sub[k];
    // shape index from the compact table, Minkowski radius from the 4th component of the first world vertex: no particle sector
    const ShapeDev& si = A.S.sh[A.ptype[ij.x]];
    const ShapeDev& sj = A.S.sh[A.ptype[ij.y]];
    const double rmi = ldg4(A.vert + (size_t)ij.x * A.S.nvmax).w, rmj = ldg4(A.vert + (size_t)ij.y * A.S.nvmax).w;
    const Contact c = detect_poly_record(A, t, ij, sb, si, sj, rmi, rmj);
    need[u] = c.hit;
    if (LAW == LAW_COHESIVE) {
      double dncut = A.prm.pp.dncut;
      if (A.prm.tpp) dncut = cp_pp(A.prm, group_of(ldg4(A.P.wg + ij.x)), group_of(ldg4(A.P.wg + ij.y))).dncut;
      need[u] = need[u] || c.dn <= dncut;
    }
  }
  uint32_t c = 0;
#pragma unroll
  for (int u = 0; u < PDPT; u++) c += need[u] ? 1u : 0u;
  __shared__ uint32_t s_base;
  uint32_t tot;
  uint32_t pos = block_excl_scan(c, &tot);
  if (threadIdx.x == 0) s_base = tot ? atomicAdd(L.count, tot) : 0u;
  __syncthreads();
  pos += s_base;
#pragma unroll
  for (int u = 0; u < PDPT; u++)
    if (need[u]) L.ent[pos++] = ent[u];
}
template <int LAW>
__global__ void __launch_bounds__(BLK, 2) k_law_poly_pp(ContactArgs A, PolyWaves PW, NeedList L) {
  if (gated(A.gate)) return;
  const uint32_t cnt = *L.count;
  for (uint32_t x = blockIdx.x * BLK + threadIdx.x; x < cnt; x += gridDim.x * BLK) {
    const uint32_t e = L.ent[x];
    A.type = (int)(e >> 28);
    A.w = PW.w[e >> 28];
    poly_pp_record<LAW>(A, e & 0x0fffffffu);
  }
}

// detection of an RShape (mesh) driver record, waves 7..12: detect<7..12> (shape_detection_api.hpp:356-406) on the particle
// (sphere centre / polyhedron world vertices) and the cached world vertices of the mesh (quirk Q13: both reference paths agree)
template <bool POLY>
__device__ __forceinline__ Contact detect_mesh_record(const ContactArgs& A, int t, uint32_t k, const uint2 ij, const Driver& D, const double4& pi,
                                                      const double4& wi, double& reff) {
  const MeshDev md = A.M.md[D.shape];
  const VA mv{A.M.verts + md.v0};
  const int* me = A.M.edges + 2 * md.e0;
  const uint2 sb = A.w.sub[k];
  const double rd = md.radius;
  if (!POLY) {
    const v3 ri = mk3(pi.x, pi.y, pi.z);
    reff = pi.w;
    if (t == 7) return detect_vv(ri, pi.w, mv[sb.y], rd);
    if (t == 8) return detect_ve(ri, pi.w, mv[me[2 * sb.y]], mv[me[2 * sb.y + 1]], rd);
    return detect_vf(ri, pi.w, mv, rd, A.M.fidx + md.fi0 + A.M.foff[md.f0 + sb.y], A.M.fn[md.f0 + sb.y]);
  }
  const ShapeDev sh = A.S.sh[type_of(wi)];
  const double ra = A.P.h[ij.x] * sh.radius;
  const VA vi{A.vert + (size_t)ij.x * A.S.nvmax};
  const int* ei = A.S.edges + 2 * sh.e0;
  reff = (ra * rd) / (ra + rd);  // contact.hpp:388
  if (t == 7) return detect_vv(vi[sb.x], ra, mv[sb.y], rd);
  if (t == 8) return detect_ve(vi[sb.x], ra, mv[me[2 * sb.y]], mv[me[2 * sb.y + 1]], rd);
  if (t == 9) return detect_vf(vi[sb.x], ra, mv, rd, A.M.fidx + md.fi0 + A.M.foff[md.f0 + sb.y], A.M.fn[md.f0 + sb.y]);
  if (t == 10) return detect_ee(vi[ei[2 * sb.x]], vi[ei[2 * sb.x + 1]], ra, mv[me[2 * sb.y]], mv[me[2 * sb.y + 1]], rd);
  if (t == 11) return detect_ve(mv[sb.y], rd, vi[ei[2 * sb.x]], vi[ei[2 * sb.x + 1]], ra);
  return detect_vf(mv[sb.y], rd, vi, ra, A.S.fidx + sh.fi0 + A.S.foff[sh.f0 + sb.x], A.S.fn[sh.f0 + sb.x]);
}
// waves 7..12: sphere::ContactLawRShapeDriverFunc (contact_sphere.hpp:306-396), polyhedron::ContactLawRShapeDriverFunc
// (polyhedron/contact.hpp:309-432).  Types 11/12 are "j to i": the driver plays i in the law and the stored friction flips sign.
template <int LAW, bool POLY>
__global__ void __launch_bounds__(BLK) k_contact_mesh(ContactArgs A) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.w.n || gated(A.gate)) return;
  const uint2 ij = A.w.ij[k];
  const uint8_t fl = A.w.flg[k];
  const bool hot = fl & FLG_HOT;
  const Driver D = A.drv[ij.y];
  const double4 pi = ldg4(A.P.pr + ij.x), wi = ldg4(A.P.wg + ij.x);
  double reff;
  const Contact c = detect_mesh_record<POLY>(A, A.type, k, ij, D, pi, wi, reff);
  const Params& cp = cp_pd(A.prm, group_of(wi), ij.y);
  bool contact = c.hit;
  if (LAW == LAW_COHESIVE) contact = contact || c.dn <= cp.dncut;
  if (!contact && !hot) return;
  Hist H;
  H.ft = ld4(A.w.ft + k);
  H.mo = (fl & FLG_MOM) ? ld4(A.w.mo + k) : make_double4(0, 0, 0, 0);
  const double4 vi = ldg4(A.P.vm + ij.x);
  const v3 ri = mk3(pi.x, pi.y, pi.z), v_i = mk3(vi.x, vi.y, vi.z), w_i = mk3(wi.x, wi.y, wi.z);
  if (A.type <= 10 || !contact) {
    finish_record<LAW>(A, k, fl, contact, c, cp, vi.w, reff, ri, v_i, w_i, D.center, D.vel, D.vrot, H);
  } else {
    v3 ft = mk3(H.ft.x, H.ft.y, H.ft.z), mom = mk3(H.mo.x, H.mo.y, H.mo.z), f = mk3(0, 0, 0);
    contact_force<LAW>(c.dn, c.n, A.dt, cp, vi.w, reff, ft, c.cp, D.center, D.vel, f, mom, D.vrot, ri, v_i, w_i);
    // stored: fd = f of the law (the particle receives -f, see the gather), friction negated (contact.hpp:417)
    st4(A.w.ft + k, make_double4(-ft.x, -ft.y, -ft.z, H.ft.w));
    st4(A.w.fd(k), make_double4(f.x, f.y, f.z, c.dn));
    st4(A.w.cp(k), make_double4(c.cp.x, c.cp.y, c.cp.z, 0));
    const bool live = !is_zero(mom);
    if (live || (fl & FLG_MOM)) st4(A.w.mo + k, make_double4(mom.x, mom.y, mom.z, 0.0));
    const uint8_t nf = (uint8_t)((fl | FLG_HOT) & ~FLG_MOM) | (live ? FLG_MOM : 0);
    if (nf != fl) A.w.flg[k] = nf;
  }
}

// ContactState view of a wave (classifier/contact_state.hpp:22-51): dn, cp, fn = f - ft (contact_sphere.hpp:180), ft.
// Records that are not hot hold no stored state: dn = fn = ft = 0 and cp is what the detection returns for the current
// positions (what the reference stores too, since it runs the detection for every record).
template <bool POLY>
__global__ void k_state(ContactArgs A, double* __restrict__ dn, double* __restrict__ cp, double* __restrict__ fn, double* __restrict__ ft) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.w.n) return;
  const uint8_t fl = A.w.flg[k];
  if (fl & FLG_HOT) {
    const double4 fd = ld4(A.w.fd(k)), c4 = ld4(A.w.cp(k)), f4 = ld4(A.w.ft + k);
    dn[k] = fd.w;
    cp[3 * k] = c4.x; cp[3 * k + 1] = c4.y; cp[3 * k + 2] = c4.z;
    if (A.type == 11 || A.type == 12) {  // fn = friction - f with the friction BEFORE its sign flip (contact.hpp:409,417)
      fn[3 * k] = -f4.x - fd.x; fn[3 * k + 1] = -f4.y - fd.y; fn[3 * k + 2] = -f4.z - fd.z;
    } else {
      fn[3 * k] = fd.x - f4.x; fn[3 * k + 1] = fd.y - f4.y; fn[3 * k + 2] = fd.z - f4.z;
    }
    ft[3 * k] = f4.x; ft[3 * k + 1] = f4.y; ft[3 * k + 2] = f4.z;
    return;
  }
  const uint2 ij = A.w.ij[k];
  Contact c;
  if (A.type <= 3) {
    if (!POLY) {
      const double4 pi = ldg4(A.P.pr + ij.x), pj = ldg4(A.P.pr + ij.y);
      c = detect_vv(mk3(pi.x, pi.y, pi.z), pi.w, mk3(pj.x, pj.y, pj.z), pj.w);
    } else {
      const double4 wi = ldg4(A.P.wg + ij.x), wj = ldg4(A.P.wg + ij.y);
      const ShapeDev si = A.S.sh[type_of(wi)], sj = A.S.sh[type_of(wj)];
      c = detect_poly_record(A, A.type, ij, A.w.sub[k], si, sj, A.P.h[ij.x] * si.radius, A.P.h[ij.y] * sj.radius);
    }
  } else if (A.type <= 6) {
    double reff;
    c = detect_driver_record<POLY>(A, k, ij, A.drv[ij.y], ldg4(A.P.pr + ij.x), ldg4(A.P.wg + ij.x), reff);
  } else {
    double reff;
    c = detect_mesh_record<POLY>(A, A.type, k, ij, A.drv[ij.y], ldg4(A.P.pr + ij.x), ldg4(A.P.wg + ij.x), reff);
  }
  dn[k] = 0.0;
  cp[3 * k] = c.cp.x; cp[3 * k + 1] = c.cp.y; cp[3 * k + 2] = c.cp.z;
  fn[3 * k] = 0.0; fn[3 * k + 1] = 0.0; fn[3 * k + 2] = 0.0;
  ft[3 * k] = 0.0; ft[3 * k + 1] = 0.0; ft[3 * k + 2] = 0.0;
}

// ---------------------------------------------------------------------------------------------------------------
// deterministic force / torque gather
// ---------------------------------------------------------------------------------------------------------------
struct GatherArgs {
  PView P;
  WTab W;
  int order[NT];     // wave order of the operator (sphere: 0,4,5,6,7,8,9 ; polyhedron: 0..3,7..12,4,5,6)
  int norder;
  const uint32_t* inc_off;   // per real particle: range of its incidence list (kernels_rebuild.cuh: k_inc_owner / k_incidence)
  const uint32_t* inc;       // codes (position in `order` << 28 | record index), canonical (serial-reference) order
  uint32_t nreal;
  int npp;           // number of particle-particle waves (1 spheres, 4 polyhedra): their owner ranges decide owner / partner
  int has_swap;      // polyhedra: role of the owner depends on the swap bit
  double g[3];
  double hdt;
  // FUSE 2 only: the prolog of the NEXT step and its rebuild trigger
  double dt, dt2;
  const double4* pr_backup;            // positions at the last neighbour build (nullptr: no trigger evaluation)
  const double4* q_backup;             // polyhedra: orientations at the last build (vertex_displ_over, backup_dem.hpp:66-100)
  ShapesView S;                        // polyhedra: FUSE 2 also runs compute_vertices for the real particles ...
  double4* vert;                       // ... into the world-vertex array (nullptr for spheres)
  double thr2;
  unsigned long long* next_flag;       // += number of particles with |r - r_backup|^2 >= thr2 after the next prolog
  const unsigned long long* gate;      // see ContactArgs::gate
  uint32_t* reset_queue;               // FUSE 2: zeroed for the next step's k_contact_sphere_pp (saves a memset per step)
  // FUSE 1 / 2: start the sums from the force / moment ARRAYS instead of from gravity — they hold reset_force_moment + gravity_force +
  // inner_bond_polyhedron of this step (k_reset_gravity, k_bond_gather), in the operator order of the fragmentation macro
  int base_from_arrays;
};

// role_i: particle acts as "i" of the record: F += f, M += (cp - r) x f + mom;  as "j": F += -f, M += (cp - r) x (-f) + (-mom)
__device__ __forceinline__ void add_wrench(const double4& fd, const double4& cp, const double4& mo, bool role_i, v3 r, v3& F, v3& M) {
  v3 f = mk3(fd.x, fd.y, fd.z), m = mk3(mo.x, mo.y, mo.z);
  if (!role_i) {
    f = -f;
    m = -m;
  }
  const v3 t = moments(mk3(cp.x, cp.y, cp.z), r, f, m);
  M.x += t.x; M.y += t.y; M.z += t.z;   // lockAndAdd(mom, ...) first, then f (contact_sphere.hpp:183-189)
  F.x += f.x; F.y += f.y; F.z += f.z;
}

// One thread per real particle walks its incidence list — every record in which the particle is owner or partner, in the order
// the serial reference loop would have applied them (wave by wave, record index ascending) — GU entries at a time: GU codes,
// then GU flag bytes, then the wrenches of the hot ones are all in flight together (three dependent round trips per GU records
// instead of three per record), and the sum is taken in list order, so the result is bit-reproducible.
// FUSE 0: [contact_*] only: f, mom += sum        (drop-in operator semantics: accumulate on top of gravity etc.)
// FUSE 1: reset_force_moment + gravity_force + contact + force_to_accel + combined_compute_epilog in one pass
// FUSE 2: FUSE 1 + combined_compute_prolog of the NEXT step + particle_displ_over of the next step, all on the registers that
//         already hold the particle: one pass over the particle arrays per time step instead of three.  The intermediate
//         acceleration / moment / angular acceleration are not stored (nothing reads them before the next gather overwrites
//         them; the last step of edb_run_verlet runs FUSE 1, so the state a caller can observe is complete).
#ifndef EDB_G_BLK
#define EDB_G_BLK 128      // profiles/r2_sweep_gather_block.log: 128 x 6 CTAs 0.1105 ms, 192 x 4 0.1120, 256 x 3 0.1127, 384 x 2 0.1145
#endif
constexpr int G_BLK = EDB_G_BLK;   // threads per block of k_gather
#ifndef EDB_G_MINB
#define EDB_G_MINB 6
#endif
constexpr int GU = EDB_GU;
template <int FUSE, bool POLY>
__global__ void __launch_bounds__(G_BLK, EDB_G_MINB) k_gather(const __grid_constant__ GatherArgs A, const __grid_constant__ FanArgs FAN) {
#if EDB_G_REVERSE
  const uint32_t p = (gridDim.x - 1 - blockIdx.x) * blockDim.x + threadIdx.x;
#else
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
#endif
  if (gated(A.gate)) return;
  if (p >= A.nreal) return;   // (exited lanes are not waited for by the full-mask reduction of the FUSE 2 trigger)
  const L2Pol pol = l2_policies();
  constexpr bool KEEP = EDB_L2_KEEP && FUSE == 2 && !POLY, STREAM = EDB_L2_STREAM && FUSE == 2 && !POLY;
  const uint32_t e0 = A.inc_off[p], e1 = A.inc_off[p + 1];
#if EDB_G_PREFETCH
  // the particle's own sectors that are only needed after the incidence walk (and the backup position of the trigger): start their
  // trip from DRAM now — a prefetch holds no register, a hoisted load would cost 22 and a resident CTA
  if (FUSE) {
    prefetch_l1(A.P.wg + p); prefetch_l1(A.P.q + p);
    prefetch_l1(A.P.ix + p); prefetch_l1(A.P.iy + p); prefetch_l1(A.P.iz + p);
    if (FUSE == 2 && A.pr_backup && A.next_flag) prefetch_l1(A.pr_backup + p);
  }
#endif
  constexpr int NPP = POLY ? 4 : 1;   // particle-particle waves of the operator (spheres: wave 0 only)
  uint32_t oa[NPP], ob[NPP];          // owner ranges of the pp waves (order positions 0..npp-1 = waves 0..npp-1)
#pragma unroll
  for (int t = 0; t < NPP; t++) {
    oa[t] = 0; ob[t] = 0;
    if (t < A.npp && A.W.w[t].n) { oa[t] = A.W.w[t].off[p]; ob[t] = A.W.w[t].off[p + 1]; }
  }
  // FUSE 2 stores pr (own slot and ghost images) in this kernel: the non-coherent path is only legal for data that is read-only
  // for the whole kernel
  const double4 pr = KEEP ? ld4h(A.P.pr + p, pol.keep) : FUSE == 2 ? ld4(A.P.pr + p) : ldg4(A.P.pr + p);
  const v3 r = mk3(pr.x, pr.y, pr.z);
  v3 F, M;
  double4 vm;
  if (FUSE) {
    vm = KEEP ? ld4h(A.P.vm + p, pol.keep) : ld4(A.P.vm + p);
    if (A.base_from_arrays) {
      F = mk3(A.P.fx[p], A.P.fy[p], A.P.fz[p]);
      M = mk3(A.P.mx[p], A.P.my[p], A.P.mz[p]);
    } else {
      F = mk3(0.0 + A.g[0] * vm.w, 0.0 + A.g[1] * vm.w, 0.0 + A.g[2] * vm.w);
      M = mk3(0, 0, 0);
    }
  } else {
    F = mk3(A.P.fx[p], A.P.fy[p], A.P.fz[p]);
    M = mk3(A.P.mx[p], A.P.my[p], A.P.mz[p]);
  }
  if (FUSE == 2 && p == 0 && A.reset_queue) *A.reset_queue = 0;   // chunk queue of the next step's contact kernel
  // sum of one hot record's wrench in the role this particle plays in it
  auto accumulate = [&](uint32_t code, uint8_t flg, const double4& fd, const double4& cp) {
    const uint32_t o = code >> INC_SHIFT, k = code & INC_MASK;
    const double4 mo = (flg & FLG_MOM) ? ld4(A.W.w[A.order[o]].mo + k) : make_double4(0, 0, 0, 0);
    bool role_i;
    if ((int)o < A.npp) {
      uint32_t a = oa[0], b = ob[0];
      if (POLY) {
        a = o == 0 ? oa[0] : o == 1 ? oa[NPP > 1 ? 1 : 0] : o == 2 ? oa[NPP > 2 ? 2 : 0] : oa[NPP > 3 ? 3 : 0];
        b = o == 0 ? ob[0] : o == 1 ? ob[NPP > 1 ? 1 : 0] : o == 2 ? ob[NPP > 2 ? 2 : 0] : ob[NPP > 3 ? 3 : 0];
      }
      const bool owner = k >= a && k < b;
      const bool owner_is_i = (POLY && A.has_swap) ? !(flg & FLG_SWAP) : true;
      role_i = owner ? owner_is_i : !owner_is_i;
    } else {
      // mesh waves 11/12 ("j to i", contact.hpp:405-421): the particle receives -f, -mom
      const int t = A.order[o];
      role_i = !(t == 11 || t == 12);
    }
    add_wrench(fd, cp, mo, role_i, r, F, M);
  };
  if (POLY) {
    // polyhedra: a particle has a dozen-plus incident records of which few are in contact — eight codes, then eight flag bytes in
    // flight at a time; the rare hot record fetches its wrench on demand
    constexpr int PU = 8;
    for (uint32_t e = e0; e < e1; e += PU) {
      uint32_t code[PU];
      uint8_t fl[PU];
#pragma unroll
      for (int u = 0; u < PU; u++) code[u] = (e + u < e1) ? A.inc[e + u] : 0xffffffffu;
#pragma unroll
      for (int u = 0; u < PU; u++) {
        fl[u] = 0;
        if (code[u] != 0xffffffffu) fl[u] = A.W.w[A.order[code[u] >> INC_SHIFT]].flg[code[u] & INC_MASK];
      }
#pragma unroll
      for (int u = 0; u < PU; u++)
        if (fl[u] & FLG_HOT) {
          const WView& w = A.W.w[A.order[code[u] >> INC_SHIFT]];
          const double4 fd = ld4(w.fd(code[u] & INC_MASK)), cp = ld4(w.cp(code[u] & INC_MASK));
          accumulate(code[u], fl[u], fd, cp);
        }
    }
  } else {
#if EDB_G_WIDE
    // spheres: all the codes of the particle (8 at a time; a particle of a packing has ~6 incident records), then all their flag bytes,
    // then the wrenches of the hot ones pair by pair — three dependent round trips per particle instead of three per pair of records
    constexpr int CU = EDB_G_CU;
    for (uint32_t e = e0; e < e1; e += CU) {
      uint32_t code[CU];
      uint8_t fl[CU];
#pragma unroll
      for (int u = 0; u < CU; u++) code[u] = (e + u < e1) ? (STREAM ? ldu32h(A.inc + e + u, pol.stream) : A.inc[e + u]) : 0xffffffffu;
#pragma unroll
      for (int u = 0; u < CU; u++) {
        fl[u] = 0;
        if (code[u] != 0xffffffffu) fl[u] = A.W.w[A.order[code[u] >> INC_SHIFT]].flg[code[u] & INC_MASK];
      }
#pragma unroll
      for (int u = 0; u < CU; u += GU) {
        double4 fd[GU], cp[GU];
#pragma unroll
        for (int v = 0; v < GU; v++)
          if (fl[u + v] & FLG_HOT) {
            const WView& w = A.W.w[A.order[code[u + v] >> INC_SHIFT]];
            fd[v] = ld4(w.fd(code[u + v] & INC_MASK));
            cp[v] = ld4(w.cp(code[u + v] & INC_MASK));
          }
#pragma unroll
        for (int v = 0; v < GU; v++)
          if (fl[u + v] & FLG_HOT) accumulate(code[u + v], fl[u + v], fd[v], cp[v]);
      }
    }
#else
    for (uint32_t e = e0; e < e1; e += GU) {
      uint32_t code[GU];
      uint8_t fl[GU];
      double4 fd[GU], cp[GU];
#pragma unroll
      for (int u = 0; u < GU; u++) code[u] = (e + u < e1) ? (STREAM ? ldu32h(A.inc + e + u, pol.stream) : A.inc[e + u]) : 0xffffffffu;
#pragma unroll
      for (int u = 0; u < GU; u++) {
        fl[u] = 0;
        if (code[u] != 0xffffffffu) fl[u] = A.W.w[A.order[code[u] >> INC_SHIFT]].flg[code[u] & INC_MASK];
      }
#pragma unroll
      for (int u = 0; u < GU; u++)
        if (fl[u] & FLG_HOT) {
          const WView& w = A.W.w[A.order[code[u] >> INC_SHIFT]];
          fd[u] = ld4(w.fd(code[u] & INC_MASK));
          cp[u] = ld4(w.cp(code[u] & INC_MASK));
        }
#pragma unroll
      for (int u = 0; u < GU; u++)
        if (fl[u] & FLG_HOT) accumulate(code[u], fl[u], fd[u], cp[u]);
    }
#endif
  }
  if (!FUSE) {
    A.P.fx[p] = F.x; A.P.fy[p] = F.y; A.P.fz[p] = F.z;
    A.P.mx[p] = M.x; A.P.my[p] = M.y; A.P.mz[p] = M.z;
    return;
  }
  // force_to_accel, then combined_compute_epilog
  const v3 acc = mk3(F.x / vm.w, F.y / vm.w, F.z / vm.w);
  double4 w = KEEP ? ld4h(A.P.wg + p, pol.keep) : ld4(A.P.wg + p);
  double4 Q = STREAM ? ld4h(A.P.q + p, pol.stream) : ld4(A.P.q + p);
  const v3 I = STREAM ? mk3(ld1h(A.P.ix + p, pol.stream), ld1h(A.P.iy + p, pol.stream), ld1h(A.P.iz + p, pol.stream))
                      : mk3(A.P.ix[p], A.P.iy[p], A.P.iz[p]);
  const v3 ar = ang_accel(Q, M, mk3(w.x, w.y, w.z), I);
  const v3 wn = mk3(w.x, w.y, w.z) + ar * A.hdt;
  w.x = wn.x; w.y = wn.y; w.z = wn.z;
  vm.x += acc.x * A.hdt; vm.y += acc.y * A.hdt; vm.z += acc.z * A.hdt;
  if (FUSE == 1) {
    A.P.fx[p] = acc.x; A.P.fy[p] = acc.y; A.P.fz[p] = acc.z;
    A.P.mx[p] = M.x; A.P.my[p] = M.y; A.P.mz[p] = M.z;
    A.P.ax[p] = ar.x; A.P.ay[p] = ar.y; A.P.az[p] = ar.z;
    st4(A.P.wg + p, w); st4(A.P.vm + p, vm);
    return;
  }
  // combined_compute_prolog of the next step (same expressions as k_prolog, on the values k_prolog would have re-read)
  double4 rn = pr;
  rn.x += vm.x * A.dt + acc.x * A.dt2; rn.y += vm.y * A.dt + acc.y * A.dt2; rn.z += vm.z * A.dt + acc.z * A.dt2;
  vm.x += acc.x * A.hdt; vm.y += acc.y * A.hdt; vm.z += acc.z * A.hdt;
  const q4 q{Q.x, Q.y, Q.z, Q.w};
  const q4 dq = qdot(q, mk3(w.x, w.y, w.z));
  const q4 nq{q.w + dq.w * A.dt, q.x + dq.x * A.dt, q.y + dq.y * A.dt, q.z + dq.z * A.dt};
  const double nn = sqrt(nq.w * nq.w + nq.x * nq.x + nq.y * nq.y + nq.z * nq.z);
  Q = make_double4(nq.w / nn, nq.x / nn, nq.y / nn, nq.z / nn);
  const v3 w2 = mk3(w.x, w.y, w.z) + A.hdt * ar;
  w.x = w2.x; w.y = w2.y; w.z = w2.z;
  if (KEEP) { st4h(A.P.pr + p, rn, pol.keep); st4h(A.P.vm + p, vm, pol.keep); st4h(A.P.wg + p, w, pol.keep); }
  else { st4(A.P.pr + p, rn); st4(A.P.vm + p, vm); st4(A.P.wg + p, w); }
  if (STREAM) st4h(A.P.q + p, Q, pol.stream);
  else st4(A.P.q + p, Q);
  // ghost images of this particle (own periodic copies, neighbour bricks): one bitmap word per warp says who has any
  if (FAN.bits && ((FAN.bits[p >> 5] >> (p & 31)) & 1u)) fan_out(FAN, A.P, p, rn, vm, w, Q);
  unsigned int over = 0;
  if (POLY && A.vert) {
    // polyhedra: compute_vertices (compute_vertices.hpp:26-67) of the step that has just been started, and vertex_displ_over
    // (vertex_displ_over.cu:74-134) on the same vertices — same expressions as k_compute_vertices / k_displ_over<true>
    const ShapeDev& sh = A.S.sh[type_of(w)];
    const double h = A.P.h[p];
    const q4 qn{Q.x, Q.y, Q.z, Q.w};
    bool trig = A.next_flag && A.q_backup;
    double4 rb = make_double4(0, 0, 0, 0), Qb = make_double4(1, 0, 0, 0);
    if (trig) {
      rb = ldg4(A.pr_backup + p); Qb = ldg4(A.q_backup + p);
      // Conservative early-out of vertex_displ_over: a vertex moved by at most |r - r_b| + 2 |sin(theta/2)| |lv|, theta the rotation
      // between the backup and the current orientation (|vec(q q_b^-1)| = |sin(theta/2)|), |lv| <= h * vmax.  When that bound (with a
      // 1e-9 relative margin, orders above the rounding of either side) is below the threshold no vertex can count, and the six
      // backup-vertex rotations of the exact test are skipped; in a packing that is almost every particle on almost every step.
      const v3 dr = mk3(rn.x, rn.y, rn.z) - mk3(rb.x, rb.y, rb.z);
      const v3 un = mk3(Q.y, Q.z, Q.w), ub = mk3(Qb.y, Qb.z, Qb.w);
      const v3 vq = Qb.x * un - Q.x * ub - cross(un, ub);
      const double bound = (norm(dr) + 2.0 * norm(vq) * h * sh.vmax) * 1.000000001;
      if (bound * bound < A.thr2) trig = false;
    }
    const q4 qb{Qb.x, Qb.y, Qb.z, Qb.w};
    const int nv = sh.nv, v0 = sh.v0;
    for (int k = 0; k < nv; k++) {
      const v3 lv = h * A.S.verts[v0 + k];
      const v3 vn = mk3(rn.x, rn.y, rn.z) + qrot(qn, lv);
      A.vert[(size_t)p * A.S.nvmax + k] = make_double4(vn.x, vn.y, vn.z, h * sh.radius);
      if (trig) {
        const v3 vo = mk3(rb.x, rb.y, rb.z) + qrot(qb, lv);
        const v3 d = vn - vo;
        over += dot(d, d) >= A.thr2 ? 1u : 0u;
      }
    }
  }
  if (A.next_flag) {   // particle_displ_over (backup_dem.hpp:66-100) for the step that has just been started
    if (!(POLY && A.vert) && A.pr_backup) {
      const double4 rb = STREAM ? ldg4h(A.pr_backup + p, pol.stream) : ldg4(A.pr_backup + p);
      const v3 d = mk3(rn.x, rn.y, rn.z) - mk3(rb.x, rb.y, rb.z);
      over = dot(d, d) >= A.thr2 ? 1u : 0u;
    }
    over = __reduce_add_sync(0xffffffffu, over);
    if ((threadIdx.x & 31) == 0 && over) atomicAdd(A.next_flag, (unsigned long long)over);
  }
  if (FUSE == 2 && FAN.nmail) fan_mail(FAN, A.next_flag);
}

// stress_tensor (src/interaction/stress_tensor.cu:41-81): per particle, sum over the records in contact (dn < 0) of
// tensor(fij, cp - r_i) on the i side and, for the particle-particle waves, tensor(-fij, cp - r_j) on the j side, with
// fij = fn + ft of the ContactState.  One thread per real particle walks its incidence list (fixed order -> reproducible, no
// atomics; the reference uses mat3d_atomic_add_contribution).  out: 9 doubles per particle, row-major, indexed by orig[p].
__global__ void __launch_bounds__(BLK) k_stress(GatherArgs A, double* __restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.nreal) return;
  const double4 pr = ldg4(A.P.pr + p);
  const v3 r = mk3(pr.x, pr.y, pr.z);
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const uint32_t e1 = A.inc_off[p + 1];
  for (uint32_t e = A.inc_off[p]; e < e1; e++) {
    const uint32_t code = A.inc[e], o = code >> INC_SHIFT, k = code & INC_MASK;
    const int t = A.order[o];
    const WView& w = A.W.w[t];
    const uint8_t fl = w.flg[k];
    if (!(fl & FLG_HOT)) continue;
    const double4 fd = ld4(w.fd(k));
    if (!(fd.w < 0.0)) continue;
    const double4 f4 = ld4(w.ft + k), c4 = ld4(w.cp(k));
    const v3 ft = mk3(f4.x, f4.y, f4.z);
    // ContactState fn as the reference stores it (contact_sphere.hpp:180; contact.hpp:409 for the "j to i" mesh waves)
    const v3 fn = (t == 11 || t == 12) ? (-ft) - mk3(fd.x, fd.y, fd.z) : mk3(fd.x, fd.y, fd.z) - ft;
    v3 fij = fn + ft;
    if (t <= 3 && w.ij[k].x != p) fij = -fij;   // this particle is the j of the record
    const v3 c = mk3(c4.x, c4.y, c4.z) - r;
    S[0] += fij.x * c.x; S[1] += fij.x * c.y; S[2] += fij.x * c.z;
    S[3] += fij.y * c.x; S[4] += fij.y * c.y; S[5] += fij.y * c.z;
    S[6] += fij.z * c.x; S[7] += fij.z * c.y; S[8] += fij.z * c.z;
  }
  double* dst = out + (size_t)A.P.orig[p] * 9;
#pragma unroll
  for (int q = 0; q < 9; q++) dst[q] = S[q];
}

// simulation_state (src/io/simulation_state.cu:80-175): particle part — per block, a fixed-order tree sum of
// {0.5 m v^2 (xyz), 0.5 m vrot^2 (xyz), m}; the host adds the block partials in block order (reproducible).
__global__ void __launch_bounds__(BLK) k_sim_state_particles(PView P, uint32_t n, double* __restrict__ partial) {
  __shared__ double sm[7][BLK];
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  if (p < n) {
    const double4 vm = ld4(P.vm + p), w = ld4(P.wg + p);
    const double m = vm.w;
    v[0] = 0.5 * vm.x * vm.x * m; v[1] = 0.5 * vm.y * vm.y * m; v[2] = 0.5 * vm.z * vm.z * m;   // simulation_state.hpp:52-55
    v[3] = 0.5 * w.x * w.x * m; v[4] = 0.5 * w.y * w.y * m; v[5] = 0.5 * w.z * w.z * m;
    v[6] = m;
  }
#pragma unroll
  for (int q = 0; q < 7; q++) sm[q][threadIdx.x] = v[q];
  __syncthreads();
  for (int s = BLK / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int q = 0; q < 7; q++) sm[q][threadIdx.x] += sm[q][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 7) partial[(size_t)blockIdx.x * 7 + threadIdx.x] = sm[threadIdx.x][0];
}
// interaction part (itools/counter.hpp:125-152): out[0] += total, out[1] += active (integer atomics: order-independent),
// out[2] = min over the active records of dn, kept as the bit pattern of a non-positive double (larger pattern = smaller value)
__global__ void __launch_bounds__(BLK) k_sim_state_wave(WView w, unsigned long long coef, unsigned long long* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int tot = 0, act = 0;
  double dn = 0.0;
  if (k < w.n) {
    const uint8_t fl = w.flg[k];
    if (((fl >> 1) & 3) != 2) {   // PartnerGhost twins are not counted
      tot = 1;
      if (fl & FLG_HOT) {
        const double4 fd = ld4(w.fd(k)), ft = ld4(w.ft + k);
        bool active = fd.w < 0.0 || ft.x != 0.0 || ft.y != 0.0 || ft.z != 0.0;
        if (!active && (fl & FLG_MOM)) {
          const double4 mo = ld4(w.mo + k);
          active = mo.x != 0.0 || mo.y != 0.0 || mo.z != 0.0;
        }
        if (active) { act = 1; dn = fd.w; }
      }
    }
  }
  tot = __reduce_add_sync(0xffffffffu, tot);
  act = __reduce_add_sync(0xffffffffu, act);
  // min(0, dn): dn <= 0 values order inversely to their bit patterns
  unsigned long long bits = dn < 0.0 ? (unsigned long long)__double_as_longlong(dn) : 0ull;
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
    bits = other > bits ? other : bits;
  }
  if ((threadIdx.x & 31) == 0) {
    if (tot) atomicAdd(out, coef * tot);
    if (act) atomicAdd(out + 1, coef * act);
    if (bits) atomicMax(out + 2, bits);
  }
}

// compute_driver_force_moment.cu:34-54 — per-driver sum of -(fn+ft) and of (cp - center) x (-(fn+ft)) - moment over the driver waves.
// Deterministic: one block per driver wave chunk writes a partial; partials are summed in order on the host side buffer.
__global__ void __launch_bounds__(BLK) k_driver_force(WView w, int type, const Driver* __restrict__ drv, int ndrv, double* __restrict__ partial) {
  // partial layout: [block][driver][6]
  extern __shared__ double sm[];
  double* acc = sm;  // ndrv*6 per warp
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < nw * ndrv * 6; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  v3 f = mk3(0, 0, 0), m = mk3(0, 0, 0);
  int d = -1;
  if (k < w.n) {
    if (w.flg[k] & FLG_HOT) {
      const double4 mo = w.mo[k], fd = *w.fd(k), cp = *w.cp(k), ft4 = w.ft[k];
      d = (int)w.ij[k].y;
      const v3 ft = mk3(ft4.x, ft4.y, ft4.z);
      // ContactState fn (contact_sphere.hpp:180; contact.hpp:409 for the "j to i" mesh waves whose stored friction is negated)
      const v3 fn = (type == 11 || type == 12) ? (-ft) - mk3(fd.x, fd.y, fd.z) : mk3(fd.x, fd.y, fd.z) - ft;
      f = -(fn + ft);                              // compute_driver_force_moment.cu:46-47
      m = cross(mk3(cp.x, cp.y, cp.z) - drv[d].center, f) + -mk3(mo.x, mo.y, mo.z);  // :48-49
    }
  }
  // serial, lane-ordered accumulation inside each warp keeps the sum order fixed
  for (int l = 0; l < 32; l++) {
    const int dl = __shfl_sync(0xffffffffu, d, l);
    const double fx = __shfl_sync(0xffffffffu, f.x, l), fy = __shfl_sync(0xffffffffu, f.y, l), fz = __shfl_sync(0xffffffffu, f.z, l);
    const double mx = __shfl_sync(0xffffffffu, m.x, l), my = __shfl_sync(0xffffffffu, m.y, l), mz = __shfl_sync(0xffffffffu, m.z, l);
    if (lane == 0 && dl >= 0) {
      double* a = acc + (wid * ndrv + dl) * 6;
      a[0] += fx; a[1] += fy; a[2] += fz; a[3] += mx; a[4] += my; a[5] += mz;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ndrv * 6; i += blockDim.x) {
    double s = 0;
    for (int ww = 0; ww < nw; ww++) s += acc[ww * ndrv * 6 + i];
    partial[(size_t)blockIdx.x * ndrv * 6 + i] = s;
  }
}

}  // namespace edb
