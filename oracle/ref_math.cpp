// oracle/_ref/libexadem_refmath.so — TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
//
// This translation unit does not restate anything: it #includes the reference's OWN header-only
// arithmetic where it lies under /root/reference (never copied into this repo) and exposes it
// through a flat C ABI so that tests can pin the CPU restatement (oracle/dem_oracle.cpp) bit-for-bit
// against the real reference functions on random inputs.
//
// Headers compiled in place:
//   src/shape/include/exaDEM/shape_detection_core.hpp   (filters + detection cores)
//   src/shape/include/exaDEM/shape_detection_api.hpp    (vertex-array overloads, detect<N>)
//   src/shape/include/exaDEM/type/OBB.hpp, basic_types.hpp (OBB::intersect, compute_obb)
//   src/forcefield/include/exaDEM/forcefield/{common_kernels,contact_force}.hpp (Hooke law)
//   src/driver/include/exaDEM/{cylinder,surface,ball}.hpp (driver filter/detector)
//   src/driver/include/exaDEM/{cylinder,surface,ball,rshape,driver_params}.hpp (driver motion integrators: push_f_v_r, push_f_v,
//     push_av_to_quat, force_to_accel / f_ra) — applied exactly as the operators of src/driver/push_*_driver.cpp, force_to_accel_driver.cpp do
// The Onika/exaNBody/yaml-cpp/tinyexpr/MPI headers they include are absent from /root/reference;
// oracle/shim/ provides minimal stand-ins (vector algebra = working definition, parity risk R1).
#include <exaDEM/shape_detection.hpp>
#include <exaDEM/forcefield/contact_force.hpp>
#include <exaDEM/basic_types.hpp>
#include <exaDEM/cylinder.hpp>
#include <exaDEM/surface.hpp>
#include <exaDEM/ball.hpp>
#include <exaDEM/rshape.hpp>
#include <exaDEM/interaction/interaction_pair.hpp>
#include <exaDEM/forcefield/inner_bond_parameters.hpp>
#include <exaDEM/forcefield/inner_bond_force.hpp>
#include "../include/exadem_b200.h"
#include <cstring>

using exanb::Vec3d;
using exanb::Quaternion;

static inline Vec3d V(const double* p) { return Vec3d{p[0], p[1], p[2]}; }
static inline void S(double* o, const Vec3d& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static inline int pack(const exaDEM::contact& c, double* out) {
  out[0] = c.dn; S(out + 1, c.normal); S(out + 4, c.position);
  return c.is_contact ? 1 : 0;
}

// Build a reference `shape` from flat arrays (vertices xyz, edges pairs, faces as [n, v...] stream).
static exaDEM::shape make_shape(int nv, const double* verts, int ne, const int* edges, int nf, const int* faces_stream,
                                int faces_len, double radius) {
  exaDEM::shape s;
  s.vertices_.resize(nv);
  for (int i = 0; i < nv; i++) s.vertices_[i] = V(verts + 3 * i);
  s.edges_.assign(edges, edges + 2 * ne);
  s.faces_.clear();
  s.faces_.push_back(nf);
  s.faces_.insert(s.faces_.end(), faces_stream, faces_stream + faces_len);
  s.radius_ = radius;
  s.compute_offset_faces();
  return s;
}

extern "C" {

int ref_filter_vertex_vertex(double rVerlet, const double* vi, double ri, const double* vj, double rj) {
  return exaDEM::filter_vertex_vertex(rVerlet, V(vi), ri, V(vj), rj);
}
int ref_detection_vertex_vertex_core(const double* pi, double ri, const double* pj, double rj, double* out7) {
  return pack(exaDEM::detection_vertex_vertex_core(V(pi), ri, V(pj), rj), out7);
}
int ref_filter_vertex_edge_core(double rVerlet, const double* vi, double ri, const double* vf, const double* vs, double rj) {
  return exaDEM::filter_vertex_edge_core(rVerlet, V(vi), ri, V(vf), V(vs), rj);
}
int ref_detection_vertex_edge_core(const double* vi, double ri, const double* vf, const double* vs, double rj, double* out7) {
  return pack(exaDEM::detection_vertex_edge_core(V(vi), ri, V(vf), V(vs), rj), out7);
}
int ref_filter_edge_edge_core(double rVerlet, const double* vfi, const double* vsi, double ri, const double* vfj,
                              const double* vsj, double rj) {
  return exaDEM::filter_edge_edge_core(rVerlet, V(vfi), V(vsi), ri, V(vfj), V(vsj), rj);
}
int ref_detection_edge_edge_core(const double* vfi, const double* vsi, double ri, const double* vfj, const double* vsj,
                                 double rj, double* out7) {
  return pack(exaDEM::detection_edge_edge_core(V(vfi), V(vsi), ri, V(vfj), V(vsj), rj), out7);
}

// vertex-face against world-space vertex array `vaj` (nvj*3) of a shape given by its face stream.
int ref_filter_vertex_face(double rVerlet, const double* vi, double ri, int nvj, const double* vaj, double hj, int j,
                           int nf, const int* faces_stream, int faces_len, double radius_j) {
  std::vector<int> e;
  std::vector<double> dummy(3 * nvj, 0.0);
  exaDEM::shape s = make_shape(nvj, dummy.data(), 0, e.data(), nf, faces_stream, faces_len, radius_j);
  std::vector<Vec3d> va(nvj);
  for (int k = 0; k < nvj; k++) va[k] = V(vaj + 3 * k);
  return exaDEM::filter_vertex_face(rVerlet, V(vi), ri, va, hj, j, &s);
}
int ref_detection_vertex_face_core(const double* vi, double ri, int nvj, const double* vaj, double hj, int j, int nf,
                                   const int* faces_stream, int faces_len, double radius_j, double* out7) {
  std::vector<int> e;
  std::vector<double> dummy(3 * nvj, 0.0);
  exaDEM::shape s = make_shape(nvj, dummy.data(), 0, e.data(), nf, faces_stream, faces_len, radius_j);
  std::vector<Vec3d> va(nvj);
  for (int k = 0; k < nvj; k++) va[k] = V(vaj + 3 * k);
  return pack(exaDEM::detection_vertex_face_core(V(vi), ri, va, hj, j, &s), out7);
}

// shape::get_vertex (world vertex p + q*(h*v)) — pins the shim's quaternion rotation against quat.hpp too.
void ref_get_vertex(const double* local_v, const double* p, double h, const double* q_wxyz, double* out3) {
  exaDEM::shape s;
  s.vertices_.push_back(V(local_v));
  S(out3, s.get_vertex(0, V(p), h, Quaternion{q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]}));
}
// the vendored Rockable quat rotation (type/quat.hpp:79-90), for cross-checking the shim definition
void ref_rockable_quat_rotate(const double* q_wxyz, const double* v, double* out3) {
  quat q(q_wxyz[1], q_wxyz[2], q_wxyz[3], q_wxyz[0]);
  vec3r r = q * vec3r(v[0], v[1], v[2]);
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// OBB layout (15 doubles): center, e1, e2, e3, extent
static OBB obb_from(const double* a) {
  OBB o;
  o.center.set(a[0], a[1], a[2]); o.e1.set(a[3], a[4], a[5]); o.e2.set(a[6], a[7], a[8]); o.e3.set(a[9], a[10], a[11]);
  o.extent.set(a[12], a[13], a[14]);
  return o;
}
static void obb_to(const OBB& o, double* a) {
  a[0] = o.center.x; a[1] = o.center.y; a[2] = o.center.z; a[3] = o.e1.x; a[4] = o.e1.y; a[5] = o.e1.z;
  a[6] = o.e2.x; a[7] = o.e2.y; a[8] = o.e2.z; a[9] = o.e3.x; a[10] = o.e3.y; a[11] = o.e3.z;
  a[12] = o.extent.x; a[13] = o.extent.y; a[14] = o.extent.z;
}
int ref_obb_intersect(const double* a15, const double* b15) { return obb_from(a15).intersect(obb_from(b15)); }
void ref_compute_obb(const double* in15, const double* pos, const double* q_wxyz, double h, double enlarge, double* out15) {
  OBB o = exaDEM::compute_obb(obb_from(in15), V(pos), Quaternion{q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]}, h);
  o.enlarge(enlarge);
  obb_to(o, out15);
}

// contact_force_core<Hooke, law>: law 0 = None, 1 = Cohesive, 2 = DMT. params = {kn,kt,kr,mu,damp_rate,fc,dncut,gamma}
void ref_contact_force_core(int law, double dn, const double* n, double dt, const double* params, double meff, double reff,
                            double* ft_inout, const double* cp, const double* pos_i, const double* vel_i, double* f_out,
                            double* mom_inout, const double* vrot_i, const double* pos_j, const double* vel_j,
                            const double* vrot_j) {
  exaDEM::ContactParams hkp;
  hkp.kn_ = params[0]; hkp.kt_ = params[1]; hkp.kr_ = params[2]; hkp.mu_ = params[3];
  hkp.damp_rate_ = params[4]; hkp.fc_ = params[5]; hkp.dncut_ = params[6]; hkp.gamma_ = params[7];
  Vec3d ft = V(ft_inout), f = V(f_out), mom = V(mom_inout);
  if (law == 0)
    exaDEM::contact_force_core<ContactLawType::Hooke, CohesiveLawType::None>(dn, V(n), dt, hkp, meff, reff, ft, V(cp), V(pos_i), V(vel_i), f, mom, V(vrot_i), V(pos_j), V(vel_j), V(vrot_j));
  else if (law == 1)
    exaDEM::contact_force_core<ContactLawType::Hooke, CohesiveLawType::Cohesive>(dn, V(n), dt, hkp, meff, reff, ft, V(cp), V(pos_i), V(vel_i), f, mom, V(vrot_i), V(pos_j), V(vel_j), V(vrot_j));
  else
    exaDEM::contact_force_core<ContactLawType::Hooke, CohesiveLawType::DMT>(dn, V(n), dt, hkp, meff, reff, ft, V(cp), V(pos_i), V(vel_i), f, mom, V(vrot_i), V(pos_j), V(vel_j), V(vrot_j));
  S(ft_inout, ft); S(f_out, f); S(mom_inout, mom);
}
void ref_compute_moments(const double* cp, const double* p, const double* f, const double* m, double* out3) {
  S(out3, exaDEM::compute_moments(V(cp), V(p), V(f), V(m)));
}

// drivers: kind 0 = cylinder {radius, axis[3], center[3] (already multiplied by axis)}, 1 = surface {offset, normal[3]},
// 2 = ball {radius, center[3]}
int ref_driver_filter(int kind, const double* prm, double rcut, const double* p) {
  if (kind == 0) { exaDEM::Cylinder d; d.fields_.radius_ = prm[0]; d.fields_.axis_ = V(prm + 1); d.fields_.center_ = V(prm + 4); return d.filter(rcut, V(p)); }
  if (kind == 1) { exaDEM::Surface d; d.fields_.offset_ = prm[0]; d.fields_.normal_ = V(prm + 1); d.fields_.center_proj_ = d.fields_.offset_ * d.fields_.normal_; return d.filter(rcut, V(p)); }
  exaDEM::BallFields bf; bf.radius_ = prm[0]; bf.center_ = V(prm + 1); exaDEM::MotionType mt = exaDEM::STATIONARY; exaDEM::Ball d(bf, mt); return d.filter(rcut, V(p));
}
int ref_driver_detector(int kind, const double* prm, double rcut, const double* p, double* out7) {
  std::tuple<bool, double, Vec3d, Vec3d> r;
  if (kind == 0) { exaDEM::Cylinder d; d.fields_.radius_ = prm[0]; d.fields_.axis_ = V(prm + 1); d.fields_.center_ = V(prm + 4); r = d.detector(rcut, V(p)); }
  else if (kind == 1) { exaDEM::Surface d; d.fields_.offset_ = prm[0]; d.fields_.normal_ = V(prm + 1); d.fields_.center_proj_ = d.fields_.offset_ * d.fields_.normal_; r = d.detector(rcut, V(p)); }
  else { exaDEM::BallFields bf; bf.radius_ = prm[0]; bf.center_ = V(prm + 1); exaDEM::MotionType mt = exaDEM::STATIONARY; exaDEM::Ball d(bf, mt); r = d.detector(rcut, V(p)); }
  out7[0] = std::get<1>(r); S(out7 + 1, std::get<2>(r)); S(out7 + 4, std::get<3>(r));
  return std::get<0>(r) ? 1 : 0;
}

// ---- driver motion: one operator of the time loop applied to ONE driver by the reference's own member functions ----
// op 0 = push_f_v_r_driver (PushFVRDriverFunc, push_f_v_r_driver.cpp:28-37), 1 = push_f_v_driver (push_f_v_driver.cpp:28-35),
// 2 = push_av_to_quat_driver (push_av_to_quat_driver.cpp:27-36), 3 = force_to_accel_driver's ForceToAccelDriverFunc
// (force_to_accel_driver.cpp:30-45) with fields_.forces_ = forces.  In/out through the edb_driver / edb_driver_motion structs.
}  // extern "C"
static exaDEM::Driver_params mk_params(const edb_driver_motion& m) {
  exaDEM::Driver_params p;
  p.motion_vector_ = V(m.motion_vector); p.motion_start_threshold_ = m.motion_start_threshold; p.motion_end_threshold_ = m.motion_end_threshold;
  p.const_vel_ = m.const_vel; p.const_force_ = m.const_force; p.sigma_ = m.sigma; p.damprate_ = m.damprate;
  for (int k = 0; k < m.n_tab; k++) { p.tab_time_.push_back(m.tab_time[k]); p.tab_pos_.push_back(V(m.tab_pos + 3 * k)); }
  p.omega_ = m.omega; p.amplitude_ = m.amplitude; p.shaker_dir_ = V(m.shaker_dir);
  p.pendulum_anchor_point_ = V(m.pendulum_anchor_point); p.pendulum_initial_position_ = V(m.pendulum_initial_position);
  p.pendulum_swing_dir_ = V(m.pendulum_swing_dir);
  return p;
}
template <class D>
static void apply_op(D& drv, exaDEM::Driver_params& p, int op, double time, double dt, double system_mass) {
  if (op == 0) { if (p.is_motion_triggered(time + dt)) drv.push_f_v_r(p, time, dt); }
  else if (op == 1) drv.push_f_v(p, dt);
  else if (op == 3) {
    if constexpr (std::is_same_v<D, exaDEM::Ball>) { p.mass_ = system_mass; drv.f_ra(p, dt); }
    if constexpr (std::is_same_v<D, exaDEM::Surface>) { p.mass_ = system_mass; }
    drv.force_to_accel(p);
  }
}
extern "C" {
int ref_driver_motion_op(int op, double time, double dt, double system_mass, edb_driver* d, edb_driver_motion* m) {
  exaDEM::Driver_params p = mk_params(*m);
  const exaDEM::MotionType mt = (exaDEM::MotionType)m->motion_type;
  if (d->kind == EDB_DRIVER_CYLINDER) {
    exaDEM::Cylinder c; c.motion_type_ = mt;
    c.fields_.radius_ = d->radius; c.fields_.axis_ = V(d->axis); c.fields_.center_ = V(d->center); c.fields_.vel_ = V(d->vel); c.fields_.vrot_ = V(d->vrot);
    apply_op(c, p, op, time, dt, system_mass);
    S(d->center, c.fields_.center_); S(d->vel, c.fields_.vel_);
  } else if (d->kind == EDB_DRIVER_SURFACE) {
    exaDEM::Surface s; s.motion_type_ = mt;
    s.fields_.offset_ = d->offset; s.fields_.normal_ = V(d->normal); s.fields_.center_ = V(d->center); s.fields_.vel_ = V(d->vel); s.fields_.vrot_ = V(d->vrot);
    s.fields_.forces_ = V(m->forces); s.fields_.mass_ = m->mass; s.fields_.surface_ = m->surface; s.fields_.acc_ = m->acc[0];
    s.fields_.center_proj_ = V(m->mom);   // center_proj_ travels in the (otherwise unused for a surface) mom slot of this test interface
    apply_op(s, p, op, time, dt, system_mass);
    d->offset = s.fields_.offset_; S(d->normal, s.fields_.normal_); S(d->center, s.fields_.center_); S(d->vel, s.fields_.vel_);
    m->acc[0] = s.fields_.acc_; S(m->mom, s.fields_.center_proj_);
  } else if (d->kind == EDB_DRIVER_BALL) {
    exaDEM::BallFields bf; bf.radius_ = d->radius; bf.center_ = V(d->center); bf.vel_ = V(d->vel); bf.vrot_ = V(d->vrot);
    bf.forces_ = V(m->forces); bf.mass_ = m->mass; bf.rv_ = m->rv; bf.ra_ = m->ra; bf.acc_ = V(m->acc);
    exaDEM::MotionType mtv = mt;
    exaDEM::Ball b(bf, mtv);
    apply_op(b, p, op, time, dt, system_mass);
    d->radius = b.fields_.radius_; S(d->center, b.fields_.center_); S(d->vel, b.fields_.vel_);
    m->rv = b.fields_.rv_; m->ra = b.fields_.ra_; S(m->acc, b.fields_.acc_); S(m->forces, b.fields_.forces_);
  } else if (d->kind == EDB_DRIVER_RSHAPE) {
    exaDEM::RShapeDriver r; r.motion_type_ = mt;
    auto& f = r.fields_;
    f.center_ = V(d->center); f.vel_ = V(d->vel); f.vrot_ = V(d->vrot); f.forces_ = V(m->forces);
    f.quat_ = Quaternion{d->quat[0], d->quat[1], d->quat[2], d->quat[3]};
    f.acc_ = V(m->acc); f.surface_ = m->surface; f.mass_ = m->mass; f.drive_by_mom_ = m->drive_by_mom != 0;
    f.applied_mom_ = V(m->applied_mom); f.mom_ = V(m->mom); f.inertia_ = V(m->inertia); f.mom_axis_ = V(m->mom_axis);
    if (op == 2) r.push_av_to_quat(p, time, dt);
    else apply_op(r, p, op, time, dt, system_mass);
    S(d->center, f.center_); S(d->vel, f.vel_); S(d->vrot, f.vrot_);
    d->quat[0] = f.quat_.w; d->quat[1] = f.quat_.x; d->quat[2] = f.quat_.y; d->quat[3] = f.quat_.z;
    S(m->acc, f.acc_); S(m->forces, f.forces_); S(m->mom, f.mom_); S(m->applied_mom, f.applied_mom_); S(m->mom_axis, f.mom_axis_);
  }
  return 0;
}

// inner bond law (src/forcefield/include/exaDEM/forcefield/inner_bond_force.hpp:26-70), the reference's own force_law_core
void ref_inner_bond_force_core(double dn, const double* n, double dn0, double weight, double dt, const edb_inner_bond_params* p, double meff,
                               double* en_et, double* tds, double* ft, const double* cp, const double* pi, const double* vi, double* fi,
                               const double* wi, const double* pj, const double* vj, const double* wj) {
  exaDEM::InnerBondParams P;
  P.kn_ = p->kn; P.kt_ = p->kt; P.damp_rate_ = p->damp_rate; P.mode_ = (exaDEM::RuptureMode)p->mode; P.crit1_ = p->crit1; P.crit2_ = p->crit2;
  Vec3d T = V(tds), F = V(ft), f{0, 0, 0};
  exaDEM::force_law_core(dn, V(n), dn0, weight, dt, P, meff, en_et[0], T, en_et[1], F, V(cp), V(pi), V(vi), f, V(wi), V(pj), V(vj), V(wj));
  S(tds, T); S(ft, F); S(fi, f);
}

// InteractionPair ordering / equality (src/interaction/include/exaDEM/interaction/interaction_pair.hpp:186-249): the sort key of the
// per-particle interaction lists and the match key of the history carry-over.  rec = {id_i, id_j, sub_i, sub_j, type, swap}
static exaDEM::InteractionPair mk_pair(const unsigned long long* r) {
  exaDEM::InteractionPair p;
  p.pi_.id_ = r[0]; p.pj_.id_ = r[1]; p.pi_.sub_ = (uint32_t)r[2]; p.pj_.sub_ = (uint32_t)r[3]; p.type_ = (uint16_t)r[4]; p.swap_ = (uint8_t)r[5];
  p.pi_.cell_ = 0; p.pj_.cell_ = 0; p.pi_.p_ = 0; p.pj_.p_ = 0; p.ghost_ = 0;
  return p;
}
int ref_pair_less(const unsigned long long* a, const unsigned long long* b) { return mk_pair(a) < mk_pair(b) ? 1 : 0; }
int ref_pair_equal(const unsigned long long* a, const unsigned long long* b) { return mk_pair(a) == mk_pair(b) ? 1 : 0; }

// face areas of a shape (shape.hpp:673-697 compute_face_areas, :359-362 get_face_area(i, h)) — input of the rupture criterion
void ref_face_areas(int nv, const double* verts, int ne, const int* edges, int nf, const int* faces_stream, int faces_len, double h, double* out) {
  exaDEM::shape s = make_shape(nv, verts, ne, edges, nf, faces_stream, faces_len, 0.0);
  s.compute_face_areas();
  for (int f = 0; f < nf; f++) out[f] = s.get_face_area(f, h);
}

}  // extern "C"
