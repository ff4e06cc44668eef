// tests/host_check_driver.cu — TEST HELPER: the product's driver integrators (exadem_b200/csrc/kernels_driver.cuh, __host__ __device__)
// compiled for the HOST, behind the same single-driver interface as oracle/dem_oracle.cpp:orc_driver_motion_op, so that the CPU suite can
// compare the code the GPU runs with the oracle bit for bit without a GPU (device sin/cos aside).  Built by tests/test_survey8f_rank3_driver_motion.py:
//   nvcc -O2 -std=c++17 -fmad=false -Xcompiler -fPIC -shared tests/host_check_driver.cu -o tests/_build/libhost_check_driver.so
#include "../exadem_b200/csrc/kernels_driver.cuh"
using namespace edb;
static v3 V(const double* p) { return mk3(p[0], p[1], p[2]); }
static void S(double* o, v3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static void to_dev(const edb_driver* d, const edb_driver_motion* s, Driver& D, DriverMotion& m);
static void from_dev(const Driver& D, const DriverMotion& m, edb_driver* d, edb_driver_motion* s);
// the fused step-boundary kernel body (driver_cycle_body, what k_driver_cycle runs per driver) on one driver
extern "C" int chk_driver_cycle(edb_driver* d, edb_driver_motion* s, const double* g, const double* sums6, int sphere_mode, double system_mass,
                                double dt, double time_next, int do_end, int do_begin) {
  Driver D{};
  DriverMotion m{};
  to_dev(d, s, D, m);
  DriverCycleArgs A{&D, &m, 1, sums6, sphere_mode, do_end, do_begin, dt, system_mass, time_next, nullptr, mk3(g[0], g[1], g[2])};
  driver_cycle_body(A, 0);
  from_dev(D, m, d, s);
  return 0;
}
extern "C" int chk_driver_motion_op(int op, double time, double dt, double system_mass, edb_driver* d, edb_driver_motion* s) {
  Driver D{};
  DriverMotion m{};
  to_dev(d, s, D, m);
  if (op == 0) { if (motion_triggered(m, time + dt)) driver_push_f_v_r(D, m, time, dt); }
  else if (op == 1) driver_push_f_v(D, m, dt);
  else if (op == 2) driver_push_av_to_quat(D, m, time, dt);
  else driver_force_to_accel(D, m, system_mass);
  from_dev(D, m, d, s);
  return 0;
}
static void to_dev(const edb_driver* d, const edb_driver_motion* s, Driver& D, DriverMotion& m) {
  D.kind = d->kind; D.radius = d->radius; D.axis = V(d->axis); D.center = V(d->center); D.offset = d->offset; D.normal = V(d->normal);
  D.center_proj = D.offset * D.normal; D.vel = V(d->vel); D.vrot = V(d->vrot);
  m.quat = q4{d->quat[0], d->quat[1], d->quat[2], d->quat[3]};
  m.motion = s->motion_type; m.n_tab = s->motion_type == EDB_MOTION_TABULATED ? s->n_tab : 0;
  m.expr_use_v = s->expr_use_v; m.expr_use_vrot = s->expr_use_vrot; m.expr_use_mom = s->expr_use_mom; m.drive_by_mom = s->drive_by_mom;
  m.motion_vector = V(s->motion_vector); m.t_start = s->motion_start_threshold; m.t_end = s->motion_end_threshold;
  m.const_vel = s->const_vel; m.const_force = s->const_force; m.sigma = s->sigma; m.damprate = s->damprate;
  m.omega = s->omega; m.amplitude = s->amplitude; m.shaker_dir = V(s->shaker_dir);
  m.pend_anchor = V(s->pendulum_anchor_point); m.pend_init = V(s->pendulum_initial_position); m.pend_dir = V(s->pendulum_swing_dir);
  m.mass = s->mass; m.surface = s->surface; m.rv = s->rv; m.ra = s->ra;
  m.acc = V(s->acc); m.inertia = V(s->inertia); m.applied_mom = V(s->applied_mom); m.mom_axis = V(s->mom_axis); m.mom = V(s->mom);
  m.forces = V(s->forces);
  for (int k = 0; k < m.n_tab; k++) { m.tab_time[k] = s->tab_time[k]; m.tab_pos[k] = V(s->tab_pos + 3 * k); }
  if (D.kind == EDB_DRIVER_SURFACE) { D.center_proj = V(s->mom); m.mom = mk3(0, 0, 0); }   // test interface: center_proj_ in the mom slot
}
static void from_dev(const Driver& D, const DriverMotion& m, edb_driver* d, edb_driver_motion* s) {
  d->radius = D.radius; S(d->center, D.center); d->offset = D.offset; S(d->normal, D.normal); S(d->vel, D.vel); S(d->vrot, D.vrot);
  d->quat[0] = m.quat.w; d->quat[1] = m.quat.x; d->quat[2] = m.quat.y; d->quat[3] = m.quat.z;
  S(s->acc, m.acc); s->rv = m.rv; s->ra = m.ra; S(s->mom, D.kind == EDB_DRIVER_SURFACE ? D.center_proj : m.mom); S(s->forces, m.forces);
  S(s->applied_mom, m.applied_mom); S(s->mom_axis, m.mom_axis);
}
