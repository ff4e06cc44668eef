// kernels_driver.cuh — the `drivers` slot lives on the device: motion integrators of the time loop, mesh vertices, rebuild trigger
// and the per-cell primitive lists of the RShape meshes.
//
// Replaces the reference's host operators push_f_v_r_driver / push_f_v_driver / push_av_to_quat_driver / reset_force_driver /
// force_to_accel_driver (src/driver/*.cpp, config_numerical_schemes.msp:13-24), driver_vertices (driver_vertices.cu), driver_displ_over
// (driver_displ_over.cu:60-243) and grid_rshape_driver (grid_stl_mesh.cpp:89-258).  The integrators are O(#drivers): one thread per
// driver, kept on the device only so that the time loop needs no host round trip; the mesh kernels are one thread per vertex / primitive.
#pragma once
#include "common.cuh"

namespace edb {

#define EDB_MAX_TAB_DEV 64   // == EDB_MAX_TAB (include/exadem_b200.h)

enum Motion { M_STATIONARY, M_LINEAR_MOTION, M_COMPRESSIVE_FORCE, M_LINEAR_FORCE_MOTION, M_PARTICLE, M_LINEAR_COMPRESSIVE_MOTION, M_TABULATED,
              M_SHAKER, M_PENDULUM_MOTION, M_EXPRESSION, M_UNKNOWN };

// Driver_params (driver_params.hpp:33-66) + the integrator state of SurfaceFields / BallFields / RShapeDriverFields that the contact
// kernels do not read (those stay in `Driver`).
struct DriverMotion {
  int motion, n_tab;
  int expr_use_v, expr_use_vrot, expr_use_mom, drive_by_mom;
  v3 motion_vector;
  double t_start, t_end, const_vel, const_force, sigma, damprate, pmass;
  double omega, amplitude;
  v3 shaker_dir, pend_anchor, pend_init, pend_dir;
  v3 expr_v, expr_vrot, expr_mom;
  double mass, surface, rv, ra;
  v3 acc, inertia, applied_mom, mom_axis, mom, forces;
  q4 quat;
  double tab_time[EDB_MAX_TAB_DEV];
  v3 tab_pos[EDB_MAX_TAB_DEV];
};

// motion_type.hpp:84-141
EDB_HD bool m_compressive(int m) { return m == M_COMPRESSIVE_FORCE || m == M_LINEAR_COMPRESSIVE_MOTION; }
EDB_HD bool m_force_motion(int m) { return m == M_PARTICLE || m == M_LINEAR_FORCE_MOTION; }
EDB_HD bool m_need_forces(int m) { return m_compressive(m) || m == M_PARTICLE || m == M_LINEAR_FORCE_MOTION; }
EDB_HD bool motion_triggered(const DriverMotion& m, double t) { return t >= m.t_start && t <= m.t_end; }   // driver_params.hpp:82-84
EDB_HD bool rshape_need_moment(const DriverMotion& m) { return m.drive_by_mom || m.motion == M_PARTICLE; }   // rshape.hpp:406-413
EDB_HD bool driver_moves(const Driver& d, const DriverMotion& m) {
  return m.motion != M_STATIONARY || (d.kind == EDB_DRIVER_RSHAPE && rshape_need_moment(m));
}

// driver_params.hpp:277-310.  std::lower_bound = first entry >= time; time <= tab_time[0] indexes tab_time_[-1] in the reference (UB):
// defined here as "first entry, not moving yet".
EDB_HD int tab_lower(const DriverMotion& m, double time) {
  int i = 0;
  while (i < m.n_tab && m.tab_time[i] < time) i++;
  return i;
}
EDB_HD v3 tab_to_position(const DriverMotion& m, double time) {
  const int it = tab_lower(m, time);
  if (it == m.n_tab) return m.tab_pos[m.n_tab - 1];
  if (it == 0) return m.tab_pos[0];
  const int lo = it - 1, up = it;
  const double Dt = (time - m.tab_time[lo]) / (m.tab_time[up] - m.tab_time[lo]);
  return (m.tab_pos[up] - m.tab_pos[lo]) * Dt + m.tab_pos[lo];
}
EDB_HD v3 tab_to_velocity(const DriverMotion& m, double time) {
  const int it = tab_lower(m, time);
  if (it == m.n_tab || it == 0) return mk3(0, 0, 0);
  const int lo = it - 1, up = it;
  const double Dt = m.tab_time[up] - m.tab_time[lo];
  return (m.tab_pos[up] - m.tab_pos[lo]) / Dt;
}
// driver_params.hpp:313-329, 366-370
EDB_HD double shaker_signal(const DriverMotion& m, double time) { time -= m.t_start; return m.amplitude * sin(m.omega * time); }
EDB_HD v3 shaker_velocity(const DriverMotion& m, double time) { time -= m.t_start; return m.amplitude * m.omega * cos(m.omega * time) * m.shaker_dir; }
// driver_params.hpp:340-364
EDB_HD void pendulum_offset_normal(const DriverMotion& m, double time, double& offset, v3& normal) {
  const v3 v1 = m.pend_anchor, v2 = m.pend_init;
  const v3 p3 = m.pend_init + m.pend_dir * shaker_signal(m, time);
  if (dot(p3, v2) < 1e-16) { offset = 0.0; normal = m.pend_dir; return; }
  v3 v1v3 = p3 - v1;
  v1v3 = v1v3 / norm(v1v3);
  const v3 proj = dot(v1v3, v2 - v1) * v1v3 + v1;
  const v3 dir = proj - v2;
  normal = dir / norm(dir);
  offset = dot(v1, normal);
}
// driver_params.hpp:68-74
EDB_HD void update_forces(const DriverMotion& m, v3& forces) {
  if (m.motion == M_LINEAR_FORCE_MOTION) forces = (dot(forces, m.motion_vector) + m.const_force) * m.motion_vector;
  else if (m.motion != M_PARTICLE) forces = mk3(0, 0, 0);
}

// push_f_v_r: cylinder.hpp:160-173, surface.hpp:262-299, ball.hpp:203-213, rshape.hpp:283-320
EDB_HD void driver_push_f_v_r(Driver& d, DriverMotion& m, double time, double dt) {
  const int mt = m.motion;
  if (d.kind == EDB_DRIVER_CYLINDER) {
    if (mt == M_SHAKER) {
      const double displ = shaker_signal(m, time + dt) - shaker_signal(m, time);
      d.center = d.center + displ * m.shaker_dir;
      d.vel = shaker_velocity(m, time + dt);
    }
  } else if (d.kind == EDB_DRIVER_SURFACE) {
    if (mt == M_STATIONARY) return;
    if (mt == M_PENDULUM_MOTION) {
      double off;
      v3 nrm;
      pendulum_offset_normal(m, time + dt, off, nrm);
      d.normal = nrm;
      d.offset = off;
      d.vel = mk3(0, 0, 0);
      d.center_proj = d.normal * d.offset;
      return;
    }
    double displ = dt * dot(d.vel, d.normal) + 0.5 * dt * dt * m.acc.x;
    if (mt == M_SHAKER) {
      const double sn = shaker_signal(m, time + dt), sc = shaker_signal(m, time);
      const double af = dot(m.shaker_dir, d.normal);
      displ = (sn - sc) * af;
      d.vel = shaker_velocity(m, time + dt);
    }
    d.center = d.center + displ * d.normal;
    d.offset += displ;
    d.center_proj = d.center_proj + displ * d.normal;
  } else if (d.kind == EDB_DRIVER_BALL) {
    if (mt == M_TABULATED) {
      d.center = tab_to_position(m, time);
      d.vel = tab_to_velocity(m, time);
    } else if (mt != M_STATIONARY) {
      if (m_compressive(mt)) d.radius += dt * m.rv + 0.5 * dt * dt * m.ra;
      d.center = d.center + dt * d.vel;
    }
  } else if (d.kind == EDB_DRIVER_RSHAPE) {
    if (mt == M_TABULATED) {
      d.center = tab_to_position(m, time);
      d.vel = tab_to_velocity(m, time);
    } else if (mt != M_STATIONARY) {
      if (mt == M_SHAKER) {
        d.vel = cmul(shaker_velocity(m, time + dt), m.shaker_dir);   // Vec3d * Vec3d, rshape.hpp:302
        m.acc = mk3(0, 0, 0);
      }
      if (mt == M_EXPRESSION && motion_triggered(m, time)) {
        if (m.expr_use_v) d.vel = m.expr_v;
        if (m.expr_use_vrot) d.vrot = m.expr_vrot;
      }
      d.center = d.center + (dt * d.vel + 0.5 * dt * dt * m.acc);
    }
  }
}
// push_f_v: cylinder.hpp:149-151 (nothing), surface.hpp:232-252, ball.hpp:219-234, rshape.hpp:262-281
EDB_HD void driver_push_f_v(Driver& d, DriverMotion& m, double dt) {
  const int mt = m.motion;
  if (d.kind == EDB_DRIVER_SURFACE) {
    if (mt == M_STATIONARY) d.vel = mk3(0, 0, 0);
    else if (m_compressive(mt)) { if (m.sigma != 0) d.vel = d.vel + (dt * m.acc.x) * m.motion_vector; }
    else if (mt == M_LINEAR_MOTION) d.vel = m.const_vel * m.motion_vector;
  } else if (d.kind == EDB_DRIVER_BALL) {
    if (m_force_motion(mt)) d.vel = m.acc * dt;
    if (mt == M_LINEAR_MOTION) d.vel = m.motion_vector * m.const_vel;
    if (m_compressive(mt)) {
      if (mt == M_COMPRESSIVE_FORCE) d.vel = mk3(0, 0, 0);
      if (m.sigma != 0) m.rv += dt * m.ra;
    }
  } else if (d.kind == EDB_DRIVER_RSHAPE) {
    if (mt == M_STATIONARY) d.vel = mk3(0, 0, 0);
    else {
      if (m_force_motion(mt)) d.vel = d.vel + m.acc * dt;
      if (mt == M_LINEAR_MOTION) d.vel = m.motion_vector * m.const_vel;
      if (m_compressive(mt) && m.sigma != 0) d.vel = d.vel + 0.5 * dt * m.acc;
    }
  }
}
// push_av_to_quat (RShape only): rshape.hpp:337-377 + driver_numerical_scheme_kernel.hpp:26-56
EDB_HD void driver_push_quat(DriverMotion& m, v3 vrot, double dt) {
  const q4 dq = qdot(m.quat, vrot);
  const q4 Q{m.quat.w + dq.w * dt, m.quat.x + dq.x * dt, m.quat.y + dq.y * dt, m.quat.z + dq.z * dt};
  const double nn = sqrt(Q.w * Q.w + Q.x * Q.x + Q.y * Q.y + Q.z * Q.z);
  m.quat = q4{Q.w / nn, Q.x / nn, Q.y / nn, Q.z / nn};
}
EDB_HD void driver_push_av_to_quat(Driver& d, DriverMotion& m, double time, double dt) {
  if (d.kind != EDB_DRIVER_RSHAPE) return;
  if (rshape_need_moment(m)) {
    if (m.motion == M_EXPRESSION && motion_triggered(m, time) && m.expr_use_mom) {
      m.applied_mom = m.expr_mom;
      const double mn = norm(m.applied_mom);
      m.mom_axis = mn > 1e-12 ? m.applied_mom / mn : mk3(0, 0, 0);
    }
    const v3 pm = m.motion == M_PARTICLE ? m.mom : dot(m.applied_mom + m.mom, m.mom_axis) * m.mom_axis;
    const q4 Qi = qconj(m.quat);
    const v3 o = qrot(Qi, d.vrot), M = qrot(Qi, pm), I = m.inertia;
    const v3 dw{(M.x - (I.z - I.y) * o.y * o.z) / I.x, (M.y - (I.x - I.z) * o.z * o.x) / I.y, (M.z - (I.y - I.x) * o.x * o.y) / I.z};
    const v3 arot = qrot(m.quat, dw);
    d.vrot = d.vrot + arot * (dt * 0.5);
    driver_push_quat(m, d.vrot, dt);
    d.vrot = d.vrot + (dt * 0.5) * arot;
    m.mom = mk3(0, 0, 0);
  } else {
    driver_push_quat(m, d.vrot, dt);
  }
}
// force_to_accel_driver.cpp:30-45 (+ Ball::f_ra ball.hpp:273-283, Surface surface.hpp:208-225, RShape rshape.hpp:220-259)
EDB_HD void driver_force_to_accel(Driver& d, DriverMotion& m, double system_mass) {
  const int mt = m.motion;
  if (d.kind == EDB_DRIVER_BALL) {
    m.pmass = system_mass;
    if (m_compressive(mt) && m.pmass != 0) {
      const double pi = 3.141592653589793;   // 4 * atan(1)
      const double s = 4 * pi * d.radius * d.radius;
      m.ra = (norm(m.forces) - m.sigma * s - (m.damprate * m.rv)) / (m.pmass * 0.5);
    }
    if (m_force_motion(mt)) { update_forces(m, m.forces); m.acc = m.forces / m.mass; }
    else m.acc = mk3(0, 0, 0);
  } else if (d.kind == EDB_DRIVER_SURFACE) {
    m.pmass = system_mass;
    if (m_compressive(mt)) {
      if (m.pmass != 0) {
        const v3 tmp = (m.forces - m.sigma * m.surface * m.motion_vector) / (m.pmass * 0.5);
        m.acc.x = dot(tmp, m.motion_vector);
      } else m.acc.x = 0;
    }
  } else if (d.kind == EDB_DRIVER_RSHAPE) {
    if (m_compressive(mt)) {
      if (m.mass != 0) {
        const double ss = m.sigma * m.surface;   // Vec3d - double: component-wise (rshape.hpp:230)
        const v3 dv = m.damprate * d.vel;
        const v3 num{m.forces.x - ss - dv.x, m.forces.y - ss - dv.y, m.forces.z - ss - dv.z};
        const v3 tmp = num / (m.mass * 0.5);
        m.acc = dot(tmp, m.motion_vector) * m.motion_vector;
      }
    } else if (m_force_motion(mt)) {
      update_forces(m, m.forces);
      m.acc = m.forces / m.mass;
    } else m.acc = mk3(0, 0, 0);
  }
}

// gravity_force.hpp:26-37 (GravityForceDriverFunctor, applied by gravity_force.cu:72-83): gravity acts on the RShape drivers whose motion
// type is PARTICLE (the reference adds it on rank 0 before the all-reduce of the driver forces; here every rank holds the reduced value)
EDB_HD void driver_gravity(const Driver& d, DriverMotion& m, v3 g) {
  if (d.kind == EDB_DRIVER_RSHAPE && m.motion == M_PARTICLE) m.forces = m.forces + g * m.mass;
}
__global__ void k_driver_gravity(const Driver* __restrict__ drv, DriverMotion* __restrict__ mo, int n, v3 g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) driver_gravity(drv[i], mo[i], g);
}

enum DriverOp { DOP_PUSH_F_V_R = 0, DOP_PUSH_F_V = 1, DOP_PUSH_AV_TO_QUAT = 2, DOP_RESET_FORCE = 3, DOP_FORCE_TO_ACCEL = 4 };

// One operator of the time loop over the drivers, one thread per driver.
__global__ void k_driver_op(Driver* __restrict__ drv, DriverMotion* __restrict__ mo, int n, int op, double time, double dt, double system_mass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Driver d = drv[i];
  DriverMotion& m = mo[i];
  if (op == DOP_PUSH_F_V_R) { if (motion_triggered(m, time + dt)) driver_push_f_v_r(d, m, time, dt); }   // push_f_v_r_driver.cpp:33-35
  else if (op == DOP_PUSH_F_V) driver_push_f_v(d, m, dt);
  else if (op == DOP_PUSH_AV_TO_QUAT) driver_push_av_to_quat(d, m, time, dt);
  else if (op == DOP_RESET_FORCE) m.forces = mk3(0, 0, 0);                                                // reset_force_driver.cpp:38-42
  else if (op == DOP_FORCE_TO_ACCEL) driver_force_to_accel(d, m, system_mass);
  drv[i] = d;
}

// Adds the block partials of k_driver_force (one wave) to the drivers that accumulate their reaction: forces when
// need_forces(motion_type) (contact_sphere.hpp:288-290,384-386, contact.hpp:295-297,402-404,422-424), the moment only for spheres
// against an RShape that needs it (contact_sphere.hpp:387-389).  Fixed block order => run-to-run identical sums.
__global__ void k_driver_force_sum(const double* __restrict__ partial, uint32_t nblocks, int ndrv, double* __restrict__ sums /* ndrv*6, accumulated */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ndrv * 6) return;
  double s = 0;
  for (uint32_t b = 0; b < nblocks; b++) s += partial[(size_t)b * ndrv * 6 + i];
  sums[i] += s;
}
// after the (optional) all-reduce of `sums` over the ranks
__global__ void k_driver_force_apply(const double* __restrict__ sums, int ndrv, const Driver* __restrict__ drv, DriverMotion* __restrict__ mo,
                                     int sphere_mode) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= ndrv) return;
  DriverMotion& m = mo[d];
  if (m_need_forces(m.motion)) m.forces = m.forces + mk3(sums[d * 6], sums[d * 6 + 1], sums[d * 6 + 2]);
  if (sphere_mode && drv[d].kind == EDB_DRIVER_RSHAPE && rshape_need_moment(m)) m.mom = m.mom + mk3(sums[d * 6 + 3], sums[d * 6 + 4], sums[d * 6 + 5]);
}

// The driver operators between two contact evaluations, fused for the speculative step loop of edb_run_verlet (one thread per driver):
//   end of step n   : reset_force_driver + the reaction sums of the step's contacts, force_to_accel_driver, push_f_v_driver(dt/2)
//   begin of step n+1: push_f_v_r_driver(t, dt) when is_motion_triggered(t + dt), push_f_v_driver(dt/2), push_av_to_quat_driver(t, dt)
// in the order of config_numerical_schemes.msp:13-15,23-24.  Gated like every state-changing kernel of the loop: a step that turns out to
// need a neighbour rebuild first does nothing and is enqueued again.
struct DriverCycleArgs {
  Driver* drv;
  DriverMotion* mo;
  int n;
  const double* sums;   // n * 6 reaction sums of step n (nullptr: no driver needs them)
  int sphere_mode, do_end, do_begin;
  double dt, system_mass, time_next;
  const unsigned long long* gate;
  v3 g;                 // gravity of the step (gravity_force acts on PARTICLE RShape drivers)
};
EDB_HD void driver_cycle_body(const DriverCycleArgs& A, int i) {
  Driver d = A.drv[i];
  DriverMotion& m = A.mo[i];
  if (A.do_end) {
    m.forces = mk3(0, 0, 0);   // reset_force_driver ran before the forces of the step
    driver_gravity(d, m, A.g);  // gravity_force comes first in compute_force
    if (A.sums) {
      if (m_need_forces(m.motion)) m.forces = m.forces + mk3(A.sums[i * 6], A.sums[i * 6 + 1], A.sums[i * 6 + 2]);
      if (A.sphere_mode && d.kind == EDB_DRIVER_RSHAPE && rshape_need_moment(m)) m.mom = m.mom + mk3(A.sums[i * 6 + 3], A.sums[i * 6 + 4], A.sums[i * 6 + 5]);
    }
    driver_force_to_accel(d, m, A.system_mass);
    driver_push_f_v(d, m, 0.5 * A.dt);
  }
  if (A.do_begin) {
    if (motion_triggered(m, A.time_next + A.dt)) driver_push_f_v_r(d, m, A.time_next, A.dt);
    driver_push_f_v(d, m, 0.5 * A.dt);
    driver_push_av_to_quat(d, m, A.time_next, A.dt);
  }
  A.drv[i] = d;
}
__global__ void k_driver_cycle(DriverCycleArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n || (A.gate && *A.gate)) return;
  driver_cycle_body(A, i);
}

// driver_vertices.cu:28-35: vertices_[i] = center + quat * shp_vertices[i] for the non-stationary meshes (rshape.hpp:419); `all` forces
// every mesh (initialisation, RShapeDriver::initialize rshape.hpp:160-165 uses update_vertex = get_vertex(i, center, 1.0, quat))
__global__ void k_driver_vertices(const MeshDev* __restrict__ md, int nmesh, const Driver* __restrict__ drv, const DriverMotion* __restrict__ mo,
                                  const double4* __restrict__ local, double4* __restrict__ world, int all) {
  const int m = blockIdx.y;
  const MeshDev M = md[m];
  const Driver& D = drv[M.drv];
  const DriverMotion& DM = mo[M.drv];
  if (!all && DM.motion == M_STATIONARY && is_zero(D.vrot)) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M.nv; i += gridDim.x * blockDim.x) {
    const v3 lv = xyz(local[M.v0 + i]);
    const v3 w = D.center + qrot(DM.quat, lv);   // homothety 1.0 * v == v
    world[M.v0 + i] = mk4(w, 0.0);
  }
}

// driver_displ_over.cu:60-243: number of drivers (mesh: vertices) that moved >= sqrt(r2) since the backup.  The reference stops at the
// first driver that moved (:233-236); only "> 0" is consumed, so every driver is counted here.
__global__ void k_driver_displ(const Driver* __restrict__ drv, const DriverMotion* __restrict__ mo, const Driver* __restrict__ bdrv,
                               const DriverMotion* __restrict__ bmo, int n, const MeshDev* __restrict__ md, const double4* __restrict__ local,
                               double r2, unsigned long long* __restrict__ out, const unsigned long long* gate) {
  const int d = blockIdx.y;
  if (d >= n || (gate && *gate)) return;
  const Driver &a = drv[d], &b = bdrv[d];
  const DriverMotion &ma = mo[d], &mb = bmo[d];
  unsigned int cnt = 0;
  if (a.kind == EDB_DRIVER_RSHAPE) {
    const bool same = a.center.x == b.center.x && a.center.y == b.center.y && a.center.z == b.center.z && ma.quat.w == mb.quat.w &&
                      ma.quat.x == mb.quat.x && ma.quat.y == mb.quat.y && ma.quat.z == mb.quat.z;
    if (!same) {
      const MeshDev M = md[a.shape];
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M.nv; i += gridDim.x * blockDim.x) {
        const v3 lv = 1.0 * xyz(local[M.v0 + i]);
        const v3 dd = (a.center + qrot(ma.quat, lv)) - (b.center + qrot(mb.quat, lv));
        if (dot(dd, dd) >= r2) cnt++;
      }
    }
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.kind == EDB_DRIVER_SURFACE) {
      v3 dd = a.center_proj - b.center_proj;
      if (mb.motion == M_PENDULUM_MOTION) {
        // intersection of the line (initial position, swing direction) with the plane (anchor, normal), :75-100
        const v3 da = ma.pend_anchor - ma.pend_init;
        const double ta = dot(da, a.normal) / dot(ma.pend_dir, a.normal), tb = dot(da, b.normal) / dot(ma.pend_dir, b.normal);
        const v3 pa = ma.pend_init + ta * ma.pend_dir, pb = ma.pend_init + tb * ma.pend_dir;
        dd = pb - pa;
      }
      cnt = dot(dd, dd) >= r2 ? 1 : 0;
    } else if (a.kind == EDB_DRIVER_BALL) {
      const v3 dd = a.center - b.center;
      if (m_compressive(ma.motion)) { const double rd = a.radius - b.radius; cnt = rd * rd >= r2 ? 1 : 0; }
      else cnt = dot(dd, dd) >= r2 ? 1 : 0;
    }   // cylinder: 0 ("should not move", :66-68, quirk Q14)
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

// ---- grid_rshape_driver on the device ---------------------------------------------------------------------------------------------
// A primitive (vertex / edge / face of mesh m) is listed in every real cell that its bounding box, enlarged by the Minkowski radius +
// the cell size, overlaps: conservative for any particle of the cell because cell_size >= 2 r_max + rcut_inc (check_rcut.cpp:72).  The
// candidate SET is decided later by the reference's filters (nbh_sphere.hpp:43-80, nbh_polyhedron_driver.hpp:61-201); the lists only
// bound the search like RShapeDriverGridCellIndexes (rshape_grid.hpp:125-149, grid_stl_mesh.cpp:150-251).
// Pass 0 counts into cnt[(m * ncell + cell) * 3 + kind]; pass 1 fills data[off[...] + cursor]; k_mesh_list_sort orders every list by
// primitive index so that the enumeration (and with it the summation order of the forces) does not depend on the atomics.
struct MeshGridArgs {
  const MeshDev* md;
  const double4* verts;
  const int* edges;
  const int* foff;
  const int* fn;
  const int* fidx;
  int nmesh;
  uint32_t nprim_total;        // sum over meshes of nv + ne + nf
  const uint32_t* prim_start;  // nmesh + 1: first global primitive index of each mesh
  GridDev G;
  double rmax;                 // max(cell_size, effective cell sizes)
};
template <int FILL>
__global__ void k_mesh_cells(MeshGridArgs A, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off, int* __restrict__ data) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nprim_total) return;
  int m = 0;
  while (m + 1 < A.nmesh && t >= A.prim_start[m + 1]) m++;
  const MeshDev M = A.md[m];
  int id = (int)(t - A.prim_start[m]), kind = 0;
  if (id >= M.nv) { id -= M.nv; kind = 1; }
  if (kind == 1 && id >= M.ne) { id -= M.ne; kind = 2; }
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  auto grow = [&](int v) {
    const double4 p = A.verts[M.v0 + v];
    lo[0] = fmin(lo[0], p.x); hi[0] = fmax(hi[0], p.x);
    lo[1] = fmin(lo[1], p.y); hi[1] = fmax(hi[1], p.y);
    lo[2] = fmin(lo[2], p.z); hi[2] = fmax(hi[2], p.z);
  };
  if (kind == 0) grow(id);
  else if (kind == 1) { grow(A.edges[2 * (M.e0 + id)]); grow(A.edges[2 * (M.e0 + id) + 1]); }
  else {
    const int n = A.fn[M.f0 + id], o = M.fi0 + A.foff[M.f0 + id];
    for (int k = 0; k < n; k++) grow(A.fidx[o + k]);
  }
  int a[3], b[3];
  const double e = M.radius + A.rmax;
  for (int d = 0; d < 3; d++) {
    a[d] = (int)floor((lo[d] - e - A.G.lo[d]) / A.G.cs[d]) + 1;
    b[d] = (int)floor((hi[d] + e - A.G.lo[d]) / A.G.cs[d]) + 1;
    a[d] = max(a[d], 1);
    b[d] = min(b[d], A.G.nc[d]);
    if (a[d] > b[d]) return;
  }
  for (int z = a[2]; z <= b[2]; z++)
    for (int y = a[1]; y <= b[1]; y++)
      for (int x = a[0]; x <= b[0]; x++) {
        const size_t cell = (size_t)x + (size_t)A.G.g[0] * ((size_t)y + (size_t)A.G.g[1] * (size_t)z);
        const size_t slot = ((size_t)m * A.G.ncell + cell) * 3 + kind;
        const uint32_t k = atomicAdd(cnt + slot, 1u);
        if (FILL) data[off[slot] + k] = id;
      }
}
// one thread per list: insertion sort (lists hold the few primitives near one cell)
__global__ void k_mesh_list_sort(const uint32_t* __restrict__ off, uint32_t nlists, int* __restrict__ data) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlists) return;
  const uint32_t b = off[l], e = off[l + 1];
  for (uint32_t i = b + 1; i < e; i++) {
    const int v = data[i];
    uint32_t j = i;
    while (j > b && data[j - 1] > v) { data[j] = data[j - 1]; j--; }
    data[j] = v;
  }
}

}  // namespace edb
