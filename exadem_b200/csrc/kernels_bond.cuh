// kernels_bond.cuh — inner bonds (fragmentation, wave 13), SURVEY 8(f) rank 4: the per-step operators.
//
// Replaces inner_bond_polyhedron (src/polyhedron/inner_bond_force.cu, include/exaDEM/polyhedron/inner_bond.hpp:106-185, law
// src/forcefield/include/exaDEM/forcefield/inner_bond_force.hpp:26-70) and apply_interface_fracture_criterion
// (src/interaction/include/exaDEM/interface/apply_interface_fracture_criterion.hpp:40-73).  Bonds are a separate SoA wave sorted by
// (owner slot, partner id, sub_i, sub_j) so that the bonds of one interface are contiguous (interface.hpp:28-33); the wrench of a bond is
// stored per bond and summed per particle through an incidence list in a fixed order — no float atomics (the reference uses 12 per bond).
#pragma once
#include "common.cuh"

namespace edb {

struct BondParams {   // InnerBondParams (inner_bond_parameters.hpp:33-45)
  double kn, kt, damp_rate;
  int mode, pad;
  double crit1, crit2;
};
struct BondView {
  uint32_t n;
  const uint2* ij;        // particle slots (0xffffffff: unresolved)
  const uint2* sub;       // glued vertices
  const double4* par;     // dn0, weight, criteria_1, criteria_2
  const uint8_t* mode;
  uint8_t* unbroken;
  double4* fr;            // friction xyz, en
  double4* tds;           // cumulative tangential displacement xyz, et
  double4* fd;            // f_i xyz, dn      (ContactState: fn = f - friction)
  double4* cp;            // contact point xyz
};

// inner_bond_force.hpp:26-70 — same operation order as the reference expression by expression
EDB_HD void inner_bond_force_core(double dn, v3 n, double dn0, double weight, double dt, const BondParams& p, double meff, double& En, v3& tds,
                                  double& Et, v3& ft, v3 cp, v3 pi, v3 vi, v3& fi, v3 wi, v3 pj, v3 vj, v3 wj) {
  const double damp = p.damp_rate * 2.0 * sqrt(p.kn * meff);
  const v3 vel = (vj - cross(cp - pj, wj)) - (vi - cross(cp - pi, wi));
  const double vn = dot(vel, n);
  const double delta = dn - dn0;
  const double kn_w = p.kn * weight, kt_w = p.kt * weight;
  const double fne = -kn_w * delta, fnv = damp * vn;
  const double fn = fne + fnv;
  const v3 vfn = fn * n;
  const v3 ds = dt * (vel - vn * n);
  tds = tds + ds;
  ft = kt_w * tds;
  fi = vfn + ft;
  En = (fne <= 0 ? 1.0 : 0.0) * (0.5 * kn_w * delta * delta);
  Et = 0.5 * kt_w * dot(tds, tds);
}
// apply_interface_fracture_criterion.hpp:40-66
EDB_HD bool interface_breaks(int mode, double c1, double c2, double En, double Et, double S) {
  if (mode == 0) return En + Et > c1;
  if (mode == 1) return En > c1 || Et > c2;
  if (mode == 2) return (En > c1) && (S > c2);
  return false;
}

// one thread per bond: vertex-vertex detection on the world vertices (detection_vertex_vertex, shape_detection_api.hpp) + the law
__global__ void __launch_bounds__(BLK) k_bond_force(BondView B, PView P, const double4* __restrict__ vert, ShapesView S, const BondParams* __restrict__ ibp,
                                                    int n_groups, double dt) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B.n) return;
  const uint2 ij = B.ij[k];
  // a broken bond no longer acts (its conversion into a vertex-vertex contact at the next rebuild is not built yet, see the header)
  if (ij.x == 0xffffffffu || ij.y == 0xffffffffu || !B.unbroken[k]) {
    B.fd[k] = make_double4(0, 0, 0, 0);
    B.cp[k] = make_double4(0, 0, 0, 0);
    return;
  }
  const uint2 sb = B.sub[k];
  const double4 pi = ldg4(P.pr + ij.x), pj = ldg4(P.pr + ij.y), wi = ldg4(P.wg + ij.x), wj = ldg4(P.wg + ij.y);
  const double4 vi = ldg4(P.vm + ij.x), vj = ldg4(P.vm + ij.y);
  const uint32_t ti = (uint32_t)(((unsigned long long)__double_as_longlong(wi.w)) & 0xffffffffu), tj = (uint32_t)(((unsigned long long)__double_as_longlong(wj.w)) & 0xffffffffu);
  const uint32_t gi = (uint32_t)(((unsigned long long)__double_as_longlong(wi.w)) >> 32), gj = (uint32_t)(((unsigned long long)__double_as_longlong(wj.w)) >> 32);
  const double ri = P.h[ij.x] * S.sh[ti].radius, rj = P.h[ij.y] * S.sh[tj].radius;
  const Contact c = detect_vv(xyz(ldg4(vert + (size_t)ij.x * S.nvmax + sb.x)), ri, xyz(ldg4(vert + (size_t)ij.y * S.nvmax + sb.y)), rj);
  const BondParams p = ibp[(size_t)gi * n_groups + gj];
  const double meff = (vi.w * vj.w) / (vi.w + vj.w);   // compute_effective_mass, common_kernels.hpp:93-97
  const double4 par = B.par[k];
  double4 fr = B.fr[k], td = B.tds[k];
  v3 tds = xyz(td), ft = xyz(fr), fi = mk3(0, 0, 0);
  double En, Et;
  inner_bond_force_core(c.dn, c.n, par.x, par.y, dt, p, meff, En, tds, Et, ft, c.cp, xyz(pi), xyz(vi), fi, xyz(wi), xyz(pj), xyz(vj), xyz(wj));
  B.fr[k] = mk4(ft, En);
  B.tds[k] = mk4(tds, Et);
  B.fd[k] = mk4(fi, c.dn);
  B.cp[k] = mk4(c.cp, 0.0);
}

// one thread per particle that carries bonds: f += sum(+-f_i), mom += sum(cross(cp - r, +-f_i)) over its incident bonds in list order
// (inner_bond.hpp:166-178).  inc = bond index << 1 | side (0: particle is i, 1: particle is j).
__global__ void __launch_bounds__(BLK) k_bond_gather(BondView B, PView P, const uint32_t* __restrict__ plist, const uint32_t* __restrict__ inc_off,
                                                     const uint32_t* __restrict__ inc, uint32_t nlist) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nlist) return;
  const uint32_t p = plist[t];
  const v3 r = xyz(ldg4(P.pr + p));
  v3 f = mk3(P.fx[p], P.fy[p], P.fz[p]), m = mk3(P.mx[p], P.my[p], P.mz[p]);
  for (uint32_t q = inc_off[t]; q < inc_off[t + 1]; q++) {
    const uint32_t code = inc[q], k = code >> 1;
    const double4 fd = B.fd[k], cp = B.cp[k];
    const v3 fi = (code & 1) ? -xyz(fd) : xyz(fd);
    m = m + moments(xyz(cp), r, fi, mk3(0, 0, 0));
    f = f + fi;
  }
  P.fx[p] = f.x; P.fy[p] = f.y; P.fz[p] = f.z;
  P.mx[p] = m.x; P.my[p] = m.y; P.mz[p] = m.z;
}

// broken_interface_friction_updater.cpp:58-98: the bonds of a broken interface will become vertex-vertex contacts; their cumulated
// friction is capped by the Coulomb threshold of the CONTACT parameters (mu * |fn|) and zeroed when the two vertices are apart
__global__ void __launch_bounds__(BLK) k_bond_friction_updater(BondView B, const uint2* __restrict__ itf, uint32_t nitf, const uint8_t* __restrict__ broken,
                                                               PView P, ParamsDev prm) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nitf || !broken[t]) return;
  const uint2 I = itf[t];
  const uint2 ij0 = B.ij[I.x];
  if (ij0.x == 0xffffffffu || ij0.y == 0xffffffffu) return;
  const double4 wi = ldg4(P.wg + ij0.x), wj = ldg4(P.wg + ij0.y);
  const uint32_t ga = (uint32_t)(((unsigned long long)__double_as_longlong(wi.w)) >> 32), gb = (uint32_t)(((unsigned long long)__double_as_longlong(wj.w)) >> 32);
  const double mu = (prm.tpp ? prm.tpp[ga * prm.n_groups + gb] : prm.pp).mu;
  for (uint32_t j = I.x; j < I.x + I.y; j++) {
    B.unbroken[j] = 0;
    double4 fr = B.fr[j];
    const double4 fd = B.fd[j];
    const v3 ftv = xyz(fr), fn = xyz(fd) - ftv;
    const double ft = norm(ftv);
    const double thr = mu * norm(fn);
    v3 nf = ftv;
    if (ft > thr && ft > 0) nf = ftv * thr / ft;
    if (fd.w > 0) nf = mk3(0, 0, 0);
    B.fr[j] = mk4(nf, fr.w);
  }
}

// unclassify_interactions.cpp:135-145 + broke_interaction (inner_bond_interaction.hpp:202-210): at the rebuild that follows a break, a
// broken bond with a non-zero friction becomes the history of the vertex-vertex contact record between the same two vertices.
struct BondConv {
  uint32_t owner;              // slot of particle i in the new order
  uint32_t sub_i, sub_j;
  long long pid;               // id of particle j
  double fx, fy, fz;
};
__global__ void k_bond_convert(const BondConv* __restrict__ cv, uint32_t n, WView w) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const BondConv b = cv[t];
  for (uint32_t s = w.off[b.owner]; s < w.off[b.owner + 1]; s++) {
    const double4 ft = w.ft[s];
    if (__double_as_longlong(ft.w) != b.pid || (w.flg[s] & 1)) continue;
    const uint2 sb = w.sub[s];
    if (sb.x != b.sub_i || sb.y != b.sub_j) continue;
    w.ft[s] = make_double4(b.fx, b.fy, b.fz, ft.w);
    w.flg[s] = w.flg[s] | 8;   // FLG_HOT: the record carries history
    return;
  }
}

// one thread per interface (loc, size): sums in bond order, breaks every bond of the interface when the criterion is exceeded
__global__ void __launch_bounds__(BLK) k_interface_criterion(BondView B, const uint2* __restrict__ itf, uint32_t nitf, uint8_t* __restrict__ broken,
                                                             unsigned long long* __restrict__ count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int nb = 0;
  if (t < nitf) {
    const uint2 I = itf[t];
    double En = 0.0, Et = 0.0, Sn = 0.0;
    for (uint32_t j = I.x; j < I.x + I.y; j++) {
      const double4 fr = B.fr[j], td = B.tds[j], fd = B.fd[j];
      En += fr.w;
      Et += td.w;
      if (fd.w > 0.0) Sn += norm(xyz(fd) - xyz(fr));   // fn = f - friction
    }
    Sn /= I.y;
    const double4 par = B.par[I.x];
    if (interface_breaks(B.mode[I.x], par.z, par.w, En, Et, Sn)) {
      broken[t] = 1;
      nb = 1;
      for (uint32_t j = I.x; j < I.x + I.y; j++) B.unbroken[j] = 0;
    }
  }
  nb = __reduce_add_sync(0xffffffffu, nb);
  if ((threadIdx.x & 31) == 0 && nb) atomicAdd(count, (unsigned long long)nb);
}

}  // namespace edb
