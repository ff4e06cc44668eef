// kernels_bond.cuh — inner bonds (fragmentation, wave 13), SURVEY 8(f) rank 4: the per-step operators.
//
// Replaces inner_bond_polyhedron (src/polyhedron/inner_bond_force.cu, include/exaDEM/polyhedron/inner_bond.hpp:106-185, law
// src/forcefield/include/exaDEM/forcefield/inner_bond_force.hpp:26-70) and apply_interface_fracture_criterion
// (src/interaction/include/exaDEM/interface/apply_interface_fracture_criterion.hpp:40-73).  Bonds are a separate SoA wave sorted by
// (owner slot, partner id, sub_i, sub_j) so that the bonds of one interface are contiguous (interface.hpp:28-33); the wrench of a bond is
// stored per bond and summed per particle through an incidence list in a fixed order — no float atomics (the reference uses 12 per bond).
#pragma once
#include "common.cuh"
#include "kernels_rebuild.cuh"

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
                                                    int n_groups, double dt, const unsigned long long* gate) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B.n || (gate && *gate)) return;
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
                                                     const uint32_t* __restrict__ inc, uint32_t nlist, const unsigned long long* gate) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nlist || (gate && *gate)) return;
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

// ---- stick_polyhedra (src/polyhedron/stick_polyhedra.cpp:88-360) on the device ------------------------------------------------------
// One thread per real particle a walks its neighbour cells like the contact enumeration does (same ContactNeighborFilterFunc); for a
// partner b of the same cluster with id_a < id_b it looks for the first pair of faces (i of a, j of b) with the same vertex count, areas
// within 1 %, parallel normals, and glues every vertex of face i to the first vertex of face j closer than the sticking threshold;
// faces with fewer than 3 glued vertices are rejected and the search goes on; the glued vertices of a must be coplanar
// (check_stiked_face, interaction_manager.hpp:226-276).  Pass 0 counts, pass 1 writes at the particle's offset.  Partners in ghost cells
// (periodic images / other bricks) are skipped: OwnerGhost bonds are not built.
constexpr int STICK_MAXV = 16;   // vertices per face handled
struct StickBond {
  uint32_t a, b, sub_i, sub_j;
  double dn0, weight, c1, c2;
  int mode, pad;
};
struct StickArgs {
  NbhArgs N;
  const double* cluster;      // per slot
  const double* farea;        // face areas at homothety 1, pooled like the face tables (ShapeDev::f0)
  const BondParams* ibp;
  int n_groups;
  double dn_crit;
};
template <int FILL>
__global__ void __launch_bounds__(64) k_stick(StickArgs T, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off, StickBond* __restrict__ out) {
  const NbhArgs& A = T.N;
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.nreal) return;
  const double4 ra = A.P.pr[a], wga = A.P.wg[a];
  const uint64_t ida = A.P.id[a];
  const uint32_t ta = (uint32_t)((unsigned long long)__double_as_longlong(wga.w) & 0xffffffffu), ga = (uint32_t)(((unsigned long long)__double_as_longlong(wga.w)) >> 32);
  const ShapeDev sa = A.S.sh[ta];
  const double ha = A.P.h[a], ca = T.cluster[a];
  const VA va{A.vert + (size_t)a * A.S.nvmax};
  uint32_t nout = 0;
  for_each_neighbor(A, a, ra, [&](uint32_t b, bool gc) {
    if (gc || b >= A.nreal) return;
    if (T.cluster[b] != ca) return;
    const uint64_t idb = A.P.id[b];
    if (ida >= idb) return;
    const double4 wgb = A.P.wg[b];
    const uint32_t tb = (uint32_t)((unsigned long long)__double_as_longlong(wgb.w) & 0xffffffffu), gb = (uint32_t)(((unsigned long long)__double_as_longlong(wgb.w)) >> 32);
    const ShapeDev sb = A.S.sh[tb];
    const double hb = A.P.h[b];
    const VA vb{A.vert + (size_t)b * A.S.nvmax};
    const BondParams P = T.ibp[(size_t)ga * T.n_groups + gb];
    StickBond loc[STICK_MAXV];
    int nloc = 0;
    bool found = false;
    for (int i = 0; i < sa.nf && !found; i++) {
      const int ni_ = A.S.fn[sa.f0 + i];
      if (ni_ < 3 || ni_ > STICK_MAXV) continue;
      const int* fi = A.S.fidx + sa.fi0 + A.S.foff[sa.f0 + i];
      v3 ni = cross(va[fi[1]] - va[fi[0]], va[fi[2]] - va[fi[0]]);
      ni = ni * (1.0 / norm(ni));
      const double area_i = T.farea[sa.f0 + i];
      for (int j = 0; j < sb.nf && !found; j++) {
        const int nj_ = A.S.fn[sb.f0 + j];
        if (nj_ < 3 || ni_ != nj_) continue;
        const double ratio = area_i / T.farea[sb.f0 + j];
        if (ratio > 1.01 || ratio < 0.99) continue;
        const int* fj = A.S.fidx + sb.fi0 + A.S.foff[sb.f0 + j];
        v3 nj = cross(vb[fj[1]] - vb[fj[0]], vb[fj[2]] - vb[fj[0]]);
        nj = nj * (1.0 / norm(nj));
        if (norm(cross(ni, nj)) > 1.e-10) continue;
        const double area = area_i * ha * ha;   // shape::get_face_area(i, h), pi.id_ < pj.id_ here
        StickBond B{a, b, 0, 0, 0.0, 1.0, 0.0, 0.0, P.mode, 0};
        if (P.mode == 0) B.c1 = 2 * area * P.crit1;
        else if (P.mode == 1) { B.c1 = 2 * area * P.crit1; B.c2 = 2 * area * P.crit2; }
        else if (P.mode == 2) { B.c1 = 2 * area * P.crit1; B.c2 = area * P.crit2; }
        found = true;
        nloc = 0;
        for (int iv = 0; iv < ni_; iv++)
          for (int jv = 0; jv < nj_; jv++) {
            const Contact c = detect_vv(va[fi[iv]], ha * sa.radius, vb[fj[jv]], hb * sb.radius);
            if (c.dn <= T.dn_crit) {
              B.sub_i = (uint32_t)fi[iv]; B.sub_j = (uint32_t)fj[jv]; B.dn0 = c.dn;
              loc[nloc++] = B;
              break;
            }
          }
        if (nloc <= 2) { nloc = 0; found = false; }
      }
    }
    if (nloc == 0) return;
    bool ok = true;
    if (nloc >= 3) {
      const v3 p0 = va[loc[0].sub_i], nrm = cross(va[loc[1].sub_i] - p0, va[loc[2].sub_i] - p0);
      for (int k = 3; k < nloc; k++)
        if (fabs(dot(va[loc[k].sub_i] - p0, nrm)) > 1.e-10) ok = false;
    }
    if (!ok) return;
    if (FILL) {
      const double w = 1. / static_cast<double>(nloc);
      for (int k = 0; k < nloc; k++) { loc[k].weight = w; out[off[a] + nout + k] = loc[k]; }
    }
    nout += (uint32_t)nloc;
  });
  if (!FILL) cnt[a] = nout;
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
