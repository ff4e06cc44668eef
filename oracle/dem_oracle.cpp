// oracle/dem_oracle.cpp -> oracle/liboracle_dem.so — CPU restatement of the exaDEM contact hot path.
//
// TEST INFRASTRUCTURE ONLY.  This is the ORACLE (the checker): only tests/, __graft_entry__.smoke() and the
// cpu_baseline / `--impl reference` legs of bench.py may load it.  The product (exadem_b200/) never does.
//
// Parity status: the arithmetic layer (dem_math.hpp) is pinned bit-for-bit against the reference's own headers
// compiled in place (oracle/_ref, tests/test_oracle_vs_ref.py).  The operator loops below depend on exaNBody /
// Onika code that is NOT in /root/reference (cell grid, neighbour visit, PushVec3 functors, quaternion algebra):
// they are restated from the call-site contract and the reference ships no golden vectors for
// them  =>  "parity unpinned" for that layer (risks R1-R3).
//
// Reference loops restated here (paths relative to /root/reference):
//   neighbour filter              src/grid/chunk_neighbors_contact.cpp:46-57
//   nbh_sphere                    src/sphere/nbh_sphere.cu:64-309
//   nbh_polyhedron                src/polyhedron/nbh_polyhedron.cpp:75-384, include/exaDEM/polyhedron/nbh_polyhedron_driver.hpp:205-238
//   history carry-over            src/interaction/include/exaDEM/interaction/interaction_manager.hpp:88-112,205-217,
//                                 placeholder_interaction.hpp:398-408, interaction_pair.hpp:204-249, interaction.hpp:129-136
//   classify (wave order)         src/interaction/classify_interactions.cpp:52-136
//   wave loop                     src/interaction/include/exaDEM/classifier/classifier_for_all.hpp:116-156
//   sphere kernels                src/sphere/include/exaDEM/sphere/contact_sphere.hpp:121-297
//   polyhedron kernels            src/polyhedron/include/exaDEM/polyhedron/contact.hpp:124-304
//   compute_vertices              src/polyhedron/include/exaDEM/polyhedron/compute_vertices.hpp:26-67, shape.hpp:259-262
//   integrators                   src/numericalscheme/combined_compute_{prolog,epilog}.cu, include/exaDEM/{push_to_quaternion,
//                                 angular_acceleration,angular_velocity,force_to_accel}.hpp, forcefield/gravity_force.hpp:38-46
//   ghost refresh field list      src/grid/update_ghosts.cu:46-58
#include "dem_math.hpp"
#include "../include/exadem_b200.h"

#include <unordered_map>
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <tuple>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

struct Shape {
  std::vector<v3> v;
  std::vector<int> edges;               // 2*ne
  std::vector<int> fidx;                // concatenated face vertex indices
  std::vector<int> foff, fn;            // per face offset into fidx, vertex count
  Obb obb;
  double radius = 0;
  int nv() const { return (int)v.size(); }
  int ne() const { return (int)edges.size() / 2; }
  int nf() const { return (int)fn.size(); }
};

struct Inter {
  uint32_t i = 0, j = 0;      // particle indices (j unused for drivers)
  uint64_t id_i = 0, id_j = 0;
  uint32_t sub_i = 0, sub_j = 0;
  uint16_t type = 0;
  uint8_t swap = 0, ghost = 0;
  uint8_t foreign = 0;        // virtual bricks: the partner is a REAL slot that another rank would own (force on it discarded)
  uint8_t drop_owner = 0;     // PartnerGhost twin (update_interaction_ghost.cpp): the owner is a ghost copy, only the partner's share counts
  v3 ft{0, 0, 0}, mom{0, 0, 0};
  // contact state (classifier/contact_state.hpp)
  double dn = 0;
  v3 cp{0, 0, 0}, fn{0, 0, 0};
  uint32_t owner() const { return swap ? j : i; }
};

// ordering key of InteractionPair::operator< (interaction_pair.hpp:222-249): owner.id, partner.id, type, swap, owner.sub, partner.sub
using Key = std::tuple<uint64_t, uint64_t, uint16_t, uint8_t, uint32_t, uint32_t>;
static Key key_of(const Inter& I) {
  if (!I.swap) return Key{I.id_i, I.id_j, I.type, I.swap, I.sub_i, I.sub_j};
  return Key{I.id_j, I.id_i, I.type, I.swap, I.sub_j, I.sub_i};
}

struct Sim {
  edb_config cfg{};
  int law = 0;
  bool multimat = false;
  int n_groups = 1;
  std::vector<Params> pp{Params()}, pd{Params()};
  std::vector<Driver> drivers;
  std::vector<Shape> shapes;
  int nvmax = 0;
  // inner bonds (wave 13, InnerBondInteraction inner_bond_interaction.hpp:49-66) + interfaces (interface.hpp:28-33)
  struct Bond {
    uint64_t id_i = 0, id_j = 0;
    uint32_t sub_i = 0, sub_j = 0, i = 0, j = 0;
    v3 friction{0, 0, 0}, tds{0, 0, 0};
    double en = 0, et = 0, dn0 = 0, weight = 1, c1 = 0, c2 = 0;
    int mode = 3;
    uint8_t unbroken = 1, ghost = 0;
    double dn = 0;                 // ContactState of the wave
    v3 cp{0, 0, 0}, fn{0, 0, 0};
  };
  std::vector<Bond> bonds;
  std::set<std::pair<uint64_t, uint64_t>> glued;   // (id_i, id_j) of the unbroken bonds at the last rebuild (InteractionManager::ignore_)
  std::vector<BondParams> ibp{BondParams()};
  int ibp_groups = 1;
  std::vector<double> cluster;                  // field::cluster per particle (index = particle index)
  std::vector<std::pair<size_t, size_t>> interfaces;   // (loc, size)
  std::vector<uint8_t> interface_broken;
  std::vector<Driver> drivers_backup;           // backup_drvs slot (backup_drivers.cu)
  double physical_time = 0, system_mass = 0;
  bool have_system_mass = false;
  std::vector<Shape> mesh_shapes;               // RShapeDriver::shp_ (rshape.hpp)
  std::vector<std::vector<v3>> mesh_verts;      // per driver: RShapeDriver::vertices_ = position + orientation * v (rshape.hpp:381-384)

  // particles: [0,nreal) real, [nreal,ntot) ghosts
  size_t nreal = 0, ntot = 0;
  std::vector<v3> r, v, f, mom, w, arot, inertia;
  std::vector<q4> q;
  std::vector<double> mass, h, radius;
  std::vector<uint64_t> id;
  std::vector<edb_region> regions;   // particle regions of the `region` slots
  std::vector<uint32_t> type, group;
  std::vector<uint32_t> gsrc;
  std::vector<v3> gshift;
  std::vector<v3> r_backup;
  std::vector<q4> q_backup;
  std::vector<v3> verts;  // ntot*nvmax world vertices

  // grid (with 1 ghost layer)
  int nc[3]{1, 1, 1}, g[3]{3, 3, 3};
  double cs[3]{1, 1, 1};  // effective cell size per axis: (hi-lo)/nc >= cfg.cell_size
  std::vector<uint32_t> cstart, citems;

  std::vector<Inter> inter;                 // "ges": owner-major, enumeration order
  std::vector<uint32_t> wave[EDB_NTYPES];   // "ic": indices into inter per type, classify order
  double rcut_inc = 0;
  int symetric = 1;
  std::string err;
  int omp_threads = 1;  // 1 = canonical serial accumulation order
  // Virtual bricks (multi-rank emulation, used only by the multi-GPU parity worker): the single domain is cut into
  // vb[0] x vb[1] x vb[2] bricks of whole cells and a neighbour cell of ANOTHER brick is treated exactly like the ghost
  // cell it would be on the owning MPI rank of the reference (pair enumerated from both sides, partner force discarded).
  int vb[3]{1, 1, 1};
  std::vector<int> vcut[3];   // non-uniform virtual bricks: vb[d] + 1 cell indices (0 .. nc[d]) per axis, empty = equal parts
  int brick_of_cell(int i, int j, int k) const {
    const int c[3] = {i, j, k};
    int r = 0, mult = 1;
    for (int d = 0; d < 3; d++) {
      int b;
      if (!vcut[d].empty()) {
        b = 0;
        for (int q = 1; q < vb[d]; q++) b += (c[d] - 1) >= vcut[d][q] ? 1 : 0;
      } else {
        b = (c[d] - 1) / (nc[d] / vb[d]);
      }
      b = b < 0 ? 0 : (b > vb[d] - 1 ? vb[d] - 1 : b);
      r += mult * b;
      mult *= vb[d];
    }
    return r;
  }
  uint64_t interaction_updates = 0, particle_steps = 0, rebuilds = 0;

  void resize_particles(size_t n) {
    r.resize(n); v.resize(n); f.resize(n); mom.resize(n); w.resize(n); arot.resize(n); inertia.resize(n);
    q.resize(n); mass.resize(n); h.resize(n); radius.resize(n); id.resize(n); type.resize(n); group.resize(n);
    ntot = n;
  }
  const Params& cp_pp(uint32_t ga, uint32_t gb) const { return multimat ? pp[ga * n_groups + gb] : pp[0]; }
  const Params& cp_pd(uint32_t ga, uint32_t d) const { return multimat ? pd[ga * drivers.size() + d] : pd[0]; }
  size_t cell_of(int i, int j, int k) const { return (size_t)i + (size_t)g[0] * ((size_t)j + (size_t)g[1] * (size_t)k); }
  bool is_ghost_cell(int i, int j, int k) const {
    return i == 0 || j == 0 || k == 0 || i == g[0] - 1 || j == g[1] - 1 || k == g[2] - 1;
  }
  const v3* pv(uint32_t p) const { return &verts[(size_t)p * nvmax]; }
};

static inline double L(const Sim& s, int d) { return s.cfg.ghi[d] - s.cfg.glo[d]; }

// --- grid + ghosts -----------------------------------------------------------------------------------------
static void wrap_positions(Sim& s) {
  for (size_t p = 0; p < s.nreal; p++) {
    double* c = &s.r[p].x;
    for (int d = 0; d < 3; d++) {
      if (!s.cfg.periodic[d]) continue;
      if (c[d] >= s.cfg.ghi[d]) c[d] -= L(s, d);
      else if (c[d] < s.cfg.glo[d]) c[d] += L(s, d);
    }
  }
}
static inline int cell_coord(const Sim& s, double x, int d) {
  int c = (int)std::floor((x - s.cfg.lo[d]) / s.cs[d]);
  if (c < 0) c = 0;
  if (c > s.nc[d] - 1) c = s.nc[d] - 1;
  return c;
}
static void build_grid(Sim& s) {
  for (int d = 0; d < 3; d++) {
    // the grid covers the brick exactly: nc cells of size (hi-lo)/nc >= cell_size (exaNBody: grid_dims*cell_size = bounds)
    s.nc[d] = std::max(1, (int)std::floor((s.cfg.hi[d] - s.cfg.lo[d]) / s.cfg.cell_size + 1e-9));
    s.cs[d] = (s.cfg.hi[d] - s.cfg.lo[d]) / s.nc[d];
    s.g[d] = s.nc[d] + 2;
  }
  for (int d = 0; d < 3; d++) {
    if (!s.vcut[d].empty()) {
      if ((int)s.vcut[d].size() != s.vb[d] + 1 || s.vcut[d].front() != 0 || s.vcut[d].back() != s.nc[d]) { s.err = "virtual brick cuts do not span the grid"; s.vcut[d].clear(); s.vb[d] = 1; }
      continue;
    }
    if (s.nc[d] % s.vb[d]) { s.err = "virtual bricks must hold a whole number of cells"; s.vb[d] = 1; }
  }
  s.resize_particles(s.nreal);
  wrap_positions(s);
  const size_t ncell = (size_t)s.g[0] * s.g[1] * s.g[2];
  std::vector<std::vector<uint32_t>> cells(ncell);
  for (size_t p = 0; p < s.nreal; p++) {
    const int ci = cell_coord(s, s.r[p].x, 0) + 1, cj = cell_coord(s, s.r[p].y, 1) + 1, ck = cell_coord(s, s.r[p].z, 2) + 1;
    cells[s.cell_of(ci, cj, ck)].push_back((uint32_t)p);
  }
  // deterministic order inside a cell: by particle id
  for (auto& c : cells) std::sort(c.begin(), c.end(), [&](uint32_t a, uint32_t b) { return s.id[a] < s.id[b]; });
  // ghosts: every ghost-shell cell mirrors the periodic source cell (whole cell is copied, exaNBody ghost layer)
  s.gsrc.clear();
  s.gshift.clear();
  std::vector<std::pair<size_t, uint32_t>> gcell;  // (ghost cell, ghost particle index)
  for (int k = 0; k < s.g[2]; k++)
    for (int j = 0; j < s.g[1]; j++)
      for (int i = 0; i < s.g[0]; i++) {
        if (!s.is_ghost_cell(i, j, k)) continue;
        int src[3] = {i, j, k};
        double sh[3] = {0, 0, 0};
        bool ok = true;
        for (int d = 0; d < 3; d++) {
          if (src[d] == 0) {
            if (!s.cfg.periodic[d]) { ok = false; break; }
            src[d] = s.nc[d];
            sh[d] = -L(s, d);
          } else if (src[d] == s.g[d] - 1) {
            if (!s.cfg.periodic[d]) { ok = false; break; }
            src[d] = 1;
            sh[d] = L(s, d);
          }
        }
        if (!ok) continue;
        const size_t cs = s.cell_of(src[0], src[1], src[2]), cg = s.cell_of(i, j, k);
        for (uint32_t p : cells[cs]) {
          if (p >= s.nreal) continue;
          gcell.push_back({cg, (uint32_t)(s.nreal + s.gsrc.size())});
          s.gsrc.push_back(p);
          s.gshift.push_back({sh[0], sh[1], sh[2]});
        }
      }
  s.resize_particles(s.nreal + s.gsrc.size());
  for (auto& gc : gcell) cells[gc.first].push_back(gc.second);
  s.cstart.assign(ncell + 1, 0);
  s.citems.clear();
  for (size_t c = 0; c < ncell; c++) {
    s.cstart[c] = (uint32_t)s.citems.size();
    s.citems.insert(s.citems.end(), cells[c].begin(), cells[c].end());
  }
  s.cstart[ncell] = (uint32_t)s.citems.size();
}
// ghost_update_rq field list (update_ghosts.cu:46-51) + static fields on (re)build
static void ghost_refresh(Sim& s, bool all_fields) {
  for (size_t gI = 0; gI < s.gsrc.size(); gI++) {
    const size_t p = s.nreal + gI, src = s.gsrc[gI];
    s.r[p] = s.r[src] + s.gshift[gI];
    s.v[p] = s.v[src]; s.w[p] = s.w[src]; s.q[p] = s.q[src];
    s.type[p] = s.type[src]; s.radius[p] = s.radius[src]; s.mass[p] = s.mass[src];
    if (all_fields) {
      s.h[p] = s.h[src]; s.id[p] = s.id[src]; s.group[p] = s.group[src]; s.inertia[p] = s.inertia[src];
      s.arot[p] = s.arot[src]; s.f[p] = {0, 0, 0}; s.mom[p] = {0, 0, 0};
    }
  }
}

static void compute_vertices(Sim& s) {
  if (s.cfg.mode != EDB_MODE_POLYHEDRON) return;
  s.verts.resize(s.ntot * (size_t)s.nvmax);
  for (size_t p = 0; p < s.ntot; p++) {
    const Shape& sh = s.shapes[s.type[p]];
    for (int k = 0; k < sh.nv(); k++) s.verts[p * s.nvmax + k] = s.r[p] + qrot(s.q[p], s.h[p] * sh.v[k]);  // shape.hpp:259-262
  }
}

// --- candidate enumeration -----------------------------------------------------------------------------------
template <class F>
static void for_each_neighbor(const Sim& s, int ci, int cj, int ck, uint32_t a, F&& fn) {
  const double rcut = 0;  // nbh_dist: implied by the radius filter under check_rcut.cpp constraints
  (void)rcut;
  for (int dk = -1; dk <= 1; dk++)
    for (int dj = -1; dj <= 1; dj++)
      for (int di = -1; di <= 1; di++) {
        const int i = ci + di, j = cj + dj, k = ck + dk;
        if (i < 0 || j < 0 || k < 0 || i >= s.g[0] || j >= s.g[1] || k >= s.g[2]) continue;
        const size_t cb = s.cell_of(i, j, k);
        const bool ghostc = s.is_ghost_cell(i, j, k) || s.brick_of_cell(i, j, k) != s.brick_of_cell(ci, cj, ck);
        for (uint32_t it = s.cstart[cb]; it < s.cstart[cb + 1]; it++) {
          const uint32_t b = s.citems[it];
          if (b == a) continue;
          // ContactNeighborFilterFunc (chunk_neighbors_contact.cpp:46-57)
          const v3 d = s.r[b] - s.r[a];
          const double d2 = dot(d, d);
          const double rmax = s.radius[a] + s.radius[b] + s.rcut_inc;
          if (d2 > 0.0 && d2 < rmax * rmax) fn(b, ghostc);
        }
      }
}

static void enumerate_sphere(Sim& s, uint32_t a, int ci, int cj, int ck, std::vector<Inter>& out) {
  Inter it;
  it.i = a; it.id_i = s.id[a]; it.sub_i = 0; it.sub_j = 0; it.swap = 0; it.ghost = 0;
  // drivers first (nbh_sphere.cu:182-231)
  for (size_t d = 0; d < s.drivers.size(); d++) {
    const Driver& D = s.drivers[d];
    if (D.kind == DRV_UNDEFINED || D.kind == DRV_RSHAPE) continue;
    it.type = (uint16_t)(4 + D.kind);
    it.id_j = d; it.j = 0xffffffffu;
    if (driver_filter(D, s.radius[a] + s.rcut_inc, s.r[a])) out.push_back(it);
  }
  // RShape (mesh) drivers: detection_sphere_driver (src/sphere/include/exaDEM/sphere/nbh_sphere.hpp:25-80).  The reference walks
  // the per-cell index grid of the mesh (rshape_grid.hpp:125-149), a conservative accelerator: the set is defined by the filters.
  for (size_t d = 0; d < s.drivers.size(); d++) {
    const Driver& D = s.drivers[d];
    if (D.kind != DRV_RSHAPE) continue;
    const Shape& sd = s.mesh_shapes[D.shape];
    const v3* mv = s.mesh_verts[d].data();
    const double rd = sd.radius;  // minkowski(1.0)
    it.id_j = d; it.j = 0xffffffffu; it.sub_i = 0;
    it.type = 7;
    for (int j = 0; j < sd.nv(); j++) if (filter_vv(s.rcut_inc, s.r[a], s.radius[a], mv[j], rd)) { it.sub_j = j; out.push_back(it); }
    it.type = 8;
    for (int j = 0; j < sd.ne(); j++) if (filter_ve(s.rcut_inc, s.r[a], s.radius[a], mv[sd.edges[2 * j]], mv[sd.edges[2 * j + 1]], rd)) { it.sub_j = j; out.push_back(it); }
    it.type = 9;
    for (int j = 0; j < sd.nf(); j++) if (filter_vf(s.rcut_inc, s.r[a], s.radius[a], mv, rd, &sd.fidx[sd.foff[j]], sd.fn[j])) { it.sub_j = j; out.push_back(it); }
  }
  it.sub_i = 0; it.sub_j = 0;
  it.type = 0;
  for_each_neighbor(s, ci, cj, ck, a, [&](uint32_t b, bool ghostc) {
    Inter n = it;
    n.ghost = 0;
    if (s.symetric) {  // nbh_sphere.cu:239-263
      if (s.id[a] >= s.id[b] && !ghostc) return;
    } else {           // :266-294
      if (s.id[a] >= s.id[b]) return;
      if (ghostc) n.ghost = 1;
    }
    n.j = b; n.id_j = s.id[b];
    n.foreign = (ghostc && b < s.nreal) ? 1 : 0;
    out.push_back(n);
  });
}

static void enumerate_polyhedron(Sim& s, uint32_t a, int ci, int cj, int ck, std::vector<Inter>& out) {
  const Shape& sa = s.shapes[s.type[a]];
  const double ha = s.h[a], rV = s.rcut_inc;
  const v3* va = s.pv(a);
  Inter it;
  it.i = a; it.id_i = s.id[a]; it.swap = 0; it.ghost = 0;
  // drivers (nbh_polyhedron_driver.hpp:205-238): one candidate per vertex passing driver.filter(rVerlet + minkowski, vertex)
  for (size_t d = 0; d < s.drivers.size(); d++) {
    const Driver& D = s.drivers[d];
    if (D.kind == DRV_UNDEFINED || D.kind == DRV_RSHAPE) continue;
    it.type = (uint16_t)(4 + D.kind);
    it.id_j = d; it.j = 0xffffffffu; it.sub_j = (uint32_t)-1;
    for (int k = 0; k < sa.nv(); k++)
      if (driver_filter(D, rV + ha * sa.radius, va[k])) { it.sub_i = k; out.push_back(it); }
  }
  // RShape (mesh) drivers: add_driver_interaction (nbh_polyhedron_driver.hpp:61-201), types 7..12.  The OBB pre-culls of the
  // reference (:103-110, :134-139, :163-169) are conservative accelerators; the set is defined by the filters.
  for (size_t d = 0; d < s.drivers.size(); d++) {
    const Driver& D = s.drivers[d];
    if (D.kind != DRV_RSHAPE) continue;
    const Shape& sd = s.mesh_shapes[D.shape];
    const v3* mv = s.mesh_verts[d].data();
    const double rd = sd.radius, ra = ha * sa.radius;
    it.id_j = d; it.j = 0xffffffffu;
    for (int i = 0; i < sa.nv(); i++) {
      it.sub_i = i;
      it.type = 7;
      for (int j = 0; j < sd.nv(); j++) if (filter_vv(rV, va[i], ra, mv[j], rd)) { it.sub_j = j; out.push_back(it); }
      it.type = 8;
      for (int j = 0; j < sd.ne(); j++) if (filter_ve(rV, va[i], ra, mv[sd.edges[2 * j]], mv[sd.edges[2 * j + 1]], rd)) { it.sub_j = j; out.push_back(it); }
      it.type = 9;
      for (int j = 0; j < sd.nf(); j++) if (filter_vf(rV, va[i], ra, mv, rd, &sd.fidx[sd.foff[j]], sd.fn[j])) { it.sub_j = j; out.push_back(it); }
    }
    it.type = 10;
    for (int i = 0; i < sa.ne(); i++)
      for (int j = 0; j < sd.ne(); j++)
        if (filter_ee(rV, va[sa.edges[2 * i]], va[sa.edges[2 * i + 1]], ra, mv[sd.edges[2 * j]], mv[sd.edges[2 * j + 1]], rd)) { it.sub_i = i; it.sub_j = j; out.push_back(it); }
    for (int j = 0; j < sd.nv(); j++) {
      it.sub_j = j;
      it.type = 11;   // mesh vertex vs particle edge
      for (int i = 0; i < sa.ne(); i++) if (filter_ve(rV, mv[j], rd, va[sa.edges[2 * i]], va[sa.edges[2 * i + 1]], ra)) { it.sub_i = i; out.push_back(it); }
      it.type = 12;   // mesh vertex vs particle face
      for (int i = 0; i < sa.nf(); i++) if (filter_vf(rV, mv[j], rd, va, ra, &sa.fidx[sa.foff[i]], sa.fn[i])) { it.sub_i = i; out.push_back(it); }
    }
  }
  for_each_neighbor(s, ci, cj, ck, a, [&](uint32_t b, bool ghostc) {
    if (s.id[a] >= s.id[b]) return;  // nbh_polyhedron.cpp:234-236
    // InteractionManager::skip_ignored_interactions (interaction_manager.hpp:114-160): no regular contact with a partner glued by an
    // unbroken inner bond
    if (!s.glued.empty() && s.glued.count({s.id[a], s.id[b]})) return;
    Inter n;
    n.ghost = ghostc ? 1 : 0;        // :237-239 OwnerGhost
    n.foreign = (ghostc && b < s.nreal) ? 1 : 0;
    const Shape& sb = s.shapes[s.type[b]];
    const double hb = s.h[b];
    const v3* vb = s.pv(b);
    Obb oi = obb_world(sa.obb, s.r[a], s.q[a], ha), oj = obb_world(sb.obb, s.r[b], s.q[b], hb);
    obb_enlarge(oi, 0.5 * rV); obb_enlarge(oj, 0.5 * rV);   // :268-269
    if (!obb_intersect(oi, oj)) return;
    obb_enlarge(oi, 0.5 * rV); obb_enlarge(oj, 0.5 * rV);   // :275-276
    const double ra = ha * sa.radius, rb = hb * sb.radius;
    // A -> B, swap = false
    n.swap = 0; n.i = a; n.id_i = s.id[a]; n.j = b; n.id_j = s.id[b];
    for (int i = 0; i < sa.nv(); i++) {
      if (!obb_intersect(oj, obb_point(va[i], ra))) continue;   // :305-310
      n.sub_i = i;
      n.type = 0;
      for (int j = 0; j < sb.nv(); j++) if (filter_vv(rV, va[i], ra, vb[j], rb)) { n.sub_j = j; out.push_back(n); }
      n.type = 1;
      for (int j = 0; j < sb.ne(); j++) if (filter_ve(rV, va[i], ra, vb[sb.edges[2 * j]], vb[sb.edges[2 * j + 1]], rb)) { n.sub_j = j; out.push_back(n); }
      n.type = 2;
      for (int j = 0; j < sb.nf(); j++) if (filter_vf(rV, va[i], ra, vb, rb, &sb.fidx[sb.foff[j]], sb.fn[j])) { n.sub_j = j; out.push_back(n); }
    }
    n.type = 3;
    for (int i = 0; i < sa.ne(); i++)
      for (int j = 0; j < sb.ne(); j++)
        if (filter_ee(rV, va[sa.edges[2 * i]], va[sa.edges[2 * i + 1]], ra, vb[sb.edges[2 * j]], vb[sb.edges[2 * j + 1]], rb)) {
          n.sub_i = i; n.sub_j = j; out.push_back(n);
        }
    // B -> A, swap = true: pi = b, pj = a (owner() stays a, interaction_pair.hpp:75-90)
    n.swap = 1; n.i = b; n.id_i = s.id[b]; n.j = a; n.id_j = s.id[a];
    for (int j = 0; j < sb.nv(); j++) {
      if (!obb_intersect(oi, obb_point(vb[j], rb))) continue;
      n.sub_i = j;
      n.type = 1;
      for (int i = 0; i < sa.ne(); i++) if (filter_ve(rV, vb[j], rb, va[sa.edges[2 * i]], va[sa.edges[2 * i + 1]], ra)) { n.sub_j = i; out.push_back(n); }
      n.type = 2;
      for (int i = 0; i < sa.nf(); i++) if (filter_vf(rV, vb[j], rb, va, ra, &sa.fidx[sa.foff[i]], sa.fn[i])) { n.sub_j = i; out.push_back(n); }
    }
  });
}

static void classify(Sim& s) {
  for (auto& w : s.wave) w.clear();
  for (size_t k = 0; k < s.inter.size(); k++) s.wave[s.inter[k].type].push_back((uint32_t)k);
}

static void resolve_bonds(Sim& s) {   // persistent interactions follow their particles by id (update_persistent_interactions.cpp)
  std::unordered_map<uint64_t, uint32_t> slot;
  for (uint32_t p = 0; p < s.nreal; p++) slot[s.id[p]] = p;
  for (auto& b : s.bonds) {
    auto a = slot.find(b.id_i), c = slot.find(b.id_j);
    b.i = a == slot.end() ? 0xffffffffu : a->second;
    b.j = c == slot.end() ? 0xffffffffu : c->second;
  }
  std::stable_sort(s.bonds.begin(), s.bonds.end(), [](const Sim::Bond& x, const Sim::Bond& y) {
    return std::tie(x.i, x.id_j, x.sub_i, x.sub_j) < std::tie(y.i, y.id_j, y.sub_i, y.sub_j);
  });
}
static void rebuild(Sim& s, double rcut_inc, int symetric) {
  s.rcut_inc = rcut_inc;
  s.symetric = symetric;
  // history = active records (interaction.hpp:129-136), PartnerGhost excluded; first record wins on duplicate keys
  std::map<Key, std::pair<v3, v3>> hist;
  for (const Inter& I : s.inter) {
    if (I.ghost == 2) continue;
    if (is_zero(I.ft) && is_zero(I.mom)) continue;
    hist.emplace(key_of(I), std::make_pair(I.ft, I.mom));
  }
  build_grid(s);
  ghost_refresh(s, true);
  compute_vertices(s);
  if (!s.bonds.empty()) {   // persistent interactions follow their particles (nbh_polyhedron.cpp:146-149); glued pairs are ignored below
    // broken bonds are no longer persistent: unclassify turns them into vertex-vertex interactions (unclassify_interactions.cpp:135-145,
    // broke_interaction inner_bond_interaction.hpp:202-210) whose friction, if any, is the history of that record in this rebuild
    std::vector<Sim::Bond> keep;
    for (const auto& B : s.bonds) {
      if (B.unbroken) { keep.push_back(B); continue; }
      if (!is_zero(B.friction)) hist.emplace(Key{B.id_i, B.id_j, 0, 0, B.sub_i, B.sub_j}, std::make_pair(B.friction, v3{0, 0, 0}));
    }
    s.bonds.swap(keep);
    resolve_bonds(s);
    s.glued.clear();
    for (const auto& B : s.bonds)
      if (B.unbroken) s.glued.insert({B.id_i, B.id_j});
  }
  s.r_backup.assign(s.r.begin(), s.r.begin() + s.nreal);
  s.q_backup.assign(s.q.begin(), s.q.begin() + s.nreal);
  s.inter.clear();
  for (int k = 1; k < s.g[2] - 1; k++)
    for (int j = 1; j < s.g[1] - 1; j++)
      for (int i = 1; i < s.g[0] - 1; i++) {
        const size_t c = s.cell_of(i, j, k);
        for (uint32_t itp = s.cstart[c]; itp < s.cstart[c + 1]; itp++) {
          const uint32_t a = s.citems[itp];
          if (s.cfg.mode == EDB_MODE_SPHERE) enumerate_sphere(s, a, i, j, k, s.inter);
          else enumerate_polyhedron(s, a, i, j, k, s.inter);
        }
      }
  for (Inter& I : s.inter) {
    auto h = hist.find(key_of(I));
    if (h != hist.end()) { I.ft = h->second.first; I.mom = h->second.second; }
  }
  // update_interaction_ghost (src/interaction/update_interaction_ghost.cpp:58-285): every OwnerGhost record is shipped to the
  // sub-domain (MPI rank, or periodic image) where its partner is a real particle; there the owner is re-located in the ghost
  // cell (:206-231), the partner by id among the neighbouring NON-ghost cells (:242-266), and the record is stored, payload
  // included, as PartnerGhost.  It is then evaluated like any other record (classify_interactions walks traversal_all), so the
  // partner receives its share of the wrench from its own rank's copy.
  {
    const size_t n0 = s.inter.size();
    std::multimap<uint32_t, uint32_t> images;   // real slot -> its ghost copies
    for (size_t q = 0; q < s.gsrc.size(); q++) images.emplace(s.gsrc[q], (uint32_t)q);
    for (size_t k = 0; k < n0; k++) {
      if (s.inter[k].ghost != 1) continue;
      Inter T = s.inter[k];
      const uint32_t a = T.swap ? T.j : T.i, b = T.swap ? T.i : T.j;   // owner, partner
      uint32_t g = a, bp = b;
      if (b >= s.nreal) {   // periodic image: the twin lives where the partner is real, its owner is the opposite image of a
        bp = s.gsrc[b - s.nreal];
        const v3 sh = s.gshift[b - s.nreal];
        bool found = false;
        auto rng = images.equal_range(a);
        for (auto it = rng.first; it != rng.second && !found; ++it) {
          const uint32_t q = it->second;
          if (s.gshift[q].x == -sh.x && s.gshift[q].y == -sh.y && s.gshift[q].z == -sh.z) { g = (uint32_t)(s.nreal + q); found = true; }
        }
        if (!found) continue;   // "can append cause not all particles are copied in ghost cells" (:229-231)
      }
      if (T.swap) { T.j = g; T.i = bp; } else { T.i = g; T.j = bp; }
      T.ghost = 2; T.drop_owner = 1; T.foreign = 0;
      s.inter.push_back(T);
    }
  }
  classify(s);
  s.rebuilds++;
}

// --- contact waves ---------------------------------------------------------------------------------------
struct Acc {  // force / torque accumulation target
  Sim& s;
  bool atomic;
  void add(v3& dst, v3 a) {
    if (!atomic) { dst.x += a.x; dst.y += a.y; dst.z += a.z; return; }
#pragma omp atomic
    dst.x += a.x;
#pragma omp atomic
    dst.y += a.y;
#pragma omp atomic
    dst.z += a.z;
  }
};

static inline void contact_one(Sim& s, Inter& I, double dt, Acc& acc) {
  const int t = I.type;
  const bool poly = s.cfg.mode == EDB_MODE_POLYHEDRON;
  const uint32_t i = I.i;
  Contact c;
  const Params* cp;
  double meff, reff;
  v3 pj, vj, wj;
  const v3 ri = s.r[i];
  if (t <= 3) {
    const uint32_t j = I.j;
    cp = &s.cp_pp(s.group[i], s.group[j]);
    if (!poly) {
      c = detect_vv(ri, s.radius[i], s.r[j], s.radius[j]);  // contact_sphere.hpp:143
    } else {
      const Shape &si = s.shapes[s.type[i]], &sj = s.shapes[s.type[j]];
      const v3 *vi = s.pv(i), *vj_ = s.pv(j);
      const double rmi = s.h[i] * si.radius, rmj = s.h[j] * sj.radius;
      if (t == 0) c = detect_vv(vi[I.sub_i], rmi, vj_[I.sub_j], rmj);
      else if (t == 1) c = detect_ve(vi[I.sub_i], rmi, vj_[sj.edges[2 * I.sub_j]], vj_[sj.edges[2 * I.sub_j + 1]], rmj);
      else if (t == 2) c = detect_vf(vi[I.sub_i], rmi, vj_, rmj, &sj.fidx[sj.foff[I.sub_j]], sj.fn[I.sub_j]);
      else c = detect_ee(vi[si.edges[2 * I.sub_i]], vi[si.edges[2 * I.sub_i + 1]], rmi, vj_[sj.edges[2 * I.sub_j]], vj_[sj.edges[2 * I.sub_j + 1]], rmj);
    }
    pj = s.r[j]; vj = s.v[j]; wj = s.w[j];
    if (!poly) { meff = (s.mass[i] * s.mass[j]) / (s.mass[i] + s.mass[j]); reff = (s.radius[i] * s.radius[j]) / (s.radius[i] + s.radius[j]); }
    else {
      const double rmi = s.h[i] * s.shapes[s.type[i]].radius, rmj = s.h[j] * s.shapes[s.type[j]].radius;
      meff = (s.mass[i] * s.mass[j]) / (s.mass[i] + s.mass[j]); reff = (rmi * rmj) / (rmi + rmj);  // contact.hpp:183
    }
  } else if (t >= 7) {
    // RShape mesh waves: sphere::ContactLawRShapeDriverFunc (contact_sphere.hpp:306-396), polyhedron:: (contact.hpp:309-432)
    const Driver& D = s.drivers[I.id_j];
    const Shape& sd = s.mesh_shapes[D.shape];
    const v3* mv = s.mesh_verts[I.id_j].data();
    const double rd = sd.radius;
    cp = &s.cp_pd(s.group[i], (uint32_t)I.id_j);
    meff = s.mass[i];
    pj = D.center; vj = D.vel; wj = D.vrot;
    if (!poly) {
      const double ra = s.radius[i];
      if (t == 7) c = detect_vv(ri, ra, mv[I.sub_j], rd);
      else if (t == 8) c = detect_ve(ri, ra, mv[sd.edges[2 * I.sub_j]], mv[sd.edges[2 * I.sub_j + 1]], rd);
      else c = detect_vf(ri, ra, mv, rd, &sd.fidx[sd.foff[I.sub_j]], sd.fn[I.sub_j]);
      reff = ra;
    } else {
      const Shape& si = s.shapes[s.type[i]];
      const v3* vi = s.pv(i);
      const double ra = s.h[i] * si.radius;
      if (t == 7) c = detect_vv(vi[I.sub_i], ra, mv[I.sub_j], rd);
      else if (t == 8) c = detect_ve(vi[I.sub_i], ra, mv[sd.edges[2 * I.sub_j]], mv[sd.edges[2 * I.sub_j + 1]], rd);
      else if (t == 9) c = detect_vf(vi[I.sub_i], ra, mv, rd, &sd.fidx[sd.foff[I.sub_j]], sd.fn[I.sub_j]);
      else if (t == 10) c = detect_ee(vi[si.edges[2 * I.sub_i]], vi[si.edges[2 * I.sub_i + 1]], ra, mv[sd.edges[2 * I.sub_j]], mv[sd.edges[2 * I.sub_j + 1]], rd);
      else if (t == 11) c = detect_ve(mv[I.sub_j], rd, vi[si.edges[2 * I.sub_i]], vi[si.edges[2 * I.sub_i + 1]], ra);   // detect<11>: roles inverted
      else c = detect_vf(mv[I.sub_j], rd, vi, ra, &si.fidx[si.foff[I.sub_i]], si.fn[I.sub_i]);                          // detect<12>
      reff = (ra * rd) / (ra + rd);   // contact.hpp:388 compute_effective_radius
    }
  } else {
    const Driver& D = s.drivers[I.id_j];
    cp = &s.cp_pd(s.group[i], (uint32_t)I.id_j);
    if (!poly) { c = driver_detect(D, s.radius[i], ri); reff = s.radius[i]; }              // contact_sphere.hpp:248,270
    else { const Shape& si = s.shapes[s.type[i]]; const double rm = s.h[i] * si.radius; c = driver_detect(D, rm, s.pv(i)[I.sub_i]); reff = rm; }  // contact.hpp:262,285
    meff = s.mass[i];
    pj = D.center; vj = D.vel; wj = D.vrot;
  }
  bool contact = c.hit;
  if (s.law == LAW_COHESIVE) contact = (contact || c.dn <= cp->dncut);
  v3 fn{0, 0, 0};
  double dn = c.dn;
  if (contact && (t == 11 || t == 12)) {
    // "j to i" (contact.hpp:405-421): the driver plays i in the law, the particle receives -f; the stored friction flips sign
    v3 f{0, 0, 0};
    contact_force(s.law, c.dn, c.n, dt, *cp, meff, reff, I.ft, c.cp, pj, vj, f, I.mom, wj, ri, s.v[i], s.w[i]);
    fn = I.ft - f;
    acc.add(s.mom[i], moments(c.cp, ri, -f, -I.mom));
    acc.add(s.f[i], -f);
    I.ft = -I.ft;
    if (s.drivers[I.id_j].need_forces) acc.add(s.drivers[I.id_j].forces, f);
  } else if (contact) {
    v3 f{0, 0, 0};
    contact_force(s.law, c.dn, c.n, dt, *cp, meff, reff, I.ft, c.cp, ri, s.v[i], f, I.mom, s.w[i], pj, vj, wj);
    fn = f - I.ft;
    // virtual bricks: the PARTNER (j, or i for swapped polyhedron records) is another rank's particle: its share is dropped
    if (!(I.foreign && I.swap) && !(I.drop_owner && !I.swap)) {
      acc.add(s.mom[i], moments(c.cp, ri, f, I.mom));
      acc.add(s.f[i], f);
    }
    if (t <= 3) {
      if ((poly || s.symetric) && !(I.foreign && !I.swap) && !(I.drop_owner && I.swap)) {
        const uint32_t j = I.j;
        acc.add(s.mom[j], moments(c.cp, pj, -f, -I.mom));
        acc.add(s.f[j], -f);
      }
    } else {
      Driver& D = s.drivers[I.id_j];
      if (D.need_forces) acc.add(D.forces, -f);
      // contact_sphere.hpp:387-389 (spheres against an RShape that needs its moment)
      if (!poly && t >= 7 && D.kind == DRV_RSHAPE && rshape_need_moment(D)) acc.add(D.mom, moments(c.cp, D.center, -f, -I.mom));
    }
  } else {
    I.ft = {0, 0, 0}; I.mom = {0, 0, 0};
    dn = 0;
  }
  I.dn = dn; I.cp = c.cp; I.fn = fn;
}

static void contact_all(Sim& s, double dt) {
  const bool par = s.omp_threads > 1;
  Acc acc{s, par};
  // wave order of the operators: contact_force_sphere.cu.in:159-170 (0,4,5,6,7..9), contact_force_polyhedron.cu.in:186-216 (0..3,7..12,4,5,6)
  static const int order_s[] = {0, 4, 5, 6, 7, 8, 9, 1, 2, 3, 10, 11, 12, 13};
  static const int order_p[] = {0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 4, 5, 6, 13};
  const int* order = s.cfg.mode == EDB_MODE_POLYHEDRON ? order_p : order_s;
  for (int o = 0; o < EDB_NTYPES; o++) {
    const int t = order[o];
    auto& w = s.wave[t];
    const long n = (long)w.size();
    if (!par) {
      for (long k = 0; k < n; k++) contact_one(s, s.inter[w[k]], dt, acc);
    } else {
#pragma omp parallel for schedule(static) num_threads(s.omp_threads)
      for (long k = 0; k < n; k++) contact_one(s, s.inter[w[k]], dt, acc);
    }
    s.interaction_updates += (uint64_t)n;
  }
}

// --- per-particle operators --------------------------------------------------------------------------------
#define PFOR(n) _Pragma("omp parallel for schedule(static) num_threads(s.omp_threads) if (s.omp_threads > 1)") for (long p = 0; p < (long)(n); p++)

static void prolog(Sim& s, double dt) {
  const double dt2 = dt * dt * 0.5, hdt = dt * 0.5;
  PFOR(s.nreal) {
    // PushVec3SecondOrderFunctor{dt, dt^2/2}: r += v*dt + a*dt2 ; PushVec3FirstOrderFunctor{dt/2}: v += a*dt/2  (exaNBody, unpinned R2)
    v3 &r = s.r[p], &v = s.v[p];
    const v3 a = s.f[p];
    r.x += v.x * dt + a.x * dt2; r.y += v.y * dt + a.y * dt2; r.z += v.z * dt + a.z * dt2;
    v.x += a.x * hdt; v.y += a.y * hdt; v.z += a.z * hdt;
    // push_to_quaternion.hpp:26-38
    q4& Q = s.q[p];
    const q4 dq = qdot(Q, s.w[p]);
    Q = {Q.w + dq.w * dt, Q.x + dq.x * dt, Q.y + dq.y * dt, Q.z + dq.z * dt};
    const double nn = std::sqrt(Q.w * Q.w + Q.x * Q.x + Q.y * Q.y + Q.z * Q.z);
    Q = {Q.w / nn, Q.x / nn, Q.y / nn, Q.z / nn};
    s.w[p] = s.w[p] + hdt * s.arot[p];
  }
}
static void push_to_quaternion(Sim& s, double dt) {
  const double hdt = dt * 0.5;
  PFOR(s.nreal) {
    q4& Q = s.q[p];
    const q4 dq = qdot(Q, s.w[p]);
    Q = {Q.w + dq.w * dt, Q.x + dq.x * dt, Q.y + dq.y * dt, Q.z + dq.z * dt};
    const double nn = std::sqrt(Q.w * Q.w + Q.x * Q.x + Q.y * Q.y + Q.z * Q.z);
    Q = {Q.w / nn, Q.x / nn, Q.y / nn, Q.z / nn};
    s.w[p] = s.w[p] + hdt * s.arot[p];
  }
}
static void angular_accel(Sim& s) {  // angular_acceleration.hpp:27-43
  PFOR(s.nreal) {
    const q4 Qi = qconj(s.q[p]);
    const v3 o = qrot(Qi, s.w[p]), M = qrot(Qi, s.mom[p]), I = s.inertia[p];
    const v3 dw{(M.x - (I.z - I.y) * o.y * o.z) / I.x, (M.y - (I.x - I.z) * o.z * o.x) / I.y, (M.z - (I.y - I.x) * o.x * o.y) / I.z};
    s.arot[p] = qrot(s.q[p], dw);
  }
}
static void angular_vel(Sim& s, double dt) {  // angular_velocity.hpp:27-33
  const double hdt = dt * 0.5;
  PFOR(s.nreal) s.w[p] = s.w[p] + s.arot[p] * hdt;
}
static void epilog(Sim& s, double dt) {
  const double hdt = dt * 0.5;
  PFOR(s.nreal) {
    const q4 Qi = qconj(s.q[p]);
    const v3 o = qrot(Qi, s.w[p]), M = qrot(Qi, s.mom[p]), I = s.inertia[p];
    const v3 dw{(M.x - (I.z - I.y) * o.y * o.z) / I.x, (M.y - (I.x - I.z) * o.z * o.x) / I.y, (M.z - (I.y - I.x) * o.x * o.y) / I.z};
    s.arot[p] = qrot(s.q[p], dw);
    s.w[p] = s.w[p] + s.arot[p] * hdt;
    v3& v = s.v[p];
    const v3 a = s.f[p];
    v.x += a.x * hdt; v.y += a.y * hdt; v.z += a.z * hdt;
  }
}
static void reset_fm(Sim& s) {
  PFOR(s.ntot) { s.f[p] = {0, 0, 0}; s.mom[p] = {0, 0, 0}; }
  for (auto& d : s.drivers) d.forces = {0, 0, 0};
}
static void gravity(Sim& s, const double* g) {
  PFOR(s.nreal) { s.f[p].x += g[0] * s.mass[p]; s.f[p].y += g[1] * s.mass[p]; s.f[p].z += g[2] * s.mass[p]; }
  // drivers (gravity_force.cu:72-83, GravityForceDriverFunctor gravity_force.hpp:26-37): only RShape drivers with the PARTICLE motion
  for (Driver& d : s.drivers)
    if (d.kind == DRV_RSHAPE && d.motion == M_PARTICLE) d.forces = d.forces + v3{g[0], g[1], g[2]} * d.mass;
}
static void force_to_accel(Sim& s) {
  PFOR(s.nreal) { s.f[p].x /= s.mass[p]; s.f[p].y /= s.mass[p]; s.f[p].z /= s.mass[p]; }
}
// particle_displ_over (exaNBody, unpinned R3) / vertex_displ_over (backup_dem.hpp:66-100): count of particles
// (vertices) whose displacement since the last rebuild is >= threshold
static uint64_t displ_over(Sim& s, double thr) {
  uint64_t n = 0;
  const double t2 = thr * thr;
  for (size_t p = 0; p < s.nreal && p < s.r_backup.size(); p++) {
    if (s.cfg.mode == EDB_MODE_SPHERE) {
      const v3 d = s.r[p] - s.r_backup[p];
      if (dot(d, d) >= t2) n++;
    } else {
      const Shape& sh = s.shapes[s.type[p]];
      for (int k = 0; k < sh.nv(); k++) {
        const v3 vo = s.r_backup[p] + qrot(s.q_backup[p], s.h[p] * sh.v[k]);
        const v3 vn = s.r[p] + qrot(s.q[p], s.h[p] * sh.v[k]);
        const v3 d = vn - vo;
        if (dot(d, d) >= t2) n++;
      }
    }
  }
  return n;
}

}  // namespace

// =============================================================================================================
extern "C" {

struct orc_sim;  // opaque == Sim

int orc_create(const edb_config* cfg, orc_sim** out) {
  Sim* s = new Sim();
  s->cfg = *cfg;
  *out = (orc_sim*)s;
  return 0;
}
int orc_destroy(orc_sim* h) { delete (Sim*)h; return 0; }
int orc_set_virtual_bricks(orc_sim* h, const int32_t* pg) {
  Sim& s = *(Sim*)h;
  for (int d = 0; d < 3; d++) s.vb[d] = pg[d] < 1 ? 1 : pg[d];
  return 0;
}
// virtual bricks bounded by cell-index cuts (vb[d] + 1 entries per axis; NULL = equal parts): the emulation of edb_set_brick_cuts
int orc_set_virtual_cuts(orc_sim* h, const int32_t* cx, const int32_t* cy, const int32_t* cz) {
  Sim& s = *(Sim*)h;
  const int32_t* in[3] = {cx, cy, cz};
  for (int d = 0; d < 3; d++) {
    s.vcut[d].clear();
    if (in[d]) s.vcut[d].assign(in[d], in[d] + s.vb[d] + 1);
  }
  return 0;
}
int orc_set_threads(orc_sim* h, int n) { ((Sim*)h)->omp_threads = n < 1 ? 1 : n; return 0; }
// threads a parallel region really gets when `n` are requested (affinity masks, OMP_THREAD_LIMIT ... can make it fewer)
int orc_threads_in_parallel(int n) {
  int got = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n < 1 ? 1 : n)
  {
#pragma omp single
    got = omp_get_num_threads();
  }
#endif
  return got;
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static Params to_params(const edb_contact_params& c) {
  Params p; p.kn = c.kn; p.kt = c.kt; p.kr = c.kr; p.mu = c.mu; p.damp = c.damp_rate; p.fc = c.fc; p.dncut = c.dncut; p.gamma = c.gamma;
  return p;
}
int orc_set_contact_params(orc_sim* h, const edb_contact_params* a, const edb_contact_params* b) {
  Sim& s = *(Sim*)h;
  s.multimat = false;
  s.pp.assign(1, to_params(*a));
  s.pd.assign(1, b ? to_params(*b) : Params());
  return 0;
}
int orc_set_contact_params_multimat(orc_sim* h, int32_t n_groups, const edb_contact_params* pp, int32_t n_drivers, const edb_contact_params* pd) {
  Sim& s = *(Sim*)h;
  s.multimat = true;
  s.n_groups = n_groups;
  s.pp.clear(); s.pd.clear();
  for (int i = 0; i < n_groups * n_groups; i++) s.pp.push_back(to_params(pp[i]));
  for (int i = 0; i < n_groups * n_drivers; i++) s.pd.push_back(to_params(pd[i]));
  return 0;
}
int orc_set_law(orc_sim* h, int32_t law) { ((Sim*)h)->law = law; return 0; }
int orc_set_drivers(orc_sim* h, int32_t n, const edb_driver* d) {
  Sim& s = *(Sim*)h;
  s.drivers.clear();
  for (int i = 0; i < n; i++) {
    Driver D;
    D.kind = d[i].kind; D.need_forces = d[i].need_forces; D.radius = d[i].radius;
    D.axis = {d[i].axis[0], d[i].axis[1], d[i].axis[2]};
    D.center = {d[i].center[0], d[i].center[1], d[i].center[2]};
    if (D.kind == DRV_CYLINDER) D.center = cmul(D.axis, D.center);  // cylinder.hpp:104
    D.offset = d[i].offset; D.normal = {d[i].normal[0], d[i].normal[1], d[i].normal[2]};
    D.vel = {d[i].vel[0], d[i].vel[1], d[i].vel[2]}; D.vrot = {d[i].vrot[0], d[i].vrot[1], d[i].vrot[2]};
    D.shape = d[i].shape;
    D.quat = {d[i].quat[0], d[i].quat[1], d[i].quat[2], d[i].quat[3]};
    D.cproj = D.offset * D.normal;   // surface.hpp:149
    s.drivers.push_back(D);
  }
  s.drivers_backup.clear();
  s.mesh_verts.assign(s.drivers.size(), {});
  for (size_t k = 0; k < s.drivers.size(); k++) {
    const Driver& D = s.drivers[k];
    if (D.kind != DRV_RSHAPE) continue;
    if (D.shape < 0 || D.shape >= (int)s.mesh_shapes.size()) { s.err = "RSHAPE driver without a mesh shape"; return EDB_ERR_INVALID; }
    for (const v3& v : s.mesh_shapes[D.shape].v) s.mesh_verts[k].push_back(D.center + qrot(D.quat, 1.0 * v));  // shape.hpp:259-262, h = 1
  }
  return 0;
}
static void load_shapes(std::vector<Shape>& dst, int32_t n, const edb_shape* sh) {
  dst.clear();
  for (int i = 0; i < n; i++) {
    Shape S;
    for (int k = 0; k < sh[i].nv; k++) S.v.push_back({sh[i].vertices[3 * k], sh[i].vertices[3 * k + 1], sh[i].vertices[3 * k + 2]});
    S.edges.assign(sh[i].edges, sh[i].edges + 2 * sh[i].ne);
    int pos = 0;
    for (int f = 0; f < sh[i].nf; f++) {
      const int nfv = sh[i].faces[pos++];
      S.foff.push_back((int)S.fidx.size()); S.fn.push_back(nfv);
      for (int k = 0; k < nfv; k++) S.fidx.push_back(sh[i].faces[pos++]);
    }
    const double* o = sh[i].obb;
    S.obb = Obb{{o[0], o[1], o[2]}, {o[3], o[4], o[5]}, {o[6], o[7], o[8]}, {o[9], o[10], o[11]}, {o[12], o[13], o[14]}};
    S.radius = sh[i].radius;
    dst.push_back(S);
  }
}
int orc_set_driver_shapes(orc_sim* h, int32_t n, const edb_shape* sh) {
  load_shapes(((Sim*)h)->mesh_shapes, n, sh);
  return 0;
}
int orc_set_shapes(orc_sim* h, int32_t n, const edb_shape* sh) {
  Sim& s = *(Sim*)h;
  s.shapes.clear();
  s.nvmax = 0;
  for (int i = 0; i < n; i++) {
    Shape S;
    for (int k = 0; k < sh[i].nv; k++) S.v.push_back({sh[i].vertices[3 * k], sh[i].vertices[3 * k + 1], sh[i].vertices[3 * k + 2]});
    S.edges.assign(sh[i].edges, sh[i].edges + 2 * sh[i].ne);
    int pos = 0;
    for (int f = 0; f < sh[i].nf; f++) {
      const int nfv = sh[i].faces[pos++];
      S.foff.push_back((int)S.fidx.size()); S.fn.push_back(nfv);
      for (int k = 0; k < nfv; k++) S.fidx.push_back(sh[i].faces[pos++]);
    }
    const double* o = sh[i].obb;
    S.obb = Obb{{o[0], o[1], o[2]}, {o[3], o[4], o[5]}, {o[6], o[7], o[8]}, {o[9], o[10], o[11]}, {o[12], o[13], o[14]}};
    S.radius = sh[i].radius;
    s.nvmax = std::max(s.nvmax, S.nv());
    s.shapes.push_back(S);
  }
  return 0;
}

#define LD3(dst, a, b, c) if (P->a && P->b && P->c) for (size_t p = 0; p < n; p++) s.dst[p] = {P->a[p], P->b[p], P->c[p]}
#define ST3(src, a, b, c) if (P->a && P->b && P->c) for (size_t p = 0; p < n; p++) { P->a[p] = s.src[p].x; P->b[p] = s.src[p].y; P->c[p] = s.src[p].z; }
int orc_upload_particles(orc_sim* h, const edb_particles* P) {
  Sim& s = *(Sim*)h;
  const size_t n = P->n;
  s.nreal = n;
  s.gsrc.clear(); s.gshift.clear();
  s.resize_particles(n);
  for (size_t p = 0; p < n; p++) {
    s.f[p] = {0, 0, 0}; s.mom[p] = {0, 0, 0}; s.w[p] = {0, 0, 0}; s.arot[p] = {0, 0, 0}; s.v[p] = {0, 0, 0};
    s.inertia[p] = {1, 1, 1}; s.q[p] = {1, 0, 0, 0}; s.h[p] = 1.0; s.mass[p] = 1.0; s.radius[p] = 0.5;
    s.id[p] = p; s.type[p] = 0; s.group[p] = 0;
  }
  LD3(r, rx, ry, rz); LD3(v, vx, vy, vz); LD3(f, fx, fy, fz); LD3(mom, momx, momy, momz); LD3(w, vrotx, vroty, vrotz);
  LD3(arot, arotx, aroty, arotz); LD3(inertia, inertiax, inertiay, inertiaz);
  if (P->qw) for (size_t p = 0; p < n; p++) s.q[p] = {P->qw[p], P->qx[p], P->qy[p], P->qz[p]};
  if (P->mass) std::copy(P->mass, P->mass + n, s.mass.begin());
  if (P->homothety) std::copy(P->homothety, P->homothety + n, s.h.begin());
  if (P->radius) std::copy(P->radius, P->radius + n, s.radius.begin());
  if (P->id) std::copy(P->id, P->id + n, s.id.begin());
  if (P->type) std::copy(P->type, P->type + n, s.type.begin());
  if (P->group) std::copy(P->group, P->group + n, s.group.begin());
  s.inter.clear();
  for (auto& w : s.wave) w.clear();
  return 0;
}
int orc_download_particles(orc_sim* h, edb_particles* P) {
  Sim& s = *(Sim*)h;
  const size_t n = s.nreal;
  P->n = n;
  ST3(r, rx, ry, rz) ST3(v, vx, vy, vz) ST3(f, fx, fy, fz) ST3(mom, momx, momy, momz) ST3(w, vrotx, vroty, vrotz)
  ST3(arot, arotx, aroty, arotz) ST3(inertia, inertiax, inertiay, inertiaz)
  if (P->qw) for (size_t p = 0; p < n; p++) { P->qw[p] = s.q[p].w; P->qx[p] = s.q[p].x; P->qy[p] = s.q[p].y; P->qz[p] = s.q[p].z; }
  if (P->mass) std::copy(s.mass.begin(), s.mass.begin() + n, P->mass);
  if (P->homothety) std::copy(s.h.begin(), s.h.begin() + n, P->homothety);
  if (P->radius) std::copy(s.radius.begin(), s.radius.begin() + n, P->radius);
  if (P->id) std::copy(s.id.begin(), s.id.begin() + n, P->id);
  if (P->type) std::copy(s.type.begin(), s.type.begin() + n, P->type);
  if (P->group) std::copy(s.group.begin(), s.group.begin() + n, P->group);
  return 0;
}
int orc_num_particles(orc_sim* h, uint64_t* nr, uint64_t* ng) {
  Sim& s = *(Sim*)h;
  if (nr) *nr = s.nreal;
  if (ng) *ng = s.ntot - s.nreal;
  return 0;
}

int orc_update_particle_neighbors(orc_sim* h, double rcut_inc, int32_t symetric) { rebuild(*(Sim*)h, rcut_inc, symetric); return 0; }
int orc_ghost_update_rq(orc_sim* h) { ghost_refresh(*(Sim*)h, false); return 0; }
int orc_compute_vertices(orc_sim* h) { compute_vertices(*(Sim*)h); return 0; }
int orc_reset_force_moment(orc_sim* h) { reset_fm(*(Sim*)h); return 0; }
int orc_gravity_force(orc_sim* h, const double* g) { gravity(*(Sim*)h, g); return 0; }
int orc_contact_sphere(orc_sim* h, double dt) { contact_all(*(Sim*)h, dt); return 0; }
int orc_contact_polyhedron(orc_sim* h, double dt) { contact_all(*(Sim*)h, dt); return 0; }
int orc_force_to_accel(orc_sim* h) { force_to_accel(*(Sim*)h); return 0; }
int orc_combined_compute_prolog(orc_sim* h, double dt) { prolog(*(Sim*)h, dt); return 0; }
int orc_combined_compute_epilog(orc_sim* h, double dt) { epilog(*(Sim*)h, dt); return 0; }
int orc_push_to_quaternion(orc_sim* h, double dt) { push_to_quaternion(*(Sim*)h, dt); return 0; }
int orc_push_to_angular_velocity(orc_sim* h, double dt) { angular_vel(*(Sim*)h, dt); return 0; }
int orc_push_to_angular_acceleration(orc_sim* h) { angular_accel(*(Sim*)h); return 0; }
int orc_freeze_particles(orc_sim* h) {
  Sim& s = *(Sim*)h;
  for (size_t p = 0; p < s.nreal; p++) { s.f[p] = {0, 0, 0}; s.mom[p] = {0, 0, 0}; s.v[p] = {0, 0, 0}; s.w[p] = {0, 0, 0}; }
  return 0;
}
// the `region` slot of reset_force_moment / freeze_particles (reset_force_moment.cu:36-57, freeze_particles.cu:47-66): exaNBody's
// ParticleRegion::contains (box, id range, quadric) under a ParticleRegionCSG expression, here in reverse Polish notation
static bool orc_in_region(const edb_region& R, const v3& r, uint64_t id) {
  if (R.use_bounds && !(r.x >= R.lo[0] && r.x < R.hi[0] && r.y >= R.lo[1] && r.y < R.hi[1] && r.z >= R.lo[2] && r.z < R.hi[2])) return false;
  if (R.use_ids && !(id >= R.id_start && id < R.id_end)) return false;
  if (R.use_quadric) {
    const double v[4] = {r.x, r.y, r.z, 1.0};
    double s = 0.0;
    for (int i = 0; i < 4; i++) {
      double t = 0.0;
      for (int j = 0; j < 4; j++) t += R.quadric[4 * i + j] * v[j];
      s += v[i] * t;
    }
    if (!(s <= 0.0)) return false;
  }
  return true;
}
static bool orc_csg(const Sim& s, int n, const int32_t* tok, const v3& r, uint64_t id) {
  if (n == 0) return true;
  std::vector<bool> st;
  for (int q = 0; q < n; q++) {
    const int t = tok[q];
    if (t >= 0) st.push_back(orc_in_region(s.regions[t], r, id));
    else if (t == EDB_CSG_NOT) st.back() = !st.back();
    else { const bool b = st.back(); st.pop_back(); const bool a = st.back(); st.back() = t == EDB_CSG_AND ? (a && b) : (a || b); }
  }
  return st.back();
}
int orc_set_particle_regions(orc_sim* h, int32_t n, const edb_region* r) {
  Sim& s = *(Sim*)h;
  s.regions.assign(r, r + n);
  return 0;
}
int orc_reset_force_moment_region(orc_sim* h, int32_t n, const int32_t* tok) {
  Sim& s = *(Sim*)h;
  for (size_t p = 0; p < s.nreal; p++)
    if (orc_csg(s, n, tok, s.r[p], s.id[p])) { s.f[p] = {0, 0, 0}; s.mom[p] = {0, 0, 0}; }
  return 0;
}
int orc_freeze_particles_region(orc_sim* h, int32_t n, const int32_t* tok) {
  Sim& s = *(Sim*)h;
  for (size_t p = 0; p < s.nreal; p++)
    if (orc_csg(s, n, tok, s.r[p], s.id[p])) { s.f[p] = {0, 0, 0}; s.mom[p] = {0, 0, 0}; s.v[p] = {0, 0, 0}; s.w[p] = {0, 0, 0}; }
  return 0;
}
int orc_displ_over(orc_sim* h, double thr, uint64_t* over) { *over = displ_over(*(Sim*)h, thr); return 0; }
int orc_compute_driver_force_moment(orc_sim* h, double* forces, double* moments) {
  // compute_driver_force_moment.cu:34-54: per driver, sum of -(fn+ft) and of cross(cp - center, -(fn+ft)) over driver waves
  Sim& s = *(Sim*)h;
  for (size_t d = 0; d < s.drivers.size(); d++) for (int k = 0; k < 3; k++) { forces[3 * d + k] = 0; moments[3 * d + k] = 0; }
  for (int t = 4; t <= 12; t++)
    for (uint32_t k : s.wave[t]) {
      const Inter& I = s.inter[k];
      const size_t d = I.id_j;
      const v3 f = -(I.fn + I.ft);
      const v3 m = cross(I.cp - s.drivers[d].center, f) + -I.mom;
      forces[3 * d] += f.x; forces[3 * d + 1] += f.y; forces[3 * d + 2] += f.z;
      moments[3 * d] += m.x; moments[3 * d + 1] += m.y; moments[3 * d + 2] += m.z;
    }
  return 0;
}

// stress_tensor.cu:41-81 (wave loop from NTypes-1 down to 0, :117-120): records with dn < 0 add tensor(fn + ft, cp - r_i) to
// particle i and, for the particle-particle waves, tensor(-(fn + ft), cp - r_j) to particle j; shares of ghost copies are dropped
int orc_stress_tensor(orc_sim* h, double* out) {
  Sim& s = *(Sim*)h;
  for (size_t q = 0; q < 9 * s.nreal; q++) out[q] = 0.0;
  auto add = [&](uint32_t p, v3 f, v3 c) {
    double* S = out + 9 * (size_t)p;
    S[0] += f.x * c.x; S[1] += f.x * c.y; S[2] += f.x * c.z;
    S[3] += f.y * c.x; S[4] += f.y * c.y; S[5] += f.y * c.z;
    S[6] += f.z * c.x; S[7] += f.z * c.y; S[8] += f.z * c.z;
  };
  for (int t = EDB_NTYPES - 1; t >= 0; t--)
    for (uint32_t k : s.wave[t]) {
      const Inter& I = s.inter[k];
      if (!(I.dn < 0.0)) continue;
      const v3 fij = I.fn + I.ft;
      if (I.i < s.nreal && !(I.foreign && I.swap) && !(I.drop_owner && !I.swap)) add(I.i, fij, I.cp - s.r[I.i]);
      if (t <= 3 && I.j < s.nreal && !(I.foreign && !I.swap) && !(I.drop_owner && I.swap)) add(I.j, -fij, I.cp - s.r[I.j]);
    }
  return 0;
}

// simulation_state.cu:80-175 + simulation_state.hpp:48-57 + itools/counter.hpp:125-152
int orc_simulation_state(orc_sim* h, int32_t symetric, edb_sim_state* out) {
  Sim& s = *(Sim*)h;
  memset(out, 0, sizeof(*out));
  for (size_t p = 0; p < s.nreal; p++) {
    const double m = s.mass[p];
    out->mass += m;
    out->rotation_energy[0] += 0.5 * s.w[p].x * s.w[p].x * m; out->rotation_energy[1] += 0.5 * s.w[p].y * s.w[p].y * m;
    out->rotation_energy[2] += 0.5 * s.w[p].z * s.w[p].z * m;
    out->kinetic_energy[0] += 0.5 * s.v[p].x * s.v[p].x * m; out->kinetic_energy[1] += 0.5 * s.v[p].y * s.v[p].y * m;
    out->kinetic_energy[2] += 0.5 * s.v[p].z * s.v[p].z * m;
    out->n_particles += 1;
  }
  for (int t = 0; t < EDB_NTYPES; t++) {
    const uint64_t coef = (t <= 3 && symetric) ? 2 : 1;
    for (uint32_t k : s.wave[t]) {
      const Inter& I = s.inter[k];
      if (I.ghost == 2) continue;   // PartnerGhost
      out->n_total_interactions += coef;
      const bool active = I.ft.x != 0 || I.ft.y != 0 || I.ft.z != 0 || I.mom.x != 0 || I.mom.y != 0 || I.mom.z != 0;
      if (I.dn < 0.0 || active) {
        out->n_active_interactions += coef;
        out->min_dn = std::min(out->min_dn, I.dn);
      }
    }
  }
  return 0;
}

// ---- inner bonds: SURVEY 8(f) rank 4 (checker side) ----
static double face_area(const Shape& sh, int f, double h = 1.0) {   // shape.hpp:673-697 (fan triangulation), :359-362 (h^2 scaling)
  double surface = 0;
  const int* idx = &sh.fidx[sh.foff[f]];
  const v3 v0 = sh.v[idx[0]];
  for (int j = 1; j < sh.fn[f] - 1; j++) surface += 0.5 * norm(cross(sh.v[idx[j]] - v0, sh.v[idx[j + 1]] - v0));
  return surface * h * h;
}
// stick_polyhedra.cpp:88-360: bonds between the matching faces of two polyhedra of the same cluster (id_i < id_j), one bond per
// vertex pair closer than the sticking threshold, weight 1/#bonds, rupture criterion from the face area
static void stick_polyhedra(Sim& s, double dn_crit) {
  s.bonds.clear();
  for (int ck = 1; ck <= s.nc[2]; ck++)
    for (int cj = 1; cj <= s.nc[1]; cj++)
      for (int ci = 1; ci <= s.nc[0]; ci++) {
        const size_t ca = s.cell_of(ci, cj, ck);
        for (uint32_t ia = s.cstart[ca]; ia < s.cstart[ca + 1]; ia++) {
          const uint32_t a = s.citems[ia];
          if (a >= s.nreal) continue;
          const Shape& sa = s.shapes[s.type[a]];
          const v3* va = s.pv(a);
          for_each_neighbor(s, ci, cj, ck, a, [&](uint32_t b, bool ghostc) {
            if (s.cluster[b] != s.cluster[a]) return;
            if (s.id[a] >= s.id[b]) return;
            const Shape& sb = s.shapes[s.type[b]];
            const v3* vb = s.pv(b);
            const BondParams& P = s.ibp[(size_t)s.group[a] * s.ibp_groups + s.group[b]];
            std::vector<Sim::Bond> local;
            bool found = false;
            for (int i = 0; i < sa.nf() && !found; i++) {
              const int* fi = &sa.fidx[sa.foff[i]];
              const int ni_ = sa.fn[i];
              if (ni_ < 3) continue;
              v3 ni = cross(va[fi[1]] - va[fi[0]], va[fi[2]] - va[fi[0]]);
              ni = ni * (1.0 / norm(ni));
              const double area_i = face_area(sa, i);
              for (int j = 0; j < sb.nf() && !found; j++) {
                const int* fj = &sb.fidx[sb.foff[j]];
                const int nj_ = sb.fn[j];
                if (nj_ < 3 || ni_ != nj_) continue;
                const double ratio = area_i / face_area(sb, j);
                if (ratio > 1.01 || ratio < 0.99) continue;
                v3 nj = cross(vb[fj[1]] - vb[fj[0]], vb[fj[2]] - vb[fj[0]]);
                nj = nj * (1.0 / norm(nj));
                if (norm(cross(ni, nj)) > 1.e-10) continue;
                const double area = face_area(sa, i, s.h[a]);   // pi.id_ < pj.id_ always holds here
                Sim::Bond B;
                B.id_i = s.id[a]; B.id_j = s.id[b]; B.i = a; B.j = b; B.ghost = ghostc ? 1 : 0; B.mode = P.mode;
                if (P.mode == 0) B.c1 = 2 * area * P.crit1;
                else if (P.mode == 1) { B.c1 = 2 * area * P.crit1; B.c2 = 2 * area * P.crit2; }
                else if (P.mode == 2) { B.c1 = 2 * area * P.crit1; B.c2 = area * P.crit2; }
                found = true;
                for (int iv = 0; iv < ni_; iv++)
                  for (int jv = 0; jv < nj_; jv++) {
                    const Contact c = detect_vv(va[fi[iv]], s.h[a] * sa.radius, vb[fj[jv]], s.h[b] * sb.radius);
                    if (c.dn <= dn_crit) { B.sub_i = fi[iv]; B.sub_j = fj[jv]; B.dn0 = c.dn; local.push_back(B); break; }
                  }
                if (local.size() <= 2) { local.clear(); found = false; }
              }
            }
            for (auto& B : local) B.weight = 1. / static_cast<double>(local.size());
            // check_stiked_face (interaction_manager.hpp:226-276): the glued vertices of i are coplanar
            bool ok = true;
            if (local.size() >= 3) {
              const v3 p0 = va[local[0].sub_i], nrm = cross(va[local[1].sub_i] - p0, va[local[2].sub_i] - p0);
              for (size_t k = 3; k < local.size(); k++)
                if (std::fabs(dot(va[local[k].sub_i] - p0, nrm)) > 1.e-10) ok = false;
            }
            if (ok) s.bonds.insert(s.bonds.end(), local.begin(), local.end());
          });
        }
      }
  resolve_bonds(s);
}
// update_interfaces.cpp:44-84: maximal runs of bonds between the same two particles (PartnerGhost records skipped)
static void update_interfaces(Sim& s) {
  s.interfaces.clear();
  size_t loc = 0;
  const size_t n = s.bonds.size();
  while (loc < n) {
    if (s.bonds[loc].ghost == 2) { loc++; continue; }
    size_t m = 1;
    while (loc + m < n && s.bonds[loc + m].id_i == s.bonds[loc].id_i && s.bonds[loc + m].id_j == s.bonds[loc].id_j) m++;
    s.interfaces.push_back({loc, m});
    loc += m;
  }
  s.interface_broken.assign(s.interfaces.size(), 0);
}
// inner_bond.hpp:106-185 (polyhedron::inner_bond_law) over wave 13, serial canonical order
static void inner_bond_force(Sim& s, double dt) {
  for (auto& B : s.bonds) {
    if (B.i == 0xffffffffu || B.j == 0xffffffffu || !B.unbroken) { B.dn = 0; B.cp = {0, 0, 0}; B.fn = -B.friction; continue; }
    const uint32_t i = B.i, j = B.j;
    const Shape &si = s.shapes[s.type[i]], &sj = s.shapes[s.type[j]];
    const Contact c = detect_vv(s.pv(i)[B.sub_i], s.h[i] * si.radius, s.pv(j)[B.sub_j], s.h[j] * sj.radius);
    const BondParams& P = s.ibp[(size_t)s.group[i] * s.ibp_groups + s.group[j]];
    const double meff = (s.mass[i] * s.mass[j]) / (s.mass[i] + s.mass[j]);
    v3 fi{0, 0, 0};
    inner_bond_force_core(c.dn, c.n, B.dn0, B.weight, dt, P, meff, B.en, B.tds, B.et, B.friction, c.cp, s.r[i], s.v[i], fi, s.w[i], s.r[j],
                          s.v[j], s.w[j]);
    B.fn = fi - B.friction;
    B.dn = c.dn; B.cp = c.cp;
    s.mom[i] = s.mom[i] + moments(c.cp, s.r[i], fi, v3{0, 0, 0});
    s.f[i] = s.f[i] + fi;
    s.mom[j] = s.mom[j] + moments(c.cp, s.r[j], -fi, v3{0, 0, 0});
    s.f[j] = s.f[j] + -fi;
  }
}
// broken_interface_friction_updater.cpp:58-98
static void broken_interface_friction_updater(Sim& s) {
  for (size_t k = 0; k < s.interfaces.size(); k++) {
    if (!s.interface_broken[k]) continue;
    const auto [loc, size] = s.interfaces[k];
    const auto& B0 = s.bonds[loc];
    if (B0.i == 0xffffffffu || B0.j == 0xffffffffu) continue;
    const double mu = s.cp_pp(s.group[B0.i], s.group[B0.j]).mu;
    for (size_t j = 0; j < size; j++) {
      auto& B = s.bonds[loc + j];
      B.unbroken = 0;
      const double ft = norm(B.friction);
      const double thr = mu * norm(B.fn);
      if (ft > thr && ft > 0) B.friction = B.friction * thr / ft;
      if (B.dn > 0) B.friction = v3{0, 0, 0};
    }
  }
}
static uint64_t apply_interface_fracture_criterion(Sim& s) {
  uint64_t broken = 0;
  for (size_t k = 0; k < s.interfaces.size(); k++) {
    const auto [loc, size] = s.interfaces[k];
    std::vector<double> en(size), et(size), dn(size);
    std::vector<v3> fn(size);
    for (size_t j = 0; j < size; j++) { const auto& B = s.bonds[loc + j]; en[j] = B.en; et[j] = B.et; dn[j] = B.dn; fn[j] = B.fn; }
    const auto& B0 = s.bonds[loc];
    if (interface_breaks(B0.mode, B0.c1, B0.c2, en.data(), et.data(), dn.data(), fn.data(), size)) {
      s.interface_broken[k] = 1;
      broken++;
      for (size_t j = 0; j < size; j++) s.bonds[loc + j].unbroken = 0;
    }
  }
  return broken;
}

// ---- driver operators of the time loop (src/driver/*.cpp), SURVEY 8(f) rank 3 ----
static bool drivers_move(const Sim& s) {
  for (const Driver& d : s.drivers) if (d.motion != M_STATIONARY || (d.kind == DRV_RSHAPE && rshape_need_moment(d))) return true;
  return false;
}
static void push_f_v_r_driver(Sim& s, double time, double dt) {   // push_f_v_r_driver.cpp:28-37
  for (Driver& d : s.drivers) if (motion_triggered(d, time + dt)) driver_push_f_v_r(d, time, dt);
}
static void push_f_v_driver(Sim& s, double dt) { for (Driver& d : s.drivers) driver_push_f_v(d, dt); }   // push_f_v_driver.cpp:28-35
static void push_av_to_quat_driver(Sim& s, double time, double dt) { for (Driver& d : s.drivers) driver_push_av_to_quat(d, time, dt); }
static void reset_force_driver(Sim& s) { for (Driver& d : s.drivers) d.forces = {0, 0, 0}; }   // reset_force_driver.cpp:38-42
static void force_to_accel_driver(Sim& s, double system_mass) { for (Driver& d : s.drivers) driver_force_to_accel(d, system_mass); }
// driver_vertices.cu:28-35,62-78: vertices = center + quat * v for the non-stationary meshes (rshape.hpp:419)
static void driver_vertices(Sim& s) {
  for (size_t k = 0; k < s.drivers.size(); k++) {
    const Driver& D = s.drivers[k];
    if (D.kind != DRV_RSHAPE) continue;
    if (D.motion == M_STATIONARY && is_zero(D.vrot)) continue;
    const auto& lv = s.mesh_shapes[D.shape].v;
    for (size_t i = 0; i < lv.size(); i++) s.mesh_verts[k][i] = D.center + qrot(D.quat, lv[i]);
  }
}
// driver_displ_over.cu:60-243 (stops at the first driver that moved, :233-236)
static uint64_t driver_displ_over(Sim& s, double thr) {
  const double r2 = thr * thr;
  uint64_t local = 0;
  for (size_t i = 0; i < s.drivers_backup.size() && local == 0; i++) {
    const Driver &a = s.drivers[i], &b = s.drivers_backup[i];
    if (a.kind == DRV_SURFACE) {
      v3 d = a.cproj - b.cproj;
      if (b.motion == M_PENDULUM_MOTION) {
        auto isect = [](v3 pp, v3 pn, v3 lp, v3 ld) { const v3 delta = pp - lp; const double t = dot(delta, pn) / dot(ld, pn); return lp + t * ld; };
        const v3 pa = isect(a.pend_anchor, a.normal, a.pend_init, a.pend_dir), pb = isect(a.pend_anchor, b.normal, a.pend_init, a.pend_dir);
        d = pb - pa;
      }
      local += dot(d, d) >= r2 ? 1 : 0;
    } else if (a.kind == DRV_BALL) {
      const v3 d = a.center - b.center;
      if (m_compressive(a.motion)) { const double rd = a.radius - b.radius; local += rd * rd >= r2 ? 1 : 0; }
      else local += dot(d, d) >= r2 ? 1 : 0;
    } else if (a.kind == DRV_RSHAPE) {
      if (a.center.x == b.center.x && a.center.y == b.center.y && a.center.z == b.center.z && a.quat.w == b.quat.w && a.quat.x == b.quat.x &&
          a.quat.y == b.quat.y && a.quat.z == b.quat.z) continue;
      for (const v3& lv : s.mesh_shapes[a.shape].v) {
        const v3 d = (a.center + qrot(a.quat, 1.0 * lv)) - (b.center + qrot(b.quat, 1.0 * lv));
        if (dot(d, d) >= r2) local++;
      }
    }   // cylinder: 0 (quirk Q14)
  }
  return local;
}
static double mass_sum(const Sim& s) { double m = 0; for (size_t p = 0; p < s.nreal; p++) m += s.mass[p]; return m; }

int orc_run_verlet(orc_sim* h, double dt, const double* g, double rcut_inc, int32_t symetric, int32_t nsteps, int32_t rebuild_every, uint64_t* n_rebuilds) {
  // numerical_scheme_verlet_macro (config_numerical_schemes.msp:6-30) with, when present, the driver operators (:13-15,17,23-24) and the
  // fragmentation operators (config_fragmentation.msp:11-41 inside check_and_update_particles; inner_bond_polyhedron in compute_force):
  // the fracture criterion is evaluated on the state of the previous step; a broken interface caps its friction and forces the rebuild
  Sim& s = *(Sim*)h;
  uint64_t nb = 0;
  const bool moving = drivers_move(s);
  if (moving && !s.have_system_mass) { s.system_mass = mass_sum(s); s.have_system_mass = true; }
  if (moving && s.drivers_backup.empty()) s.drivers_backup = s.drivers;
  for (int it = 0; it < nsteps; it++) {
    const bool bonds = !s.bonds.empty();
    prolog(s, dt);
    if (moving) {
      push_f_v_r_driver(s, s.physical_time, dt);
      push_f_v_driver(s, 0.5 * dt);
      push_av_to_quat_driver(s, s.physical_time, dt);
    }
    const uint64_t broken = bonds ? apply_interface_fracture_criterion(s) : 0;
    if (broken) broken_interface_friction_updater(s);
    bool rb = broken > 0;   // displ_or_break
    if (rebuild_every > 0) rb = rb || ((it + 1) % rebuild_every) == 0;
    else if (rebuild_every == 0) rb = rb || displ_over(s, 0.5 * rcut_inc) > 0 || (moving && driver_displ_over(s, 0.5 * rcut_inc) > 0);
    if (rb) {
      if (moving) { s.drivers_backup = s.drivers; driver_vertices(s); }   // backup_drivers; ghost_full_update -> compute_all_vertices
      rebuild(s, rcut_inc, symetric); nb++;
      if (bonds) update_interfaces(s);
    } else { ghost_refresh(s, false); compute_vertices(s); if (moving) driver_vertices(s); }
    if (moving) reset_force_driver(s);
    reset_fm(s);
    gravity(s, g);
    if (bonds) inner_bond_force(s, dt);
    contact_all(s, dt);
    force_to_accel(s);
    if (moving) { force_to_accel_driver(s, s.system_mass); push_f_v_driver(s, 0.5 * dt); }
    epilog(s, dt);
    s.particle_steps += s.nreal;
    s.physical_time += dt;
  }
  if (n_rebuilds) *n_rebuilds = nb;
  return 0;
}

static inline v3 V3(const double* p) { return {p[0], p[1], p[2]}; }
static inline void S3(double* o, v3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
int orc_set_driver_motions(orc_sim* h, int32_t n, const edb_driver_motion* m) {
  Sim& s = *(Sim*)h;
  if (n != (int)s.drivers.size()) { s.err = "edb_set_driver_motions: n differs from the driver count"; return EDB_ERR_INVALID; }
  for (int i = 0; i < n; i++) {
    Driver& D = s.drivers[i];
    D.motion = m[i].motion_type;
    D.motion_vector = V3(m[i].motion_vector);
    D.t_start = m[i].motion_start_threshold; D.t_end = m[i].motion_end_threshold;
    D.const_vel = m[i].const_vel; D.const_force = m[i].const_force; D.sigma = m[i].sigma; D.damprate = m[i].damprate;
    D.tab_time.assign(m[i].tab_time, m[i].tab_time + (m[i].tab_time ? m[i].n_tab : 0));
    D.tab_pos.clear();
    for (int k = 0; k < (m[i].tab_pos ? m[i].n_tab : 0); k++) D.tab_pos.push_back(V3(m[i].tab_pos + 3 * k));
    D.omega = m[i].omega; D.amplitude = m[i].amplitude; D.shaker_dir = V3(m[i].shaker_dir);
    D.pend_anchor = V3(m[i].pendulum_anchor_point); D.pend_init = V3(m[i].pendulum_initial_position); D.pend_dir = V3(m[i].pendulum_swing_dir);
    D.expr_use_v = m[i].expr_use_v; D.expr_use_vrot = m[i].expr_use_vrot; D.expr_use_mom = m[i].expr_use_mom; D.drive_by_mom = m[i].drive_by_mom;
    D.mass = m[i].mass; D.surface = m[i].surface; D.acc = V3(m[i].acc); D.rv = m[i].rv; D.ra = m[i].ra;
    D.inertia = V3(m[i].inertia); D.applied_mom = V3(m[i].applied_mom); D.mom_axis = V3(m[i].mom_axis); D.mom = V3(m[i].mom);
    D.need_forces = m_need_forces(D.motion) ? 1 : 0;
  }
  return 0;
}
int orc_set_driver_expr(orc_sim* h, int32_t d, const double* v, const double* vrot, const double* mom) {
  Sim& s = *(Sim*)h;
  if (d < 0 || d >= (int)s.drivers.size()) return EDB_ERR_INVALID;
  if (v) s.drivers[d].expr_v = V3(v);
  if (vrot) s.drivers[d].expr_vrot = V3(vrot);
  if (mom) s.drivers[d].expr_mom = V3(mom);
  return 0;
}
int orc_get_drivers(orc_sim* h, int32_t n, edb_driver* out, edb_driver_motion* mo) {
  Sim& s = *(Sim*)h;
  if (n != (int)s.drivers.size()) return EDB_ERR_INVALID;
  for (int i = 0; i < n; i++) {
    const Driver& D = s.drivers[i];
    if (out) {
      edb_driver& o = out[i];
      o = edb_driver{};
      o.kind = D.kind; o.need_forces = D.need_forces; o.radius = D.radius; S3(o.axis, D.axis); S3(o.center, D.center);
      o.offset = D.offset; S3(o.normal, D.normal); S3(o.vel, D.vel); S3(o.vrot, D.vrot); o.shape = D.shape;
      o.quat[0] = D.quat.w; o.quat[1] = D.quat.x; o.quat[2] = D.quat.y; o.quat[3] = D.quat.z;
    }
    if (mo) {
      edb_driver_motion& o = mo[i];
      o = edb_driver_motion{};
      o.motion_type = D.motion; o.n_tab = (int)D.tab_time.size(); S3(o.motion_vector, D.motion_vector);
      o.motion_start_threshold = D.t_start; o.motion_end_threshold = D.t_end; o.const_vel = D.const_vel; o.const_force = D.const_force;
      o.sigma = D.sigma; o.damprate = D.damprate; o.omega = D.omega; o.amplitude = D.amplitude; S3(o.shaker_dir, D.shaker_dir);
      S3(o.pendulum_anchor_point, D.pend_anchor); S3(o.pendulum_initial_position, D.pend_init); S3(o.pendulum_swing_dir, D.pend_dir);
      o.expr_use_v = D.expr_use_v; o.expr_use_vrot = D.expr_use_vrot; o.expr_use_mom = D.expr_use_mom; o.drive_by_mom = D.drive_by_mom;
      o.mass = D.mass; o.surface = D.surface; S3(o.acc, D.acc); o.rv = D.rv; o.ra = D.ra;
      S3(o.inertia, D.inertia); S3(o.applied_mom, D.applied_mom); S3(o.mom_axis, D.mom_axis); S3(o.mom, D.mom); S3(o.forces, D.forces);
    }
  }
  return 0;
}
// One operator on one driver through the structs of the ABI (test interface mirrored by oracle/ref_math.cpp:ref_driver_motion_op; for a
// surface center_proj_ travels in the mom slot).  op 0 push_f_v_r (trigger-gated), 1 push_f_v, 2 push_av_to_quat, 3 force_to_accel.
int orc_driver_motion_op(int op, double time, double dt, double system_mass, edb_driver* d, edb_driver_motion* m) {
  Sim s;
  edb_driver din = *d;
  if (din.kind == EDB_DRIVER_RSHAPE) { s.mesh_shapes.resize(1); din.shape = 0; }
  double cyl_center[3] = {d->center[0], d->center[1], d->center[2]};
  if (orc_set_drivers((orc_sim*)&s, 1, &din)) return EDB_ERR_INVALID;
  Driver& D = s.drivers[0];
  if (D.kind == DRV_CYLINDER) D.center = V3(cyl_center);   // the test passes the initialised centre (already multiplied by the axis)
  orc_set_driver_motions((orc_sim*)&s, 1, m);
  D.forces = V3(m->forces);
  if (D.kind == DRV_SURFACE) { D.cproj = V3(m->mom); D.mom = {0, 0, 0}; }
  if (op == 0) { if (motion_triggered(D, time + dt)) driver_push_f_v_r(D, time, dt); }
  else if (op == 1) driver_push_f_v(D, dt);
  else if (op == 2) driver_push_av_to_quat(D, time, dt);
  else driver_force_to_accel(D, system_mass);
  const v3 cproj = D.cproj;
  orc_get_drivers((orc_sim*)&s, 1, d, m);
  if (D.kind == DRV_SURFACE) S3(m->mom, cproj);
  return 0;
}
int orc_set_inner_bond_params(orc_sim* h, int32_t n_groups, const edb_inner_bond_params* p) {
  Sim& s = *(Sim*)h;
  s.ibp_groups = n_groups;
  s.ibp.clear();
  for (int i = 0; i < n_groups * n_groups; i++) s.ibp.push_back(BondParams{p[i].kn, p[i].kt, p[i].damp_rate, p[i].mode, p[i].crit1, p[i].crit2});
  return 0;
}
int orc_set_particle_clusters(orc_sim* h, uint64_t n, const uint64_t* ids, const double* cluster) {   // field::cluster (oracle-only: stick_polyhedra)
  Sim& s = *(Sim*)h;
  std::unordered_map<uint64_t, double> m;
  for (uint64_t k = 0; k < n; k++) m[ids[k]] = cluster[k];
  s.cluster.assign(s.ntot, -1.0);
  for (size_t p = 0; p < s.ntot; p++) { auto it = m.find(s.id[p]); if (it != m.end()) s.cluster[p] = it->second; }
  return 0;
}
int orc_stick_polyhedra(orc_sim* h, double threshold) {
  Sim& s = *(Sim*)h;
  if (s.cluster.size() != s.ntot) { s.err = "stick_polyhedra: clusters not set (or stale after a rebuild)"; return EDB_ERR_STATE; }
  compute_vertices(s);
  stick_polyhedra(s, threshold);
  return 0;
}
int orc_import_inner_bonds(orc_sim* h, const edb_inner_bond* in, uint64_t n) {
  Sim& s = *(Sim*)h;
  s.bonds.clear();
  for (uint64_t k = 0; k < n; k++) {
    Sim::Bond B;
    B.id_i = in[k].id_i; B.id_j = in[k].id_j; B.sub_i = in[k].sub_i; B.sub_j = in[k].sub_j;
    B.friction = V3(in[k].friction); B.tds = V3(in[k].tds); B.en = in[k].en; B.et = in[k].et; B.dn0 = in[k].dn0; B.weight = in[k].weight;
    B.c1 = in[k].criteria_1; B.c2 = in[k].criteria_2; B.mode = in[k].mode; B.unbroken = in[k].unbroken; B.ghost = in[k].ghost;
    s.bonds.push_back(B);
  }
  resolve_bonds(s);
  return 0;
}
int orc_export_inner_bonds(orc_sim* h, edb_inner_bond* out, uint64_t cap, uint64_t* n) {
  Sim& s = *(Sim*)h;
  *n = s.bonds.size();
  if (!out) return 0;
  if (cap < s.bonds.size()) return EDB_ERR_CAPACITY;
  for (size_t k = 0; k < s.bonds.size(); k++) {
    const auto& B = s.bonds[k];
    edb_inner_bond& o = out[k];
    o = edb_inner_bond{};
    o.id_i = B.id_i; o.id_j = B.id_j; o.sub_i = B.sub_i; o.sub_j = B.sub_j; S3(o.friction, B.friction); o.en = B.en; S3(o.tds, B.tds);
    o.et = B.et; o.dn0 = B.dn0; o.weight = B.weight; o.criteria_1 = B.c1; o.criteria_2 = B.c2; o.mode = B.mode; o.unbroken = B.unbroken;
    o.ghost = B.ghost;
  }
  return 0;
}
int orc_update_interfaces(orc_sim* h, uint64_t* n) { Sim& s = *(Sim*)h; resolve_bonds(s); update_interfaces(s); if (n) *n = s.interfaces.size(); return 0; }
int orc_inner_bond_polyhedron(orc_sim* h, double dt) { Sim& s = *(Sim*)h; inner_bond_force(s, dt); return 0; }
int orc_apply_interface_fracture_criterion(orc_sim* h, uint64_t* n_broken) { *n_broken = apply_interface_fracture_criterion(*(Sim*)h); return 0; }
int orc_broken_interface_friction_updater(orc_sim* h) { broken_interface_friction_updater(*(Sim*)h); return 0; }
int orc_get_inner_bond_state(orc_sim* h, double* dn, double* cp, double* fn, double* ft) {
  Sim& s = *(Sim*)h;
  for (size_t k = 0; k < s.bonds.size(); k++) {
    const auto& B = s.bonds[k];
    if (dn) dn[k] = B.dn;
    if (cp) S3(cp + 3 * k, B.cp);
    if (fn) S3(fn + 3 * k, B.fn);
    if (ft) S3(ft + 3 * k, B.friction);
  }
  return 0;
}
// the law alone through flat arrays (mirrored by oracle/ref_math.cpp:ref_inner_bond_force_core)
void orc_inner_bond_force_core(double dn, const double* n, double dn0, double weight, double dt, const edb_inner_bond_params* p, double meff,
                               double* en_et /* 2 */, double* tds, double* ft, const double* cp, const double* pi, const double* vi, double* fi,
                               const double* wi, const double* pj, const double* vj, const double* wj) {
  BondParams P{p->kn, p->kt, p->damp_rate, p->mode, p->crit1, p->crit2};
  v3 T = V3(tds), F = V3(ft), f{0, 0, 0};
  inner_bond_force_core(dn, V3(n), dn0, weight, dt, P, meff, en_et[0], T, en_et[1], F, V3(cp), V3(pi), V3(vi), f, V3(wi), V3(pj), V3(vj), V3(wj));
  S3(tds, T); S3(ft, F); S3(fi, f);
}
// the oracle's sort / match key against InteractionPair::operator< and operator== (mirrored by oracle/ref_math.cpp:ref_pair_less/equal)
static Inter mk_inter(const unsigned long long* r) {
  Inter I;
  I.id_i = r[0]; I.id_j = r[1]; I.sub_i = (uint32_t)r[2]; I.sub_j = (uint32_t)r[3]; I.type = (uint16_t)r[4]; I.swap = (uint8_t)r[5];
  return I;
}
int orc_pair_less(const unsigned long long* a, const unsigned long long* b) { return key_of(mk_inter(a)) < key_of(mk_inter(b)) ? 1 : 0; }
int orc_pair_equal(const unsigned long long* a, const unsigned long long* b) { return key_of(mk_inter(a)) == key_of(mk_inter(b)) ? 1 : 0; }
// face areas through the flat shape arrays (mirrored by oracle/ref_math.cpp:ref_face_areas)
void orc_face_areas(int nv, const double* verts, int ne, const int* edges, int nf, const int* faces_stream, int faces_len, double h, double* out) {
  edb_shape sh{};
  sh.nv = nv; sh.ne = ne; sh.nf = nf; sh.faces_len = faces_len; sh.vertices = verts; sh.edges = edges; sh.faces = faces_stream;
  std::vector<Shape> v;
  load_shapes(v, 1, &sh);
  for (int f = 0; f < nf; f++) out[f] = face_area(v[0], f, h);
}
int orc_interface_breaks(int mode, double c1, double c2, const double* en, const double* et, const double* dn, const double* fn3, uint64_t size) {
  std::vector<v3> fn(size);
  for (uint64_t k = 0; k < size; k++) fn[k] = V3(fn3 + 3 * k);
  return interface_breaks(mode, c1, c2, en, et, dn, fn.data(), size) ? 1 : 0;
}
// One step boundary on one driver, composed from the operators in the macro's order (config_numerical_schemes.msp:13-24): the reference
// sequence that the product's fused k_driver_cycle has to reproduce.  sums6 = what the contacts of the step add to forces / moment.
int orc_driver_cycle(edb_driver* d, edb_driver_motion* m, const double* g, const double* sums6, int sphere_mode, double system_mass, double dt,
                     double time_next, int do_end, int do_begin) {
  Sim s;
  edb_driver din = *d;
  if (din.kind == EDB_DRIVER_RSHAPE) { s.mesh_shapes.resize(1); din.shape = 0; }
  const double c0[3] = {d->center[0], d->center[1], d->center[2]};
  if (orc_set_drivers((orc_sim*)&s, 1, &din)) return EDB_ERR_INVALID;
  Driver& D = s.drivers[0];
  if (D.kind == DRV_CYLINDER) D.center = V3(c0);
  orc_set_driver_motions((orc_sim*)&s, 1, m);
  D.forces = V3(m->forces);
  if (D.kind == DRV_SURFACE) { D.cproj = V3(m->mom); D.mom = {0, 0, 0}; }
  if (do_end) {
    reset_force_driver(s);                                   // (before the forces of the step)
    gravity(s, g);                                           // compute_force: gravity_force first (no particles here)
    if (sums6) {                                             // ... then the contacts
      if (m_need_forces(D.motion)) D.forces = D.forces + V3(sums6);
      if (sphere_mode && D.kind == DRV_RSHAPE && rshape_need_moment(D)) D.mom = D.mom + V3(sums6 + 3);
    }
    force_to_accel_driver(s, system_mass);
    push_f_v_driver(s, 0.5 * dt);
  }
  if (do_begin) {
    push_f_v_r_driver(s, time_next, dt);
    push_f_v_driver(s, 0.5 * dt);
    push_av_to_quat_driver(s, time_next, dt);
  }
  const v3 cproj = D.cproj;
  orc_get_drivers((orc_sim*)&s, 1, d, m);
  if (D.kind == DRV_SURFACE) S3(m->mom, cproj);
  return 0;
}
int orc_push_f_v_r_driver(orc_sim* h, double t, double dt) { push_f_v_r_driver(*(Sim*)h, t, dt); return 0; }
int orc_push_f_v_driver(orc_sim* h, double dt) { push_f_v_driver(*(Sim*)h, dt); return 0; }
int orc_push_av_to_quat_driver(orc_sim* h, double t, double dt) { push_av_to_quat_driver(*(Sim*)h, t, dt); return 0; }
int orc_reset_force_driver(orc_sim* h) { reset_force_driver(*(Sim*)h); return 0; }
int orc_force_to_accel_driver(orc_sim* h, double, double system_mass) { force_to_accel_driver(*(Sim*)h, system_mass); return 0; }
int orc_driver_vertices(orc_sim* h) { driver_vertices(*(Sim*)h); return 0; }
int orc_grid_rshape_driver(orc_sim*) { return 0; }   // the oracle enumerates mesh primitives exhaustively: the set is defined by the filters
int orc_backup_drivers(orc_sim* h) { Sim& s = *(Sim*)h; s.drivers_backup = s.drivers; return 0; }
int orc_driver_displ_over(orc_sim* h, double thr, uint64_t* r) {
  Sim& s = *(Sim*)h;
  if (s.drivers_backup.empty()) s.drivers_backup = s.drivers;
  *r = driver_displ_over(s, thr);
  return 0;
}
int orc_set_physical_time(orc_sim* h, double t) { ((Sim*)h)->physical_time = t; return 0; }
int orc_get_physical_time(orc_sim* h, double* t) { *t = ((Sim*)h)->physical_time; return 0; }
int orc_set_system_mass(orc_sim* h, double m) { Sim& s = *(Sim*)h; s.system_mass = m; s.have_system_mass = true; return 0; }

int orc_interaction_counts(orc_sim* h, uint64_t* counts, uint64_t* active) {
  Sim& s = *(Sim*)h;
  for (int t = 0; t < EDB_NTYPES; t++) {
    counts[t] = s.wave[t].size();
    uint64_t a = 0;
    for (uint32_t k : s.wave[t]) if (!(is_zero(s.inter[k].ft) && is_zero(s.inter[k].mom))) a++;
    if (active) active[t] = a;
  }
  return 0;
}
// dem_cost_model.cpp:63-92: inner (non-ghost) cells only; cost = 3 N + 1 + sum of the type weights of the records stored with the
// cell's particles (PartnerGhost copies live in ghost cells and are not seen); no list at all: 3 N + 1
int orc_dem_cost_model(orc_sim* h, int32_t* grid_dims, double* costs, uint64_t cap) {
  Sim& s = *(Sim*)h;
  if (grid_dims) for (int d = 0; d < 3; d++) grid_dims[d] = s.nc[d];
  if (!costs) return 0;
  const size_t ncell = (size_t)s.nc[0] * s.nc[1] * s.nc[2];
  if (cap < ncell) { s.err = "dem_cost_model: cost buffer smaller than the real-cell grid"; return EDB_ERR_CAPACITY; }
  if (s.cstart.empty()) { s.err = "dem_cost_model: no cell grid yet"; return EDB_ERR_STATE; }
  static const double weight[EDB_NTYPES] = {1, 3, 5, 4, 1, 1, 1, 1, 3, 5, 4, 3, 5, 1};
  std::vector<double> per_particle(s.nreal, 0.0);
  const bool skip_interaction = s.inter.empty() && s.bonds.empty();
  if (!skip_interaction) {
    for (const Inter& I : s.inter) {
      if (I.ghost == 2) continue;   // stored in a ghost cell
      const uint32_t o = I.owner();
      if (o < s.nreal) per_particle[o] += weight[I.type];
    }
    for (const Sim::Bond& b : s.bonds)
      if (b.ghost != 2 && b.i < s.nreal) per_particle[b.i] += weight[13];
  }
  for (int k = 1; k <= s.nc[2]; k++)
    for (int j = 1; j <= s.nc[1]; j++)
      for (int i = 1; i <= s.nc[0]; i++) {
        const size_t c = s.cell_of(i, j, k);
        const size_t n = s.cstart[c + 1] - s.cstart[c];
        double cost = 3.0 * (double)n + 1.0;
        for (size_t x = s.cstart[c]; x < s.cstart[c + 1]; x++) cost += per_particle[s.citems[x]];
        costs[(size_t)(i - 1) + (size_t)s.nc[0] * ((size_t)(j - 1) + (size_t)s.nc[1] * (size_t)(k - 1))] = cost;
      }
  return 0;
}
int orc_export_interactions(orc_sim* h, edb_interaction* out, uint64_t cap, uint64_t* n) {
  Sim& s = *(Sim*)h;
  uint64_t k = 0;
  for (int t = 0; t < EDB_NTYPES; t++)
    for (uint32_t idx : s.wave[t]) {
      if (k < cap) {
        const Inter& I = s.inter[idx];
        edb_interaction e;
        memset(&e, 0, sizeof(e));
        e.id_i = I.id_i; e.id_j = I.id_j; e.sub_i = I.sub_i; e.sub_j = I.sub_j; e.type = I.type; e.swap = I.swap; e.ghost = I.ghost;
        e.friction[0] = I.ft.x; e.friction[1] = I.ft.y; e.friction[2] = I.ft.z;
        e.moment[0] = I.mom.x; e.moment[1] = I.mom.y; e.moment[2] = I.mom.z;
        out[k] = e;
      }
      k++;
    }
  *n = k;
  return 0;
}
int orc_import_interactions(orc_sim* h, const edb_interaction* in, uint64_t n) {
  Sim& s = *(Sim*)h;
  s.inter.clear();
  for (uint64_t k = 0; k < n; k++) {
    Inter I;
    I.id_i = in[k].id_i; I.id_j = in[k].id_j; I.sub_i = in[k].sub_i; I.sub_j = in[k].sub_j; I.type = in[k].type; I.swap = in[k].swap; I.ghost = in[k].ghost;
    I.ft = {in[k].friction[0], in[k].friction[1], in[k].friction[2]};
    I.mom = {in[k].moment[0], in[k].moment[1], in[k].moment[2]};
    s.inter.push_back(I);
  }
  for (auto& w : s.wave) w.clear();
  return 0;
}
int orc_get_contact_state(orc_sim* h, int32_t type, double* dn, double* cp, double* fn, double* ft) {
  Sim& s = *(Sim*)h;
  size_t k = 0;
  for (uint32_t idx : s.wave[type]) {
    const Inter& I = s.inter[idx];
    if (dn) dn[k] = I.dn;
    if (cp) { cp[3 * k] = I.cp.x; cp[3 * k + 1] = I.cp.y; cp[3 * k + 2] = I.cp.z; }
    if (fn) { fn[3 * k] = I.fn.x; fn[3 * k + 1] = I.fn.y; fn[3 * k + 2] = I.fn.z; }
    if (ft) { ft[3 * k] = I.ft.x; ft[3 * k + 1] = I.ft.y; ft[3 * k + 2] = I.ft.z; }
    k++;
  }
  return 0;
}
int orc_get_counters(orc_sim* h, edb_counters* c) {
  Sim& s = *(Sim*)h;
  memset(c, 0, sizeof(*c));
  c->interaction_updates = s.interaction_updates; c->particle_steps = s.particle_steps; c->rebuilds = s.rebuilds;
  return 0;
}
int orc_event_record(orc_sim*, int32_t) { return 0; }
int orc_reset_counters(orc_sim* h) { Sim& s = *(Sim*)h; s.interaction_updates = s.particle_steps = s.rebuilds = 0; return 0; }
// world vertices of particle p (testing compute_vertices)
int orc_get_vertices(orc_sim* h, double* out, uint64_t cap_doubles) {
  Sim& s = *(Sim*)h;
  const size_t n = s.nreal * (size_t)s.nvmax;
  for (size_t k = 0; k < n && 3 * k + 2 < cap_doubles; k++) { out[3 * k] = s.verts[k].x; out[3 * k + 1] = s.verts[k].y; out[3 * k + 2] = s.verts[k].z; }
  return s.nvmax;
}

// ---- function-level entry points (pinned against oracle/_ref in tests/test_oracle_vs_ref.py) ----------------
static v3 V(const double* p) { return {p[0], p[1], p[2]}; }
static int packc(const Contact& c, double* o) { o[0] = c.dn; o[1] = c.n.x; o[2] = c.n.y; o[3] = c.n.z; o[4] = c.cp.x; o[5] = c.cp.y; o[6] = c.cp.z; return c.hit; }
int orc_filter_vertex_vertex(double rV, const double* vi, double ri, const double* vj, double rj) { return filter_vv(rV, V(vi), ri, V(vj), rj); }
int orc_detection_vertex_vertex_core(const double* pi, double ri, const double* pj, double rj, double* o) { return packc(detect_vv(V(pi), ri, V(pj), rj), o); }
int orc_filter_vertex_edge_core(double rV, const double* vi, double ri, const double* vf, const double* vs, double rj) { return filter_ve(rV, V(vi), ri, V(vf), V(vs), rj); }
int orc_detection_vertex_edge_core(const double* vi, double ri, const double* vf, const double* vs, double rj, double* o) { return packc(detect_ve(V(vi), ri, V(vf), V(vs), rj), o); }
int orc_filter_edge_edge_core(double rV, const double* a, const double* b, double ri, const double* c, const double* d, double rj) { return filter_ee(rV, V(a), V(b), ri, V(c), V(d), rj); }
int orc_detection_edge_edge_core(const double* a, const double* b, double ri, const double* c, const double* d, double rj, double* o) { return packc(detect_ee(V(a), V(b), ri, V(c), V(d), rj), o); }
int orc_filter_vertex_face(double rV, const double* vi, double ri, int nvj, const double* vaj, double rj, const int* idx, int nf) {
  std::vector<v3> va(nvj);
  for (int k = 0; k < nvj; k++) va[k] = V(vaj + 3 * k);
  return filter_vf(rV, V(vi), ri, va, rj, idx, nf);
}
int orc_detection_vertex_face_core(const double* vi, double ri, int nvj, const double* vaj, double rj, const int* idx, int nf, double* o) {
  std::vector<v3> va(nvj);
  for (int k = 0; k < nvj; k++) va[k] = V(vaj + 3 * k);
  return packc(detect_vf(V(vi), ri, va, rj, idx, nf), o);
}
void orc_get_vertex(const double* lv, const double* p, double hh, const double* q, double* o) {
  const v3 r = V(p) + qrot(q4{q[0], q[1], q[2], q[3]}, hh * V(lv));
  o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
static Obb obbf(const double* a) { return Obb{{a[0], a[1], a[2]}, {a[3], a[4], a[5]}, {a[6], a[7], a[8]}, {a[9], a[10], a[11]}, {a[12], a[13], a[14]}}; }
int orc_obb_intersect(const double* a, const double* b) { return obb_intersect(obbf(a), obbf(b)); }
void orc_compute_obb(const double* in, const double* pos, const double* q, double hh, double enl, double* o) {
  Obb r = obb_world(obbf(in), V(pos), q4{q[0], q[1], q[2], q[3]}, hh);
  obb_enlarge(r, enl);
  const v3* f = &r.c;
  for (int k = 0; k < 5; k++) { o[3 * k] = f[k].x; o[3 * k + 1] = f[k].y; o[3 * k + 2] = f[k].z; }
}
void orc_contact_force_core(int law, double dn, const double* n, double dt, const double* prm, double meff, double reff, double* ft, const double* cp,
                            const double* pi, const double* vi, double* f, double* mom, const double* wi, const double* pj, const double* vj, const double* wj) {
  Params p; p.kn = prm[0]; p.kt = prm[1]; p.kr = prm[2]; p.mu = prm[3]; p.damp = prm[4]; p.fc = prm[5]; p.dncut = prm[6]; p.gamma = prm[7];
  v3 FT = V(ft), F = V(f), M = V(mom);
  contact_force(law, dn, V(n), dt, p, meff, reff, FT, V(cp), V(pi), V(vi), F, M, V(wi), V(pj), V(vj), V(wj));
  ft[0] = FT.x; ft[1] = FT.y; ft[2] = FT.z; f[0] = F.x; f[1] = F.y; f[2] = F.z; mom[0] = M.x; mom[1] = M.y; mom[2] = M.z;
}
void orc_compute_moments(const double* cp, const double* p, const double* f, const double* m, double* o) {
  const v3 r = moments(V(cp), V(p), V(f), V(m));
  o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
static Driver mkdrv(int kind, const double* prm) {
  Driver d; d.kind = kind;
  if (kind == 0) { d.radius = prm[0]; d.axis = V(prm + 1); d.center = V(prm + 4); }
  else if (kind == 1) { d.offset = prm[0]; d.normal = V(prm + 1); d.cproj = d.offset * d.normal; }
  else { d.radius = prm[0]; d.center = V(prm + 1); }
  return d;
}
int orc_driver_filter(int kind, const double* prm, double rcut, const double* p) { return driver_filter(mkdrv(kind, prm), rcut, V(p)); }
int orc_driver_detector(int kind, const double* prm, double rcut, const double* p, double* o) { return packc(driver_detect(mkdrv(kind, prm), rcut, V(p)), o); }

}  // extern "C"
