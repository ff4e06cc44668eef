// kernels_rebuild.cuh — neighbour / interaction rebuild on the device (runs every 10-100 steps).
//   cell binning (counting sort by cell, id-ordered inside a cell) -> ghost layer -> candidate enumeration
//   (count + scan + fill, owner-sorted per interaction type) -> friction/moment history carry-over -> partner
//   incidence lists for the deterministic force gather.
// Reference behaviour reproduced: src/grid/chunk_neighbors_contact.cpp:46-57, src/sphere/nbh_sphere.cu:64-309,
// src/polyhedron/nbh_polyhedron.cpp:75-384, polyhedron/nbh_polyhedron_driver.hpp:205-238,
// src/interaction/include/exaDEM/interaction/interaction_manager.hpp:88-112 + placeholder_interaction.hpp:398-408.
#pragma once
#include "common.cuh"

namespace edb {

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan (uint32) : reduce tiles -> scan tile sums (one block) -> apply
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256;
constexpr int SCAN_I = 8;
constexpr int SCAN_TILE = SCAN_T * SCAN_I;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}
// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t ws[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t inc = warp_incl_scan(v);
  if (lane == 31) ws[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = (lane < (blockDim.x >> 5)) ? ws[lane] : 0;
    s = warp_incl_scan(s);
    ws[lane] = s;
  }
  __syncthreads();
  const uint32_t base = wid ? ws[wid - 1] : 0;
  *total = ws[(blockDim.x >> 5) - 1];
  __syncthreads();
  return base + inc - v;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_reduce(const uint32_t* __restrict__ in, uint32_t* __restrict__ bsum, uint32_t n) {
  const uint32_t base = blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_I; k++) {
    const uint32_t i = base + k * SCAN_T + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_bsum(uint32_t* __restrict__ bsum, uint32_t nb, uint32_t* __restrict__ total) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nb; base += SCAN_T) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? bsum[i] : 0;
    uint32_t tot;
    const uint32_t ex = block_excl_scan(v, &tot);
    if (i < nb) bsum[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total = carry;
}
// out has n+1 entries (out[n] = total); in-place allowed
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const uint32_t* in, uint32_t* out, const uint32_t* __restrict__ bsum,
                                                       const uint32_t* __restrict__ total, uint32_t n) {
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_I;
  uint32_t v[SCAN_I], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_I; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_I; k++) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = *total;
}

// ---------------------------------------------------------------------------------------------------------------
// cell binning
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(const GridDev& G, double x, int d) {
  int c = (int)floor((x - G.lo[d]) / G.cs[d]);
  c = c < 0 ? 0 : c;
  c = c > G.nc[d] - 1 ? G.nc[d] - 1 : c;
  return c;
}
__device__ __forceinline__ uint32_t cell_lin(const GridDev& G, int i, int j, int k) {
  return (uint32_t)i + (uint32_t)G.g[0] * ((uint32_t)j + (uint32_t)G.g[1] * (uint32_t)k);
}
// wrap periodic coordinates into the global box, assign the (ghost-shelled) cell index, histogram
// dest (multi-GPU only): >= 0 marks a particle that migrated away — its slot is dead and is dropped by the sort
constexpr uint32_t DEAD_CELL = 0xffffffffu;
__global__ void k_cell_assign(double4* __restrict__ pr, uint32_t n, GridDev G, uint32_t* __restrict__ cidx,
                              uint32_t* __restrict__ ccount, const int* __restrict__ dest) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (dest && dest[p] >= 0) {
    cidx[p] = DEAD_CELL;
    return;
  }
  double4 r = pr[p];
  double c[3] = {r.x, r.y, r.z};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (G.periodic[d]) {
      const double L = G.ghi[d] - G.glo[d];
      if (c[d] >= G.ghi[d]) c[d] -= L;
      else if (c[d] < G.glo[d]) c[d] += L;
    }
  }
  r.x = c[0]; r.y = c[1]; r.z = c[2];
  pr[p] = r;
  const uint32_t lin = cell_lin(G, cell_coord(G, c[0], 0) + 1, cell_coord(G, c[1], 1) + 1, cell_coord(G, c[2], 2) + 1);
  cidx[p] = lin;
  atomicAdd(&ccount[lin], 1u);
}
__global__ void k_cell_scatter(const uint32_t* __restrict__ cidx, uint32_t n, const uint32_t* __restrict__ cstart,
                               uint32_t* __restrict__ cfill, uint32_t* __restrict__ tmp) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t c = cidx[p];
  if (c == DEAD_CELL) return;
  tmp[cstart[c] + atomicAdd(&cfill[c], 1u)] = p;
}
// deterministic order inside a cell: rank by particle id (ids are unique)
__global__ void k_cell_rank(const uint32_t* __restrict__ cidx, const uint64_t* __restrict__ id, uint32_t n,
                            const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ tmp,
                            uint32_t* __restrict__ new2old, uint32_t* __restrict__ old2new) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t c = cidx[p];
  if (c == DEAD_CELL) {
    old2new[p] = DEAD_CELL;
    return;
  }
  const uint64_t me = id[p];
  uint32_t rank = 0;
  for (uint32_t s = cstart[c]; s < cstart[c + 1]; s++) rank += (id[tmp[s]] < me) ? 1u : 0u;
  const uint32_t np = cstart[c] + rank;
  new2old[np] = p;
  old2new[p] = np;
}
__global__ void k_permute(PView dst, PView src, const uint32_t* __restrict__ new2old, const uint32_t* __restrict__ cidx_old,
                          uint32_t* __restrict__ cidx_new, uint32_t n) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t o = new2old[s];
  dst.pr[s] = src.pr[o]; dst.vm[s] = src.vm[o]; dst.wg[s] = src.wg[o]; dst.q[s] = src.q[o];
  dst.fx[s] = src.fx[o]; dst.fy[s] = src.fy[o]; dst.fz[s] = src.fz[o];
  dst.mx[s] = src.mx[o]; dst.my[s] = src.my[o]; dst.mz[s] = src.mz[o];
  dst.ax[s] = src.ax[o]; dst.ay[s] = src.ay[o]; dst.az[s] = src.az[o];
  dst.ix[s] = src.ix[o]; dst.iy[s] = src.iy[o]; dst.iz[s] = src.iz[o];
  dst.h[s] = src.h[o]; dst.id[s] = src.id[o]; dst.orig[s] = src.orig[o];
  cidx_new[s] = cidx_old[o];
}
__global__ void k_cell_ranges(const uint32_t* __restrict__ cstart, uint32_t ncell, uint32_t* __restrict__ cbeg,
                              uint32_t* __restrict__ cend) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  cbeg[c] = cstart[c];
  cend[c] = cstart[c + 1];
}

// ---------------------------------------------------------------------------------------------------------------
// ghost layer (periodic images of the local brick; remote bricks are handled by the exchange in comm.cuh)
// ghost table entry k: ghost cell gcell[k] mirrors local real cell gsrc_cell[k] shifted by gshift[k]
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_ghost_count(const uint32_t* __restrict__ gsrc_cell, uint32_t nent, const uint32_t* __restrict__ cstart,
                              uint32_t* __restrict__ gcount) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nent) return;
  const uint32_t c = gsrc_cell[k];
  gcount[k] = cstart[c + 1] - cstart[c];
}
__global__ void k_ghost_fill(const uint32_t* __restrict__ gcell, const uint32_t* __restrict__ gsrc_cell, uint32_t nent,
                             const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ goff, uint32_t base,
                             uint32_t* __restrict__ gsrc, uint32_t* __restrict__ gent, uint32_t* __restrict__ cbeg,
                             uint32_t* __restrict__ cend) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nent) return;
  const uint32_t c = gsrc_cell[k];
  const uint32_t s0 = cstart[c], cnt = cstart[c + 1] - s0, o = goff[k];
  for (uint32_t t = 0; t < cnt; t++) {
    gsrc[o + t] = s0 + t;
    gent[o + t] = k;
  }
  cbeg[gcell[k]] = base + o;
  cend[gcell[k]] = base + o + cnt;
}
// ghost refresh.  ALL=false: the per-step field list of ghost_update_rq (src/grid/update_ghosts.cu:46-51):
// r (+shift), v, vrot, orient (+ type/radius/mass which ride in the same 32-byte sectors).
template <bool ALL>
__global__ void k_ghost_refresh(PView P, const uint32_t* __restrict__ gsrc, const uint32_t* __restrict__ gent,
                                const double* __restrict__ gshift, uint32_t nreal, uint32_t nghost, uint32_t* __restrict__ cidx,
                                const uint32_t* __restrict__ gcell) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const uint32_t s = gsrc[g], k = gent[g], d = nreal + g;
  double4 r = ld4(P.pr + s);
  r.x = r.x + gshift[3 * k]; r.y = r.y + gshift[3 * k + 1]; r.z = r.z + gshift[3 * k + 2];
  st4(P.pr + d, r);
  st4(P.vm + d, ld4(P.vm + s));
  st4(P.wg + d, ld4(P.wg + s));
  st4(P.q + d, ld4(P.q + s));
  if (ALL) {
    P.h[d] = P.h[s];
    P.id[d] = P.id[s];
    cidx[d] = gcell[k];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// candidate enumeration
// ---------------------------------------------------------------------------------------------------------------
struct CntTab {
  uint32_t* c[NT];  // per type: count (count pass) / CSR offsets (fill pass) per owner slot, nullptr = type inactive
};
struct NbhArgs {
  PView P;
  const uint32_t* cidx;
  const uint32_t* cbeg;
  const uint32_t* cend;
  GridDev G;
  const Driver* drv;
  int ndrv;
  double rcut_inc;
  int symetric;
  uint32_t nreal;
  const double4* vert;  // polyhedra: world vertices
  ShapesView S;
  MeshView M;
  int skip_pp;   // polyhedra: the particle-particle waves are enumerated by k_nbh_poly_cell (kernels_nbh_poly.cuh)
  // InteractionManager::ignore_ (interaction_manager.hpp:114-160): per owner slot, the ids of the partners it holds an unbroken inner bond
  // with — no regular contact is enumerated between the two while the bond holds.  CSR over the real slots; nullptr without bonds.
  const uint32_t* ign_off;
  const unsigned long long* ign_ids;
  // spheres: the count pass also RECORDS what it finds — candidate s of particle a at pp_tmp[s * nreal + a] = (partner slot, flag
  // byte), up to PP_TMP_CAP per particle — so that the fill pass is a copy (k_nbh_sphere_copy) instead of a second walk over the
  // 27-cell stencil; a particle with more candidates raises *pp_overflow and the fill pass enumerates again.
  uint2* pp_tmp;
  uint32_t* pp_overflow;
  // largest pr.w over reals + ghosts (device scalar, k_max_radius): lets for_each_neighbor skip the stencil cells that lie wholly
  // beyond ra.w + max radius + rcut_inc of the particle (nullptr: every cell of the stencil is visited)
  const double* rmax_all;
};
__global__ void k_max_radius(const double4* __restrict__ pr, uint32_t n, unsigned long long* __restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  double r = p < n ? pr[p].w : 0.0;
  for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
  if ((threadIdx.x & 31) == 0 && r > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(r));   // positive doubles order like their bits
}
constexpr uint32_t PP_TMP_CAP = 16;
__device__ __forceinline__ bool bond_ignored(const NbhArgs& A, uint32_t owner, unsigned long long partner_id) {
  if (!A.ign_off) return false;
  for (uint32_t q = A.ign_off[owner]; q < A.ign_off[owner + 1]; q++)
    if (A.ign_ids[q] == partner_id) return true;
  return false;
}

__device__ __forceinline__ void cell_ijk(const GridDev& G, uint32_t lin, int& i, int& j, int& k) {
  i = lin % G.g[0];
  const uint32_t t = lin / G.g[0];
  j = t % G.g[1];
  k = t / G.g[1];
}
__device__ __forceinline__ bool is_ghost_cell(const GridDev& G, int i, int j, int k) {
  return i == 0 || j == 0 || k == 0 || i == G.g[0] - 1 || j == G.g[1] - 1 || k == G.g[2] - 1;
}

// Visit the 3x3x3 stencil of particle a in the order (dk, dj, di) x cell order, applying ContactNeighborFilterFunc
// (chunk_neighbors_contact.cpp:46-57).  f(b, ghost_cell)
template <class F>
__device__ __forceinline__ void for_each_neighbor(const NbhArgs& A, uint32_t a, const double4 ra, F&& f) {
  int ci, cj, ck;
  cell_ijk(A.G, A.cidx[a], ci, cj, ck);
  // Distance from the particle to the face it shares with the neighbour cell in each direction (cell c of the shelled grid covers
  // [lo + (c-1) cs, lo + c cs)): a stencil cell whose nearest point is farther than the largest possible filter radius holds no
  // candidate.  The visiting order of the remaining cells is unchanged, and so is the order of the records.  With cells of 2.5 and a
  // reach of 1.4 (the drum) a particle needs 9.5 of its 27 cells on average — but the 32 particles of a warp sit all over their cell and
  // the warp walks the union, so the 1 M-sphere build stays at 1.8 ms (measured); it pays for sparse or strongly polydisperse sets.
  double reach2 = 1.7976931348623157e308, gx[3], gy[3], gz[3];
  if (A.rmax_all) {
    const double reach = (ra.w + *A.rmax_all + A.rcut_inc) * 1.000000001;
    reach2 = reach * reach;
  }
  gx[1] = gy[1] = gz[1] = 0.0;
  gx[0] = fmax(ra.x - (A.G.lo[0] + (ci - 1) * A.G.cs[0]), 0.0); gx[2] = fmax((A.G.lo[0] + ci * A.G.cs[0]) - ra.x, 0.0);
  gy[0] = fmax(ra.y - (A.G.lo[1] + (cj - 1) * A.G.cs[1]), 0.0); gy[2] = fmax((A.G.lo[1] + cj * A.G.cs[1]) - ra.y, 0.0);
  gz[0] = fmax(ra.z - (A.G.lo[2] + (ck - 1) * A.G.cs[2]), 0.0); gz[2] = fmax((A.G.lo[2] + ck * A.G.cs[2]) - ra.z, 0.0);
  for (int dk = -1; dk <= 1; dk++)
    for (int dj = -1; dj <= 1; dj++)
      for (int di = -1; di <= 1; di++) {
        const int i = ci + di, j = cj + dj, k = ck + dk;
        if (i < 0 || j < 0 || k < 0 || i >= A.G.g[0] || j >= A.G.g[1] || k >= A.G.g[2]) continue;
        if (gx[di + 1] * gx[di + 1] + gy[dj + 1] * gy[dj + 1] + gz[dk + 1] * gz[dk + 1] > reach2) continue;
        const uint32_t cb = cell_lin(A.G, i, j, k);
        const bool gc = is_ghost_cell(A.G, i, j, k);
        const uint32_t e = A.cend[cb];
        for (uint32_t b = A.cbeg[cb]; b < e; b++) {
          if (b == a) continue;
          const double4 rb = ldg4(A.P.pr + b);
          const v3 d = mk3(rb.x, rb.y, rb.z) - mk3(ra.x, ra.y, ra.z);
          const double d2 = dot(d, d);
          const double rmax = ra.w + rb.w + A.rcut_inc;
          if (d2 > 0.0 && d2 < rmax * rmax) f(b, gc);
        }
      }
}

struct VA {  // world-vertex accessor of one particle
  const double4* p;
  __device__ __forceinline__ v3 operator[](int i) const {
    const double4 t = ldg4(p + i);
    return mk3(t.x, t.y, t.z);
  }
};

// FILL=false: count candidates per type;  FILL=true: write records at the owner's CSR offset
struct WTab {
  WView w[NT];
};

template <bool FILL>
__global__ void __launch_bounds__(128) k_nbh_sphere(NbhArgs A, CntTab C, WTab W) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.nreal) return;
  const double4 ra = A.P.pr[a];
  const uint64_t ida = A.P.id[a];
  const v3 pa = mk3(ra.x, ra.y, ra.z);
  // drivers first (nbh_sphere.cu:182-231)
  uint32_t nd[3] = {0, 0, 0};
  for (int d = 0; d < A.ndrv; d++) {
    const Driver D = A.drv[d];
    if (D.kind > 2) continue;
    if (driver_filter(D, ra.w + A.rcut_inc, pa)) {
      const int t = 4 + D.kind;
      if (FILL) {
        const uint32_t k = C.c[t][a] + nd[D.kind];
        W.w[t].ij[k] = make_uint2(a, (uint32_t)d);
        W.w[t].flg[k] = 0;
        W.w[t].ft[k] = make_double4(0, 0, 0, __longlong_as_double((long long)d));
        W.w[t].mo[k] = make_double4(0, 0, 0, 0);
      }
      nd[D.kind]++;
    }
  }
  // RShape (mesh) drivers: detection_sphere_driver (sphere/nbh_sphere.hpp:25-80) over the per-cell primitive lists
  uint32_t nm[3] = {0, 0, 0};
  if (A.M.nmesh) {
    const uint32_t cell = A.cidx[a];
    for (int d = 0; d < A.ndrv; d++) {
      const Driver D = A.drv[d];
      if (D.kind != 3) continue;
      const MeshDev md = A.M.md[D.shape];
      const VA mv{A.M.verts + md.v0};
      const uint32_t* off = A.M.cell_off + ((size_t)D.shape * A.M.ncell + cell) * 3;
      for (int kind = 0; kind < 3; kind++) {
        const int t = 7 + kind;
        for (uint32_t s = off[kind]; s < off[kind + 1]; s++) {
          const int j = A.M.cell_data[s];
          bool ok;
          if (kind == 0) ok = filter_vv(A.rcut_inc, pa, ra.w, mv[j], md.radius);
          else if (kind == 1) ok = filter_ve(A.rcut_inc, pa, ra.w, mv[A.M.edges[2 * (md.e0 + j)]], mv[A.M.edges[2 * (md.e0 + j) + 1]], md.radius);
          else ok = filter_vf(A.rcut_inc, pa, ra.w, mv, md.radius, A.M.fidx + md.fi0 + A.M.foff[md.f0 + j], A.M.fn[md.f0 + j]);
          if (!ok) continue;
          if (FILL) {
            const uint32_t k = C.c[t][a] + nm[kind];
            W.w[t].ij[k] = make_uint2(a, (uint32_t)d);
            W.w[t].sub[k] = make_uint2(0u, (uint32_t)j);
            W.w[t].flg[k] = 0;
            W.w[t].ft[k] = make_double4(0, 0, 0, __longlong_as_double((long long)d));
            W.w[t].mo[k] = make_double4(0, 0, 0, 0);
          }
          nm[kind]++;
        }
      }
    }
  }
  uint32_t n0 = 0;
  const uint32_t base0 = FILL ? C.c[0][a] : 0;
  if (!A.skip_pp) for_each_neighbor(A, a, ra, [&](uint32_t b, bool gc) {
    const uint64_t idb = A.P.id[b];
    uint8_t ghost = 0;
    if (A.symetric) {  // nbh_sphere.cu:245-249
      if (ida >= idb && !gc) return;
    } else {           // :272-279
      if (ida >= idb) return;
      if (gc) ghost = 1;
    }
    if (!FILL && A.pp_tmp && n0 < PP_TMP_CAP) A.pp_tmp[(size_t)n0 * A.nreal + a] = make_uint2(b, (uint32_t)(ghost << 1));
    if (FILL) {
      const uint32_t k = base0 + n0;
      W.w[0].ij[k] = make_uint2(a, b);
      W.w[0].flg[k] = (uint8_t)(ghost << 1);
      W.w[0].ft[k] = make_double4(0, 0, 0, __longlong_as_double((long long)idb));
      W.w[0].mo[k] = make_double4(0, 0, 0, 0);
    }
    n0++;
  });
  if (!FILL) {
    if (!A.skip_pp) {
      C.c[0][a] = n0;
      if (A.pp_tmp && n0 > PP_TMP_CAP) *A.pp_overflow = 1u;
    }
    if (C.c[4]) C.c[4][a] = nd[0];
    if (C.c[5]) C.c[5][a] = nd[1];
    if (C.c[6]) C.c[6][a] = nd[2];
    if (C.c[7]) C.c[7][a] = nm[0];
    if (C.c[8]) C.c[8][a] = nm[1];
    if (C.c[9]) C.c[9][a] = nm[2];
  }
}

// Count pass of wave 0 with a WARP per particle (spheres): the lanes take 32 consecutive candidates of a stencil cell at a time — one
// coalesced 1 KB load instead of 32 dependent trips through a per-thread loop —, the cell-level reach test is warp-uniform (one
// particle), so 9.5 instead of 27 cells are visited, and a ballot keeps the accepted candidates in the serial loop's order
// (cell order, then slot order): same records, same order, same pp_tmp contents as k_nbh_sphere<false>.
// Measured (drum, 1.05 M spheres): 2.40 ms per rebuild against 1.78 ms for the thread-per-particle count: a cell holds ~4 spheres,
// so 28 of 32 lanes idle per stencil cell. Kept behind EDB_NBH_WARP (parity-green) for dense cells; off by default.
__global__ void __launch_bounds__(256) k_nbh_sphere_warp(NbhArgs A, uint32_t* __restrict__ cnt0) {
  const int lane = threadIdx.x & 31;
  const uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (a >= A.nreal) return;
  const double4 ra = A.P.pr[a];
  const uint64_t ida = A.P.id[a];
  int ci, cj, ck;
  cell_ijk(A.G, A.cidx[a], ci, cj, ck);
  double reach2 = 1.7976931348623157e308, gx[3], gy[3], gz[3];
  if (A.rmax_all) {
    const double reach = (ra.w + *A.rmax_all + A.rcut_inc) * 1.000000001;
    reach2 = reach * reach;
  }
  gx[1] = gy[1] = gz[1] = 0.0;
  gx[0] = fmax(ra.x - (A.G.lo[0] + (ci - 1) * A.G.cs[0]), 0.0); gx[2] = fmax((A.G.lo[0] + ci * A.G.cs[0]) - ra.x, 0.0);
  gy[0] = fmax(ra.y - (A.G.lo[1] + (cj - 1) * A.G.cs[1]), 0.0); gy[2] = fmax((A.G.lo[1] + cj * A.G.cs[1]) - ra.y, 0.0);
  gz[0] = fmax(ra.z - (A.G.lo[2] + (ck - 1) * A.G.cs[2]), 0.0); gz[2] = fmax((A.G.lo[2] + ck * A.G.cs[2]) - ra.z, 0.0);
  uint32_t n0 = 0;
  for (int dk = -1; dk <= 1; dk++)
    for (int dj = -1; dj <= 1; dj++)
      for (int di = -1; di <= 1; di++) {
        const int i = ci + di, j = cj + dj, k = ck + dk;
        if (i < 0 || j < 0 || k < 0 || i >= A.G.g[0] || j >= A.G.g[1] || k >= A.G.g[2]) continue;
        if (gx[di + 1] * gx[di + 1] + gy[dj + 1] * gy[dj + 1] + gz[dk + 1] * gz[dk + 1] > reach2) continue;
        const uint32_t cb = cell_lin(A.G, i, j, k);
        const bool gc = is_ghost_cell(A.G, i, j, k);
        const uint32_t e = A.cend[cb];
        for (uint32_t b0 = A.cbeg[cb]; b0 < e; b0 += 32) {
          const uint32_t b = b0 + lane;
          bool ok = false;
          uint8_t ghost = 0;
          if (b < e && b != a) {
            const double4 rb = ldg4(A.P.pr + b);
            const v3 d = mk3(rb.x, rb.y, rb.z) - mk3(ra.x, ra.y, ra.z);
            const double d2 = dot(d, d);
            const double rmax = ra.w + rb.w + A.rcut_inc;
            if (d2 > 0.0 && d2 < rmax * rmax) {   // ContactNeighborFilterFunc (chunk_neighbors_contact.cpp:46-57)
              const uint64_t idb = A.P.id[b];
              if (A.symetric) ok = !(ida >= idb && !gc);           // nbh_sphere.cu:245-249
              else { ok = ida < idb; ghost = gc ? 1 : 0; }          // :272-279
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok) {
            const uint32_t pos = n0 + __popc(m & ((1u << lane) - 1u));
            if (A.pp_tmp && pos < PP_TMP_CAP) A.pp_tmp[(size_t)pos * A.nreal + a] = make_uint2(b, (uint32_t)(ghost << 1));
          }
          n0 += __popc(m);
        }
      }
  if (lane == 0) {
    cnt0[a] = n0;
    if (A.pp_tmp && n0 > PP_TMP_CAP) *A.pp_overflow = 1u;
  }
}

// fill pass of wave 0 from the candidates recorded by the count pass: off = the scanned counts (CSR by owner)
__global__ void __launch_bounds__(BLK) k_nbh_sphere_copy(PView P, const uint2* __restrict__ tmp, const uint32_t* __restrict__ off, uint32_t n, WView w) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const uint32_t k0 = off[a], cnt = off[a + 1] - k0;
  for (uint32_t s = 0; s < cnt; s++) {
    const uint2 e = tmp[(size_t)s * n + a];
    const uint32_t k = k0 + s;
    w.ij[k] = make_uint2(a, e.x);
    w.flg[k] = (uint8_t)e.y;
    w.ft[k] = make_double4(0, 0, 0, __longlong_as_double((long long)P.id[e.x]));
    w.mo[k] = make_double4(0, 0, 0, 0);
  }
}

template <bool FILL>
__device__ __forceinline__ void emit(const CntTab& C, const WTab& W, uint32_t owner, uint32_t* cnt, int t, uint32_t i, uint32_t j,
                                     uint32_t si, uint32_t sj, uint8_t flg, uint64_t partner_id) {
  if (FILL) {
    const uint32_t k = C.c[t][owner] + cnt[t];
    W.w[t].ij[k] = make_uint2(i, j);
    W.w[t].sub[k] = make_uint2(si, sj);
    W.w[t].flg[k] = flg;
    W.w[t].ft[k] = make_double4(0, 0, 0, __longlong_as_double((long long)partner_id));
    W.w[t].mo[k] = make_double4(0, 0, 0, 0);
  }
  cnt[t]++;
}

template <bool FILL>
__global__ void __launch_bounds__(64) k_nbh_poly(NbhArgs A, CntTab C, WTab W) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.nreal) return;
  const double4 ra = A.P.pr[a];
  const uint64_t ida = A.P.id[a];
  const double4 wga = A.P.wg[a];
  const uint32_t ta = (uint32_t)((unsigned long long)__double_as_longlong(wga.w) & 0xffffffffu);
  const ShapeDev sa = A.S.sh[ta];
  const double ha = A.P.h[a], rV = A.rcut_inc;
  const VA va{A.vert + (size_t)a * A.S.nvmax};
  const double rma = ha * sa.radius;
  uint32_t cnt[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // drivers (nbh_polyhedron_driver.hpp:205-238): one record per vertex with sub_j = -1
  for (int d = 0; d < A.ndrv; d++) {
    const Driver D = A.drv[d];
    if (D.kind > 2) continue;
    const int t = 4 + D.kind;
    for (int k = 0; k < sa.nv; k++)
      if (driver_filter(D, rV + rma, va[k])) emit<FILL>(C, W, a, cnt, t, a, (uint32_t)d, (uint32_t)k, 0xffffffffu, 0, (uint64_t)d);
  }
  const int* ea = A.S.edges + 2 * sa.e0;
  // RShape (mesh) drivers: add_driver_interaction (nbh_polyhedron_driver.hpp:61-201), types 7..12
  if (A.M.nmesh) {
    const uint32_t cell = A.cidx[a];
    for (int d = 0; d < A.ndrv; d++) {
      const Driver D = A.drv[d];
      if (D.kind != 3) continue;
      const MeshDev md = A.M.md[D.shape];
      const VA mv{A.M.verts + md.v0};
      const uint32_t* off = A.M.cell_off + ((size_t)D.shape * A.M.ncell + cell) * 3;
      const int* me = A.M.edges + 2 * md.e0;
      const double rd = md.radius;
      for (int i = 0; i < sa.nv; i++) {
        const v3 vi = va[i];
        for (uint32_t s = off[0]; s < off[1]; s++) {
          const int j = A.M.cell_data[s];
          if (filter_vv(rV, vi, rma, mv[j], rd)) emit<FILL>(C, W, a, cnt, 7, a, (uint32_t)d, (uint32_t)i, (uint32_t)j, 0, (uint64_t)d);
        }
        for (uint32_t s = off[1]; s < off[2]; s++) {
          const int j = A.M.cell_data[s];
          if (filter_ve(rV, vi, rma, mv[me[2 * j]], mv[me[2 * j + 1]], rd)) emit<FILL>(C, W, a, cnt, 8, a, (uint32_t)d, (uint32_t)i, (uint32_t)j, 0, (uint64_t)d);
        }
        for (uint32_t s = off[2]; s < off[3]; s++) {
          const int j = A.M.cell_data[s];
          if (filter_vf(rV, vi, rma, mv, rd, A.M.fidx + md.fi0 + A.M.foff[md.f0 + j], A.M.fn[md.f0 + j]))
            emit<FILL>(C, W, a, cnt, 9, a, (uint32_t)d, (uint32_t)i, (uint32_t)j, 0, (uint64_t)d);
        }
      }
      for (int i = 0; i < sa.ne; i++)
        for (uint32_t s = off[1]; s < off[2]; s++) {
          const int j = A.M.cell_data[s];
          if (filter_ee(rV, va[ea[2 * i]], va[ea[2 * i + 1]], rma, mv[me[2 * j]], mv[me[2 * j + 1]], rd))
            emit<FILL>(C, W, a, cnt, 10, a, (uint32_t)d, (uint32_t)i, (uint32_t)j, 0, (uint64_t)d);
        }
      for (uint32_t s = off[0]; s < off[1]; s++) {
        const int j = A.M.cell_data[s];
        const v3 vj = mv[j];
        for (int i = 0; i < sa.ne; i++)
          if (filter_ve(rV, vj, rd, va[ea[2 * i]], va[ea[2 * i + 1]], rma)) emit<FILL>(C, W, a, cnt, 11, a, (uint32_t)d, (uint32_t)i, (uint32_t)j, 0, (uint64_t)d);
        for (int i = 0; i < sa.nf; i++)
          if (filter_vf(rV, vj, rd, va, rma, A.S.fidx + sa.fi0 + A.S.foff[sa.f0 + i], A.S.fn[sa.f0 + i]))
            emit<FILL>(C, W, a, cnt, 12, a, (uint32_t)d, (uint32_t)i, (uint32_t)j, 0, (uint64_t)d);
      }
    }
  }
  const double4 qa4 = A.P.q[a];
  const q4 qa{qa4.x, qa4.y, qa4.z, qa4.w};
  if (!A.skip_pp) for_each_neighbor(A, a, ra, [&](uint32_t b, bool gc) {
    const uint64_t idb = A.P.id[b];
    if (ida >= idb) return;  // nbh_polyhedron.cpp:234-236
    if (bond_ignored(A, a, idb)) return;   // InteractionManager::skip_ignored_interactions
    const uint8_t ghost = gc ? 1 : 0;
    const double4 rb = A.P.pr[b];
    const double4 wgb = A.P.wg[b];
    const uint32_t tb = (uint32_t)((unsigned long long)__double_as_longlong(wgb.w) & 0xffffffffu);
    const ShapeDev sb = A.S.sh[tb];
    const double hb = A.P.h[b];
    const double4 qb4 = A.P.q[b];
    const q4 qb{qb4.x, qb4.y, qb4.z, qb4.w};
    Obb oi = obb_world(sa.obb, mk3(ra.x, ra.y, ra.z), qa, ha), oj = obb_world(sb.obb, mk3(rb.x, rb.y, rb.z), qb, hb);
    obb_enlarge(oi, 0.5 * rV); obb_enlarge(oj, 0.5 * rV);
    if (!obb_intersect(oi, oj)) return;
    obb_enlarge(oi, 0.5 * rV); obb_enlarge(oj, 0.5 * rV);
    const VA vb{A.vert + (size_t)b * A.S.nvmax};
    const double rmb = hb * sb.radius;
    const int* eb = A.S.edges + 2 * sb.e0;
    // A -> B (swap = 0)
    const uint8_t f0 = (uint8_t)(ghost << 1);
    for (int i = 0; i < sa.nv; i++) {
      const v3 vi = va[i];
      if (!obb_intersect(oj, obb_point(vi, rma))) continue;
      for (int j = 0; j < sb.nv; j++)
        if (filter_vv(rV, vi, rma, vb[j], rmb)) emit<FILL>(C, W, a, cnt, 0, a, b, i, j, f0, idb);
      for (int j = 0; j < sb.ne; j++)
        if (filter_ve(rV, vi, rma, vb[eb[2 * j]], vb[eb[2 * j + 1]], rmb)) emit<FILL>(C, W, a, cnt, 1, a, b, i, j, f0, idb);
      for (int j = 0; j < sb.nf; j++)
        if (filter_vf(rV, vi, rma, vb, rmb, A.S.fidx + sb.fi0 + A.S.foff[sb.f0 + j], A.S.fn[sb.f0 + j])) emit<FILL>(C, W, a, cnt, 2, a, b, i, j, f0, idb);
    }
    for (int i = 0; i < sa.ne; i++)
      for (int j = 0; j < sb.ne; j++)
        if (filter_ee(rV, va[ea[2 * i]], va[ea[2 * i + 1]], rma, vb[eb[2 * j]], vb[eb[2 * j + 1]], rmb)) emit<FILL>(C, W, a, cnt, 3, a, b, i, j, f0, idb);
    // B -> A (swap = 1): pi = b, pj = a; the owner stays a (interaction_pair.hpp:75-90)
    const uint8_t f1 = (uint8_t)(1 | (ghost << 1));
    for (int j = 0; j < sb.nv; j++) {
      const v3 vj = vb[j];
      if (!obb_intersect(oi, obb_point(vj, rmb))) continue;
      for (int i = 0; i < sa.ne; i++)
        if (filter_ve(rV, vj, rmb, va[ea[2 * i]], va[ea[2 * i + 1]], rma)) emit<FILL>(C, W, a, cnt, 1, b, a, j, i, f1, idb);
      for (int i = 0; i < sa.nf; i++)
        if (filter_vf(rV, vj, rmb, va, rma, A.S.fidx + sa.fi0 + A.S.foff[sa.f0 + i], A.S.fn[sa.f0 + i])) emit<FILL>(C, W, a, cnt, 2, b, a, j, i, f1, idb);
    }
  });
  if (!FILL) {
    for (int t = A.skip_pp ? 4 : 0; t < 13; t++)
      if (C.c[t]) C.c[t][a] = cnt[t];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// history carry-over: new record -> same key in the previous owner segment (key = partner id, subs, swap; the
// type is implied by the wave, the owner by the segment).  interaction_pair.hpp:204-216, placeholder_interaction.hpp:398-408
// ---------------------------------------------------------------------------------------------------------------
// Immigrants (multi-GPU migration) have no old owner segment: their history arrives as a CSR table by immigrant index.
struct ImmRec {  // = MigRec of kernels_migrate.cuh
  uint32_t m;
  uint8_t type, flg;
  uint16_t pad;
  uint32_t sub_i, sub_j;
  long long pid;
  double ft[3], mo[3];
};
struct ImmHist {
  const uint32_t* off;  // n_imm + 1
  const ImmRec* rec;
  uint32_t n_imm;
};
__global__ void k_history(WView nw, WView ow, const uint32_t* __restrict__ new2old, uint32_t n_old_real, bool has_sub, int type,
                          ImmHist H) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nw.n) return;
  const uint8_t fl = nw.flg[k];
  const uint2 ij = nw.ij[k];
  const uint32_t owner = (fl & 1) ? ij.y : ij.x;
  const uint32_t oo = new2old[owner];
  const double4 ft = nw.ft[k];
  const long long pid = __double_as_longlong(ft.w);
  uint2 sb = make_uint2(0, 0);
  if (has_sub) sb = nw.sub[k];
  if (oo >= n_old_real) {
    const uint32_t m = oo - n_old_real;
    if (m >= H.n_imm) return;
    const uint32_t e = H.off[m + 1];
    for (uint32_t s = H.off[m]; s < e; s++) {
      const ImmRec& r = H.rec[s];
      if (r.type != type || r.pid != pid || (r.flg & 1) != (fl & 1)) continue;
      if (has_sub && (r.sub_i != sb.x || r.sub_j != sb.y)) continue;
      nw.ft[k] = make_double4(r.ft[0], r.ft[1], r.ft[2], ft.w);
      nw.mo[k] = make_double4(r.mo[0], r.mo[1], r.mo[2], 0.0);
      nw.flg[k] = fl | 8 | ((r.mo[0] != 0.0 || r.mo[1] != 0.0 || r.mo[2] != 0.0) ? 16 : 0);  // FLG_HOT | FLG_MOM (kernels_step.cuh)
      return;
    }
    return;
  }
  if (ow.n == 0) return;
  const uint32_t e = ow.off[oo + 1];
  for (uint32_t s = ow.off[oo]; s < e; s++) {
    const double4 oft = ow.ft[s];
    if (__double_as_longlong(oft.w) != pid) continue;
    const uint8_t ofl = ow.flg[s];
    if ((ofl & 1) != (fl & 1)) continue;
    if (((ofl >> 1) & 3) == 2) continue;  // PartnerGhost records are not history sources (interaction_pair.hpp:108-109)
    if (has_sub) {
      const uint2 osb = ow.sub[s];
      if (osb.x != sb.x || osb.y != sb.y) continue;
    }
    if (!(ofl & 8)) return;  // the old record was not in contact: no history (its payload is zero)
    nw.ft[k] = make_double4(oft.x, oft.y, oft.z, ft.w);
    if (ofl & 16) nw.mo[k] = ow.mo[s];
    *nw.fd(k) = make_double4(0, 0, 0, 0);  // the wrench of the old evaluation does not survive the rebuild
    *nw.cp(k) = make_double4(0, 0, 0, 0);
    nw.flg[k] = fl | 8 | (ofl & 16);
    return;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// incidence lists: for every real particle, ALL records it takes part in — as owner (any wave) or as partner (particle-
// particle waves) — as codes (position of the wave in the operator's wave order << 28 | record index), sorted ascending, which
// is the order in which the reference's serial wave-by-wave loop would add their wrenches to the particle.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t INC_SHIFT = 28;
constexpr uint32_t INC_MASK = (1u << INC_SHIFT) - 1;
struct OrderTab {
  int order[NT];
  int norder;
};

// owner entries.  FILL=false: cnt[p] = number of owned records;  FILL=true: writes them at the head of the particle's list and
// leaves cnt[p] = that number (the cursor at which k_incidence<true> appends the partner entries).
template <bool FILL>
__global__ void k_inc_owner(WTab W, OrderTab O, uint32_t nreal, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ inc_off,
                            uint32_t* __restrict__ inc) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nreal) return;
  uint32_t s = 0;
  const uint32_t base = FILL ? inc_off[p] : 0;
  for (int o = 0; o < O.norder; o++) {
    const WView& w = W.w[O.order[o]];
    if (w.n == 0) continue;
    const uint32_t a = w.off[p], b = w.off[p + 1];
    if (FILL)
      for (uint32_t k = a; k < b; k++) inc[base + s + (k - a)] = ((uint32_t)o << INC_SHIFT) | k;
    s += b - a;
  }
  cnt[p] = s;
}
// partner entries of the particle-particle wave at order position `o`
template <bool FILL>
__global__ void k_incidence(WView w, int o, uint32_t nreal, uint32_t* __restrict__ cnt_or_fill,
                            const uint32_t* __restrict__ inc_off, uint32_t* __restrict__ inc) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= w.n) return;
  const uint2 ij = w.ij[k];
  const uint32_t partner = (w.flg[k] & 1) ? ij.x : ij.y;
  if (partner >= nreal) return;  // forces on ghost copies are discarded (quirk Q6)
  const uint32_t s = atomicAdd(&cnt_or_fill[partner], 1u);
  if (FILL) inc[inc_off[partner] + s] = ((uint32_t)o << INC_SHIFT) | k;
}
__global__ void k_incidence_sort(const uint32_t* __restrict__ inc_off, uint32_t* __restrict__ inc, uint32_t nreal) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nreal) return;
  const uint32_t b = inc_off[p], e = inc_off[p + 1];
  for (uint32_t i = b + 1; i < e; i++) {
    const uint32_t v = inc[i];
    uint32_t j = i;
    while (j > b && inc[j - 1] > v) {
      inc[j] = inc[j - 1];
      j--;
    }
    inc[j] = v;
  }
}

__global__ void k_backup(const double4* __restrict__ pr, const double4* __restrict__ q, double4* __restrict__ pr_b,
                         double4* __restrict__ q_b, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  pr_b[p] = pr[p];
  if (q_b) q_b[p] = q[p];
}

}  // namespace edb
