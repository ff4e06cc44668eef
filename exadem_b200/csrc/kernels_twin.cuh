// kernels_twin.cuh — update_interaction_ghost on the device (src/interaction/update_interaction_ghost.cpp:58-285).
//
// Polyhedron pairs (and sphere pairs with symetric:false) that straddle a sub-domain or periodic boundary are owned by the
// lower-id side only and flagged OwnerGhost there (nbh_polyhedron.cpp:237-239, nbh_sphere.cu:272-279); the force on the ghost
// partner is discarded.  The reference then ships every OwnerGhost record, payload included, to the sub-domain (another rank,
// or the same rank through a periodic image) where the partner is a REAL particle; there the owner is located in the ghost
// layer by id, the partner by id in the neighbouring non-ghost cells (:206-266), and the record is stored as PartnerGhost.
// It is evaluated like any other record, so the partner receives its share of the wrench from that copy ("twin").
//
// Device scheme: count per destination -> [host: sizes, all ranks exchange counts] -> order-preserving pack (flag, scan,
// fill) -> NCCL all-to-all (remote) / in-place (periodic images of the own brick) -> resolve: partner = real slot named by
// the message, owner = the one ghost copy with the owner's id among the ghost cells around the partner's cell (unique when a
// periodic axis spans >= 3 cells) -> per-type order-preserving append behind the owner-sorted part of the wave
// (index >= off[nreal]): the gather reaches twins through the partner incidence lists only.
#pragma once
#include "common.cuh"
#include "kernels_rebuild.cuh"

namespace edb {

struct TwinMsg {  // 72 B wire format
  uint64_t id_owner;
  uint32_t partner;  // receiver-side reference: real slot (own brick) or index in the receiver's send list to the sender
  uint32_t sub_i, sub_j;
  uint8_t type, flg;  // flg: bit0 swap, bit3 HOT, bit4 MOM
  uint16_t pad;
  double ft[3], mo[3];
};
static_assert(sizeof(TwinMsg) == 72, "twin wire format");

constexpr int TWIN_MAXD = 28;  // own brick + up to 26 neighbouring ranks
struct TwinDest {
  int n;
  uint32_t lo[TWIN_MAXD], hi[TWIN_MAXD];  // ghost-slot range fed by destination d (d = 0: periodic images of the own brick)
};

__device__ __forceinline__ bool twin_source(const WView& w, uint32_t k, uint32_t& owner, uint32_t& partner, uint8_t& fl) {
  fl = w.flg[k];
  if (((fl >> 1) & 3) != 1) return false;  // OwnerGhost only
  const uint2 ij = w.ij[k];
  owner = (fl & 1) ? ij.y : ij.x;
  partner = (fl & 1) ? ij.x : ij.y;
  return true;
}

__global__ void k_twin_count(WView w, TwinDest D, uint32_t* __restrict__ cnt) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= w.n) return;
  uint32_t a, b;
  uint8_t fl;
  if (!twin_source(w, k, a, b, fl)) return;
  for (int d = 0; d < D.n; d++)
    if (b >= D.lo[d] && b < D.hi[d]) {
      atomicAdd(&cnt[d], 1u);
      return;
    }
}
__global__ void k_twin_flag(WView w, uint32_t lo, uint32_t hi, uint32_t* __restrict__ flag) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= w.n) return;
  uint32_t a, b;
  uint8_t fl;
  flag[k] = (twin_source(w, k, a, b, fl) && b >= lo && b < hi) ? 1u : 0u;
}
// g_src != nullptr: destination is the own brick, the partner reference is the real slot the ghost mirrors
__global__ void k_twin_pack(WView w, int type, bool has_sub, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ pos,
                            const uint64_t* __restrict__ id, const uint32_t* __restrict__ g_src, uint32_t nreal,
                            TwinMsg* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= w.n) return;
  uint32_t a, b;
  uint8_t fl;
  if (!twin_source(w, k, a, b, fl) || b < lo || b >= hi) return;
  TwinMsg m;
  m.id_owner = id[a];
  m.partner = g_src ? g_src[b - nreal] : b - lo;
  uint2 sb = make_uint2(0, 0);
  if (has_sub) sb = w.sub[k];
  m.sub_i = sb.x; m.sub_j = sb.y;
  m.type = (uint8_t)type;
  m.flg = fl & (uint8_t)(1 | 8 | 16);
  m.pad = 0;
  const double4 ft = (fl & 8) ? w.ft[k] : make_double4(0, 0, 0, 0);
  const double4 mo = (fl & 16) ? w.mo[k] : make_double4(0, 0, 0, 0);
  m.ft[0] = ft.x; m.ft[1] = ft.y; m.ft[2] = ft.z;
  m.mo[0] = mo.x; m.mo[1] = mo.y; m.mo[2] = mo.z;
  out[pos[k]] = m;
}

// own[m] = ghost slot holding the owner's copy next to the partner (0xffffffff: none — "not all particles are copied in
// ghost cells", update_interaction_ghost.cpp:229-231); par[m] = real slot of the partner
__global__ void k_twin_resolve(const TwinMsg* __restrict__ msg, uint32_t n, const uint32_t* __restrict__ smap, uint32_t nsmap,
                               const uint32_t* __restrict__ cidx, const uint32_t* __restrict__ cbeg, const uint32_t* __restrict__ cend,
                               GridDev G, const uint64_t* __restrict__ id, uint32_t nreal, uint32_t* __restrict__ own,
                               uint32_t* __restrict__ par) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const TwinMsg t = msg[m];
  uint32_t found = 0xffffffffu, bp = 0xffffffffu;
  if (smap ? t.partner < nsmap : t.partner < nreal) {
    bp = smap ? smap[t.partner] : t.partner;
    int ci, cj, ck;
    cell_ijk(G, cidx[bp], ci, cj, ck);
    for (int dk = -1; dk <= 1; dk++)
      for (int dj = -1; dj <= 1; dj++)
        for (int di = -1; di <= 1; di++) {
          const int i = ci + di, j = cj + dj, k = ck + dk;
          if (i < 0 || j < 0 || k < 0 || i >= G.g[0] || j >= G.g[1] || k >= G.g[2]) continue;
          if (!is_ghost_cell(G, i, j, k)) continue;
          const uint32_t c = cell_lin(G, i, j, k);
          for (uint32_t s = cbeg[c]; s < cend[c]; s++)
            if (id[s] == t.id_owner) found = s;
        }
  }
  own[m] = found;
  par[m] = bp;
}
__global__ void k_twin_tflag(const TwinMsg* __restrict__ msg, const uint32_t* __restrict__ own, uint32_t n, int type,
                             uint32_t* __restrict__ flag) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  flag[m] = (msg[m].type == type && own[m] != 0xffffffffu) ? 1u : 0u;
}
__global__ void k_twin_append(const TwinMsg* __restrict__ msg, const uint32_t* __restrict__ own, const uint32_t* __restrict__ par,
                              uint32_t n, int type, const uint32_t* __restrict__ pos, WView w, uint32_t base, bool has_sub,
                              const uint64_t* __restrict__ id) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const TwinMsg t = msg[m];
  if (t.type != type || own[m] == 0xffffffffu) return;
  const uint32_t k = base + pos[m], g = own[m], bp = par[m];
  w.ij[k] = (t.flg & 1) ? make_uint2(bp, g) : make_uint2(g, bp);
  if (has_sub) w.sub[k] = make_uint2(t.sub_i, t.sub_j);
  w.flg[k] = (uint8_t)((t.flg & (1 | 8 | 16)) | (2u << 1));  // PartnerGhost
  w.ft[k] = make_double4(t.ft[0], t.ft[1], t.ft[2], __longlong_as_double((long long)id[bp]));
  w.mo[k] = make_double4(t.mo[0], t.mo[1], t.mo[2], 0.0);
  *w.fd(k) = make_double4(0, 0, 0, 0);
  *w.cp(k) = make_double4(0, 0, 0, 0);
}

}  // namespace edb
