// kernels_migrate.cuh — rebuild-time particle migration between the bricks of the GPUs of one box (replaces the MPI path
// of migrate_cell_particles_interaction, src/interaction/migrate_cell_particles_interaction.cpp:25-32: particles carry
// their interaction records — here only the ACTIVE ones, i.e. the friction/moment history, interaction.hpp:129-136).
//   classify (wrap + destination rank) -> count history -> [host: count all-gather] -> pack -> NCCL all-to-all -> unpack
//   -> immigrant history table (CSR by immigrant) consulted by k_history.
#pragma once
#include "common.cuh"
#include "kernels_rebuild.cuh"

namespace edb {

struct MigPart {  // 240 B: the full DEMFieldSet of one particle
  double4 pr, vm, wg, q;
  double f[3], m[3], a[3], I[3], h;
  uint64_t id;
};
struct MigRec {  // 72 B: one active interaction record of a migrating owner
  uint32_t m;    // sender: index of the owner in the sender's list for this destination; receiver: immigrant index
  uint8_t type, flg;
  uint16_t pad;
  uint32_t sub_i, sub_j;
  long long pid;  // partner id (driver index for driver waves)
  double ft[3], mo[3];
};
static_assert(sizeof(MigPart) == 240 && sizeof(MigRec) == 72, "migration wire format");

constexpr int MAX_CUTS = 17;   // pgrid[d] <= 16
struct BrickMap {
  double glo[3], ghi[3], len[3];  // global box, brick length per axis (uniform bricks)
  double cut[3][MAX_CUTS];        // edb_set_brick_cuts: ncut[d] = pg[d] + 1 positions, 0 = uniform along d
  int ncut[3];
  int pg[3], periodic[3];
  int me;
};

__host__ __device__ __forceinline__ int brick_of(const BrickMap& B, const double c[3]) {
  int r = 0, mult = 1;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    int k;
    if (B.ncut[d]) {   // last cut <= c (bricks are [cut_k, cut_k+1))
      k = 0;
      for (int q = 1; q < B.ncut[d] - 1; q++) k += c[d] >= B.cut[d][q] ? 1 : 0;
    } else {
      k = (int)floor((c[d] - B.glo[d]) / B.len[d]);
    }
    k = k < 0 ? 0 : (k > B.pg[d] - 1 ? B.pg[d] - 1 : k);
    r += mult * k;
    mult *= B.pg[d];
  }
  return r;
}

// dest[p] = destination rank of a leaver, -1 for a particle that stays; midx[p] = its index in the list for that rank
__global__ void k_mig_classify(double4* __restrict__ pr, uint32_t n, BrickMap B, int* __restrict__ dest, uint32_t* __restrict__ midx,
                               uint32_t* __restrict__ cnt_part) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double4 r = pr[p];
  double c[3] = {r.x, r.y, r.z};
#pragma unroll
  for (int d = 0; d < 3; d++)
    if (B.periodic[d]) {
      const double L = B.ghi[d] - B.glo[d];
      if (c[d] >= B.ghi[d]) c[d] -= L;
      else if (c[d] < B.glo[d]) c[d] += L;
    }
  r.x = c[0]; r.y = c[1]; r.z = c[2];
  pr[p] = r;
  const int b = brick_of(B, c);
  if (b == B.me) {
    dest[p] = -1;
  } else {
    dest[p] = b;
    midx[p] = atomicAdd(&cnt_part[b], 1u);
  }
}

__device__ __forceinline__ bool rec_active(const double4& ft, const double4& mo, uint8_t flg) {
  if (((flg >> 1) & 3) == 2) return false;  // PartnerGhost copies are not history (interaction_pair.hpp:108-109)
  return ft.x != 0.0 || ft.y != 0.0 || ft.z != 0.0 || mo.x != 0.0 || mo.y != 0.0 || mo.z != 0.0;
}

struct WTabC {
  WView w[NT];
};

// PACK=false: count active records per destination; PACK=true: write MigPart + MigRec into the per-destination ranges
template <bool PACK>
__global__ void k_mig_pack(PView P, uint32_t n, const int* __restrict__ dest, const uint32_t* __restrict__ midx, WTabC W, bool has_sub,
                           uint32_t* __restrict__ cnt_hist, const uint32_t* __restrict__ part_off, const uint32_t* __restrict__ hist_off,
                           MigPart* __restrict__ out_part, MigRec* __restrict__ out_hist) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int d = dest[p];
  if (d < 0) return;
  if (PACK) {
    MigPart o;
    o.pr = P.pr[p]; o.vm = P.vm[p]; o.wg = P.wg[p]; o.q = P.q[p];
    o.f[0] = P.fx[p]; o.f[1] = P.fy[p]; o.f[2] = P.fz[p];
    o.m[0] = P.mx[p]; o.m[1] = P.my[p]; o.m[2] = P.mz[p];
    o.a[0] = P.ax[p]; o.a[1] = P.ay[p]; o.a[2] = P.az[p];
    o.I[0] = P.ix[p]; o.I[1] = P.iy[p]; o.I[2] = P.iz[p];
    o.h = P.h[p];
    o.id = P.id[p];
    out_part[part_off[d] + midx[p]] = o;
  }
  for (int t = 0; t < NT; t++) {
    const WView& w = W.w[t];
    if (w.n == 0 || w.off == nullptr) continue;
    const uint32_t e = w.off[p + 1];
    for (uint32_t k = w.off[p]; k < e; k++) {
      const uint8_t fl = w.flg[k];
      if (!(fl & 8)) continue;  // not FLG_HOT: no history
      const double4 ft = w.ft[k], mo = (fl & 16) ? w.mo[k] : make_double4(0, 0, 0, 0);
      if (!rec_active(ft, mo, fl)) continue;
      const uint32_t s = atomicAdd(&cnt_hist[d], 1u);
      if (PACK) {
        MigRec r;
        r.m = midx[p]; r.type = (uint8_t)t; r.flg = fl; r.pad = 0;
        uint2 sb = make_uint2(0, 0);
        if (w.sub && (has_sub || t >= 7)) sb = w.sub[k];
        r.sub_i = sb.x; r.sub_j = sb.y;
        r.pid = __double_as_longlong(ft.w);
        r.ft[0] = ft.x; r.ft[1] = ft.y; r.ft[2] = ft.z;
        r.mo[0] = mo.x; r.mo[1] = mo.y; r.mo[2] = mo.z;
        out_hist[hist_off[d] + s] = r;
      }
    }
  }
}

// immigrants land at slots [base, base + n)
__global__ void k_mig_unpack(PView P, uint32_t base, uint32_t n, const MigPart* __restrict__ in, int* __restrict__ dest) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const MigPart o = in[i];
  const uint32_t p = base + i;
  P.pr[p] = o.pr; P.vm[p] = o.vm; P.wg[p] = o.wg; P.q[p] = o.q;
  P.fx[p] = o.f[0]; P.fy[p] = o.f[1]; P.fz[p] = o.f[2];
  P.mx[p] = o.m[0]; P.my[p] = o.m[1]; P.mz[p] = o.m[2];
  P.ax[p] = o.a[0]; P.ay[p] = o.a[1]; P.az[p] = o.a[2];
  P.ix[p] = o.I[0]; P.iy[p] = o.I[1]; P.iz[p] = o.I[2];
  P.h[p] = o.h;
  P.id[p] = o.id;
  P.orig[p] = p;
  dest[p] = -1;
}

// immigrant history table: records [r0, r0+n) came from one source whose immigrants start at index m0
template <bool FILL>
__global__ void k_imm_hist(MigRec* __restrict__ rec, uint32_t r0, uint32_t n, uint32_t m0, uint32_t* __restrict__ cnt_or_cur,
                           const uint32_t* __restrict__ ioff, MigRec* __restrict__ table) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  MigRec r = rec[r0 + i];
  const uint32_t m = r.m + (FILL ? 0u : m0);
  if (!FILL) {
    rec[r0 + i].m = m;  // globalise once (count pass)
    atomicAdd(&cnt_or_cur[m], 1u);
  } else {
    table[ioff[m] + atomicAdd(&cnt_or_cur[m], 1u)] = r;
  }
}

__global__ void k_iota(uint32_t* __restrict__ a, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

}  // namespace edb
