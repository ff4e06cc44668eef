// kernels_bond_rebuild.cuh — inner bonds across a neighbour rebuild, on the device.
//
// The reference re-resolves its persistent interactions by particle id at every update_particles_full
// (src/polyhedron/update_persistent_interactions.cpp, interaction_manager.hpp:114-160) and drops the broken ones
// (unclassify_interactions.cpp:135-145).  Here the bonds stay in device memory: the particle slots follow the re-binning
// permutation (old2new), a stable radix sort by the new owner slot restores the order (owner slot, partner id, sub_i, sub_j) — the
// order inside an owner does not change —, broken bonds sort to the end and are cut off (those with friction are queued as
// vertex-vertex history, k_bond_convert), and the three derived tables (ignore list of the contact enumeration, per-particle
// incidence lists of the bond gather, interfaces) are rebuilt with flag / scan / scatter passes.  The host only learns four counts
// and the permutation of the static bond fields it mirrors for edb_export_inner_bonds.
#pragma once
#include "kernels_bond.cuh"

namespace edb {

struct BondArrays {   // one complete set of per-bond arrays (two sets: the rebuild permutes from one into the other)
  uint2 *ij, *sub;
  ulonglong2* ids;    // (id_i, id_j)
  double4 *par, *fr, *tds, *fd, *cp;
  uint8_t *mode, *unbroken, *ghost;
};

// key = new owner slot of the bonds that survive, 0xffffffff for the broken ones; cnt[0] += broken, and the broken bonds that carry
// friction are appended to conv (cnt[1] = their number)
__global__ void __launch_bounds__(BLK) k_bond_rekey(BondArrays B, uint32_t n, const uint32_t* __restrict__ old2new, uint32_t* __restrict__ key,
                                                    uint32_t* __restrict__ idx, uint32_t* __restrict__ cnt, BondConv* __restrict__ conv) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint2 ij = B.ij[k];
  const uint32_t owner = ij.x == 0xffffffffu ? 0xffffffffu : old2new[ij.x];
  idx[k] = k;
  if (B.unbroken[k]) {
    key[k] = owner == 0xffffffffu ? 0xfffffffeu : owner;   // (an unresolved owner keeps its bond, after every resolved one)
    return;
  }
  key[k] = 0xffffffffu;
  atomicAdd(cnt, 1u);
  const double4 fr = B.fr[k];
  if ((fr.x != 0.0 || fr.y != 0.0 || fr.z != 0.0) && owner != 0xffffffffu) {   // Interaction::active (interaction.hpp:129-136)
    const uint2 sb = B.sub[k];
    conv[atomicAdd(cnt + 1, 1u)] = BondConv{owner, sb.x, sb.y, (long long)B.ids[k].y, fr.x, fr.y, fr.z};
  }
}
// new set <- old set through the sorted permutation; the wrench of the old evaluation does not survive the rebuild
__global__ void __launch_bounds__(BLK) k_bond_permute(BondArrays D, BondArrays S, uint32_t n_keep, const uint32_t* __restrict__ key,
                                                      const uint32_t* __restrict__ idx, const uint32_t* __restrict__ old2new) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_keep) return;
  const uint32_t s = idx[k];
  const uint2 ij = S.ij[s];
  D.ij[k] = make_uint2(key[k] == 0xfffffffeu ? 0xffffffffu : key[k], ij.y == 0xffffffffu ? 0xffffffffu : old2new[ij.y]);
  D.sub[k] = S.sub[s]; D.ids[k] = S.ids[s]; D.par[k] = S.par[s]; D.fr[k] = S.fr[s]; D.tds[k] = S.tds[s];
  D.fd[k] = make_double4(0, 0, 0, 0); D.cp[k] = make_double4(0, 0, 0, 0);
  D.mode[k] = S.mode[s]; D.unbroken[k] = 1; D.ghost[k] = S.ghost[s];
}
// flags of the three derived tables, one pass over the bonds (sorted):
//   ign[k]   first bond of a (owner, partner id) pair with both slots resolved -> one entry of the owner's ignore list; cnt_ign[owner]++
//   itf[k]   first bond of an interface: maximal run with the same (id_i, id_j), PartnerGhost records (ghost == 2) never start one
//            (update_interfaces.cpp:44-84)
//   inc keys particle of entry 2k (side i) and 2k+1 (side j), 0xffffffff when the bond is not resolved on both sides
__global__ void __launch_bounds__(BLK) k_bond_flags(BondArrays B, uint32_t n, uint32_t* __restrict__ ign, uint32_t* __restrict__ cnt_ign,
                                                    uint32_t* __restrict__ itf, uint32_t* __restrict__ inc_key, uint32_t* __restrict__ inc_val,
                                                    uint32_t* __restrict__ n_inc_valid) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint2 ij = B.ij[k];
  const ulonglong2 id = B.ids[k];
  const bool both = ij.x != 0xffffffffu && ij.y != 0xffffffffu;
  bool first_pair = both, first_itf = B.ghost[k] != 2;
  if (k > 0) {
    const uint2 pj = B.ij[k - 1];
    const ulonglong2 pid = B.ids[k - 1];
    const bool prev_both = pj.x != 0xffffffffu && pj.y != 0xffffffffu;
    if (prev_both && pj.x == ij.x && pid.y == id.y) first_pair = false;
    if (pid.x == id.x && pid.y == id.y && B.ghost[k - 1] != 2) first_itf = false;
  }
  ign[k] = first_pair ? 1u : 0u;
  if (first_pair) atomicAdd(cnt_ign + ij.x, 1u);
  itf[k] = first_itf ? 1u : 0u;
  inc_key[2 * k] = both ? ij.x : 0xffffffffu;
  inc_key[2 * k + 1] = both ? ij.y : 0xffffffffu;
  inc_val[2 * k] = 2 * k;
  inc_val[2 * k + 1] = 2 * k + 1;
  if (both) atomicAdd(n_inc_valid, 2u);
}
// scatter of the flagged bonds: ignore ids in flag order (= by owner, then partner id: the CSR of the per-owner counts lines up with
// it), interface starts
__global__ void __launch_bounds__(BLK) k_bond_scatter(BondArrays B, uint32_t n, const uint32_t* __restrict__ ign, const uint32_t* __restrict__ ign_rank,
                                                      unsigned long long* __restrict__ ign_ids, const uint32_t* __restrict__ itf,
                                                      const uint32_t* __restrict__ itf_rank, uint2* __restrict__ itf_out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (ign[k]) ign_ids[ign_rank[k]] = B.ids[k].y;
  if (itf[k]) itf_out[itf_rank[k]].x = k;
}
__global__ void __launch_bounds__(BLK) k_itf_lengths(uint2* __restrict__ itf, uint32_t nitf, uint32_t n_bonds) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nitf) return;
  const uint32_t a = itf[t].x, b = t + 1 < nitf ? itf[t + 1].x : n_bonds;
  itf[t].y = b - a;
}
// incidence lists from the (particle, entry) pairs sorted by particle: run starts -> plist / inc_off
__global__ void __launch_bounds__(BLK) k_bond_inc_flags(const uint32_t* __restrict__ key, uint32_t n_valid, uint32_t* __restrict__ start) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_valid) return;
  start[e] = (e == 0 || key[e - 1] != key[e]) ? 1u : 0u;
}
__global__ void __launch_bounds__(BLK) k_bond_inc_lists(const uint32_t* __restrict__ key, const uint32_t* __restrict__ start, const uint32_t* __restrict__ rank,
                                                        uint32_t n_valid, uint32_t* __restrict__ plist, uint32_t* __restrict__ inc_off) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > n_valid) return;
  if (e == n_valid) { inc_off[rank[n_valid]] = n_valid; return; }   // rank[n_valid] = number of lists (total of the scan)
  if (start[e]) { plist[rank[e]] = key[e]; inc_off[rank[e]] = e; }
}

}  // namespace edb
