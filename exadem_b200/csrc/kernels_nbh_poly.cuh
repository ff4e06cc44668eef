// kernels_nbh_poly.cuh — polyhedron candidate enumeration (nbh_polyhedron, src/polyhedron/nbh_polyhedron.cpp:75-384), cell-block
// version for sm_100a.  This is the expensive kernel of a polyhedron neighbour rebuild: ~420 filter evaluations (V-V, V-E, V-F,
// E-E in both directions) per pair of octahedra whose enlarged OBBs intersect.
//
//   block  = one real cell of the grid; 16 warps; the warps take the cell's owner particles from a shared counter.
//            (8 warps x 2 blocks per SM with orientations / ids read from global memory measured 13 % slower.)
//   stage  = the world vertices, positions+radius and orientations of every particle of the 3x3x3 cell stencil are copied into
//            shared memory by the TMA engine: one cp.async.bulk per (cell, array) — cell-sorted particles make each cell's block
//            contiguous — completing on one mbarrier.  ids, homothety, shape type and slot follow with ordinary loads.
//   warp   = one owner particle of the cell at a time.  Lanes scan the stencil 32 candidates at a time (distance + id-order
//            filter of chunk_neighbors_contact.cpp:46-57 and nbh_polyhedron.cpp:234-236, then the 15-axis OBB test, :268-276);
//            the survivors are taken in order (ballot), and for each pair the lanes SHARE the pair's filter tests: test index =
//            base + lane, ballot + popc give every accepted test its place in the owner's record segment, so the records come out
//            in exactly the order of the reference's nested loops (the gather's summation order depends on it).
//   passes = the COUNT pass counts per (owner, wave) and records what it found — per pair a header and the 32-bit ballot of every
//            round of tests — as a word stream in a chunked pool; after the scan the FILL pass (k_nbh_poly_replay, warp per owner)
//            replays the streams and writes the records without evaluating a single filter.  (First rebuild, or pool overflow:
//            the fill pass repeats the tests instead, MODE 2.)
// All arithmetic is the same dem_math.cuh functions on the same inputs as the per-thread kernel (k_nbh_poly), so membership is
// bit-identical; only the work distribution changes.
#pragma once
#include "kernels_rebuild.cuh"

namespace edb {

#ifndef EDB_NPC_THREADS
#define EDB_NPC_THREADS 512
#endif
#ifndef EDB_NPC_MINB
#define EDB_NPC_MINB 1
#endif
#ifndef EDB_NPC_PARTS
#define EDB_NPC_PARTS 1
#endif
#ifndef EDB_NPC_NI
#define EDB_NPC_NI 0
#endif
// EDB_NPC_HOIST bit 0: the owner's world OBB is computed once per owner and kept in shared memory (every OBB round and every
// vertex-of-b prefilter round used to rebuild it per lane: four quaternion rotations); bit 1: the candidates' world OBBs of an OBB
// round stay in the registers of the lane that tested them and the vertex-of-a prefilter rounds fetch them by shuffle instead of
// rebuilding them per lane.  Same values from the same functions on the same inputs: the candidate set is unchanged.
// EDB_NPC_MERGE: the two vertex-edge groups of a pair (a's vertices on b's edges, b's vertices on a's edges) share one test index
// space, and so do the two vertex-face groups: both directions go to the same wave and are adjacent there, so the records and their
// order do not change, but two octahedra need one vertex-face round instead of two (16 + 16 tests) and the kernel holds one inlined
// copy of each filter instead of two.
#ifndef EDB_NPC_MERGE
#define EDB_NPC_MERGE 1
#endif
// EDB_NPC_VECULL (needs EDB_NPC_MERGE): the vertex-edge tests run over the edges within reach only, like the edge-edge tests.
// filter_ve accepts (vertex v of a, edge e of b) only if v is within R = ri + rj + rV of the point of e's LINE at a parameter in
// [-rV, 1 + rV], i.e. within R + rV |e| of the segment; v lies within bnd_a of a's centre, so an edge of b farther than
// bnd_a + R + rV |e| from a's centre cannot pass against any vertex of a.  Same margin as the edge-edge cull; ascending edge order kept.
// Parity-green but SLOWER: 17.4 ms per build against 15.8 (profiles/r2_sweep_poly_vecull.log) — the merged group already fits two
// rounds, and the n-th-set-bit decode of the edge index in every test costs more than the tests it saves.  Off.
#ifndef EDB_NPC_VECULL
#define EDB_NPC_VECULL 0
#endif
// EDB_NPC_TAB: the kept vertices / edges of a pair (set bits of mA, mB, eA, eB and of the vertex-edge reach masks) are written once
// per pair into small per-warp tables, ascending, so that a test decodes its vertex or edge with one shared-memory load instead of an
// n-th-set-bit loop.  EDB_NPC_FDIV: test index -> (row, column) by a float reciprocal (exact for these sizes: indices < 2^10, divisors
// <= 32, the quotient is at least 1/64 away from the next integer) instead of a 32-bit integer division per test.
#ifndef EDB_NPC_TAB
#define EDB_NPC_TAB 1
#endif
#ifndef EDB_NPC_FDIV
#define EDB_NPC_FDIV 1
#endif
#ifndef EDB_NPC_HOIST
#define EDB_NPC_HOIST 1
#endif
constexpr int NPC_THREADS = EDB_NPC_THREADS;
// EDB_NPC_PARTS > 1: a warp's work item in the count pass is 1/NPC_PARTS of an owner's stencil (a contiguous range of 32-candidate
// chunks) instead of the whole owner; every part writes its own stream, the counts are added per owner, the replay walks the parts
// in order, so the records and their order are unchanged (parity-green at 4).  The idea was load balance — a cell holds 8 / 12 / 18 /
// 27 octahedra on 16 warps, ncu shows 11 of 16 resident warps active on average —, the measurement says no: 19.6 / 23.6 / 29.8 /
// 36.2 / 40.7 ms per build at 1 / 2 / 4 / 8 / 14 parts (profiles/r2_sweep_poly_parts.log): every part pays its own OBB round and
// prefilter rounds on a handful of lanes, and that costs more than the idle warps did.  Default 1.
constexpr uint32_t NPC_PARTS = EDB_NPC_PARTS;
constexpr int CHW = 64;                       // words per pool chunk; word 0 links to the next chunk
constexpr uint32_t NPC_NONE = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct SV {  // world-vertex accessor over the staged (shared-memory) copy
  const double4* p;
  __device__ __forceinline__ v3 operator[](int i) const {
    const double4 t = p[i];
    return mk3(t.x, t.y, t.z);
  }
};

__device__ __forceinline__ uint32_t type_of_bits(double w) { return (uint32_t)((unsigned long long)__double_as_longlong(w) & 0xffffffffu); }
__host__ __device__ inline size_t npc_bytes_per_particle(int nvmax) { return (size_t)nvmax * 32 + 32 + 32 + 8 + 8 + 4 + 4 + 1; }

// Conservative reach test used to cull edge-edge filter evaluations: squared distance from point c to the segment p0-p1.
// filter_ee accepts a pair of edges only if two points ON the segments are within R = ri + rj + rV of each other (ee_closest returns
// parameters strictly inside (0,1)), and every point of b's edges lies within bnd_b of b's centre, so an edge of a whose distance to
// b's centre exceeds bnd_b + R cannot pass against ANY edge of b.  The cull compares with a relative + absolute safety margin many
// orders above the rounding of either computation, so it never removes a test the reference would have accepted: the candidate set
// is unchanged, only the number of evaluated edge pairs drops (octahedra in a packing: ~16-25 of 144).
__device__ __forceinline__ double seg_point_d2(v3 p0, v3 p1, v3 c) {
  const v3 d = p1 - p0, w = c - p0;
  const double dd = dot(d, d);
  double t = dd > 0.0 ? dot(w, d) / dd : 0.0;
  t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  const v3 q = w - d * t;
  return dot(q, q);
}
__device__ __forceinline__ double reach2(double bnd, double R, v3 c) {
  const double m = fmax(fmax(fabs(c.x), fabs(c.y)), fmax(fabs(c.z), 1.0));
  const double r = (bnd + R) * 1.000001 + 1e-9 * m;
  return r * r;
}

// index of the n-th (0-based) set bit of m.  __fns is a ~40-instruction library routine and was 10 % of the instructions of the
// neighbour build (profiles/r2_ncu_polynbh_lines.txt); the masks here have a handful of bits set.
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t m, uint32_t n) {
  for (uint32_t k = 0; k < n; k++) m &= m - 1;
  return (uint32_t)__ffs((int)m) - 1u;
}

// EDB_NPC_NI (bit 0: OBB test, bit 1: filters): out-of-line copies of the big inlined bodies, to shrink the kernel's code (ncu: 16 % of
// the warp samples are instruction-fetch stalls on 6700 SASS instructions)
__device__ __noinline__ bool obb_intersect_ni(const Obb& a, const Obb& b) { return obb_intersect(a, b); }
__device__ __noinline__ bool filter_ve_ni(double rV, v3 vi, double ri, v3 vf, v3 vs, double rj) { return filter_ve(rV, vi, ri, vf, vs, rj); }
__device__ __noinline__ bool filter_ee_ni(double rV, v3 fi, v3 si, double ri, v3 fj, v3 sj, double rj) { return filter_ee(rV, fi, si, ri, fj, sj, rj); }
__device__ __noinline__ bool filter_vf_ni(double rV, v3 vi, double ri, const double4* vb, double rj, const int* idx, int nf) {
  return filter_vf(rV, vi, ri, SV{vb}, rj, idx, nf);
}
#if EDB_NPC_NI & 1
#define NPC_OBB_INTERSECT obb_intersect_ni
#else
#define NPC_OBB_INTERSECT obb_intersect
#endif
#if EDB_NPC_NI & 2
#define NPC_FILTER_VE filter_ve_ni
#define NPC_FILTER_EE filter_ee_ni
#define NPC_FILTER_VF(rV, vi, ri, vb, rj, idx, nf) filter_vf_ni(rV, vi, ri, (vb).p, rj, idx, nf)
#else
#define NPC_FILTER_VE filter_ve
#define NPC_FILTER_EE filter_ee
#define NPC_FILTER_VF(rV, vi, ri, vb, rj, idx, nf) filter_vf(rV, vi, ri, vb, rj, idx, nf)
#endif

struct NbhPool {
  uint32_t* words;                // cap_chunks * CHW
  uint32_t* cursor;               // next free chunk
  uint32_t cap_chunks;
  uint32_t* first;                // per (owner slot, part): first chunk of its stream (NPC_NONE: no pair)
  unsigned long long* need;       // += chunks every owner would need (sizes the pool for the next rebuild)
  uint32_t* overflow;             // set when the pool ran out: the fill pass falls back to recomputation
};

// max over the real cells of the number of particles in the 3x3x3 stencil (sizes the staging area)
__global__ void k_stencil_max(GridDev G, const uint32_t* __restrict__ cbeg, const uint32_t* __restrict__ cend, unsigned long long* __restrict__ out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nreal_cells = (uint32_t)G.nc[0] * G.nc[1] * G.nc[2];
  unsigned int s = 0;
  if (c < nreal_cells) {
    const int ci = c % G.nc[0] + 1, cj = (c / G.nc[0]) % G.nc[1] + 1, ck = c / (G.nc[0] * G.nc[1]) + 1;
    if (cend[cell_lin(G, ci, cj, ck)] > cbeg[cell_lin(G, ci, cj, ck)])
      for (int dk = -1; dk <= 1; dk++)
        for (int dj = -1; dj <= 1; dj++)
          for (int di = -1; di <= 1; di++) {
            const uint32_t cb = cell_lin(G, ci + di, cj + dj, ck + dk);   // real cells are 1..nc: the shell always exists
            s += cend[cb] - cbeg[cb];
          }
  }
  s = __reduce_max_sync(0xffffffffu, s);
  if ((threadIdx.x & 31) == 0 && s) atomicMax(out, (unsigned long long)s);
}

// per-pair test groups, shared by the count pass and the replay: (wave, number of tests, decode of a test index)
struct PairDims {
  uint32_t nva, nea, nfa, nvb, neb, nfb;
};

// MODE 0: count;  MODE 1: count + record the streams;  MODE 2: fill by repeating the tests
template <int MODE>
__global__ void __launch_bounds__(NPC_THREADS, EDB_NPC_MINB) k_nbh_poly_cell(NbhArgs A, CntTab C, WTab W, uint32_t cap, NbhPool PL) {
  constexpr bool FILL = MODE == 2;
  extern __shared__ __align__(128) unsigned char smem[];
  const int nvmax = A.S.nvmax;
  double4* s_vert = reinterpret_cast<double4*>(smem);
  double4* s_pr = s_vert + (size_t)cap * nvmax;
  double4* s_q = s_pr + cap;
  unsigned long long* s_id = reinterpret_cast<unsigned long long*>(s_q + cap);
  double* s_h = reinterpret_cast<double*>(s_id + cap);
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(s_h + cap);
  uint32_t* s_type = s_slot + cap;
  uint8_t* s_gc = reinterpret_cast<uint8_t*>(s_type + cap);
  __shared__ unsigned long long s_bar;
  __shared__ uint32_t s_cbase[28], s_cbeg[27], s_next;
  __shared__ uint8_t s_cgc[27];
  __shared__ uint16_t s_queue[NPC_THREADS / 32][64];
  __shared__ uint32_t s_vmask[NPC_THREADS / 32][2][32];   // vertex-vs-OBB prefilter masks (mA, mB) of the pairs of one OBB round
#if EDB_NPC_HOIST & 1
  __shared__ double s_oa[NPC_THREADS / 32][16];           // the owner's world OBB (c, e1, e2, e3, ext), not enlarged
#endif
#if EDB_NPC_TAB
  __shared__ uint8_t s_tab[NPC_THREADS / 32][6][32];      // per pair: kept vertices of a, of b, edges of a, of b, V-E edges of a, of b
#endif

  const GridDev& G = A.G;
  const uint32_t c = blockIdx.x;
  const int ci = c % G.nc[0] + 1, cj = (c / G.nc[0]) % G.nc[1] + 1, ck = c / (G.nc[0] * G.nc[1]) + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  // ---- stage the stencil ------------------------------------------------------------------------------------------------
  if (wid == 0) {
    if (lane == 0) { mbar_init(&s_bar, 1); s_next = 0; }
    uint32_t cnt = 0, beg = 0;
    if (lane < 27) {
      const int dk = lane / 9 - 1, dj = (lane / 3) % 3 - 1, di = lane % 3 - 1;   // same visiting order as for_each_neighbor
      const int i = ci + di, j = cj + dj, k = ck + dk;
      const uint32_t cb = cell_lin(G, i, j, k);
      beg = A.cbeg[cb];
      cnt = A.cend[cb] - beg;
      s_cbeg[lane] = beg;
      s_cgc[lane] = is_ghost_cell(G, i, j, k) ? 1 : 0;
    }
    const uint32_t inc = warp_incl_scan(cnt);
    if (lane < 27) s_cbase[lane] = inc - cnt;
    const uint32_t nst0 = __shfl_sync(0xffffffffu, inc, 31);
    const bool owners = __shfl_sync(0xffffffffu, cnt, 13) != 0;   // empty centre cell: nothing to own, nothing is staged
    if (lane == 0) {
      s_cbase[27] = nst0;
      if (owners) mbar_expect_tx(&s_bar, nst0 * (uint32_t)(nvmax * 32 + 64));
    }
    __syncwarp();
    if (owners && lane < 27 && cnt) {
      const uint32_t o = inc - cnt;
      tma_load_1d(s_vert + (size_t)o * nvmax, A.vert + (size_t)beg * nvmax, cnt * (uint32_t)nvmax * 32u, &s_bar);
      tma_load_1d(s_pr + o, A.P.pr + beg, cnt * 32u, &s_bar);
      tma_load_1d(s_q + o, A.P.q + beg, cnt * 32u, &s_bar);
    }
  }
  __syncthreads();
  if (s_cbase[14] == s_cbase[13]) return;   // uniform for the block
  const uint32_t nst = s_cbase[27];
  // one pass over the staged particles (a thread finds its stencil cell in the prefix table) — a loop over the 27 cells would
  // serialise 27 global-load latencies on the ~17 threads a cell keeps busy while the other warps wait at the barrier
  for (uint32_t x = threadIdx.x; x < nst; x += NPC_THREADS) {
    int t = 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
      if (t + s < 27 && s_cbase[t + s] <= x) t += s;
    const uint32_t slot = s_cbeg[t] + (x - s_cbase[t]);
    s_slot[x] = slot;
    s_id[x] = A.P.id[slot];
    s_h[x] = A.P.h[slot];
    s_type[x] = type_of_bits(A.P.wg[slot].w);
    s_gc[x] = s_cgc[t];
  }
  mbar_wait(&s_bar, 0);
  __syncthreads();

  // ---- owners (taken from a shared counter: the work per owner varies with its number of neighbours) -----------------------
  const double rV = A.rcut_inc;
  const uint32_t own0 = s_cbase[13], nown = s_cbase[14] - own0;
  const uint32_t nparts = FILL ? 1u : NPC_PARTS;   // the recomputing fill needs the whole owner (its offsets run through the parts)
  const uint32_t nchunk = (nst + 31) / 32;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(&s_next, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nown * nparts) break;
    const uint32_t oi_ = item / nparts, part = item % nparts;
    const uint32_t ch_lo = nchunk * part / nparts * 32, ch_hi = min(nst, nchunk * (part + 1) / nparts * 32);
    const uint32_t La = own0 + oi_;
    const uint32_t a = s_slot[La];
    const double4 ra = s_pr[La], qa4 = s_q[La];
    const unsigned long long ida = s_id[La];
    const ShapeDev* __restrict__ psa = A.S.sh + s_type[La];
    const struct { int nv, ne, nf, e0, f0, fi0; double radius; } sa{psa->nv, psa->ne, psa->nf, psa->e0, psa->f0, psa->fi0, psa->radius};
    const double ha = s_h[La];
    const q4 qa{qa4.x, qa4.y, qa4.z, qa4.w};
    const SV va{s_vert + (size_t)La * nvmax};
    const double rma = ha * sa.radius;
    const int* ea = A.S.edges + 2 * sa.e0;
    // bounding radius of the owner's vertex set about its position (edge-edge cull)
    double bnda = 0.0;
    if (lane < sa.nv) bnda = norm(va[lane] - mk3(ra.x, ra.y, ra.z));
    for (int o = 16; o > 0; o >>= 1) bnda = fmax(bnda, __shfl_xor_sync(0xffffffffu, bnda, o));
#if EDB_NPC_HOIST & 1
    {
      __syncwarp();   // (the previous owner's readers are done)
      const Obb o = obb_world(psa->obb, mk3(ra.x, ra.y, ra.z), qa, ha);
      if (lane == 0) {
        double* d = s_oa[wid];
        d[0] = o.c.x; d[1] = o.c.y; d[2] = o.c.z; d[3] = o.e1.x; d[4] = o.e1.y; d[5] = o.e1.z; d[6] = o.e2.x; d[7] = o.e2.y; d[8] = o.e2.z;
        d[9] = o.e3.x; d[10] = o.e3.y; d[11] = o.e3.z; d[12] = o.ext.x; d[13] = o.ext.y; d[14] = o.ext.z;
      }
      __syncwarp();
    }
    auto owner_obb = [&]() {
      const double* d = s_oa[wid];
      return Obb{mk3(d[0], d[1], d[2]), mk3(d[3], d[4], d[5]), mk3(d[6], d[7], d[8]), mk3(d[9], d[10], d[11]), mk3(d[12], d[13], d[14])};
    };
#else
    auto owner_obb = [&]() { return obb_world(psa->obb, mk3(ra.x, ra.y, ra.z), qa, ha); };
#endif
    uint32_t run[4] = {0, 0, 0, 0};
    uint32_t base[4] = {0, 0, 0, 0};
    if (FILL) {
#pragma unroll
      for (int t = 0; t < 4; t++) base[t] = C.c[t][a];
    }
    // stream writer (lane 0 holds the state)
    uint32_t nw = 0, wchunk = NPC_NONE, wfill = CHW;
    bool wovf = false;
    auto emit = [&](uint32_t wv) {
      nw++;
      if (MODE == 1 && lane == 0 && !wovf) {
        if (wfill == CHW) {
          const uint32_t nc = atomicAdd(PL.cursor, 1u);
          if (nc >= PL.cap_chunks) { wovf = true; atomicExch(PL.overflow, 1u); return; }
          if (wchunk == NPC_NONE) PL.first[(size_t)a * NPC_PARTS + part] = nc;
          else PL.words[(size_t)wchunk * CHW] = nc;
          wchunk = nc;
          wfill = 1;
        }
        PL.words[(size_t)wchunk * CHW + wfill++] = wv;
      }
    };
    // one group of tests of a pair: test index x in [0, N), lanes take x = b0 + lane; accepted tests are written in x order
    auto group = [&](int t, uint32_t N, auto&& test, auto&& rec) {
      for (uint32_t b0 = 0; b0 < N; b0 += 32) {
        const uint32_t x = b0 + lane;
        const bool ok = x < N && test(x);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (FILL && ok) rec(x, base[t] + run[t] + __popc(m & ((1u << lane) - 1u)));
        if (!FILL) emit(m);
        run[t] += __popc(m);
      }
    };
    // Candidates that pass the cheap distance + id-order filter are queued (stencil order kept) and the 15-axis OBB test runs on a
    // full warp of queued candidates at a time instead of on the one or two lanes of a chunk that happen to be in range.
    uint16_t* qu = s_queue[wid];
    uint32_t nq = 0;
    for (uint32_t ch = ch_lo; ch < ch_hi; ch += 32) {
      const uint32_t L = ch + lane;
      bool near = false;
      if (L < nst && L != La) {
        const double4 rb = s_pr[L];
        const v3 d = mk3(rb.x, rb.y, rb.z) - mk3(ra.x, ra.y, ra.z);
        const double d2 = dot(d, d);
        const double rmax = ra.w + rb.w + rV;
        near = d2 > 0.0 && d2 < rmax * rmax && ida < s_id[L];   // ContactNeighborFilterFunc, then nbh_polyhedron.cpp:234-236
        if (near && A.ign_off) near = !bond_ignored(A, a, s_id[L]);   // bonded pair: InteractionManager::skip_ignored_interactions
      }
      const unsigned nm = __ballot_sync(0xffffffffu, near);
      if (near) qu[nq + __popc(nm & ((1u << lane) - 1u))] = (uint16_t)L;
      nq += __popc(nm);
      __syncwarp();
      const bool last_chunk = ch + 32 >= ch_hi;
      while (nq >= 32 || (last_chunk && nq > 0)) {
        const uint32_t take = nq < 32 ? nq : 32;
        const uint32_t Lc = lane < take ? qu[lane] : 0;
        bool pass = false;
#if EDB_NPC_HOIST & 2
        Obb ojk{};   // this lane's candidate, enlarged once: the prefilter rounds below read it by shuffle
#endif
        if (lane < take) {
          const double4 rb = s_pr[Lc], qb4 = s_q[Lc];
          Obb oi = owner_obb();
          Obb oj = obb_world(A.S.sh[s_type[Lc]].obb, mk3(rb.x, rb.y, rb.z), q4{qb4.x, qb4.y, qb4.z, qb4.w}, s_h[Lc]);
          obb_enlarge(oi, 0.5 * rV); obb_enlarge(oj, 0.5 * rV);
          pass = NPC_OBB_INTERSECT(oi, oj);
#if EDB_NPC_HOIST & 2
          ojk = oj;
#endif
        }
        unsigned pm = __ballot_sync(0xffffffffu, pass);
        const uint32_t rest = nq - take;   // < 32: a chunk adds at most 32 and the queue is drained as soon as it holds 32
        const uint16_t keep = lane < rest ? qu[take + lane] : (uint16_t)0;
        __syncwarp();
        if (lane < rest) qu[lane] = keep;
        __syncwarp();
        nq = rest;
        // Vertex-vs-OBB prefilters of ALL the pairs that passed this OBB round, batched: a pair needs nv_a (resp. nv_b) lanes, so a
        // round serves 32 / S pairs (S = 8, 16 or 32 lanes per pair) instead of one — with octahedra 4 pairs per round.  The OBBs are
        // enlarged in two half steps like the reference (:268-276, :305-310); same functions on the same inputs as before, only computed
        // ahead of the pair loop.
        {
          const int npass = __popc(pm);
          uint32_t* vmA = s_vmask[wid][0];
          uint32_t* vmB = s_vmask[wid][1];
          const int SA = sa.nv <= 8 ? 8 : (sa.nv <= 16 ? 16 : 32), SB = nvmax <= 8 ? 8 : (nvmax <= 16 ? 16 : 32);
          for (int q0 = 0; q0 < npass; q0 += 32 / SA) {
            const int q = q0 + lane / SA, v = lane % SA;
            const int l = (int)nth_set_bit(pm, (uint32_t)(q < npass ? q : 0));
            const uint32_t Lq = __shfl_sync(0xffffffffu, Lc, l);
            bool ok = false;
#if EDB_NPC_HOIST & 2
            auto sh3 = [&](v3 a) { return mk3(__shfl_sync(0xffffffffu, a.x, l), __shfl_sync(0xffffffffu, a.y, l), __shfl_sync(0xffffffffu, a.z, l)); };
            Obb oj{sh3(ojk.c), sh3(ojk.e1), sh3(ojk.e2), sh3(ojk.e3), sh3(ojk.ext)};   // enlarged once by the lane that tested it
            if (q < npass && v < sa.nv) {
              obb_enlarge(oj, 0.5 * rV);
              ok = NPC_OBB_INTERSECT(oj, obb_point(va[v], rma));
            }
#else
            if (q < npass && v < sa.nv) {
              const double4 rb = s_pr[Lq], qb4 = s_q[Lq];
              Obb oj = obb_world(A.S.sh[s_type[Lq]].obb, mk3(rb.x, rb.y, rb.z), q4{qb4.x, qb4.y, qb4.z, qb4.w}, s_h[Lq]);
              obb_enlarge(oj, 0.5 * rV); obb_enlarge(oj, 0.5 * rV);
              ok = NPC_OBB_INTERSECT(oj, obb_point(va[v], rma));
            }
#endif
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (q < npass && v == 0) vmA[q] = SA == 32 ? m : ((m >> (lane & ~(SA - 1))) & ((1u << SA) - 1u));
          }
          for (int q0 = 0; q0 < npass; q0 += 32 / SB) {
            const int q = q0 + lane / SB, v = lane % SB;
            const int l = (int)nth_set_bit(pm, (uint32_t)(q < npass ? q : 0));
            const uint32_t Lq = __shfl_sync(0xffffffffu, Lc, l);
            bool ok = false;
            if (q < npass) {
              const ShapeDev* __restrict__ psq = A.S.sh + s_type[Lq];
              if (v < psq->nv) {
                Obb oi = owner_obb();
                obb_enlarge(oi, 0.5 * rV); obb_enlarge(oi, 0.5 * rV);
                const SV vq{s_vert + (size_t)Lq * nvmax};
                ok = NPC_OBB_INTERSECT(oi, obb_point(vq[v], s_h[Lq] * psq->radius));
              }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (q < npass && v == 0) vmB[q] = SB == 32 ? m : ((m >> (lane & ~(SB - 1))) & ((1u << SB) - 1u));
          }
          __syncwarp();
        }
        int qi = 0;
        while (pm) {
          const int l = __ffs(pm) - 1;
          pm &= pm - 1;
          const uint32_t Lb = __shfl_sync(0xffffffffu, Lc, l);   // warp-uniform
          const uint32_t b = s_slot[Lb];
          const double4 rb = s_pr[Lb];
          const unsigned long long idb = s_id[Lb];
          const ShapeDev* __restrict__ psb = A.S.sh + s_type[Lb];
          const struct { int nv, ne, nf, e0, f0, fi0; double radius; } sb{psb->nv, psb->ne, psb->nf, psb->e0, psb->f0, psb->fi0, psb->radius};
          const double hb = s_h[Lb];
          const SV vb{s_vert + (size_t)Lb * nvmax};
          const double rmb = hb * sb.radius;
          const int* eb = A.S.edges + 2 * sb.e0;
          const unsigned mA = s_vmask[wid][0][qi], mB = s_vmask[wid][1][qi];
          qi++;
          // edge reach masks (see seg_point_d2): lanes [0, ne_a) test a's edges against b's bounding sphere, lanes [16, 16 + ne_b)
          // b's edges against a's; shapes with more than 16 edges keep every edge
          unsigned eA = sa.ne >= 32 ? 0xffffffffu : ((1u << sa.ne) - 1u), eB = sb.ne >= 32 ? 0xffffffffu : ((1u << sb.ne) - 1u);
          unsigned eAv = eA, eBv = eB;   // the same for the vertex-edge tests (wider reach)
          if (sa.ne <= 16 && sb.ne <= 16) {
            double bndb = 0.0;
            if (lane < sb.nv) bndb = norm(vb[lane] - mk3(rb.x, rb.y, rb.z));
            for (int o = 16; o > 0; o >>= 1) bndb = fmax(bndb, __shfl_xor_sync(0xffffffffu, bndb, o));
            const double Rr = rma + rmb + rV;
            bool keep = false, keepv = false;
            if (lane < sa.ne) {
              const v3 p0 = va[ea[2 * lane]], p1 = va[ea[2 * lane + 1]], cb = mk3(rb.x, rb.y, rb.z);
              const double d2 = seg_point_d2(p0, p1, cb);
              keep = d2 <= reach2(bndb, Rr, cb);
              keepv = d2 <= reach2(bndb, Rr + rV * norm(p1 - p0), cb);
            } else if (lane >= 16 && lane - 16 < sb.ne) {
              const v3 p0 = vb[eb[2 * (lane - 16)]], p1 = vb[eb[2 * (lane - 16) + 1]], ca = mk3(ra.x, ra.y, ra.z);
              const double d2 = seg_point_d2(p0, p1, ca);
              keep = d2 <= reach2(bnda, Rr, ca);
              keepv = d2 <= reach2(bnda, Rr + rV * norm(p1 - p0), ca);
            }
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            eA = km & 0xffffu;
            eB = km >> 16;
#if EDB_NPC_MERGE && EDB_NPC_VECULL
            const unsigned kv = __ballot_sync(0xffffffffu, keepv);
            eAv = kv & 0xffffu;
            eBv = kv >> 16;
#endif
          }
          // a pair whose prefilters leave no test at all (no vertex of either inside the other's enlarged OBB, no two edges within reach)
          // produces no record: it is neither evaluated nor written to the stream (the replay never hears of it)
          if ((mA | mB) == 0u && (eA == 0u || eB == 0u)) continue;
          const uint8_t ghost = s_gc[Lb];
          const uint8_t f0 = (uint8_t)(ghost << 1), f1 = (uint8_t)(1 | (ghost << 1));
          if (!FILL) {   // pair header of the stream
            emit(b); emit((uint32_t)idb); emit((uint32_t)(idb >> 32)); emit((uint32_t)ghost | (s_type[Lb] << 8)); emit(mA); emit(mB);
            emit(eA); emit(eB);
#if EDB_NPC_MERGE && EDB_NPC_VECULL
            emit(eAv); emit(eBv);
#endif
          }
          auto write = [&](int t, uint32_t k, uint32_t i, uint32_t j, uint32_t si, uint32_t sj, uint8_t flg) {
            W.w[t].ij[k] = make_uint2(i, j);
            W.w[t].sub[k] = make_uint2(si, sj);
            W.w[t].flg[k] = flg;
            W.w[t].ft[k] = make_double4(0, 0, 0, __longlong_as_double((long long)idb));
            W.w[t].mo[k] = make_double4(0, 0, 0, 0);
          };
          const uint32_t nvb = sb.nv, neb = sb.ne, nfb = sb.nf, nea = sa.ne, nfa = sa.nf;
          // The vertex-vs-OBB prefilter usually keeps 1-3 of the vertices: the test index space is built over the KEPT vertices only
          // (n-th set bit of the mask, ascending, so the emission order is unchanged) and the lanes stay busy.
          const uint32_t nA = __popc(mA), nB = __popc(mB);
          const uint32_t nEA = __popc(eA), nEB = __popc(eB);
#if EDB_NPC_TAB
          {
            __syncwarp();   // (the previous pair's readers are done)
            const unsigned below = (1u << lane) - 1u;
            const unsigned ms[6] = {mA, mB, eA, eB, eAv, eBv};
#pragma unroll
            for (int k = 0; k < 6; k++)
              if ((ms[k] >> lane) & 1u) s_tab[wid][k][__popc(ms[k] & below)] = (uint8_t)lane;
            __syncwarp();
          }
          auto vtxA = [&](uint32_t n) { return (uint32_t)s_tab[wid][0][n]; };
          auto vtxB = [&](uint32_t n) { return (uint32_t)s_tab[wid][1][n]; };
          auto edgA = [&](uint32_t n) { return (uint32_t)s_tab[wid][2][n]; };
          auto edgB = [&](uint32_t n) { return (uint32_t)s_tab[wid][3][n]; };
          auto edgAv = [&](uint32_t n) { return (uint32_t)s_tab[wid][4][n]; };
          auto edgBv = [&](uint32_t n) { return (uint32_t)s_tab[wid][5][n]; };
#else
          auto vtxA = [&](uint32_t n) { return nth_set_bit(mA, n); };
          auto vtxB = [&](uint32_t n) { return nth_set_bit(mB, n); };
          auto edgA = [&](uint32_t n) { return nth_set_bit(eA, n); };
          auto edgB = [&](uint32_t n) { return nth_set_bit(eB, n); };
          auto edgAv = [&](uint32_t n) { return nth_set_bit(eAv, n); };
          auto edgBv = [&](uint32_t n) { return nth_set_bit(eBv, n); };
#endif
          (void)edgAv; (void)edgBv;
          // x -> (x / d, x % d)
#if EDB_NPC_FDIV
          auto rcpf = [](uint32_t d) { return 1.0f / (float)(d ? d : 1u); };
          auto divm = [](uint32_t x, uint32_t d, float rc, uint32_t& q, uint32_t& r) { q = (uint32_t)(((float)x + 0.5f) * rc); r = x - q * d; };
#else
          auto rcpf = [](uint32_t) { return 0.0f; };
          auto divm = [](uint32_t x, uint32_t d, float, uint32_t& q, uint32_t& r) { q = x / d; r = x % d; };
#endif
          // A -> B (swap = 0): vertices of a against vertices / edges / faces of b
          {
            const float rc = rcpf(nvb);
            group(0, nA * nvb,
                  [&](uint32_t x) { uint32_t q, j; divm(x, nvb, rc, q, j); return filter_vv(rV, va[vtxA(q)], rma, vb[j], rmb); },
                  [&](uint32_t x, uint32_t k) { write(0, k, a, b, vtxA(x / nvb), x % nvb, f0); });
          }
#if EDB_NPC_MERGE
          {
            // tests [0, nAB) are A -> B, the rest B -> A (swap = 1: pi = b, pj = a; the owner stays a, interaction_pair.hpp:75-90)
#if EDB_NPC_VECULL
            const uint32_t nEBv = __popc(eBv), nEAv = __popc(eAv);   // (a's vertices meet b's edges within reach, and the other way round)
            const uint32_t nAB = nA * nEBv;
            const float rcB = rcpf(nEBv), rcA = rcpf(nEAv);
            group(1, nAB + nB * nEAv,
                  [&](uint32_t x) {
                    const bool fw = x < nAB;
                    const uint32_t y = fw ? x : x - nAB, nE = fw ? nEBv : nEAv;
                    uint32_t q, r;
                    divm(y, nE, fw ? rcB : rcA, q, r);
                    const uint32_t e = fw ? edgBv(r) : edgAv(r);
                    const uint32_t v = fw ? vtxA(q) : vtxB(q);
                    const SV& P = fw ? va : vb;
                    const SV& Q = fw ? vb : va;
                    const int* E = fw ? eb : ea;
                    return NPC_FILTER_VE(rV, P[v], fw ? rma : rmb, Q[E[2 * e]], Q[E[2 * e + 1]], fw ? rmb : rma);
                  },
                  [&](uint32_t x, uint32_t k) {
                    if (x < nAB) write(1, k, a, b, vtxA(x / nEBv), edgBv(x % nEBv), f0);
                    else write(1, k, b, a, vtxB((x - nAB) / nEAv), edgAv((x - nAB) % nEAv), f1);
                  });
#else
            const uint32_t nAB = nA * neb;
            const float rcB = rcpf(neb), rcA = rcpf(nea);
            group(1, nAB + nB * nea,
                  [&](uint32_t x) {
                    const bool fw = x < nAB;
                    const uint32_t y = fw ? x : x - nAB, nE = fw ? neb : nea;
                    uint32_t q, e;
                    divm(y, nE, fw ? rcB : rcA, q, e);
                    const uint32_t v = fw ? vtxA(q) : vtxB(q);
                    const SV& P = fw ? va : vb;
                    const SV& Q = fw ? vb : va;
                    const int* E = fw ? eb : ea;
                    return NPC_FILTER_VE(rV, P[v], fw ? rma : rmb, Q[E[2 * e]], Q[E[2 * e + 1]], fw ? rmb : rma);
                  },
                  [&](uint32_t x, uint32_t k) {
                    if (x < nAB) write(1, k, a, b, vtxA(x / neb), x % neb, f0);
                    else write(1, k, b, a, vtxB((x - nAB) / nea), (x - nAB) % nea, f1);
                  });
#endif
            const uint32_t nAF = nA * nfb;
            const float rfB = rcpf(nfb), rfA = rcpf(nfa);
            group(2, nAF + nB * nfa,
                  [&](uint32_t x) {
                    const bool fw = x < nAF;
                    const uint32_t y = fw ? x : x - nAF, nF = fw ? nfb : nfa;
                    uint32_t q, f;
                    divm(y, nF, fw ? rfB : rfA, q, f);
                    const uint32_t v = fw ? vtxA(q) : vtxB(q);
                    const SV& P = fw ? va : vb;
                    const SV& Q = fw ? vb : va;
                    const int f0s = fw ? sb.f0 : sa.f0, fi0s = fw ? sb.fi0 : sa.fi0;
                    return NPC_FILTER_VF(rV, P[v], fw ? rma : rmb, Q, fw ? rmb : rma, A.S.fidx + fi0s + A.S.foff[f0s + f], A.S.fn[f0s + f]);
                  },
                  [&](uint32_t x, uint32_t k) {
                    if (x < nAF) write(2, k, a, b, vtxA(x / nfb), x % nfb, f0);
                    else write(2, k, b, a, vtxB((x - nAF) / nfa), (x - nAF) % nfa, f1);
                  });
            const float rce = rcpf(nEB);
            group(3, nEA * nEB,
                  [&](uint32_t x) {
                    uint32_t q, r;
                    divm(x, nEB, rce, q, r);
                    const uint32_t i = edgA(q), j = edgB(r);
                    return NPC_FILTER_EE(rV, va[ea[2 * i]], va[ea[2 * i + 1]], rma, vb[eb[2 * j]], vb[eb[2 * j + 1]], rmb);
                  },
                  [&](uint32_t x, uint32_t k) { write(3, k, a, b, edgA(x / nEB), edgB(x % nEB), f0); });
            continue;
          }
#endif
          group(1, nA * neb,
                [&](uint32_t x) { const uint32_t i = vtxA(x / neb), j = x % neb; return NPC_FILTER_VE(rV, va[i], rma, vb[eb[2 * j]], vb[eb[2 * j + 1]], rmb); },
                [&](uint32_t x, uint32_t k) { write(1, k, a, b, vtxA(x / neb), x % neb, f0); });
          group(2, nA * nfb,
                [&](uint32_t x) {
                  const uint32_t i = vtxA(x / nfb), j = x % nfb;
                  return NPC_FILTER_VF(rV, va[i], rma, vb, rmb, A.S.fidx + sb.fi0 + A.S.foff[sb.f0 + j], A.S.fn[sb.f0 + j]);
                },
                [&](uint32_t x, uint32_t k) { write(2, k, a, b, vtxA(x / nfb), x % nfb, f0); });
          // edge-edge over the edges within reach only (ascending in both indices: the emission order of the full double loop)
          group(3, nEA * nEB,
                [&](uint32_t x) {
                  const uint32_t i = edgA(x / nEB), j = edgB(x % nEB);
                  return NPC_FILTER_EE(rV, va[ea[2 * i]], va[ea[2 * i + 1]], rma, vb[eb[2 * j]], vb[eb[2 * j + 1]], rmb);
                },
                [&](uint32_t x, uint32_t k) { write(3, k, a, b, edgA(x / nEB), edgB(x % nEB), f0); });
          // B -> A (swap = 1): pi = b, pj = a; the owner stays a (interaction_pair.hpp:75-90)
          group(1, nB * nea,
                [&](uint32_t x) { const uint32_t j = vtxB(x / nea), i = x % nea; return NPC_FILTER_VE(rV, vb[j], rmb, va[ea[2 * i]], va[ea[2 * i + 1]], rma); },
                [&](uint32_t x, uint32_t k) { write(1, k, b, a, vtxB(x / nea), x % nea, f1); });
          group(2, nB * nfa,
                [&](uint32_t x) {
                  const uint32_t j = vtxB(x / nfa), i = x % nfa;
                  return NPC_FILTER_VF(rV, vb[j], rmb, va, rma, A.S.fidx + sa.fi0 + A.S.foff[sa.f0 + i], A.S.fn[sa.f0 + i]);
                },
                [&](uint32_t x, uint32_t k) { write(2, k, b, a, vtxB(x / nfa), x % nfa, f1); });
        }
      }
    }
    if (!FILL) {
      if (nw) emit(NPC_NONE);   // end of this part's stream
      if (lane == 0) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
          if (NPC_PARTS == 1) C.c[t][a] = run[t];
          else if (run[t]) atomicAdd(&C.c[t][a], run[t]);   // the host zeroed the counts; integer sums: order-independent
        }
        if (nw) atomicAdd(PL.need, (unsigned long long)((nw + CHW - 2) / (CHW - 1)));
        if (MODE == 1 && wchunk == NPC_NONE) PL.first[(size_t)a * NPC_PARTS + part] = NPC_NONE;
      }
    }
  }
}

// FILL pass by replay: warp per owner, reads the owner's stream (pair headers + ballots of the count pass) and writes the records
// at the owner's CSR offsets — the same control flow as k_nbh_poly_cell with every filter test replaced by its recorded bit.
__global__ void __launch_bounds__(256) k_nbh_poly_replay(NbhArgs A, CntTab C, WTab W, NbhPool PL) {
  const int lane = threadIdx.x & 31;
  const uint32_t a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= A.nreal) return;
  uint32_t chunk = NPC_NONE, pos = 1;
  auto rd = [&]() {
    if (pos == CHW) { chunk = PL.words[(size_t)chunk * CHW]; pos = 1; }
    return PL.words[(size_t)chunk * CHW + pos++];
  };
  const ShapeDev* __restrict__ psa = A.S.sh + type_of_bits(A.P.wg[a].w);
  const uint32_t nea = psa->ne, nfa = psa->nf;
  uint32_t run[4] = {0, 0, 0, 0}, base[4];
#pragma unroll
  for (int t = 0; t < 4; t++) base[t] = C.c[t][a];
  for (uint32_t part = 0; part < NPC_PARTS; part++) {   // the parts of the owner's stencil in order: the order of the unsplit walk
   chunk = PL.first[(size_t)a * NPC_PARTS + part];
   if (chunk == NPC_NONE) continue;
   pos = 1;
   for (;;) {
    const uint32_t b = rd();
    if (b == NPC_NONE) break;
    const uint32_t idlo = rd(), idhi = rd(), gt = rd(), mA = rd(), mB = rd(), eA = rd(), eB = rd();
#if EDB_NPC_MERGE && EDB_NPC_VECULL
    const uint32_t eAv = rd(), eBv = rd();
#endif
    const unsigned long long idb = ((unsigned long long)idhi << 32) | idlo;
    const uint8_t ghost = (uint8_t)(gt & 0xffu);
    const ShapeDev* __restrict__ psb = A.S.sh + (gt >> 8);
    const uint32_t nvb = psb->nv, neb = psb->ne, nfb = psb->nf;
    const uint8_t f0 = (uint8_t)(ghost << 1), f1 = (uint8_t)(1 | (ghost << 1));
    const uint32_t nA = __popc(mA), nB = __popc(mB);
    auto group = [&](int t, uint32_t N, auto&& rec) {
      for (uint32_t b0 = 0; b0 < N; b0 += 32) {
        const unsigned m = rd();
        if ((m >> lane) & 1u) rec(b0 + lane, base[t] + run[t] + __popc(m & ((1u << lane) - 1u)));
        run[t] += __popc(m);
      }
    };
    auto write = [&](int t, uint32_t k, uint32_t i, uint32_t j, uint32_t si, uint32_t sj, uint8_t flg) {
      W.w[t].ij[k] = make_uint2(i, j);
      W.w[t].sub[k] = make_uint2(si, sj);
      W.w[t].flg[k] = flg;
      W.w[t].ft[k] = make_double4(0, 0, 0, __longlong_as_double((long long)idb));
      W.w[t].mo[k] = make_double4(0, 0, 0, 0);
    };
    auto vtxA = [&](uint32_t n) { return nth_set_bit(mA, n); };
    auto vtxB = [&](uint32_t n) { return nth_set_bit(mB, n); };
    group(0, nA * nvb, [&](uint32_t x, uint32_t k) { write(0, k, a, b, vtxA(x / nvb), x % nvb, f0); });
#if EDB_NPC_MERGE
    {
      const uint32_t nAF = nA * nfb, nEBm = __popc(eB);
#if EDB_NPC_VECULL
      const uint32_t nEBv = __popc(eBv), nEAv = __popc(eAv), nAB = nA * nEBv;
      group(1, nAB + nB * nEAv, [&](uint32_t x, uint32_t k) {
        if (x < nAB) write(1, k, a, b, vtxA(x / nEBv), nth_set_bit(eBv, x % nEBv), f0);
        else write(1, k, b, a, vtxB((x - nAB) / nEAv), nth_set_bit(eAv, (x - nAB) % nEAv), f1);
      });
#else
      const uint32_t nAB = nA * neb;
      group(1, nAB + nB * nea, [&](uint32_t x, uint32_t k) {
        if (x < nAB) write(1, k, a, b, vtxA(x / neb), x % neb, f0);
        else write(1, k, b, a, vtxB((x - nAB) / nea), (x - nAB) % nea, f1);
      });
#endif
      group(2, nAF + nB * nfa, [&](uint32_t x, uint32_t k) {
        if (x < nAF) write(2, k, a, b, vtxA(x / nfb), x % nfb, f0);
        else write(2, k, b, a, vtxB((x - nAF) / nfa), (x - nAF) % nfa, f1);
      });
      group(3, __popc(eA) * nEBm,
            [&](uint32_t x, uint32_t k) { write(3, k, a, b, nth_set_bit(eA, x / nEBm), nth_set_bit(eB, x % nEBm), f0); });
      continue;
    }
#endif
    group(1, nA * neb, [&](uint32_t x, uint32_t k) { write(1, k, a, b, vtxA(x / neb), x % neb, f0); });
    group(2, nA * nfb, [&](uint32_t x, uint32_t k) { write(2, k, a, b, vtxA(x / nfb), x % nfb, f0); });
    const uint32_t nEB = __popc(eB);
    group(3, __popc(eA) * nEB,
          [&](uint32_t x, uint32_t k) { write(3, k, a, b, nth_set_bit(eA, x / nEB), nth_set_bit(eB, x % nEB), f0); });
    group(1, nB * nea, [&](uint32_t x, uint32_t k) { write(1, k, b, a, vtxB(x / nea), x % nea, f1); });
    group(2, nB * nfa, [&](uint32_t x, uint32_t k) { write(2, k, b, a, vtxB(x / nfa), x % nfa, f1); });
   }
  }
}

}  // namespace edb
