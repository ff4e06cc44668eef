// comm.cuh — spatial decomposition over the GPUs of one box: ghost-layer plan + NCCL exchange (replaces the MPI path of
// ghost_update_rq / UpdateGhostsNode, src/grid/update_ghosts.cu:46-63, and of migrate_cell_particles_interaction).
//
// Decomposition: uniform bricks on a pgrid[0] x pgrid[1] x pgrid[2] process grid, rank = x + px*(y + py*z).  Because all
// bricks have the same cell grid, every rank can compute every other rank's ghost table locally, so the send lists need
// no handshake: rank A sends, in B's enumeration order, exactly the cells B's table names as sources on A.
// NCCL is resolved at run time (dlopen of the libnccl already loaded by torch, else the system one) so that the library
// has no link-time dependency and a single NCCL instance lives in the process.
#pragma once
#include <dlfcn.h>

#include <map>

#include "common.cuh"

namespace edb {

// minimal NCCL ABI (nccl.h 2.x): only what the exchange needs
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process (torch)?
      if (h) break;
    }
    if (!h)
      for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
      }
    if (!h) return false;
#define EDB_SYM(f) f = (decltype(f))dlsym(h, "nccl" #f)
    EDB_SYM(GetUniqueId); EDB_SYM(CommInitRank); EDB_SYM(CommDestroy); EDB_SYM(GroupStart); EDB_SYM(GroupEnd);
    EDB_SYM(Send); EDB_SYM(Recv); EDB_SYM(AllReduce); EDB_SYM(AllGather); EDB_SYM(GetErrorString);
#undef EDB_SYM
    return GetUniqueId && CommInitRank && Send && Recv && GroupStart && GroupEnd && AllReduce && AllGather;
  }
};
static NcclApi g_nccl;

#define NCK(call)                                                                                  \
  do {                                                                                             \
    ncclResult_t _r = (call);                                                                      \
    if (_r != 0) throw edb::StateError{EDB_ERR_NCCL, std::string("NCCL: ") + #call + " -> " +      \
                                       (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?")}; \
  } while (0)

struct GhostEntry {
  uint32_t gcell;   // ghost cell (receiver's grid)
  uint32_t scell;   // source real cell (sender's grid)
  double shift[3];  // added to the source position
};

// pack kernel: one thread per outgoing ghost particle; src slots come from the send map built at rebuild
__global__ void k_pack_ghosts(PView P, const uint32_t* __restrict__ smap, const uint32_t* __restrict__ sent, const double* __restrict__ sshift,
                              uint32_t n, double4* __restrict__ o_pr, double4* __restrict__ o_vm, double4* __restrict__ o_wg,
                              double4* __restrict__ o_q, double* __restrict__ o_h, uint64_t* __restrict__ o_id, bool all) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t s = smap[k], e = sent[k];
  double4 r = ld4(P.pr + s);
  r.x = r.x + sshift[3 * e]; r.y = r.y + sshift[3 * e + 1]; r.z = r.z + sshift[3 * e + 2];
  st4(o_pr + k, r); st4(o_vm + k, ld4(P.vm + s)); st4(o_wg + k, ld4(P.wg + s)); st4(o_q + k, ld4(P.q + s));
  if (all) { o_h[k] = P.h[s]; o_id[k] = P.id[s]; }
}
__global__ void k_send_map(const uint32_t* __restrict__ scell, uint32_t nent, const uint32_t* __restrict__ cstart,
                           const uint32_t* __restrict__ soff, uint32_t* __restrict__ smap, uint32_t* __restrict__ sent) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nent) return;
  const uint32_t c = scell[k], s0 = cstart[c], cnt = cstart[c + 1] - s0, o = soff[k];
  for (uint32_t t = 0; t < cnt; t++) { smap[o + t] = s0 + t; sent[o + t] = k; }
}
__global__ void k_recv_cells(const uint32_t* __restrict__ gcell, uint32_t nent, const uint32_t* __restrict__ roff, const uint32_t* __restrict__ rcnt,
                             uint32_t base, uint32_t* __restrict__ cbeg, uint32_t* __restrict__ cend, uint32_t* __restrict__ cidx) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nent) return;
  const uint32_t b = base + roff[k], e = b + rcnt[k];
  cbeg[gcell[k]] = b; cend[gcell[k]] = e;
  for (uint32_t s = b; s < e; s++) cidx[s] = gcell[k];
}

struct Peer {
  int rank;
  std::vector<GhostEntry> recv;  // our ghost cells fed by this peer (our table order)
  std::vector<GhostEntry> send;  // the peer's ghost cells fed by us (the peer's table order)
  DevBuf<uint32_t> d_rgcell, d_rcnt, d_roff, d_scell, d_scnt, d_soff, d_smap, d_sent;
  DevBuf<double> d_sshift;
  uint32_t n_send = 0, n_recv = 0, recv_base = 0;  // particles; recv_base = offset inside the ghost region
  DevBuf<double4> s_pr, s_vm, s_wg, s_q;
  DevBuf<double> s_h;
  DevBuf<uint64_t> s_id;
};

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int pg[3] = {1, 1, 1}, rc[3] = {0, 0, 0};
  uint32_t n_local_ghost = 0;
  std::map<int, Peer> peers;
  DevBuf<unsigned long long> d_red;
  DevBuf<uint32_t> d_gather;

  static int unique_id(uint8_t out[128]) {
    if (!g_nccl.load()) return EDB_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return EDB_ERR_NCCL;
    memcpy(out, id.internal, 128);
    return EDB_OK;
  }
  void init(const edb_config& cfg, const uint8_t id_bytes[128]) {
    if (cfg.nranks <= 1) return;
    if (!g_nccl.load()) throw StateError{EDB_ERR_NCCL, "libnccl.so.2 not found"};
    ncclUniqueId id;
    memcpy(id.internal, id_bytes, 128);
    NCK(g_nccl.CommInitRank(&comm, cfg.nranks, id, cfg.rank));
  }
  void destroy() {
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    comm = nullptr;
    for (auto& kv : peers) {
      Peer& p = kv.second;
      p.d_rgcell.release(); p.d_rcnt.release(); p.d_roff.release(); p.d_scell.release(); p.d_scnt.release(); p.d_soff.release();
      p.d_smap.release(); p.d_sent.release(); p.d_sshift.release(); p.s_pr.release(); p.s_vm.release(); p.s_wg.release(); p.s_q.release();
      p.s_h.release(); p.s_id.release();
    }
    d_red.release();
    d_gather.release();
  }
  static void coords(const int pg[3], int r, int out[3]) {
    out[0] = r % pg[0];
    out[1] = (r / pg[0]) % pg[1];
    out[2] = r / (pg[0] * pg[1]);
  }
  void plan_reset(const edb_config& cfg, const GridDev&) {
    rank = cfg.rank; nranks = cfg.nranks < 1 ? 1 : cfg.nranks;
    for (int d = 0; d < 3; d++) pg[d] = cfg.pgrid[d] < 1 ? 1 : cfg.pgrid[d];
    if (pg[0] * pg[1] * pg[2] != nranks) throw StateError{EDB_ERR_INVALID, "pgrid product != nranks"};
    coords(pg, rank, rc);
    peers.clear();
  }
  // source of ghost cell idx[] of the brick at coords `who` : returns false when there is none (non-periodic wall)
  static bool source_of(const edb_config& cfg, const GridDev& G, const int pg[3], const int who[3], const int idx[3], int src[3],
                        double shift[3], int* peer) {
    int rb[3];
    for (int d = 0; d < 3; d++) {
      rb[d] = who[d]; src[d] = idx[d]; shift[d] = 0.0;
      if (idx[d] == 0) {
        rb[d] = who[d] - 1; src[d] = G.nc[d];
        if (rb[d] < 0) {
          if (!cfg.periodic[d]) return false;
          rb[d] = pg[d] - 1; shift[d] = -(cfg.ghi[d] - cfg.glo[d]);
        }
      } else if (idx[d] == G.g[d] - 1) {
        rb[d] = who[d] + 1; src[d] = 1;
        if (rb[d] >= pg[d]) {
          if (!cfg.periodic[d]) return false;
          rb[d] = 0; shift[d] = cfg.ghi[d] - cfg.glo[d];
        }
      }
    }
    *peer = rb[0] + pg[0] * (rb[1] + pg[1] * rb[2]);
    return true;
  }
  bool ghost_source(const edb_config& cfg, const GridDev& G, const int idx[3], int src[3], double shift[3], int* peer) const {
    return source_of(cfg, G, pg, rc, idx, src, shift, peer);
  }
  void plan_add_remote(int peer, uint32_t gcell, uint32_t scell, const double shift[3]) {
    Peer& p = peers[peer];
    p.rank = peer;
    p.recv.push_back(GhostEntry{gcell, scell, {shift[0], shift[1], shift[2]}});
  }
  // send lists: replay every neighbour's table and keep the entries that name us as the source
  void plan_send_lists(const edb_config& cfg, const GridDev& G) {
    for (auto& kv : peers) kv.second.send.clear();
    for (int pr = 0; pr < nranks; pr++) {
      if (pr == rank) continue;
      int who[3];
      coords(pg, pr, who);
      for (int k = 0; k < G.g[2]; k++)
        for (int j = 0; j < G.g[1]; j++)
          for (int i = 0; i < G.g[0]; i++) {
            if (!(i == 0 || j == 0 || k == 0 || i == G.g[0] - 1 || j == G.g[1] - 1 || k == G.g[2] - 1)) continue;
            const int idx[3] = {i, j, k};
            int src[3], peer;
            double shift[3];
            if (!source_of(cfg, G, pg, who, idx, src, shift, &peer)) continue;
            if (peer != rank) continue;
            Peer& p = peers[pr];
            p.rank = pr;
            const uint32_t cg = (uint32_t)i + G.g[0] * ((uint32_t)j + G.g[1] * (uint32_t)k);
            const uint32_t cs = (uint32_t)src[0] + G.g[0] * ((uint32_t)src[1] + G.g[1] * (uint32_t)src[2]);
            p.send.push_back(GhostEntry{cg, cs, {shift[0], shift[1], shift[2]}});
          }
    }
  }
  void plan_finalize() {
    for (auto& kv : peers) {
      Peer& p = kv.second;
      std::vector<uint32_t> a, b;
      std::vector<double> sh;
      for (auto& e : p.recv) a.push_back(e.gcell);
      for (auto& e : p.send) { b.push_back(e.scell); sh.push_back(e.shift[0]); sh.push_back(e.shift[1]); sh.push_back(e.shift[2]); }
      p.d_rgcell.ensure(a.size() + 1); p.d_rcnt.ensure(a.size() + 2); p.d_roff.ensure(a.size() + 2);
      p.d_scell.ensure(b.size() + 1); p.d_scnt.ensure(b.size() + 2); p.d_soff.ensure(b.size() + 2); p.d_sshift.ensure(sh.size() + 3);
      if (!a.empty()) CK(cudaMemcpy(p.d_rgcell.p, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
      if (!b.empty()) {
        CK(cudaMemcpy(p.d_scell.p, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.d_sshift.p, sh.data(), sh.size() * 8, cudaMemcpyHostToDevice));
      }
    }
  }
  // every rank contributes n uint32; out (host) receives nranks*n in rank order
  void allgather_u32(const uint32_t* h_in, uint32_t n, uint32_t* h_out, cudaStream_t st) {
    if (!comm) throw StateError{EDB_ERR_STATE, "multi-GPU context without edb_comm_init"};
    d_gather.ensure((size_t)n * (nranks + 1));
    CK(cudaMemcpyAsync(d_gather.p, h_in, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    NCK(g_nccl.AllGather(d_gather.p, d_gather.p + n, n, ncclUint32, comm, st));
    CK(cudaMemcpyAsync(h_out, d_gather.p + n, (size_t)n * nranks * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  // variable-size all-to-all of raw bytes: (s/r)off, (s/r)cnt in ELEMENTS of `elem` bytes per rank
  void alltoallv_bytes(const void* sbuf, const uint32_t* soff, const uint32_t* scnt, void* rbuf, const uint32_t* roff,
                       const uint32_t* rcnt, size_t elem, cudaStream_t st) {
    if (!comm) throw StateError{EDB_ERR_STATE, "multi-GPU context without edb_comm_init"};
    NCK(g_nccl.GroupStart());
    for (int r = 0; r < nranks; r++) {
      if (r == rank) continue;
      if (scnt[r]) NCK(g_nccl.Send((const char*)sbuf + (size_t)soff[r] * elem, (size_t)scnt[r] * elem, ncclUint8, r, comm, st));
      if (rcnt[r]) NCK(g_nccl.Recv((char*)rbuf + (size_t)roff[r] * elem, (size_t)rcnt[r] * elem, ncclUint8, r, comm, st));
    }
    NCK(g_nccl.GroupEnd());
  }
  // rebuild: per-entry particle counts of the ghost cells travel from the source rank to the ghost owner
  void exchange_ghost_counts(cudaStream_t st) {
    if (!comm) throw StateError{EDB_ERR_STATE, "multi-GPU context without edb_comm_init"};
    NCK(g_nccl.GroupStart());
    for (auto& kv : peers) {
      Peer& p = kv.second;
      if (!p.send.empty()) NCK(g_nccl.Send(p.d_scnt.p, p.send.size(), ncclUint32, p.rank, comm, st));
      if (!p.recv.empty()) NCK(g_nccl.Recv(p.d_rcnt.p, p.recv.size(), ncclUint32, p.rank, comm, st));
    }
    NCK(g_nccl.GroupEnd());
  }
  // in place, on the stream, no host round trip (rebuild trigger of the speculative step loop)
  void allreduce_sum_u64_dev(unsigned long long* d, cudaStream_t st) {
    if (!comm) return;
    NCK(g_nccl.AllReduce(d, d, 1, ncclUint64, ncclSum, comm, st));
  }
  uint64_t allreduce_sum_u64(uint64_t v, cudaStream_t st) {
    if (!comm) return v;
    d_red.ensure(2);
    unsigned long long hv = v;
    CK(cudaMemcpyAsync(d_red.p, &hv, 8, cudaMemcpyHostToDevice, st));
    NCK(g_nccl.AllReduce(d_red.p, d_red.p + 1, 1, ncclUint64, ncclSum, comm, st));
    CK(cudaMemcpyAsync(&hv, d_red.p + 1, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return hv;
  }
  // per-step refresh (all=false) or full rebuild refresh (all=true) of the remote ghosts; data lands in place
  void exchange_ghosts(PView P, uint32_t nreal, bool all, cudaStream_t st, uint64_t* launches) {
    if (!comm) throw StateError{EDB_ERR_STATE, "multi-GPU context without edb_comm_init"};
    for (auto& kv : peers) {
      Peer& p = kv.second;
      if (!p.n_send) continue;
      k_pack_ghosts<<<(p.n_send + BLK - 1) / BLK, BLK, 0, st>>>(P, p.d_smap.p, p.d_sent.p, p.d_sshift.p, p.n_send, p.s_pr.p, p.s_vm.p,
                                                              p.s_wg.p, p.s_q.p, p.s_h.p, p.s_id.p, all);
      (*launches)++;
    }
    NCK(g_nccl.GroupStart());
    for (auto& kv : peers) {
      Peer& p = kv.second;
      const uint32_t base = nreal + n_local_ghost + p.recv_base;
      if (p.n_send) {
        NCK(g_nccl.Send(p.s_pr.p, (size_t)p.n_send * 4, ncclFloat64, p.rank, comm, st));
        NCK(g_nccl.Send(p.s_vm.p, (size_t)p.n_send * 4, ncclFloat64, p.rank, comm, st));
        NCK(g_nccl.Send(p.s_wg.p, (size_t)p.n_send * 4, ncclFloat64, p.rank, comm, st));
        NCK(g_nccl.Send(p.s_q.p, (size_t)p.n_send * 4, ncclFloat64, p.rank, comm, st));
        if (all) {
          NCK(g_nccl.Send(p.s_h.p, (size_t)p.n_send, ncclFloat64, p.rank, comm, st));
          NCK(g_nccl.Send(p.s_id.p, (size_t)p.n_send, ncclUint64, p.rank, comm, st));
        }
      }
      if (p.n_recv) {
        NCK(g_nccl.Recv(P.pr + base, (size_t)p.n_recv * 4, ncclFloat64, p.rank, comm, st));
        NCK(g_nccl.Recv(P.vm + base, (size_t)p.n_recv * 4, ncclFloat64, p.rank, comm, st));
        NCK(g_nccl.Recv(P.wg + base, (size_t)p.n_recv * 4, ncclFloat64, p.rank, comm, st));
        NCK(g_nccl.Recv(P.q + base, (size_t)p.n_recv * 4, ncclFloat64, p.rank, comm, st));
        if (all) {
          NCK(g_nccl.Recv(P.h + base, (size_t)p.n_recv, ncclFloat64, p.rank, comm, st));
          NCK(g_nccl.Recv(P.id + base, (size_t)p.n_recv, ncclUint64, p.rank, comm, st));
        }
      }
    }
    NCK(g_nccl.GroupEnd());
  }
};

}  // namespace edb
