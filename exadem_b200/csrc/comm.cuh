// comm.cuh — spatial decomposition over the GPUs of one box: ghost-layer plan + NCCL exchange (replaces the MPI path of
// ghost_update_rq / UpdateGhostsNode, src/grid/update_ghosts.cu:46-63, and of migrate_cell_particles_interaction).
//
// Decomposition: rectilinear bricks on a pgrid[0] x pgrid[1] x pgrid[2] process grid, rank = x + px*(y + py*z): uniform by default,
// or cut at the per-axis positions of edb_set_brick_cuts (cost-balanced, exadem_b200.parallel.balanced_cuts).  The cuts are global
// knowledge and a brick's cell grid follows from its bounds, so every rank can compute every other rank's ghost table locally and the
// send lists need no handshake: rank A sends, in B's enumeration order, exactly the cells B's table names as sources on A.  (Tensor-
// product cuts keep two face neighbours on the same transverse cell grid, so ghost cells still map one to one onto source cells.)
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

// ---------------------------------------------------------------------------------------------------------------
// Per-step ghost refresh over NVLink peer memory (no NCCL on the per-step path).
//
// Every rank owns one "mailbox" allocation, exported to the other ranks with CUDA IPC:
//     [ mail[2][64] u64 | landing[0] : pr | vm | wg | q (cap double4 each) | landing[1] : ... ]
// k_ghost_push   (sender)   stores the refreshed sectors of its boundary particles straight into the receivers' landing zones
//                           (256-bit stores over NVLink), and its last block then stores {epoch, local rebuild-trigger count}
//                           into the mailbox of EVERY rank (release at system scope).
// k_ghost_land   (receiver) waits until every rank's mail of this epoch has arrived, sums the trigger counts (this is the
//                           all-reduce of `trigger_move_particles`, for free) and copies the landing zone into the ghost slots.
// Landing zones and mail are double-buffered by epoch parity: a sender can be at most one exchange ahead of a receiver
// (its next push comes after its own land, which needs the receiver's push of the current epoch), so no "ready" handshake is
// needed.  The rebuild-time exchange (ids, homothety, variable counts) stays on NCCL.
// ---------------------------------------------------------------------------------------------------------------
constexpr int P2P_MAXR = 64;     // mailbox slots
constexpr int P2P_MAXPEER = 26;  // neighbouring bricks
struct PushPeer {
  const uint32_t* smap;    // source slot per outgoing ghost
  const uint32_t* sent;    // send-table entry per outgoing ghost (-> shift)
  const double* sshift;
  uint32_t start, n;       // range inside the combined launch
  double4* land;           // receiver's landing zone of this epoch, already offset to our range; arrays are `cap` apart
  uint32_t cap;
};
struct PushArgs {
  PView P;
  PushPeer peer[P2P_MAXPEER];
  int npeer;
  uint32_t total;
  unsigned long long* mail[P2P_MAXR];  // mailbox slot [parity][me] of every other rank
  int nmail;
  unsigned int* done;                  // block counter (device, zero between launches)
  const unsigned long long* count;     // local trigger count (nullptr: 0)
  uint32_t epoch;
};
__global__ void __launch_bounds__(BLK) k_ghost_push(const __grid_constant__ PushArgs A) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < A.total) {
    int pi = 0;
    while (pi + 1 < A.npeer && k >= A.peer[pi + 1].start) pi++;
    const PushPeer& p = A.peer[pi];
    const uint32_t l = k - p.start;
    const uint32_t s = p.smap[l], e = p.sent[l];
    double4 r = ld4(A.P.pr + s);
    r.x = r.x + p.sshift[3 * e]; r.y = r.y + p.sshift[3 * e + 1]; r.z = r.z + p.sshift[3 * e + 2];
    st4(p.land + l, r);
    st4(p.land + (size_t)p.cap + l, ld4(A.P.vm + s));
    st4(p.land + 2 * (size_t)p.cap + l, ld4(A.P.wg + s));
    st4(p.land + 3 * (size_t)p.cap + l, ld4(A.P.q + s));
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(A.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x == 0) *A.done = 0;
  if ((int)threadIdx.x < A.nmail) {
    unsigned long long c = A.count ? *A.count : 0ull;
    if (c > 0xffffffffull) c = 0xffffffffull;
    const unsigned long long v = ((unsigned long long)A.epoch << 32) | c;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.mail[threadIdx.x]), "l"(v) : "memory");
  }
}
struct LandArgs {
  PView P;
  const unsigned long long* mail;   // our mailbox row of this parity: one slot per rank
  int nranks, me;
  uint32_t epoch;
  unsigned long long* count;        // in: local trigger count, out: sum over the ranks (nullptr: not reduced)
  const double4* land;              // landing zone of this parity
  uint32_t cap, n, base;            // n remote ghosts -> slots [base, base + n)
  // fused push (k_gather<2> stored the ghosts): this launch also SENDS our mail, before it waits for the others'.  The gather has
  // completed (stream order), a system-scope fence by the mailing threads then orders its peer stores before the mail.
  unsigned long long* out_mail[P2P_MAXR];
  int n_out;
  unsigned long long* host_flag;    // mapped host copy of the summed trigger count (nullptr: the caller publishes it itself)
};
__global__ void __launch_bounds__(BLK) k_ghost_land(const __grid_constant__ LandArgs A) {
  __shared__ unsigned long long s_cnt[P2P_MAXR];
  if (blockIdx.x == 0 && (int)threadIdx.x < A.n_out) {
    unsigned long long c = A.count ? *(volatile const unsigned long long*)A.count : 0ull;
    if (c > 0xffffffffull) c = 0xffffffffull;
    const unsigned long long v = ((unsigned long long)A.epoch << 32) | c;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.out_mail[threadIdx.x]), "l"(v) : "memory");
  }
  __syncthreads();   // (block 0: the local count has been read before thread 0 overwrites it with the sum below)
  if ((int)threadIdx.x < A.nranks) {
    unsigned long long v = 0;
    if ((int)threadIdx.x != A.me) {
      for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(A.mail + threadIdx.x) : "memory");
        if ((int32_t)((uint32_t)(v >> 32) - A.epoch) >= 0) break;
        __nanosleep(64);
      }
    }
    s_cnt[threadIdx.x] = v & 0xffffffffull;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0 && A.count) {
    unsigned long long t = *A.count;
    for (int r = 0; r < A.nranks; r++)
      if (r != A.me) t += s_cnt[r];
    *A.count = t;
    if (A.host_flag) { *A.host_flag = t; __threadfence_system(); }
  }
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < A.n; k += gridDim.x * blockDim.x) {
    st4(A.P.pr + A.base + k, ld4(A.land + k));
    st4(A.P.vm + A.base + k, ld4(A.land + (size_t)A.cap + k));
    st4(A.P.wg + A.base + k, ld4(A.land + 2 * (size_t)A.cap + k));
    st4(A.P.q + A.base + k, ld4(A.land + 3 * (size_t)A.cap + k));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Ghost fan-out fused into the gather (k_gather<2>, kernels_step.cuh): every real particle that has images — periodic images in
// the own brick and/or copies in neighbouring bricks — carries a list of destinations; the thread that has just finished the
// particle's prolog stores the four refreshed sectors to each of them from its registers (own ghost slots, or the neighbours'
// landing zones over NVLink).  The mail {epoch, trigger count} is sent by the k_ghost_land launch that follows in the stream.  Per
// step this removes the ghost-refresh / push launches and overlaps the NVLink traffic with the rest of the gather.
// ---------------------------------------------------------------------------------------------------------------
struct FanEnt {
  double sx, sy, sz;   // shift added to the position
  uint32_t dst;        // peer < 0: ghost slot in the own arrays;  peer >= 0: index inside that peer's landing range
  int32_t peer;        // index into FanArgs::peer, or -1
};
struct FanPeer {
  double4* land;       // receiver's landing zone of this epoch, offset to our range; arrays `cap` apart
  uint32_t cap, pad;
};
struct FanArgs {
  const uint32_t* bits;  // one bit per real particle: has images (one word per warp; nullptr: no fan-out in this launch)
  const uint32_t* off;   // nreal + 1
  const FanEnt* ent;
  FanPeer peer[P2P_MAXPEER];
  // several ranks: the LAST block of the gather to finish mails {epoch, local rebuild-trigger count} to every other rank — the
  // neighbours' landing launch finds it already there instead of waiting for this rank's next launch to send it
  unsigned long long* mail[P2P_MAXR];
  int nmail;
  unsigned int* done;    // block counter (zero between launches)
  uint32_t epoch;
};
// Tail of k_gather<2>: every block fences its peer stores at system scope and counts itself; the last one sends the mail.
__device__ __forceinline__ void fan_mail(const FanArgs& F, const unsigned long long* count) {
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(F.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) *F.done = 0;
  if ((int)threadIdx.x < F.nmail) {
    unsigned long long c = count ? *(volatile const unsigned long long*)count : 0ull;
    if (c > 0xffffffffull) c = 0xffffffffull;
    const unsigned long long v = ((unsigned long long)F.epoch << 32) | c;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(F.mail[threadIdx.x]), "l"(v) : "memory");
  }
}
// returns true when the particle was stored to a neighbour GPU (the caller then fences at system scope before the block signals)
__device__ __forceinline__ bool fan_out(const FanArgs& F, const PView& P, uint32_t p, const double4& r, const double4& vm, const double4& wg,
                                        const double4& q) {
  bool remote = false;
  const uint32_t e1 = F.off[p + 1];
  for (uint32_t e = F.off[p]; e < e1; e++) {
    const FanEnt d = F.ent[e];
    const double4 rs = make_double4(r.x + d.sx, r.y + d.sy, r.z + d.sz, r.w);
    if (d.peer < 0) {
      st4(P.pr + d.dst, rs); st4(P.vm + d.dst, vm); st4(P.wg + d.dst, wg); st4(P.q + d.dst, q);
    } else {
      const FanPeer& t = F.peer[d.peer];
      st4(t.land + d.dst, rs); st4(t.land + (size_t)t.cap + d.dst, vm); st4(t.land + 2 * (size_t)t.cap + d.dst, wg);
      st4(t.land + 3 * (size_t)t.cap + d.dst, q);
      remote = true;
    }
  }
  return remote;
}
// fan-out table construction (rebuild): count / fill over the local image list and over each peer's send list
template <bool FILL>
__global__ void k_fan_build(const uint32_t* __restrict__ src, const uint32_t* __restrict__ ent, const double* __restrict__ shift, uint32_t n,
                            uint32_t dst_base, int peer, uint32_t* __restrict__ cursor, const uint32_t* __restrict__ off, FanEnt* __restrict__ out,
                            uint32_t* __restrict__ bits) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t s = src[k];
  const uint32_t c = atomicAdd(&cursor[s], 1u);
  if (!FILL && c == 0) atomicOr(&bits[s >> 5], 1u << (s & 31));
  if (FILL) {
    const uint32_t e = ent[k];
    out[off[s] + c] = FanEnt{shift[3 * e], shift[3 * e + 1], shift[3 * e + 2], dst_base + k, peer};
  }
}

struct Peer {
  int rank;
  std::vector<GhostEntry> recv;  // our ghost cells fed by this peer (our table order)
  std::vector<GhostEntry> send;  // the peer's ghost cells fed by us (the peer's table order)
  DevBuf<uint32_t> d_rgcell, d_rcnt, d_roff, d_scell, d_scnt, d_soff, d_smap, d_sent;
  DevBuf<double> d_sshift;
  uint32_t n_send = 0, n_recv = 0, recv_base = 0;  // particles; recv_base = offset inside the ghost region
  uint32_t remote_base = 0, remote_cap = 0;         // P2P: where our data lands in the peer's landing zone / its array stride
  int fan_index = -1;                                // slot in FanArgs::peer
  DevBuf<double4> s_pr, s_vm, s_wg, s_q;
  DevBuf<double> s_h;
  DevBuf<uint64_t> s_id;
};

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int pg[3] = {1, 1, 1}, rc[3] = {0, 0, 0};
  std::vector<double> cuts[3];   // per axis pg[d] + 1 cut positions (edb_set_brick_cuts); empty = uniform bricks
  uint32_t n_local_ghost = 0;
  std::map<int, Peer> peers;
  DevBuf<unsigned long long> d_red;
  DevBuf<uint32_t> d_gather;
  // NVLink peer-memory ghost refresh (see k_ghost_push / k_ghost_land)
  struct P2P {
    bool on = false, tried = false;
    std::string why;                 // why it is off
    char* base = nullptr;            // own mailbox allocation
    uint32_t cap = 0;                // landing capacity (ghost particles per parity)
    char* peer_base[P2P_MAXR] = {};  // opened peer allocations
    uint32_t peer_cap[P2P_MAXR] = {};
    cudaIpcMemHandle_t peer_handle[P2P_MAXR] = {};
    uint32_t epoch = 0;
    unsigned int* d_done = nullptr;
    uint64_t exchanges = 0;
  } p2p;
  static size_t p2p_mail_bytes() { return 2 * P2P_MAXR * sizeof(unsigned long long); }
  static size_t p2p_bytes(uint32_t cap) { return p2p_mail_bytes() + 2 * 4 * (size_t)cap * sizeof(double4); }
  static unsigned long long* p2p_mail(char* b, int parity) { return (unsigned long long*)b + parity * P2P_MAXR; }
  static double4* p2p_land(char* b, uint32_t cap, int parity) { return (double4*)(b + p2p_mail_bytes()) + (size_t)parity * 4 * cap; }
  void p2p_close() {
    for (int r = 0; r < P2P_MAXR; r++)
      if (p2p.peer_base[r]) { cudaIpcCloseMemHandle(p2p.peer_base[r]); p2p.peer_base[r] = nullptr; }
  }
  // Collective, called at every rebuild once the ghost layout of all ranks is known: (re)allocates the mailbox when the ghost
  // count outgrew it, exchanges IPC handles + landing offsets (one NCCL all-gather), opens what changed.
  void p2p_prepare(uint32_t n_remote_ghost, cudaStream_t st) {
    if (!comm || nranks > P2P_MAXR || (int)peers.size() > P2P_MAXPEER) { p2p.on = false; p2p.why = "rank/peer count over the mailbox size"; return; }
    if (p2p.tried && !p2p.on) return;
    if (getenv("EDB_NO_P2P")) { p2p.tried = true; p2p.on = false; p2p.why = "EDB_NO_P2P set"; return; }
    CK(cudaStreamSynchronize(st));   // every push/land of the previous layout has completed on this rank
    bool ok = true;
    // 1. capacity: a rank reallocates only when it must; the others learn the new handle from the all-gather
    bool fresh = false;
    if (!p2p.base || n_remote_ghost > p2p.cap) {
      if (p2p.base) {
        // peers may still map the old allocation: it is released after the all-gather barrier below, when nobody pushes to it
      }
      const uint32_t ncap = n_remote_ghost + n_remote_ghost / 4 + 4096;
      char* nb = nullptr;
      if (cudaMalloc(&nb, p2p_bytes(ncap)) != cudaSuccess) { cudaGetLastError(); ok = false; }
      else {
        CK(cudaMemset(nb, 0, p2p_bytes(ncap)));
        old_base = p2p.base;
        p2p.base = nb;
        p2p.cap = ncap;
        fresh = true;
      }
      if (!p2p.d_done) { CK(cudaMalloc(&p2p.d_done, 4)); CK(cudaMemset(p2p.d_done, 0, 4)); }
    }
    // 2. blob: [ok][cap][handle 16 u32][recv_base per rank]
    const uint32_t nb32 = 2 + 16 + (uint32_t)nranks;
    std::vector<uint32_t> mine(nb32, 0), all((size_t)nb32 * nranks, 0);
    cudaIpcMemHandle_t h{};
    if (ok && cudaIpcGetMemHandle(&h, p2p.base) != cudaSuccess) { cudaGetLastError(); ok = false; }
    mine[0] = ok ? 1u : 0u;
    mine[1] = p2p.cap;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(&mine[2], &h, 64);
    for (auto& kv : peers) mine[18 + kv.first] = kv.second.recv_base;
    allgather_u32(mine.data(), nb32, all.data(), st);
    for (int r = 0; r < nranks; r++) ok = ok && all[(size_t)r * nb32] == 1u;
    // 3. open what is new
    if (ok)
      for (int r = 0; r < nranks && ok; r++) {
        if (r == rank) continue;
        cudaIpcMemHandle_t hr;
        memcpy(&hr, &all[(size_t)r * nb32 + 2], 64);
        if (p2p.peer_base[r] && memcmp(&hr, &p2p.peer_handle[r], 64) == 0) continue;
        if (p2p.peer_base[r]) { cudaIpcCloseMemHandle(p2p.peer_base[r]); p2p.peer_base[r] = nullptr; }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        p2p.peer_base[r] = (char*)ptr;
        p2p.peer_handle[r] = hr;
        p2p.peer_cap[r] = all[(size_t)r * nb32 + 1];
      }
    // 4. agree (a second tiny all-gather doubles as the barrier after which replaced allocations can be freed)
    std::vector<uint32_t> okv(nranks, 0);
    const uint32_t myok = ok ? 1u : 0u;
    allgather_u32(&myok, 1, okv.data(), st);
    for (int r = 0; r < nranks; r++) ok = ok && okv[r] == 1u;
    if (old_base) { cudaFree(old_base); old_base = nullptr; }
    (void)fresh;
    p2p.tried = true;
    p2p.on = ok;
    if (!ok) { p2p.why = "CUDA IPC peer mapping unavailable"; p2p_close(); return; }
    for (auto& kv : peers) {
      Peer& p = kv.second;
      p.remote_base = all[(size_t)p.rank * nb32 + 18 + rank];
      p.remote_cap = p2p.peer_cap[p.rank];
    }
  }
  char* old_base = nullptr;
  // fused path: the gather of this step pushes (FanArgs) under a new epoch; the NEXT ghost refresh only lands
  uint32_t fan_epoch_pending = 0;
  bool fan_mailed = false;   // the pending epoch's mail was sent by the gather itself
  void fan_begin(FanArgs& F, bool mail_from_gather) {
    if (!(comm && p2p.on)) return;
    const uint32_t e = ++p2p.epoch;
    const int par = (int)(e & 1u);
    for (auto& kv : peers) {
      Peer& p = kv.second;   // fan_index was fixed when the fan-out table was built (rebuild)
      if (p.fan_index >= 0) F.peer[p.fan_index] = FanPeer{p2p_land(p2p.peer_base[p.rank], p.remote_cap, par) + p.remote_base, p.remote_cap, 0};
    }
    fan_epoch_pending = e;
    fan_mailed = mail_from_gather;
    F.nmail = 0;
    if (mail_from_gather) {
      for (int r = 0; r < nranks; r++)
        if (r != rank) F.mail[F.nmail++] = p2p_mail(p2p.peer_base[r], par) + rank;
      F.done = p2p.d_done;
      F.epoch = e;
    }
  }
  const double4* landing_of(uint32_t epoch) const { return p2p_land(p2p.base, p2p.cap, (int)(epoch & 1u)); }
  // the remote ghosts of `epoch` (already landed: its mail was waited for) copied into the ghost slots; no mail, no count
  void land_copy(PView P, uint32_t nreal, uint32_t epoch, cudaStream_t st, uint64_t* launches) {
    uint32_t nrem = 0;
    for (auto& kv : peers) nrem += kv.second.n_recv;
    if (!nrem) return;
    LandArgs L{};
    L.P = P;
    L.mail = p2p_mail(p2p.base, (int)(epoch & 1u));
    L.nranks = nranks; L.me = rank; L.epoch = epoch; L.count = nullptr;
    L.land = landing_of(epoch);
    L.cap = p2p.cap; L.n = nrem; L.base = nreal + n_local_ghost;
    k_ghost_land<<<std::min<uint32_t>((nrem + BLK - 1) / BLK, 592u), BLK, 0, st>>>(L);
    (*launches)++;
  }
  // copy = false: only the mail is exchanged (our mail out, everybody's in, trigger counts summed); the caller reads the ghosts in
  // the landing zone.  Returns the epoch that landed.
  uint32_t land_pending(PView P, uint32_t nreal, unsigned long long* count, cudaStream_t st, uint64_t* launches, bool copy = true,
                        unsigned long long* host_flag = nullptr) {
    const uint32_t e = fan_epoch_pending;
    fan_epoch_pending = 0;
    const int par = (int)(e & 1u);
    uint32_t nrem = 0;
    for (auto& kv : peers) nrem += kv.second.n_recv;
    LandArgs L{};
    L.P = P;
    L.mail = p2p_mail(p2p.base, par);
    L.nranks = nranks; L.me = rank; L.epoch = e; L.count = count;
    L.land = p2p_land(p2p.base, p2p.cap, par);
    L.cap = p2p.cap; L.n = copy ? nrem : 0; L.base = nreal + n_local_ghost;
    L.host_flag = count ? host_flag : nullptr;
    if (!fan_mailed)
      for (int r = 0; r < nranks; r++)
        if (r != rank) L.out_mail[L.n_out++] = p2p_mail(p2p.peer_base[r], par) + rank;
    fan_mailed = false;
    const uint32_t nb = L.n ? std::min<uint32_t>((L.n + BLK - 1) / BLK, 592u) : 1u;
    k_ghost_land<<<nb, BLK, 0, st>>>(L);
    (*launches)++;
    p2p.exchanges++;
    return e;
  }
  // per-step refresh over peer memory; `count` (device, nullable) is all-reduced (sum) in place as a by-product
  void exchange_ghosts_p2p(PView P, uint32_t nreal, unsigned long long* count, cudaStream_t st, uint64_t* launches) {
    const uint32_t e = ++p2p.epoch;
    const int par = (int)(e & 1u);
    PushArgs A{};
    A.P = P;
    uint32_t tot = 0, nrem = 0;
    for (auto& kv : peers) {
      Peer& p = kv.second;
      nrem += p.n_recv;
      if (!p.n_send) continue;
      PushPeer& q = A.peer[A.npeer++];
      q.smap = p.d_smap.p; q.sent = p.d_sent.p; q.sshift = p.d_sshift.p;
      q.start = tot; q.n = p.n_send;
      q.cap = p.remote_cap;
      q.land = p2p_land(p2p.peer_base[p.rank], p.remote_cap, par) + p.remote_base;
      tot += p.n_send;
    }
    A.total = tot;
    for (int r = 0; r < nranks; r++)
      if (r != rank) A.mail[A.nmail++] = p2p_mail(p2p.peer_base[r], par) + rank;
    A.done = p2p.d_done;
    A.count = count;
    A.epoch = e;
    k_ghost_push<<<tot ? (tot + BLK - 1) / BLK : 1, BLK, 0, st>>>(A);
    LandArgs L{};
    L.P = P;
    L.mail = p2p_mail(p2p.base, par);
    L.nranks = nranks; L.me = rank; L.epoch = e; L.count = count;
    L.land = p2p_land(p2p.base, p2p.cap, par);
    L.cap = p2p.cap; L.n = nrem; L.base = nreal + n_local_ghost;
    const uint32_t nb = nrem ? std::min<uint32_t>((nrem + BLK - 1) / BLK, 592u) : 1u;
    k_ghost_land<<<nb, BLK, 0, st>>>(L);
    (*launches) += 2;
    p2p.exchanges++;
  }

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
    p2p_close();
    if (p2p.base) { cudaFree(p2p.base); p2p.base = nullptr; }
    if (p2p.d_done) { cudaFree(p2p.d_done); p2p.d_done = nullptr; }
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    comm = nullptr;
    release_peers();
    d_red.release();
    d_gather.release();
  }
  void release_peers() {
    for (auto& kv : peers) {
      Peer& p = kv.second;
      p.d_rgcell.release(); p.d_rcnt.release(); p.d_roff.release(); p.d_scell.release(); p.d_scnt.release(); p.d_soff.release();
      p.d_smap.release(); p.d_sent.release(); p.d_sshift.release(); p.s_pr.release(); p.s_vm.release(); p.s_wg.release(); p.s_q.release();
      p.s_h.release(); p.s_id.release();
    }
    peers.clear();
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
  // real cells along axis d of the bricks at coordinate `coord` of that axis (same rule as setup_grid; uniform bricks share ours)
  int nc_axis(const edb_config& cfg, const GridDev& G, int d, int coord) const {
    if (cuts[d].empty()) return G.nc[d];
    const double len = cuts[d][coord + 1] - cuts[d][coord];
    return std::max(1, (int)std::floor(len / cfg.cell_size + 1e-9));
  }
  // source of ghost cell idx[] of the brick at coords `who` : returns false when there is none (non-periodic wall).  src[] is a cell
  // of the SOURCE brick's grid, src_g[] that grid's dimensions (with the ghost shell)
  bool source_of(const edb_config& cfg, const GridDev& G, const int who[3], const int idx[3], int src[3], int src_g[3], double shift[3],
                 int* peer) const {
    int rb[3];
    for (int d = 0; d < 3; d++) {
      rb[d] = who[d]; src[d] = idx[d]; shift[d] = 0.0;
      const int gw = nc_axis(cfg, G, d, who[d]) + 2;
      if (idx[d] == 0) {
        rb[d] = who[d] - 1;
        if (rb[d] < 0) {
          if (!cfg.periodic[d]) return false;
          rb[d] = pg[d] - 1; shift[d] = -(cfg.ghi[d] - cfg.glo[d]);
        }
        src[d] = nc_axis(cfg, G, d, rb[d]);
      } else if (idx[d] == gw - 1) {
        rb[d] = who[d] + 1; src[d] = 1;
        if (rb[d] >= pg[d]) {
          if (!cfg.periodic[d]) return false;
          rb[d] = 0; shift[d] = cfg.ghi[d] - cfg.glo[d];
        }
      }
      src_g[d] = nc_axis(cfg, G, d, rb[d]) + 2;
    }
    *peer = rb[0] + pg[0] * (rb[1] + pg[1] * rb[2]);
    return true;
  }
  bool ghost_source(const edb_config& cfg, const GridDev& G, const int idx[3], int src[3], int src_g[3], double shift[3], int* peer) const {
    return source_of(cfg, G, rc, idx, src, src_g, shift, peer);
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
      int who[3], gw[3];
      coords(pg, pr, who);
      for (int d = 0; d < 3; d++) gw[d] = nc_axis(cfg, G, d, who[d]) + 2;   // the RECEIVER's grid
      for (int k = 0; k < gw[2]; k++)
        for (int j = 0; j < gw[1]; j++)
          for (int i = 0; i < gw[0]; i++) {
            if (!(i == 0 || j == 0 || k == 0 || i == gw[0] - 1 || j == gw[1] - 1 || k == gw[2] - 1)) continue;
            const int idx[3] = {i, j, k};
            int src[3], src_g[3], peer;
            double shift[3];
            if (!source_of(cfg, G, who, idx, src, src_g, shift, &peer)) continue;
            if (peer != rank) continue;
            if (src_g[0] != G.g[0] || src_g[1] != G.g[1] || src_g[2] != G.g[2])
              throw StateError{EDB_ERR_INVALID, "brick cuts: this rank's bounds do not match the cuts (cell grids disagree)"};
            Peer& p = peers[pr];
            p.rank = pr;
            const uint32_t cg = (uint32_t)i + gw[0] * ((uint32_t)j + gw[1] * (uint32_t)k);
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
  // driver reactions (force_to_accel_driver.cpp:128-139: MPI_Allreduce of 6 doubles per driver), in place on the stream
  void allreduce_sum_f64_dev(double* d, size_t n, cudaStream_t st) {
    if (!comm || n == 0) return;
    NCK(g_nccl.AllReduce(d, d, n, ncclFloat64, ncclSum, comm, st));
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
