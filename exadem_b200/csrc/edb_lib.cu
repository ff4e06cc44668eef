// edb_lib.cu — libexadem_b200.so: context management + the C ABI of include/exadem_b200.h.
// All compute is in the CUDA kernels of kernels_rebuild.cuh / kernels_step.cuh (sm_100a).  No CPU fallback.
#include <algorithm>
#include <cstring>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels_rebuild.cuh"
#include "kernels_step.cuh"
#include "kernels_nbh_poly.cuh"
#include "kernels_migrate.cuh"
#include "kernels_twin.cuh"
#include "kernels_driver.cuh"
#include "kernels_bond.cuh"
#include "kernels_bond_rebuild.cuh"
#include <cub/device/device_radix_sort.cuh>
#include "comm.cuh"

using namespace edb;

static thread_local std::string g_create_error;
static_assert(sizeof(edb_config) == 152 && sizeof(edb_interaction) == 104 && sizeof(edb_contact_params) == 64 &&
                  sizeof(edb_particles) == 8 + 31 * 8 && sizeof(edb_driver) == 184 && sizeof(edb_driver_motion) == 400 &&
                  sizeof(edb_inner_bond) == 128 && sizeof(edb_inner_bond_params) == 48,
              "C ABI struct layout changed: update include/exadem_b200.h users and exadem_b200/capi.py");

// wave order of the contact operators (classifier_for_all.hpp:116-156 as called by contact_force_sphere.cu.in:132-188 and
// contact_force_polyhedron.cu.in:148-245)
static const int ORDER_SPHERE[] = {0, 4, 5, 6, 7, 8, 9};
static const int ORDER_POLY[] = {0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 4, 5, 6};

enum Bucket { B_CONTACT = 0, B_GATHER, B_PROLOG, B_EPILOG, B_REBUILD, B_GHOST, B_CONTACT_PP, B_N };

struct edb_context {
  edb_config cfg{};
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;          // side stream: driver / mesh waves of the sphere operator run beside the wave-0 kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  int law = 0;
  bool timing = false;

  GridDev G{};
  // particles
  PArr pa, pb;            // current / scratch (double-buffered for the cell sort)
  uint32_t nreal = 0, nghost = 0;
  uint32_t n_old_real = 0;  // #real slots the `prev` waves refer to
  DevBuf<uint32_t> cidx, cidx2, ccount, cstart, cfill, tmp, new2old, old2new, cbeg, cend;
  DevBuf<uint32_t> scan_bsum, scan_total;
  DevBuf<double4> pr_backup, q_backup, vert;
  DevBuf<RegionDev> d_regions;       // particle regions of the `region` slots (edb_set_particle_regions)
  int n_regions = 0;
  DevBuf<uint16_t> ptype;            // polyhedra: shape index per slot (reals + ghosts), refreshed with the world vertices
  bool have_backup = false;
  int shape_nemax = 0;               // largest edge count over the shapes (the cell-block neighbour kernel holds edge sets in 32-bit masks)
  bool have_cells = false;           // cbeg / cend / cidx describe the current particle set (set by the rebuild)
  // ghost table (host-built, static for a given grid)
  uint32_t n_gent = 0;
  DevBuf<uint32_t> g_cell, g_src_cell, g_count, g_off, g_src, g_ent;
  DevBuf<double> g_shift;
  // interactions
  Wave wave[NT], prev[NT];
  bool have_prev = false;
  DevBuf<uint32_t> inc_cnt, inc_off, inc;
  bool have_inc = false;
  bool queue_clean = false;   // d_need_count is zero (left so by k_gather<2>)
  // ghost fan-out fused into k_gather<2> (comm.cuh): table built at rebuild; ghost_pushed = the last gather refreshed the ghosts
  DevBuf<uint32_t> fan_cnt, fan_off, fan_bits;
  // polyhedron neighbour build: stream pool of the count pass (kernels_nbh_poly.cuh)
  DevBuf<uint32_t> nbh_pool, nbh_first, nbh_ctr;
  DevBuf<unsigned long long> rmax_dev;   // bits of the largest pr.w over reals + ghosts (k_max_radius), read by for_each_neighbor
  DevBuf<uint2> pp_tmp;              // spheres: candidates recorded by the count pass of the neighbour build (kernels_rebuild.cuh)
  uint32_t nbh_pool_chunks = 0;
  uint32_t* h_nbh = nullptr;   // pinned copy of nbh_ctr
  uint64_t nbh_replays = 0;
  DevBuf<FanEnt> fan_ent;
  bool have_fan = false, ghost_pushed = false;
  bool flag_published = false;     // the landing launch of this step already stored the trigger count into h_sflag
  uint32_t ghost_in_landing = 0;   // epoch whose remote ghosts live in the landing zone only (0: the ghost slots are current)
  // debugging switches read once at creation: EDB_NO_FAN (separate ghost kernels), EDB_NBH_POLY_PER_THREAD, EDB_NBH_NO_REPLAY
  bool env_no_fan = false, env_nbh_per_thread = false, env_nbh_no_replay = false;
  bool env_nbh_warp = false;         // EDB_NBH_WARP: spheres count with a warp per particle (measured 2.40 ms per drum rebuild against 1.78 ms for a thread per particle: off)
  bool env_nbh_no_cull = false;      // EDB_NBH_NO_CULL: visit all 27 stencil cells per particle (no cell-level reach test)
  bool env_nbh_two_pass = false;     // EDB_NBH_TWO_PASS: spheres enumerate twice (count, fill) like round 1
  bool env_gather_mail = false;       // EDB_GATHER_MAIL: the gather's last block sends the mail instead of the landing launch
  bool env_land_copy = false;         // EDB_LAND_COPY: always copy the landing zone into the ghost slots (the round-1 path)
  bool env_drivers_unfused = false;   // EDB_DRIVERS_UNFUSED: moving drivers through the operator-by-operator loop on one rank too
  uint32_t* d_need_count = nullptr;  // chunk queue of the persistent sphere contact kernel / length of the polyhedron work list
  DevBuf<uint32_t> need;             // work list of k_detect_poly_pp -> k_law_poly_pp
  unsigned long long* d_sflag = nullptr;   // rebuild trigger of the speculative step loop (2 slots, device)
  unsigned long long* h_sflag = nullptr;   // ... pinned host copies (mapped: the device stores them, no copy engine involved)
  unsigned long long* h_sflag_dev = nullptr;
  cudaEvent_t ev_sflag[2] = {};
  int n_sm = 148;
  size_t smem_optin = 99 * 1024;   // cudaDevAttrMaxSharedMemoryPerBlockOptin
  double rcut_inc = 0;
  int symetric = 1;
  double mean_face_nv = 3;  // mean vertex count of the particle shapes' faces (algorithmic bytes of the vertex-face waves)
  double rmax_upload = 0;   // largest `radius` given to edb_upload_particles (check_rcut.cpp:72-100)
  // parameters
  ParamsDev prm{};
  DevBuf<Params> tpp, tpd;
  std::vector<Driver> h_drv;
  DevBuf<Driver> drv;
  // shapes
  std::vector<ShapeDev> h_sh;
  DevBuf<ShapeDev> d_sh;
  DevBuf<v3> d_sv;
  DevBuf<int> d_se, d_sfoff, d_sfn, d_sfidx;
  ShapesView S{};
  // RShape (mesh) drivers
  std::vector<edb_shape> h_mesh_shape;            // as given (pointers re-homed into the pools below)
  std::vector<std::vector<double>> h_mesh_v;      // local vertices per mesh shape
  std::vector<std::vector<int>> h_mesh_e, h_mesh_foff, h_mesh_fn, h_mesh_fidx;
  std::vector<double> h_mesh_radius;
  std::vector<MeshDev> h_md;                      // one per RSHAPE driver
  std::vector<double4> h_mverts;                  // pooled world vertices
  std::vector<int> h_medges, h_mfoff, h_mfn, h_mfidx;
  DevBuf<MeshDev> d_md;
  DevBuf<double4> d_mverts;
  DevBuf<int> d_medges, d_mfoff, d_mfn, d_mfidx, d_mcell_data;
  DevBuf<uint32_t> d_mcell_off;
  MeshView M{};
  bool mesh_dirty = false;
  // device-resident `drivers` slot: motions + integrator state (kernels_driver.cuh), backup for driver_displ_over
  std::vector<DriverMotion> h_mo;
  std::vector<int> h_drv_shape;                   // per driver: the caller's mesh-shape index (Driver::shape holds the mesh slot)
  DevBuf<DriverMotion> mo, mo_backup;
  DevBuf<Driver> drv_backup;
  DevBuf<double4> d_mlocal;                       // pooled local mesh vertices (same offsets as d_mverts)
  std::vector<double4> h_mlocal;
  DevBuf<uint32_t> d_mprim_start, d_mcell_cnt;
  DevBuf<double> drv_sums;
  bool drivers_moving = false, have_drv_backup = false;
  bool drivers_need_reaction = false, mesh_moving = false;   // set with the motions
  // inner bonds (wave 13, kernels_bond.cuh): host mirror in device order + SoA on the device
  std::vector<edb_inner_bond> h_bonds;
  bool bonds_on_device = false;                   // the dynamic fields (friction, tds, en, et, unbroken) of h_bonds are stale
  DevBuf<uint2> b_ij, b_sub, b_itf;
  DevBuf<double4> b_par, b_fr, b_tds, b_fd, b_cp;
  DevBuf<uint8_t> b_mode, b_unbroken, b_itf_broken;
  DevBuf<ulonglong2> b_ids;                       // (id_i, id_j) per bond
  DevBuf<uint8_t> b_ghost;
  // second set + scratch of the device-side rebuild (kernels_bond_rebuild.cuh)
  DevBuf<uint2> b2_ij, b2_sub;
  DevBuf<double4> b2_par, b2_fr, b2_tds, b2_fd, b2_cp;
  DevBuf<uint8_t> b2_mode, b2_unbroken, b2_ghost, cub_tmp;
  DevBuf<ulonglong2> b2_ids;
  DevBuf<uint32_t> b_key, b_key2, b_idx, b_idx2, b_fl_ign, b_rk_ign, b_fl_itf, b_rk_itf, b_ikey, b_ikey2, b_ival, b_ist, b_irk, b_cnt;
  uint32_t n_conv_dev = 0;                        // broken bonds queued on the device as vertex-vertex history (d_conv)
  bool env_bonds_unfused = false;                 // EDB_BONDS_UNFUSED: edb_run_verlet with bonds runs the operator-by-operator loop
  bool env_bonds_host = false;                    // EDB_BONDS_HOST: bonds go through the host at every rebuild (the round-1 path)
  DevBuf<uint32_t> b_plist, b_inc_off, b_inc;
  uint32_t n_bond_particles = 0, n_interfaces = 0;
  DevBuf<uint32_t> b_ign_off;
  DevBuf<unsigned long long> b_ign_ids;
  bool have_ignore = false;
  std::vector<double> h_farea;                    // face areas of the particle shapes at homothety 1, pooled by ShapeDev::f0
  std::vector<BondConv> h_conv;                   // broken bonds waiting to become vertex-vertex history at this rebuild
  DevBuf<BondConv> d_conv;
  DevBuf<BondParams> d_ibp;
  int ibp_groups = 0;
  double physical_time = 0, system_mass = 0;
  bool have_system_mass = false;
  // asynchronous host <-> device path of a host-owned grid (edb_upload_dynamic_async / edb_download_dynamic_async): own copy
  // streams so that the download of step n, the upload of step n+1 and the kernels of step n+1 overlap on a full-duplex link
  cudaStream_t st_up = nullptr, st_dn = nullptr;
  cudaStream_t st_flag = nullptr;      // high-priority side stream that publishes the rebuild trigger while asynchronous host copies run
  cudaEvent_t ev_flag_src = nullptr;
  cudaEvent_t ev_up_done = nullptr, ev_up_consumed = nullptr, ev_dn_ready[2] = {}, ev_dn_done[2] = {}, ev_field[22] = {};
  bool up_consumed_valid = false, dn_done_valid[2] = {false, false}, field_valid[22] = {};
  int dn_parity = 0;
  // EDB_TRACE_ASYNC: timed events at the ends of every upload / step / download of the asynchronous copy pipeline, printed at destroy
  bool env_flag_on_stream = false;   // EDB_FLAG_ON_STREAM: the round-2 first form (8-byte copy on the compute stream), for comparison
  bool trace_async = false;
  std::vector<std::pair<const char*, cudaEvent_t>> trace_ev;
  DevBuf<double> stage_up, stage_dn[2];
  // staging for host transfers
  DevBuf<double> stage;
  DevBuf<uint64_t> stage_u64;
  DevBuf<uint32_t> stage_u32;
  // pinned scalars
  unsigned long long* h_flag = nullptr;
  unsigned long long* d_flag = nullptr;
  uint32_t* h_tot = nullptr;  // pinned, NT+4 entries
  // counters / timing
  edb_counters ctr{};
  std::vector<EventPair> ev_pending, ev_free;
  // L2 flush scratch
  DevBuf<double4> l2buf;
  cudaEvent_t user_ev[8] = {};
  // multi-GPU
  Comm comm;
  DevBuf<int> mig_dest;
  DevBuf<uint32_t> mig_midx, mig_cnt, mig_off, imm_cnt, imm_off;
  DevBuf<MigPart> mig_sp, mig_rp;
  DevBuf<MigRec> mig_sh, mig_rh, imm_tab;
  uint32_t n_imm = 0;
  uint64_t migrated_out = 0, migrated_in = 0;
  // update_interaction_ghost (PartnerGhost twins)
  DevBuf<uint32_t> tw_cnt, tw_flag, tw_pos, tw_own, tw_par, tw_tpos;
  DevBuf<TwinMsg> tw_s, tw_r;
  uint64_t twins = 0;

  bool poly() const { return cfg.mode == EDB_MODE_POLYHEDRON; }
};

#define LAUNCH_ON(ctx, stream, kern, grid, block, smem, ...)       \
  do {                                                            \
    if ((grid) > 0) {                                             \
      kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);   \
      (ctx)->ctr.kernel_launches++;                               \
    }                                                             \
  } while (0)
#define LAUNCH(ctx, kern, grid, block, smem, ...)                 \
  do {                                                            \
    if ((grid) > 0) {                                             \
      kern<<<(grid), (block), (smem), (ctx)->st>>>(__VA_ARGS__);  \
      (ctx)->ctr.kernel_launches++;                               \
    }                                                             \
  } while (0)

static inline uint32_t nblk(uint64_t n, int b = BLK) { return (uint32_t)((n + b - 1) / b); }

struct ScopedTimer {
  edb_context* c;
  EventPair ep;
  bool on;
  ScopedTimer(edb_context* c_, int bucket) : c(c_), on(c_->timing) {
    if (!on) return;
    if (!c->ev_free.empty()) { ep = c->ev_free.back(); c->ev_free.pop_back(); }
    else { CK(cudaEventCreate(&ep.a)); CK(cudaEventCreate(&ep.b)); }
    ep.bucket = bucket;
    CK(cudaEventRecord(ep.a, c->st));
  }
  ~ScopedTimer() {
    if (!on) return;
    cudaEventRecord(ep.b, c->st);
    c->ev_pending.push_back(ep);
  }
};

static void resolve_timers(edb_context* c) {
  if (c->ev_pending.empty()) return;
  CK(cudaStreamSynchronize(c->st));
  for (auto& ep : c->ev_pending) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ep.a, ep.b));
    double* b[B_N] = {&c->ctr.contact_ms, &c->ctr.gather_ms, &c->ctr.prolog_ms, &c->ctr.epilog_ms, &c->ctr.rebuild_ms, &c->ctr.ghost_ms, &c->ctr.contact_pp_ms};
    *b[ep.bucket] += ms;
    c->ev_free.push_back(ep);
  }
  c->ev_pending.clear();
}

// exclusive scan of n uint32 (out has n+1 entries); returns nothing, total stays on device in out[n]
static void scan_u32(edb_context* c, const uint32_t* in, uint32_t* out, uint32_t n) {
  const uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  c->scan_bsum.ensure(nb + 1);
  c->scan_total.ensure(1);
  if (n == 0) {
    CK(cudaMemsetAsync(out, 0, sizeof(uint32_t), c->st));
    return;
  }
  LAUNCH(c, k_scan_reduce, nb, SCAN_T, 0, in, c->scan_bsum.p, n);
  LAUNCH(c, k_scan_bsum, 1, SCAN_T, 0, c->scan_bsum.p, nb, c->scan_total.p);
  LAUNCH(c, k_scan_apply, nb, SCAN_T, 0, in, out, c->scan_bsum.p, c->scan_total.p, n);
}

static int fail(edb_context* c, int code, const std::string& m) {
  if (c) c->err = m;
  else g_create_error = m;
  return code;
}

#define API_BEGIN(ctx)                                                \
  if (!(ctx)) return EDB_ERR_INVALID;                                 \
  try {                                                               \
    CK(cudaSetDevice((ctx)->cfg.device));
#define API_END(ctx)                                                                                            \
  }                                                                                                             \
  catch (const CudaError& e) {                                                                                  \
    char b[512];                                                                                                \
    snprintf(b, sizeof b, "CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e.e), cudaGetErrorString(e.e), \
             e.file, e.line, e.what);                                                                           \
    return fail((ctx), EDB_ERR_CUDA, b);                                                                        \
  }                                                                                                             \
  catch (const StateError& e) { return fail((ctx), e.code, e.msg); }                                            \
  catch (const std::exception& e) { return fail((ctx), EDB_ERR_INVALID, e.what()); }                            \
  return EDB_OK;

// ---------------------------------------------------------------------------------------------------------------
static void setup_grid(edb_context* c) {
  GridDev& G = c->G;
  for (int d = 0; d < 3; d++) {
    const double len = c->cfg.hi[d] - c->cfg.lo[d];
    if (!(len > 0) || !(c->cfg.cell_size > 0)) throw StateError{EDB_ERR_INVALID, "empty domain or cell_size <= 0"};
    G.nc[d] = std::max(1, (int)std::floor(len / c->cfg.cell_size + 1e-9));
    G.cs[d] = len / G.nc[d];
    G.g[d] = G.nc[d] + 2;
    G.lo[d] = c->cfg.lo[d];
    G.glo[d] = c->cfg.glo[d];
    G.ghi[d] = c->cfg.ghi[d];
    G.periodic[d] = c->cfg.periodic[d];
  }
  G.ncell = (uint32_t)G.g[0] * G.g[1] * G.g[2];
  // ghost table: every shell cell mirrors a source cell of some brick (possibly our own, shifted by the period)
  std::vector<uint32_t> gc, gs;
  std::vector<double> sh;
  c->comm.plan_reset(c->cfg, G);
  for (int k = 0; k < G.g[2]; k++)
    for (int j = 0; j < G.g[1]; j++)
      for (int i = 0; i < G.g[0]; i++) {
        const int idx[3] = {i, j, k};
        if (!(i == 0 || j == 0 || k == 0 || i == G.g[0] - 1 || j == G.g[1] - 1 || k == G.g[2] - 1)) continue;
        int src[3], src_g[3];
        double shift[3];
        int peer;
        if (!c->comm.ghost_source(c->cfg, G, idx, src, src_g, shift, &peer)) continue;
        const uint32_t cg = (uint32_t)i + G.g[0] * ((uint32_t)j + G.g[1] * (uint32_t)k);
        const uint32_t csrc = (uint32_t)src[0] + src_g[0] * ((uint32_t)src[1] + src_g[1] * (uint32_t)src[2]);   // in the source brick's grid
        if (peer == c->cfg.rank) {
          gc.push_back(cg); gs.push_back(csrc);
          sh.push_back(shift[0]); sh.push_back(shift[1]); sh.push_back(shift[2]);
        } else {
          c->comm.plan_add_remote(peer, cg, csrc, shift);
        }
      }
  c->n_gent = (uint32_t)gc.size();
  c->g_cell.ensure(gc.size() + 1); c->g_src_cell.ensure(gc.size() + 1); c->g_shift.ensure(sh.size() + 3);
  c->g_count.ensure(gc.size() + 2); c->g_off.ensure(gc.size() + 2);
  if (!gc.empty()) {
    CK(cudaMemcpy(c->g_cell.p, gc.data(), gc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->g_src_cell.p, gs.data(), gs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->g_shift.p, sh.data(), sh.size() * 8, cudaMemcpyHostToDevice));
  }
  if (c->cfg.nranks > 1) c->comm.plan_send_lists(c->cfg, G);
  c->comm.plan_finalize();
  c->ccount.ensure(G.ncell + 1); c->cstart.ensure(G.ncell + 2); c->cfill.ensure(G.ncell + 1);
  c->cbeg.ensure(G.ncell + 1); c->cend.ensure(G.ncell + 1);
}

extern "C" int edb_version(void) { return 100; }

extern "C" const char* edb_last_error(const edb_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int edb_create(const edb_config* cfg, edb_context** out) {
  if (!cfg || !out) return EDB_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, EDB_ERR_NO_DEVICE, "no CUDA device: libexadem_b200 has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, EDB_ERR_NO_DEVICE, "invalid device ordinal");
  edb_context* c = new edb_context();
  c->cfg = *cfg;
  if (c->cfg.nranks < 1) c->cfg.nranks = 1;
  c->env_no_fan = getenv("EDB_NO_FAN") != nullptr;
  c->env_bonds_host = getenv("EDB_BONDS_HOST") != nullptr;
  c->env_bonds_unfused = getenv("EDB_BONDS_UNFUSED") != nullptr;
  c->env_nbh_two_pass = getenv("EDB_NBH_TWO_PASS") != nullptr;
  c->env_nbh_no_cull = getenv("EDB_NBH_NO_CULL") != nullptr;
  c->env_nbh_warp = getenv("EDB_NBH_WARP") != nullptr;
  c->trace_async = getenv("EDB_TRACE_ASYNC") != nullptr;
  c->env_flag_on_stream = getenv("EDB_FLAG_ON_STREAM") != nullptr;
  c->env_land_copy = getenv("EDB_LAND_COPY") != nullptr;
  c->env_gather_mail = getenv("EDB_GATHER_MAIL") != nullptr;
  c->env_nbh_per_thread = getenv("EDB_NBH_POLY_PER_THREAD") != nullptr;
  c->env_nbh_no_replay = getenv("EDB_NBH_NO_REPLAY") != nullptr;
  c->env_drivers_unfused = getenv("EDB_DRIVERS_UNFUSED") != nullptr;
  try {
    CK(cudaSetDevice(cfg->device));
    CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CK(cudaMallocHost(&c->h_flag, sizeof(unsigned long long) * 4));
    CK(cudaMalloc(&c->d_flag, sizeof(unsigned long long) * 4));
    CK(cudaMalloc(&c->d_need_count, 16));
    CK(cudaMalloc(&c->d_sflag, 16));
    CK(cudaHostAlloc(&c->h_sflag, 16, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&c->h_sflag_dev, c->h_sflag, 0));
    CK(cudaMallocHost(&c->h_nbh, 32));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&c->ev_sflag[i], cudaEventDisableTiming));
    { int dev = 0, v = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) c->n_sm = v;
      if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0) c->smem_optin = (size_t)v; }
    CK(cudaMallocHost(&c->h_tot, sizeof(uint32_t) * (NT + 8)));
    for (int d = 0; d < 3; d++)
      if (c->cfg.periodic[d] && c->cfg.pgrid[d] <= 1) {
        const int nc = std::max(1, (int)std::floor((c->cfg.hi[d] - c->cfg.lo[d]) / c->cfg.cell_size + 1e-9));
        if (nc < 3) throw StateError{EDB_ERR_INVALID, "a periodic axis needs at least 3 cells"};
      }
    setup_grid(c);
    Params z{};
    c->prm.pp = z; c->prm.pd = z; c->prm.tpp = nullptr; c->prm.tpd = nullptr; c->prm.n_groups = 1; c->prm.n_drivers = 0;
    c->drv.ensure(1);
  } catch (const CudaError& e) {
    char b[512];
    snprintf(b, sizeof b, "CUDA error %s at %s:%d in %s", cudaGetErrorString(e.e), e.file, e.line, e.what);
    g_create_error = b;
    delete c;
    return EDB_ERR_CUDA;
  } catch (const StateError& e) {
    g_create_error = e.msg;
    delete c;
    return e.code;
  }
  *out = c;
  return EDB_OK;
}

// Cost-balanced decomposition: the bricks of the process grid are cut at these positions instead of uniformly (the reference's
// load_balance_rcb fed by dem_cost_model; here tensor-product cuts, see comm.cuh).  Re-plans the ghost layer of this rank.
extern "C" int edb_set_brick_cuts(edb_context* c, const double* cx, const double* cy, const double* cz) {
  API_BEGIN(c)
  if (c->nreal || c->comm.comm) throw StateError{EDB_ERR_STATE, "edb_set_brick_cuts: call before edb_comm_init and before the first upload"};
  const double* in[3] = {cx, cy, cz};
  int rc[3];
  Comm::coords(c->comm.pg, c->cfg.rank, rc);
  for (int d = 0; d < 3; d++) {
    c->comm.cuts[d].clear();
    if (!in[d]) continue;
    const int pg = c->comm.pg[d];
    if (pg + 1 > MAX_CUTS) throw StateError{EDB_ERR_INVALID, "edb_set_brick_cuts: more than 16 bricks along an axis"};
    const double L = c->cfg.ghi[d] - c->cfg.glo[d], tol = 1e-9 * L;
    if (std::fabs(in[d][0] - c->cfg.glo[d]) > tol || std::fabs(in[d][pg] - c->cfg.ghi[d]) > tol)
      throw StateError{EDB_ERR_INVALID, "edb_set_brick_cuts: the first / last cut must be the global bounds"};
    for (int k = 0; k < pg; k++)
      if (!(in[d][k + 1] - in[d][k] >= c->cfg.cell_size * (1.0 - 1e-9)))
        throw StateError{EDB_ERR_INVALID, "edb_set_brick_cuts: every brick must be at least one cell_size wide"};
    if (std::fabs(in[d][rc[d]] - c->cfg.lo[d]) > tol || std::fabs(in[d][rc[d] + 1] - c->cfg.hi[d]) > tol)
      throw StateError{EDB_ERR_INVALID, "edb_set_brick_cuts: this rank's lo / hi are not the cuts of its brick"};
    c->comm.cuts[d].assign(in[d], in[d] + pg + 1);
    c->comm.cuts[d][rc[d]] = c->cfg.lo[d];       // the brick's own bounds bit for bit (its cell grid is computed from them)
    c->comm.cuts[d][rc[d] + 1] = c->cfg.hi[d];
  }
  c->comm.release_peers();
  setup_grid(c);
  API_END(c)
}

static void trace_dump(edb_context* c);
extern "C" int edb_destroy(edb_context* c) {
  if (!c) return EDB_ERR_INVALID;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->st);
  if (c->st2) { cudaStreamSynchronize(c->st2); cudaStreamDestroy(c->st2); }
  trace_dump(c);
  if (c->st_flag) { cudaStreamSynchronize(c->st_flag); cudaStreamDestroy(c->st_flag); cudaEventDestroy(c->ev_flag_src); }
  if (c->st_up) {
    cudaStreamSynchronize(c->st_up); cudaStreamSynchronize(c->st_dn);
    cudaStreamDestroy(c->st_up); cudaStreamDestroy(c->st_dn);
    cudaEventDestroy(c->ev_up_done); cudaEventDestroy(c->ev_up_consumed);
    for (int k = 0; k < 2; k++) { cudaEventDestroy(c->ev_dn_ready[k]); cudaEventDestroy(c->ev_dn_done[k]); }
    for (int k = 0; k < 22; k++) cudaEventDestroy(c->ev_field[k]);
    c->stage_up.release(); c->stage_dn[0].release(); c->stage_dn[1].release();
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  c->comm.destroy();
  c->pa.release(); c->pb.release();
  for (int t = 0; t < NT; t++) { c->wave[t].release(); c->prev[t].release(); }
  DevBuf<uint32_t>* u32s[] = {&c->cidx, &c->cidx2, &c->ccount, &c->cstart, &c->cfill, &c->tmp, &c->new2old, &c->old2new, &c->cbeg, &c->cend,
                              &c->scan_bsum, &c->scan_total, &c->g_cell, &c->g_src_cell, &c->g_count, &c->g_off, &c->g_src, &c->g_ent,
                              &c->inc_cnt, &c->inc_off, &c->inc, &c->stage_u32, &c->fan_cnt, &c->fan_off, &c->fan_bits, &c->need, &c->nbh_pool, &c->nbh_first, &c->nbh_ctr};
  for (auto* b : u32s) b->release();
  c->pr_backup.release(); c->q_backup.release(); c->vert.release(); c->g_shift.release(); c->tpp.release(); c->tpd.release();
  c->drv.release(); c->d_sh.release(); c->d_sv.release(); c->d_se.release(); c->d_sfoff.release(); c->d_sfn.release(); c->d_sfidx.release();
  c->stage.release(); c->stage_u64.release(); c->l2buf.release();
  c->d_md.release(); c->d_mverts.release(); c->d_medges.release(); c->d_mfoff.release(); c->d_mfn.release(); c->d_mfidx.release();
  c->d_mcell_data.release(); c->d_mcell_off.release();
  c->mo.release(); c->mo_backup.release(); c->drv_backup.release(); c->d_mlocal.release(); c->d_mprim_start.release();
  c->d_mcell_cnt.release(); c->drv_sums.release();
  c->b_ij.release(); c->b_sub.release(); c->b_itf.release(); c->b_par.release(); c->b_fr.release(); c->b_tds.release(); c->b_fd.release();
  c->b_cp.release(); c->b_mode.release(); c->b_unbroken.release(); c->b_itf_broken.release(); c->b_plist.release(); c->b_inc_off.release();
  c->b_inc.release(); c->d_ibp.release(); c->b_ign_off.release(); c->b_ign_ids.release(); c->d_conv.release();
  c->mig_dest.release(); c->mig_midx.release(); c->mig_cnt.release(); c->mig_off.release(); c->imm_cnt.release(); c->imm_off.release();
  c->mig_sp.release(); c->mig_rp.release(); c->mig_sh.release(); c->mig_rh.release(); c->imm_tab.release();
  c->tw_cnt.release(); c->tw_flag.release(); c->tw_pos.release(); c->tw_own.release(); c->tw_par.release(); c->tw_tpos.release();
  c->tw_s.release(); c->tw_r.release(); c->fan_ent.release();
  for (auto& e : c->ev_free) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (auto& e : c->ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (auto& e : c->user_ev) if (e) cudaEventDestroy(e);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  if (c->d_flag) cudaFree(c->d_flag);
  if (c->d_need_count) cudaFree(c->d_need_count);
  if (c->d_sflag) cudaFree(c->d_sflag);
  if (c->h_sflag) cudaFreeHost(c->h_sflag);
  c->rmax_dev.release();
  c->pp_tmp.release();
  if (c->h_nbh) cudaFreeHost(c->h_nbh);
  for (int i = 0; i < 2; i++) if (c->ev_sflag[i]) cudaEventDestroy(c->ev_sflag[i]);
  if (c->h_tot) cudaFreeHost(c->h_tot);
  cudaStreamDestroy(c->st);
  delete c;
  return EDB_OK;
}

// Host setters that touch device-resident state (driver / motion / parameter / mesh tables) wait for the kernels already enqueued:
// the library streams are non-blocking, so a blocking cudaMemcpy on the legacy stream is NOT ordered behind them.
static void quiesce(edb_context* c) {
  if (c->st) CK(cudaStreamSynchronize(c->st));
  if (c->st2) CK(cudaStreamSynchronize(c->st2));
  if (c->st_up) CK(cudaStreamSynchronize(c->st_up));
  if (c->st_dn) CK(cudaStreamSynchronize(c->st_dn));
}

extern "C" int edb_sync(edb_context* c) {
  API_BEGIN(c)
  CK(cudaStreamSynchronize(c->st));
  if (c->st_up) CK(cudaStreamSynchronize(c->st_up));   // asynchronous host copies, if any were issued
  if (c->st_dn) CK(cudaStreamSynchronize(c->st_dn));
  API_END(c)
}

// ---- parameters -------------------------------------------------------------------------------------------------
static Params to_params(const edb_contact_params& p) { return Params{p.kn, p.kt, p.kr, p.mu, p.damp_rate, p.fc, p.dncut, p.gamma}; }

extern "C" int edb_set_contact_params(edb_context* c, const edb_contact_params* a, const edb_contact_params* b) {
  API_BEGIN(c)
  if (!a) throw StateError{EDB_ERR_INVALID, "config is required"};
  c->prm.pp = to_params(*a);
  c->prm.pd = b ? to_params(*b) : Params{};
  c->prm.tpp = nullptr; c->prm.tpd = nullptr; c->prm.n_groups = 1;
  API_END(c)
}
extern "C" int edb_set_contact_params_multimat(edb_context* c, int32_t ng, const edb_contact_params* pp, int32_t nd, const edb_contact_params* pd) {
  API_BEGIN(c)
  if (ng < 1 || !pp) throw StateError{EDB_ERR_INVALID, "multimat table is empty"};
  quiesce(c);
  std::vector<Params> A, B;
  for (int i = 0; i < ng * ng; i++) A.push_back(to_params(pp[i]));
  for (int i = 0; i < ng * nd; i++) B.push_back(to_params(pd[i]));
  c->tpp.ensure(A.size()); c->tpd.ensure(B.size() + 1);
  CK(cudaMemcpy(c->tpp.p, A.data(), A.size() * sizeof(Params), cudaMemcpyHostToDevice));
  if (!B.empty()) CK(cudaMemcpy(c->tpd.p, B.data(), B.size() * sizeof(Params), cudaMemcpyHostToDevice));
  c->prm.tpp = c->tpp.p; c->prm.tpd = c->tpd.p; c->prm.n_groups = ng; c->prm.n_drivers = nd;
  API_END(c)
}
extern "C" int edb_set_law(edb_context* c, int32_t law) {
  API_BEGIN(c)
  if (law < 0 || law > 2) throw StateError{EDB_ERR_INVALID, "unknown law"};
  c->law = law;
  API_END(c)
}
// Static tables of the RShape meshes (topology, local vertices) + the world vertices at the pose given to edb_set_drivers.  From here on
// the world vertices are owned by the device (k_driver_vertices follows the drivers' motion).
static void upload_mesh_tables(edb_context* c) {
  const int nm = (int)c->h_md.size();
  const MeshView old = c->M;
  c->M = MeshView{};
  if (nm == 0) return;
  c->d_md.ensure(nm); c->d_mverts.ensure(c->h_mverts.size() + 1); c->d_mlocal.ensure(c->h_mlocal.size() + 1);
  c->d_medges.ensure(c->h_medges.size() + 1);
  c->d_mfoff.ensure(c->h_mfoff.size() + 1); c->d_mfn.ensure(c->h_mfn.size() + 1); c->d_mfidx.ensure(c->h_mfidx.size() + 1);
  CK(cudaMemcpy(c->d_md.p, c->h_md.data(), nm * sizeof(MeshDev), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_mverts.p, c->h_mverts.data(), c->h_mverts.size() * sizeof(double4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_mlocal.p, c->h_mlocal.data(), c->h_mlocal.size() * sizeof(double4), cudaMemcpyHostToDevice));
  if (!c->h_medges.empty()) CK(cudaMemcpy(c->d_medges.p, c->h_medges.data(), c->h_medges.size() * 4, cudaMemcpyHostToDevice));
  if (!c->h_mfn.empty()) {
    CK(cudaMemcpy(c->d_mfoff.p, c->h_mfoff.data(), c->h_mfoff.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_mfn.p, c->h_mfn.data(), c->h_mfn.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_mfidx.p, c->h_mfidx.data(), c->h_mfidx.size() * 4, cudaMemcpyHostToDevice));
  }
  std::vector<uint32_t> ps(nm + 1, 0);
  for (int m = 0; m < nm; m++) ps[m + 1] = ps[m] + (uint32_t)(c->h_md[m].nv + c->h_md[m].ne + c->h_md[m].nf);
  c->d_mprim_start.ensure(nm + 1);
  CK(cudaMemcpy(c->d_mprim_start.p, ps.data(), (nm + 1) * 4, cudaMemcpyHostToDevice));
  // a host-driven update of the same drivers between two rebuilds (edb_set_drivers called again): the contact operators keep working on
  // the cell lists of the last rebuild with the refreshed tables; anything else waits for the next rebuild (mesh_dirty)
  if (old.nmesh == nm && old.cell_off)
    c->M = MeshView{c->d_md.p, c->d_mverts.p, c->d_medges.p, c->d_mfoff.p, c->d_mfn.p, c->d_mfidx.p, old.cell_off, old.cell_data, nm, old.ncell};
}

// [grid_rshape_driver] per-cell primitive lists of the RShape meshes, built on the device from the current world vertices
// (kernels_driver.cuh: count -> scan -> fill -> per-list sort).  The reference builds them on the host with OpenMP locks
// (grid_stl_mesh.cpp:150-251) at every update_particles_full (config_move_particles.msp:92-94).
static void build_mesh_grid(edb_context* c) {
  const GridDev& G = c->G;
  const int nm = (int)c->h_md.size();
  c->M = MeshView{};
  c->mesh_dirty = false;
  if (nm == 0) return;
  uint32_t nprim = 0;
  for (int m = 0; m < nm; m++) nprim += (uint32_t)(c->h_md[m].nv + c->h_md[m].ne + c->h_md[m].nf);
  const size_t nlists = (size_t)nm * G.ncell * 3;
  if (nlists + 1 > 0xffffffffull) throw StateError{EDB_ERR_CAPACITY, "mesh cell table exceeds 2^32 lists"};
  MeshGridArgs A{c->d_md.p, c->d_mverts.p, c->d_medges.p, c->d_mfoff.p, c->d_mfn.p, c->d_mfidx.p, nm, nprim, c->d_mprim_start.p, G,
                 std::max(c->cfg.cell_size, std::max(G.cs[0], std::max(G.cs[1], G.cs[2])))};
  c->d_mcell_cnt.ensure(nlists + 1);
  c->d_mcell_off.ensure(nlists + 2);
  CK(cudaMemsetAsync(c->d_mcell_cnt.p, 0, (nlists + 1) * 4, c->st));
  LAUNCH(c, k_mesh_cells<0>, nblk(nprim), BLK, 0, A, c->d_mcell_cnt.p, nullptr, nullptr);
  scan_u32(c, c->d_mcell_cnt.p, c->d_mcell_off.p, (uint32_t)nlists);
  uint32_t total = 0;
  CK(cudaMemcpyAsync(&total, c->d_mcell_off.p + nlists, 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  c->d_mcell_data.ensure((size_t)total + 1);
  CK(cudaMemsetAsync(c->d_mcell_cnt.p, 0, (nlists + 1) * 4, c->st));
  LAUNCH(c, k_mesh_cells<1>, nblk(nprim), BLK, 0, A, c->d_mcell_cnt.p, c->d_mcell_off.p, c->d_mcell_data.p);
  LAUNCH(c, k_mesh_list_sort, nblk(nlists), BLK, 0, c->d_mcell_off.p, (uint32_t)nlists, c->d_mcell_data.p);
  c->M = MeshView{c->d_md.p, c->d_mverts.p, c->d_medges.p, c->d_mfoff.p, c->d_mfn.p, c->d_mfidx.p, c->d_mcell_off.p, c->d_mcell_data.p, nm,
                  (uint32_t)G.ncell};
}

extern "C" int edb_set_drivers(edb_context* c, int32_t n, const edb_driver* d) {
  API_BEGIN(c)
  quiesce(c);
  c->h_drv.clear();
  for (int i = 0; i < n; i++) {
    Driver D{};
    D.kind = d[i].kind; D.need_forces = d[i].need_forces; D.radius = d[i].radius;
    D.axis = mk3(d[i].axis[0], d[i].axis[1], d[i].axis[2]);
    D.center = mk3(d[i].center[0], d[i].center[1], d[i].center[2]);
    if (D.kind == EDB_DRIVER_CYLINDER) D.center = cmul(D.axis, D.center);  // cylinder.hpp:104 (quirk Q11)
    D.offset = d[i].offset;
    D.normal = mk3(d[i].normal[0], d[i].normal[1], d[i].normal[2]);
    D.center_proj = D.offset * D.normal;                                   // surface.hpp:149
    D.vel = mk3(d[i].vel[0], d[i].vel[1], d[i].vel[2]);
    D.vrot = mk3(d[i].vrot[0], d[i].vrot[1], d[i].vrot[2]);
    D.shape = d[i].shape;
    c->h_drv.push_back(D);
  }
  // RShape drivers: world vertices = position + orientation * v (rshape.hpp:381-384), pooled per driver
  c->h_drv_shape.assign(n, -1);
  c->h_md.clear(); c->h_mverts.clear(); c->h_mlocal.clear(); c->h_medges.clear(); c->h_mfoff.clear(); c->h_mfn.clear(); c->h_mfidx.clear();
  for (int i = 0; i < n; i++) {
    Driver& D = c->h_drv[i];
    if (D.kind != EDB_DRIVER_RSHAPE) continue;
    const int sidx = d[i].shape;
    if (sidx < 0 || sidx >= (int)c->h_mesh_v.size()) throw StateError{EDB_ERR_INVALID, "RSHAPE driver: shape index outside the edb_set_driver_shapes table"};
    MeshDev md{};
    const auto& lv = c->h_mesh_v[sidx];
    md.nv = (int)lv.size() / 3; md.ne = (int)c->h_mesh_e[sidx].size() / 2; md.nf = (int)c->h_mesh_fn[sidx].size();
    md.drv = i;
    md.v0 = (int)c->h_mverts.size(); md.e0 = (int)c->h_medges.size() / 2; md.f0 = (int)c->h_mfn.size(); md.fi0 = (int)c->h_mfidx.size();
    md.radius = c->h_mesh_radius[sidx];
    const q4 q{d[i].quat[0], d[i].quat[1], d[i].quat[2], d[i].quat[3]};
    for (int k = 0; k < md.nv; k++) {
      const v3 w = D.center + qrot(q, 1.0 * mk3(lv[3 * k], lv[3 * k + 1], lv[3 * k + 2]));
      c->h_mverts.push_back(make_double4(w.x, w.y, w.z, 0.0));
      c->h_mlocal.push_back(make_double4(lv[3 * k], lv[3 * k + 1], lv[3 * k + 2], 0.0));
    }
    c->h_medges.insert(c->h_medges.end(), c->h_mesh_e[sidx].begin(), c->h_mesh_e[sidx].end());
    c->h_mfoff.insert(c->h_mfoff.end(), c->h_mesh_foff[sidx].begin(), c->h_mesh_foff[sidx].end());
    c->h_mfn.insert(c->h_mfn.end(), c->h_mesh_fn[sidx].begin(), c->h_mesh_fn[sidx].end());
    c->h_mfidx.insert(c->h_mfidx.end(), c->h_mesh_fidx[sidx].begin(), c->h_mesh_fidx[sidx].end());
    c->h_drv_shape[i] = sidx;
    D.shape = (int)c->h_md.size();   // device: slot of this driver's mesh
    c->h_md.push_back(md);
  }
  c->mesh_dirty = true;
  c->drv.ensure(n + 1);
  if (n) CK(cudaMemcpy(c->drv.p, c->h_drv.data(), n * sizeof(Driver), cudaMemcpyHostToDevice));
  if (!c->prm.tpd) c->prm.n_drivers = n;
  // every driver starts STATIONARY (no edb_set_driver_motions yet); the orientation of an RShape lives with the motion state
  c->h_mo.assign(n, DriverMotion{});
  for (int i = 0; i < n; i++) {
    DriverMotion& m = c->h_mo[i];
    m.t_end = 1e300; m.shaker_dir = mk3(0, 0, 1); m.mass = 1.7976931348623157e308 / 4; m.surface = -1;
    m.quat = q4{d[i].quat[0], d[i].quat[1], d[i].quat[2], d[i].quat[3]};
  }
  c->mo.ensure(n + 1);
  if (n) CK(cudaMemcpy(c->mo.p, c->h_mo.data(), n * sizeof(DriverMotion), cudaMemcpyHostToDevice));
  c->drivers_moving = false; c->drivers_need_reaction = false; c->mesh_moving = false;
  c->have_drv_backup = false;
  upload_mesh_tables(c);
  API_END(c)
}
static bool any_driver_moves(const edb_context* c) {
  for (size_t i = 0; i < c->h_drv.size(); i++)
    if (driver_moves(c->h_drv[i], c->h_mo[i])) return true;
  return false;
}
extern "C" int edb_set_driver_motions(edb_context* c, int32_t n, const edb_driver_motion* mo) {
  API_BEGIN(c)
  if (n != (int)c->h_drv.size()) throw StateError{EDB_ERR_INVALID, "edb_set_driver_motions: n differs from the driver count of edb_set_drivers"};
  quiesce(c);
  // the orientation of an RShape lives with the motion state on the device and may have moved since edb_set_drivers: keep the device's
  std::vector<DriverMotion> dev_mo(n);
  if (n) CK(cudaMemcpy(dev_mo.data(), c->mo.p, n * sizeof(DriverMotion), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    const edb_driver_motion& s = mo[i];
    DriverMotion& m = c->h_mo[i];
    const q4 keep = dev_mo[i].quat;
    if (s.motion_type < 0 || s.motion_type > EDB_MOTION_UNKNOWN) throw StateError{EDB_ERR_INVALID, "unknown motion type"};
    if (s.motion_type == EDB_MOTION_TABULATED && (s.n_tab < 1 || s.n_tab > EDB_MAX_TAB || !s.tab_time || !s.tab_pos))
      throw StateError{EDB_ERR_INVALID, "TABULATED motion needs 1..EDB_MAX_TAB (time, position) entries"};
    m = DriverMotion{};
    m.quat = keep;
    m.motion = s.motion_type; m.n_tab = s.motion_type == EDB_MOTION_TABULATED ? s.n_tab : 0;
    m.expr_use_v = s.expr_use_v; m.expr_use_vrot = s.expr_use_vrot; m.expr_use_mom = s.expr_use_mom; m.drive_by_mom = s.drive_by_mom;
    auto V = [](const double* p) { return mk3(p[0], p[1], p[2]); };
    m.motion_vector = V(s.motion_vector); m.t_start = s.motion_start_threshold; m.t_end = s.motion_end_threshold;
    m.const_vel = s.const_vel; m.const_force = s.const_force; m.sigma = s.sigma; m.damprate = s.damprate;
    m.omega = s.omega; m.amplitude = s.amplitude; m.shaker_dir = V(s.shaker_dir);
    m.pend_anchor = V(s.pendulum_anchor_point); m.pend_init = V(s.pendulum_initial_position); m.pend_dir = V(s.pendulum_swing_dir);
    m.mass = s.mass; m.surface = s.surface; m.rv = s.rv; m.ra = s.ra;
    m.acc = V(s.acc); m.inertia = V(s.inertia); m.applied_mom = V(s.applied_mom); m.mom_axis = V(s.mom_axis); m.mom = V(s.mom);
    for (int k = 0; k < m.n_tab; k++) { m.tab_time[k] = s.tab_time[k]; m.tab_pos[k] = V(s.tab_pos + 3 * k); }
    c->h_drv[i].need_forces = m_need_forces(m.motion) ? 1 : 0;   // motion_type.hpp:133-141
  }
  // the device copy of the drivers may already have moved: only the need_forces flag is refreshed there
  if (n) {
    std::vector<Driver> cur(n);
    CK(cudaMemcpy(cur.data(), c->drv.p, n * sizeof(Driver), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) cur[i].need_forces = c->h_drv[i].need_forces;
    CK(cudaMemcpy(c->drv.p, cur.data(), n * sizeof(Driver), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->mo.p, c->h_mo.data(), n * sizeof(DriverMotion), cudaMemcpyHostToDevice));
  }
  c->drivers_moving = any_driver_moves(c);
  c->drivers_need_reaction = false; c->mesh_moving = false;
  for (int i = 0; i < n; i++) {
    const bool mesh = c->h_drv[i].kind == EDB_DRIVER_RSHAPE;
    if (m_need_forces(c->h_mo[i].motion) || (mesh && rshape_need_moment(c->h_mo[i]))) c->drivers_need_reaction = true;
    if (mesh && driver_moves(c->h_drv[i], c->h_mo[i])) c->mesh_moving = true;
  }
  API_END(c)
}
extern "C" int edb_set_driver_expr(edb_context* c, int32_t d, const double v[3], const double vrot[3], const double mom[3]) {
  API_BEGIN(c)
  if (d < 0 || d >= (int)c->h_drv.size()) throw StateError{EDB_ERR_INVALID, "driver index out of range"};
  quiesce(c);
  DriverMotion m;
  CK(cudaMemcpy(&m, c->mo.p + d, sizeof(DriverMotion), cudaMemcpyDeviceToHost));
  if (v) m.expr_v = mk3(v[0], v[1], v[2]);
  if (vrot) m.expr_vrot = mk3(vrot[0], vrot[1], vrot[2]);
  if (mom) m.expr_mom = mk3(mom[0], mom[1], mom[2]);
  CK(cudaMemcpy(c->mo.p + d, &m, sizeof(DriverMotion), cudaMemcpyHostToDevice));
  API_END(c)
}
extern "C" int edb_get_drivers(edb_context* c, int32_t n, edb_driver* out, edb_driver_motion* mo) {
  API_BEGIN(c)
  if (n != (int)c->h_drv.size()) throw StateError{EDB_ERR_INVALID, "edb_get_drivers: n differs from the driver count"};
  if (n == 0) return EDB_OK;
  CK(cudaStreamSynchronize(c->st));
  std::vector<Driver> D(n);
  std::vector<DriverMotion> M(n);
  CK(cudaMemcpy(D.data(), c->drv.p, n * sizeof(Driver), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(M.data(), c->mo.p, n * sizeof(DriverMotion), cudaMemcpyDeviceToHost));
  auto S = [](double* o, v3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
  for (int i = 0; i < n; i++) {
    if (out) {
      edb_driver& o = out[i];
      o = edb_driver{};
      o.kind = D[i].kind; o.need_forces = D[i].need_forces; o.radius = D[i].radius; S(o.axis, D[i].axis); S(o.center, D[i].center);
      o.offset = D[i].offset; S(o.normal, D[i].normal); S(o.vel, D[i].vel); S(o.vrot, D[i].vrot);
      o.shape = D[i].kind == EDB_DRIVER_RSHAPE ? c->h_drv_shape[i] : D[i].shape;
      o.quat[0] = M[i].quat.w; o.quat[1] = M[i].quat.x; o.quat[2] = M[i].quat.y; o.quat[3] = M[i].quat.z;
    }
    if (mo) {
      edb_driver_motion& o = mo[i];
      o = edb_driver_motion{};
      const DriverMotion& m = M[i];
      o.motion_type = m.motion; o.n_tab = m.n_tab; S(o.motion_vector, m.motion_vector);
      o.motion_start_threshold = m.t_start; o.motion_end_threshold = m.t_end; o.const_vel = m.const_vel; o.const_force = m.const_force;
      o.sigma = m.sigma; o.damprate = m.damprate; o.omega = m.omega; o.amplitude = m.amplitude; S(o.shaker_dir, m.shaker_dir);
      S(o.pendulum_anchor_point, m.pend_anchor); S(o.pendulum_initial_position, m.pend_init); S(o.pendulum_swing_dir, m.pend_dir);
      o.expr_use_v = m.expr_use_v; o.expr_use_vrot = m.expr_use_vrot; o.expr_use_mom = m.expr_use_mom; o.drive_by_mom = m.drive_by_mom;
      o.mass = m.mass; o.surface = m.surface; S(o.acc, m.acc); o.rv = m.rv; o.ra = m.ra;
      S(o.inertia, m.inertia); S(o.applied_mom, m.applied_mom); S(o.mom_axis, m.mom_axis); S(o.mom, m.mom); S(o.forces, m.forces);
    }
  }
  API_END(c)
}
extern "C" int edb_set_driver_shapes(edb_context* c, int32_t n, const edb_shape* sh) {
  API_BEGIN(c)
  quiesce(c);
  c->h_mesh_v.clear(); c->h_mesh_e.clear(); c->h_mesh_foff.clear(); c->h_mesh_fn.clear(); c->h_mesh_fidx.clear(); c->h_mesh_radius.clear();
  for (int i = 0; i < n; i++) {
    c->h_mesh_v.emplace_back(sh[i].vertices, sh[i].vertices + 3 * (size_t)sh[i].nv);
    c->h_mesh_e.emplace_back(sh[i].edges, sh[i].edges + 2 * (size_t)sh[i].ne);
    std::vector<int> foff, fn, fidx;
    int pos = 0;
    for (int f = 0; f < sh[i].nf; f++) {
      const int nfv = sh[i].faces[pos++];
      if (nfv < 3 || pos + nfv > sh[i].faces_len) throw StateError{EDB_ERR_INVALID, "malformed face stream"};
      foff.push_back((int)fidx.size());
      fn.push_back(nfv);
      for (int k = 0; k < nfv; k++) fidx.push_back(sh[i].faces[pos++]);
    }
    c->h_mesh_foff.push_back(foff); c->h_mesh_fn.push_back(fn); c->h_mesh_fidx.push_back(fidx);
    c->h_mesh_radius.push_back(sh[i].radius);
  }
  API_END(c)
}
extern "C" int edb_set_shapes(edb_context* c, int32_t n, const edb_shape* sh) {
  API_BEGIN(c)
  quiesce(c);
  std::vector<v3> sv;
  std::vector<int> se, sfoff, sfn, sfidx;
  c->h_sh.clear();
  c->h_farea.clear();
  int nvmax = 0;
  for (int i = 0; i < n; i++) {
    ShapeDev S{};
    S.nv = sh[i].nv; S.ne = sh[i].ne; S.nf = sh[i].nf;
    S.v0 = (int)sv.size(); S.e0 = (int)se.size() / 2; S.f0 = (int)sfn.size(); S.fi0 = (int)sfidx.size();
    for (int k = 0; k < S.nv; k++) sv.push_back(mk3(sh[i].vertices[3 * k], sh[i].vertices[3 * k + 1], sh[i].vertices[3 * k + 2]));
    for (int k = 0; k < 2 * S.ne; k++) se.push_back(sh[i].edges[k]);
    int pos = 0;
    for (int f = 0; f < S.nf; f++) {
      const int nfv = sh[i].faces[pos++];
      if (nfv < 3 || pos + nfv > sh[i].faces_len) throw StateError{EDB_ERR_INVALID, "malformed face stream"};
      sfoff.push_back((int)sfidx.size() - S.fi0);
      sfn.push_back(nfv);
      for (int k = 0; k < nfv; k++) sfidx.push_back(sh[i].faces[pos++]);
    }
    // face areas at homothety 1 (shape.hpp:673-697: fan triangulation from the first vertex), used by stick_polyhedra
    for (int f = 0; f < S.nf; f++) {
      const int* idx = &sfidx[S.fi0 + sfoff[S.f0 + f]];
      const v3 v0 = sv[S.v0 + idx[0]];
      double surface = 0;
      for (int j = 1; j < sfn[S.f0 + f] - 1; j++) surface += 0.5 * norm(cross(sv[S.v0 + idx[j]] - v0, sv[S.v0 + idx[j + 1]] - v0));
      c->h_farea.push_back(surface);
    }
    const double* o = sh[i].obb;
    S.obb = Obb{mk3(o[0], o[1], o[2]), mk3(o[3], o[4], o[5]), mk3(o[6], o[7], o[8]), mk3(o[9], o[10], o[11]), mk3(o[12], o[13], o[14])};
    S.radius = sh[i].radius;
    S.vmax = 0.0;
    for (int k = 0; k < S.nv; k++) S.vmax = std::max(S.vmax, norm(sv[S.v0 + k]));
    nvmax = std::max(nvmax, S.nv);
    c->shape_nemax = i == 0 ? S.ne : std::max(c->shape_nemax, S.ne);
    c->h_sh.push_back(S);
  }
  c->d_sh.ensure(n + 1); c->d_sv.ensure(sv.size() + 1); c->d_se.ensure(se.size() + 1); c->d_sfoff.ensure(sfoff.size() + 1);
  c->d_sfn.ensure(sfn.size() + 1); c->d_sfidx.ensure(sfidx.size() + 1);
  CK(cudaMemcpy(c->d_sh.p, c->h_sh.data(), n * sizeof(ShapeDev), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_sv.p, sv.data(), sv.size() * sizeof(v3), cudaMemcpyHostToDevice));
  if (!se.empty()) CK(cudaMemcpy(c->d_se.p, se.data(), se.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (!sfn.empty()) {
    CK(cudaMemcpy(c->d_sfoff.p, sfoff.data(), sfoff.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_sfn.p, sfn.data(), sfn.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_sfidx.p, sfidx.data(), sfidx.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  c->S = ShapesView{c->d_sh.p, c->d_sv.p, c->d_se.p, c->d_sfoff.p, c->d_sfn.p, c->d_sfidx.p, nvmax};
  c->mean_face_nv = sfn.empty() ? 3.0 : (double)sfidx.size() / (double)sfn.size();
  API_END(c)
}

// ---- host <-> device particle transfer ------------------------------------------------------------------------------
// staging layout: field-major blocks of n doubles
__global__ void k_pack_upload(PView P, const double* __restrict__ s, const uint64_t* __restrict__ id, const uint32_t* __restrict__ type,
                              const uint32_t* __restrict__ group, uint32_t n, unsigned long long have) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  auto F = [&](int f, double dflt) { return (have >> f) & 1ull ? s[(size_t)f * n + p] : dflt; };
  // field order = edb_particles: rx ry rz vx vy vz fx fy fz mass homothety radius qw qx qy qz mom(3) vrot(3) arot(3) inertia(3)
  P.pr[p] = make_double4(F(0, 0), F(1, 0), F(2, 0), F(11, 0.5));
  P.vm[p] = make_double4(F(3, 0), F(4, 0), F(5, 0), F(9, 1.0));
  const unsigned long long tg = (unsigned long long)(type ? type[p] : 0u) | ((unsigned long long)(group ? group[p] : 0u) << 32);
  P.wg[p] = make_double4(F(19, 0), F(20, 0), F(21, 0), __longlong_as_double((long long)tg));
  P.q[p] = make_double4(F(12, 1.0), F(13, 0), F(14, 0), F(15, 0));
  P.fx[p] = F(6, 0); P.fy[p] = F(7, 0); P.fz[p] = F(8, 0);
  P.mx[p] = F(16, 0); P.my[p] = F(17, 0); P.mz[p] = F(18, 0);
  P.ax[p] = F(22, 0); P.ay[p] = F(23, 0); P.az[p] = F(24, 0);
  P.ix[p] = F(25, 1.0); P.iy[p] = F(26, 1.0); P.iz[p] = F(27, 1.0);
  P.h[p] = F(10, 1.0);
  P.id[p] = id ? id[p] : (uint64_t)p;
  P.orig[p] = p;
}
__global__ void k_unpack_download(PView P, double* __restrict__ s, uint64_t* __restrict__ id, uint32_t* __restrict__ type,
                                  uint32_t* __restrict__ group, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t o = P.orig[p];
  const double4 r = P.pr[p], v = P.vm[p], w = P.wg[p], q = P.q[p];
  auto W = [&](int f, double val) { s[(size_t)f * n + o] = val; };
  W(0, r.x); W(1, r.y); W(2, r.z); W(3, v.x); W(4, v.y); W(5, v.z); W(6, P.fx[p]); W(7, P.fy[p]); W(8, P.fz[p]);
  W(9, v.w); W(10, P.h[p]); W(11, r.w); W(12, q.x); W(13, q.y); W(14, q.z); W(15, q.w);
  W(16, P.mx[p]); W(17, P.my[p]); W(18, P.mz[p]); W(19, w.x); W(20, w.y); W(21, w.z);
  W(22, P.ax[p]); W(23, P.ay[p]); W(24, P.az[p]); W(25, P.ix[p]); W(26, P.iy[p]); W(27, P.iz[p]);
  id[o] = P.id[p];
  const unsigned long long tg = (unsigned long long)__double_as_longlong(w.w);
  type[o] = (uint32_t)(tg & 0xffffffffu);
  group[o] = (uint32_t)(tg >> 32);
}
// dynamic subset in upload order: 13 inputs (r, v, vrot, q)
__global__ void k_scatter_dynamic(PView P, const double* __restrict__ s, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const size_t o = P.orig[p];
  double4 r = P.pr[p], v = P.vm[p], w = P.wg[p];
  r.x = s[0 * (size_t)n + o]; r.y = s[1 * (size_t)n + o]; r.z = s[2 * (size_t)n + o];
  v.x = s[3 * (size_t)n + o]; v.y = s[4 * (size_t)n + o]; v.z = s[5 * (size_t)n + o];
  w.x = s[6 * (size_t)n + o]; w.y = s[7 * (size_t)n + o]; w.z = s[8 * (size_t)n + o];
  P.pr[p] = r; P.vm[p] = v; P.wg[p] = w;
  P.q[p] = make_double4(s[9 * (size_t)n + o], s[10 * (size_t)n + o], s[11 * (size_t)n + o], s[12 * (size_t)n + o]);
}
// 22 outputs: r v vrot q f mom arot
__global__ void k_gather_dynamic(PView P, double* __restrict__ s, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const size_t o = P.orig[p];
  const double4 r = P.pr[p], v = P.vm[p], w = P.wg[p], q = P.q[p];
  const double vals[22] = {r.x, r.y, r.z, v.x, v.y, v.z, w.x, w.y, w.z, q.x, q.y, q.z, q.w, P.fx[p], P.fy[p], P.fz[p],
                           P.mx[p], P.my[p], P.mz[p], P.ax[p], P.ay[p], P.az[p]};
#pragma unroll
  for (int f = 0; f < 22; f++) s[(size_t)f * n + o] = vals[f];
}

static double* const* field_ptrs(const edb_particles* P, double* out[28]) {
  double* f[28] = {P->rx, P->ry, P->rz, P->vx, P->vy, P->vz, P->fx, P->fy, P->fz, P->mass, P->homothety, P->radius, P->qw, P->qx,
                   P->qy, P->qz, P->momx, P->momy, P->momz, P->vrotx, P->vroty, P->vrotz, P->arotx, P->aroty, P->arotz,
                   P->inertiax, P->inertiay, P->inertiaz};
  for (int i = 0; i < 28; i++) out[i] = f[i];
  return out;
}

static void clear_interactions(edb_context* c) {
  for (int t = 0; t < NT; t++) { c->wave[t].n = 0; c->prev[t].n = 0; }
  c->have_prev = false;
  c->have_inc = false;
  c->have_backup = false;
  // tables of the last rebuild refer to slots that no longer exist
  c->have_fan = false; c->ghost_pushed = false; c->bonds_on_device = false; c->have_drv_backup = false;
}

extern "C" int edb_upload_particles(edb_context* c, const edb_particles* P) {
  API_BEGIN(c)
  if (!P || !P->rx || !P->ry || !P->rz) throw StateError{EDB_ERR_INVALID, "positions are required"};
  const uint64_t n = P->n;
  if (n >= (1ull << 31)) throw StateError{EDB_ERR_CAPACITY, "too many particles for one GPU context"};
  c->pa.ensure(n + 1, false, c->st);
  c->stage.ensure(28 * n + 1); c->stage_u64.ensure(n + 1); c->stage_u32.ensure(2 * n + 1);
  double* f[28];
  field_ptrs(P, f);
  unsigned long long have = 0;
  for (int i = 0; i < 28; i++)
    if (f[i]) {
      have |= 1ull << i;
      CK(cudaMemcpyAsync(c->stage.p + (size_t)i * n, f[i], n * 8, cudaMemcpyHostToDevice, c->st));
    }
  if (P->id) CK(cudaMemcpyAsync(c->stage_u64.p, P->id, n * 8, cudaMemcpyHostToDevice, c->st));
  if (P->type) CK(cudaMemcpyAsync(c->stage_u32.p, P->type, n * 4, cudaMemcpyHostToDevice, c->st));
  if (P->group) CK(cudaMemcpyAsync(c->stage_u32.p + n, P->group, n * 4, cudaMemcpyHostToDevice, c->st));
  LAUNCH(c, k_pack_upload, nblk(n), BLK, 0, view(c->pa), c->stage.p, P->id ? c->stage_u64.p : nullptr,
         P->type ? c->stage_u32.p : nullptr, P->group ? c->stage_u32.p + n : nullptr, (uint32_t)n, have);
  CK(cudaStreamSynchronize(c->st));
  c->rmax_upload = 0;
  if (P->radius)
    for (uint64_t i = 0; i < n; i++) c->rmax_upload = std::max(c->rmax_upload, P->radius[i]);
  c->nreal = (uint32_t)n;
  c->have_cells = false;
  c->nghost = 0;
  clear_interactions(c);
  API_END(c)
}

extern "C" int edb_download_particles(edb_context* c, edb_particles* P) {
  API_BEGIN(c)
  const uint64_t n = c->nreal;
  if (!P) throw StateError{EDB_ERR_INVALID, "null particle view"};
  P->n = n;
  c->stage.ensure(28 * n + 1); c->stage_u64.ensure(n + 1); c->stage_u32.ensure(2 * n + 1);
  LAUNCH(c, k_unpack_download, nblk(n), BLK, 0, view(c->pa), c->stage.p, c->stage_u64.p, c->stage_u32.p, c->stage_u32.p + n, (uint32_t)n);
  double* f[28];
  field_ptrs(P, f);
  for (int i = 0; i < 28; i++)
    if (f[i]) CK(cudaMemcpyAsync(f[i], c->stage.p + (size_t)i * n, n * 8, cudaMemcpyDeviceToHost, c->st));
  if (P->id) CK(cudaMemcpyAsync(P->id, c->stage_u64.p, n * 8, cudaMemcpyDeviceToHost, c->st));
  if (P->type) CK(cudaMemcpyAsync(P->type, c->stage_u32.p, n * 4, cudaMemcpyDeviceToHost, c->st));
  if (P->group) CK(cudaMemcpyAsync(P->group, c->stage_u32.p + n, n * 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  API_END(c)
}

extern "C" int edb_upload_dynamic(edb_context* c, const edb_particles* P) {
  API_BEGIN(c)
  const uint64_t n = c->nreal;
  if (!P || P->n != n) throw StateError{EDB_ERR_INVALID, "particle count mismatch"};
  c->stage.ensure(28 * n + 1);
  double* f[13] = {P->rx, P->ry, P->rz, P->vx, P->vy, P->vz, P->vrotx, P->vroty, P->vrotz, P->qw, P->qx, P->qy, P->qz};
  for (int i = 0; i < 13; i++) {
    if (!f[i]) throw StateError{EDB_ERR_INVALID, "upload_dynamic needs r, v, vrot and orient"};
    CK(cudaMemcpyAsync(c->stage.p + (size_t)i * n, f[i], n * 8, cudaMemcpyHostToDevice, c->st));
  }
  LAUNCH(c, k_scatter_dynamic, nblk(n), BLK, 0, view(c->pa), c->stage.p, (uint32_t)n);
  API_END(c)
}
extern "C" int edb_download_dynamic(edb_context* c, edb_particles* P) {
  API_BEGIN(c)
  const uint64_t n = c->nreal;
  if (!P || P->n != n) throw StateError{EDB_ERR_INVALID, "particle count mismatch"};
  c->stage.ensure(28 * n + 1);
  LAUNCH(c, k_gather_dynamic, nblk(n), BLK, 0, view(c->pa), c->stage.p, (uint32_t)n);
  double* f[22] = {P->rx, P->ry, P->rz, P->vx, P->vy, P->vz, P->vrotx, P->vroty, P->vrotz, P->qw, P->qx, P->qy, P->qz,
                   P->fx, P->fy, P->fz, P->momx, P->momy, P->momz, P->arotx, P->aroty, P->arotz};
  for (int i = 0; i < 22; i++)
    if (f[i]) CK(cudaMemcpyAsync(f[i], c->stage.p + (size_t)i * n, n * 8, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  API_END(c)
}

// ---- asynchronous forms (pinned host buffers) ------------------------------------------------------------------------------------
static void trace_mark(edb_context* c, const char* what, cudaStream_t s) {
  if (!c->trace_async || c->trace_ev.size() > 4000) return;
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  CK(cudaEventRecord(e, s));
  c->trace_ev.push_back({what, e});
}
static void trace_dump(edb_context* c) {
  if (c->trace_ev.empty()) return;
  cudaDeviceSynchronize();
  const size_t n = c->trace_ev.size(), from = n > 160 ? n - 160 : 0;
  for (size_t k = from; k < n; k++) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->trace_ev[from].second, c->trace_ev[k].second);
    fprintf(stderr, "[edb trace] %10.4f ms  %s\n", ms, c->trace_ev[k].first);
  }
  for (auto& e : c->trace_ev) cudaEventDestroy(e.second);
  c->trace_ev.clear();
}
static void async_streams(edb_context* c) {
  if (c->st_up) return;
  CK(cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->st_dn, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_up_done, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_up_consumed, cudaEventDisableTiming));
  for (int k = 0; k < 2; k++) {
    CK(cudaEventCreateWithFlags(&c->ev_dn_ready[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_dn_done[k], cudaEventDisableTiming));
  }
  for (int k = 0; k < 22; k++) CK(cudaEventCreateWithFlags(&c->ev_field[k], cudaEventDisableTiming));
}
extern "C" int edb_upload_dynamic_async(edb_context* c, const edb_particles* P) {
  API_BEGIN(c)
  const uint64_t n = c->nreal;
  if (!P || P->n != n) throw StateError{EDB_ERR_INVALID, "particle count mismatch"};
  async_streams(c);
  double* f[13] = {P->rx, P->ry, P->rz, P->vx, P->vy, P->vz, P->vrotx, P->vroty, P->vrotz, P->qw, P->qx, P->qy, P->qz};
  for (int i = 0; i < 13; i++)
    if (!f[i]) throw StateError{EDB_ERR_INVALID, "upload_dynamic needs r, v, vrot and orient"};
  if (c->stage_up.cap < 13 * n + 1) { quiesce(c); c->stage_up.ensure(13 * n + 1); }
  // the staging area is free once the scatter kernel of the previous upload has read it
  if (c->up_consumed_valid) CK(cudaStreamWaitEvent(c->st_up, c->ev_up_consumed, 0));
  if (c->field_valid[0]) CK(cudaStreamWaitEvent(c->st_up, c->ev_field[0], 0));
  trace_mark(c, "upload begin (field 0 of the previous download landed)", c->st_up);
  for (int i = 0; i < 13; i++) {
    // the same host array may still be the target of the previous asynchronous download: field by field, so that on a full-duplex
    // link this upload runs one field behind that download instead of after it
    if (c->field_valid[i]) CK(cudaStreamWaitEvent(c->st_up, c->ev_field[i], 0));
    CK(cudaMemcpyAsync(c->stage_up.p + (size_t)i * n, f[i], n * 8, cudaMemcpyHostToDevice, c->st_up));
  }
  CK(cudaEventRecord(c->ev_up_done, c->st_up));
  trace_mark(c, "upload end", c->st_up);
  CK(cudaStreamWaitEvent(c->st, c->ev_up_done, 0));
  LAUNCH(c, k_scatter_dynamic, nblk(n), BLK, 0, view(c->pa), c->stage_up.p, (uint32_t)n);
  trace_mark(c, "scatter end", c->st);
  CK(cudaEventRecord(c->ev_up_consumed, c->st));
  c->up_consumed_valid = true;
  API_END(c)
}
extern "C" int edb_download_dynamic_async(edb_context* c, edb_particles* P, uint32_t field_mask) {
  API_BEGIN(c)
  const uint64_t n = c->nreal;
  if (!P || P->n != n) throw StateError{EDB_ERR_INVALID, "particle count mismatch"};
  async_streams(c);
  const int par = (c->dn_parity ^= 1);
  if (c->stage_dn[par].cap < 22 * n + 1) { quiesce(c); c->stage_dn[par].ensure(22 * n + 1); }
  // two staging areas: the snapshot of step n+1 is taken while the copy of step n is still draining
  if (c->dn_done_valid[par]) CK(cudaStreamWaitEvent(c->st, c->ev_dn_done[par], 0));
  trace_mark(c, "step end (download requested)", c->st);
  LAUNCH(c, k_gather_dynamic, nblk(n), BLK, 0, view(c->pa), c->stage_dn[par].p, (uint32_t)n);
  CK(cudaEventRecord(c->ev_dn_ready[par], c->st));
  CK(cudaStreamWaitEvent(c->st_dn, c->ev_dn_ready[par], 0));
  trace_mark(c, "download begin", c->st_dn);
  double* f[22] = {P->rx, P->ry, P->rz, P->vx, P->vy, P->vz, P->vrotx, P->vroty, P->vrotz, P->qw, P->qx, P->qy, P->qz,
                   P->fx, P->fy, P->fz, P->momx, P->momy, P->momz, P->arotx, P->aroty, P->arotz};
  // groups: bit 0 r, 1 v, 2 vrot, 3 orient, 4 f (acceleration), 5 mom, 6 arot
  static const int group_of_field[22] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 6};
  for (int i = 0; i < 22; i++) {
    if (!f[i] || !((field_mask >> group_of_field[i]) & 1u)) continue;
    CK(cudaMemcpyAsync(f[i], c->stage_dn[par].p + (size_t)i * n, n * 8, cudaMemcpyDeviceToHost, c->st_dn));
    CK(cudaEventRecord(c->ev_field[i], c->st_dn));
    c->field_valid[i] = true;
  }
  CK(cudaEventRecord(c->ev_dn_done[par], c->st_dn));
  trace_mark(c, "download end", c->st_dn);
  c->dn_done_valid[par] = true;
  API_END(c)
}

extern "C" int edb_num_particles(edb_context* c, uint64_t* nr, uint64_t* ng) {
  if (!c) return EDB_ERR_INVALID;
  if (nr) *nr = c->nreal;
  if (ng) *ng = c->nghost;
  return EDB_OK;
}

// ---- rebuild -----------------------------------------------------------------------------------------------------
// trig (device, nullable): the local rebuild-trigger count of the step, summed over the ranks in place on the way
// Copies the remote ghosts of the last landed epoch from the landing zone into the ghost slots (they were being read in place).
static void ghosts_home(edb_context* c) {
  if (!c->ghost_in_landing) return;
  ScopedTimer tm(c, B_GHOST);
  c->comm.land_copy(view(c->pa), c->nreal, c->ghost_in_landing, c->st, &c->ctr.kernel_launches);
  c->ghost_in_landing = 0;
}
// in_place: the caller's contact kernels can read remote ghosts from the landing zone (the wave-0 sphere kernel inside
// edb_run_verlet) — the landing launch then only exchanges the mail (epoch + rebuild-trigger count) and copies nothing.
static void ghost_refresh(edb_context* c, bool all, unsigned long long* trig = nullptr, bool in_place = false) {
  if (c->nghost == 0 && c->cfg.nranks == 1) return;
  if (!all && c->ghost_pushed) {   // the previous k_gather<2> stored the images itself; remote ones only have to land
    c->ghost_pushed = false;
    if (c->cfg.nranks > 1) {
      ScopedTimer tm(c, B_GHOST);
      const bool stay = in_place && !c->poly() && !c->env_land_copy;
      // the landing launch also stores the summed trigger count into the host's mapped copy (saves the publishing launch)
      unsigned long long* hf = nullptr;
      if (in_place && trig && !c->st_up && trig >= c->d_sflag && trig < c->d_sflag + 2) hf = c->h_sflag_dev + (trig - c->d_sflag);
      const uint32_t e = c->comm.land_pending(view(c->pa), c->nreal, trig, c->st, &c->ctr.kernel_launches, !stay, hf);
      c->flag_published = hf != nullptr;
      c->ghost_in_landing = stay ? e : 0;
    }
    return;
  }
  ghosts_home(c);
  c->ghost_pushed = false;
  c->comm.fan_epoch_pending = 0;
  ScopedTimer tm(c, B_GHOST);
  const uint32_t nloc = c->comm.n_local_ghost;
  if (nloc) {
    if (all) LAUNCH(c, k_ghost_refresh<true>, nblk(nloc), BLK, 0, view(c->pa), c->g_src.p, c->g_ent.p, c->g_shift.p, c->nreal, nloc, c->cidx.p, c->g_cell.p);
    else LAUNCH(c, k_ghost_refresh<false>, nblk(nloc), BLK, 0, view(c->pa), c->g_src.p, c->g_ent.p, c->g_shift.p, c->nreal, nloc, c->cidx.p, c->g_cell.p);
  }
  if (c->cfg.nranks > 1) {
    if (!all && c->comm.p2p.on) {
      c->comm.exchange_ghosts_p2p(view(c->pa), c->nreal, trig, c->st, &c->ctr.kernel_launches);
    } else {
      c->comm.exchange_ghosts(view(c->pa), c->nreal, all, c->st, &c->ctr.kernel_launches);
      if (trig) c->comm.allreduce_sum_u64_dev(trig, c->st);
    }
  }
}

// ghosts_only: the vertices of the real particles have just been written by k_gather<2> (fused compute_vertices)
static void compute_vertices(edb_context* c, bool ghosts_only = false) {
  if (!c->poly()) return;
  const uint32_t ntot = c->nreal + c->nghost;
  if (c->S.nvmax <= 0) throw StateError{EDB_ERR_STATE, "polyhedron mode needs edb_set_shapes"};
  c->vert.ensure((size_t)ntot * c->S.nvmax + 1, true, c->st);
  const uint32_t first = ghosts_only ? c->nreal : 0;
  c->ptype.ensure((size_t)ntot + 1, true, c->st);
  LAUNCH(c, k_compute_vertices, nblk(ntot - first), BLK, 0, view(c->pa), first, ntot, c->S, c->vert.p, c->ptype.p);
}

// Multi-GPU: particles that left the brick move to their new owner together with their active interaction records.
// Returns the number of occupied slots (stayers + leavers' dead slots + immigrants); *n_live = particles now owned.
static uint32_t migrate(edb_context* c, uint32_t* n_live) {
  const int R = c->cfg.nranks, me = c->cfg.rank;
  const uint32_t n = c->nreal;
  const bool poly = c->poly();
  BrickMap B{};
  for (int d = 0; d < 3; d++) {
    B.glo[d] = c->cfg.glo[d]; B.ghi[d] = c->cfg.ghi[d]; B.pg[d] = c->comm.pg[d]; B.periodic[d] = c->cfg.periodic[d];
    B.len[d] = (c->cfg.ghi[d] - c->cfg.glo[d]) / c->comm.pg[d];
    B.ncut[d] = (int)c->comm.cuts[d].size();
    for (int k = 0; k < B.ncut[d]; k++) B.cut[d][k] = c->comm.cuts[d][k];
  }
  B.me = me;
  c->mig_dest.ensure(n + 1); c->mig_midx.ensure(n + 1);
  c->mig_cnt.ensure(2 * R + 2); c->mig_off.ensure(2 * R + 2);
  uint32_t* d_cp = c->mig_cnt.p;      // particles per destination
  uint32_t* d_ch = c->mig_cnt.p + R;  // history records per destination
  CK(cudaMemsetAsync(c->mig_cnt.p, 0, 2 * R * 4, c->st));
  LAUNCH(c, k_mig_classify, nblk(n), BLK, 0, c->pa.pr.p, n, B, c->mig_dest.p, c->mig_midx.p, d_cp);
  WTabC W{};
  const bool had = c->have_prev;
  for (int t = 0; t < NT; t++) {
    W.w[t] = view(c->wave[t]);
    if (!had) W.w[t].n = 0;
  }
  LAUNCH(c, k_mig_pack<false>, nblk(n), BLK, 0, view(c->pa), n, c->mig_dest.p, c->mig_midx.p, W, poly, d_ch, nullptr, nullptr, nullptr, nullptr);
  std::vector<uint32_t> mine(2 * R), all((size_t)2 * R * R);
  CK(cudaMemcpyAsync(mine.data(), c->mig_cnt.p, 2 * R * 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  c->comm.allgather_u32(mine.data(), 2 * R, all.data(), c->st);
  // send side: offsets per destination; receive side: what every source has for me
  std::vector<uint32_t> sp_off(R + 1, 0), sh_off(R + 1, 0), rp_cnt(R), rh_cnt(R), rp_off(R + 1, 0), rh_off(R + 1, 0);
  for (int r = 0; r < R; r++) {
    sp_off[r + 1] = sp_off[r] + mine[r];
    sh_off[r + 1] = sh_off[r] + mine[R + r];
    rp_cnt[r] = all[(size_t)r * 2 * R + me];
    rh_cnt[r] = all[(size_t)r * 2 * R + R + me];
    rp_off[r + 1] = rp_off[r] + rp_cnt[r];
    rh_off[r + 1] = rh_off[r] + rh_cnt[r];
  }
  const uint32_t n_out = sp_off[R], h_out = sh_off[R], n_in = rp_off[R], h_in = rh_off[R];
  c->n_imm = n_in;
  c->migrated_out += n_out; c->migrated_in += n_in;
  *n_live = n - n_out + n_in;
  // any rank migrating => all ranks take part in the exchange (counts are globally known, so everyone agrees)
  uint64_t any = 0;
  for (size_t i = 0; i < all.size(); i++) any += all[i];
  if (any == 0) return n;
  c->mig_sp.ensure(n_out + 1); c->mig_sh.ensure(h_out + 1); c->mig_rp.ensure(n_in + 1); c->mig_rh.ensure(h_in + 1);
  CK(cudaMemcpyAsync(c->mig_off.p, sp_off.data(), R * 4, cudaMemcpyHostToDevice, c->st));
  CK(cudaMemcpyAsync(c->mig_off.p + R, sh_off.data(), R * 4, cudaMemcpyHostToDevice, c->st));
  CK(cudaMemsetAsync(d_ch, 0, R * 4, c->st));
  if (n_out)
    LAUNCH(c, k_mig_pack<true>, nblk(n), BLK, 0, view(c->pa), n, c->mig_dest.p, c->mig_midx.p, W, poly, d_ch, c->mig_off.p, c->mig_off.p + R,
           c->mig_sp.p, c->mig_sh.p);
  c->comm.alltoallv_bytes(c->mig_sp.p, sp_off.data(), mine.data(), c->mig_rp.p, rp_off.data(), rp_cnt.data(), sizeof(MigPart), c->st);
  c->comm.alltoallv_bytes(c->mig_sh.p, sh_off.data(), mine.data() + R, c->mig_rh.p, rh_off.data(), rh_cnt.data(), sizeof(MigRec), c->st);
  // immigrants are appended behind the old real range (the ghost range is rebuilt anyway)
  const uint32_t n_all = n + n_in;
  c->pa.ensure(n_all + 1, true, c->st);
  c->mig_dest.ensure(n_all + 1, true, c->st);
  if (n_in) LAUNCH(c, k_mig_unpack, nblk(n_in), BLK, 0, view(c->pa), n, n_in, c->mig_rp.p, c->mig_dest.p);
  // immigrant history table: CSR by immigrant index
  c->imm_cnt.ensure(n_in + 2); c->imm_off.ensure(n_in + 2); c->imm_tab.ensure(h_in + 1);
  CK(cudaMemsetAsync(c->imm_cnt.p, 0, (n_in + 1) * 4, c->st));
  for (int r = 0; r < R; r++)
    if (rh_cnt[r]) LAUNCH(c, k_imm_hist<false>, nblk(rh_cnt[r]), BLK, 0, c->mig_rh.p, rh_off[r], rh_cnt[r], rp_off[r], c->imm_cnt.p, nullptr, nullptr);
  scan_u32(c, c->imm_cnt.p, c->imm_off.p, n_in);
  CK(cudaMemsetAsync(c->imm_cnt.p, 0, (n_in + 1) * 4, c->st));
  for (int r = 0; r < R; r++)
    if (rh_cnt[r]) LAUNCH(c, k_imm_hist<true>, nblk(rh_cnt[r]), BLK, 0, c->mig_rh.p, rh_off[r], rh_cnt[r], rp_off[r], c->imm_cnt.p, c->imm_off.p, c->imm_tab.p);
  return n_all;
}

// update_interaction_ghost (src/interaction/update_interaction_ghost.cpp:58-285): ship every OwnerGhost record to the brick
// where its partner is real and append it there as a PartnerGhost record (see kernels_twin.cuh).  Runs after the history
// carry-over so the payload travels with the record.  No OwnerGhost records exist for spheres with symetric:true.
static void update_interaction_ghost(edb_context* c) {
  const bool poly = c->poly();
  if (!poly && c->symetric) return;
  const int tmax = poly ? 3 : 0;
  const int R = c->cfg.nranks, me = c->cfg.rank;
  const bool multi = R > 1;
  const uint32_t n = c->nreal, nloc = c->comm.n_local_ghost;
  // destinations: 0 = periodic images of the own brick, then the peers in rank order
  TwinDest D{};
  std::vector<Peer*> dpeer;
  D.n = 1; D.lo[0] = n; D.hi[0] = n + nloc;
  dpeer.push_back(nullptr);
  if (multi)
    for (auto& kv : c->comm.peers) {
      Peer& p = kv.second;
      if (!p.n_recv) continue;
      if (D.n >= TWIN_MAXD) throw StateError{EDB_ERR_CAPACITY, "more than 27 neighbouring ranks"};
      D.lo[D.n] = n + nloc + p.recv_base; D.hi[D.n] = D.lo[D.n] + p.n_recv;
      dpeer.push_back(&p);
      D.n++;
    }
  const int nd = D.n;
  c->tw_cnt.ensure((size_t)(tmax + 1) * TWIN_MAXD);
  CK(cudaMemsetAsync(c->tw_cnt.p, 0, (size_t)(tmax + 1) * TWIN_MAXD * 4, c->st));
  uint64_t nmax = 0;
  for (int t = 0; t <= tmax; t++) {
    nmax = std::max<uint64_t>(nmax, c->wave[t].n);
    if (c->wave[t].n) LAUNCH(c, k_twin_count, nblk(c->wave[t].n), BLK, 0, view(c->wave[t]), D, c->tw_cnt.p + (size_t)t * TWIN_MAXD);
  }
  std::vector<uint32_t> cnt((size_t)(tmax + 1) * TWIN_MAXD);
  CK(cudaMemcpyAsync(cnt.data(), c->tw_cnt.p, cnt.size() * 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  // send layout: per destination, ordered by wave then record index
  std::vector<uint32_t> doff(nd + 1, 0);
  for (int d = 0; d < nd; d++) {
    uint32_t s = 0;
    for (int t = 0; t <= tmax; t++) s += cnt[(size_t)t * TWIN_MAXD + d];
    doff[d + 1] = doff[d] + s;
  }
  const uint32_t n_out = doff[nd];
  c->tw_s.ensure(n_out + 1);
  c->tw_flag.ensure(nmax + 2); c->tw_pos.ensure(nmax + 2);
  for (int d = 0; d < nd; d++) {
    uint32_t base = doff[d];
    for (int t = 0; t <= tmax; t++) {
      const uint32_t m = cnt[(size_t)t * TWIN_MAXD + d];
      if (!m) continue;
      Wave& w = c->wave[t];
      LAUNCH(c, k_twin_flag, nblk(w.n), BLK, 0, view(w), D.lo[d], D.hi[d], c->tw_flag.p);
      scan_u32(c, c->tw_flag.p, c->tw_pos.p, (uint32_t)w.n);
      LAUNCH(c, k_twin_pack, nblk(w.n), BLK, 0, view(w), t, poly, D.lo[d], D.hi[d], c->tw_pos.p, c->pa.id.p, d == 0 ? c->g_src.p : nullptr, n,
             c->tw_s.p + base);
      base += m;
    }
  }
  // receive layout: own-brick messages first, then one range per source rank
  std::vector<uint32_t> scnt(R, 0), soff(R, 0), rcnt(R, 0), roff(R + 1, 0);
  uint32_t n_in = doff[1];
  if (multi) {
    for (int d = 1; d < nd; d++) { scnt[dpeer[d]->rank] = doff[d + 1] - doff[d]; soff[dpeer[d]->rank] = doff[d]; }
    std::vector<uint32_t> all((size_t)R * R);
    c->comm.allgather_u32(scnt.data(), R, all.data(), c->st);
    for (int r = 0; r < R; r++) {
      rcnt[r] = all[(size_t)r * R + me];
      roff[r] = n_in;
      n_in += rcnt[r];
    }
  }
  c->tw_r.ensure(n_in + 1);
  if (doff[1]) CK(cudaMemcpyAsync(c->tw_r.p, c->tw_s.p, (size_t)doff[1] * sizeof(TwinMsg), cudaMemcpyDeviceToDevice, c->st));
  if (multi) c->comm.alltoallv_bytes(c->tw_s.p, soff.data(), scnt.data(), c->tw_r.p, roff.data(), rcnt.data(), sizeof(TwinMsg), c->st);
  if (!n_in) return;
  // resolve the slots on this side
  c->tw_own.ensure(n_in + 1); c->tw_par.ensure(n_in + 1);
  if (doff[1])
    LAUNCH(c, k_twin_resolve, nblk(doff[1]), BLK, 0, c->tw_r.p, doff[1], nullptr, 0u, c->cidx.p, c->cbeg.p, c->cend.p, c->G, c->pa.id.p, n,
           c->tw_own.p, c->tw_par.p);
  if (multi)
    for (auto& kv : c->comm.peers) {
      Peer& p = kv.second;
      const uint32_t m = rcnt[p.rank];
      if (!m) continue;
      LAUNCH(c, k_twin_resolve, nblk(m), BLK, 0, c->tw_r.p + roff[p.rank], m, p.d_smap.p, p.n_send, c->cidx.p, c->cbeg.p, c->cend.p, c->G,
             c->pa.id.p, n, c->tw_own.p + roff[p.rank], c->tw_par.p + roff[p.rank]);
    }
  // per-type order-preserving append
  c->tw_flag.ensure(n_in + 2);
  c->tw_tpos.ensure((size_t)(tmax + 1) * (n_in + 2));
  for (int t = 0; t <= tmax; t++) {
    uint32_t* pos = c->tw_tpos.p + (size_t)t * (n_in + 2);
    LAUNCH(c, k_twin_tflag, nblk(n_in), BLK, 0, c->tw_r.p, c->tw_own.p, n_in, t, c->tw_flag.p);
    scan_u32(c, c->tw_flag.p, pos, n_in);
    CK(cudaMemcpyAsync(c->h_tot + t, pos + n_in, 4, cudaMemcpyDeviceToHost, c->st));
  }
  CK(cudaStreamSynchronize(c->st));
  for (int t = 0; t <= tmax; t++) {
    const uint32_t m = c->h_tot[t];
    if (!m) continue;
    Wave& w = c->wave[t];
    if (w.n + m >= (1ull << INC_SHIFT)) throw StateError{EDB_ERR_CAPACITY, "wave larger than 2^28 records"};
    w.ensure_keep(w.n + m + 1, poly, c->st);
    LAUNCH(c, k_twin_append, nblk(n_in), BLK, 0, c->tw_r.p, c->tw_own.p, c->tw_par.p, n_in, t, c->tw_tpos.p + (size_t)t * (n_in + 2), view(w),
           (uint32_t)w.n, poly, c->pa.id.p);
    w.n += m;
    c->twins += m;
  }
}

static void driver_op(edb_context* c, int op, double time, double dt, double system_mass);
static void driver_gravity_launch(edb_context* c, const double g[3]);
static void driver_force_reduce(edb_context* c);
static void driver_vertices(edb_context* c, bool all);
static void backup_drivers(edb_context* c);
static void driver_displ_launch(edb_context* c, double thr, unsigned long long* d_out, const unsigned long long* gate);
static void driver_force_sums(edb_context* c);
static void bonds_download(edb_context* c);
static void bonds_upload(edb_context* c);
static void bonds_retire_broken(edb_context* c);
static void bonds_rebuild_device(edb_context* c);
static void rebuild(edb_context* c, double rcut_inc, int symetric) {
  ScopedTimer tm(c, B_REBUILD);
  const GridDev& G = c->G;
  c->rcut_inc = rcut_inc;
  c->symetric = symetric;
  const bool poly = c->poly();
  if (poly && c->h_sh.empty()) throw StateError{EDB_ERR_STATE, "polyhedron mode needs edb_set_shapes"};
  // [check_rcut] (src/field/check_rcut.cpp:72-100): the neighbour kernels, the ghost layer and the mesh cell lists assume a one-cell
  // stencil, i.e. a pair within 2 r_max + rcut_inc never spans more than one cell.  A smaller cell would silently drop contacts.
  {
    const double need = 2.0 * c->rmax_upload + rcut_inc, have = std::min(G.cs[0], std::min(G.cs[1], G.cs[2]));
    if (c->rmax_upload > 0 && have < need * (1.0 - 1e-12)) {
      char b[256];
      snprintf(b, sizeof b, "cell size %.6g < 2 r_max + rcut_inc = %.6g: the one-cell neighbour stencil would miss contacts (check_rcut)", have, need);
      throw StateError{EDB_ERR_INVALID, b};
    }
  }

  // 0. multi-GPU: particles that left the brick migrate (with their interaction history) before re-binning
  const bool multi = c->cfg.nranks > 1;
  c->ghost_pushed = false;
  c->ghost_in_landing = 0;   // the ghost layer is rebuilt from scratch below
  c->comm.fan_epoch_pending = 0;
  c->have_fan = false;
  const uint32_t n_before = c->nreal;   // slots the previous waves refer to
  uint32_t n_all = c->nreal, n = c->nreal;
  c->n_imm = 0;
  if (multi) n_all = migrate(c, &n);
  const int* dead = multi ? c->mig_dest.p : nullptr;
  // 1. bin reals into cells, id-ordered inside each cell
  c->cidx.ensure(n_all + 1); c->tmp.ensure(n_all + 1); c->new2old.ensure(n_all + 1); c->old2new.ensure(n_all + 1);
  CK(cudaMemsetAsync(c->ccount.p, 0, (G.ncell + 1) * 4, c->st));
  CK(cudaMemsetAsync(c->cfill.p, 0, (G.ncell + 1) * 4, c->st));
  LAUNCH(c, k_cell_assign, nblk(n_all), BLK, 0, c->pa.pr.p, n_all, G, c->cidx.p, c->ccount.p, dead);
  scan_u32(c, c->ccount.p, c->cstart.p, G.ncell);
  LAUNCH(c, k_cell_scatter, nblk(n_all), BLK, 0, c->cidx.p, n_all, c->cstart.p, c->cfill.p, c->tmp.p);
  LAUNCH(c, k_cell_rank, nblk(n_all), BLK, 0, c->cidx.p, c->pa.id.p, n_all, c->cstart.p, c->tmp.p, c->new2old.p, c->old2new.p);
  // 2. ghosts: count, size, permute into the (possibly grown) scratch set, fill
  LAUNCH(c, k_ghost_count, nblk(c->n_gent), BLK, 0, c->g_src_cell.p, c->n_gent, c->cstart.p, c->g_count.p);
  scan_u32(c, c->g_count.p, c->g_off.p, c->n_gent);
  CK(cudaMemcpyAsync(c->h_tot, c->g_off.p + c->n_gent, 4, cudaMemcpyDeviceToHost, c->st));
  std::vector<uint32_t> h_peer;
  if (multi) {
    // remote ghosts: the source rank counts the particles of each cell it feeds and tells the ghost owner
    for (auto& kv : c->comm.peers) {
      Peer& p = kv.second;
      if (!p.send.empty()) LAUNCH(c, k_ghost_count, nblk(p.send.size()), BLK, 0, p.d_scell.p, (uint32_t)p.send.size(), c->cstart.p, p.d_scnt.p);
    }
    c->comm.exchange_ghost_counts(c->st);
    h_peer.assign(2 * c->comm.peers.size() + 2, 0);
    size_t i = 0;
    c->stage_u32.ensure(2 * c->comm.peers.size() + 2);
    for (auto& kv : c->comm.peers) {
      Peer& p = kv.second;
      scan_u32(c, p.d_scnt.p, p.d_soff.p, (uint32_t)p.send.size());
      CK(cudaMemcpyAsync(c->stage_u32.p + i, p.d_soff.p + p.send.size(), 4, cudaMemcpyDeviceToDevice, c->st));
      scan_u32(c, p.d_rcnt.p, p.d_roff.p, (uint32_t)p.recv.size());
      CK(cudaMemcpyAsync(c->stage_u32.p + i + 1, p.d_roff.p + p.recv.size(), 4, cudaMemcpyDeviceToDevice, c->st));
      i += 2;
    }
    CK(cudaMemcpyAsync(h_peer.data(), c->stage_u32.p, i * 4, cudaMemcpyDeviceToHost, c->st));
  }
  CK(cudaStreamSynchronize(c->st));
  const uint32_t nloc_ghost = c->n_gent ? c->h_tot[0] : 0;
  c->comm.n_local_ghost = nloc_ghost;
  c->nghost = nloc_ghost;
  if (multi) {
    size_t i = 0;
    uint32_t base = 0;
    for (auto& kv : c->comm.peers) {
      Peer& p = kv.second;
      p.n_send = h_peer[i]; p.n_recv = h_peer[i + 1]; p.recv_base = base;
      base += p.n_recv;
      i += 2;
      p.d_smap.ensure(p.n_send + 1); p.d_sent.ensure(p.n_send + 1);
      p.s_pr.ensure(p.n_send + 1); p.s_vm.ensure(p.n_send + 1); p.s_wg.ensure(p.n_send + 1); p.s_q.ensure(p.n_send + 1);
      p.s_h.ensure(p.n_send + 1); p.s_id.ensure(p.n_send + 1);
    }
    c->nghost += base;
    c->comm.p2p_prepare(base, c->st);
  }
  const uint32_t ntot = n + c->nghost;
  c->pb.ensure(ntot + 1, false, c->st);
  c->cidx2.ensure(ntot + 1);
  LAUNCH(c, k_permute, nblk(n), BLK, 0, view(c->pb), view(c->pa), c->new2old.p, c->cidx.p, c->cidx2.p, n);
  std::swap(c->pa, c->pb);
  std::swap(c->cidx, c->cidx2);
  c->nreal = n;
  if (multi) LAUNCH(c, k_iota, nblk(n), BLK, 0, c->pa.orig.p, n);  // download order = device order once particles migrate
  LAUNCH(c, k_cell_ranges, nblk(G.ncell), BLK, 0, c->cstart.p, G.ncell, c->cbeg.p, c->cend.p);
  c->g_src.ensure(nloc_ghost + 1); c->g_ent.ensure(nloc_ghost + 1);
  LAUNCH(c, k_ghost_fill, nblk(c->n_gent), BLK, 0, c->g_cell.p, c->g_src_cell.p, c->n_gent, c->cstart.p, c->g_off.p, n, c->g_src.p,
         c->g_ent.p, c->cbeg.p, c->cend.p);
  if (multi)
    for (auto& kv : c->comm.peers) {
      Peer& p = kv.second;
      if (!p.send.empty()) LAUNCH(c, k_send_map, nblk(p.send.size()), BLK, 0, p.d_scell.p, (uint32_t)p.send.size(), c->cstart.p, p.d_soff.p, p.d_smap.p, p.d_sent.p);
      if (!p.recv.empty()) LAUNCH(c, k_recv_cells, nblk(p.recv.size()), BLK, 0, p.d_rgcell.p, (uint32_t)p.recv.size(), p.d_roff.p, p.d_rcnt.p,
                                  n + nloc_ghost + p.recv_base, c->cbeg.p, c->cend.p, c->cidx.p);
    }
  ghost_refresh(c, true);
  compute_vertices(c);
  // 2b. ghost fan-out table for the fused per-step refresh (single GPU: periodic images; multi-GPU: needs the peer-memory path)
  if ((!multi && nloc_ghost) || (multi && c->comm.p2p.on)) {
    c->fan_cnt.ensure(n + 2); c->fan_off.ensure(n + 2); c->fan_bits.ensure(n / 32 + 2);
    CK(cudaMemsetAsync(c->fan_bits.p, 0, (n / 32 + 1) * 4, c->st));
    int pi = 0;
    if (multi)
      for (auto& kv : c->comm.peers) kv.second.fan_index = kv.second.n_send ? pi++ : -1;
    for (int fill = 0; fill < 2; fill++) {
      CK(cudaMemsetAsync(c->fan_cnt.p, 0, (n + 1) * 4, c->st));
      if (nloc_ghost) {
        if (!fill) LAUNCH(c, k_fan_build<false>, nblk(nloc_ghost), BLK, 0, c->g_src.p, c->g_ent.p, c->g_shift.p, nloc_ghost, n, -1, c->fan_cnt.p, nullptr, nullptr, c->fan_bits.p);
        else LAUNCH(c, k_fan_build<true>, nblk(nloc_ghost), BLK, 0, c->g_src.p, c->g_ent.p, c->g_shift.p, nloc_ghost, n, -1, c->fan_cnt.p, c->fan_off.p, c->fan_ent.p, c->fan_bits.p);
      }
      if (multi)
        for (auto& kv : c->comm.peers) {
          Peer& p = kv.second;
          if (!p.n_send) continue;
          if (!fill) LAUNCH(c, k_fan_build<false>, nblk(p.n_send), BLK, 0, p.d_smap.p, p.d_sent.p, p.d_sshift.p, p.n_send, 0u, p.fan_index, c->fan_cnt.p, nullptr, nullptr, c->fan_bits.p);
          else LAUNCH(c, k_fan_build<true>, nblk(p.n_send), BLK, 0, p.d_smap.p, p.d_sent.p, p.d_sshift.p, p.n_send, 0u, p.fan_index, c->fan_cnt.p, c->fan_off.p, c->fan_ent.p, c->fan_bits.p);
        }
      if (!fill) {
        scan_u32(c, c->fan_cnt.p, c->fan_off.p, n);
        CK(cudaMemcpyAsync(c->h_tot, c->fan_off.p + n, 4, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        c->fan_ent.ensure((size_t)c->h_tot[0] + 1);
      }
    }
    c->have_fan = true;
  }
  // 3. backup for the displacement trigger (backup_r / backup_dem)
  c->pr_backup.ensure(n + 1);
  if (poly) c->q_backup.ensure(n + 1);
  LAUNCH(c, k_backup, nblk(n), BLK, 0, c->pa.pr.p, c->pa.q.p, c->pr_backup.p, poly ? c->q_backup.p : nullptr, n);
  c->have_backup = true;

  // 4. enumerate candidates: count -> scan -> fill (per type, owner-sorted)
  for (int t = 0; t < NT; t++) std::swap(c->wave[t], c->prev[t]);
  const bool had = c->have_prev;  // prev now holds the waves of the previous build (valid if had)
  bool active[NT] = {};
  bool has_drv[3] = {false, false, false};
  for (auto& D : c->h_drv)
    if (D.kind >= 0 && D.kind <= 2) has_drv[D.kind] = true;
  active[0] = true;
  if (poly) active[1] = active[2] = active[3] = true;
  for (int k = 0; k < 3; k++) active[4 + k] = has_drv[k];
  if (c->drivers_moving) backup_drivers(c);                     // update_particles_full: backup_drivers (config_move_particles.msp:150-153)
  if (c->mesh_dirty || c->drivers_moving) build_mesh_grid(c);   // grid_rshape_driver { force_reset: true } (:92-94)
  if (c->M.nmesh)
    for (int t = 7; t <= (poly ? 12 : 9); t++) active[t] = true;
  CntTab C{};
  for (int t = 0; t < NT; t++) {
    C.c[t] = nullptr;
    c->wave[t].n = 0;
    if (active[t]) {
      c->wave[t].off.ensure(n + 2);
      C.c[t] = c->wave[t].off.p;
    }
  }
  NbhArgs A{view(c->pa), c->cidx.p, c->cbeg.p, c->cend.p, G, c->drv.p, (int)c->h_drv.size(), rcut_inc, symetric, n, c->vert.p, c->S, c->M, 0,
            nullptr, nullptr};
  if (!c->env_nbh_no_cull) {
    c->rmax_dev.ensure(2);
    CK(cudaMemsetAsync(c->rmax_dev.p, 0, 8, c->st));
    LAUNCH(c, k_max_radius, nblk(n + c->nghost), BLK, 0, c->pa.pr.p, n + c->nghost, c->rmax_dev.p);
    A.rmax_all = reinterpret_cast<const double*>(c->rmax_dev.p);
  }
  if (!c->h_bonds.empty()) {
    // persistent interactions (nbh_polyhedron.cpp:146-149): the bonds follow their particles into the new order by id, and the pairs
    // they glue are ignored by the contact enumeration while they hold
    if (c->bonds_on_device && !multi && !c->env_bonds_host) {
      bonds_rebuild_device(c);  // slots follow old2new, stable sort by new owner, broken bonds cut off / queued, tables rebuilt: all on the device
    } else {
      bonds_download(c);
      bonds_retire_broken(c);   // broken bonds leave wave 13; those with friction are queued as vertex-vertex history (below, after k_history)
      bonds_upload(c);
    }
    if (c->have_ignore) { A.ign_off = c->b_ign_off.p; A.ign_ids = c->b_ign_ids.p; }
  }
  WTab W{};
  // polyhedra: the particle-particle waves go through the cell-block kernel (TMA-staged stencil, warp-shared pair tests) when the
  // largest stencil fits the shared memory of an SM; drivers / meshes stay on the per-particle kernel
  uint32_t cell_cap = 0;
  size_t cell_smem = 0;
  const uint32_t n_real_cells = (uint32_t)G.nc[0] * G.nc[1] * G.nc[2];
  bool drv_waves = false;
  for (int t = 4; t < NT; t++) drv_waves = drv_waves || active[t];
  if (poly && c->S.nvmax <= 32 && c->shape_nemax <= 32 && !c->env_nbh_per_thread) {
    CK(cudaMemsetAsync(c->d_flag + 1, 0, 8, c->st));
    LAUNCH(c, k_stencil_max, nblk(n_real_cells), BLK, 0, G, c->cbeg.p, c->cend.p, c->d_flag + 1);
    CK(cudaMemcpyAsync(c->h_flag + 1, c->d_flag + 1, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint32_t cap = (uint32_t)c->h_flag[1] + 1;
    const size_t bytes = (size_t)cap * npc_bytes_per_particle(c->S.nvmax) + 64;
    if (bytes + 4096 <= c->smem_optin) {
      cell_cap = cap;
      cell_smem = bytes;
      CK(cudaFuncSetAttribute(k_nbh_poly_cell<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      CK(cudaFuncSetAttribute(k_nbh_poly_cell<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      CK(cudaFuncSetAttribute(k_nbh_poly_cell<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
  }
  A.skip_pp = cell_cap ? 1 : 0;
  // the count pass records its findings (pair headers + test ballots) when the pool sized by the previous rebuild is in place;
  // the fill pass then replays them instead of repeating the tests
  NbhPool PL{};
  bool record = false;
  if (cell_cap) {
    c->nbh_ctr.ensure(8);   // [0] cursor  [1] overflow  [2,3] need (u64)
    CK(cudaMemsetAsync(c->nbh_ctr.p, 0, 32, c->st));
    c->nbh_first.ensure((size_t)n * NPC_PARTS + 1);
    record = c->nbh_pool_chunks > 0 && !c->env_nbh_no_replay;
    PL = NbhPool{c->nbh_pool.p, c->nbh_ctr.p, record ? c->nbh_pool_chunks : 0u, c->nbh_first.p,
                 reinterpret_cast<unsigned long long*>(c->nbh_ctr.p + 2), c->nbh_ctr.p + 1};
  }
  if (!poly) {
    if (!c->env_nbh_two_pass && (size_t)n * PP_TMP_CAP * sizeof(uint2) <= ((size_t)3 << 30)) {
      c->pp_tmp.ensure((size_t)n * PP_TMP_CAP + 1);
      c->nbh_ctr.ensure(8);
      CK(cudaMemsetAsync(c->nbh_ctr.p, 0, 32, c->st));
      A.pp_tmp = c->pp_tmp.p;
      A.pp_overflow = c->nbh_ctr.p + 1;
    }
    if (A.pp_tmp && c->env_nbh_warp) {
      // wave 0 by a warp per particle; the driver / mesh waves by the per-particle kernel without its stencil walk
      LAUNCH(c, k_nbh_sphere_warp, nblk((uint64_t)n * 32, 256), 256, 0, A, C.c[0]);
      bool other = false;
      for (int t = 4; t < NT; t++) other = other || active[t];
      if (other) {
        A.skip_pp = 1;
        LAUNCH(c, k_nbh_sphere<false>, nblk(n, 128), 128, 0, A, C, W);
        A.skip_pp = 0;
      }
    } else {
      LAUNCH(c, k_nbh_sphere<false>, nblk(n, 128), 128, 0, A, C, W);
    }
    if (A.pp_tmp) CK(cudaMemcpyAsync(c->h_nbh, c->nbh_ctr.p, 32, cudaMemcpyDeviceToHost, c->st));
  } else {
    if (cell_cap) {
      if (NPC_PARTS > 1)   // the parts of an owner add their counts
        for (int t = 0; t < 4; t++)
          if (C.c[t]) CK(cudaMemsetAsync(C.c[t], 0, (size_t)(n + 1) * 4, c->st));
      if (record) LAUNCH(c, k_nbh_poly_cell<1>, n_real_cells, NPC_THREADS, cell_smem, A, C, W, cell_cap, PL);
      else LAUNCH(c, k_nbh_poly_cell<0>, n_real_cells, NPC_THREADS, cell_smem, A, C, W, cell_cap, PL);
      CK(cudaMemcpyAsync(c->h_nbh, c->nbh_ctr.p, 32, cudaMemcpyDeviceToHost, c->st));
    }
    if (!cell_cap || drv_waves) LAUNCH(c, k_nbh_poly<false>, nblk(n, 64), 64, 0, A, C, W);
  }
  for (int t = 0; t < NT; t++)
    if (active[t]) {
      scan_u32(c, C.c[t], C.c[t], n);
      CK(cudaMemcpyAsync(c->h_tot + t, C.c[t] + n, 4, cudaMemcpyDeviceToHost, c->st));
    }
  CK(cudaStreamSynchronize(c->st));
  for (int t = 0; t < NT; t++) {
    if (!active[t]) continue;
    c->wave[t].n = c->h_tot[t];
    if (c->wave[t].n >= (1ull << INC_SHIFT)) throw StateError{EDB_ERR_CAPACITY, "wave larger than 2^28 records"};
    c->wave[t].ensure(c->wave[t].n + 1, poly || t >= 7);
    W.w[t] = view(c->wave[t]);
  }
  if (!poly) {
    // wave 0 is a copy of what the count pass recorded (unless a particle overflowed its slots); the driver / mesh waves are
    // enumerated again — a loop over the drivers per particle
    const bool copy0 = A.pp_tmp && c->h_nbh[1] == 0;
    bool other = false;
    for (int t = 4; t < NT; t++) other = other || (active[t] && c->wave[t].n);
    A.skip_pp = copy0 ? 1 : 0;
    if (!copy0 || other) LAUNCH(c, k_nbh_sphere<true>, nblk(n, 128), 128, 0, A, C, W);
    if (copy0 && c->wave[0].n) LAUNCH(c, k_nbh_sphere_copy, nblk(n), BLK, 0, view(c->pa), c->pp_tmp.p, C.c[0], n, W.w[0]);
  } else {
    if (cell_cap) {
      const uint32_t overflow = c->h_nbh[1];
      const unsigned long long need = *reinterpret_cast<unsigned long long*>(c->h_nbh + 2);
      if (record && !overflow) {
        LAUNCH(c, k_nbh_poly_replay, nblk(n, 8), 256, 0, A, C, W, PL);
        c->nbh_replays++;
      } else {
        LAUNCH(c, k_nbh_poly_cell<2>, n_real_cells, NPC_THREADS, cell_smem, A, C, W, cell_cap, PL);
      }
      if (c->trace_async)
        fprintf(stderr, "[edb trace] poly rebuild %llu: cell_cap %u smem %zu record %d overflow %u need %llu pool %u chunks\n",
                (unsigned long long)c->ctr.rebuilds, cell_cap, cell_smem, (int)record, overflow, need, c->nbh_pool_chunks);
      // size the pool for the next rebuild (25 % head room)
      const unsigned long long want = need + need / 4 + 1024;
      if (need + need / 16 + 64 > c->nbh_pool_chunks && want * CHW < (1ull << 32)) {   // grow only when the head room is used up
        CK(cudaStreamSynchronize(c->st));
        c->nbh_pool.ensure((size_t)want * CHW);
        c->nbh_pool_chunks = (uint32_t)want;
      }
    }
    if (!cell_cap || drv_waves) LAUNCH(c, k_nbh_poly<true>, nblk(n, 64), 64, 0, A, C, W);
  }
  // 5. history carry-over
  if (had || c->n_imm) {
    const ImmHist H{c->imm_off.p, (const ImmRec*)c->imm_tab.p, c->n_imm};
    for (int t = 0; t < NT; t++)
      if (active[t] && c->wave[t].n && ((had && c->prev[t].n) || c->n_imm)) {
        WView ow = view(c->prev[t]);
        if (!had) ow.n = 0;
        LAUNCH(c, k_history, nblk(c->wave[t].n), BLK, 0, view(c->wave[t]), ow, c->new2old.p, had ? c->n_old_real : n_before, poly || t >= 7, t, H);
      }
  }
  if (c->n_conv_dev) {
    if (c->wave[0].n) LAUNCH(c, k_bond_convert, nblk(c->n_conv_dev), BLK, 0, c->d_conv.p, c->n_conv_dev, view(c->wave[0]));
    c->n_conv_dev = 0;
  }
  if (!c->h_conv.empty()) {
    if (c->wave[0].n) {
      c->d_conv.ensure(c->h_conv.size());
      CK(cudaMemcpyAsync(c->d_conv.p, c->h_conv.data(), c->h_conv.size() * sizeof(BondConv), cudaMemcpyHostToDevice, c->st));
      LAUNCH(c, k_bond_convert, nblk(c->h_conv.size()), BLK, 0, c->d_conv.p, (uint32_t)c->h_conv.size(), view(c->wave[0]));
      CK(cudaStreamSynchronize(c->st));   // (h_conv is pageable host memory)
    }
    c->h_conv.clear();
  }
  c->have_prev = true;
  c->n_old_real = n;
  // 5b. OwnerGhost records -> PartnerGhost twins where the partner is real
  update_interaction_ghost(c);
  // 6. incidence lists (owner + partner entries, canonical order) for the deterministic gather
  {
    const bool partner = poly || symetric;
    c->have_inc = false;
    c->inc_cnt.ensure(n + 2); c->inc_off.ensure(n + 2);
    OrderTab O{};
    const int* order = poly ? ORDER_POLY : ORDER_SPHERE;
    O.norder = poly ? 13 : 7;
    for (int o = 0; o < O.norder; o++) O.order[o] = order[o];
    WTab WT{};
    uint64_t tot = 0;
    for (int t = 0; t < NT; t++) { WT.w[t] = view(c->wave[t]); tot += c->wave[t].n; }
    const int tmax = partner ? (poly ? 3 : 0) : -1;
    if (partner) for (int t = 0; t <= tmax; t++) tot += c->wave[t].n;
    LAUNCH(c, k_inc_owner<false>, nblk(n), BLK, 0, WT, O, n, c->inc_cnt.p, nullptr, nullptr);
    for (int t = 0; t <= tmax; t++)
      if (c->wave[t].n) LAUNCH(c, k_incidence<false>, nblk(c->wave[t].n), BLK, 0, view(c->wave[t]), t, n, c->inc_cnt.p, nullptr, nullptr);
    scan_u32(c, c->inc_cnt.p, c->inc_off.p, n);
    c->inc.ensure(tot + 1);
    LAUNCH(c, k_inc_owner<true>, nblk(n), BLK, 0, WT, O, n, c->inc_cnt.p, c->inc_off.p, c->inc.p);
    for (int t = 0; t <= tmax; t++)
      if (c->wave[t].n) LAUNCH(c, k_incidence<true>, nblk(c->wave[t].n), BLK, 0, view(c->wave[t]), t, n, c->inc_cnt.p, c->inc_off.p, c->inc.p);
    if (partner) LAUNCH(c, k_incidence_sort, nblk(n), BLK, 0, c->inc_off.p, c->inc.p, n);
    c->have_inc = true;
  }
  c->ctr.rebuilds++;
  c->have_cells = true;
}

extern "C" int edb_update_particle_neighbors(edb_context* c, double rcut_inc, int32_t symetric) {
  API_BEGIN(c)
  rebuild(c, rcut_inc, symetric);
  API_END(c)
}
extern "C" int edb_ghost_update_rq(edb_context* c) {
  API_BEGIN(c)
  ghost_refresh(c, false);
  API_END(c)
}
extern "C" int edb_compute_vertices(edb_context* c) {
  API_BEGIN(c)
  compute_vertices(c);
  API_END(c)
}

// ---- per-step operators -----------------------------------------------------------------------------------------
extern "C" int edb_reset_force_moment(edb_context* c) {
  API_BEGIN(c)
  LAUNCH(c, k_reset_fm, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal);
  API_END(c)
}
extern "C" int edb_gravity_force(edb_context* c, const double g[3]) {
  API_BEGIN(c)
  LAUNCH(c, k_gravity, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, g[0], g[1], g[2]);
  driver_gravity_launch(c, g);   // "particles and drivers": PARTICLE RShape drivers (gravity_force.cu:72-83)
  API_END(c)
}
extern "C" int edb_force_to_accel(edb_context* c) {
  API_BEGIN(c)
  LAUNCH(c, k_force_to_accel, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal);
  API_END(c)
}
static void prolog(edb_context* c, double dt) {
  ScopedTimer tm(c, B_PROLOG);
  LAUNCH(c, k_prolog, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, dt, dt * dt * 0.5, dt * 0.5);
}
extern "C" int edb_combined_compute_prolog(edb_context* c, double dt) {
  API_BEGIN(c)
  prolog(c, dt);
  API_END(c)
}
extern "C" int edb_combined_compute_epilog(edb_context* c, double dt) {
  API_BEGIN(c)
  ScopedTimer tm(c, B_EPILOG);
  LAUNCH(c, k_epilog<2>, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, dt * 0.5);
  API_END(c)
}
extern "C" int edb_push_to_quaternion(edb_context* c, double dt) {
  API_BEGIN(c)
  LAUNCH(c, k_push_to_quaternion, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, dt, dt * 0.5);
  API_END(c)
}
extern "C" int edb_push_to_angular_velocity(edb_context* c, double dt) {
  API_BEGIN(c)
  LAUNCH(c, k_epilog<1>, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, dt * 0.5);
  API_END(c)
}
extern "C" int edb_push_to_angular_acceleration(edb_context* c) {
  API_BEGIN(c)
  LAUNCH(c, k_epilog<0>, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, 0.0);
  API_END(c)
}
extern "C" int edb_set_particle_regions(edb_context* c, int32_t n, const edb_region* r) {
  API_BEGIN(c)
  if (n < 0 || (n > 0 && !r)) throw StateError{EDB_ERR_INVALID, "edb_set_particle_regions: bad region table"};
  quiesce(c);
  std::vector<RegionDev> h((size_t)n);
  for (int i = 0; i < n; i++) {
    RegionDev& d = h[i];
    for (int k = 0; k < 3; k++) { d.lo[k] = r[i].lo[k]; d.hi[k] = r[i].hi[k]; }
    for (int k = 0; k < 16; k++) d.Q[k] = r[i].quadric[k];
    d.id0 = r[i].id_start; d.id1 = r[i].id_end;
    d.use_bounds = r[i].use_bounds; d.use_ids = r[i].use_ids; d.use_quadric = r[i].use_quadric; d.pad = 0;
  }
  c->d_regions.ensure((size_t)n + 1);
  if (n) CK(cudaMemcpy(c->d_regions.p, h.data(), (size_t)n * sizeof(RegionDev), cudaMemcpyHostToDevice));
  c->n_regions = n;
  API_END(c)
}
static RegionExpr region_expr(edb_context* c, int32_t n_expr, const int32_t* expr) {
  if (n_expr < 0 || n_expr > EDB_CSG_MAX || (n_expr > 0 && !expr)) throw StateError{EDB_ERR_INVALID, "region expression: 0..EDB_CSG_MAX tokens"};
  RegionExpr E{};
  E.reg = c->d_regions.p;
  E.n = n_expr;
  int sp = 0;
  for (int q = 0; q < n_expr; q++) {
    const int t = expr[q];
    if (t >= 0) { if (t >= c->n_regions) throw StateError{EDB_ERR_INVALID, "region expression names a region that was not set"}; sp++; }
    else if (t == EDB_CSG_NOT) { if (sp < 1) throw StateError{EDB_ERR_INVALID, "malformed region expression"}; }
    else if (t == EDB_CSG_AND || t == EDB_CSG_OR) { if (sp < 2) throw StateError{EDB_ERR_INVALID, "malformed region expression"}; sp--; }
    else throw StateError{EDB_ERR_INVALID, "unknown token in region expression"};
    E.tok[q] = t;
  }
  if (n_expr && sp != 1) throw StateError{EDB_ERR_INVALID, "malformed region expression"};
  return E;
}
extern "C" int edb_reset_force_moment_region(edb_context* c, int32_t n_expr, const int32_t* expr) {
  API_BEGIN(c)
  const RegionExpr E = region_expr(c, n_expr, expr);
  LAUNCH(c, k_region_set<false>, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, E);
  API_END(c)
}
extern "C" int edb_freeze_particles_region(edb_context* c, int32_t n_expr, const int32_t* expr) {
  API_BEGIN(c)
  const RegionExpr E = region_expr(c, n_expr, expr);
  LAUNCH(c, k_region_set<true>, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, E);
  API_END(c)
}
extern "C" int edb_freeze_particles(edb_context* c) {
  API_BEGIN(c)
  LAUNCH(c, k_freeze, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal);
  API_END(c)
}


template <int LAW>
static void launch_contact_waves(edb_context* c, double dt, const unsigned long long* gate, unsigned long long* zero_flag) {
  const bool poly = c->poly();
  ContactArgs A{view(c->pa), WView{}, c->prm, c->drv.p, c->vert.p, c->S, dt, 0, c->symetric, c->M, gate, zero_flag,
                nullptr, nullptr, nullptr, 0xffffffffu, c->ptype.p};
  if (c->ghost_in_landing) {   // spheres on several GPUs, inside edb_run_verlet: remote ghosts are read from the landing zone
    const uint32_t lo = c->nreal + c->comm.n_local_ghost;
    const double4* land = c->comm.landing_of(c->ghost_in_landing);
    A.rem_lo = lo;
    A.g_pr = land - lo; A.g_vm = land + (size_t)c->comm.p2p.cap - lo; A.g_wg = land + 2 * (size_t)c->comm.p2p.cap - lo;
  }
  // the gather of the previous step left the chunk queue at zero when it ran as k_gather<2>; otherwise it is cleared here
  const bool queue_clean = c->queue_clean;
  c->queue_clean = false;
  const int* order = poly ? ORDER_POLY : ORDER_SPHERE;
  const int norder = poly ? 13 : 7;
  bool pp_done = false;
  // spheres: the driver / mesh waves are small (one launch of a few blocks each) and independent of wave 0 — they run on a side
  // stream beside the persistent wave-0 kernel instead of after it
  bool side = false;
  if (!poly && c->wave[0].n) {
    for (int t = 4; t <= 9; t++) side = side || c->wave[t].n > 0;
    if (side) {
      CK(cudaEventRecord(c->ev_fork, c->st));
      CK(cudaStreamWaitEvent(c->st2, c->ev_fork, 0));
    }
  }
  cudaStream_t sd = side ? c->st2 : c->st;
  for (int o = 0; o < norder; o++) {
    const int t = order[o];
    Wave& w = c->wave[t];
    if (!w.n) continue;
    A.w = view(w);
    A.type = t;
    if (t <= 3) {
      ScopedTimer tpp(c, B_CONTACT_PP);
      if (poly) {
        if (t == order[0] || !pp_done) {   // one detection + one law launch cover the four particle-particle waves
          PolyWaves PW{};
          uint32_t tot = 0;
          for (int q = 0; q < 4; q++) { PW.w[q] = view(c->wave[q]); PW.start[q] = tot; tot += (uint32_t)c->wave[q].n; }
          PW.start[4] = tot;
          c->need.ensure((size_t)tot + 1);
          const NeedList L{c->need.p, c->d_need_count};
          CK(cudaMemsetAsync(c->d_need_count, 0, 4, c->st));
          LAUNCH(c, k_detect_poly_pp<LAW>, nblk(tot, BLK * PDPT), BLK, 0, A, PW, L);
          LAUNCH(c, k_law_poly_pp<LAW>, (uint32_t)std::min<uint64_t>(nblk(tot), (uint64_t)c->n_sm * 2), BLK, 0, A, PW, L);
          pp_done = true;
        }
      } else {
        if (!queue_clean) CK(cudaMemsetAsync(c->d_need_count, 0, 4, c->st));
#if EDB_PP_PIPE
        LAUNCH(c, k_contact_sphere_pp_pipe<LAW>, (uint32_t)std::min<uint64_t>(nblk(w.n, CHUNK * (PP_BLK / 32)), (uint64_t)c->n_sm * PP_MINB), PP_BLK, 0, A, c->d_need_count);
#else
        LAUNCH(c, k_contact_sphere_pp<LAW>, (uint32_t)std::min<uint64_t>(nblk(w.n, CHUNK * (PP_BLK / 32)), (uint64_t)c->n_sm * PP_MINB), PP_BLK, 0, A, c->d_need_count);
#endif
      }
      c->ctr.contact_pp_launches++;
      c->ctr.contact_pp_records += w.n;
    } else if (t <= 6) {
      if (poly) LAUNCH(c, (k_contact_driver<LAW, true>), nblk(w.n), BLK, 0, A);
      else LAUNCH_ON(c, sd, (k_contact_driver<LAW, false>), nblk(w.n), BLK, 0, A);
    } else if (t <= 12) {
      if (poly) LAUNCH(c, (k_contact_mesh<LAW, true>), nblk(w.n), BLK, 0, A);
      else LAUNCH_ON(c, sd, (k_contact_mesh<LAW, false>), nblk(w.n), BLK, 0, A);
    }
    c->ctr.interaction_updates += w.n;
    c->ctr.contact_launches++;
    // algorithmic bytes of this launch (SURVEY 8d / DESIGN.md): per record 8 (ij) + 48 r + 48 w history + 80 contact state;
    // per distinct particle touched 92 B of gathers (r, radius, vrot, group, v, mass) counted once per wave.  Polyhedra add per
    // record 8 (sub) + 24 k of vertices, k = 2 V-V, 3 V-E, 1 + n_face V-F, 4 E-E (mesh waves alike; 1 vertex for a driver wave).
    uint64_t per_rec = 184;
    if (poly) {
      const double nf = c->mean_face_nv;
      const double k = (t == 0 || t == 7) ? 2 : (t == 1 || t == 8 || t == 11) ? 3 : (t == 2 || t == 9 || t == 12) ? 1 + nf : (t == 3 || t == 10) ? 4 : 1;
      per_rec += 8 + (uint64_t)(24.0 * k);
    }
    c->ctr.bytes_contact_algorithmic += w.n * per_rec + (uint64_t)c->nreal * 92ull;
  }
  if (side) {
    CK(cudaEventRecord(c->ev_join, c->st2));
    CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
  }
}

// zero_flag: returns through *zeroed whether a launched kernel clears it (spheres with a wave-0 launch); else the caller memsets
static void contact_waves(edb_context* c, double dt, const unsigned long long* gate = nullptr, unsigned long long* zero_flag = nullptr,
                          bool* zeroed = nullptr) {
  ScopedTimer tm(c, B_CONTACT);
  if (!c->have_prev) throw StateError{EDB_ERR_STATE, "contact operator before update_particle_neighbors"};
  const bool can_zero = zero_flag && !c->poly() && c->wave[0].n > 0;   // (a gated-off step neither zeroes nor accumulates)
  if (zeroed) *zeroed = can_zero;
  unsigned long long* zf = can_zero ? zero_flag : nullptr;
  if (c->law == 0) launch_contact_waves<0>(c, dt, gate, zf);
  else if (c->law == 1) launch_contact_waves<1>(c, dt, gate, zf);
  else launch_contact_waves<2>(c, dt, gate, zf);
}

// next_flag (FUSE 2 only): where the rebuild trigger of the next step is accumulated (nullptr: not evaluated)
template <int FUSE>
static void gather(edb_context* c, const double* g, double dt, const unsigned long long* gate = nullptr,
                   unsigned long long* next_flag = nullptr, double thr = 0.0, bool base_from_arrays = false) {
  ScopedTimer tm(c, B_GATHER);
  const bool poly = c->poly();
  GatherArgs A{};
  A.P = view(c->pa);
  for (int t = 0; t < NT; t++) A.W.w[t] = view(c->wave[t]);
  const int* order = poly ? ORDER_POLY : ORDER_SPHERE;
  A.norder = poly ? 13 : 7;
  for (int o = 0; o < A.norder; o++) A.order[o] = order[o];
  if (!c->have_inc) throw StateError{EDB_ERR_STATE, "contact operator before update_particle_neighbors"};
  A.inc_off = c->inc_off.p;
  A.inc = c->inc.p;
  A.npp = poly ? 4 : 1;
  A.nreal = c->nreal;
  A.has_swap = poly ? 1 : 0;
  if (g) { A.g[0] = g[0]; A.g[1] = g[1]; A.g[2] = g[2]; }
  A.hdt = dt * 0.5;
  A.dt = dt;
  A.dt2 = dt * dt * 0.5;
  A.gate = gate;
  A.next_flag = next_flag;
  A.base_from_arrays = base_from_arrays ? 1 : 0;
  A.reset_queue = (FUSE == 2 && !poly) ? c->d_need_count : nullptr;
  if (A.reset_queue) c->queue_clean = true;
  A.pr_backup = next_flag ? c->pr_backup.p : nullptr;
  A.q_backup = (next_flag && poly) ? c->q_backup.p : nullptr;
  A.S = c->S;
  A.vert = (FUSE == 2 && poly) ? c->vert.p : nullptr;
  A.thr2 = thr * thr;
  FanArgs F{};
  if (FUSE == 2 && c->have_fan && !c->env_no_fan) {
    F.bits = c->fan_bits.p;
    F.off = c->fan_off.p;
    F.ent = c->fan_ent.p;
    // EDB_GATHER_MAIL: the gather's last block mails the trigger count itself (measured on 2 GPUs: the per-block system-scope fence
    // it needs costs the gather 0.065 ms, against 0.006 ms saved in the landing launch) — default: the landing launch sends the mail
    c->comm.fan_begin(F, c->cfg.nranks > 1 && next_flag != nullptr && c->env_gather_mail);
    c->ghost_pushed = true;
  }
  if (poly) LAUNCH(c, (k_gather<FUSE, true>), nblk(c->nreal, G_BLK), G_BLK, 0, A, F);
  else LAUNCH(c, (k_gather<FUSE, false>), nblk(c->nreal, G_BLK), G_BLK, 0, A, F);
  c->ctr.gather_launches++;
}

extern "C" int edb_contact_sphere(edb_context* c, double dt) {
  API_BEGIN(c)
  if (c->poly()) throw StateError{EDB_ERR_STATE, "contact_sphere on a polyhedron context"};
  contact_waves(c, dt);
  gather<0>(c, nullptr, dt);
  API_END(c)
}
extern "C" int edb_contact_polyhedron(edb_context* c, double dt) {
  API_BEGIN(c)
  if (!c->poly()) throw StateError{EDB_ERR_STATE, "contact_polyhedron on a sphere context"};
  contact_waves(c, dt);
  gather<0>(c, nullptr, dt);
  API_END(c)
}

static uint64_t displ_over(edb_context* c, double thr) {
  if (!c->have_backup) throw StateError{EDB_ERR_STATE, "displ_over before update_particle_neighbors"};
  CK(cudaMemsetAsync(c->d_flag, 0, 8, c->st));
  if (c->poly()) LAUNCH(c, k_displ_over<true>, nblk(c->nreal), BLK, 0, view(c->pa), c->pr_backup.p, c->q_backup.p, c->nreal, c->S, thr * thr, c->d_flag, nullptr);
  else LAUNCH(c, k_displ_over<false>, nblk(c->nreal), BLK, 0, view(c->pa), c->pr_backup.p, nullptr, c->nreal, c->S, thr * thr, c->d_flag, nullptr);
  CK(cudaMemcpyAsync(c->h_flag, c->d_flag, 8, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  uint64_t v = c->h_flag[0];
  if (c->cfg.nranks > 1) v = c->comm.allreduce_sum_u64(v, c->st);
  return v;
}
extern "C" int edb_displ_over(edb_context* c, double thr, uint64_t* over) {
  API_BEGIN(c)
  *over = displ_over(c, thr);
  API_END(c)
}

// Rebuild trigger of a step: the producer (k_gather<2> of the previous step, or k_displ_over) leaves the local count in
// d_sflag[slot]; the ghost refresh of the step sums it over the ranks in place; trigger_publish copies it to h_sflag[slot]
// (pinned), valid once ev_sflag[slot] has completed.  The gated kernels of the step read d_sflag[slot] on the device.
__global__ void k_publish_flag(unsigned long long* __restrict__ host_mapped, const unsigned long long* __restrict__ dev) {
  *host_mapped = *dev;
  __threadfence_system();
}
static void trigger_publish(edb_context* c, int slot) {
  // A one-thread kernel stores the flag into mapped host memory: an 8-byte cudaMemcpyAsync between two kernels costs the stream a
  // round trip through the copy engine (measured: 0.219 -> 0.210 ms per step on the 1 M drum).
  // While the asynchronous host copies of a host-owned grid are in flight neither form may sit on the compute stream: the 8-byte copy
  // queues in the copy engine behind the 176 MB download of the previous step and the SM's store (with its system-scope fence) behind
  // the same bytes on the link, and every kernel of the step waits behind it (EDB_TRACE_ASYNC: the step ended 0.22 ms after the
  // previous download instead of 1 ms before it; with the store on the stream: e2e 4.2 -> 6.4 ms per step).  The store then goes to a
  // high-priority side stream: the step's kernels run at once, only the host's wait for the flag sees the link.
  if (c->flag_published) c->flag_published = false;   // (k_ghost_land of this step stored it)
  else if (c->st_up && !c->env_flag_on_stream) {
    if (!c->st_flag) {
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&c->st_flag, cudaStreamNonBlocking, hi));
      CK(cudaEventCreateWithFlags(&c->ev_flag_src, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(c->ev_flag_src, c->st));
    CK(cudaStreamWaitEvent(c->st_flag, c->ev_flag_src, 0));
    k_publish_flag<<<1, 1, 0, c->st_flag>>>(c->h_sflag_dev + slot, c->d_sflag + slot);
    CK(cudaGetLastError());
    c->ctr.kernel_launches++;
    CK(cudaEventRecord(c->ev_sflag[slot], c->st_flag));
    return;
  }
  else if (c->st_up) CK(cudaMemcpyAsync(c->h_sflag + slot, c->d_sflag + slot, 8, cudaMemcpyDeviceToHost, c->st));
  else LAUNCH(c, k_publish_flag, 1, 1, 0, c->h_sflag_dev + slot, c->d_sflag + slot);
  CK(cudaEventRecord(c->ev_sflag[slot], c->st));
}
static void trigger_standalone(edb_context* c, int slot, double thr) {
  if (!c->have_backup) throw StateError{EDB_ERR_STATE, "displ_over before update_particle_neighbors"};
  CK(cudaMemsetAsync(c->d_sflag + slot, 0, 8, c->st));
  if (c->poly()) LAUNCH(c, k_displ_over<true>, nblk(c->nreal), BLK, 0, view(c->pa), c->pr_backup.p, c->q_backup.p, c->nreal, c->S, thr * thr, c->d_sflag + slot, nullptr);
  else LAUNCH(c, k_displ_over<false>, nblk(c->nreal), BLK, 0, view(c->pa), c->pr_backup.p, nullptr, c->nreal, c->S, thr * thr, c->d_sflag + slot, nullptr);
}

// The Verlet loop of numerical_scheme_verlet_macro (config_numerical_schemes.msp:6-30).  Per step the device runs
//   [ghost refresh] -> contact waves -> gather fused with gravity, force_to_accel, the epilog AND the prolog + rebuild trigger of
//   the next step (k_gather<2>).
// With the displacement trigger (rebuild_every = 0) the host does not wait for the trigger before it enqueues the step: the
// kernels that change state are launched gated on the device-side flag, and the host reads the flag (already copied to pinned
// memory, one step behind the launch front) afterwards.  When the flag says "rebuild", the gated kernels have done nothing and
// the step is enqueued again behind the rebuild.  The GPU never idles on the host in the common no-rebuild case.

static void bond_force_launch(edb_context* c, double dt, const unsigned long long* gate = nullptr);
static uint64_t interface_criterion_launch(edb_context* c);
static void interface_criterion_async(edb_context* c, unsigned long long* flag);
static void bond_friction_updater_launch(edb_context* c);
// The same macro with moving drivers (edb_set_driver_motions): operator by operator in the reference's order
// (config_numerical_schemes.msp:12-29, check_and_update_particles of config_move_particles.msp:186-189), drivers integrated on the
// device.  One host read of the rebuild trigger per step (particle_displ_over | driver_displ_over, all-reduced).
static uint64_t run_verlet_moving(edb_context* c, double dt, const double g[3], double rcut_inc, int symetric, int nsteps, int rebuild_every) {
  uint64_t nb = 0;
  if (!c->have_system_mass) {
    if (c->cfg.nranks > 1) throw StateError{EDB_ERR_STATE, "moving drivers on several ranks: call edb_set_system_mass first"};
    const uint32_t n = c->nreal, nbk = nblk(n);
    c->stage.ensure((size_t)nbk * 7 + 8);
    std::vector<double> part((size_t)nbk * 7);
    if (nbk) {
      LAUNCH(c, k_sim_state_particles, nbk, BLK, 0, view(c->pa), n, c->stage.p);
      CK(cudaMemcpyAsync(part.data(), c->stage.p, part.size() * 8, cudaMemcpyDeviceToHost, c->st));
      CK(cudaStreamSynchronize(c->st));
    }
    double m = 0;
    for (uint32_t b = 0; b < nbk; b++) m += part[(size_t)b * 7 + 6];
    c->system_mass = m;
    c->have_system_mass = true;
  }
  if (!c->have_drv_backup) backup_drivers(c);
  const double thr = 0.5 * rcut_inc;
  for (int it = 0; it < nsteps; it++) {
    const double t = c->physical_time;
    prolog(c, dt);
    driver_op(c, DOP_PUSH_F_V_R, t, dt, 0.0);
    driver_op(c, DOP_PUSH_F_V, 0.0, 0.5 * dt, 0.0);
    driver_op(c, DOP_PUSH_AV_TO_QUAT, t, dt, 0.0);
    // fragmentation operators when bonds exist (config_fragmentation.msp:11-41): criterion on the previous step's energies, a broken
    // interface caps its friction and forces the rebuild (displ_or_break)
    const bool bonds = !c->h_bonds.empty();
    const uint64_t broken = bonds ? interface_criterion_launch(c) : 0;
    if (broken) bond_friction_updater_launch(c);
    bool rb;
    if (rebuild_every > 0) rb = ((it + 1) % rebuild_every) == 0;
    else if (rebuild_every < 0) rb = false;
    else {
      uint64_t over = displ_over(c, thr);          // all-reduced over the ranks
      if (!over) {
        CK(cudaMemsetAsync(c->d_flag, 0, 8, c->st));
        driver_displ_launch(c, thr, c->d_flag, nullptr);
        CK(cudaMemcpyAsync(c->h_flag, c->d_flag, 8, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        over = c->h_flag[0];
      }
      rb = over > 0;
    }
    rb = rb || broken > 0;
    if (rb) {
      backup_drivers(c);
      driver_vertices(c, false);       // ghost_full_update -> compute_all_vertices -> driver_vertices
      rebuild(c, rcut_inc, symetric);  // grid_rshape_driver inside (drivers_moving)
      nb++;
    } else {
      ghost_refresh(c, false);
      compute_vertices(c);
      driver_vertices(c, false);
    }
    driver_op(c, DOP_RESET_FORCE, 0.0, 0.0, 0.0);
    driver_gravity_launch(c, g);       // gravity_force also acts on the PARTICLE RShape drivers (gravity_force.cu:72-83)
    if (!bonds) {
      contact_waves(c, dt);
      gather<1>(c, g, dt);             // reset + gravity + contact wrenches + force_to_accel + epilog of the particles
    } else {                           // compute_force with inner_bond_polyhedron: the particle operators one by one
      LAUNCH(c, k_reset_fm, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal);
      LAUNCH(c, k_gravity, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, g[0], g[1], g[2]);
      bond_force_launch(c, dt);
      contact_waves(c, dt);
      gather<0>(c, nullptr, dt);
      LAUNCH(c, k_force_to_accel, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal);
    }
    driver_force_reduce(c);            // force_to_accel_driver: reaction sums, then Ball::f_ra / force_to_accel
    driver_op(c, DOP_FORCE_TO_ACCEL, 0.0, dt, c->system_mass);
    driver_op(c, DOP_PUSH_F_V, 0.0, 0.5 * dt, 0.0);
    if (bonds) LAUNCH(c, k_epilog<2>, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, dt * 0.5);
    c->ctr.particle_steps += c->nreal;
    c->physical_time += dt;
  }
  return nb;
}

// The macro of a fragmentation deck (config_fragmentation.msp:11-41 inside check_and_update_particles + compute_force with
// inner_bond_polyhedron): operator by operator, one host read per step (interfaces broken | particles moved).  A broken interface caps
// its friction and forces the neighbour rebuild (displ_or_break), where its bonds turn into vertex-vertex contacts.
static uint64_t run_verlet_bonds(edb_context* c, double dt, const double g[3], double rcut_inc, int symetric, int nsteps, int rebuild_every) {
  if (c->cfg.nranks > 1) throw StateError{EDB_ERR_STATE, "edb_run_verlet: inner bonds on several ranks are not supported yet"};
  uint64_t nb = 0;
  const double thr = 0.5 * rcut_inc;
  for (int it = 0; it < nsteps; it++) {
    prolog(c, dt);
    const uint64_t broken = interface_criterion_launch(c);
    if (broken) bond_friction_updater_launch(c);
    bool rb = broken > 0;
    if (rebuild_every > 0) rb = rb || ((it + 1) % rebuild_every) == 0;
    else if (rebuild_every == 0) rb = rb || displ_over(c, thr) > 0;
    if (rb) { rebuild(c, rcut_inc, symetric); nb++; }   // bonds follow their particles, broken ones retire (update_interfaces inside bonds_upload)
    else { ghost_refresh(c, false); compute_vertices(c); }
    LAUNCH(c, k_reset_fm, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal);
    LAUNCH(c, k_gravity, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, g[0], g[1], g[2]);
    bond_force_launch(c, dt);
    contact_waves(c, dt);
    gather<0>(c, nullptr, dt);
    LAUNCH(c, k_force_to_accel, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal);
    LAUNCH(c, k_epilog<2>, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, dt * 0.5);
    c->ctr.particle_steps += c->nreal;
    c->physical_time += dt;
  }
  return nb;
}

static void ensure_system_mass(edb_context* c) {
  if (c->have_system_mass) return;
  if (c->cfg.nranks > 1) throw StateError{EDB_ERR_STATE, "moving drivers on several ranks: call edb_set_system_mass first"};
  const uint32_t n = c->nreal, nbk = nblk(n);
  c->stage.ensure((size_t)nbk * 7 + 8);
  std::vector<double> part((size_t)nbk * 7);
  if (nbk) {
    LAUNCH(c, k_sim_state_particles, nbk, BLK, 0, view(c->pa), n, c->stage.p);
    CK(cudaMemcpyAsync(part.data(), c->stage.p, part.size() * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
  }
  double m = 0;
  for (uint32_t b = 0; b < nbk; b++) m += part[(size_t)b * 7 + 6];
  c->system_mass = m;
  c->have_system_mass = true;
}
// the fused driver operators of one step boundary (k_driver_cycle) + what follows them: mesh vertices and the drivers' share of the
// rebuild trigger of the next step
static void driver_cycle(edb_context* c, const double g[3], double dt, double time_next, bool do_end, bool do_begin, const unsigned long long* gate,
                         unsigned long long* next_flag, double thr) {
  const int nd = (int)c->h_drv.size();
  if (!nd) return;
  const bool sums = do_end && c->drivers_need_reaction;
  if (sums) {
    driver_force_sums(c);   // reads the contact state of the step; harmless on a gated-off step (the cycle kernel is gated)
    // force_to_accel_driver.cpp:128-139 (MPI_Allreduce of 6 doubles per driver): stream-ordered NCCL all-reduce, no host round trip.
    // Every rank enqueues the same sequence because the rebuild trigger that steers the loop is itself all-reduced.
    if (c->cfg.nranks > 1) c->comm.allreduce_sum_f64_dev(c->drv_sums.p, (size_t)nd * 6, c->st);
  }
  DriverCycleArgs A{c->drv.p, c->mo.p, nd, sums ? c->drv_sums.p : nullptr, c->poly() ? 0 : 1, do_end ? 1 : 0, do_begin ? 1 : 0,
                    dt, c->system_mass, time_next, gate, g ? mk3(g[0], g[1], g[2]) : mk3(0, 0, 0)};
  LAUNCH(c, k_driver_cycle, nblk(nd, 64), 64, 0, A);
  if (do_begin) {
    if (c->mesh_moving) driver_vertices(c, false);   // idempotent on a gated-off step (drivers unchanged)
    if (next_flag) driver_displ_launch(c, thr, next_flag, gate);
  }
}

extern "C" int edb_run_verlet(edb_context* c, double dt, const double g[3], double rcut_inc, int32_t symetric, int32_t nsteps,
                              int32_t rebuild_every, uint64_t* n_rebuilds) {
  API_BEGIN(c)
  uint64_t nb = 0;
  if (!c->have_prev) { rebuild(c, rcut_inc, symetric); }
  const bool bonds = !c->h_bonds.empty();
  if (bonds) {
    if (c->cfg.nranks > 1) throw StateError{EDB_ERR_STATE, "edb_run_verlet: inner bonds on several ranks are not supported yet"};
    // moving drivers, a fixed rebuild period, or EDB_BONDS_UNFUSED: the fragmentation macro operator by operator (one host read per step);
    // otherwise the speculative loop below carries the bond operators (the fracture criterion feeds the device-side rebuild trigger)
    if (c->drivers_moving || rebuild_every > 0 || c->env_bonds_unfused) {
      nb = c->drivers_moving ? run_verlet_moving(c, dt, g, rcut_inc, symetric, nsteps, rebuild_every)
                             : run_verlet_bonds(c, dt, g, rcut_inc, symetric, nsteps, rebuild_every);
      if (n_rebuilds) *n_rebuilds = nb;
      return EDB_OK;
    }
  }
  const bool mv = c->drivers_moving;
  if (mv && c->env_drivers_unfused) {
    nb = run_verlet_moving(c, dt, g, rcut_inc, symetric, nsteps, rebuild_every);
    if (n_rebuilds) *n_rebuilds = nb;
    return EDB_OK;
  }
  // with bonds a break must rebuild even when the caller asked for no displacement trigger (rebuild_every < 0): the trigger slot
  // then only counts broken interfaces (an infinite displacement threshold)
  const bool trig = rebuild_every == 0 || bonds;
  const double thr = rebuild_every == 0 ? 0.5 * rcut_inc : 1e300;
  const bool poly = c->poly();
  double tcur = c->physical_time;   // physical time of the step being enqueued (advanced by repeated addition like next_time_step)
  if (mv) {
    ensure_system_mass(c);
    if (!c->have_drv_backup) backup_drivers(c);
  }
  if (nsteps > 0) {
    prolog(c, dt);
    if (trig) trigger_standalone(c, 0, thr);
    if (mv) driver_cycle(c, g, dt, tcur, false, true, nullptr, trig ? c->d_sflag + 0 : nullptr, thr);
  }
  for (int it = 0; it < nsteps; it++) {
    const bool last = it + 1 == nsteps;
    const int slot = it & 1, nslot = slot ^ 1;
    auto step_kernels = [&](const unsigned long long* gate) {
      bool zeroed = false;
      if (bonds) {   // reset_force_moment, gravity_force, inner_bond_polyhedron into the force arrays; the gather starts from them
        LAUNCH(c, k_reset_gravity, nblk(c->nreal), BLK, 0, view(c->pa), c->nreal, g[0], g[1], g[2], gate);
        bond_force_launch(c, dt, gate);
      }
      contact_waves(c, dt, gate, (!last && trig) ? c->d_sflag + nslot : nullptr, &zeroed);
      if (last) {
        gather<1>(c, g, dt, gate, nullptr, 0.0, bonds);
        if (mv) driver_cycle(c, g, dt, tcur + dt, true, false, gate, nullptr, thr);
        return;
      }
      if (!trig) {
        gather<2>(c, g, dt, gate);
        if (mv) driver_cycle(c, g, dt, tcur + dt, true, true, gate, nullptr, thr);
        return;
      }
      if (!zeroed) CK(cudaMemsetAsync(c->d_sflag + nslot, 0, 8, c->st));
      gather<2>(c, g, dt, gate, c->d_sflag + nslot, thr, bonds);
      if (mv) driver_cycle(c, g, dt, tcur + dt, true, true, gate, c->d_sflag + nslot, thr);
    };
    if (!trig) {
      const bool rb = rebuild_every > 0 && ((it + 1) % rebuild_every) == 0;
      if (rb) { rebuild(c, rcut_inc, symetric); nb++; }
      else { ghost_refresh(c, false, nullptr, true); compute_vertices(c, it > 0); }
      step_kernels(nullptr);
    } else {
      const edb_counters snap = c->ctr;
      if (bonds) {   // apply_interface_fracture_criterion + broken_interface_friction_updater of this step: a break is a rebuild trigger
        interface_criterion_async(c, c->d_sflag + slot);
        bond_friction_updater_launch(c);
      }
      ghost_refresh(c, false, c->d_sflag + slot, true);
      trigger_publish(c, slot);
      compute_vertices(c, it > 0);
      step_kernels(c->d_sflag + slot);
      CK(cudaEventSynchronize(c->ev_sflag[slot]));
      if (c->h_sflag[slot] > 0) {
        const uint64_t launches = c->ctr.kernel_launches;
        c->ctr = snap;   // the gated launches did nothing
        c->ctr.kernel_launches = launches;
        rebuild(c, rcut_inc, symetric);   // with moving drivers: backup_drivers + grid_rshape_driver inside
        nb++;
        c->queue_clean = false;   // the gated-off gather did not clear the chunk queue
        step_kernels(nullptr);
      }
    }
    c->ctr.particle_steps += c->nreal;
    tcur += dt;
  }
  ghosts_home(c);   // the ghost slots are current again for whoever reads them next (operators, views, the next rebuild)
  c->physical_time = tcur;
  if (n_rebuilds) *n_rebuilds = nb;
  API_END(c)
}

// ---- views ---------------------------------------------------------------------------------------------------------
__global__ void k_count_active(WView w, unsigned long long* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int a = 0;
  if (k < w.n) {
    const uint8_t fl = w.flg[k];
    if (fl & FLG_HOT) {
      const double4 f = w.ft[k], m = (fl & FLG_MOM) ? w.mo[k] : make_double4(0, 0, 0, 0);
      a = (f.x != 0.0 || f.y != 0.0 || f.z != 0.0 || m.x != 0.0 || m.y != 0.0 || m.z != 0.0) ? 1 : 0;
    }
  }
  a = __reduce_add_sync(0xffffffffu, a);
  if ((threadIdx.x & 31) == 0 && a) atomicAdd(out, (unsigned long long)a);
}
extern "C" int edb_interaction_counts(edb_context* c, uint64_t counts[EDB_NTYPES], uint64_t active[EDB_NTYPES]) {
  API_BEGIN(c)
  for (int t = 0; t < NT; t++) {
    counts[t] = c->wave[t].n;
    if (active) {
      active[t] = 0;
      if (c->wave[t].n) {
        CK(cudaMemsetAsync(c->d_flag, 0, 8, c->st));
        LAUNCH(c, k_count_active, nblk(c->wave[t].n), BLK, 0, view(c->wave[t]), c->d_flag);
        CK(cudaMemcpyAsync(c->h_flag, c->d_flag, 8, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        active[t] = c->h_flag[0];
      }
    }
  }
  API_END(c)
}

__global__ void k_export(WView w, const uint64_t* __restrict__ id, int type, bool has_sub, bool driver, edb_interaction* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= w.n) return;
  const uint2 ij = w.ij[k];
  const uint8_t fl = w.flg[k];
  const double4 ft = (fl & FLG_HOT) ? w.ft[k] : make_double4(0, 0, 0, 0);
  const double4 mo = (fl & FLG_MOM) ? w.mo[k] : make_double4(0, 0, 0, 0);
  edb_interaction e;
  memset(&e, 0, sizeof(e));
  e.id_i = id[ij.x];
  e.id_j = driver ? (uint64_t)ij.y : id[ij.y];
  if (has_sub) { const uint2 s = w.sub[k]; e.sub_i = s.x; e.sub_j = s.y; }
  e.type = (uint16_t)type; e.swap = fl & 1; e.ghost = (fl >> 1) & 3;
  e.friction[0] = ft.x; e.friction[1] = ft.y; e.friction[2] = ft.z;
  e.moment[0] = mo.x; e.moment[1] = mo.y; e.moment[2] = mo.z;
  out[k] = e;
}
extern "C" int edb_export_interactions(edb_context* c, edb_interaction* out, uint64_t cap, uint64_t* n) {
  API_BEGIN(c)
  uint64_t tot = 0;
  for (int t = 0; t < NT; t++) tot += c->wave[t].n;
  *n = tot;
  if (tot > cap) throw StateError{EDB_ERR_CAPACITY, "export buffer too small"};
  edb_interaction* d = nullptr;
  if (tot) {
    CK(cudaMalloc(&d, tot * sizeof(edb_interaction)));
    uint64_t o = 0;
    for (int t = 0; t < NT; t++) {
      if (!c->wave[t].n) continue;
      LAUNCH(c, k_export, nblk(c->wave[t].n), BLK, 0, view(c->wave[t]), c->pa.id.p, t, c->poly() || t >= 7, t >= 4, d + o);
      o += c->wave[t].n;
    }
    CK(cudaMemcpyAsync(out, d, tot * sizeof(edb_interaction), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaFree(d));
  }
  API_END(c)
}

extern "C" int edb_get_contact_state(edb_context* c, int32_t type, double* dn, double* cp, double* fn, double* ft) {
  API_BEGIN(c)
  if (type < 0 || type >= NT) throw StateError{EDB_ERR_INVALID, "bad type"};
  const uint64_t n = c->wave[type].n;
  if (n) {
    c->stage.ensure(10 * n + 1);
    double* s = c->stage.p;
    ContactArgs A{view(c->pa), view(c->wave[type]), c->prm, c->drv.p, c->vert.p, c->S, 0.0, type, c->symetric, c->M};
    if (c->poly()) LAUNCH(c, k_state<true>, nblk(n), BLK, 0, A, s, s + n, s + 4 * n, s + 7 * n);
    else LAUNCH(c, k_state<false>, nblk(n), BLK, 0, A, s, s + n, s + 4 * n, s + 7 * n);
    if (dn) CK(cudaMemcpyAsync(dn, s, n * 8, cudaMemcpyDeviceToHost, c->st));
    if (cp) CK(cudaMemcpyAsync(cp, s + n, 3 * n * 8, cudaMemcpyDeviceToHost, c->st));
    if (fn) CK(cudaMemcpyAsync(fn, s + 4 * n, 3 * n * 8, cudaMemcpyDeviceToHost, c->st));
    if (ft) CK(cudaMemcpyAsync(ft, s + 7 * n, 3 * n * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
  }
  API_END(c)
}

extern "C" int edb_import_interactions(edb_context* c, const edb_interaction* in, uint64_t n) {
  API_BEGIN(c)
  // Restart path: the records become the history source of the NEXT rebuild.  Owner slots are resolved on the host
  // through the id -> slot map of the current device order (rebuild-frequency operation, not on the per-step path).
  const uint32_t np = c->nreal;
  std::vector<uint64_t> ids(np);
  CK(cudaMemcpy(ids.data(), c->pa.id.p, (size_t)np * 8, cudaMemcpyDeviceToHost));
  std::unordered_map<uint64_t, uint32_t> slot;
  slot.reserve(np * 2);
  for (uint32_t p = 0; p < np; p++) slot[ids[p]] = p;
  const bool poly = c->poly();
  std::vector<std::vector<uint32_t>> per_type(NT);
  for (uint64_t k = 0; k < n; k++)
    if (in[k].type < NT) per_type[in[k].type].push_back((uint32_t)k);
  for (int t = 0; t < NT; t++) {
    Wave& w = c->wave[t];
    auto& idx = per_type[t];
    // owner slot per record; records whose owner is unknown are dropped
    std::vector<std::pair<uint32_t, uint32_t>> key;  // (owner slot, record)
    for (uint32_t r : idx) {
      const auto& e = in[r];
      const uint64_t oid = e.swap ? e.id_j : e.id_i;
      auto it = slot.find(oid);
      if (it != slot.end()) key.push_back({it->second, r});
    }
    std::stable_sort(key.begin(), key.end(), [](auto& a, auto& b) { return a.first < b.first; });
    const size_t m = key.size();
    w.n = m;
    w.off.ensure(np + 2);
    w.ensure(m + 1, poly || t >= 7);
    std::vector<uint32_t> off(np + 1, 0);
    std::vector<uint2> ij(m), sub(m);
    std::vector<uint8_t> flg(m);
    std::vector<double4> ft(m), mo(m);
    for (size_t s = 0; s < m; s++) {
      const auto& e = in[key[s].second];
      off[key[s].first + 1]++;
      const uint64_t pid = (t >= 4) ? e.id_j : (e.swap ? e.id_i : e.id_j);
      uint32_t si = 0, sj = 0;
      auto a = slot.find(e.id_i), b = slot.find(e.id_j);
      if (a != slot.end()) si = a->second;
      if (t >= 4) sj = (uint32_t)e.id_j; else if (b != slot.end()) sj = b->second;
      ij[s] = make_uint2(si, sj);
      sub[s] = make_uint2(e.sub_i, e.sub_j);
      const bool live = e.friction[0] != 0.0 || e.friction[1] != 0.0 || e.friction[2] != 0.0 || e.moment[0] != 0.0 || e.moment[1] != 0.0 ||
                        e.moment[2] != 0.0;
      const bool mlive = e.moment[0] != 0.0 || e.moment[1] != 0.0 || e.moment[2] != 0.0;
      flg[s] = (uint8_t)((e.swap & 1) | ((e.ghost & 3) << 1) | (live ? FLG_HOT : 0) | (mlive ? FLG_MOM : 0));
      long long pb = (long long)pid;
      double pidd;
      memcpy(&pidd, &pb, 8);
      ft[s] = make_double4(e.friction[0], e.friction[1], e.friction[2], pidd);
      mo[s] = make_double4(e.moment[0], e.moment[1], e.moment[2], 0.0);
    }
    for (uint32_t p = 0; p < np; p++) off[p + 1] += off[p];
    CK(cudaMemcpy(w.off.p, off.data(), (np + 1) * 4, cudaMemcpyHostToDevice));
    if (m) {
      CK(cudaMemcpy(w.ij.p, ij.data(), m * sizeof(uint2), cudaMemcpyHostToDevice));
      if (poly || t >= 7) CK(cudaMemcpy(w.sub.p, sub.data(), m * sizeof(uint2), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(w.flg.p, flg.data(), m, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(w.ft.p, ft.data(), m * sizeof(double4), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(w.mo.p, mo.data(), m * sizeof(double4), cudaMemcpyHostToDevice));
      CK(cudaMemset(w.fc.p, 0, 2 * m * sizeof(double4)));
    }
  }
  c->have_prev = true;   // valid history source; contact operators still require a rebuild (have_inc = false)
  c->n_old_real = np;
  c->have_inc = false;
  // identity permutation until the next rebuild re-bins
  API_END(c)
}

static void driver_force_sums(edb_context* c);
extern "C" int edb_compute_driver_force_moment(edb_context* c, double* forces, double* moments) {
  API_BEGIN(c)
  const int nd = (int)c->h_drv.size();
  if (nd == 0) return EDB_OK;
  driver_force_sums(c);   // per wave: block partials (k_driver_force) summed in block order on the device (k_driver_force_sum)
  std::vector<double> acc((size_t)nd * 6, 0.0);
  CK(cudaMemcpyAsync(acc.data(), c->drv_sums.p, acc.size() * 8, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  for (int d = 0; d < nd; d++)
    for (int k = 0; k < 3; k++) {
      forces[3 * d + k] = acc[d * 6 + k];
      moments[3 * d + k] = acc[d * 6 + 3 + k];
    }
  API_END(c)
}

// ---- driver operators of the time loop (config_numerical_schemes.msp:13-24), device-resident drivers ----------------------------
static void driver_op(edb_context* c, int op, double time, double dt, double system_mass) {
  const int nd = (int)c->h_drv.size();
  if (nd) LAUNCH(c, k_driver_op, nblk(nd, 64), 64, 0, c->drv.p, c->mo.p, nd, op, time, dt, system_mass);
}
static void driver_gravity_launch(edb_context* c, const double g[3]) {
  const int nd = (int)c->h_drv.size();
  if (nd && c->drivers_moving) LAUNCH(c, k_driver_gravity, nblk(nd, 64), 64, 0, c->drv.p, c->mo.p, nd, mk3(g[0], g[1], g[2]));
}
// reaction of the driver waves of the last contact operator -> DriverMotion::forces / mom (deterministic block-ordered sums)
// block partials of the driver waves summed into drv_sums[nd * 6] in a fixed order (this rank's share)
static void driver_force_sums(edb_context* c) {
  const int nd = (int)c->h_drv.size();
  if (!nd) return;
  c->drv_sums.ensure((size_t)nd * 6 + 1);
  CK(cudaMemsetAsync(c->drv_sums.p, 0, (size_t)nd * 6 * sizeof(double), c->st));
  for (int t = 4; t <= 12; t++) {
    Wave& w = c->wave[t];
    if (!w.n) continue;
    const uint32_t nb = nblk(w.n);
    c->stage.ensure((size_t)nb * nd * 6 + 1);
    const size_t smem = (size_t)(BLK / 32) * nd * 6 * sizeof(double);
    LAUNCH(c, k_driver_force, nb, BLK, smem, view(w), t, c->drv.p, nd, c->stage.p);
    LAUNCH(c, k_driver_force_sum, nblk((uint64_t)nd * 6, 64), 64, 0, c->stage.p, nb, nd, c->drv_sums.p);
  }
}
static void driver_force_reduce(edb_context* c) {
  const int nd = (int)c->h_drv.size();
  if (!nd) return;
  driver_force_sums(c);
  if (c->cfg.nranks > 1) c->comm.allreduce_sum_f64_dev(c->drv_sums.p, (size_t)nd * 6, c->st);
  LAUNCH(c, k_driver_force_apply, nblk(nd, 64), 64, 0, c->drv_sums.p, nd, c->drv.p, c->mo.p, c->poly() ? 0 : 1);
}
static void driver_vertices(edb_context* c, bool all) {
  const int nm = (int)c->h_md.size();
  if (!nm) return;
  int nvmax = 1;
  for (auto& m : c->h_md) nvmax = std::max(nvmax, m.nv);
  dim3 grid(std::min<uint32_t>(nblk(nvmax), 1024u), nm);
  k_driver_vertices<<<grid, BLK, 0, c->st>>>(c->d_md.p, nm, c->drv.p, c->mo.p, c->d_mlocal.p, c->d_mverts.p, all ? 1 : 0);
  c->ctr.kernel_launches++;
}
static void backup_drivers(edb_context* c) {
  const int nd = (int)c->h_drv.size();
  c->drv_backup.ensure(nd + 1); c->mo_backup.ensure(nd + 1);
  if (nd) {
    CK(cudaMemcpyAsync(c->drv_backup.p, c->drv.p, nd * sizeof(Driver), cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(c->mo_backup.p, c->mo.p, nd * sizeof(DriverMotion), cudaMemcpyDeviceToDevice, c->st));
  }
  c->have_drv_backup = true;
}
// adds the count to *d_out (device); the caller zeroes / reads it
static void driver_displ_launch(edb_context* c, double thr, unsigned long long* d_out, const unsigned long long* gate = nullptr) {
  const int nd = (int)c->h_drv.size();
  if (!nd) return;
  if (!c->have_drv_backup) backup_drivers(c);
  int nvmax = 1;
  for (auto& m : c->h_md) nvmax = std::max(nvmax, m.nv);
  dim3 grid(std::min<uint32_t>(nblk(nvmax), 1024u), nd);
  k_driver_displ<<<grid, BLK, 0, c->st>>>(c->drv.p, c->mo.p, c->drv_backup.p, c->mo_backup.p, nd, c->d_md.p, c->d_mlocal.p, thr * thr, d_out, gate);
  c->ctr.kernel_launches++;
}
extern "C" int edb_push_f_v_r_driver(edb_context* c, double time, double dt) {
  API_BEGIN(c)
  driver_op(c, DOP_PUSH_F_V_R, time, dt, 0.0);
  API_END(c)
}
extern "C" int edb_push_f_v_driver(edb_context* c, double dt) {
  API_BEGIN(c)
  driver_op(c, DOP_PUSH_F_V, 0.0, dt, 0.0);
  API_END(c)
}
extern "C" int edb_push_av_to_quat_driver(edb_context* c, double time, double dt) {
  API_BEGIN(c)
  driver_op(c, DOP_PUSH_AV_TO_QUAT, time, dt, 0.0);
  API_END(c)
}
extern "C" int edb_reset_force_driver(edb_context* c) {
  API_BEGIN(c)
  driver_op(c, DOP_RESET_FORCE, 0.0, 0.0, 0.0);
  API_END(c)
}
extern "C" int edb_force_to_accel_driver(edb_context* c, double dt, double system_mass) {
  API_BEGIN(c)
  driver_force_reduce(c);
  driver_op(c, DOP_FORCE_TO_ACCEL, 0.0, dt, system_mass);
  API_END(c)
}
extern "C" int edb_driver_vertices(edb_context* c) {
  API_BEGIN(c)
  driver_vertices(c, false);
  API_END(c)
}
extern "C" int edb_grid_rshape_driver(edb_context* c) {
  API_BEGIN(c)
  build_mesh_grid(c);
  API_END(c)
}
extern "C" int edb_backup_drivers(edb_context* c) {
  API_BEGIN(c)
  backup_drivers(c);
  API_END(c)
}
extern "C" int edb_driver_displ_over(edb_context* c, double thr, uint64_t* result) {
  API_BEGIN(c)
  CK(cudaMemsetAsync(c->d_flag, 0, 8, c->st));
  driver_displ_launch(c, thr, c->d_flag);
  CK(cudaMemcpyAsync(c->h_flag, c->d_flag, 8, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  // every rank holds the same drivers: no reduction needed (the reference's MPI_Allreduce sums identical per-rank flags, :238)
  *result = c->h_flag[0];
  API_END(c)
}
extern "C" int edb_set_physical_time(edb_context* c, double t) {
  API_BEGIN(c)
  c->physical_time = t;
  API_END(c)
}
extern "C" int edb_get_physical_time(edb_context* c, double* t) {
  API_BEGIN(c)
  *t = c->physical_time;
  API_END(c)
}
extern "C" int edb_set_system_mass(edb_context* c, double m) {
  API_BEGIN(c)
  c->system_mass = m;
  c->have_system_mass = true;
  API_END(c)
}

// ---- inner bonds (per-step operators) --------------------------------------------------------------------------
static BondView bond_view(edb_context* c) {
  return BondView{(uint32_t)c->h_bonds.size(), c->b_ij.p, c->b_sub.p, c->b_par.p, c->b_mode.p, c->b_unbroken.p, c->b_fr.p, c->b_tds.p, c->b_fd.p, c->b_cp.p};
}
// device -> host mirror of the fields the kernels change
static void bonds_download(edb_context* c) {
  const size_t n = c->h_bonds.size();
  if (!c->bonds_on_device || n == 0) return;
  std::vector<double4> fr(n), td(n);
  std::vector<uint8_t> ub(n);
  CK(cudaStreamSynchronize(c->st));
  CK(cudaMemcpy(fr.data(), c->b_fr.p, n * sizeof(double4), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(td.data(), c->b_tds.p, n * sizeof(double4), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ub.data(), c->b_unbroken.p, n, cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < n; k++) {
    edb_inner_bond& b = c->h_bonds[k];
    b.friction[0] = fr[k].x; b.friction[1] = fr[k].y; b.friction[2] = fr[k].z; b.en = fr[k].w;
    b.tds[0] = td[k].x; b.tds[1] = td[k].y; b.tds[2] = td[k].z; b.et = td[k].w;
    b.unbroken = ub[k];
  }
}
// Resolves the particle slots by id in the current device order (rebuild-frequency, like update_persistent_interactions), sorts the
// bonds by (owner slot, partner id, sub_i, sub_j), uploads the SoA, the per-particle incidence lists and the interfaces.
static void bonds_upload(edb_context* c) {
  const size_t n = c->h_bonds.size();
  c->n_bond_particles = 0; c->n_interfaces = 0;
  c->bonds_on_device = false; c->have_ignore = false;
  if (n == 0) return;
  const uint32_t np = c->nreal;
  std::vector<uint64_t> ids(np);
  CK(cudaStreamSynchronize(c->st));
  if (np) CK(cudaMemcpy(ids.data(), c->pa.id.p, (size_t)np * 8, cudaMemcpyDeviceToHost));
  std::unordered_map<uint64_t, uint32_t> slot;
  slot.reserve((size_t)np * 2);
  for (uint32_t p = 0; p < np; p++) slot[ids[p]] = p;
  std::vector<uint2> ij(n);
  std::vector<uint32_t> order(n);
  auto find = [&](uint64_t id) { auto it = slot.find(id); return it == slot.end() ? 0xffffffffu : it->second; };
  std::vector<uint32_t> si(n), sj(n);
  for (size_t k = 0; k < n; k++) { si[k] = find(c->h_bonds[k].id_i); sj[k] = find(c->h_bonds[k].id_j); order[k] = (uint32_t)k; }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    const edb_inner_bond &x = c->h_bonds[a], &y = c->h_bonds[b];
    return std::tie(si[a], x.id_j, x.sub_i, x.sub_j) < std::tie(si[b], y.id_j, y.sub_i, y.sub_j);
  });
  std::vector<edb_inner_bond> sorted(n);
  for (size_t k = 0; k < n; k++) { sorted[k] = c->h_bonds[order[k]]; ij[k] = make_uint2(si[order[k]], sj[order[k]]); }
  c->h_bonds.swap(sorted);
  std::vector<uint2> sub(n);
  std::vector<double4> par(n), fr(n), td(n);
  std::vector<uint8_t> mode(n), ub(n);
  for (size_t k = 0; k < n; k++) {
    const edb_inner_bond& b = c->h_bonds[k];
    sub[k] = make_uint2(b.sub_i, b.sub_j);
    par[k] = make_double4(b.dn0, b.weight, b.criteria_1, b.criteria_2);
    fr[k] = make_double4(b.friction[0], b.friction[1], b.friction[2], b.en);
    td[k] = make_double4(b.tds[0], b.tds[1], b.tds[2], b.et);
    mode[k] = (uint8_t)b.mode; ub[k] = b.unbroken;
  }
  c->b_ij.ensure(n); c->b_sub.ensure(n); c->b_par.ensure(n); c->b_fr.ensure(n); c->b_tds.ensure(n); c->b_fd.ensure(n); c->b_cp.ensure(n);
  c->b_mode.ensure(n); c->b_unbroken.ensure(n); c->b_ids.ensure(n); c->b_ghost.ensure(n);
  {
    std::vector<ulonglong2> ids2(n);
    std::vector<uint8_t> gh(n);
    for (size_t k = 0; k < n; k++) { ids2[k] = make_ulonglong2(c->h_bonds[k].id_i, c->h_bonds[k].id_j); gh[k] = (uint8_t)c->h_bonds[k].ghost; }
    CK(cudaMemcpy(c->b_ids.p, ids2.data(), n * sizeof(ulonglong2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->b_ghost.p, gh.data(), n, cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(c->b_ij.p, ij.data(), n * sizeof(uint2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->b_sub.p, sub.data(), n * sizeof(uint2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->b_par.p, par.data(), n * sizeof(double4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->b_fr.p, fr.data(), n * sizeof(double4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->b_tds.p, td.data(), n * sizeof(double4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->b_mode.p, mode.data(), n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->b_unbroken.p, ub.data(), n, cudaMemcpyHostToDevice));
  CK(cudaMemset(c->b_fd.p, 0, n * sizeof(double4)));
  CK(cudaMemset(c->b_cp.p, 0, n * sizeof(double4)));
  // ignore lists of the neighbour build (InteractionManager::update_ignore_interaction): unique partner ids of the unbroken bonds per owner
  {
    std::vector<uint32_t> ioff((size_t)np + 1, 0);
    std::vector<unsigned long long> iids;
    size_t k = 0;
    for (uint32_t p = 0; p < np; p++) {
      ioff[p] = (uint32_t)iids.size();
      const size_t first = iids.size();
      for (; k < n && ij[k].x == p; k++) {
        if (!c->h_bonds[k].unbroken || ij[k].y == 0xffffffffu) continue;
        bool dup = false;
        for (size_t q = first; q < iids.size(); q++) dup = dup || iids[q] == c->h_bonds[k].id_j;
        if (!dup) iids.push_back(c->h_bonds[k].id_j);
      }
    }
    ioff[np] = (uint32_t)iids.size();
    c->b_ign_off.ensure(ioff.size() + 1); c->b_ign_ids.ensure(iids.size() + 1);
    CK(cudaMemcpy(c->b_ign_off.p, ioff.data(), ioff.size() * 4, cudaMemcpyHostToDevice));
    if (!iids.empty()) CK(cudaMemcpy(c->b_ign_ids.p, iids.data(), iids.size() * 8, cudaMemcpyHostToDevice));
    c->have_ignore = !iids.empty();
  }
  // incidence lists: for every particle that carries bonds, its bonds in bond order (side 0 = i, 1 = j)
  std::vector<std::pair<uint32_t, uint32_t>> inc;   // (particle, code)
  for (size_t k = 0; k < n; k++) {
    if (ij[k].x == 0xffffffffu || ij[k].y == 0xffffffffu) continue;
    inc.push_back({ij[k].x, (uint32_t)(k << 1)});
    inc.push_back({ij[k].y, (uint32_t)(k << 1) | 1u});
  }
  std::stable_sort(inc.begin(), inc.end(), [](auto& a, auto& b) { return a.first < b.first; });
  std::vector<uint32_t> plist, off, codes;
  for (size_t q = 0; q < inc.size(); q++) {
    if (q == 0 || inc[q].first != inc[q - 1].first) { plist.push_back(inc[q].first); off.push_back((uint32_t)q); }
    codes.push_back(inc[q].second);
  }
  off.push_back((uint32_t)inc.size());
  c->n_bond_particles = (uint32_t)plist.size();
  c->b_plist.ensure(plist.size() + 1); c->b_inc_off.ensure(off.size() + 1); c->b_inc.ensure(codes.size() + 1);
  if (!plist.empty()) {
    CK(cudaMemcpy(c->b_plist.p, plist.data(), plist.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->b_inc_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->b_inc.p, codes.data(), codes.size() * 4, cudaMemcpyHostToDevice));
  }
  // update_interfaces.cpp:44-84: maximal runs of bonds between the same two particles, PartnerGhost records skipped
  std::vector<uint2> itf;
  size_t loc = 0;
  while (loc < n) {
    if (c->h_bonds[loc].ghost == 2) { loc++; continue; }
    size_t m = 1;
    while (loc + m < n && c->h_bonds[loc + m].id_i == c->h_bonds[loc].id_i && c->h_bonds[loc + m].id_j == c->h_bonds[loc].id_j) m++;
    itf.push_back(make_uint2((uint32_t)loc, (uint32_t)m));
    loc += m;
  }
  c->n_interfaces = (uint32_t)itf.size();
  c->b_itf.ensure(itf.size() + 1); c->b_itf_broken.ensure(itf.size() + 1);
  if (!itf.empty()) {
    CK(cudaMemcpy(c->b_itf.p, itf.data(), itf.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CK(cudaMemset(c->b_itf_broken.p, 0, itf.size()));
  }
  c->bonds_on_device = true;
}
// The device-side form of bonds_download + bonds_retire_broken + bonds_upload (kernels_bond_rebuild.cuh): called by rebuild() once the
// re-binning permutation old2new is known.  Two host reads of a few counters; the host mirror follows through the permutation of the
// surviving bonds (its dynamic fields stay stale until bonds_download).
static void bonds_rebuild_device(edb_context* c) {
  const uint32_t n = (uint32_t)c->h_bonds.size();
  const uint32_t np = c->nreal;
  c->n_conv_dev = 0;
  if (n == 0) return;
  auto arrays = [&](bool second) {
    return second ? BondArrays{c->b2_ij.p, c->b2_sub.p, c->b2_ids.p, c->b2_par.p, c->b2_fr.p, c->b2_tds.p, c->b2_fd.p, c->b2_cp.p, c->b2_mode.p, c->b2_unbroken.p, c->b2_ghost.p}
                  : BondArrays{c->b_ij.p, c->b_sub.p, c->b_ids.p, c->b_par.p, c->b_fr.p, c->b_tds.p, c->b_fd.p, c->b_cp.p, c->b_mode.p, c->b_unbroken.p, c->b_ghost.p};
  };
  c->b2_ij.ensure(n); c->b2_sub.ensure(n); c->b2_ids.ensure(n); c->b2_par.ensure(n); c->b2_fr.ensure(n); c->b2_tds.ensure(n); c->b2_fd.ensure(n);
  c->b2_cp.ensure(n); c->b2_mode.ensure(n); c->b2_unbroken.ensure(n); c->b2_ghost.ensure(n);
  c->b_key.ensure(n + 1); c->b_key2.ensure(n + 1); c->b_idx.ensure(n + 1); c->b_idx2.ensure(n + 1); c->b_cnt.ensure(8);
  c->d_conv.ensure(n);
  CK(cudaMemsetAsync(c->b_cnt.p, 0, 32, c->st));
  LAUNCH(c, k_bond_rekey, nblk(n), BLK, 0, arrays(false), n, c->old2new.p, c->b_key.p, c->b_idx.p, c->b_cnt.p, c->d_conv.p);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, c->b_key.p, c->b_key2.p, c->b_idx.p, c->b_idx2.p, (int)n, 0, 32, c->st);
  size_t tb2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, c->b_key.p, c->b_key2.p, c->b_idx.p, c->b_idx2.p, (int)(2 * n), 0, 32, c->st);
  c->cub_tmp.ensure(std::max(tb, tb2) + 16);
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, c->b_key.p, c->b_key2.p, c->b_idx.p, c->b_idx2.p, (int)n, 0, 32, c->st));
  c->ctr.kernel_launches += 4;
  uint32_t h_cnt[8] = {};
  CK(cudaMemcpyAsync(h_cnt, c->b_cnt.p, 32, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  const uint32_t n_keep = n - h_cnt[0];
  c->n_conv_dev = h_cnt[1];
  // host mirror: the surviving bonds in their new order (static fields; the dynamic ones are read back on demand)
  {
    std::vector<uint32_t> perm(n_keep);
    if (n_keep) CK(cudaMemcpy(perm.data(), c->b_idx2.p, (size_t)n_keep * 4, cudaMemcpyDeviceToHost));
    std::vector<edb_inner_bond> nh(n_keep);
    for (uint32_t k = 0; k < n_keep; k++) { nh[k] = c->h_bonds[perm[k]]; nh[k].unbroken = 1; }
    c->h_bonds.swap(nh);
  }
  c->n_bond_particles = 0; c->n_interfaces = 0; c->have_ignore = false;
  if (n_keep == 0) { c->bonds_on_device = false; return; }
  LAUNCH(c, k_bond_permute, nblk(n_keep), BLK, 0, arrays(true), arrays(false), n_keep, c->b_key2.p, c->b_idx2.p, c->old2new.p);
  std::swap(c->b_ij, c->b2_ij); std::swap(c->b_sub, c->b2_sub); std::swap(c->b_ids, c->b2_ids); std::swap(c->b_par, c->b2_par);
  std::swap(c->b_fr, c->b2_fr); std::swap(c->b_tds, c->b2_tds); std::swap(c->b_fd, c->b2_fd); std::swap(c->b_cp, c->b2_cp);
  std::swap(c->b_mode, c->b2_mode); std::swap(c->b_unbroken, c->b2_unbroken); std::swap(c->b_ghost, c->b2_ghost);
  // derived tables
  const uint32_t m = n_keep;
  c->b_fl_ign.ensure(m + 2); c->b_rk_ign.ensure(m + 2); c->b_fl_itf.ensure(m + 2); c->b_rk_itf.ensure(m + 2);
  c->b_ikey.ensure(2 * (size_t)m + 2); c->b_ikey2.ensure(2 * (size_t)m + 2); c->b_ival.ensure(2 * (size_t)m + 2); c->b_inc.ensure(2 * (size_t)m + 2);
  c->b_ist.ensure(2 * (size_t)m + 2); c->b_irk.ensure(2 * (size_t)m + 3);
  c->b_ign_off.ensure((size_t)np + 2); c->b_ign_ids.ensure((size_t)m + 1);
  c->b_itf.ensure((size_t)m + 1); c->b_itf_broken.ensure((size_t)m + 1);
  c->b_plist.ensure(2 * (size_t)m + 2); c->b_inc_off.ensure(2 * (size_t)m + 3);
  c->tmp.ensure((size_t)np + 2);
  CK(cudaMemsetAsync(c->tmp.p, 0, ((size_t)np + 1) * 4, c->st));            // per-owner count of ignore entries
  CK(cudaMemsetAsync(c->b_cnt.p, 0, 32, c->st));
  LAUNCH(c, k_bond_flags, nblk(m), BLK, 0, arrays(false), m, c->b_fl_ign.p, c->tmp.p, c->b_fl_itf.p, c->b_ikey.p, c->b_ival.p, c->b_cnt.p);
  scan_u32(c, c->tmp.p, c->b_ign_off.p, np);
  scan_u32(c, c->b_fl_ign.p, c->b_rk_ign.p, m);
  scan_u32(c, c->b_fl_itf.p, c->b_rk_itf.p, m);
  LAUNCH(c, k_bond_scatter, nblk(m), BLK, 0, arrays(false), m, c->b_fl_ign.p, c->b_rk_ign.p, c->b_ign_ids.p, c->b_fl_itf.p, c->b_rk_itf.p, c->b_itf.p);
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb2, c->b_ikey.p, c->b_ikey2.p, c->b_ival.p, c->b_inc.p, (int)(2 * m), 0, 32, c->st));
  c->ctr.kernel_launches += 4;
  CK(cudaMemcpyAsync(h_cnt, c->b_cnt.p, 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaMemcpyAsync(h_cnt + 1, c->b_rk_ign.p + m, 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaMemcpyAsync(h_cnt + 2, c->b_rk_itf.p + m, 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  const uint32_t n_valid = h_cnt[0], n_ign = h_cnt[1], n_itf = h_cnt[2];
  c->have_ignore = n_ign > 0;
  c->n_interfaces = n_itf;
  if (n_itf) {
    LAUNCH(c, k_itf_lengths, nblk(n_itf), BLK, 0, c->b_itf.p, n_itf, m);
    CK(cudaMemsetAsync(c->b_itf_broken.p, 0, n_itf, c->st));
  }
  if (n_valid) {
    LAUNCH(c, k_bond_inc_flags, nblk(n_valid), BLK, 0, c->b_ikey2.p, n_valid, c->b_ist.p);
    scan_u32(c, c->b_ist.p, c->b_irk.p, n_valid);
    LAUNCH(c, k_bond_inc_lists, nblk(n_valid + 1), BLK, 0, c->b_ikey2.p, c->b_ist.p, c->b_irk.p, n_valid, c->b_plist.p, c->b_inc_off.p);
    CK(cudaMemcpyAsync(h_cnt, c->b_irk.p + n_valid, 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    c->n_bond_particles = h_cnt[0];
  }
  c->bonds_on_device = true;
}
// At a rebuild: broken bonds are no longer persistent (inner_bond_interaction.hpp:125).  They leave the bond list; the ones whose
// friction is non-zero (Interaction::active, interaction.hpp:129-136) are queued to become the history of the vertex-vertex record of
// the same two vertices in the new wave 0 (unclassify_interactions.cpp:135-145, broke_interaction).  Needs the particles' new order.
static void bonds_retire_broken(edb_context* c) {
  c->h_conv.clear();
  bool any = false;
  for (auto& b : c->h_bonds) any = any || !b.unbroken;
  if (!any) return;
  const uint32_t np = c->nreal;
  std::vector<uint64_t> ids(np);
  CK(cudaStreamSynchronize(c->st));
  if (np) CK(cudaMemcpy(ids.data(), c->pa.id.p, (size_t)np * 8, cudaMemcpyDeviceToHost));
  std::unordered_map<uint64_t, uint32_t> slot;
  slot.reserve((size_t)np * 2);
  for (uint32_t p = 0; p < np; p++) slot[ids[p]] = p;
  std::vector<edb_inner_bond> keep;
  for (auto& b : c->h_bonds) {
    if (b.unbroken) { keep.push_back(b); continue; }
    if (b.friction[0] == 0.0 && b.friction[1] == 0.0 && b.friction[2] == 0.0) continue;
    auto it = slot.find(b.id_i);
    if (it == slot.end()) continue;
    c->h_conv.push_back(BondConv{it->second, b.sub_i, b.sub_j, (long long)b.id_j, b.friction[0], b.friction[1], b.friction[2]});
  }
  c->h_bonds.swap(keep);
}
// [stick_polyhedra] on the device.  `cluster` (field::cluster, not part of the resident field set) is given per particle id.
extern "C" int edb_stick_polyhedra(edb_context* c, double sticking_threshold, uint64_t n, const uint64_t* ids, const double* cluster) {
  API_BEGIN(c)
  if (!c->poly()) throw StateError{EDB_ERR_STATE, "stick_polyhedra needs a polyhedron context"};
  if (!c->have_inc) throw StateError{EDB_ERR_STATE, "stick_polyhedra after chunk_neighbors_contact: call edb_update_particle_neighbors first"};
  if (!c->ibp_groups) throw StateError{EDB_ERR_STATE, "stick_polyhedra before edb_set_inner_bond_params"};
  const uint32_t np = c->nreal;
  std::vector<uint64_t> pid(np);
  CK(cudaStreamSynchronize(c->st));
  if (np) CK(cudaMemcpy(pid.data(), c->pa.id.p, (size_t)np * 8, cudaMemcpyDeviceToHost));
  std::unordered_map<uint64_t, double> cl;
  cl.reserve((size_t)n * 2);
  for (uint64_t k = 0; k < n; k++) cl[ids[k]] = cluster[k];
  std::vector<double> hc(np, -1.0);
  for (uint32_t p = 0; p < np; p++) {
    auto it = cl.find(pid[p]);
    if (it == cl.end()) throw StateError{EDB_ERR_INVALID, "stick_polyhedra: a particle has no cluster value"};
    hc[p] = it->second;
  }
  const std::vector<double>& fa = c->h_farea;
  DevBuf<double> d_cl, d_fa;
  DevBuf<uint32_t> d_cnt, d_off;
  DevBuf<StickBond> d_out;
  d_cl.ensure(np + 1); d_fa.ensure(fa.size() + 1); d_cnt.ensure(np + 1); d_off.ensure(np + 2);
  if (np) CK(cudaMemcpy(d_cl.p, hc.data(), (size_t)np * 8, cudaMemcpyHostToDevice));
  if (!fa.empty()) CK(cudaMemcpy(d_fa.p, fa.data(), fa.size() * 8, cudaMemcpyHostToDevice));
  StickArgs T{NbhArgs{view(c->pa), c->cidx.p, c->cbeg.p, c->cend.p, c->G, c->drv.p, 0, c->rcut_inc, c->symetric, np, c->vert.p, c->S, MeshView{}, 0,
                      nullptr, nullptr},
              d_cl.p, d_fa.p, c->d_ibp.p, c->ibp_groups, sticking_threshold};
  LAUNCH(c, k_stick<0>, nblk(np, 64), 64, 0, T, d_cnt.p, nullptr, nullptr);
  scan_u32(c, d_cnt.p, d_off.p, np);
  uint32_t total = 0;
  CK(cudaMemcpyAsync(&total, d_off.p + np, 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  std::vector<StickBond> hb(total);
  if (total) {
    d_out.ensure(total);
    LAUNCH(c, k_stick<1>, nblk(np, 64), 64, 0, T, d_cnt.p, d_off.p, d_out.p);
    CK(cudaStreamSynchronize(c->st));
    CK(cudaMemcpy(hb.data(), d_out.p, (size_t)total * sizeof(StickBond), cudaMemcpyDeviceToHost));
  }
  c->h_bonds.clear();
  for (const StickBond& b : hb) {
    edb_inner_bond o{};
    o.id_i = pid[b.a]; o.id_j = pid[b.b]; o.sub_i = b.sub_i; o.sub_j = b.sub_j; o.dn0 = b.dn0; o.weight = b.weight;
    o.criteria_1 = b.c1; o.criteria_2 = b.c2; o.mode = b.mode; o.unbroken = 1; o.ghost = 0;
    c->h_bonds.push_back(o);
  }
  d_cl.release(); d_fa.release(); d_cnt.release(); d_off.release(); d_out.release();
  bonds_upload(c);
  API_END(c)
}
static void bond_friction_updater_launch(edb_context* c);
extern "C" int edb_broken_interface_friction_updater(edb_context* c) {
  API_BEGIN(c)
  bond_friction_updater_launch(c);
  API_END(c)
}
extern "C" int edb_set_inner_bond_params(edb_context* c, int32_t ng, const edb_inner_bond_params* p) {
  API_BEGIN(c)
  if (ng < 1) throw StateError{EDB_ERR_INVALID, "n_groups < 1"};
  std::vector<BondParams> t((size_t)ng * ng);
  for (size_t i = 0; i < t.size(); i++) t[i] = BondParams{p[i].kn, p[i].kt, p[i].damp_rate, p[i].mode, 0, p[i].crit1, p[i].crit2};
  c->d_ibp.ensure(t.size());
  CK(cudaMemcpy(c->d_ibp.p, t.data(), t.size() * sizeof(BondParams), cudaMemcpyHostToDevice));
  c->ibp_groups = ng;
  API_END(c)
}
extern "C" int edb_import_inner_bonds(edb_context* c, const edb_inner_bond* in, uint64_t n) {
  API_BEGIN(c)
  if (!c->poly()) throw StateError{EDB_ERR_STATE, "inner bonds need a polyhedron context"};
  c->h_bonds.assign(in, in + n);
  bonds_upload(c);
  API_END(c)
}
extern "C" int edb_export_inner_bonds(edb_context* c, edb_inner_bond* out, uint64_t cap, uint64_t* n) {
  API_BEGIN(c)
  *n = c->h_bonds.size();
  if (!out) return EDB_OK;
  if (cap < c->h_bonds.size()) throw StateError{EDB_ERR_CAPACITY, "edb_export_inner_bonds: capacity too small"};
  bonds_download(c);
  if (!c->h_bonds.empty()) memcpy(out, c->h_bonds.data(), c->h_bonds.size() * sizeof(edb_inner_bond));
  API_END(c)
}
extern "C" int edb_update_interfaces(edb_context* c, uint64_t* n_interfaces) {
  API_BEGIN(c)
  bonds_download(c);
  bonds_upload(c);
  if (n_interfaces) *n_interfaces = c->n_interfaces;
  API_END(c)
}
static void bond_force_launch(edb_context* c, double dt, const unsigned long long* gate) {
  const uint32_t n = (uint32_t)c->h_bonds.size();
  if (n == 0) return;
  if (!c->bonds_on_device) throw StateError{EDB_ERR_STATE, "inner_bond_polyhedron before edb_import_inner_bonds / edb_update_interfaces"};
  if (!c->ibp_groups) throw StateError{EDB_ERR_STATE, "inner_bond_polyhedron before edb_set_inner_bond_params"};
  if (c->S.nvmax <= 0 || !c->vert.p) throw StateError{EDB_ERR_STATE, "inner_bond_polyhedron needs the world vertices (edb_compute_vertices)"};
  LAUNCH(c, k_bond_force, nblk(n), BLK, 0, bond_view(c), view(c->pa), c->vert.p, c->S, c->d_ibp.p, c->ibp_groups, dt, gate);
  LAUNCH(c, k_bond_gather, nblk(c->n_bond_particles), BLK, 0, bond_view(c), view(c->pa), c->b_plist.p, c->b_inc_off.p, c->b_inc.p, c->n_bond_particles, gate);
  c->ctr.interaction_updates += n;
}
static uint64_t interface_criterion_launch(edb_context* c) {
  if (c->n_interfaces == 0) return 0;
  CK(cudaMemsetAsync(c->d_flag, 0, 8, c->st));
  LAUNCH(c, k_interface_criterion, nblk(c->n_interfaces), BLK, 0, bond_view(c), c->b_itf.p, c->n_interfaces, c->b_itf_broken.p, c->d_flag);
  CK(cudaMemcpyAsync(c->h_flag, c->d_flag, 8, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  return c->h_flag[0];
}
// speculative loop: the number of interfaces that break is ADDED to the rebuild trigger of the step (device flag), no host read
static void interface_criterion_async(edb_context* c, unsigned long long* flag) {
  if (c->n_interfaces == 0) return;
  LAUNCH(c, k_interface_criterion, nblk(c->n_interfaces), BLK, 0, bond_view(c), c->b_itf.p, c->n_interfaces, c->b_itf_broken.p, flag);
}
static void bond_friction_updater_launch(edb_context* c) {
  if (c->n_interfaces == 0 || !c->bonds_on_device) return;
  LAUNCH(c, k_bond_friction_updater, nblk(c->n_interfaces), BLK, 0, bond_view(c), c->b_itf.p, c->n_interfaces, c->b_itf_broken.p, view(c->pa), c->prm);
}
extern "C" int edb_inner_bond_polyhedron(edb_context* c, double dt) {
  API_BEGIN(c)
  bond_force_launch(c, dt);
  API_END(c)
}
extern "C" int edb_apply_interface_fracture_criterion(edb_context* c, uint64_t* n_broken) {
  API_BEGIN(c)
  *n_broken = interface_criterion_launch(c);
  API_END(c)
}
extern "C" int edb_get_inner_bond_state(edb_context* c, double* dn, double* cp, double* fn, double* ft) {
  API_BEGIN(c)
  const size_t n = c->h_bonds.size();
  if (n == 0 || !c->bonds_on_device) return EDB_OK;
  std::vector<double4> fd(n), cpv(n), fr(n);
  CK(cudaStreamSynchronize(c->st));
  CK(cudaMemcpy(fd.data(), c->b_fd.p, n * sizeof(double4), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cpv.data(), c->b_cp.p, n * sizeof(double4), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(fr.data(), c->b_fr.p, n * sizeof(double4), cudaMemcpyDeviceToHost));
  for (size_t k = 0; k < n; k++) {
    if (dn) dn[k] = fd[k].w;
    if (cp) { cp[3 * k] = cpv[k].x; cp[3 * k + 1] = cpv[k].y; cp[3 * k + 2] = cpv[k].z; }
    if (fn) { fn[3 * k] = fd[k].x - fr[k].x; fn[3 * k + 1] = fd[k].y - fr[k].y; fn[3 * k + 2] = fd[k].z - fr[k].z; }
    if (ft) { ft[3 * k] = fr[k].x; ft[3 * k + 1] = fr[k].y; ft[3 * k + 2] = fr[k].z; }
  }
  API_END(c)
}

extern "C" int edb_stress_tensor(edb_context* c, double* stress9) {
  API_BEGIN(c)
  if (!c->have_inc) throw StateError{EDB_ERR_STATE, "stress_tensor before update_particle_neighbors"};
  const bool poly = c->poly();
  GatherArgs A{};
  A.P = view(c->pa);
  for (int t = 0; t < NT; t++) A.W.w[t] = view(c->wave[t]);
  const int* order = poly ? ORDER_POLY : ORDER_SPHERE;
  A.norder = poly ? 13 : 7;
  for (int o = 0; o < A.norder; o++) A.order[o] = order[o];
  A.inc_off = c->inc_off.p;
  A.inc = c->inc.p;
  A.nreal = c->nreal;
  c->stage.ensure((size_t)c->nreal * 9 + 1);
  LAUNCH(c, k_stress, nblk(c->nreal), BLK, 0, A, c->stage.p);
  CK(cudaMemcpyAsync(stress9, c->stage.p, (size_t)c->nreal * 9 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  API_END(c)
}

// ---- dem_cost_model (src/interaction/dem_cost_model.cpp:63-92) ------------------------------------------------------------
struct CostOff {
  const uint32_t* off[NT];
};
__constant__ double c_cost_weight[NT] = {1, 3, 5, 4, 1, 1, 1, 1, 3, 5, 4, 3, 5, 1};
// thread per real cell: the owners of a cell are contiguous slots of the cell-sorted arrays, so the records they own in a wave are
// one difference of the wave's CSR offsets
__global__ void k_cell_costs(GridDev G, const uint32_t* __restrict__ cbeg, const uint32_t* __restrict__ cend, CostOff O, double* __restrict__ out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (uint32_t)G.nc[0] * G.nc[1] * G.nc[2]) return;
  const int ci = c % G.nc[0] + 1, cj = (c / G.nc[0]) % G.nc[1] + 1, ck = c / (G.nc[0] * G.nc[1]) + 1;
  const uint32_t lin = cell_lin(G, ci, cj, ck), b = cbeg[lin], e = cend[lin];
  double cost = 3.0 * (double)(e - b) + 1.0;
  if (e > b)
    for (int t = 0; t < NT; t++)
      if (O.off[t]) cost += c_cost_weight[t] * (double)(O.off[t][e] - O.off[t][b]);
  out[c] = cost;
}
// inner bonds (type 13) live in their own arrays: one per bond, in the owner's cell (integer-valued sums: order-independent)
__global__ void k_cell_costs_bonds(GridDev G, const uint2* __restrict__ ij, uint32_t n, const uint32_t* __restrict__ cidx, uint32_t nreal,
                                   double* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t a = ij[k].x;
  if (a >= nreal) return;
  int ci, cj, ck;
  cell_ijk(G, cidx[a], ci, cj, ck);
  atomicAdd(&out[(uint32_t)(ci - 1) + G.nc[0] * ((uint32_t)(cj - 1) + G.nc[1] * (uint32_t)(ck - 1))], c_cost_weight[13]);
}
extern "C" int edb_dem_cost_model(edb_context* c, int32_t grid_dims[3], double* costs, uint64_t cap) {
  API_BEGIN(c)
  const GridDev& G = c->G;
  if (grid_dims)
    for (int d = 0; d < 3; d++) grid_dims[d] = G.nc[d];
  if (costs) {
    const uint64_t ncell = (uint64_t)G.nc[0] * G.nc[1] * G.nc[2];
    if (cap < ncell) throw StateError{EDB_ERR_CAPACITY, "dem_cost_model: cost buffer smaller than the real-cell grid"};
    if (!c->have_cells) throw StateError{EDB_ERR_STATE, "dem_cost_model: no cell grid yet (edb_update_particle_neighbors first)"};
    quiesce(c);
    CostOff O{};
    for (int t = 0; t < NT; t++) O.off[t] = c->wave[t].n ? c->wave[t].off.p : nullptr;
    DevBuf<double> d_cost;
    d_cost.ensure(ncell + 1);
    LAUNCH(c, k_cell_costs, nblk(ncell), BLK, 0, G, c->cbeg.p, c->cend.p, O, d_cost.p);
    if (!c->h_bonds.empty() && c->bonds_on_device)
      LAUNCH(c, k_cell_costs_bonds, nblk(c->h_bonds.size()), BLK, 0, G, c->b_ij.p, (uint32_t)c->h_bonds.size(), c->cidx.p, c->nreal, d_cost.p);
    CK(cudaMemcpyAsync(costs, d_cost.p, ncell * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    d_cost.release();
  }
  API_END(c)
}

extern "C" int edb_simulation_state(edb_context* c, int32_t symetric, edb_sim_state* out) {
  API_BEGIN(c)
  memset(out, 0, sizeof(*out));
  const uint32_t n = c->nreal, nb = nblk(n);
  c->stage.ensure((size_t)nb * 7 + 8);
  std::vector<double> part((size_t)nb * 7);
  if (nb) {
    LAUNCH(c, k_sim_state_particles, nb, BLK, 0, view(c->pa), n, c->stage.p);
    CK(cudaMemcpyAsync(part.data(), c->stage.p, part.size() * 8, cudaMemcpyDeviceToHost, c->st));
  }
  CK(cudaMemsetAsync(c->d_flag, 0, 24, c->st));
  for (int t = 0; t < NT; t++)
    if (c->wave[t].n) LAUNCH(c, k_sim_state_wave, nblk(c->wave[t].n), BLK, 0, view(c->wave[t]), (unsigned long long)((t <= 3 && symetric) ? 2 : 1), c->d_flag);
  CK(cudaMemcpyAsync(c->h_flag, c->d_flag, 24, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (uint32_t b = 0; b < nb; b++)
    for (int q = 0; q < 7; q++) acc[q] += part[(size_t)b * 7 + q];
  for (int q = 0; q < 3; q++) { out->kinetic_energy[q] = acc[q]; out->rotation_energy[q] = acc[3 + q]; }
  out->mass = acc[6];
  out->n_particles = n;
  out->n_total_interactions = c->h_flag[0];
  out->n_active_interactions = c->h_flag[1];
  long long bits = (long long)c->h_flag[2];
  double dn;
  memcpy(&dn, &bits, 8);
  out->min_dn = c->h_flag[2] ? dn : 0.0;
  API_END(c)
}

// ---- multi-GPU -----------------------------------------------------------------------------------------------------
extern "C" int edb_comm_unique_id(uint8_t id_out[128]) { return Comm::unique_id(id_out); }
extern "C" int edb_comm_init(edb_context* c, const uint8_t id[128]) {
  API_BEGIN(c)
  c->comm.init(c->cfg, id);
  API_END(c)
}

// ---- measurement ---------------------------------------------------------------------------------------------------
extern "C" int edb_set_timing(edb_context* c, int32_t on) {
  API_BEGIN(c)
  resolve_timers(c);
  c->timing = on != 0;
  API_END(c)
}
extern "C" int edb_get_counters(edb_context* c, edb_counters* out) {
  API_BEGIN(c)
  resolve_timers(c);
  *out = c->ctr;
  API_END(c)
}
extern "C" int edb_reset_counters(edb_context* c) {
  API_BEGIN(c)
  resolve_timers(c);
  c->ctr = edb_counters{};
  API_END(c)
}
__global__ void k_flush(double4* p, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = make_double4(0, 0, 0, 0);
}
extern "C" int edb_event_record(edb_context* c, int32_t slot) {
  API_BEGIN(c)
  if (slot < 0 || slot >= 8) throw StateError{EDB_ERR_INVALID, "event slot out of range"};
  if (!c->user_ev[slot]) CK(cudaEventCreate(&c->user_ev[slot]));
  CK(cudaEventRecord(c->user_ev[slot], c->st));
  API_END(c)
}
extern "C" int edb_event_elapsed_ms(edb_context* c, int32_t a, int32_t b, double* ms) {
  API_BEGIN(c)
  if (a < 0 || a >= 8 || b < 0 || b >= 8 || !c->user_ev[a] || !c->user_ev[b]) throw StateError{EDB_ERR_INVALID, "event slot not recorded"};
  CK(cudaEventSynchronize(c->user_ev[b]));
  float f = 0;
  CK(cudaEventElapsedTime(&f, c->user_ev[a], c->user_ev[b]));
  *ms = f;
  API_END(c)
}
extern "C" int edb_flush_l2(edb_context* c) {
  API_BEGIN(c)
  const size_t n = (size_t)256 << 20 >> 5;  // 256 MiB > 126 MB L2
  c->l2buf.ensure(n);
  k_flush<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>(c->l2buf.p, n);
  API_END(c)
}
