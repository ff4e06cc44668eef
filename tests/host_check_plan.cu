// tests/host_check_plan.cu — TEST HELPER: the product's ghost-layer plan (exadem_b200/csrc/comm.cuh: Comm::source_of, plan_send_lists,
// the enumeration of setup_grid) run on the HOST for every rank of a process grid, uniform or cut at given positions
// (edb_set_brick_cuts), so that the CPU suite can compare it with its numpy mirror exadem_b200.parallel.ghost_plan / send_plan.
//   nvcc -O2 -std=c++17 -Xcompiler -fPIC -shared tests/host_check_plan.cu -o libhost_check_plan.so
#include "../exadem_b200/csrc/comm.cuh"
#include "../exadem_b200/csrc/kernels_migrate.cuh"
using namespace edb;

static void grid_of(const edb_config& cfg, GridDev& G) {   // as setup_grid (edb_lib.cu)
  for (int d = 0; d < 3; d++) {
    const double len = cfg.hi[d] - cfg.lo[d];
    G.nc[d] = std::max(1, (int)std::floor(len / cfg.cell_size + 1e-9));
    G.cs[d] = len / G.nc[d];
    G.g[d] = G.nc[d] + 2;
    G.lo[d] = cfg.lo[d]; G.glo[d] = cfg.glo[d]; G.ghi[d] = cfg.ghi[d]; G.periodic[d] = cfg.periodic[d];
  }
}
// recv entries of `cfg->rank` (ghost cell ijk, source rank, source cell ijk, shift: 10 doubles each) and its send entries
// (destination rank, ghost cell lin in the receiver's grid, source cell lin in our grid, shift: 6 doubles each)
extern "C" int chk_plan(const edb_config* cfg, const double* cx, const double* cy, const double* cz, double* recv, int cap_recv, int* n_recv,
                        double* send, int cap_send, int* n_send) {
  try {
    GridDev G{};
    grid_of(*cfg, G);
    Comm cm;
    cm.plan_reset(*cfg, G);
    const double* in[3] = {cx, cy, cz};
    for (int d = 0; d < 3; d++)
      if (in[d]) {
        cm.cuts[d].assign(in[d], in[d] + cm.pg[d] + 1);
        cm.cuts[d][cm.rc[d]] = cfg->lo[d];
        cm.cuts[d][cm.rc[d] + 1] = cfg->hi[d];
      }
    int nr = 0;
    for (int k = 0; k < G.g[2]; k++)
      for (int j = 0; j < G.g[1]; j++)
        for (int i = 0; i < G.g[0]; i++) {
          if (!(i == 0 || j == 0 || k == 0 || i == G.g[0] - 1 || j == G.g[1] - 1 || k == G.g[2] - 1)) continue;
          const int idx[3] = {i, j, k};
          int src[3], src_g[3], peer;
          double shift[3];
          if (!cm.ghost_source(*cfg, G, idx, src, src_g, shift, &peer)) continue;
          if (nr < cap_recv) {
            double* o = recv + 10 * nr;
            o[0] = i; o[1] = j; o[2] = k; o[3] = peer; o[4] = src[0]; o[5] = src[1]; o[6] = src[2]; o[7] = shift[0]; o[8] = shift[1]; o[9] = shift[2];
          }
          nr++;
        }
    *n_recv = nr;
    int ns = 0;
    if (cfg->nranks > 1) {
      cm.plan_send_lists(*cfg, G);
      for (auto& kv : cm.peers)
        for (auto& e : kv.second.send) {
          if (ns < cap_send) {
            double* o = send + 6 * ns;
            o[0] = kv.first; o[1] = e.gcell; o[2] = e.scell; o[3] = e.shift[0]; o[4] = e.shift[1]; o[5] = e.shift[2];
          }
          ns++;
        }
    }
    *n_send = ns;
    return 0;
  } catch (const StateError&) {
    return 1;
  }
}
// brick_of (kernels_migrate.cuh) on the host: destination rank of n positions
extern "C" int chk_brick_of(const edb_config* cfg, const double* cx, const double* cy, const double* cz, const double* pos, int n, int* out) {
  BrickMap B{};
  const double* in[3] = {cx, cy, cz};
  for (int d = 0; d < 3; d++) {
    B.glo[d] = cfg->glo[d]; B.ghi[d] = cfg->ghi[d]; B.pg[d] = cfg->pgrid[d]; B.periodic[d] = cfg->periodic[d];
    B.len[d] = (cfg->ghi[d] - cfg->glo[d]) / cfg->pgrid[d];
    B.ncut[d] = in[d] ? cfg->pgrid[d] + 1 : 0;
    for (int k = 0; k < B.ncut[d]; k++) B.cut[d][k] = in[d][k];
  }
  for (int i = 0; i < n; i++) out[i] = brick_of(B, pos + 3 * i);
  return 0;
}
