// tests/host_check_bond.cu — TEST HELPER: the product's inner-bond law and fracture criterion (exadem_b200/csrc/kernels_bond.cuh,
// __host__ __device__) compiled for the HOST behind the same flat interface as the oracle's, so the CPU suite compares the code the GPU
// runs with the oracle (and, through it, with the reference's force_law_core) bit for bit.
#include "../exadem_b200/csrc/kernels_bond.cuh"
using namespace edb;
static v3 V(const double* p) { return mk3(p[0], p[1], p[2]); }
static void S(double* o, v3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
extern "C" void chk_inner_bond_force_core(double dn, const double* n, double dn0, double weight, double dt, const edb_inner_bond_params* p,
                                          double meff, double* en_et, double* tds, double* ft, const double* cp, const double* pi,
                                          const double* vi, double* fi, const double* wi, const double* pj, const double* vj, const double* wj) {
  BondParams P{p->kn, p->kt, p->damp_rate, p->mode, 0, p->crit1, p->crit2};
  v3 T = V(tds), F = V(ft), f = mk3(0, 0, 0);
  inner_bond_force_core(dn, V(n), dn0, weight, dt, P, meff, en_et[0], T, en_et[1], F, V(cp), V(pi), V(vi), f, V(wi), V(pj), V(vj), V(wj));
  S(tds, T); S(ft, F); S(fi, f);
}
extern "C" int chk_interface_breaks(int mode, double c1, double c2, const double* en, const double* et, const double* dn, const double* fn3,
                                    unsigned long long size) {
  double En = 0.0, Et = 0.0, Sn = 0.0;
  for (unsigned long long j = 0; j < size; j++) {
    En += en[j];
    Et += et[j];
    if (dn[j] > 0.0) Sn += norm(V(fn3 + 3 * j));
  }
  Sn /= size;
  return interface_breaks(mode, c1, c2, En, Et, Sn) ? 1 : 0;
}
