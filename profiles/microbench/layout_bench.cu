// layout_bench.cu — microbenchmark behind the HBM layout decision: sparse 32-byte sector accesses spread over
// several SoA arrays vs one 128-byte line per particle / per interaction record.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ double4 ld4(const double4* p) {
  double4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st4(double4* p, double4 v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
// record pass: hot records read 2 sectors and write 4; SoA: 4 arrays, AoS: 4 consecutive sectors
template <bool AOS>
__global__ void k_rec(const unsigned char* hot, double4* a0, double4* a1, double4* a2, double4* a3, unsigned n) {
  unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || !hot[k]) return;
  double4 *p0, *p1, *p2, *p3;
  if (AOS) { p0 = a0 + 4ull * k; p1 = p0 + 1; p2 = p0 + 2; p3 = p0 + 3; }
  else { p0 = a0 + k; p1 = a1 + k; p2 = a2 + k; p3 = a3 + k; }
  double4 x = ld4(p0), y = ld4(p1);
  x.x += y.y; y.x += x.z;
  st4(p0, x); st4(p1, y); st4(p2, x); st4(p3, y);
}
// particle gather: each record gathers 3 sectors of particle i and of particle j
template <bool AOS>
__global__ void k_gat(const uint2* ij, const unsigned char* hot, const double4* a0, const double4* a1, const double4* a2, double* out, unsigned n) {
  unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint2 p = ij[k];
  const double4 *i0, *i1, *i2, *j0, *j1, *j2;
  if (AOS) { i0 = a0 + 4ull * p.x; i1 = i0 + 1; i2 = i0 + 2; j0 = a0 + 4ull * p.y; j1 = j0 + 1; j2 = j0 + 2; }
  else { i0 = a0 + p.x; i1 = a1 + p.x; i2 = a2 + p.x; j0 = a0 + p.y; j1 = a1 + p.y; j2 = a2 + p.y; }
  double4 x = ld4(i0), y = ld4(j0);
  double s = x.x - y.x;
  if (hot[k]) { double4 a = ld4(i1), b = ld4(i2), c = ld4(j1), d = ld4(j2); s += a.x + b.y + c.z + d.w; }
  if (s == 123.456) out[k] = s;
}
int main() {
  const unsigned NP = 1000000, NI = 3000000;
  std::vector<uint2> ij(NI);
  std::vector<unsigned char> hot(NI);
  srand(1);
  for (unsigned k = 0; k < NI; k++) {
    unsigned i = k / 3;
    long j = (long)i + (rand() % 3 - 1) * 10000L + (rand() % 3 - 1) * 100L + (rand() % 40 - 20);
    if (j < 0) j = 0; if (j >= NP) j = NP - 1;
    ij[k] = make_uint2(i, (unsigned)j);
    hot[k] = (rand() % 100) < 42;
  }
  uint2* d_ij; unsigned char* d_hot; double4 *P, *R; double* out;
  cudaMalloc(&d_ij, NI * 8); cudaMalloc(&d_hot, NI); cudaMalloc(&P, 4ull * NP * 32); cudaMalloc(&R, 4ull * NI * 32); cudaMalloc(&out, NI * 8);
  cudaMemcpy(d_ij, ij.data(), NI * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_hot, hot.data(), NI, cudaMemcpyHostToDevice);
  cudaMemset(P, 0, 4ull * NP * 32); cudaMemset(R, 0, 4ull * NI * 32);
  double4* flush; cudaMalloc(&flush, 512u << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto launch, double bytes) {
    float best = 1e9;
    for (int it = 0; it < 6; it++) {
      cudaMemsetAsync(flush, it, 512u << 20);
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    printf("%-28s %8.1f us   %7.1f GB/s (useful bytes %.0f MB)\n", name, best * 1e3, bytes / best / 1e6, bytes / 1e6);
  };
  const int B = 256; const unsigned gi = (NI + B - 1) / B;
  double nhot = 0; for (auto h : hot) nhot += h;
  run("records SoA (4 arrays)", [&] { k_rec<false><<<gi, B>>>(d_hot, R, R + NI, R + 2ull * NI, R + 3ull * NI, NI); }, nhot * 192 + NI);
  run("records AoS (128 B line)", [&] { k_rec<true><<<gi, B>>>(d_hot, R, 0, 0, 0, NI); }, nhot * 192 + NI);
  run("gather SoA (3 arrays)", [&] { k_gat<false><<<gi, B>>>(d_ij, d_hot, P, P + NP, P + 2ull * NP, out, NI); }, NI * 9.0 + NP * 96.0);
  run("gather AoS (128 B line)", [&] { k_gat<true><<<gi, B>>>(d_ij, d_hot, P, 0, 0, out, NI); }, NI * 9.0 + NP * 96.0);
  return 0;
}
