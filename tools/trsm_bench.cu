// Micro-benchmark of the chain kernels of plan v2 (leaf factor, leaf inverse, triangular solve) on an idle B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I causalstump_b200/csrc tools/trsm_bench.cu -o tools/trsm_bench
#include <cstdio>
#include <vector>
#include "cs_rt.h"
#include "dense.cuh"
using namespace csdense;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <class F> float time_us(F f, int reps = 20) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); for (int r = 0; r < reps; ++r) f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
  return ms * 1e3f / reps;
}

int main() {
  const int np = 4096, T = 128;
  std::vector<double> h((size_t)np * np, 0.0);
  for (int j = 0; j < np; ++j) for (int i = j; i < np; ++i) h[i + (size_t)j * np] = (i == j) ? 2.0 + 1e-3 * (j % 7) : 1e-3 * ((i * 31 + j * 17) % 13 - 6);
  double *G, *X, *ld; int* st;
  CK(cudaMalloc(&G, sizeof(double) * np * np)); CK(cudaMalloc(&X, sizeof(double) * np * np)); CK(cudaMalloc(&ld, 8 * 64)); CK(cudaMalloc(&st, 4 * 64));
  CK(cudaMemcpy(G, h.data(), sizeof(double) * np * np, cudaMemcpyHostToDevice)); CK(cudaMemset(st, 0, 4 * 64)); CK(cudaMemset(X, 0, sizeof(double) * np * np));
  configure_kernels();
  for (int nrows : {128, 384, 1024, 3968}) {
    float us = time_us([&] { launch_trsm(T, G, np, G + T, np, nrows, 0, nullptr); });
    printf("trsm tile 128, %4d rows: %7.2f us\n", nrows, us);
  }
  for (int nrows : {64, 704}) {
    float us = time_us([&] { launch_trsm(64, G, np, G + 64, np, nrows, 0, nullptr); });
    printf("trsm tile  64, %4d rows: %7.2f us\n", nrows, us);
  }
  for (int mode : {0, 1, 2}) {
    CK(cudaMemcpy(G, h.data(), sizeof(double) * np * T, cudaMemcpyHostToDevice));
    float us = time_us([&] { launch_leaf(T, G, np, X, np, 0, ld, st, np, 0, nullptr, 1, 0, mode); }, 1);
    printf("leaf tile 128 mode %d: %7.2f us\n", mode, us);
  }
  { float us = time_us([&] { cudaMemsetAsync(ld, 0, 8, 0); }); printf("memset (launch floor): %7.2f us\n", us); }
  return 0;
}
