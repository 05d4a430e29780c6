// Latency micro-benchmarks behind the leaf kernel (not part of the product library).
#include <cstdio>
#include <vector>
#include "dense.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void lat_dfma(double* out, int n) { double x = out[0], a = out[1], b = out[2]; long long t0 = clock64(); for (int i = 0; i < n; ++i) x = fma(x, a, b); long long t1 = clock64(); out[3] = x; if (threadIdx.x == 0) ((long long*)out)[4] = t1 - t0; }
__global__ void lat_drcp(double* out, int n) { double x = out[0]; long long t0 = clock64(); for (int i = 0; i < n; ++i) x = __drcp_rn(x) + 1.0; long long t1 = clock64(); out[3] = x; if (threadIdx.x == 0) ((long long*)out)[4] = t1 - t0; }
__global__ void lat_ddiv(double* out, int n) { double x = out[0]; long long t0 = clock64(); for (int i = 0; i < n; ++i) x = 1.0 / x + 1.0; long long t1 = clock64(); out[3] = x; if (threadIdx.x == 0) ((long long*)out)[4] = t1 - t0; }
__global__ void lat_dsqrt(double* out, int n) { double x = out[0]; long long t0 = clock64(); for (int i = 0; i < n; ++i) x = sqrt(x) + 2.0; long long t1 = clock64(); out[3] = x; if (threadIdx.x == 0) ((long long*)out)[4] = t1 - t0; }
__global__ void lat_bar(double* out, int n) { long long t0 = clock64(); for (int i = 0; i < n; ++i) __syncthreads(); long long t1 = clock64(); if (threadIdx.x == 0) ((long long*)out)[4] = t1 - t0; }
// smem ping-pong: one thread writes, barrier, all read, dependent fma, next
__global__ void lat_pingpong(double* out, int n) {
  __shared__ double buf[2][128];
  double x = out[0] + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if ((threadIdx.x >> 4) == (i & 15)) buf[i & 1][threadIdx.x & 127] = x;
    __syncthreads();
    x = fma(buf[i & 1][i & 127], 1.0000001, x);
  }
  long long t1 = clock64();
  out[3 + threadIdx.x % 2] = x;
  if (threadIdx.x == 0) ((long long*)out)[4] = t1 - t0;
}
__global__ void lat_lds(double* out, int n) { __shared__ double buf[256]; buf[threadIdx.x] = threadIdx.x; __syncthreads(); int idx = threadIdx.x; long long t0 = clock64(); for (int i = 0; i < n; ++i) idx = (int)buf[idx] ; long long t1 = clock64(); out[3] = idx; if (threadIdx.x == 0) ((long long*)out)[4] = t1 - t0; }

template <class F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); for (int r = 0; r < reps; ++r) f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  double* d; CK(cudaMalloc(&d, 4096)); double h[8] = {1.5, 0.999, 0.001, 0, 0, 0, 0, 0};
  const int n = 4096;
  auto run = [&](const char* name, void (*k)(double*, int), int threads) {
    CK(cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
    k<<<1, threads>>>(d, n); CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, (char*)d + 32, 8, cudaMemcpyDeviceToHost));
    printf("%-36s %7.1f cycles / iteration (%d threads)\n", name, (double)cyc / n, threads);
  };
  run("dependent DFMA", lat_dfma, 32); run("dependent DFMA (256 thr)", lat_dfma, 256);
  run("dependent __drcp_rn + add", lat_drcp, 32); run("dependent 1.0/x + add", lat_ddiv, 32);
  run("dependent sqrt + add", lat_dsqrt, 32);
  run("__syncthreads (256 thr)", lat_bar, 256); run("__syncthreads (128 thr)", lat_bar, 128);
  run("STS -> BAR -> LDS -> DFMA (256 thr)", lat_pingpong, 256);
  run("dependent LDS (pointer chase)", lat_lds, 32);
  // the leaf kernels themselves
  const int T = 128; const long long ld = 128;
  std::vector<double> A(T * T, 0.0);
  for (int j = 0; j < T; ++j) for (int i = j; i < T; ++i) A[i + j * T] = (i == j) ? T + 1.0 : 1.0 / (1 + i - j);
  double *G, *X, *ldp; int* st;
  CK(cudaMalloc(&G, sizeof(double) * T * T)); CK(cudaMalloc(&X, sizeof(double) * T * T)); CK(cudaMalloc(&ldp, 8)); CK(cudaMalloc(&st, 4));
  csdense::configure_kernels();
  for (int tile : {128, 64}) {
    float ms = time_ms([&] { CK(cudaMemcpyAsync(G, A.data(), sizeof(double) * T * T, cudaMemcpyHostToDevice));
                             csdense::launch_leaf(tile, G, ld, X, ld, 0, ldp, st, 1 << 30, 0, nullptr); }, 50);
    float ms0 = time_ms([&] { CK(cudaMemcpyAsync(G, A.data(), sizeof(double) * T * T, cudaMemcpyHostToDevice)); }, 50);
    printf("leaf_kernel<%d>: %.1f us (incl. H2D %.1f us)\n", tile, ms * 1e3, ms0 * 1e3);
  }
  return 0;
}
