// Dependent-issue latencies of the FP64 / shuffle instructions the chain kernels are made of (one warp, clock64 brackets).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/lat_bench.cu -o tools/lat_bench
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int N = 4096;
__global__ void k(double* out, long long* clk, double seed, int mode) {
  double a = seed + threadIdx.x * 1e-9, b = 1.0 + seed * 1e-9, c = seed * 1e-12;
  float fa = (float)a, fb = (float)b, fc = (float)c;
  long long t0 = clock64();
  if (mode == 0) { for (int i = 0; i < N; ++i) a = fma(a, b, c); }
  else if (mode == 1) { for (int i = 0; i < N; ++i) a = a * b; }
  else if (mode == 2) { for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (i + 1) & 31); }
  else if (mode == 3) { for (int i = 0; i < N; ++i) { double x = __shfl_sync(0xffffffffu, a, i & 31) * b; a = fma(-x, c, a); } }
  else if (mode == 4) { for (int i = 0; i < N; ++i) fa = fmaf(fa, fb, fc); a = fa; }
  else if (mode == 5) { for (int i = 0; i < N; ++i) a = a + b; }
  else if (mode == 6) { for (int i = 0; i < N; ++i) a = 1.0 / a; }
  else if (mode == 7) { for (int i = 0; i < N; ++i) a = sqrt(a) + b; }
  else if (mode == 8) { for (int i = 0; i < N; ++i) a = rsqrt(a) + b; }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  double* out; long long* clk; CK(cudaMalloc(&out, 8 * 1024 * 1024)); CK(cudaMalloc(&clk, 8));
  const char* names[] = {"DFMA", "DMUL", "SHFL.64 (2 x SHFL)", "SHFL.64 + DMUL + DFMA (one solve step)", "FFMA", "DADD", "1.0/x (f64)", "sqrt + DADD", "rsqrt + DADD"};
  for (int warps : {1, 8, 16}) for (int mode = 0; mode < 9; ++mode) {
    k<<<1, 32 * warps>>>(out, clk, 1.0 + mode, mode); CK(cudaDeviceSynchronize());
    k<<<1, 32 * warps>>>(out, clk, 1.0 + mode, mode); CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
    printf("%2d warps/CTA  %-40s %7.1f clk per dependent op\n", warps, names[mode], (double)c / N);
  }
  return 0;
}
