// Do DFMA (FP64 FMA pipe) and DMMA.8x8x4 (FP64 tensor) instructions share an execution pipe on sm_100a?
// Three kernels with identical launch shapes: FMA only, DMMA only, both interleaved.  If the mixed kernel takes about
// max(t_fma, t_dmma) the pipes are separate; if it takes about t_fma + t_dmma they are one.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>  // 1 = FMA, 2 = DMMA, 3 = both
__global__ void __launch_bounds__(256) probe(double* out, int iters, double x, double y) {
  double f[8], c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x + i; c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE & 1) {  // 4 DFMA = 8 flops/thread: 256 flops per warp-instruction x4
        f[i] = fma(f[i], x, y); f[i] = fma(f[i], y, x);
        f[i] = fma(f[i], x, y); f[i] = fma(f[i], y, x);
      }
      if (MODE & 2) dmma(c[i][0], c[i][1], x, y);  // 512 flops per warp-instruction
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(double* out, int grid, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  probe<MODE><<<grid, 256>>>(out, iters / 10, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  probe<MODE><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int grid = pr.multiProcessorCount * 4, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  const float t1 = run<1>(out, grid, iters), t2 = run<2>(out, grid, iters), t3 = run<3>(out, grid, iters);
  const double warps = (double)grid * 8, n = (double)iters * 8;
  const double f_fma = warps * n * 4 * 64, f_mma = warps * n * 512;
  printf("%s SMs=%d\n", pr.name, pr.multiProcessorCount);
  printf("FMA only : %8.3f ms  %6.2f TFLOP/s\n", t1, f_fma / t1 * 1e-9);
  printf("DMMA only: %8.3f ms  %6.2f TFLOP/s\n", t2, f_mma / t2 * 1e-9);
  printf("both     : %8.3f ms  %6.2f TFLOP/s  (sum of the two alone %.3f ms, max %.3f ms)\n", t3, (f_fma + f_mma) / t3 * 1e-9, t1 + t2,
         t1 > t2 ? t1 : t2);
  return cudaGetLastError() != cudaSuccess;
}
