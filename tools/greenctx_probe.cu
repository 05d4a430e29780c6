// Probe (not part of the product): can the critical path of the blocked Cholesky be isolated on its own
// SM partition with CUDA green contexts?  Measures the start delay of a one-CTA, register-heavy kernel while
// a bulk kernel saturates (a) the whole GPU from a low-priority stream, (b) only the other partition.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s = nullptr; cuGetErrorString(e, &s); printf("driver error %d (%s) at line %d\n", (int)e, s ? s : "?", __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }

// bulk: each CTA spins for `ns` nanoseconds; 128 threads, smem sized by the launch
__global__ void __launch_bounds__(128) bulk(unsigned long long ns, unsigned* smids) {
  extern __shared__ double sm[];
  if (threadIdx.x == 0) { smids[blockIdx.x] = smid(); sm[0] = 0; }
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < ns) { __nanosleep(200); }
}
// critical: one CTA, 256 threads, lots of registers (like the leaf kernel); records start time and smid
__global__ void __launch_bounds__(256) critical(unsigned long long* tstart, unsigned* sm_out, double* sink) {
  double r[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) r[i] = sink[i] + threadIdx.x;
  if (threadIdx.x == 0) { *tstart = gtime(); *sm_out = smid(); }
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < 20000ull) {
#pragma unroll
    for (int i = 0; i < 96; ++i) r[i] = fma(r[i], 1.0000001, 0.5);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 96; ++i) s += r[i];
  sink[threadIdx.x] = s;
}
__global__ void stamp(unsigned long long* t) { if (threadIdx.x == 0) *t = gtime(); }

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  CU(cuInit(0));
  CUdevice dev; CU(cuDeviceGet(&dev, 0));
  CUdevResource all{}; CU(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", all.sm.smCount);
  CUdevResource grpA{}, grpB{};
  unsigned nb = 1;
  CU(cuDevSmResourceSplitByCount(&grpA, &nb, &all, &grpB, 0, 8));
  printf("split: groups %u, A %u SMs, B (remaining) %u SMs\n", nb, grpA.sm.smCount, grpB.sm.smCount);
  CUdevResourceDesc dA, dB;
  CU(cuDevResourceGenerateDesc(&dA, &grpA, 1));
  CU(cuDevResourceGenerateDesc(&dB, &grpB, 1));
  CUgreenCtx gA, gB;
  CU(cuGreenCtxCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CU(cuGreenCtxCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUstream sA, sB, sB2;
  CU(cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, hi));
  CU(cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, lo));
  CU(cuGreenCtxStreamCreate(&sB2, gB, CU_STREAM_NON_BLOCKING, hi));
  cudaStream_t nHi, nLo;
  CK(cudaStreamCreateWithPriority(&nHi, cudaStreamNonBlocking, hi));
  CK(cudaStreamCreateWithPriority(&nLo, cudaStreamNonBlocking, lo));

  const int grid = 148 * 24;
  unsigned* d_sm; CK(cudaMalloc(&d_sm, sizeof(unsigned) * grid));
  unsigned long long* d_t; CK(cudaMalloc(&d_t, sizeof(unsigned long long) * 8));
  unsigned* d_csm; CK(cudaMalloc(&d_csm, sizeof(unsigned) * 8));
  double* d_sink; CK(cudaMalloc(&d_sink, sizeof(double) * 1024)); CK(cudaMemset(d_sink, 0, sizeof(double) * 1024));
  CK(cudaFuncSetAttribute(bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024));
  cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

  auto run = [&](const char* name, cudaStream_t sb, cudaStream_t sc) {
    // bulk: 3 CTAs/SM by shared memory (70 KB each), 60 us per CTA, 24 CTAs per SM queued
    bulk<<<grid, 128, 70 * 1024, sb>>>(60000ull, d_sm);
    CK(cudaEventRecord(ev, sb));   // (not waited on: just exercises event record on this stream)
    // let the bulk kernel occupy the machine, then time a critical launch 5 times
    std::vector<double> delays;
    for (int rep = 0; rep < 5; ++rep) {
      stamp<<<1, 32, 0, sc>>>(d_t + 0);   // tiny kernel: when does the stream reach this point
      critical<<<1, 256, 0, sc>>>(d_t + 1, d_csm, d_sink);
      CK(cudaStreamSynchronize(sc));
      unsigned long long h[2]; unsigned hs;
      CK(cudaMemcpy(h, d_t, sizeof h, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&hs, d_csm, sizeof hs, cudaMemcpyDeviceToHost));
      delays.push_back((double)(h[1] - h[0]) * 1e-3);
      if (rep == 0) printf("  [%s] critical ran on smid %u\n", name, hs);
    }
    CK(cudaDeviceSynchronize());
    std::vector<unsigned> hsm(grid);
    CK(cudaMemcpy(hsm.data(), d_sm, sizeof(unsigned) * grid, cudaMemcpyDeviceToHost));
    std::set<unsigned> used(hsm.begin(), hsm.end());
    printf("[%s] bulk used %zu distinct SMs (min %u max %u); stamp->critical start delay us:", name, used.size(), *used.begin(), *used.rbegin());
    for (double d : delays) printf(" %.1f", d);
    printf("\n");
  };
  // warm-up
  stamp<<<1, 32, 0, nHi>>>(d_t); critical<<<1, 256, 0, nHi>>>(d_t + 1, d_csm, d_sink); CK(cudaDeviceSynchronize());
  run("plain streams: bulk low prio, critical high prio", nLo, nHi);
  run("green ctx: bulk on B (140 SMs), critical on A (8 SMs)", (cudaStream_t)sB, (cudaStream_t)sA);
  run("green ctx: bulk on B low, critical on B high", (cudaStream_t)sB, (cudaStream_t)sB2);

  // cross-partition event dependency + graph capture across green-context streams
  {
    cudaEvent_t e1, e2; CK(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    stamp<<<1, 32, 0, (cudaStream_t)sA>>>(d_t + 2);
    CK(cudaEventRecord(e1, (cudaStream_t)sA));
    CK(cudaStreamWaitEvent((cudaStream_t)sB, e1, 0));
    stamp<<<1, 32, 0, (cudaStream_t)sB>>>(d_t + 3);
    CK(cudaDeviceSynchronize());
    unsigned long long h[2]; CK(cudaMemcpy(h, d_t + 2, sizeof h, cudaMemcpyDeviceToHost));
    printf("event A->B: ordered %s (dt %.1f us)\n", h[1] >= h[0] ? "yes" : "NO", (double)(h[1] - h[0]) * 1e-3);
    cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
    cudaError_t e = cudaStreamBeginCapture(nHi, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      stamp<<<1, 32, 0, nHi>>>(d_t + 4);
      cudaEventRecord(e1, nHi);
      cudaError_t ew = cudaStreamWaitEvent((cudaStream_t)sA, e1, 0);
      stamp<<<1, 32, 0, (cudaStream_t)sA>>>(d_t + 5);
      cudaEventRecord(e2, (cudaStream_t)sA);
      cudaStreamWaitEvent(nHi, e2, 0);
      e = cudaStreamEndCapture(nHi, &g);
      printf("graph capture across a green-context stream: wait %s, end %s\n", cudaGetErrorString(ew), cudaGetErrorString(e));
      if (e == cudaSuccess && g) {
        e = cudaGraphInstantiate(&ge, g, 0);
        printf("  instantiate: %s\n", cudaGetErrorString(e));
        if (e == cudaSuccess) { e = cudaGraphLaunch(ge, nHi); cudaError_t e2s = cudaDeviceSynchronize(); printf("  launch: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2s)); }
      }
      cudaGetLastError();
    } else printf("begin capture failed: %s\n", cudaGetErrorString(e));
  }
  printf("done\n");
  return 0;
}
