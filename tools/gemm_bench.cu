// Micro-benchmark for tuning csg::gemm_kernel on a B200 (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I causalstump_b200/csrc tools/gemm_bench.cu -o tools/gemm_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gemm.cuh"
#include "gemm_mt.cuh"
using namespace csg;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void fill(double* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) { unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; p[i] = (double)(h & 0xffff) / 65536.0 - 0.5; }
}

// raw pipe ceilings
__global__ void dmma_peak(double* out, int iters) {
  double c[16][2];
  for (int i = 0; i < 16; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) csd::dmma8x8x4(c[i], a, b);
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_peak(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-7;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  CK(cudaGetLastError());
  return best;
}

static GemmProblem* dP;
template <int TILE, int WM, int WN, bool AK, bool BK_, int KC, int NS, int MINB = 1>
void run(const char* name, GemmProblem P, double flops) {
  auto k = gemm_kernel<TILE, WM, WN, AK, BK_, KC, NS, MINB>;
  constexpr int smem = gemm_smem_bytes<TILE, AK, BK_, KC, NS>();
  if (smem > 227 * 1024) { printf("%-44s skipped (smem %d)\n", name, smem); return; }
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaMemcpy(dP, &P, sizeof P, cudaMemcpyHostToDevice));
  float ms = time_ms([&] { k<<<P.ntiles, WM * WN * 32, smem>>>(dP, 1, nullptr); });
  printf("%-44s %8.3f ms  %6.2f TFLOP/s  (smem %d KB, %d thr)\n", name, ms, flops / (ms * 1e-3) / 1e12, smem / 1024, WM * WN * 32);
}

template <int TILE, int WM, int WN, bool AK, bool BK_, int KC, int NS>
void run_mt(const char* name, GemmProblem P, double flops, int tpc) {
  auto k = gemm_kernel_mt<TILE, WM, WN, AK, BK_, KC, NS>;
  constexpr int smem = gemm_smem_bytes<TILE, AK, BK_, KC, NS>();
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaMemcpy(dP, &P, sizeof P, cudaMemcpyHostToDevice));
  const int grid = (P.ntiles + tpc - 1) / tpc;
  float ms = time_ms([&] { k<<<grid, WM * WN * 32, smem>>>(dP, 1, nullptr, tpc, P.ntiles); });
  printf("%-36s tpc %2d %8.3f ms  %6.2f TFLOP/s  (%d CTAs)\n", name, tpc, ms, flops / (ms * 1e-3) / 1e12, grid);
}

#define ALLCFG(AK, BK_, MK, FL, tag)                                                   \
  run<128, 2, 4, AK, BK_, 16, 4>(tag " 128 2x4 kc16 s4", MK(128), FL);               \
  run<128, 2, 4, AK, BK_, 32, 3>(tag " 128 2x4 kc32 s3", MK(128), FL);               \
  run<64, 2, 2, AK, BK_, 16, 4>(tag " 64 2x2 kc16 s4 (2 CTA/SM)", MK(64), FL);       \
  run<64, 2, 2, AK, BK_, 16, 3>(tag " 64 2x2 kc16 s3 (3 CTA/SM)", MK(64), FL);       \
  run<64, 2, 2, AK, BK_, 16, 2, 4>(tag " 64 2x2 kc16 s2 minb4 (shipped)", MK(64), FL); \
  run<64, 2, 2, AK, BK_, 32, 3>(tag " 64 2x2 kc32 s3 (2 CTA/SM)", MK(64), FL);       \
  run<64, 2, 2, AK, BK_, 32, 2>(tag " 64 2x2 kc32 s2 (3 CTA/SM)", MK(64), FL);       \
  run<64, 1, 4, AK, BK_, 16, 3>(tag " 64 1x4 kc16 s3", MK(64), FL);                  \
  run<64, 4, 1, AK, BK_, 16, 3>(tag " 64 4x1 kc16 s3", MK(64), FL);                  \
  run<64, 2, 4, AK, BK_, 16, 3>(tag " 64 2x4 kc16 s3 (256 thr)", MK(64), FL);

int main() {
  const int n = 8192;
  const long long ld = n;
  double *A, *B, *C, *out;
  CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
  CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 256)); CK(cudaMalloc(&dP, sizeof(GemmProblem)));
  fill<<<(n * (size_t)n + 255) / 256, 256>>>(A, (size_t)n * n, 1); fill<<<(n * (size_t)n + 255) / 256, 256>>>(B, (size_t)n * n, 2);
  CK(cudaMemset(C, 0, sizeof(double) * n * n));
  CK(cudaDeviceSynchronize());
  {  // pipe ceilings: 148*4 CTAs x 256 threads
    const int iters = 20000, ctas = 148 * 4;
    float ms = time_ms([&] { dmma_peak<<<ctas, 256>>>(out, iters); });
    printf("DMMA.8x8x4 register-only peak: %.2f TFLOP/s\n", 2.0 * 256 * 16.0 * iters * ctas * 8 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { dfma_peak<<<ctas, 256>>>(out, iters); });
    printf("DFMA register-only peak:       %.2f TFLOP/s\n", 2.0 * 16.0 * iters * ctas * 256 / (ms * 1e-3) / 1e12);
  }
  auto mk = [&](int TILE, int mt, int nt, int K, int kmode, int order, int mirror, double alpha, double beta) {
    GemmProblem P{}; P.A = A; P.B = B; P.C = C; P.lda = ld; P.ldb = ld; P.ldc = ld; P.mt = mt; P.nt = nt; P.K = K; P.kmode = kmode;
    P.order = order; P.mirror = mirror; P.set_scale(alpha, beta); P.tile_base = 0; P.ntiles = order == ORD_LOWER ? lower_tiles(mt) : mt * nt;
    return P; };
  if (getenv("MT_ONLY")) {
    printf("\n== several tiles per CTA (gemm_kernel_mt) vs one (gemm_kernel): 7680 rows lower, K = 512\n");
    { double fl = 2.0 * 64 * 64 * 512.0 * lower_tiles(120);
      auto P1 = mk(64, 120, 120, 512, KM_FULL, ORD_LOWER, 0, -1.0, 1.0);
      run<64, 2, 2, false, false, 16, 3>("NT512 beta=1 one tile per CTA", P1, fl);
      for (int tpc : {1, 2, 3, 4, 6, 8, 12}) run_mt<64, 2, 2, false, false, 16, 3>("NT512 beta=1", P1, fl, tpc);
      auto P0 = mk(64, 120, 120, 512, KM_FULL, ORD_LOWER, 0, -1.0, 0.0);
      run<64, 2, 2, false, false, 16, 3>("NT512 beta=0 one tile per CTA", P0, fl);
      for (int tpc : {1, 4, 8}) run_mt<64, 2, 2, false, false, 16, 3>("NT512 beta=0", P0, fl, tpc);
      auto PT = mk(64, 120, 120, 512, KM_FULL, ORD_LOWER, 1, 1.0, 1.0);
      run<64, 2, 2, true, true, 16, 2>("TN512 beta=1 mirror one tile per CTA", PT, fl);
      for (int tpc : {1, 2, 4, 8}) run_mt<64, 2, 2, true, true, 16, 2>("TN512 beta=1 mirror", PT, fl, tpc);
      auto PN = mk(64, 120, 120, 512, KM_FULL, ORD_LOWER, 0, -1.0, 1.0);
      run<64, 2, 2, false, true, 16, 3>("NN512 beta=1 one tile per CTA", PN, fl);
      for (int tpc : {1, 2, 4, 8}) run_mt<64, 2, 2, false, true, 16, 3>("NN512 beta=1", PN, fl, tpc);
    }
    { double fl = 2.0 * 64 * 64 * 128.0 * lower_tiles(120);
      auto P1 = mk(64, 120, 120, 128, KM_FULL, ORD_LOWER, 0, -1.0, 1.0);
      run<64, 2, 2, false, false, 16, 3>("NT128 beta=1 one tile per CTA", P1, fl);
      for (int tpc : {2, 4, 8, 16}) run_mt<64, 2, 2, false, false, 16, 3>("NT128 beta=1", P1, fl, tpc);
    }
    { double fl = (double)n * n * n / 3.0;
      auto P1 = mk(64, 128, 128, n, KM_GE_I, ORD_LOWER, 1, 1.0, 0.0);
      run<64, 2, 2, true, true, 16, 2>("TN X^T X one launch, one tile per CTA", P1, fl);
      for (int tpc : {2, 4}) run_mt<64, 2, 2, true, true, 16, 2>("TN X^T X one launch", P1, fl, tpc);
    }
    printf("\n== whole waves: 96 x 74 = 7104 = 12 x 592 tiles (4 CTAs/SM), tiles per CTA 1 / 4 / 12, K sweep\n");
    for (int K : {128, 256, 512, 1024, 2048, 4096}) {
      double fl = 2.0 * 64 * 64 * (double)K * 7104;
      char nm[64];
      auto PR = mk(64, 96, 74, K, KM_FULL, ORD_COL, 0, -1.0, 1.0);
      snprintf(nm, sizeof nm, "NT K=%d", K);
      for (int tpc : {1, 4, 12}) run_mt<64, 2, 2, false, false, 16, 3>(nm, PR, fl, tpc);
      snprintf(nm, sizeof nm, "TN K=%d (5 CTAs/SM: 9.6 waves)", K);
      for (int tpc : {1, 4}) run_mt<64, 2, 2, true, true, 16, 2>(nm, PR, fl, tpc);
    }
    printf("\n== whole waves (7104 tiles), layout x stages, K = 512 and 2048, beta = 1\n");
    for (int K : {512, 2048}) {
      double fl = 2.0 * 64 * 64 * (double)K * 7104;
      char nm[64];
      auto PR = mk(64, 96, 74, K, KM_FULL, ORD_COL, 0, -1.0, 1.0);
      snprintf(nm, sizeof nm, "NT s2 K=%d", K); run<64, 2, 2, false, false, 16, 2>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT s3 K=%d", K); run<64, 2, 2, false, false, 16, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT s4 K=%d", K); run<64, 2, 2, false, false, 16, 4>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN s2 K=%d", K); run<64, 2, 2, true, true, 16, 2>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN s3 K=%d", K); run<64, 2, 2, true, true, 16, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NN s2 K=%d", K); run<64, 2, 2, false, true, 16, 2>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NN s3 K=%d", K); run<64, 2, 2, false, true, 16, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN(A k-major, B m-major) s2 K=%d", K); run<64, 2, 2, true, false, 16, 2>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN(A k-major, B m-major) s3 K=%d", K); run<64, 2, 2, true, false, 16, 3>(nm, PR, fl);
    }
    printf("\n== whole waves (7104 tiles): gemm_kernel with 3 stages (NT / NN) or 2 (TN) vs gemm_kernel_mt with one tile per CTA and 2 stages\n");
    for (int K : {64, 128, 256, 512, 1024}) {
      double fl = 2.0 * 64 * 64 * (double)K * 7104;
      char nm[64];
      auto PR = mk(64, 96, 74, K, KM_FULL, ORD_COL, 0, -1.0, 1.0);
      snprintf(nm, sizeof nm, "NT s3 K=%d", K); run<64, 2, 2, false, false, 16, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT s2 K=%d mt", K); run_mt<64, 2, 2, false, false, 16, 2>(nm, PR, fl, 1);
      snprintf(nm, sizeof nm, "NN s3 K=%d", K); run<64, 2, 2, false, true, 16, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NN s2 K=%d mt", K); run_mt<64, 2, 2, false, true, 16, 2>(nm, PR, fl, 1);
      snprintf(nm, sizeof nm, "TN s2 K=%d", K); run<64, 2, 2, true, true, 16, 2>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN s2 K=%d mt", K); run_mt<64, 2, 2, true, true, 16, 2>(nm, PR, fl, 1);
    }
    printf("\n== 8880 tiles = 15 waves of 592 = 12 waves of 740: register cap for 5 CTAs/SM (__launch_bounds__(128, 5)), 2 stages\n");
    for (int K : {128, 512, 2048}) {
      double fl = 2.0 * 64 * 64 * (double)K * 8880;
      char nm[64];
      auto PR = mk(64, 120, 74, K, KM_FULL, ORD_COL, 0, -1.0, 1.0);
      snprintf(nm, sizeof nm, "NT K=%d minb3", K); run<64, 2, 2, false, false, 16, 2, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT K=%d minb4", K); run<64, 2, 2, false, false, 16, 2, 4>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT K=%d minb5", K); run<64, 2, 2, false, false, 16, 2, 5>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NN K=%d minb3", K); run<64, 2, 2, false, true, 16, 2, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NN K=%d minb4", K); run<64, 2, 2, false, true, 16, 2, 4>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NN K=%d minb5", K); run<64, 2, 2, false, true, 16, 2, 5>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN K=%d minb3", K); run<64, 2, 2, true, true, 16, 2, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN K=%d minb4", K); run<64, 2, 2, true, true, 16, 2, 4>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN K=%d minb5", K); run<64, 2, 2, true, true, 16, 2, 5>(nm, PR, fl);
    }
    printf("\n== 8880 tiles, 8 warps per 64 x 64 CTA (32 x 16 warp tiles), 2 stages\n");
    for (int K : {128, 512, 2048}) {
      double fl = 2.0 * 64 * 64 * (double)K * 8880;
      char nm[64];
      auto PR = mk(64, 120, 74, K, KM_FULL, ORD_COL, 0, -1.0, 1.0);
      snprintf(nm, sizeof nm, "NT K=%d 2x4 minb2", K); run<64, 2, 4, false, false, 16, 2, 2>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT K=%d 2x4 minb3", K); run<64, 2, 4, false, false, 16, 2, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT K=%d 2x4 minb4", K); run<64, 2, 4, false, false, 16, 2, 4>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT K=%d 4x2 minb3", K); run<64, 4, 2, false, false, 16, 2, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN K=%d 2x4 minb3", K); run<64, 2, 4, true, true, 16, 2, 3>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN K=%d 2x4 minb4", K); run<64, 2, 4, true, true, 16, 2, 4>(nm, PR, fl);
      snprintf(nm, sizeof nm, "NT K=%d 2x2 kc32 minb4", K); run<64, 2, 2, false, false, 32, 2, 4>(nm, PR, fl);
      snprintf(nm, sizeof nm, "TN K=%d 2x2 kc32 minb4", K); run<64, 2, 2, true, true, 32, 2, 4>(nm, PR, fl);
    }
    // results: the same launch by both kernels must agree bitwise
    { auto P1 = mk(64, 120, 120, 512, KM_FULL, ORD_LOWER, 1, -1.0, 1.0);
      std::vector<double> h0((size_t)n * n), h1((size_t)n * n);
      auto k0 = gemm_kernel<64, 2, 2, false, false, 16, 3>; auto k1 = gemm_kernel_mt<64, 2, 2, false, false, 16, 3>;
      constexpr int smem = gemm_smem_bytes<64, false, false, 16, 3>();
      CK(cudaMemcpy(dP, &P1, sizeof P1, cudaMemcpyHostToDevice));
      fill<<<(n * (size_t)n + 255) / 256, 256>>>(C, (size_t)n * n, 3);
      k0<<<P1.ntiles, 128, smem>>>(dP, 1, nullptr); CK(cudaMemcpy(h0.data(), C, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
      fill<<<(n * (size_t)n + 255) / 256, 256>>>(C, (size_t)n * n, 3);
      k1<<<(P1.ntiles + 4) / 5, 128, smem>>>(dP, 1, nullptr, 5, P1.ntiles); CK(cudaMemcpy(h1.data(), C, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
      size_t bad = 0; for (size_t i = 0; i < h0.size(); ++i) bad += h0[i] != h1[i];
      printf("bitwise comparison gemm_kernel vs gemm_kernel_mt (tpc 5, mirror): %zu differing entries of %zu\n", bad, h0.size());
    }
    return 0;
  }
  printf("\n== TN  K^-1 = X^T X  (n=8192, k >= i, lower + mirror): n^3/3 flops\n");
  { double fl = (double)n * n * n / 3.0;
    auto MK = [&](int T) { return mk(T, n / T, n / T, n, KM_GE_I, ORD_LOWER, 1, 1.0, 0.0); };
    ALLCFG(true, true, MK, fl, "TN") }
  printf("\n== NT  trailing update k=512, 7680 rows lower (beta=1)\n");
  { double fl = 2.0 * 128 * 128 * 512.0 * lower_tiles(60);
    auto MK = [&](int T) { return mk(T, 7680 / T, 7680 / T, 512, KM_FULL, ORD_LOWER, 0, -1.0, 1.0); };
    ALLCFG(false, false, MK, fl, "NT512") }
  printf("\n== layout vs K: the same 64 x 64 kernel, 4096 rows lower, K = 4096 and K = 512, NT / TN / NN, beta = 1 and beta = 0\n");
  for (int K : {4096, 512}) {
    double fl = 2.0 * 64 * 64 * (double)K * lower_tiles(4096 / 64);
    char nm[64];
    for (int beta = 1; beta >= 0; --beta) {
      auto MKb = [&](int T) { return mk(T, 4096 / T, 4096 / T, K, KM_FULL, ORD_LOWER, 0, -1.0, (double)beta); };
      snprintf(nm, sizeof nm, "NT K=%d beta=%d s3", K, beta); run<64, 2, 2, false, false, 16, 3>(nm, MKb(64), fl);
      snprintf(nm, sizeof nm, "NT K=%d beta=%d s2", K, beta); run<64, 2, 2, false, false, 16, 2>(nm, MKb(64), fl);
      snprintf(nm, sizeof nm, "TN K=%d beta=%d s2", K, beta); run<64, 2, 2, true, true, 16, 2>(nm, MKb(64), fl);
      snprintf(nm, sizeof nm, "TN K=%d beta=%d s3", K, beta); run<64, 2, 2, true, true, 16, 3>(nm, MKb(64), fl);
      snprintf(nm, sizeof nm, "NN K=%d beta=%d s3", K, beta); run<64, 2, 2, false, true, 16, 3>(nm, MKb(64), fl);
    }
  }
  printf("\n== NT  inner update k=128, 7680 x 384 rectangle (beta=1)\n");
  { double fl = 2.0 * 7680.0 * 384 * 128;
    auto MK = [&](int T) { return mk(T, 7680 / T, 384 / T, 128, KM_FULL, ORD_COL, 0, -1.0, 1.0); };
    ALLCFG(false, false, MK, fl, "NT128") }
  printf("\n== NN  4096^3 (trtri top level shape, k <= i)\n");
  { double fl = 2.0 * 4096.0 * 4096 * 4096 / 2;
    auto MK = [&](int T) { return mk(T, 4096 / T, 4096 / T, 4096, KM_LE_I, ORD_ROW_DESC, 0, -1.0, 0.0); };
    ALLCFG(false, true, MK, fl, "NN") }
  printf("\n== NN  512^3 x 8 problems is not grouped here; single 512^3 (k <= i)\n");
  { double fl = 2.0 * 512.0 * 512 * 512 / 2;
    auto MK = [&](int T) { return mk(T, 512 / T, 512 / T, 512, KM_LE_I, ORD_ROW_DESC, 0, -1.0, 0.0); };
    ALLCFG(false, true, MK, fl, "NN512") }
  return 0;
}
