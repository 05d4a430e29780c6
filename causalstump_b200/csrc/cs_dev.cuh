// Device primitives for sm_100a: FP64 tensor-core MMA (DMMA.8x8x4), cp.async staging,
// warp reductions.  Under CS_EMU (tests only) each primitive has a behavioural model.
#pragma once
#include <type_traits>

#include "cs_rt.h"

namespace csd {

// D(8x8) += A(8x4,row) * B(4x8,col), fp64.  Fragment ownership (PTX ISA, mma.m8n8k4.f64):
//   a : A[lane>>2][lane&3]        b : B[lane&3][lane>>2]
//   c[i] : C[lane>>2][2*(lane&3)+i]
__device__ __forceinline__ void dmma8x8x4(double (&c)[2], double a, double b) {
#ifdef CS_EMU
  int par = emu::next_parity();
  double* slot = static_cast<double*>(emu::xch_slot(par, emu::lane()));
  slot[0] = a; slot[1] = b;
  emu::barrier_warp();
  const int g = emu::lane() >> 2, t = emu::lane() & 3;
  for (int i = 0; i < 2; ++i) {
    const int col = 2 * t + i;
    double s = c[i];
    for (int k = 0; k < 4; ++k) {
      const double av = static_cast<double*>(emu::xch_slot(par, g * 4 + k))[0];
      const double bv = static_cast<double*>(emu::xch_slot(par, col * 4 + k))[1];
      s = std::fma(av, bv, s);
    }
    c[i] = s;
  }
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
#endif
}

// 16-byte global -> shared asynchronous copy (LDGSTS).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef CS_EMU
  std::memcpy(smem_dst, gmem_src, 16);
#else
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
#endif
}
// exp(x) through a 64-entry table of 2^(j/64) (shared memory, filled by exp_tab_fill) and a degree-5 polynomial on
// |r| <= ln2/128:  x = k ln2/64 + r,  exp(x) = 2^(k>>6) * T[k&63] * p(r).  10 FP64 instructions instead of the ~17 of the
// library exp (degree-11 polynomial) -- the two exponentials per matrix entry are a third of all instructions of the Gram
// and trace-gradient kernels.  Maximum relative error 4e-16 on [-700, 5] (checked against 60-digit arithmetic,
// tests/test_oracle.py::test_table_exp); arguments below -700 give 0, the caller's arguments never exceed a few units.
constexpr int EXP_TAB = 64;
__device__ __forceinline__ void exp_tab_fill(double* T, int tid) {
  if (tid < EXP_TAB) T[tid] = exp2((double)tid * (1.0 / EXP_TAB));
}
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ T) {
  const double magic = 6755399441055744.0;   // 1.5 * 2^52: the rounded integer sits in the low word
  const double t = fma(x, 92.33248261689366, magic);
  const double kd = t - magic;
  double r = fma(kd, -0.01083042469326756, x);      // ln2/64, high part (32 significant bits: kd * hi is exact)
  r = fma(kd, -2.9815858269852933e-12, r);          // low part
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
#ifdef CS_EMU
  long long tb; std::memcpy(&tb, &t, 8);
  const int k = (int)(unsigned)(tb & 0xffffffffLL);
  double v = T[k & (EXP_TAB - 1)] * p;
  long long vb; std::memcpy(&vb, &v, 8);
  vb += (long long)(k >> 6) << 52;
  std::memcpy(&v, &vb, 8);
#else
  const int k = __double2loint(t);
  double v = T[k & (EXP_TAB - 1)] * p;
  v = __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
#endif
  return x < -700.0 ? 0.0 : v;
}

// L2 prefetch of the 128-byte line holding p
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef CS_EMU
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
#else
  (void)p;
#endif
}
// 8-byte variant (element-granular staging of a packed triangle)
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
#ifdef CS_EMU
  std::memcpy(smem_dst, gmem_src, 8);
#else
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef CS_EMU
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef CS_EMU
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
#endif
}

// Streaming 8-byte load whose position in the instruction stream is kept (asm volatile): lets a loop put a batch of
// independent loads in flight before the first use, where the scheduler would otherwise pair each load with its FMA.
__device__ __forceinline__ double ld_stream(const double* p) {
#ifdef CS_EMU
  return *p;
#else
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#endif
}

__device__ __forceinline__ double2 ld_stream2(const double* p) {  // 16-byte aligned pair
#ifdef CS_EMU
  double2 v; v.x = p[0]; v.y = p[1];
  return v;
#else
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#endif
}

// ---- TMA (cp.async.bulk.tensor, SASS UTMALDG) + mbarrier: one elected thread stages a whole operand tile into shared
// memory with a single instruction; the consumers wait on the transaction barrier.  Product build only: the SIMT
// emulator keeps the plain copy loops (its thread blocks run to completion one after another, nothing to overlap).
#ifndef CS_EMU
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// generic-proxy reads of a shared-memory buffer are ordered before the async-proxy (TMA) writes that refill it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, unsigned long long* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
#endif

// Batched launches (cs_fit_create_batch): every per-fit device buffer sits at the same offset of a
// per-fit slab, so one launch serves B fits with blockIdx.z selecting the slab (bstride bytes apart).
template <class T>
__device__ __forceinline__ T* bshift(T* p, long long bstride) {
  if (bstride == 0 || p == nullptr) return p;
  return reinterpret_cast<T*>(reinterpret_cast<char*>(const_cast<typename std::remove_const<T>::type*>(p)) +
                              (long long)blockIdx.z * bstride);
}

// reciprocal without the slow-path subroutine call of IEEE division
__device__ __forceinline__ double fast_rcp(double x) {
#ifdef CS_EMU
  return 1.0 / x;
#else
  return __drcp_rn(x);
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum; result valid in thread 0 (and returned to all after sync).
// `red` must hold >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = (lane < nw) ? red[lane] : 0.0;
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

}  // namespace csd
