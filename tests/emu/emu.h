// Minimal SIMT emulator: TEST INFRASTRUCTURE ONLY (never shipped, never loaded by the
// causalstump_b200 package).  It lets the *same* kernel sources under
// causalstump_b200/csrc/ be compiled with g++ (-DCS_EMU) so that indexing, fragment
// layouts, barriers and the host-side launch plans can be unit-tested in the GPU-less
// build container before GPU minutes are spent.  Every CUDA thread is a user-level
// fiber; a thread block runs to completion before the next one starts; barriers and
// warp collectives are cooperative yields.  Nothing here is a CPU fallback of the
// product: libcausalstump_b200.so is built by nvcc only and fails loudly without a GPU.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

struct emu_uint3 { unsigned x, y, z; };
struct alignas(16) double2 { double x, y; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

extern emu_uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

namespace emu {
extern unsigned char* dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& fn);
void barrier_block();
void barrier_warp();
int lane();
// per-warp exchange slot for collectives (double-buffered by per-lane op parity)
void* xch_slot(int parity, int lane);
int next_parity();
extern long n_launches;
}  // namespace emu

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __ldg(p) (*(p))

inline void __syncthreads() { emu::barrier_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier_warp(); }
inline void __threadfence() {}

template <class T>
inline T emu_shfl_idx(T v, int src) {
  static_assert(sizeof(T) <= 16, "shfl payload too large");
  int par = emu::next_parity();
  std::memcpy(emu::xch_slot(par, emu::lane()), &v, sizeof(T));
  emu::barrier_warp();
  T r;
  std::memcpy(&r, emu::xch_slot(par, src & 31), sizeof(T));
  return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_idx(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl_idx(v, emu::lane() ^ m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) {
  int s = emu::lane() + d;
  return emu_shfl_idx(v, s > 31 ? emu::lane() : s);
}

inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
inline int atomicCAS(int* p, int c, int v) { int o = *p; if (o == c) *p = v; return o; }

inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __dsqrt_rn(double x) { return std::sqrt(x); }
using std::exp; using std::log; using std::sqrt; using std::fma; using std::lgamma; using std::fabs;
