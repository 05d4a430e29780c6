// Runtime shim: the product build (nvcc, sm_100a) talks to the CUDA runtime directly.
// With -DCS_EMU (g++, tests/emu only) the same sources run on the SIMT emulator so the
// kernels' index logic can be unit-tested without a GPU.  There is no CPU fallback in
// the product library: CS_EMU is never defined by causalstump_b200/csrc/build.py.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef CS_EMU
#include "emu.h"
typedef int cudaStream_t;
typedef int cudaEvent_t;
typedef int cudaError_t;
#define cudaSuccess 0
#define CS_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::dyn_smem)
#define CS_LAUNCH(kfn, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kfn(__VA_ARGS__); })
#define CS_SET_SMEM(kfn, bytes) ((void)0)
#define __grid_constant__
namespace rt {
inline int check(int e, const char*, int) { return e; }
inline int dev_malloc(void** p, size_t bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? 0 : 2; }
inline int dev_free(void* p) { std::free(p); return 0; }
inline int h2d(void* d, const void* h, size_t b, cudaStream_t) { std::memcpy(d, h, b); return 0; }
inline int d2h(void* h, const void* d, size_t b, cudaStream_t) { std::memcpy(h, d, b); return 0; }
inline int d2d(void* d, const void* s, size_t b, cudaStream_t) { std::memmove(d, s, b); return 0; }
inline int memset0(void* d, size_t b, cudaStream_t) { std::memset(d, 0, b); return 0; }
inline int stream_create(cudaStream_t* s) { *s = 0; return 0; }
inline int stream_create_low(cudaStream_t* s) { *s = 0; return 0; }
inline int event_create_fast(cudaEvent_t* e) { *e = 0; return 0; }
inline int stream_destroy(cudaStream_t) { return 0; }
inline int stream_sync(cudaStream_t) { return 0; }
inline int event_create(cudaEvent_t* e) { *e = 0; return 0; }
inline int event_destroy(cudaEvent_t) { return 0; }
inline int event_record(cudaEvent_t, cudaStream_t) { return 0; }
inline int event_sync(cudaEvent_t) { return 0; }
inline int stream_wait_event(cudaStream_t, cudaEvent_t) { return 0; }
inline int event_elapsed(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
inline int last_error() { return 0; }
inline const char* err_str(int) { return "emu"; }
inline int device_count(int* n) { *n = 1; return 0; }
inline int set_device(int) { return 0; }
inline int sm_count() { return 4; }
inline int host_alloc(void** p, size_t b) { *p = std::malloc(b ? b : 1); return *p ? 0 : 2; }
inline int host_free(void* p) { std::free(p); return 0; }
inline int partition_streams(int, cudaStream_t*, cudaStream_t*, cudaStream_t*, cudaStream_t**, const int*, int, int*, int*) { return 1; }
inline int copy2d(double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols, cudaStream_t) {
  for (long long c = 0; c < cols; ++c) std::memmove(dst + c * ldd, src + c * lds, sizeof(double) * rows);
  return 0;
}
inline int stream_create_level(cudaStream_t* s, int) { *s = 0; return 0; }
struct XTma { int on; };  // no TMA on the emulator (see cs_dev.cuh)
inline int make_x_tensormap(XTma* out, const double*, long long, int, int, long long) { out->on = 0; return 1; }
typedef int graph_exec_t;
inline int graph_begin(cudaStream_t) { return 1; }  // no graphs on the emulator: eager launches
inline int graph_end(cudaStream_t, graph_exec_t*) { return 1; }
inline int graph_launch(graph_exec_t, cudaStream_t) { return 1; }
inline int graph_destroy(graph_exec_t) { return 0; }
typedef int graph_cond_t;
template <class Body> inline int graph_while_build(cudaStream_t, Body, graph_exec_t*) { return 1; }
}  // namespace rt
#else
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>
#define CS_DYN_SMEM(T, name) \
  extern __shared__ __align__(128) unsigned char cs_dyn_smem_raw[]; \
  T* name = reinterpret_cast<T*>(cs_dyn_smem_raw)
#define CS_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CS_SET_SMEM(kfn, bytes) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
namespace rt {
inline int dev_malloc(void** p, size_t bytes) { return (int)cudaMalloc(p, bytes ? bytes : 1); }
inline int dev_free(void* p) { return (int)cudaFree(p); }
inline int h2d(void* d, const void* h, size_t b, cudaStream_t s) { return (int)cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, s); }
inline int d2h(void* h, const void* d, size_t b, cudaStream_t s) { return (int)cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, s); }
inline int d2d(void* d, const void* s_, size_t b, cudaStream_t s) { return (int)cudaMemcpyAsync(d, s_, b, cudaMemcpyDeviceToDevice, s); }
inline int memset0(void* d, size_t b, cudaStream_t s) { return (int)cudaMemsetAsync(d, 0, b, s); }
// main streams carry the latency-bound critical path: highest priority; auxiliary streams
// (bulk trailing updates) lowest, so critical-path CTAs take the next free SM slot.
inline int stream_create(cudaStream_t* s) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  return (int)cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, hi);
}
inline int stream_create_low(cudaStream_t* s) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  return (int)cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, lo);
}
// level 0 = highest priority, larger = lower (clamped to the device's range)
inline int prio_of_level(int level) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi is numerically smallest
  int p = hi + level;
  return p > lo ? lo : p;
}
inline int stream_create_level(cudaStream_t* s, int level) {
  return (int)cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, prio_of_level(level));
}
inline int event_create_fast(cudaEvent_t* e) { return (int)cudaEventCreateWithFlags(e, cudaEventDisableTiming); }
inline int stream_destroy(cudaStream_t s) { return (int)cudaStreamDestroy(s); }
inline int stream_sync(cudaStream_t s) { return (int)cudaStreamSynchronize(s); }
inline int event_create(cudaEvent_t* e) { return (int)cudaEventCreate(e); }
inline int event_destroy(cudaEvent_t e) { return (int)cudaEventDestroy(e); }
inline int event_record(cudaEvent_t e, cudaStream_t s) { return (int)cudaEventRecord(e, s); }
inline int event_sync(cudaEvent_t e) { return (int)cudaEventSynchronize(e); }
inline int stream_wait_event(cudaStream_t s, cudaEvent_t e) { return (int)cudaStreamWaitEvent(s, e, 0); }
inline int event_elapsed(float* ms, cudaEvent_t a, cudaEvent_t b) { return (int)cudaEventElapsedTime(ms, a, b); }
inline int last_error() { return (int)cudaGetLastError(); }
inline const char* err_str(int e) { return cudaGetErrorString((cudaError_t)e); }
inline int device_count(int* n) { return (int)cudaGetDeviceCount(n); }
inline int set_device(int d) { return (int)cudaSetDevice(d); }
inline int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}
inline int host_alloc(void** p, size_t b) { return (int)cudaMallocHost(p, b ? b : 1); }
inline int host_free(void* p) { return (int)cudaFreeHost(p); }

// SM partition (CUDA green contexts, driver API resolved through cudaGetDriverEntryPoint so the
// library still links only cudart).  The blocked Cholesky has a latency-bound chain of one-CTA
// kernels (diagonal-block factorizations) that must never queue behind the bulk GEMM CTAs of the
// trailing update / pipelined inverse: stream priorities are not enough, because a register-heavy
// CTA only starts once enough co-resident bulk CTAs of ONE SM have drained.  So the chain gets a
// small partition A (8 SMs, the hardware minimum on sm_90+) of its own and the bulk streams run on
// partition B (the remaining SMs).  One pair of green contexts per device, created lazily, shared
// by every handle; returns non-zero when the driver cannot partition (callers fall back to plain
// priority streams).
struct PartitionCtx { int state = 0; CUgreenCtx a = nullptr, b = nullptr; int sm_a = 0, sm_b = 0; };
inline PartitionCtx& partition_ctx(int dev) { static PartitionCtx tab[64]; return tab[dev & 63]; }
template <class F> inline bool drv_sym(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}
inline int copy2d(double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols, cudaStream_t s) {
  return (int)cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds, sizeof(double) * rows, (size_t)cols,
                                cudaMemcpyDeviceToDevice, s);
}
inline int partition_streams(int sm_crit, cudaStream_t* crit_hi, cudaStream_t* crit_hi2, cudaStream_t* crit_hi3, cudaStream_t** bulk,
                             const int* level, int nbulk, int* sm_a, int* sm_b) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  static std::mutex mu;  // handles may be created from several host threads
  std::lock_guard<std::mutex> lock(mu);
  PartitionCtx& P = partition_ctx(dev);
  CUresult (*f_devget)(CUdevice*, int) = nullptr;
  CUresult (*f_getres)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*f_split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*f_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*f_gcreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*f_gstream)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  if (!drv_sym("cuGreenCtxStreamCreate", &f_gstream)) return 1;
  if (P.state == 0) {
    P.state = -1;
    if (!drv_sym("cuDeviceGet", &f_devget) || !drv_sym("cuDeviceGetDevResource", &f_getres) ||
        !drv_sym("cuDevSmResourceSplitByCount", &f_split) || !drv_sym("cuDevResourceGenerateDesc", &f_desc) ||
        !drv_sym("cuGreenCtxCreate", &f_gcreate))
      return 1;
    cudaFree(0);  // make sure the primary context exists
    CUdevice cd;
    CUdevResource all{}, ga{}, gb{};
    unsigned int nb = 1;
    CUdevResourceDesc da = nullptr, db = nullptr;
    if (f_devget(&cd, dev) != CUDA_SUCCESS) return 1;
    if (f_getres(cd, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return 1;
    if (f_split(&ga, &nb, &all, &gb, 0, (unsigned int)sm_crit) != CUDA_SUCCESS || nb != 1) return 1;
    if (ga.sm.smCount == 0 || gb.sm.smCount == 0) return 1;
    if (f_desc(&da, &ga, 1) != CUDA_SUCCESS || f_desc(&db, &gb, 1) != CUDA_SUCCESS) return 1;
    if (f_gcreate(&P.a, da, cd, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return 1;
    if (f_gcreate(&P.b, db, cd, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return 1;
    P.sm_a = (int)ga.sm.smCount; P.sm_b = (int)gb.sm.smCount;
    P.state = 1;
  }
  if (P.state != 1) return 1;
  CUstream s0 = nullptr;
  if (f_gstream(&s0, P.a, CU_STREAM_NON_BLOCKING, prio_of_level(0)) != CUDA_SUCCESS) return 1;
  *crit_hi = (cudaStream_t)s0;
  CUstream s0b = nullptr;
  if (f_gstream(&s0b, P.a, CU_STREAM_NON_BLOCKING, prio_of_level(0)) != CUDA_SUCCESS) return 1;
  *crit_hi2 = (cudaStream_t)s0b;
  CUstream s0c = nullptr;
  if (f_gstream(&s0c, P.a, CU_STREAM_NON_BLOCKING, prio_of_level(0)) != CUDA_SUCCESS) return 1;
  *crit_hi3 = (cudaStream_t)s0c;
  for (int i = 0; i < nbulk; ++i) {
    CUstream sb = nullptr;
    if (f_gstream(&sb, P.b, CU_STREAM_NON_BLOCKING, prio_of_level(level[i])) != CUDA_SUCCESS) return 1;
    *bulk[i] = (cudaStream_t)sb;
  }
  if (sm_a) *sm_a = P.sm_a;
  if (sm_b) *sm_b = P.sm_b;
  return 0;
}
// TMA descriptor of a covariate matrix X (np rows x p columns, column-major, one per member of a batched handle `slab`
// bytes apart): a 3-D tensor {np, p, members} with box {64, p, 1}, so ONE cp.async.bulk.tensor instruction brings the
// [p][64] X tile of a 64-row block into shared memory in exactly the layout the element-wise kernels index (xs[i*64 + r]).
// Passed to the kernels by value as a __grid_constant__ parameter.  on == 0: the kernels use their plain copy loops.
struct XTma { CUtensorMap map; int on; };
inline int make_x_tensormap(XTma* out, const double* X, long long np, int p, int members, long long slab_bytes) {
  out->on = 0;
  if (p < 1 || p > 256 || np % 64 != 0 || (reinterpret_cast<unsigned long long>(X) & 15ull) || std::getenv("CS_NO_TMA")) return 1;
  CUresult (*f_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
  if (!drv_sym("cuTensorMapEncodeTiled", &f_enc)) return 1;
  const cuuint64_t dims[3] = {(cuuint64_t)np, (cuuint64_t)p, (cuuint64_t)(members > 0 ? members : 1)};
  const cuuint64_t col = (cuuint64_t)np * sizeof(double);
  const cuuint64_t strides[2] = {col, members > 1 ? (cuuint64_t)slab_bytes : col * (cuuint64_t)p};
  if (strides[1] % 16 != 0) return 1;
  const cuuint32_t box[3] = {64u, (cuuint32_t)p, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (f_enc(&out->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  out->on = 1;
  return 0;
}
// CUDA graphs: one optimizer iteration (all its launches, both streams) is captured once
// and replayed, removing per-launch CPU cost from latency-bound small-n fits.
typedef cudaGraphExec_t graph_exec_t;
inline int graph_begin(cudaStream_t s) { return (int)cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal); }
inline int graph_end(cudaStream_t s, graph_exec_t* exec) {
  cudaGraph_t g = nullptr;
  int e = (int)cudaStreamEndCapture(s, &g);
  if (e != 0 || g == nullptr) { cudaGetLastError(); return e ? e : 1; }
  e = (int)cudaGraphInstantiate(exec, g, 0);
  cudaGraphDestroy(g);
  if (e != 0) cudaGetLastError();
  return e;
}
inline int graph_launch(graph_exec_t exec, cudaStream_t s) { return (int)cudaGraphLaunch(exec, s); }
inline int graph_destroy(graph_exec_t exec) { return (int)cudaGraphExecDestroy(exec); }
// Device-side loop: a graph whose only node is a conditional WHILE node (CUDA 12.4+); `body(handle)` enqueues one
// iteration on `s` (captured into the body graph) and must end with a kernel that calls cudaGraphSetConditional(handle, go_on).
// The condition starts at 1 on every launch.  All kernels of the body must belong to one context (no green-context streams).
typedef cudaGraphConditionalHandle graph_cond_t;
template <class Body>
inline int graph_while_build(cudaStream_t s, Body body, graph_exec_t* exec) {
  cudaGraph_t g = nullptr;
  if (cudaGraphCreate(&g, 0) != cudaSuccess) { cudaGetLastError(); return 1; }
  cudaGraphConditionalHandle h;
  cudaGraphNode_t node;
  cudaGraphNodeParams np{};
  int e = (int)cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  if (e == 0) {
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = h;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    e = (int)cudaGraphAddNode(&node, g, nullptr, 0, &np);
  }
  if (e == 0) {
    cudaGraph_t bodyg = np.conditional.phGraph_out[0];
    e = (int)cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e == 0) {
      body(h);
      cudaGraph_t got = nullptr;
      e = (int)cudaStreamEndCapture(s, &got);
    }
  }
  if (e == 0) e = (int)cudaGraphInstantiate(exec, g, 0);
  cudaGraphDestroy(g);
  if (e != 0) cudaGetLastError();
  return e;
}
}  // namespace rt
#endif
