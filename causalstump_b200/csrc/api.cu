// C-ABI of libcausalstump_b200.so (declared in include/cs_b200.h): host orchestration of
// the device-resident fit.  One translation unit; compiled by nvcc for sm_100a only
// (tests/emu compiles the same file with -DCS_EMU for index-logic unit tests).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <string>
#include <vector>

#include "../../include/cs_b200.h"
#include "dense.cuh"
#include "kernels.cuh"
#include "kernels_mma.cuh"
#include "predict.cuh"
#include "dist.cuh"
#include "sim.cuh"
#ifndef CS_EMU
#include <dlfcn.h>
#endif

using csdense::DenseEngine;
using csdense::g_launches;
using csk::EvalScalars;
using csk::Layout;
using csk::LoopState;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define RT(call)                                                                              \
  do {                                                                                        \
    int e__ = (call);                                                                         \
    if (e__ != 0) return fail(e__ == 2 ? CS_ERR_NOMEM : CS_ERR_CUDA, "%s failed: %s (%s:%d)", \
                              #call, rt::err_str(e__), __FILE__, __LINE__);                   \
  } while (0)

#define KCHECK()                                                                                      \
  do {                                                                                                \
    int e__ = rt::last_error();                                                                       \
    if (e__ != 0) return fail(CS_ERR_CUDA, "kernel launch failed: %s (%s:%d)", rt::err_str(e__),     \
                              __FILE__, __LINE__);                                                    \
  } while (0)

static int need_device() {
  int n = 0;
  if (rt::device_count(&n) != 0 || n <= 0)
    return fail(CS_ERR_NODEVICE, "no CUDA device available: libcausalstump_b200 has no CPU fallback");
  return 0;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------
static __global__ void fill_kernel(double* p, long long n, double v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// identity padding of an np x np matrix whose leading n x n block holds data
static __global__ void pad_identity_kernel(double* A, long long ld, int n, int np) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)np * np;
  if (idx >= total) return;
  const int r = (int)(idx % np), c = (int)(idx / np);
  if (r >= n || c >= n) A[(long long)c * ld + r] = (r == c) ? 1.0 : 0.0;
}

static __global__ void add_noise_diag_kernel(double* A, long long ld, int n, const double* w, double esig) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < n) A[(long long)i * ld + i] += esig / w[i];
}

// Scratch arena for the predict / treatment calls: chunks are kept for the life of the handle and handed
// out again on the next call, so a replicate loop of 3 calls per fit does not pay cudaMalloc / cudaFree
// (which also synchronises the device) every time.  Active only inside an ArenaScope.
struct Arena {
  struct Chunk { char* base; size_t cap, off; };
  std::vector<Chunk> chunks;
  void reset() { for (auto& c : chunks) c.off = 0; }
  void* take(size_t bytes) {
    bytes = (bytes + 255) / 256 * 256;
    if (bytes == 0) bytes = 256;
    for (auto& c : chunks)
      if (c.cap - c.off >= bytes) { void* q = c.base + c.off; c.off += bytes; return q; }
    Chunk c{nullptr, std::max(bytes, (size_t)8 << 20), 0};
    if (rt::dev_malloc((void**)&c.base, c.cap)) return nullptr;
    c.off = bytes;
    chunks.push_back(c);
    return c.base;
  }
  void release() { for (auto& c : chunks) rt::dev_free(c.base); chunks.clear(); }
  // keep at most `cap` bytes for the next call (largest chunks go first): a replicate loop at n ~ 700 keeps its few
  // MB, one predict at n = 8192 (1.5 GB of scratch) does not stay resident for the life of the host thread
  void trim(size_t cap) {
    size_t total = 0;
    for (auto& c : chunks) total += c.cap;
    while (total > cap && !chunks.empty()) {
      size_t big = 0;
      for (size_t i = 1; i < chunks.size(); ++i) if (chunks[i].cap > chunks[big].cap) big = i;
      total -= chunks[big].cap;
      rt::dev_free(chunks[big].base);
      chunks.erase(chunks.begin() + big);
    }
  }
  ~Arena() { release(); }  // thread exit: the scratch of this host thread goes back to the device
};
static constexpr size_t ARENA_KEEP_BYTES = (size_t)256 << 20;
static thread_local Arena* g_arena = nullptr;
// one arena per host thread and device, shared by every handle: replicate loops create and destroy handles,
// the scratch survives them (predict / treatment calls are synchronous, so it is idle between calls)
static Arena* thread_arena() {
  static thread_local Arena pool[16];
  int dev = 0;
#ifndef CS_EMU
  cudaGetDevice(&dev);
#endif
  return &pool[dev & 15];
}
struct ArenaScope {
  Arena* prev;
  explicit ArenaScope(Arena* a) : prev(g_arena) { g_arena = a; if (a) a->reset(); }
  ~ArenaScope() { if (g_arena) g_arena->trim(ARENA_KEEP_BYTES); g_arena = prev; }
};

struct DevBuf {
  void* p = nullptr;
  bool owned = true;
  ~DevBuf() { if (p && owned) rt::dev_free(p); }
  int alloc(size_t bytes) {
    if (g_arena) { owned = false; p = g_arena->take(bytes); return p ? 0 : 2; }
    return rt::dev_malloc(&p, bytes);
  }
  double* d() const { return static_cast<double*>(p); }
};

// host (ld = rows) -> device (ld = ldd) column-major upload with zero padding
static int upload_matrix(double* dst, long long ldd, int rows_pad, int cols_pad, const double* src, int rows,
                         int cols, cudaStream_t st) {
  int e;
  if ((e = rt::memset0(dst, sizeof(double) * (size_t)ldd * cols_pad, st))) return e;
  (void)rows_pad;
#ifdef CS_EMU
  for (int c = 0; c < cols; ++c) std::memcpy(dst + (size_t)c * ldd, src + (size_t)c * rows, sizeof(double) * rows);
  return 0;
#else
  return (int)cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * rows, sizeof(double) * rows, cols,
                                cudaMemcpyHostToDevice, st);
#endif
}

// same, source with its own leading dimension, on the host OR on the device (cudaMemcpyDefault: unified addressing)
static int stage_matrix(double* dst, long long ldd, int cols_pad, const double* src, long long lds, int rows, int cols, cudaStream_t st) {
  int e;
  if ((e = rt::memset0(dst, sizeof(double) * (size_t)ldd * cols_pad, st))) return e;
#ifdef CS_EMU
  for (int c = 0; c < cols; ++c) std::memcpy(dst + (size_t)c * ldd, src + (size_t)c * lds, sizeof(double) * rows);
  return 0;
#else
  return (int)cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds, sizeof(double) * rows, cols, cudaMemcpyDefault, st);
#endif
}
static int copy_any(void* dst, const void* src, size_t bytes, cudaStream_t st) {
#ifdef CS_EMU
  std::memcpy(dst, src, bytes);
  return 0;
#else
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st);
#endif
}

static int download_matrix(double* dst, int rows, int cols, const double* src, long long lds, cudaStream_t st) {
#ifdef CS_EMU
  for (int c = 0; c < cols; ++c) std::memcpy(dst + (size_t)c * rows, src + (size_t)c * lds, sizeof(double) * rows);
  return 0;
#else
  return (int)cudaMemcpy2DAsync(dst, sizeof(double) * rows, src, sizeof(double) * lds, sizeof(double) * rows, cols,
                                cudaMemcpyDeviceToHost, st);
#endif
}

static int upload_vector(double* dst, int npad, const double* src, int n, double padval, cudaStream_t st) {
  int e;
  if (npad > n) {
    auto k = fill_kernel;
    ++g_launches;
    CS_LAUNCH(k, (npad + 255) / 256, 256, 0, st, dst, (long long)npad, padval);
  }
  if ((e = rt::h2d(dst, src, sizeof(double) * n, st))) return e;
  return 0;
}

// ------------------------------------------------------------------------------------
// the fit handle
// ------------------------------------------------------------------------------------
struct HostLoop { LoopState ls; int status; int pad; };

struct cs_fit {
  cs_fit_config cfg{};
  Layout lay{};
  int n = 0, p = 0, np = 0, nblk = 0;  // np: padded to the dense tile; nblk = np / 64
  cudaStream_t st{};
  DenseEngine eng;
  double *X = nullptr, *y = nullptr, *z = nullptr, *w = nullptr, *ones = nullptr;
  double* Xt = nullptr;  // packed covariate tiles of the tensor-path Gram / gradient kernels (csk::pack_x_tiles_kernel), or nullptr
  rt::XTma xt{};  // TMA descriptor of X over all members (kernels stage their covariate tiles with it); on == 0: plain loads
  double *params = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr;
  double *alpha = nullptr, *u = nullptr, *s = nullptr;
  double* mv_part = nullptr;
  int nsplit = 1, cols_per_split = 0;
  double* gpart = nullptr;
  int ncta_grad = 1;
  double* kal_part = nullptr;
  double* rs_part = nullptr;
  int nrs = 1;
  EvalScalars* sc = nullptr;
  LoopState* ls = nullptr;
  double* trace = nullptr;
  int trace_cap = 0;
  double* flush = nullptr;
  size_t flush_bytes = 0;
  double* pinned = nullptr;
  size_t pinned_bytes = 0;
  rt::graph_exec_t iter_graph{};   // one captured optimizer iteration
  bool graph_ok = false, graph_tried = false;
  rt::graph_exec_t loop_graph{};   // the whole optimisation loop as one conditional WHILE graph (no host polling)
  bool loop_ok = false;
  bool loop_running = false;
  long long launches_per_iter = 0;
  HostLoop* h_ls = nullptr;  // pinned: loop state + status read-back, one per member
  double* trace_unused = nullptr;
  int run_maxiter = 0, run_launched = 0, run_chunk = 1;
  cudaEvent_t ev[16]{};
  bool ev_ok = false;
  // Batched handle (cs_fit_create_batch): nbatch independent fits of the same (n, p, kind) share every
  // launch (blockIdx.z = member).  Each per-fit device buffer lives at the same offset of a per-member slab,
  // `bstride` bytes apart (0 for a single fit).  The pointer fields above always address member `cur`
  // (cs_fit_select rebases them); batched launches are enqueued with member 0 selected.
  int nbatch = 1, cur = 0;
  long long bstride = 0, slab_bytes = 0;
  char* slab = nullptr;
  std::vector<std::pair<void**, size_t>> slab_req;
  bool trace_in_slab = false;
  std::vector<int> member_status;  // per member, after cs_fit_run: 0 OK, > 0 failing pivot, < 0 CS_ERR_*
  // cs_fit_dense_timing: event pairs around the dense phase (Cholesky + inverse) of every evaluation enqueued while on
  bool dense_timing = false;
  std::vector<cudaEvent_t> dense_ev;
  size_t dense_used = 0;

  void want(void** field, size_t bytes) { slab_req.emplace_back(field, bytes); }
  int commit_slab() {
    size_t off = 0;
    std::vector<size_t> offs;
    for (auto& r : slab_req) { offs.push_back(off); off += (r.second + 255) / 256 * 256; }
    slab_bytes = (long long)off;
    int e = rt::dev_malloc((void**)&slab, (size_t)slab_bytes * nbatch);
    if (e) return e;
    for (size_t i = 0; i < slab_req.size(); ++i) *slab_req[i].first = slab + offs[i];
    bstride = nbatch > 1 ? slab_bytes : 0;
    cur = 0;
    return 0;
  }
  void select(int member) {
    if (member == cur) return;
    const long long delta = (long long)(member - cur) * slab_bytes;
    for (auto& r : slab_req) *r.first = static_cast<char*>(*r.first) + delta;
    cur = member;
  }
};

// batched launches address member 0; the caller's selection is restored afterwards
struct BaseMember {
  cs_fit* f; int keep;
  explicit BaseMember(cs_fit* f_) : f(f_), keep(f_->cur) { f->select(0); }
  ~BaseMember() { f->select(keep); }
};

static constexpr int TRACE_CAP_ITERS = 6000;  // per-member trace capacity inside the slab (reference default maxiter = 5000)

static int split_plan(int row_blocks, int cols, int* nsplit, int* cols_per_split) {
  int target = 4 * rt::sm_count();
  int ns = (target + row_blocks - 1) / row_blocks;
  const int cblocks = cols / csk::TB;
  if (ns > cblocks) ns = cblocks;
  if (ns < 1) ns = 1;
  int cps = ((cblocks + ns - 1) / ns) * csk::TB;
  ns = (cols + cps - 1) / cps;
  *nsplit = ns;
  *cols_per_split = cps;
  return 0;
}

// K1 / K3 launchers: the tensor-path kernels (kernels_mma.cuh) for p <= 32, the element-wise ones (kernels.cuh) otherwise
// or with CS_GRAD_OLD=1 (A/B switch)
static bool use_mma_kernels(int p) {
  static const bool old_k = std::getenv("CS_GRAD_OLD") != nullptr;
  return csk::mma_kernels_ok(p) && !old_k;
}
static int use_bulk_copies() {
#ifdef CS_EMU
  return std::getenv("CS_NO_TMA") == nullptr ? 1 : 0;   // read at every handle creation: the CPU tests run both staging paths
#else
  static const int on = std::getenv("CS_NO_TMA") == nullptr ? 1 : 0;
  return on;
#endif
}
// The packed tiles follow X, z, w of the handle: rebuilt in front of every evaluation (18 KB per 64 observations; the data
// may have been rewritten by the caller or by the simulation driver since the last one).
static void pack_tiles(cs_fit* f, cudaStream_t st, const int* skip, unsigned B) {
  if (!f->Xt || !use_mma_kernels(f->p)) return;
  auto k = csk::pack_x_tiles_kernel;
  ++g_launches;
  CS_LAUNCH(k, dim3(f->nblk, 1, B), csk::ETH, 0, st, f->X, (long long)f->np, f->z, f->w, f->params, f->lay, f->nblk, f->Xt, skip);
}
static void launch_gram_train(cs_fit* f, cudaStream_t st, const int* skip, int tile0, int ntiles, unsigned B) {
  const long long ld = f->np;
  ++g_launches;
  if (use_mma_kernels(f->p)) {
    auto k = csk::gram_train_mma_kernel;
    CS_LAUNCH(k, dim3(ntiles, 1, B), csk::ETH, csk::gram_mma_smem_bytes(), st, f->X, ld, f->z, f->w, f->params, f->lay, f->n, f->nblk,
              f->eng.G, ld, skip, tile0, (const double*)f->Xt);
  } else {
    auto k = csk::gram_train_kernel;
    CS_LAUNCH(k, dim3(ntiles, 1, B), csk::ETH, csk::gram_smem_bytes(f->p), st, f->X, ld, f->z, f->w, f->params, f->lay, f->n, f->eng.G, ld,
              skip, tile0, f->xt);
  }
}
static void launch_grad_trace(cs_fit* f, cudaStream_t st, const int* skip, unsigned B) {
  const long long ld = f->np;
  const double* nul = nullptr;
  ++g_launches;
  if (use_mma_kernels(f->p)) {
    auto k = csk::grad_trace_mma_kernel<false>;
    CS_LAUNCH(k, dim3(f->ncta_grad, 1, B), csk::ETH, csk::grad_mma_smem_bytes(), st, f->X, ld, f->z, f->params, f->lay, f->n, f->nblk,
              f->eng.C, ld, f->alpha, f->sc, f->gpart, f->kal_part, ld, nul, nul, nul, skip, (const int*)nullptr, (const double*)f->Xt);
  } else {
    // element-wise kernel; p beyond what its shared-memory accumulators hold runs as several launches, one dimension window each
    auto k = csk::grad_trace_kernel<false>;
    const int dwmax = csk::grad_window(f->p);
    --g_launches;
    for (int d0 = 0; d0 < f->p; d0 += dwmax) {
      const int dw = std::min(dwmax, f->p - d0);
      const bool pre = f->xt.on != 0 && dwmax >= f->p;
      ++g_launches;
      CS_LAUNCH(k, dim3(f->ncta_grad, 1, B), csk::ETH, csk::grad_smem_bytes(f->p, pre, dw), st, f->X, ld, f->z, f->params, f->lay, f->n,
                f->nblk, f->eng.C, ld, f->alpha, f->sc, f->gpart, f->kal_part, ld, nul, nul, nul, skip, (const int*)nullptr,
                pre ? f->xt : rt::XTma{}, d0, dw);
    }
  }
}

static int fit_enqueue_eval(cs_fit* f, const int* skip, cudaEvent_t* phase_ev) {
  const int n = f->n, np = f->np, p = f->p;
  cudaStream_t st = f->st;
  const long long ld = np;
  const unsigned B = (unsigned)f->nbatch;
  int ph = 0;
  // inside the optimisation loop (skip = the member's `done` flag) the status word is not reset per iteration: a
  // non positive definite matrix at ANY iteration of a chunk must reach the host (the reference throws there)
  f->eng.reset_status = (skip == nullptr);
  pack_tiles(f, st, skip, B);
  if (phase_ev) rt::event_record(phase_ev[ph++], st);
  {  // K1.  With the v2 plan the build is split: the tiles of the first diagonal block (rows < W*TILE) on the
     // handle's stream, the rest on the panel stream S2, which every other consumer of G is ordered behind
     // (DenseEngine::build_chol2) -- so the first diagonal block is factored while the rest is still built.
    const int total = csg::lower_tiles(f->nblk);
    int first = total;
    static const bool no_split = std::getenv("CS_NO_GRAM_SPLIT") != nullptr;
    // profile mode (cs_fit_eval_profile) builds the whole matrix on the handle's stream so that the "gram" phase brackets
    // ALL of the build (with the split, 8220 of the 8256 tiles at n = 8192 run on S2, outside any bracket on `st`)
    if (f->eng.plan_version == 2 && f->eng.aux2_ok && !no_split && !phase_ev && !f->eng.left_looking) {
      const int bw = std::min(f->nblk, (f->eng.bounds.size() > 1 ? f->eng.bounds[1] : f->eng.panel_blocks) * f->eng.tile / csk::TB);
      first = csg::lower_tiles(bw);
    }
    launch_gram_train(f, st, skip, 0, first, B);
    if (first < total) {
      rt::event_record(f->eng.ev_gram, st);   // orders S2 behind everything already on the handle's stream
      rt::stream_wait_event(f->eng.aux2, f->eng.ev_gram);
      launch_gram_train(f, f->eng.aux2, skip, first, total - first, B);
    }
  }
  if (phase_ev) rt::event_record(phase_ev[ph++], st);
  const bool dtime = f->dense_timing && !phase_ev && skip == nullptr;
  if (dtime) {
    if (f->dense_used + 2 > f->dense_ev.size()) {
      cudaEvent_t a, b;
      if (rt::event_create(&a) == 0 && rt::event_create(&b) == 0) { f->dense_ev.push_back(a); f->dense_ev.push_back(b); }
    }
    if (f->dense_used + 2 <= f->dense_ev.size()) rt::event_record(f->dense_ev[f->dense_used], st);
  }
  struct DenseStop {  // closes the bracket when the dense phase has been enqueued (both plan branches below)
    cs_fit* f; bool on; cudaStream_t st;
    void stop() { if (on && f->dense_used + 2 <= f->dense_ev.size()) { rt::event_record(f->dense_ev[f->dense_used + 1], st); f->dense_used += 2; } on = false; }
  } dstop{f, dtime, st};
  if (f->eng.inv_pipeline) {
    f->eng.factor_and_invert(st, skip);  // Cholesky with the inverse pipelined behind it
    if (phase_ev) { rt::event_record(phase_ev[ph++], st); rt::event_record(phase_ev[ph++], st); rt::event_record(phase_ev[ph++], st); }
  } else {
    f->eng.set_status_max(st, skip);
    f->eng.run_seq(f->eng.potrf, st, skip, true);
    if (phase_ev) rt::event_record(phase_ev[ph++], st);
    if (!f->eng.single_panel) f->eng.run_seq(f->eng.trtri, st, skip);
    if (phase_ev) rt::event_record(phase_ev[ph++], st);
    f->eng.run_seq(f->eng.xtx, st, skip);
    if (phase_ev) rt::event_record(phase_ev[ph++], st);
  }
  dstop.stop();
  {  // alpha, u, s
    auto k = csk::matvec_kernel<3>;
    ++g_launches;
    CS_LAUNCH(k, dim3(f->nblk, f->nsplit, B), csk::ETH, 0, st, f->eng.C, ld, f->cols_per_split, np, f->y, f->y, f->ones,
              f->params + f->lay.i_mu(), f->mv_part, ld, 3 * ld, skip, f->bstride);
    auto k2 = csk::scalars_kernel;
    ++g_launches;
    CS_LAUNCH(k2, dim3(1, 1, B), csk::SCALARS_THREADS, 0, st, f->mv_part, f->nsplit, ld, 3 * ld, n, np, f->y, f->params, f->lay,
              f->eng.logdet_part, f->eng.N, f->alpha, f->u, f->s, f->sc, skip);
  }
  if (phase_ev) rt::event_record(phase_ev[ph++], st);
  launch_grad_trace(f, st, skip, B);  // K3 / K4
  if (phase_ev) rt::event_record(phase_ev[ph++], st);
  {
    auto k0 = csk::rowstats_kernel;
    ++g_launches;
    CS_LAUNCH(k0, dim3(f->nrs, 1, B), csk::ETH, 0, st, f->kal_part, ld, f->nblk, f->eng.C, ld, f->alpha, f->y, f->w, f->params,
              f->lay, n, f->sc, f->rs_part, skip);
    auto k = csk::finalize_kernel;
    ++g_launches;
    CS_LAUNCH(k, dim3(1, 1, B), csk::ETH, 0, st, f->gpart, f->ncta_grad, f->rs_part, f->nrs, f->params, f->lay, n, f->sc, f->grad,
              skip);
  }
  if (phase_ev) rt::event_record(phase_ev[ph++], st);
  return 0;
}

static void fit_enqueue_loop_step(cs_fit* f) {
  auto k = csk::loop_step_kernel;
  ++g_launches;
  CS_LAUNCH(k, dim3(1, 1, (unsigned)f->nbatch), 64, 0, f->st, f->cfg.optim, f->lay, f->cfg.lr, f->cfg.beta1, f->cfg.beta2, f->cfg.eps,
            f->cfg.momentum, f->m, f->v, f->grad, f->params, f->sc, f->ls, f->trace, (const int*)f->eng.status);
}

extern "C" {

int cs_version(void) { return 100; }
const char* cs_last_error(void) { return g_err.c_str(); }
int cs_device_count(void) {
  int n = 0;
  if (rt::device_count(&n) != 0) return 0;
  return n;
}
int cs_set_device(int device) {
  if (need_device()) return CS_ERR_NODEVICE;
  RT(rt::set_device(device));
  return 0;
}
long long cs_launch_count(void) { return g_launches; }
int cs_release_scratch(void) {  // predict / treatment scratch retained by the calling host thread on the current device
  if (need_device()) return CS_ERR_NODEVICE;
  thread_arena()->release();
  return 0;
}

static int fit_create_impl(const cs_fit_config* cfg, int nbatch, int n, int p, const double* const* ys,
                           const double* const* Xs, const double* const* zs, const double* const* ws,
                           const double* const* inits, cs_fit** out) {
  // ys == nullptr: "device fill" -- the slabs are allocated and zeroed, the caller's kernels write the training data and the
  // initial parameters of every member (the simulation driver, sim.cuh)
  const bool device_fill = (ys == nullptr);
  if (!cfg || !out || (!device_fill && (!Xs || !zs || !inits))) return fail(CS_ERR_ARG, "cs_fit_create: null argument");
  if (nbatch < 1) return fail(CS_ERR_ARG, "cs_fit_create_batch: nbatch must be positive");
  for (int b = 0; b < nbatch && !device_fill; ++b)
    if (!ys[b] || !Xs[b] || !zs[b] || !inits[b]) return fail(CS_ERR_ARG, "cs_fit_create: null argument (member %d)", b);
  if (n < 1 || p < 1) return fail(CS_ERR_ARG, "cs_fit_create: n and p must be positive");
  if (p > CS_MAX_P) return fail(CS_ERR_UNSUPPORTED, "cs_fit_create: p=%d exceeds CS_MAX_P=%d", p, CS_MAX_P);
  if (cfg->kind != CS_GP && cfg->kind != CS_TP) return fail(CS_ERR_ARG, "cs_fit_create: bad kind");
  if (cfg->optim < 0 || cfg->optim > 2) return fail(CS_ERR_ARG, "cs_fit_create: bad optimizer");
  if (need_device()) return CS_ERR_NODEVICE;
  int tile = cfg->tile ? cfg->tile : csdense::pick_tile(n);
  if (tile != 64 && tile != 128) return fail(CS_ERR_ARG, "cs_fit_create: tile must be 0, 64 or 128");
  cs_fit* f = new cs_fit();
  *out = nullptr;
  f->cfg = *cfg;
  f->cfg.tile = tile;
  if (nbatch > 1) f->cfg.inv_pipeline = std::getenv("CS_BATCH_PIPE") ? 1 : 0;  // experiment knob; sequential inverse measured best for small-n batches
  f->lay.kind = cfg->kind;
  f->lay.p = p;
  f->n = n;
  f->p = p;
  f->nbatch = nbatch;
  auto bail = [&](int code) { cs_fit_destroy(f); return code; };
  if (rt::stream_create(&f->st)) return bail(fail(CS_ERR_CUDA, "stream create failed"));
  const int np = DenseEngine::padded(n, tile);
  f->np = np;
  f->nblk = np / csk::TB;
  const int NP = f->lay.np();
  const size_t vb = sizeof(double) * np;
  const size_t mb = sizeof(double) * (size_t)np * np;
  // ---- one slab per member: dense engine buffers first, then the vectors ----
  DenseEngine& E = f->eng;
  f->want((void**)&E.G, mb); f->want((void**)&E.X, mb); f->want((void**)&E.C, mb);
  f->want((void**)&E.logdet_part, sizeof(double) * (np / tile));
  f->want((void**)&E.Pbuf, DenseEngine::pbuf_bytes(np, tile));
  f->want((void**)&E.Rbuf, DenseEngine::pbuf_bytes(np, tile));
  f->want((void**)&E.status, sizeof(int));
  f->want((void**)&f->X, vb * p); f->want((void**)&f->y, vb); f->want((void**)&f->z, vb); f->want((void**)&f->w, vb);
  f->want((void**)&f->ones, vb);
  if (csk::mma_kernels_ok(p) && use_bulk_copies()) f->want((void**)&f->Xt, sizeof(double) * ((size_t)csk::XT * f->nblk + csk::CST_N));
  f->want((void**)&f->params, sizeof(double) * NP); f->want((void**)&f->grad, sizeof(double) * NP);
  f->want((void**)&f->m, sizeof(double) * NP); f->want((void**)&f->v, sizeof(double) * NP);
  f->want((void**)&f->alpha, vb); f->want((void**)&f->u, vb); f->want((void**)&f->s, vb);
  split_plan(f->nblk, np, &f->nsplit, &f->cols_per_split);
  f->want((void**)&f->mv_part, vb * 3 * f->nsplit);
  // persistent gradient CTAs: one per SM for a single fit, shared between the members of a batch
  f->ncta_grad = std::min(csg::lower_tiles(f->nblk), std::max(4, (nbatch > 1 ? 2 : 1) * rt::sm_count() / nbatch));
  f->want((void**)&f->gpart, sizeof(double) * (2 * p + 2) * f->ncta_grad);
  f->want((void**)&f->kal_part, vb * f->nblk);
  f->nrs = (np + csk::ETH - 1) / csk::ETH;
  f->want((void**)&f->rs_part, sizeof(double) * 2 * f->nrs);
  f->want((void**)&f->sc, sizeof(EvalScalars)); f->want((void**)&f->ls, sizeof(LoopState));
  f->trace_cap = 2 * (TRACE_CAP_ITERS + 2);
  f->want((void**)&f->trace, sizeof(double) * f->trace_cap);
  f->trace_in_slab = true;
  if (f->commit_slab()) return bail(fail(CS_ERR_NOMEM, "device allocation failed (%lld bytes x %d members)", f->slab_bytes, nbatch));
  f->lay.bstride = f->bstride;
  if (csk::grad_prefetch_fits(p)) rt::make_x_tensormap(&f->xt, f->X, np, p, nbatch, f->slab_bytes);  // member 0 is selected here
  E.external_buffers = true;
  E.nbatch = nbatch;
  E.bstride = f->bstride;
  if (int e = E.init(n, tile, cfg->gemm_tile, f->cfg.inv_pipeline, cfg->partition)) return bail(fail(e == 2 ? CS_ERR_NOMEM : CS_ERR_CUDA, "dense engine init failed: %s", rt::err_str(e)));
  cudaStream_t st = f->st;
  if (rt::memset0(f->slab, (size_t)f->slab_bytes * nbatch, st)) return bail(fail(CS_ERR_CUDA, "memset failed"));
  for (int b = 0; b < nbatch && !device_fill; ++b) {
    f->select(b);
    if (upload_matrix(f->X, np, np, p, Xs[b], n, p, st)) return bail(fail(CS_ERR_CUDA, "upload X failed"));
    if (upload_vector(f->y, np, ys[b], n, 0.0, st) || upload_vector(f->z, np, zs[b], n, 0.0, st))
      return bail(fail(CS_ERR_CUDA, "upload y/z failed"));
    auto k = fill_kernel;
    ++g_launches;
    CS_LAUNCH(k, (np + 255) / 256, 256, 0, st, f->ones, (long long)np, 1.0);
    const double* w = ws ? ws[b] : nullptr;
    if (w) { if (upload_vector(f->w, np, w, n, 1.0, st)) return bail(fail(CS_ERR_CUDA, "upload w failed")); }
    else { ++g_launches; CS_LAUNCH(k, (np + 255) / 256, 256, 0, st, f->w, (long long)np, 1.0); }
    if (rt::h2d(f->params, inits[b], sizeof(double) * NP, st)) return bail(fail(CS_ERR_CUDA, "upload params failed"));
  }
  f->select(0);
  { auto k = csk::gram_train_kernel; CS_SET_SMEM(k, csk::gram_smem_bytes(csk::PMAX)); }
  { auto k = csk::gram_cross_kernel; CS_SET_SMEM(k, csk::gram_smem_bytes(csk::PMAX)); }
  { auto k = csk::grad_trace_kernel<false>; CS_SET_SMEM(k, csk::SMEM_LIMIT); }
  { auto k = csk::grad_trace_kernel<true>; CS_SET_SMEM(k, csk::SMEM_LIMIT); }
  { auto k = csp::gram_sum_kernel; CS_SET_SMEM(k, csk::gram_smem_bytes(csk::PMAX)); }
  { auto k = csk::gram_train_mma_kernel; CS_SET_SMEM(k, csk::gram_mma_smem_bytes()); }
  { auto k = csk::grad_trace_mma_kernel<false>; CS_SET_SMEM(k, csk::grad_mma_smem_bytes()); }
  { auto k = csk::grad_trace_mma_kernel<true>; CS_SET_SMEM(k, csk::grad_mma_smem_bytes()); }
  for (int i = 0; i < 16; ++i)
    if (rt::event_create(&f->ev[i])) return bail(fail(CS_ERR_CUDA, "event create failed"));
  f->ev_ok = true;
  if (rt::stream_sync(st)) return bail(fail(CS_ERR_CUDA, "fit create sync failed: %s", rt::err_str(rt::last_error())));
  *out = f;
  return 0;
}

int cs_fit_create(const cs_fit_config* cfg, int n, int p, const double* y, const double* X, const double* z,
                  const double* w, const double* init_params, cs_fit** out) {
  const double* ws[1] = {w};
  if (!y) return fail(CS_ERR_ARG, "cs_fit_create: null argument");
  return fit_create_impl(cfg, 1, n, p, &y, &X, &z, ws, &init_params, out);
}

int cs_fit_create_batch(const cs_fit_config* cfg, int nbatch, int n, int p, const double* const* y, const double* const* X,
                        const double* const* z, const double* const* w, const double* const* init_params, cs_fit** out) {
  if (!y) return fail(CS_ERR_ARG, "cs_fit_create_batch: null argument");
  return fit_create_impl(cfg, nbatch, n, p, y, X, z, w, init_params, out);
}

int cs_fit_batch_size(const cs_fit* f) { return f ? f->nbatch : CS_ERR_ARG; }

int cs_fit_select(cs_fit* f, int member) {
  if (!f || member < 0 || member >= f->nbatch) return fail(CS_ERR_ARG, "cs_fit_select: member out of range");
  f->select(member);
  return 0;
}

void cs_fit_destroy(cs_fit* f) {
  if (!f) return;
  rt::stream_sync(f->st);
  if (f->trace && !f->trace_in_slab) rt::dev_free(f->trace);
  if (f->slab) rt::dev_free(f->slab);
  if (f->flush) rt::dev_free(f->flush);
  if (f->pinned) rt::host_free(f->pinned);
  if (f->h_ls) rt::host_free(f->h_ls);
  if (f->graph_ok) rt::graph_destroy(f->iter_graph);
  if (f->loop_ok) rt::graph_destroy(f->loop_graph);
  for (cudaEvent_t e : f->dense_ev) rt::event_destroy(e);
  f->eng.destroy();
  if (f->ev_ok) for (int i = 0; i < 16; ++i) rt::event_destroy(f->ev[i]);
  rt::stream_destroy(f->st);
  delete f;
}

int cs_fit_nparam(const cs_fit* f) { return f ? f->lay.np() : CS_ERR_ARG; }

int cs_fit_set_params(cs_fit* f, const double* params) {
  if (!f || !params) return fail(CS_ERR_ARG, "cs_fit_set_params: null argument");
  RT(rt::h2d(f->params, params, sizeof(double) * f->lay.np(), f->st));
  RT(rt::stream_sync(f->st));
  return 0;
}

int cs_fit_get_params(cs_fit* f, double* params) {
  if (!f || !params) return fail(CS_ERR_ARG, "cs_fit_get_params: null argument");
  RT(rt::d2h(params, f->params, sizeof(double) * f->lay.np(), f->st));
  RT(rt::stream_sync(f->st));
  return 0;
}

int cs_fit_set_data(cs_fit* f, const double* y, const double* X, const double* z, const double* w) {
  if (!f || !y || !X || !z) return fail(CS_ERR_ARG, "cs_fit_set_data: null argument");
  const int n = f->n, p = f->p, np = f->np;
  cudaStream_t st = f->st;
  // stage through pinned host memory so the copies are truly asynchronous DMA transfers
  const size_t need = sizeof(double) * (size_t)n * (p + 3);
  if (f->pinned_bytes < need) {
    if (f->pinned) rt::host_free(f->pinned);
    f->pinned = nullptr;
    RT(rt::host_alloc((void**)&f->pinned, need));
    f->pinned_bytes = need;
  }
  RT(rt::stream_sync(st));  // previous use of the staging buffer has drained
  double* hp = f->pinned;
  std::memcpy(hp, X, sizeof(double) * (size_t)n * p);
  std::memcpy(hp + (size_t)n * p, y, sizeof(double) * n);
  std::memcpy(hp + (size_t)n * (p + 1), z, sizeof(double) * n);
  if (w) std::memcpy(hp + (size_t)n * (p + 2), w, sizeof(double) * n);
  else for (int i = 0; i < n; ++i) hp[(size_t)n * (p + 2) + i] = 1.0;
  RT(upload_matrix(f->X, np, np, p, hp, n, p, st));
  RT(upload_vector(f->y, np, hp + (size_t)n * p, n, 0.0, st));
  RT(upload_vector(f->z, np, hp + (size_t)n * (p + 1), n, 0.0, st));
  RT(upload_vector(f->w, np, hp + (size_t)n * (p + 2), n, 1.0, st));
  return 0;
}

int cs_fit_reset_optimizer(cs_fit* f) {  // every member of a batched handle (cs_fit_run advances all of them)
  if (!f) return fail(CS_ERR_ARG, "null handle");
  BaseMember base(f);
  for (int b = 0; b < f->nbatch; ++b) {
    RT(rt::memset0(reinterpret_cast<char*>(f->m) + b * f->bstride, sizeof(double) * f->lay.np(), f->st));
    RT(rt::memset0(reinterpret_cast<char*>(f->v) + b * f->bstride, sizeof(double) * f->lay.np(), f->st));
  }
  return 0;
}

int cs_fit_eval_async(cs_fit* f) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  BaseMember base(f);
  fit_enqueue_eval(f, nullptr, nullptr);
  KCHECK();
  return 0;
}

int cs_fit_sync(cs_fit* f) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  RT(rt::stream_sync(f->st));
  return 0;
}

static int read_status(cs_fit* f) {
  int status = 0;
  RT(rt::d2h(&status, f->eng.status, sizeof(int), f->st));
  RT(rt::stream_sync(f->st));
  if (status != 0x7fffffff && status > 0) {
    fail(status, "inv_sympd(): matrix is singular or not positive definite (pivot %d)", status);
    return status;
  }
  return 0;
}

// outputs of the SELECTED member (cs_fit_select; member 0 for a single fit)
int cs_fit_fetch(cs_fit* f, double* grad, double* stats, double* mu_solution) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  EvalScalars h{};
  RT(rt::d2h(&h, f->sc, sizeof h, f->st));
  if (grad) RT(rt::d2h(grad, f->grad, sizeof(double) * f->lay.np(), f->st));
  RT(rt::stream_sync(f->st));
  if (stats) { stats[0] = h.stats[0]; stats[1] = h.stats[1]; }
  if (mu_solution) *mu_solution = h.mu_sol;
  return read_status(f);
}

int cs_fit_eval(cs_fit* f, const double* params, double* grad, double* stats, double* mu_solution) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  if (params) RT(rt::h2d(f->params, params, sizeof(double) * f->lay.np(), f->st));
  int e = cs_fit_eval_async(f);
  if (e) return e;
  return cs_fit_fetch(f, grad, stats, mu_solution);
}

// ---- the loop of R/main.R:79-88 as begin / enqueue-chunk / poll / finish, so that several
// ---- handles can be driven concurrently from one host thread (cs_fit_run_many); a batched handle
// ---- advances all of its members with every launch

static int fit_run_begin(cs_fit* f, int maxiter, double tol) {
  BaseMember base(f);
  cudaStream_t st = f->st;
  const int B = f->nbatch;
  const int cap = 2 * (maxiter + 2);
  if (f->trace_cap < cap) {
    if (B > 1) return fail(CS_ERR_UNSUPPORTED, "cs_fit_run: maxiter %d exceeds the batched trace capacity %d", maxiter, TRACE_CAP_ITERS);
    if (f->trace && !f->trace_in_slab) rt::dev_free(f->trace);
    f->trace = nullptr;
    f->trace_in_slab = false;  // single fit: a private buffer (bstride == 0, no slab addressing involved)
    for (auto& r : f->slab_req) if (r.first == (void**)&f->trace) r.first = (void**)&f->trace_unused;
    RT(rt::dev_malloc((void**)&f->trace, sizeof(double) * cap));
    f->trace_cap = cap;
    if (f->graph_ok) { rt::graph_destroy(f->iter_graph); f->graph_ok = false; }  // graph holds the old pointer
    if (f->loop_ok) { rt::graph_destroy(f->loop_graph); f->loop_ok = false; }
    f->graph_tried = false;
  }
  if (!f->h_ls) RT(rt::host_alloc((void**)&f->h_ls, sizeof(HostLoop) * B));
  for (int b = 0; b < B; ++b) {
    RT(rt::memset0(reinterpret_cast<char*>(f->trace) + b * f->bstride, sizeof(double) * cap, st));
    HostLoop& h = f->h_ls[b];
    h = HostLoop{};
    h.ls.it = 1; h.ls.done = 0; h.ls.iters = 0; h.ls.maxiter = maxiter; h.ls.tol = tol; h.ls.prev_lml = 0.0;  // stats[,1] = 0 (R/main.R:77)
    h.ls.err = 0;
    RT(rt::h2d(reinterpret_cast<char*>(f->ls) + b * f->bstride, &h.ls, sizeof(LoopState), st));
    static const int status_ok = 0x7fffffff;
    RT(rt::h2d(reinterpret_cast<char*>(f->eng.status) + b * f->bstride, &status_ok, sizeof(int), st));
  }
  RT(rt::stream_sync(st));
  f->member_status.assign(B, 0);
  f->run_maxiter = maxiter;
  f->run_launched = 0;
  f->run_chunk = f->cfg.chunk;
  if (f->run_chunk <= 0) {
    // keep each poll interval around 5-10 milliseconds of device work (iterations past convergence are skipped on the
    // device, so a longer chunk only costs launch latency, while every poll drains the handle's stream)
    const double flops = (double)f->np * f->np * f->np * B;
    f->run_chunk = (int)std::max(1.0, std::min(64.0, 4.0e10 / std::max(flops, 1.0)));
  }
  if (!f->graph_tried && f->cfg.use_graph >= 0) {
    f->graph_tried = true;
    const long long l0 = g_launches;
    if (rt::graph_begin(st) == 0) {
      fit_enqueue_eval(f, &f->ls->done, nullptr);
      fit_enqueue_loop_step(f);
      f->graph_ok = (rt::graph_end(st, &f->iter_graph) == 0);
      f->launches_per_iter = g_launches - l0;
      g_launches = l0;  // captured, not executed
    }
#ifndef CS_EMU
    // The loop itself on the device: WHILE(any member running) { evaluation; optimizer step; stopping rule }.  Kernels of a
    // conditional body must share one context, so handles whose Cholesky chain runs on green-context streams (padded
    // n >= 4096, iterations of >= 6 ms) keep the chunked host loop.
    static const bool no_while = std::getenv("CS_NO_WHILE") != nullptr;
    if (f->graph_ok && !f->eng.crit_ok && !no_while) {
      const long long l1 = g_launches;
      f->loop_ok = rt::graph_while_build(st, [&](rt::graph_cond_t h) {
        fit_enqueue_eval(f, &f->ls->done, nullptr);
        fit_enqueue_loop_step(f);
        auto kc = csk::loop_cond_kernel;
        CS_LAUNCH(kc, 1, 32, 0, st, (const LoopState*)f->ls, f->bstride, f->nbatch, h);
      }, &f->loop_graph) == 0;
      g_launches = l1;
    }
#endif
  }
  return 0;
}

// enqueue the next chunk of iterations and an asynchronous read-back of the loop state
static int fit_run_enqueue(cs_fit* f) {
  BaseMember base(f);
  cudaStream_t st = f->st;
  const bool whole_loop = f->loop_ok && f->run_launched == 0;
  if (whole_loop) {  // ONE launch runs every iteration up to convergence / maxiter of every member
    if (rt::graph_launch(f->loop_graph, st) != 0) return fail(CS_ERR_CUDA, "loop graph launch failed: %s", rt::err_str(rt::last_error()));
    f->loop_running = true;
  }
  const int todo = whole_loop ? 0 : std::min(f->run_chunk, f->run_maxiter - f->run_launched);
  if (whole_loop) f->run_launched = f->run_maxiter;
  for (int i = 0; i < todo; ++i) {
    if (f->graph_ok) {
      if (rt::graph_launch(f->iter_graph, st) != 0)
        return fail(CS_ERR_CUDA, "graph launch failed: %s", rt::err_str(rt::last_error()));
      g_launches += f->launches_per_iter;
    } else {
      fit_enqueue_eval(f, &f->ls->done, nullptr);
      fit_enqueue_loop_step(f);
    }
  }
  f->run_launched += todo;
  KCHECK();
  for (int b = 0; b < f->nbatch; ++b) {
    RT(rt::d2h(&f->h_ls[b].ls, reinterpret_cast<char*>(f->ls) + b * f->bstride, sizeof(LoopState), st));
    RT(rt::d2h(&f->h_ls[b].status, reinterpret_cast<char*>(f->eng.status) + b * f->bstride, sizeof(int), st));
  }
  return 0;
}

// wait for the chunk; *finished = 1 when every member has stopped.  A member that failed (non positive definite K_n,
// non-finite LML) has stopped itself on the device with LoopState.err set (loop_step_kernel); the others keep running --
// in the reference only the failing CausalStump() call errors, not the other replicates of the cluster.
static int fit_run_poll(cs_fit* f, int* finished) {
  RT(rt::stream_sync(f->st));
  if (f->loop_running) {  // claimed launch count of the device-side loop: iterations actually run x launches of one iteration
    int its = 0;
    for (int b = 0; b < f->nbatch; ++b) its = std::max(its, f->h_ls[b].ls.it - 1);
    g_launches += (f->launches_per_iter + 1) * (long long)its;
    f->loop_running = false;
  }
  int all_done = 1;
  for (int b = 0; b < f->nbatch; ++b)
    if (!f->h_ls[b].ls.done) all_done = 0;
  *finished = (all_done || f->run_launched >= f->run_maxiter) ? 1 : 0;
  return 0;
}

static int member_fail(int code, int member, int iteration) {
  if (code > 0)
    return fail(code, "inv_sympd(): matrix is singular or not positive definite (pivot %d) at iteration %d (member %d)", code,
                iteration, member);
  if (code == CS_ERR_NONFINITE)
    return fail(code, "log marginal likelihood is not finite at iteration %d (member %d): missing value where TRUE/FALSE needed",
                iteration, member);
  return code;
}

static int fit_run_finish_enqueue(cs_fit* f) {
  // final get_train_stats with the final parameters (R/main.R:88)
  BaseMember base(f);
  fit_enqueue_eval(f, nullptr, nullptr);
  KCHECK();
  return 0;
}

// iters[nbatch]; stats_trace = nbatch consecutive 2 x (maxiter+2) column-major blocks.  Every member is fetched; the
// status of each lands in f->member_status (cs_fit_member_status) and the first failure is the return value.
static int fit_run_finish_fetch(cs_fit* f, int* iters, double* stats_trace) {
  const int keep = f->cur;
  const int cap = 2 * (f->run_maxiter + 2);
  int first = 0, first_member = -1, first_it = 0;
  for (int b = 0; b < f->nbatch; ++b) {
    f->select(b);
    const int it = f->h_ls[b].ls.done ? f->h_ls[b].ls.iters : f->run_maxiter;
    double st2[2] = {0.0, 0.0};
    int rc = f->h_ls[b].ls.err;
    if (rc == 0) rc = cs_fit_fetch(f, nullptr, st2, nullptr);  // > 0: the final get_train_stats was not positive definite
    if (rc < 0 && rc != CS_ERR_NONFINITE) { f->select(keep); return rc; }  // CUDA / transfer failure: not a property of the member
    f->member_status[b] = rc;
    if (rc != 0 && first == 0) { first = rc; first_member = b; first_it = it; }
    if (iters) iters[b] = it;
    if (stats_trace) {
      double* tr = stats_trace + (size_t)b * cap;
      if (rt::d2h(tr, f->trace, sizeof(double) * cap, f->st) || rt::stream_sync(f->st)) { f->select(keep); return fail(CS_ERR_CUDA, "trace download failed"); }
      if (rc == 0) { tr[2 * (it + 1) + 0] = st2[0]; tr[2 * (it + 1) + 1] = st2[1]; }
    }
  }
  f->select(keep);
  return first ? member_fail(first, first_member, first_it) : 0;
}

int cs_fit_run(cs_fit* f, int maxiter, double tol, int* iters, double* stats_trace) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  if (maxiter < 1) return fail(CS_ERR_ARG, "cs_fit_run: maxiter must be >= 1");
  int e = fit_run_begin(f, maxiter, tol);
  if (e) return e;
  for (int fin = 0; !fin;) {
    if ((e = fit_run_enqueue(f))) return e;
    if ((e = fit_run_poll(f, &fin))) return e;
  }
  if ((e = fit_run_finish_enqueue(f))) return e;
  return fit_run_finish_fetch(f, iters, stats_trace);
}

int cs_fit_dense_timing(cs_fit* f, int on) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  f->dense_timing = on != 0;
  f->dense_used = 0;
  return 0;
}

int cs_fit_dense_ms(cs_fit* f, float* total_ms, int* evaluations) {
  if (!f || !total_ms || !evaluations) return fail(CS_ERR_ARG, "cs_fit_dense_ms: null argument");
  RT(rt::stream_sync(f->st));
  float sum = 0.f;
  for (size_t i = 0; i + 1 < f->dense_used; i += 2) {
    float ms = 0.f;
    RT(rt::event_elapsed(&ms, f->dense_ev[i], f->dense_ev[i + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *evaluations = (int)(f->dense_used / 2);
  f->dense_used = 0;
  return 0;
}

int cs_fit_member_status(const cs_fit* f, int* status) {
  if (!f || !status) return fail(CS_ERR_ARG, "cs_fit_member_status: null argument");
  for (int b = 0; b < f->nbatch; ++b) status[b] = b < (int)f->member_status.size() ? f->member_status[b] : 0;
  return 0;
}

// handles may be batched: iters then holds sum(batch sizes) entries in handle order, stats_traces[i] nbatch_i blocks
int cs_fit_run_many(cs_fit** fits, int nfits, int maxiter, double tol, int* iters, double** stats_traces) {
  if (!fits || nfits < 1 || maxiter < 1) return fail(CS_ERR_ARG, "cs_fit_run_many: bad argument");
  for (int i = 0; i < nfits; ++i) if (!fits[i]) return fail(CS_ERR_ARG, "cs_fit_run_many: null handle");
  int e;
  for (int i = 0; i < nfits; ++i) if ((e = fit_run_begin(fits[i], maxiter, tol))) return e;
  std::vector<int> active(nfits, 1);
  int nactive = nfits;
  // every handle is re-fed right after its own poll, so its stream never waits for the polls of the others
  for (int i = 0; i < nfits; ++i) if ((e = fit_run_enqueue(fits[i]))) return e;
  while (nactive > 0) {
    for (int i = 0; i < nfits; ++i)
      if (active[i]) {
        int fin = 0;
        if ((e = fit_run_poll(fits[i], &fin))) return e;
        if (fin) { active[i] = 0; --nactive; if ((e = fit_run_finish_enqueue(fits[i]))) return e; }
        else if ((e = fit_run_enqueue(fits[i]))) return e;
      }
  }
  // every handle is fetched even when a member of an earlier one failed; the first failure is the return value and
  // cs_fit_member_status tells which members of which handle are affected
  int off = 0, first = 0;
  std::string first_msg;
  for (int i = 0; i < nfits; ++i) {
    e = fit_run_finish_fetch(fits[i], iters ? iters + off : nullptr, stats_traces ? stats_traces[i] : nullptr);
    if (e < 0 && e != CS_ERR_NONFINITE) return e;
    if (e != 0 && first == 0) { first = e; first_msg = g_err + " (handle " + std::to_string(i) + ")"; }
    off += fits[i]->nbatch;
  }
  if (first) g_err = first_msg;
  return first;
}

int cs_fit_get_matrices(cs_fit* f, double* Kmat, double* Km, double* Ka, double* invK) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  const int n = f->n, np = f->np;
  cudaStream_t st = f->st;
  if (Kmat || Km || Ka) {
    DevBuf bK, bM, bA;
    const size_t mb = sizeof(double) * (size_t)np * np;
    if (Kmat && bK.alloc(mb)) return fail(CS_ERR_NOMEM, "alloc failed");
    if (Km && bM.alloc(mb)) return fail(CS_ERR_NOMEM, "alloc failed");
    if (Ka && bA.alloc(mb)) return fail(CS_ERR_NOMEM, "alloc failed");
    auto k = csk::gram_cross_kernel;
    ++g_launches;
    CS_LAUNCH(k, dim3(f->nblk, f->nblk), csk::ETH, csk::gram_smem_bytes(f->p), st, f->X, (long long)np, f->z, n, f->X,
              (long long)np, f->z, n, f->params, f->lay, bK.d(), bM.d(), bA.d(), (long long)np);
    KCHECK();
    if (Kmat) RT(download_matrix(Kmat, n, n, bK.d(), np, st));
    if (Km) RT(download_matrix(Km, n, n, bM.d(), np, st));
    if (Ka) RT(download_matrix(Ka, n, n, bA.d(), np, st));
    RT(rt::stream_sync(st));
  }
  if (invK) {
    RT(download_matrix(invK, n, n, f->eng.C, np, st));
    RT(rt::stream_sync(st));
  }
  return 0;
}

// test hook: verifies that the multi-stream Cholesky plan of an n x n engine has no unordered
// conflicting launches (returns the number of races found; 0 = clean)
int cs_debug_check_plan(int n, int tile, int gemm_tile, int verbose) {
  if (need_device()) return CS_ERR_NODEVICE;
  DenseEngine eng;
  if (int e = eng.init(n, tile ? tile : csdense::pick_tile(n), gemm_tile, 1)) { eng.destroy(); return fail(CS_ERR_CUDA, "engine init failed (%d)", e); }
  const int bad = eng.check_plan(eng.potrf, verbose) + eng.check_plan(eng.potrf_inv, verbose);
  eng.destroy();
  return bad;
}

int cs_fit_event_record(cs_fit* f, int slot) {
  if (!f || slot < 0 || slot >= 16) return fail(CS_ERR_ARG, "bad event slot");
  RT(rt::event_record(f->ev[slot], f->st));
  return 0;
}

int cs_fit_event_elapsed_ms(cs_fit* f, int a, int b, float* ms) {
  if (!f || !ms || a < 0 || a >= 16 || b < 0 || b >= 16) return fail(CS_ERR_ARG, "bad event slot");
  RT(rt::event_sync(f->ev[b]));
  RT(rt::event_elapsed(ms, f->ev[a], f->ev[b]));
  return 0;
}

int cs_fit_eval_profile(cs_fit* f, float phase_ms[CS_NPHASE]) {
  if (!f || !phase_ms) return fail(CS_ERR_ARG, "null argument");
  { BaseMember base(f); fit_enqueue_eval(f, nullptr, f->ev); }
  KCHECK();
  RT(rt::stream_sync(f->st));
  for (int i = 0; i < CS_NPHASE; ++i) RT(rt::event_elapsed(&phase_ms[i], f->ev[i], f->ev[i + 1]));
  return read_status(f);
}

// Timeline of the dense plan of one evaluation: every leaf / GEMM launch is bracketed by timing
// events on its own stream.  rows[6*i..] = (stream, type (0 gemm, 1 leaf), tiles, K of the first
// problem, t_begin_ms, t_end_ms) relative to the start of the factorization.
int cs_debug_timeline(cs_fit* f, int cap, int* n_out, float* rows) {
  if (!f || !n_out || !rows) return fail(CS_ERR_ARG, "null argument");
  cudaStream_t st = f->st;
  std::vector<cudaEvent_t> evs;
  pack_tiles(f, st, nullptr, (unsigned)f->nbatch);
  launch_gram_train(f, st, nullptr, 0, csg::lower_tiles(f->nblk), (unsigned)f->nbatch);  // Gram first so the factorization sees a valid matrix
  RT(rt::stream_sync(st));
  cudaEvent_t base;
  RT(rt::event_create(&base));
  RT(rt::event_record(base, st));
  f->eng.tl = &evs;
  f->eng.reset_status = true;
  f->eng.factor_and_invert(st, nullptr);
  f->eng.tl = nullptr;
  KCHECK();
  RT(rt::stream_sync(st));
  const std::vector<csdense::Launch>& seq = f->eng.inv_pipeline ? f->eng.potrf_inv : f->eng.potrf;
  int n = 0;
  size_t e = 0;
  auto emit = [&](const std::vector<csdense::Launch>& sq) {
    for (const csdense::Launch& L : sq) {
      if (L.is_leaf != csdense::OP_LEAF && L.is_leaf != csdense::OP_GEMM && L.is_leaf != csdense::OP_COPY && L.is_leaf != csdense::OP_TRSM) continue;
      if (e + 1 >= evs.size()) break;
      float t0 = 0.f, t1 = 0.f;
      rt::event_elapsed(&t0, base, evs[e]);
      rt::event_elapsed(&t1, base, evs[e + 1]);
      e += 2;
      if (n < cap) {
        float* r = rows + 6 * (size_t)n;
        r[0] = (float)L.stream; r[1] = (float)((L.is_leaf == csdense::OP_LEAF && L.leaf_mode == 2) ? 6 : L.is_leaf);  // 6: tile inverse (leaf mode 2), 5: triangular solve
        r[2] = (float)(L.is_leaf == csdense::OP_GEMM ? L.ntiles : 1);
        r[3] = (float)((L.is_leaf == csdense::OP_LEAF || L.is_leaf == csdense::OP_TRSM) ? L.blk : (L.is_leaf == csdense::OP_COPY ? 0 : f->eng.probs[L.first_prob].K));
        r[4] = t0; r[5] = t1;
        ++n;
      }
    }
  };
  emit(seq);
  if (!f->eng.inv_pipeline) { emit(f->eng.trtri); emit(f->eng.xtx); }
  for (cudaEvent_t ev : evs) rt::event_destroy(ev);
  rt::event_destroy(base);
  *n_out = n;
  return 0;
}

int cs_flush_l2(cs_fit* f) {
  if (!f) return fail(CS_ERR_ARG, "null handle");
  const size_t bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
  if (!f->flush) {
    RT(rt::dev_malloc((void**)&f->flush, bytes));
    f->flush_bytes = bytes;
  }
  RT(rt::memset0(f->flush, f->flush_bytes, f->st));
  return 0;
}

// ------------------------------------------------------------------------------------
// predict / treatment
// ------------------------------------------------------------------------------------
struct CrossWork {
  DevBuf X2, z2, Kc, V, part, q, tmp, cov;
  int n2 = 0, n2p = 0, nsplit = 1, cps = 0;
};

// shared by cs_predict (use_m = 1, z2 given) and cs_treat (use_m = 0, z2 = 1, Ka only)
// dev_io: X2 (leading dimension ldx2), z2, map and var_diag may live on the device (zero-copy callers: the simulation
// driver, handle-based wrappers); then nothing here waits for the stream -- the scratch comes from the caller's arena
static int cross_common(cs_fit* f, CrossWork& W, int n2, const double* X2, long long ldx2, const double* z2, bool dev_io, bool treat,
                        double* map, double* var_diag, double add_sigma, double add_const, double* cov_out) {
  const int n = f->n, np = f->np, p = f->p, tile = f->eng.gt;
  cudaStream_t st = f->st;
  const int n2p = round_up(n2, f->eng.tile);
  W.n2 = n2; W.n2p = n2p;
  const size_t vb2 = sizeof(double) * n2p;
  if (W.X2.alloc(vb2 * p) || W.z2.alloc(vb2) || W.Kc.alloc(vb2 * np) || W.V.alloc(vb2 * np) || W.q.alloc(vb2) ||
      W.tmp.alloc(vb2))
    return fail(CS_ERR_NOMEM, "predict workspace allocation failed");
  RT(stage_matrix(W.X2.d(), n2p, p, X2, ldx2, n2, p, st));
  if (treat) {
    auto k = fill_kernel;
    ++g_launches;
    CS_LAUNCH(k, (n2p + 255) / 256, 256, 0, st, W.z2.d(), (long long)n2p, 1.0);
  } else {
    if (n2p > n2) { auto kf = fill_kernel; ++g_launches; CS_LAUNCH(kf, (n2p + 255) / 256, 256, 0, st, W.z2.d(), (long long)n2p, 0.0); }
    RT(copy_any(W.z2.d(), z2, sizeof(double) * n2, st));
  }
  const int rb2 = n2p / csk::TB;
  {  // K_xX (or Ka_xX) : n2p x np
    auto k = csk::gram_cross_kernel;
    ++g_launches;
    double* nul = nullptr;
    CS_LAUNCH(k, dim3(rb2, f->nblk), csk::ETH, csk::gram_smem_bytes(p), st, W.X2.d(), (long long)n2p, W.z2.d(), n2, f->X,
              (long long)np, f->z, n, f->params, f->lay, treat ? nul : W.Kc.d(), nul, treat ? W.Kc.d() : nul,
              (long long)n2p);
  }
  split_plan(rb2, np, &W.nsplit, &W.cps);
  if (W.part.alloc(vb2 * W.nsplit)) return fail(CS_ERR_NOMEM, "predict workspace allocation failed");
  {  // map = Kc * alpha (+ mu)
    auto k = csk::matvec_kernel<1>;
    ++g_launches;
    const double* nul = nullptr;
    CS_LAUNCH(k, dim3(rb2, W.nsplit), csk::ETH, 0, st, W.Kc.d(), (long long)n2p, W.cps, np, f->alpha, nul, nul, nul,
              W.part.d(), (long long)n2p, (long long)n2p, (const int*)nullptr, 0LL);
    auto k2 = csp::reduce_parts_kernel;
    ++g_launches;
    CS_LAUNCH(k2, (n2p + 255) / 256, 256, 0, st, W.part.d(), W.nsplit, (long long)n2p, n2p,
              treat ? nul : (const double*)(f->params + f->lay.i_mu()), W.tmp.d());
    if (map) RT(copy_any(map, W.tmp.d(), sizeof(double) * n2, st));
  }
  {  // V = Kc * K^-1  (NT GEMM on the DMMA path; K^-1 is full symmetric)
    csg::GemmProblem P = DenseEngine::mk(W.Kc.d(), n2p, f->eng.C, np, W.V.d(), n2p, n2p / tile, np / tile, np,
                                         csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0);
    DevBuf dP;
    if (dP.alloc(sizeof P)) return fail(CS_ERR_NOMEM, "alloc failed");
    RT(rt::h2d(dP.p, &P, sizeof P, st));
    csdense::launch_gemm(tile, csdense::GK_NT, (const csg::GemmProblem*)dP.p, 1, P.ntiles, st);
    // q_i = sum_k V[i,k] Kc[i,k]
    auto k = csp::rowdot_kernel;
    ++g_launches;
    CS_LAUNCH(k, dim3(rb2, W.nsplit), csk::ETH, 0, st, W.V.d(), W.Kc.d(), (long long)n2p, W.cps, np, W.part.d(),
              (long long)n2p);
    auto k2 = csp::reduce_parts_kernel;
    ++g_launches;
    CS_LAUNCH(k2, (n2p + 255) / 256, 256, 0, st, W.part.d(), W.nsplit, (long long)n2p, n2p, (const double*)nullptr,
              W.q.d());
    auto k3 = csp::var_diag_kernel;
    ++g_launches;
    CS_LAUNCH(k3, (n2p + 255) / 256, 256, 0, st, W.q.d(), W.z2.d(), f->params, f->lay, treat ? 0 : 1, add_sigma,
              add_const, n2, W.tmp.d());
    if (var_diag) RT(copy_any(var_diag, W.tmp.d(), sizeof(double) * n2, st));
    if (!dev_io || !g_arena) RT(rt::stream_sync(st));  // dP must outlive the launch (arena scratch does)
  }
  if (cov_out) {
    if (W.cov.alloc(vb2 * n2p)) return fail(CS_ERR_NOMEM, "cov allocation failed");
    auto k = csk::gram_cross_kernel;
    ++g_launches;
    double* nul = nullptr;
    CS_LAUNCH(k, dim3(rb2, rb2), csk::ETH, csk::gram_smem_bytes(p), st, W.X2.d(), (long long)n2p, W.z2.d(), n2,
              W.X2.d(), (long long)n2p, W.z2.d(), n2, f->params, f->lay, treat ? nul : W.cov.d(), nul,
              treat ? W.cov.d() : nul, (long long)n2p);
    csg::GemmProblem P = DenseEngine::mk(W.V.d(), n2p, W.Kc.d(), n2p, W.cov.d(), n2p, n2p / tile, n2p / tile, np,
                                         csg::KM_FULL, csg::ORD_COL, 0, -1.0, 1.0);
    DevBuf dP;
    if (dP.alloc(sizeof P)) return fail(CS_ERR_NOMEM, "alloc failed");
    RT(rt::h2d(dP.p, &P, sizeof P, st));
    csdense::launch_gemm(tile, csdense::GK_NT, (const csg::GemmProblem*)dP.p, 1, P.ntiles, st);
    auto k2 = csp::cov_diag_kernel;
    ++g_launches;
    CS_LAUNCH(k2, (n2p + 255) / 256, 256, 0, st, W.cov.d(), (long long)n2p, n2, n2p, f->params, f->lay, add_sigma,
              add_const);
    RT(download_matrix(cov_out, n2, n2, W.cov.d(), n2p, st));
    RT(rt::stream_sync(st));
  }
  KCHECK();
  return 0;
}

// sum(cov) of the treatment block without n2 x n2 temporaries: scal_dev[0] = c' K^-1 c with c_j = sum_i Ka_xX[i,j],
// scal_dev[1] = sum(Ka_xx); sum(cov) = scal[1] - scal[0] (+ jitter * n2).  Enqueue only.
static int treat_sumcov_enqueue(cs_fit* f, CrossWork& W, int n2, double* scal_dev) {
  cudaStream_t st = f->st;
  const int np = f->np, n2p = W.n2p, p = f->p;
  DevBuf c, t, part, gs;
  const int rb2 = n2p / csk::TB;
  int ns = 1, cps = 0;
  split_plan(f->nblk, np, &ns, &cps);
  if (c.alloc(sizeof(double) * np) || t.alloc(sizeof(double) * np) || part.alloc(sizeof(double) * np * ns) ||
      gs.alloc(sizeof(double) * rb2 * rb2))
    return fail(CS_ERR_NOMEM, "alloc failed");
  auto k = csp::colsum_kernel;
  ++g_launches;
  CS_LAUNCH(k, (np * 32 + 255) / 256, 256, 0, st, W.Kc.d(), (long long)n2p, n2, np, c.d());
  auto k2 = csk::matvec_kernel<1>;
  ++g_launches;
  const double* nul = nullptr;
  CS_LAUNCH(k2, dim3(f->nblk, ns), csk::ETH, 0, st, f->eng.C, (long long)np, cps, np, c.d(), nul, nul, nul, part.d(),
            (long long)np, (long long)np, (const int*)nullptr, 0LL);
  auto k3 = csp::reduce_parts_kernel;
  ++g_launches;
  CS_LAUNCH(k3, (np + 255) / 256, 256, 0, st, part.d(), ns, (long long)np, np, nul, t.d());
  auto k4 = csp::dot_kernel;
  ++g_launches;
  CS_LAUNCH(k4, 1, csk::ETH, 0, st, c.d(), t.d(), f->n, scal_dev);
  auto k5 = csp::gram_sum_kernel;
  ++g_launches;
  CS_LAUNCH(k5, dim3(rb2, rb2), csk::ETH, csk::gram_smem_bytes(p), st, W.X2.d(), (long long)n2p, n2, f->params, f->lay,
            gs.d());
  ++g_launches;
  CS_LAUNCH(k4, 1, csk::ETH, 0, st, gs.d(), nul, rb2 * rb2, scal_dev + 1);
  KCHECK();
  if (!g_arena) RT(rt::stream_sync(st));  // the scratch above is freed on return unless it comes from the arena
  return 0;
}

int cs_predict(cs_fit* f, int n2, const double* X2, const double* z2, double* map, double* var_diag, double* cov) {
  if (!f || !X2 || !z2 || n2 < 1) return fail(CS_ERR_ARG, "cs_predict: bad argument");
  ArenaScope scope(thread_arena());
  CrossWork W;
  int e = cross_common(f, W, n2, X2, n2, z2, false, false, map, var_diag, 1.0, 0.0, cov);
  if (e) return e;
  RT(rt::stream_sync(f->st));
  return 0;
}

int cs_treat(cs_fit* f, int n2, const double* X2, double cov_jitter, double* map, double* var_diag, double* ate,
             double* ate_var, double* cov) {
  if (!f || !X2 || n2 < 1) return fail(CS_ERR_ARG, "cs_treat: bad argument");
  ArenaScope scope(thread_arena());
  CrossWork W;
  std::vector<double> hmap(n2);
  int e = cross_common(f, W, n2, X2, n2, nullptr, false, true, hmap.data(), var_diag, 0.0, cov_jitter, cov);
  if (e) return e;
  cudaStream_t st = f->st;
  RT(rt::stream_sync(st));
  if (map) std::copy(hmap.begin(), hmap.end(), map);
  if (ate) {
    double s = 0.0;
    for (int i = 0; i < n2; ++i) s += hmap[i];
    *ate = s / n2;  // mean(map), R/kernel_GP_SE_class.R:96
  }
  if (ate_var) {
    DevBuf scal;
    if (scal.alloc(sizeof(double) * 2)) return fail(CS_ERR_NOMEM, "alloc failed");
    if ((e = treat_sumcov_enqueue(f, W, n2, scal.d()))) return e;
    double h[2];
    RT(rt::d2h(h, scal.d(), sizeof h, st));
    RT(rt::stream_sync(st));
    const double dn = (double)f->n;
    *ate_var = (h[1] - h[0] + cov_jitter * n2) / (dn * dn);  // R/kernel_GP_SE_class.R:101 (training n)
  }
  return 0;
}

// ------------------------------------------------------------------------------------
// Student-t posterior sampler
// ------------------------------------------------------------------------------------
int cs_tp_sample(int n2, const double* cov, double df, int nsampling, unsigned long long seed, double* U,
                 double* ate_U) {
  if (n2 < 1 || !cov || !U || !(df > 0.0) || nsampling < 1) return fail(CS_ERR_ARG, "cs_tp_sample: bad argument");
  if (need_device()) return CS_ERR_NODEVICE;
  const int NS = 1000, NSP = csp::SORT_N;            // N_sample (R/kernel_TP_SE_class.R:86) padded to 1024
  const int rank = (int)std::floor(NS * 0.90) - 1;   // 900th order statistic, 1-based in R (:90)
  const int nrep = (nsampling + NS - 1) / NS;        // :88
  DenseEngine eng;
  const int tile = csdense::pick_tile(n2);
  if (int e = eng.init(n2, tile)) { eng.destroy(); return fail(e == 2 ? CS_ERR_NOMEM : CS_ERR_CUDA, "engine init failed"); }
  struct Guard { DenseEngine& e; ~Guard() { e.destroy(); } } guard{eng};
  cudaStream_t st = 0;
  const int np = eng.np;
  RT(upload_matrix(eng.G, np, np, np, cov, n2, n2, st));
  {
    auto k = pad_identity_kernel;
    ++g_launches;
    const long long tot = (long long)np * np;
    CS_LAUNCH(k, (unsigned)((tot + 255) / 256), 256, 0, st, eng.G, (long long)np, n2, np);
  }
  eng.factor(st);
  KCHECK();
  int status = 0;
  RT(rt::d2h(&status, eng.status, sizeof(int), st));
  RT(rt::stream_sync(st));
  if (status != 0x7fffffff && status > 0) return fail(status, "rmvt: covariance not positive definite (pivot %d)", status);
  DevBuf Z, Y, scale, Ud, rm, au, dP;
  if (Z.alloc(sizeof(double) * NSP * np) || Y.alloc(sizeof(double) * NSP * np) || scale.alloc(sizeof(double) * NSP) ||
      Ud.alloc(sizeof(double) * np) || rm.alloc(sizeof(double) * NSP) || au.alloc(sizeof(double)) ||
      dP.alloc(sizeof(csg::GemmProblem)))
    return fail(CS_ERR_NOMEM, "sampler allocation failed");
  RT(rt::memset0(Ud.d(), sizeof(double) * np, st));
  RT(rt::memset0(au.d(), sizeof(double), st));
  // Y[s,i] = sum_k Z[s,k] L[i,k]   (NT)
  const int gt = eng.gt;
  // L is lower triangular (k <= j); tiles above the diagonal of G still hold the uploaded
  // covariance and must never be read: KM_LE_J stops every k-range at the diagonal tile.
  csg::GemmProblem P = DenseEngine::mk(Z.d(), NSP, eng.G, np, Y.d(), NSP, NSP / gt, np / gt, np, csg::KM_LE_J,
                                       csg::ORD_COL, 0, 1.0, 0.0);
  RT(rt::h2d(dP.p, &P, sizeof P, st));
  for (int r = 0; r < nrep; ++r) {
    auto k = csp::normal_fill_kernel;
    const long long npair = ((long long)NSP * np + 1) / 2;
    ++g_launches;
    CS_LAUNCH(k, (unsigned)((npair + 255) / 256), 256, 0, st, Z.d(), (long long)NSP, NSP, np, seed, (unsigned)r);
    auto k2 = csp::chi2_scale_kernel;
    ++g_launches;
    CS_LAUNCH(k2, (NSP + 255) / 256, 256, 0, st, scale.d(), NSP, df, seed, (unsigned)r);
    csdense::launch_gemm(gt, csdense::GK_NT, (const csg::GemmProblem*)dP.p, 1, P.ntiles, st);
    auto k3 = csp::order_stat_kernel;
    ++g_launches;
    CS_LAUNCH(k3, n2, 512, 0, st, Y.d(), (long long)NSP, scale.d(), NS, rank, 1.0 / nrep, Ud.d());
    if (ate_U) {
      auto k4 = csp::rowmean_kernel;
      ++g_launches;
      CS_LAUNCH(k4, (NSP + 255) / 256, 256, 0, st, Y.d(), (long long)NSP, NSP, n2, rm.d());
      ++g_launches;
      CS_LAUNCH(k3, 1, 512, 0, st, rm.d(), (long long)NSP, scale.d(), NS, rank, 1.0 / nrep, au.d());
    }
  }
  KCHECK();
  RT(rt::d2h(U, Ud.d(), sizeof(double) * n2, st));
  if (ate_U) RT(rt::d2h(ate_U, au.d(), sizeof(double), st));
  RT(rt::stream_sync(st));
  return 0;
}

// ------------------------------------------------------------------------------------
// stateless mirrors of the registered routines
// ------------------------------------------------------------------------------------
int cs_kernmat(int n1, int n2, int p, const double* X1, const double* X2, const double* Z1, const double* Z2,
               const double* Lm, const double* La, double lambdam, double lambda_a, double* K, double* Km,
               double* Ka) {
  if (n1 < 1 || n2 < 1 || p < 1 || !X1 || !X2 || !Z1 || !Z2 || !Lm || !La) return fail(CS_ERR_ARG, "cs_kernmat: bad argument");
  if (p > CS_MAX_P) return fail(CS_ERR_UNSUPPORTED, "cs_kernmat: p=%d exceeds CS_MAX_P=%d", p, CS_MAX_P);
  if (need_device()) return CS_ERR_NODEVICE;
  cudaStream_t st = 0;
  const int n1p = round_up(n1, csk::TB), n2p = round_up(n2, csk::TB);
  Layout lay{CS_GP, p};
  std::vector<double> hp(lay.np(), 0.0);
  hp[lay.i_lm()] = lambdam;
  hp[lay.i_la()] = lambda_a;
  for (int i = 0; i < p; ++i) { hp[lay.i_Lm() + i] = Lm[i]; hp[lay.i_La() + i] = La[i]; }
  DevBuf dX1, dX2, dZ1, dZ2, dP, dK, dM, dA;
  const size_t mb = sizeof(double) * (size_t)n1p * n2p;
  if (dX1.alloc(sizeof(double) * n1p * p) || dX2.alloc(sizeof(double) * n2p * p) || dZ1.alloc(sizeof(double) * n1p) ||
      dZ2.alloc(sizeof(double) * n2p) || dP.alloc(sizeof(double) * hp.size()) || (K && dK.alloc(mb)) ||
      (Km && dM.alloc(mb)) || (Ka && dA.alloc(mb)))
    return fail(CS_ERR_NOMEM, "cs_kernmat: allocation failed");
  RT(upload_matrix(dX1.d(), n1p, n1p, p, X1, n1, p, st));
  RT(upload_matrix(dX2.d(), n2p, n2p, p, X2, n2, p, st));
  RT(upload_vector(dZ1.d(), n1p, Z1, n1, 0.0, st));
  RT(upload_vector(dZ2.d(), n2p, Z2, n2, 0.0, st));
  RT(rt::h2d(dP.p, hp.data(), sizeof(double) * hp.size(), st));
  auto k = csk::gram_cross_kernel;
  CS_SET_SMEM(k, csk::gram_smem_bytes(csk::PMAX));
  ++g_launches;
  CS_LAUNCH(k, dim3(n1p / csk::TB, n2p / csk::TB), csk::ETH, csk::gram_smem_bytes(p), st, dX1.d(), (long long)n1p,
            dZ1.d(), n1, dX2.d(), (long long)n2p, dZ2.d(), n2, dP.d(), lay, dK.d(), dM.d(), dA.d(), (long long)n1p);
  KCHECK();
  if (K) RT(download_matrix(K, n1, n2, dK.d(), n1p, st));
  if (Km) RT(download_matrix(Km, n1, n2, dM.d(), n1p, st));
  if (Ka) RT(download_matrix(Ka, n1, n2, dA.d(), n1p, st));
  RT(rt::stream_sync(st));
  return 0;
}

static int engine_status(DenseEngine& eng, cudaStream_t st) {
  int status = 0;
  RT(rt::d2h(&status, eng.status, sizeof(int), st));
  RT(rt::stream_sync(st));
  if (status != 0x7fffffff && status > 0)
    return fail(status, "inv_sympd(): matrix is singular or not positive definite (pivot %d)", status);
  return 0;
}

static int engine_logdet(DenseEngine& eng, cudaStream_t st, double* out) {
  std::vector<double> h(eng.N);
  RT(rt::d2h(h.data(), eng.logdet_part, sizeof(double) * eng.N, st));
  RT(rt::stream_sync(st));
  double s = 0.0;
  for (double v : h) s += v;
  *out = 2.0 * s;
  return 0;
}

int cs_invkernel(int n, const double* w, const double* K, double sigma, double* invK, double* logdet) {
  if (n < 1 || !K || !invK) return fail(CS_ERR_ARG, "cs_invkernel: bad argument");
  if (need_device()) return CS_ERR_NODEVICE;
  DenseEngine eng;
  if (int e = eng.init(n, csdense::pick_tile(n))) { eng.destroy(); return fail(e == 2 ? CS_ERR_NOMEM : CS_ERR_CUDA, "engine init failed"); }
  struct Guard { DenseEngine& e; ~Guard() { e.destroy(); } } guard{eng};
  cudaStream_t st = 0;
  const int np = eng.np;
  DevBuf dw;
  if (dw.alloc(sizeof(double) * np)) return fail(CS_ERR_NOMEM, "alloc failed");
  RT(upload_matrix(eng.G, np, np, np, K, n, n, st));
  if (w) RT(upload_vector(dw.d(), np, w, n, 1.0, st));
  else { auto k = fill_kernel; ++g_launches; CS_LAUNCH(k, (np + 255) / 256, 256, 0, st, dw.d(), (long long)np, 1.0); }
  {
    auto k = pad_identity_kernel;
    const long long tot = (long long)np * np;
    ++g_launches;
    CS_LAUNCH(k, (unsigned)((tot + 255) / 256), 256, 0, st, eng.G, (long long)np, n, np);
    auto k2 = add_noise_diag_kernel;
    ++g_launches;
    CS_LAUNCH(k2, (n + 255) / 256, 256, 0, st, eng.G, (long long)np, n, dw.d(), std::exp(sigma));  // :55
  }
  eng.factor(st);
  eng.invert(st);
  KCHECK();
  if (int e = engine_status(eng, st)) return e;
  RT(download_matrix(invK, n, n, eng.C, np, st));
  RT(rt::stream_sync(st));
  if (logdet) return engine_logdet(eng, st, logdet);
  return 0;
}

int cs_mu_solution(int n, const double* y, const double* invK, double* mu) {
  if (n < 1 || !y || !invK || !mu) return fail(CS_ERR_ARG, "cs_mu_solution: bad argument");
  if (need_device()) return CS_ERR_NODEVICE;
  cudaStream_t st = 0;
  const int np = round_up(n, csk::TB), nb = np / csk::TB;
  int ns = 1, cps = 0;
  split_plan(nb, np, &ns, &cps);
  DevBuf dA, dy, d1, part, out;
  if (dA.alloc(sizeof(double) * (size_t)np * np) || dy.alloc(sizeof(double) * np) || d1.alloc(sizeof(double) * np) ||
      part.alloc(sizeof(double) * np * 2 * ns) || out.alloc(sizeof(double)))
    return fail(CS_ERR_NOMEM, "alloc failed");
  RT(upload_matrix(dA.d(), np, np, np, invK, n, n, st));
  RT(upload_vector(dy.d(), np, y, n, 0.0, st));
  {
    auto k = fill_kernel; ++g_launches;
    CS_LAUNCH(k, (np + 255) / 256, 256, 0, st, d1.d(), (long long)np, 1.0);
    auto k2 = csk::matvec_kernel<2>; ++g_launches;
    const double* nul = nullptr;
    CS_LAUNCH(k2, dim3(nb, ns), csk::ETH, 0, st, dA.d(), (long long)np, cps, np, dy.d(), d1.d(), nul, nul, part.d(),
              (long long)np, (long long)(2 * np), (const int*)nullptr, 0LL);
    auto k3 = csk::mu_solution_kernel; ++g_launches;
    CS_LAUNCH(k3, 1, csk::ETH, 0, st, part.d(), ns, (long long)np, (long long)(2 * np), n, out.d());
  }
  KCHECK();
  RT(rt::d2h(mu, out.p, sizeof(double), st));
  RT(rt::stream_sync(st));
  return 0;
}

// grad / stats mirrors share one implementation: a temporary fit handle whose K^-1 is
// either computed (invK == NULL) or uploaded as given.
static int grad_stats_common(int kind, int n, int p, const double* y, const double* X, const double* z,
                             const double* w, const double* Kmat, const double* Km, const double* Ka,
                             const double* invK, const double* params, double* grad, double* stats) {
  cs_fit_config cfg{};
  cfg.kind = kind; cfg.optim = CS_OPT_NADAM; cfg.lr = 0.0; cfg.beta1 = 0.9; cfg.beta2 = 0.999; cfg.eps = 1e-8;
  cs_fit* f = nullptr;
  int e = cs_fit_create(&cfg, n, p, y, X, z, w, params, &f);
  if (e) return e;
  struct Guard { cs_fit* f; ~Guard() { cs_fit_destroy(f); } } guard{f};
  if (!invK) {
    if (Kmat || Km || Ka) return fail(CS_ERR_ARG, "Kmat/Km/Ka given without invK");
    return cs_fit_eval(f, nullptr, grad, stats, nullptr);
  }
  cudaStream_t st = f->st;
  const int np = f->np;
  const long long ld = np;
  // the given inverse: full copy in C for the traces, lower copy in G to obtain log det(invK)
  RT(upload_matrix(f->eng.C, np, np, np, invK, n, n, st));
  RT(upload_matrix(f->eng.G, np, np, np, invK, n, n, st));
  {
    auto k = pad_identity_kernel;
    const long long tot = (long long)np * np;
    ++g_launches; CS_LAUNCH(k, (unsigned)((tot + 255) / 256), 256, 0, st, f->eng.C, ld, n, np);
    ++g_launches; CS_LAUNCH(k, (unsigned)((tot + 255) / 256), 256, 0, st, f->eng.G, ld, n, np);
  }
  f->eng.factor(st);  // logdet_part now holds log det(invK)/2 per block; sign flipped below
  pack_tiles(f, st, nullptr, 1u);
  DevBuf dK, dM, dA;
  const size_t mb = sizeof(double) * (size_t)np * np;
  const bool from_mem = Kmat && Km && Ka;
  if ((Kmat || Km || Ka) && !from_mem) return fail(CS_ERR_ARG, "Kmat, Km and Ka must be given together");
  if (from_mem) {
    if (dK.alloc(mb) || dM.alloc(mb) || dA.alloc(mb)) return fail(CS_ERR_NOMEM, "alloc failed");
    RT(upload_matrix(dK.d(), np, np, np, Kmat, n, n, st));
    RT(upload_matrix(dM.d(), np, np, np, Km, n, n, st));
    RT(upload_matrix(dA.d(), np, np, np, Ka, n, n, st));
  }
  {
    auto kneg = csk::negate_kernel;
    ++g_launches; CS_LAUNCH(kneg, 1, 256, 0, st, f->eng.logdet_part, f->eng.N);
    auto k = csk::matvec_kernel<3>;
    ++g_launches;
    CS_LAUNCH(k, dim3(f->nblk, f->nsplit), csk::ETH, 0, st, f->eng.C, ld, f->cols_per_split, np, f->y, f->y, f->ones,
              f->params + f->lay.i_mu(), f->mv_part, ld, 3 * ld, (const int*)nullptr, 0LL);
    auto k2 = csk::scalars_kernel;
    ++g_launches;
    CS_LAUNCH(k2, 1, csk::SCALARS_THREADS, 0, st, f->mv_part, f->nsplit, ld, 3 * ld, n, np, f->y, f->params, f->lay,
              f->eng.logdet_part, f->eng.N, f->alpha, f->u, f->s, f->sc, (const int*)nullptr);
    if (from_mem) {
      ++g_launches;
      if (use_mma_kernels(p)) {
        auto k3 = csk::grad_trace_mma_kernel<true>;
        CS_LAUNCH(k3, f->ncta_grad, csk::ETH, csk::grad_mma_smem_bytes(), st, f->X, ld, f->z, f->params, f->lay, n, f->nblk,
                  f->eng.C, ld, f->alpha, f->sc, f->gpart, f->kal_part, ld, (const double*)dK.d(), (const double*)dM.d(),
                  (const double*)dA.d(), (const int*)nullptr, (const int*)nullptr, (const double*)f->Xt);
      } else {
        auto k3 = csk::grad_trace_kernel<true>;
        const int dwmax = csk::grad_window(p);
        --g_launches;
        for (int d0 = 0; d0 < p; d0 += dwmax) {
          const int dw = std::min(dwmax, p - d0);
          ++g_launches;
          CS_LAUNCH(k3, f->ncta_grad, csk::ETH, csk::grad_smem_bytes(p, false, dw), st, f->X, ld, f->z, f->params, f->lay, n, f->nblk,
                    f->eng.C, ld, f->alpha, f->sc, f->gpart, f->kal_part, ld, (const double*)dK.d(), (const double*)dM.d(),
                    (const double*)dA.d(), (const int*)nullptr, (const int*)nullptr, rt::XTma{}, d0, dw);
        }
      }
    } else {
      launch_grad_trace(f, st, nullptr, 1u);
    }
    auto k40 = csk::rowstats_kernel;
    ++g_launches;
    CS_LAUNCH(k40, f->nrs, csk::ETH, 0, st, f->kal_part, ld, f->nblk, f->eng.C, ld, f->alpha, f->y, f->w, f->params,
              f->lay, n, f->sc, f->rs_part, (const int*)nullptr);
    auto k4 = csk::finalize_kernel;
    ++g_launches;
    CS_LAUNCH(k4, 1, csk::ETH, 0, st, f->gpart, f->ncta_grad, f->rs_part, f->nrs, f->params, f->lay, n, f->sc,
              f->grad, (const int*)nullptr);
  }
  KCHECK();
  return cs_fit_fetch(f, grad, stats, nullptr);
}

int cs_grad(int kind, int n, int p, const double* y, const double* X, const double* z, const double* w,
            const double* Kmat, const double* Km, const double* Ka, const double* invK, const double* params,
            double* grad, double* stats) {
  if (!y || !X || !z || !params) return fail(CS_ERR_ARG, "cs_grad: null argument");
  return grad_stats_common(kind, n, p, y, X, z, w, Kmat, Km, Ka, invK, params, grad, stats);
}

int cs_stats(int kind, int n, const double* y, const double* Kmat, const double* invK, double mu, double nu,
             double* stats) {
  if (n < 1 || !y || !Kmat || !invK || !stats) return fail(CS_ERR_ARG, "cs_stats: bad argument");
  if (need_device()) return CS_ERR_NODEVICE;
  DenseEngine eng;
  if (int e = eng.init(n, csdense::pick_tile(n))) { eng.destroy(); return fail(e == 2 ? CS_ERR_NOMEM : CS_ERR_CUDA, "engine init failed"); }
  struct Guard { DenseEngine& e; ~Guard() { e.destroy(); } } guard{eng};
  cudaStream_t st = 0;
  const int np = eng.np, nb = np / csk::TB;
  const long long ld = np;
  int ns = 1, cps = 0;
  split_plan(nb, np, &ns, &cps);
  DevBuf dy, dmu, part, alpha, kal, out;
  if (dy.alloc(sizeof(double) * np) || dmu.alloc(sizeof(double)) || part.alloc(sizeof(double) * np * ns) ||
      alpha.alloc(sizeof(double) * np) || kal.alloc(sizeof(double) * np) || out.alloc(sizeof(double) * 2))
    return fail(CS_ERR_NOMEM, "alloc failed");
  RT(upload_vector(dy.d(), np, y, n, 0.0, st));
  RT(rt::h2d(dmu.p, &mu, sizeof(double), st));
  RT(upload_matrix(eng.C, np, np, np, invK, n, n, st));
  RT(upload_matrix(eng.G, np, np, np, invK, n, n, st));
  {
    auto k = pad_identity_kernel;
    const long long tot = (long long)np * np;
    ++g_launches; CS_LAUNCH(k, (unsigned)((tot + 255) / 256), 256, 0, st, eng.G, ld, n, np);
  }
  eng.factor(st);
  const double* nul = nullptr;
  auto mv = csk::matvec_kernel<1>;
  auto rp = csp::reduce_parts_kernel;
  ++g_launches;  // alpha = invK (y - mu); padded rows of C are zero, padded alpha is discarded below
  CS_LAUNCH(mv, dim3(nb, ns), csk::ETH, 0, st, eng.C, ld, cps, np, dy.d(), nul, nul, (const double*)dmu.d(), part.d(), ld,
            ld, (const int*)nullptr, 0LL);
  ++g_launches;
  CS_LAUNCH(rp, (np + 255) / 256, 256, 0, st, part.d(), ns, ld, n, nul, alpha.d());
  if (np > n) { auto kf = fill_kernel; ++g_launches; CS_LAUNCH(kf, (np - n + 255) / 256, 256, 0, st, alpha.d() + n, (long long)(np - n), 0.0); }
  RT(upload_matrix(eng.C, np, np, np, Kmat, n, n, st));  // reuse C for Kmat (stream-ordered after the mat-vec)
  ++g_launches;
  CS_LAUNCH(mv, dim3(nb, ns), csk::ETH, 0, st, eng.C, ld, cps, np, alpha.d(), nul, nul, nul, part.d(), ld, ld,
            (const int*)nullptr, 0LL);
  ++g_launches;
  CS_LAUNCH(rp, (np + 255) / 256, 256, 0, st, part.d(), ns, ld, n, nul, kal.d());
  auto ks = csk::stats_only_kernel;
  ++g_launches;
  CS_LAUNCH(ks, 1, csk::ETH, 0, st, kind, n, (const double*)dy.d(), (const double*)alpha.d(), (const double*)kal.d(),
            (const double*)eng.logdet_part, eng.N, mu, nu, out.d());
  KCHECK();
  if (int e = engine_status(eng, st)) return e;
  RT(rt::d2h(stats, out.p, sizeof(double) * 2, st));
  RT(rt::stream_sync(st));
  return 0;
}

int cs_opt_step(int optim, int nparam, double iter, double lr, double beta1, double beta2, double eps, double* m,
                double* v, const double* grad, double* para) {
  if (nparam < 1 || !m || !grad || !para) return fail(CS_ERR_ARG, "cs_opt_step: null argument");
  if (optim != CS_OPT_NESTEROV && !v) return fail(CS_ERR_ARG, "cs_opt_step: v required for Adam/Nadam");
  if (need_device()) return CS_ERR_NODEVICE;
  cudaStream_t st = 0;
  DevBuf dm, dv, dg, dp;
  const size_t b = sizeof(double) * nparam;
  if (dm.alloc(b) || dv.alloc(b) || dg.alloc(b) || dp.alloc(b)) return fail(CS_ERR_NOMEM, "alloc failed");
  RT(rt::h2d(dm.p, m, b, st));
  if (v) RT(rt::h2d(dv.p, v, b, st)); else RT(rt::memset0(dv.p, b, st));
  RT(rt::h2d(dg.p, grad, b, st));
  RT(rt::h2d(dp.p, para, b, st));
  auto k = csk::opt_step_kernel;
  ++g_launches;
  CS_LAUNCH(k, 1, 32, 0, st, optim, nparam, iter, lr, beta1, beta2, eps, /*momentum=*/beta1, dm.d(), dv.d(),
            (const double*)dg.d(), dp.d());
  KCHECK();
  RT(rt::d2h(m, dm.p, b, st));
  if (v) RT(rt::d2h(v, dv.p, b, st));
  RT(rt::d2h(para, dp.p, b, st));
  RT(rt::stream_sync(st));
  return 0;
}


}  // extern "C" (reopened below)

// ------------------------------------------------------------------------------------
// simulation driver on the device (kernels in sim.cuh)
// ------------------------------------------------------------------------------------
struct SimGroup {  // the units that share one batched fit handle
  cs_fit* f = nullptr;
  int B = 0, first = 0;
  char* buf = nullptr;  // one allocation: everything below
  css::SimArgs ga{};
  css::MetricArgs ma{};
  double *pred_map = nullptr, *tr_map = nullptr, *tr_var = nullptr, *te_map = nullptr, *te_var = nullptr, *scal = nullptr, *metrics = nullptr;
  int* units = nullptr;
  ~SimGroup() { if (f) cs_fit_destroy(f); if (buf) rt::dev_free(buf); }
};

static int sim_group_create(SimGroup& g, const cs_sim_config* cfg, int B, const int* replicate, const int* restart,
                            const double* dXshared, const double* dzshared) {
  const int n = cfg->n, p = cfg->p;
  const int nte = (int)std::ceil(n * cfg->test_frac), ntr = n - nte;
  cs_fit_config fc{};
  fc.kind = CS_GP; fc.optim = cfg->optim; fc.lr = cfg->lr; fc.beta1 = 0.9; fc.beta2 = 0.999; fc.eps = 1e-8; fc.momentum = 0.0;
  int e = fit_create_impl(&fc, B, ntr, p, nullptr, nullptr, nullptr, nullptr, nullptr, &g.f);
  if (e) return e;
  g.B = B;
  cs_fit* f = g.f;
  // carve the group's buffer
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t dn = sizeof(double) * n, dnp = dn * p;
  const size_t o_units = take(sizeof(int) * 2 * B), o_Xraw = take(dnp * B), o_z = take(dn * B), o_y = take(dn * B), o_y1 = take(dn * B),
               o_y0 = take(dn * B), o_ist = take(sizeof(int) * n * B), o_tri = take(sizeof(int) * ntr * B), o_tei = take(sizeof(int) * nte * B),
               o_Xall = take(dnp * B), o_Xte = take(sizeof(double) * nte * p * B), o_mom = take(sizeof(double) * (2 * p + 2) * B),
               o_pm = take(dn * B), o_trm = take(sizeof(double) * ntr * B), o_trv = take(sizeof(double) * ntr * B),
               o_tem = take(sizeof(double) * nte * B), o_tev = take(sizeof(double) * nte * B), o_scal = take(sizeof(double) * 4 * B),
               o_met = take(sizeof(double) * css::NMETRIC * B);
  if (rt::dev_malloc((void**)&g.buf, off)) return fail(CS_ERR_NOMEM, "simulation buffers: allocation of %zu bytes failed", off);
  auto D = [&](size_t o) { return reinterpret_cast<double*>(g.buf + o); };
  auto I = [&](size_t o) { return reinterpret_cast<int*>(g.buf + o); };
  g.units = I(o_units);
  std::vector<int> hu(2 * (size_t)B);
  for (int b = 0; b < B; ++b) { hu[2 * b] = replicate[b]; hu[2 * b + 1] = restart[b]; }
  RT(rt::h2d(g.units, hu.data(), sizeof(int) * hu.size(), f->st));
  RT(rt::stream_sync(f->st));
  css::SimArgs& a = g.ga;
  a.n = n; a.p = p; a.ntr = ntr; a.nte = nte; a.np_fit = f->np; a.seed = cfg->seed; a.units = g.units;
  a.Xshared = dXshared; a.zshared = dzshared;
  a.Xraw = D(o_Xraw); a.zraw = D(o_z); a.y = D(o_y); a.y1 = D(o_y1); a.y0 = D(o_y0);
  a.is_test = I(o_ist); a.train_idx = I(o_tri); a.test_idx = I(o_tei); a.Xall = D(o_Xall); a.Xte = D(o_Xte); a.moments = D(o_mom);
  f->select(0);
  a.fX = f->X; a.fy = f->y; a.fz = f->z; a.fw = f->w; a.fones = f->ones; a.fparams = f->params; a.bstride = f->bstride; a.lay = f->lay;
  a.lay.bstride = 0;
  g.pred_map = D(o_pm); g.tr_map = D(o_trm); g.tr_var = D(o_trv); g.te_map = D(o_tem); g.te_var = D(o_tev); g.scal = D(o_scal);
  g.metrics = D(o_met);
  css::MetricArgs& m = g.ma;
  m.n = n; m.p = p; m.ntr = ntr; m.nte = nte; m.y = a.y; m.z = a.zraw; m.y1 = a.y1; m.y0 = a.y0; m.train_idx = a.train_idx;
  m.test_idx = a.test_idx; m.moments = a.moments; m.pred_map = g.pred_map; m.tr_map = g.tr_map; m.tr_var = g.tr_var;
  m.te_map = g.te_map; m.te_var = g.te_var; m.scal = g.scal; m.metrics = g.metrics;
  return 0;
}

static int sim_group_generate(SimGroup& g) {
  auto k = css::sim_generate_kernel;
  const int smem = css::sim_smem_bytes(g.ga.n, g.ga.p);
  if (smem > 227 * 1024) return fail(CS_ERR_UNSUPPORTED, "simulation generator: n=%d does not fit the shared-memory staging", g.ga.n);
  CS_SET_SMEM(k, smem);
  ++g_launches;
  CS_LAUNCH(k, g.B, css::SIM_THREADS, smem, g.f->st, g.ga);
  KCHECK();
  return 0;
}

// predict on all rows, treatment on the training and on the test rows of every member (inputs and outputs stay on the
// device), then the metric block of the whole group as one kernel
static int sim_group_post(SimGroup& g, const std::vector<int>& status) {
  cs_fit* f = g.f;
  const css::SimArgs& a = g.ga;
  const int n = a.n, p = a.p, ntr = a.ntr, nte = a.nte;
  ArenaScope scope(thread_arena());
  const int keep = f->cur;
  for (int b = 0; b < g.B; ++b) {
    if (status[g.first + b] != 0) continue;
    f->select(b);
    if (g_arena) g_arena->reset();  // the previous member's scratch is reusable: same stream, launches are ordered
    int e;
    {
      CrossWork W;
      if ((e = cross_common(f, W, n, a.Xall + (size_t)b * n * p, n, a.zraw + (size_t)b * n, true, false, g.pred_map + (size_t)b * n,
                            nullptr, 1.0, 0.0, nullptr))) { f->select(keep); return e; }
    }
    for (int sub = 0; sub < 2; ++sub) {
      CrossWork W;
      const int n2 = sub == 0 ? ntr : nte;
      const double* X2 = sub == 0 ? f->X : a.Xte + (size_t)b * nte * p;
      const long long ld2 = sub == 0 ? f->np : nte;
      double* mp = sub == 0 ? g.tr_map + (size_t)b * ntr : g.te_map + (size_t)b * nte;
      double* vr = sub == 0 ? g.tr_var + (size_t)b * ntr : g.te_var + (size_t)b * nte;
      if ((e = cross_common(f, W, n2, X2, ld2, nullptr, true, true, mp, vr, 0.0, 0.0, nullptr))) { f->select(keep); return e; }
      DevBuf h2;
      if (h2.alloc(sizeof(double) * 2)) { f->select(keep); return fail(CS_ERR_NOMEM, "alloc failed"); }
      if ((e = treat_sumcov_enqueue(f, W, n2, h2.d()))) { f->select(keep); return e; }
      auto k = css::treat_scalars_kernel;
      ++g_launches;
      CS_LAUNCH(k, 1, csk::ETH, 0, f->st, (const double*)mp, n2, (const double*)h2.d(), 0.0, f->n, g.scal + 4 * (size_t)b + 2 * sub);
    }
  }
  f->select(keep);
  auto km = css::sim_metrics_kernel;
  ++g_launches;
  CS_LAUNCH(km, g.B, css::SIM_THREADS, 0, f->st, g.ma);
  KCHECK();
  RT(rt::stream_sync(f->st));
  return 0;
}

extern "C" {

int cs_sim_run(const cs_sim_config* cfg, int nunits, const int* replicate, const int* restart, const double* X, const double* z,
               double* metrics, int* iters, double* lml, int* status) {
  if (!cfg || nunits < 1 || !replicate || !restart || !metrics) return fail(CS_ERR_ARG, "cs_sim_run: null argument");
  if (cfg->n < 8 || cfg->p < 1 || cfg->p > CS_MAX_P || !(cfg->test_frac > 0.0 && cfg->test_frac < 0.9) || cfg->maxiter < 1)
    return fail(CS_ERR_ARG, "cs_sim_run: bad configuration");
  if ((X == nullptr) != (z == nullptr)) return fail(CS_ERR_ARG, "cs_sim_run: X and z are given together or not at all");
  if (need_device()) return CS_ERR_NODEVICE;
  const int n = cfg->n, p = cfg->p;
  const int batch = cfg->batch > 0 ? cfg->batch : 32;
  DevBuf dX, dz;
  if (X) {
    if (dX.alloc(sizeof(double) * (size_t)n * p) || dz.alloc(sizeof(double) * n)) return fail(CS_ERR_NOMEM, "alloc failed");
    RT(rt::h2d(dX.p, X, sizeof(double) * (size_t)n * p, 0));
    RT(rt::h2d(dz.p, z, sizeof(double) * n, 0));
    RT(rt::stream_sync(0));
  }
  std::vector<std::unique_ptr<SimGroup>> groups;
  int e = 0;
  for (int u0 = 0; u0 < nunits; u0 += batch) {
    const int B = std::min(batch, nunits - u0);
    groups.emplace_back(new SimGroup());
    SimGroup& g = *groups.back();
    g.first = u0;
    if ((e = sim_group_create(g, cfg, B, replicate + u0, restart + u0, X ? dX.d() : nullptr, z ? dz.d() : nullptr))) return e;
    if ((e = sim_group_generate(g))) return e;
  }
  std::vector<cs_fit*> fits;
  for (auto& g : groups) fits.push_back(g->f);
  std::vector<int> its(nunits, 0), stat(nunits, 0);
  e = cs_fit_run_many(fits.data(), (int)fits.size(), cfg->maxiter, cfg->tol, its.data(), nullptr);
  if (e < 0 && e != CS_ERR_NONFINITE) return e;   // e > 0 / non-finite: properties of single members, reported through `status`
  for (auto& g : groups) {
    std::vector<int> ms(g->B, 0);
    cs_fit_member_status(g->f, ms.data());
    for (int b = 0; b < g->B; ++b) stat[g->first + b] = ms[b];
  }
  for (auto& g : groups)
    if ((e = sim_group_post(*g, stat))) return e;
  const double nanv = std::nan("");
  for (auto& g : groups) {
    RT(rt::d2h(metrics + (size_t)g->first * css::NMETRIC, g->metrics, sizeof(double) * css::NMETRIC * g->B, g->f->st));
    RT(rt::stream_sync(g->f->st));
    for (int b = 0; b < g->B; ++b) {
      const int u = g->first + b;
      if (stat[u] != 0) for (int k = 0; k < css::NMETRIC; ++k) metrics[(size_t)u * css::NMETRIC + k] = nanv;
      if (lml) {
        csk::EvalScalars h{};
        g->f->select(b);
        RT(rt::d2h(&h, g->f->sc, sizeof h, g->f->st));
        RT(rt::stream_sync(g->f->st));
        lml[u] = stat[u] == 0 ? h.stats[1] : -HUGE_VAL;
      }
    }
  }
  if (iters) std::copy(its.begin(), its.end(), iters);
  if (status) std::copy(stat.begin(), stat.end(), status);
  return 0;
}

// test hook: the generator alone for ONE unit -- everything sim_generate_kernel produces, downloaded
int cs_sim_generate(const cs_sim_config* cfg, int replicate, int restart, const double* X, const double* z, double* Xraw,
                    double* zraw, double* y, double* y1, double* y0, int* is_test, double* Xtrain, double* ytrain, double* ztrain,
                    double* Xall, double* moments, double* params) {
  if (!cfg) return fail(CS_ERR_ARG, "cs_sim_generate: null argument");
  if ((X == nullptr) != (z == nullptr)) return fail(CS_ERR_ARG, "cs_sim_generate: X and z are given together or not at all");
  if (need_device()) return CS_ERR_NODEVICE;
  const int n = cfg->n, p = cfg->p;
  DevBuf dX, dz;
  if (X) {
    if (dX.alloc(sizeof(double) * (size_t)n * p) || dz.alloc(sizeof(double) * n)) return fail(CS_ERR_NOMEM, "alloc failed");
    RT(rt::h2d(dX.p, X, sizeof(double) * (size_t)n * p, 0));
    RT(rt::h2d(dz.p, z, sizeof(double) * n, 0));
    RT(rt::stream_sync(0));
  }
  SimGroup g;
  int e = sim_group_create(g, cfg, 1, &replicate, &restart, X ? dX.d() : nullptr, z ? dz.d() : nullptr);
  if (e) return e;
  if ((e = sim_group_generate(g))) return e;
  cs_fit* f = g.f;
  cudaStream_t st = f->st;
  const css::SimArgs& a = g.ga;
  const size_t dn = sizeof(double) * n;
  if (Xraw) RT(rt::d2h(Xraw, a.Xraw, dn * p, st));
  if (zraw) RT(rt::d2h(zraw, a.zraw, dn, st));
  if (y) RT(rt::d2h(y, a.y, dn, st));
  if (y1) RT(rt::d2h(y1, a.y1, dn, st));
  if (y0) RT(rt::d2h(y0, a.y0, dn, st));
  if (is_test) RT(rt::d2h(is_test, a.is_test, sizeof(int) * n, st));
  if (Xall) RT(rt::d2h(Xall, a.Xall, dn * p, st));
  if (moments) RT(rt::d2h(moments, a.moments, sizeof(double) * (2 * p + 2), st));
  if (Xtrain) RT(download_matrix(Xtrain, a.ntr, p, f->X, f->np, st));
  if (ytrain) RT(rt::d2h(ytrain, f->y, sizeof(double) * a.ntr, st));
  if (ztrain) RT(rt::d2h(ztrain, f->z, sizeof(double) * a.ntr, st));
  if (params) RT(rt::d2h(params, f->params, sizeof(double) * f->lay.np(), st));
  RT(rt::stream_sync(st));
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------
// config 5: distributed evaluation over a P x Q process grid
// ------------------------------------------------------------------------------------

#ifndef CS_EMU
// NCCL is resolved at run time (dlopen) so the library has no link-time dependency on it and
// binds to whichever libnccl.so.2 the process already loaded (e.g. the one bundled with torch).
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  typedef void* Comm_t;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm_t*, int, UniqueId, int) = nullptr;
  int (*CommSplit)(Comm_t, int, int, Comm_t*, void*) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, Comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, Comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(Comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* h = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return false;
#define SYM(f, name) *(void**)(&f) = dlsym(h, name); if (!f) return false
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommSplit, "ncclCommSplit");
    SYM(Broadcast, "ncclBroadcast"); SYM(AllGather, "ncclAllGather"); SYM(AllReduce, "ncclAllReduce");
    SYM(CommDestroy, "ncclCommDestroy"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
  }
};
static NcclApi g_nccl;
enum { NCCL_DOUBLE = 8, NCCL_SUM = 0 };

struct NcclComm : csdist::Comm {
  NcclApi::Comm_t world = nullptr, row = nullptr, col = nullptr;
  int err = 0;
  int chk(int e) { if (e != 0 && err == 0) err = e; return e; }
  int bcast(std::vector<csdist::DistRank*>& L, int group, int which, int root, const csdist::BufFn& buf, size_t count) override {
    csdist::DistRank& R = *L[0];
    const int mine = group == csdist::G_ROW ? R.g.pr : R.g.pc;
    if (which >= 0 && which != mine) return 0;
    const int gsize = group == csdist::G_ROW ? R.g.Q : R.g.P;
    if (gsize == 1) return 0;
    double* b = buf(R);
    return chk(g_nccl.Broadcast(b, b, count, NCCL_DOUBLE, root, group == csdist::G_ROW ? row : col, R.st));
  }
  int allgather(std::vector<csdist::DistRank*>& L, int group, int which, const csdist::BufFn& send, const csdist::BufFn& recv,
                size_t count) override {
    csdist::DistRank& R = *L[0];
    const int mine = group == csdist::G_ROW ? R.g.pr : R.g.pc;
    if (which >= 0 && which != mine) return 0;
    const int gsize = group == csdist::G_ROW ? R.g.Q : R.g.P;
    if (gsize == 1) return rt::d2d(recv(R), send(R), sizeof(double) * count, R.st);
    return chk(g_nccl.AllGather(send(R), recv(R), count, NCCL_DOUBLE, group == csdist::G_ROW ? row : col, R.st));
  }
  int allreduce_sum(std::vector<csdist::DistRank*>& L, const csdist::BufFn& buf, size_t) override {
    csdist::DistRank& R = *L[0];
    double* b = buf(R);
    return chk(g_nccl.AllReduce(b, b, (size_t)R.red_len, NCCL_DOUBLE, NCCL_SUM, world, R.st));
  }
  // share_panel: the default (all-gather inside the owning group, broadcast across the other dimension) is kept.  A
  // fused variant -- one ncclGroup of `pieces` world broadcasts -- was measured SLOWER on 8 B200s (253 vs 241 ms at
  // n = 32768): the two small sub-communicator collectives beat one world-wide group.
  ~NcclComm() override {
    if (row) g_nccl.CommDestroy(row);
    if (col) g_nccl.CommDestroy(col);
    if (world) g_nccl.CommDestroy(world);
  }
};
#endif

struct cs_dist {
  std::vector<csdist::DistRank*> ranks;  // 1 (NCCL) or P*Q (in-process)
  std::unique_ptr<csdist::Comm> comm;
  csdist::DistEval ev;
  cudaEvent_t e0{}, e1{};
  bool ev_ok = false;
  ~cs_dist() {
    for (auto* r : ranks) { r->destroy(); delete r; }
    if (ev_ok) { rt::event_destroy(e0); rt::event_destroy(e1); }
  }
};

static int dist_fetch(cs_dist* d, double* grad, double* stats, double* mu_solution) {
  csdist::DistRank& R = *d->ranks[0];
  csk::EvalScalars h{};
  RT(rt::d2h(&h, R.sc, sizeof h, R.st));
  if (grad) RT(rt::d2h(grad, R.grad, sizeof(double) * R.lay.np(), R.st));
  int status = 0;
  RT(rt::d2h(&status, R.status, sizeof(int), R.st));
  RT(rt::stream_sync(R.st));
  if (stats) { stats[0] = h.stats[0]; stats[1] = h.stats[1]; }
  if (mu_solution) *mu_solution = h.mu_sol;
  if (status != 0x7fffffff && status > 0)
    return fail(status, "inv_sympd(): matrix is singular or not positive definite (pivot %d)", status);
  return 0;
}

extern "C" {

int cs_dist_unique_id(char* id128) {
#ifdef CS_EMU
  (void)id128;
  return fail(CS_ERR_UNSUPPORTED, "NCCL is not available in the emulator build");
#else
  if (!id128) return fail(CS_ERR_ARG, "null id");
  if (!g_nccl.load()) return fail(CS_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded: %s", dlerror());
  NcclApi::UniqueId id;
  int e = g_nccl.GetUniqueId(&id);
  if (e) return fail(CS_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString(e));
  std::memcpy(id128, id.internal, 128);
  return 0;
#endif
}

int cs_dist_create(int kind, int n, int p, const double* y, const double* X, const double* z, const double* w,
                   const double* params, int rank, int world, int P, int Q, const char* id128, cs_dist** out) {
#ifdef CS_EMU
  (void)kind; (void)n; (void)p; (void)y; (void)X; (void)z; (void)w; (void)params; (void)rank; (void)world; (void)P; (void)Q; (void)id128; (void)out;
  return fail(CS_ERR_UNSUPPORTED, "NCCL is not available in the emulator build");
#else
  if (!y || !X || !z || !params || !id128 || !out) return fail(CS_ERR_ARG, "cs_dist_create: null argument");
  if (P * Q != world || rank < 0 || rank >= world) return fail(CS_ERR_ARG, "cs_dist_create: P*Q must equal world");
  if (p > 40) return fail(CS_ERR_UNSUPPORTED, "cs_dist_create: p=%d exceeds 40, the limit of the distributed engine", p);
  if (need_device()) return CS_ERR_NODEVICE;
  if (!g_nccl.load()) return fail(CS_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  std::unique_ptr<cs_dist> d(new cs_dist());
  auto* nc = new NcclComm();
  d->comm.reset(nc);
  NcclApi::UniqueId id;
  std::memcpy(id.internal, id128, 128);
  int e = g_nccl.CommInitRank(&nc->world, world, id, rank);
  if (e) return fail(CS_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(e));
  const int pr = rank / Q, pc = rank % Q;
  if ((e = g_nccl.CommSplit(nc->world, pr, pc, &nc->row, nullptr))) return fail(CS_ERR_CUDA, "ncclCommSplit(row): %s", g_nccl.GetErrorString(e));
  if ((e = g_nccl.CommSplit(nc->world, pc, pr, &nc->col, nullptr))) return fail(CS_ERR_CUDA, "ncclCommSplit(col): %s", g_nccl.GetErrorString(e));
  cudaStream_t st;
  RT(rt::stream_create(&st));
  auto* R = new csdist::DistRank();
  d->ranks.push_back(R);
  if ((e = csdist::rank_init(*R, n, p, kind, 128, P, Q, pr, pc, y, X, z, w, params, st, true)))
    return fail(e == 2 ? CS_ERR_NOMEM : CS_ERR_CUDA, "distributed rank init failed: %s", rt::err_str(e));
  d->ev.L = d->ranks; d->ev.comm = d->comm.get(); d->ev.T = 128; d->ev.N = R->g.N; d->ev.P = P; d->ev.Q = Q; d->ev.n = n;
  RT(rt::event_create(&d->e0)); RT(rt::event_create(&d->e1)); d->ev_ok = true;
  *out = d.release();
  return 0;
#endif
}

int cs_dist_eval(cs_dist* d, const double* params, double* grad, double* stats, double* mu_solution, float* ms) {
  if (!d) return fail(CS_ERR_ARG, "null handle");
  for (auto* R : d->ranks)
    if (params) RT(rt::h2d(R->params, params, sizeof(double) * R->lay.np(), R->st));
  cudaStream_t st = d->ranks[0]->st;
  if (d->ev_ok) rt::event_record(d->e0, st);
  int e = d->ev.run();
  if (e) return fail(CS_ERR_CUDA, "distributed evaluation failed (%d): %s", e, rt::err_str(rt::last_error()));
  if (d->ev_ok) rt::event_record(d->e1, st);
  KCHECK();
  e = dist_fetch(d, grad, stats, mu_solution);
  if (ms && d->ev_ok) { *ms = 0.f; rt::event_elapsed(ms, d->e0, d->e1); }
  return e;
}

void cs_dist_destroy(cs_dist* d) { delete d; }

// All P*Q ranks inside this process (one device / the emulator), collectives as device
// copies: exercises the identical driver, tables and kernels as the NCCL path.
int cs_dist_sim_eval(int kind, int n, int p, int T, int P, int Q, const double* y, const double* X, const double* z,
                     const double* w, const double* params, double* grad, double* stats, double* mu_solution) {
  if (!y || !X || !z || !params || P < 1 || Q < 1 || (T != 64 && T != 128)) return fail(CS_ERR_ARG, "cs_dist_sim_eval: bad argument");
  if (need_device()) return CS_ERR_NODEVICE;
  cs_dist d;
  d.comm.reset(new csdist::SimComm(P, Q));
  cudaStream_t st = 0;
  for (int pr = 0; pr < P; ++pr)
    for (int pc = 0; pc < Q; ++pc) {
      auto* R = new csdist::DistRank();
      d.ranks.push_back(R);
      if (int e = csdist::rank_init(*R, n, p, kind, T, P, Q, pr, pc, y, X, z, w, params, st, false))
        return fail(e == 2 ? CS_ERR_NOMEM : CS_ERR_CUDA, "rank init failed: %s", rt::err_str(e));
    }
  d.ev.L = d.ranks; d.ev.comm = d.comm.get(); d.ev.T = T; d.ev.N = d.ranks[0]->g.N; d.ev.P = P; d.ev.Q = Q; d.ev.n = n;
  int e = d.ev.run();
  if (e) return fail(CS_ERR_CUDA, "simulated distributed evaluation failed (%d)", e);
  KCHECK();
  e = dist_fetch(&d, grad, stats, mu_solution);
  if (e) return e;
  // every rank must hold the same answer
  const int NP = d.ranks[0]->lay.np();
  std::vector<double> g0(NP), gi(NP);
  RT(rt::d2h(g0.data(), d.ranks[0]->grad, sizeof(double) * NP, st));
  RT(rt::stream_sync(st));
  for (size_t r = 1; r < d.ranks.size(); ++r) {
    RT(rt::d2h(gi.data(), d.ranks[r]->grad, sizeof(double) * NP, st));
    RT(rt::stream_sync(st));
    for (int i = 0; i < NP; ++i)
      if (!(std::fabs(gi[i] - g0[i]) <= 1e-12 * (std::fabs(g0[i]) + 1e-300) + 1e-300))
        return fail(CS_ERR_CUDA, "rank %d disagrees with rank 0 on gradient entry %d", (int)r, i);
  }
  return 0;
}

}  // extern "C"
