// SIMT emulator runtime (see emu.h).  x86-64 only; single OS thread.
#include "emu.h"

#include <sys/mman.h>

#include <random>
#include <vector>

emu_uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

unsigned char* dyn_smem = nullptr;
long n_launches = 0;

enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };

struct Fiber {
  void* sp;
  State st;
  int tid, lane, warp;
  int parity;
  emu_uint3 tidx;
};

static constexpr size_t kStack = 256 * 1024;
static constexpr int kMaxThreads = 1024;
static unsigned char* g_stacks = nullptr;
static Fiber g_fib[kMaxThreads];
static void* g_sched_sp = nullptr;
static Fiber* g_cur = nullptr;
static const std::function<void()>* g_fn = nullptr;
static int g_nthreads = 0;
alignas(16) static unsigned char g_xch[kMaxThreads / 32][2][32][16];
static std::vector<unsigned char> g_smem;

int lane() { return g_cur->lane; }
void* xch_slot(int parity, int l) { return g_xch[g_cur->warp][parity][l]; }
int next_parity() { int p = g_cur->parity; g_cur->parity ^= 1; return p; }

static void yield_to_sched() { emu_switch(&g_cur->sp, g_sched_sp); }

void barrier_block() {
  g_cur->st = WAIT_BLOCK;
  yield_to_sched();
}
void barrier_warp() {
  g_cur->st = WAIT_WARP;
  yield_to_sched();
}

static void fiber_entry() {
  (*g_fn)();
  g_cur->st = DONE;
  yield_to_sched();
  std::fprintf(stderr, "emu: resumed a finished fiber\n");
  std::abort();
}

static void init_fiber(Fiber& f, int tid, const dim3& block) {
  f.tid = tid;
  f.lane = tid & 31;
  f.warp = tid >> 5;
  f.parity = 0;
  f.st = READY;
  f.tidx.x = tid % block.x;
  f.tidx.y = (tid / block.x) % block.y;
  f.tidx.z = tid / (block.x * block.y);
  unsigned char* top = g_stacks + (size_t)(tid + 1) * kStack;
  void** sp = reinterpret_cast<void**>(top);
  sp -= 2;                                   // keep (top-16): after `ret`, rsp = top-8 == 8 (mod 16)
  sp[0] = reinterpret_cast<void*>(&fiber_entry);
  sp -= 6;                                   // r15 r14 r13 r12 rbx rbp
  for (int i = 0; i < 6; ++i) sp[i] = nullptr;
  f.sp = sp;
}

static void resume(Fiber& f) {
  g_cur = &f;
  threadIdx = f.tidx;
  f.st = READY;
  emu_switch(&g_sched_sp, f.sp);
}

static void run_block(std::mt19937& rng, bool shuffle) {
  const int nwarps = (g_nthreads + 31) / 32;
  std::vector<int> order(nwarps);
  for (int i = 0; i < nwarps; ++i) order[i] = i;
  for (;;) {
    if (shuffle) std::shuffle(order.begin(), order.end(), rng);
    bool progress = false;
    for (int wi = 0; wi < nwarps; ++wi) {
      int w = order[wi];
      int lo = w * 32, hi = std::min(g_nthreads, lo + 32);
      for (;;) {
        bool ran = false;
        for (int t = lo; t < hi; ++t)
          if (g_fib[t].st == READY) { resume(g_fib[t]); ran = true; progress = true; }
        // warp barrier release: all non-done lanes waiting at a warp barrier
        int nwait = 0, nlive = 0, nblock = 0;
        for (int t = lo; t < hi; ++t) {
          if (g_fib[t].st != DONE) ++nlive;
          if (g_fib[t].st == WAIT_WARP) ++nwait;
          if (g_fib[t].st == WAIT_BLOCK) ++nblock;
        }
        if (nlive > 0 && nwait == nlive) {
          for (int t = lo; t < hi; ++t) if (g_fib[t].st == WAIT_WARP) g_fib[t].st = READY;
          continue;
        }
        if (nwait > 0 && nwait + nblock == nlive && nblock > 0) {
          std::fprintf(stderr, "emu: divergent warp/block barrier in warp %d (deadlock)\n", w);
          std::abort();
        }
        if (!ran) break;
      }
    }
    int nlive = 0, nblock = 0;
    for (int t = 0; t < g_nthreads; ++t) {
      if (g_fib[t].st != DONE) ++nlive;
      if (g_fib[t].st == WAIT_BLOCK) ++nblock;
    }
    if (nlive == 0) return;
    if (nblock == nlive) {
      for (int t = 0; t < g_nthreads; ++t) if (g_fib[t].st == WAIT_BLOCK) g_fib[t].st = READY;
      continue;
    }
    if (!progress) {
      std::fprintf(stderr, "emu: deadlock (%d live, %d at block barrier)\n", nlive, nblock);
      std::abort();
    }
  }
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& fn) {
  ++n_launches;
  if (!g_stacks) {
    g_stacks = static_cast<unsigned char*>(mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE,
                                                MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
    if (g_stacks == MAP_FAILED) { std::perror("emu mmap"); std::abort(); }
  }
  g_nthreads = (int)(block.x * block.y * block.z);
  if (g_nthreads > kMaxThreads || g_nthreads <= 0) { std::fprintf(stderr, "emu: bad block size\n"); std::abort(); }
  if (g_smem.size() < smem_bytes + 64) g_smem.resize(smem_bytes + 64);
  dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(g_smem.data()) + 15) & ~uintptr_t(15));
  blockDim = block;
  gridDim = grid;
  g_fn = &fn;
  static std::mt19937 rng(12345);
  static const bool shuffle = std::getenv("CS_EMU_SHUFFLE") != nullptr;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        // poison dynamic shared memory so reads of unwritten smem show up as NaN
        std::memset(dyn_smem, 0xFF, smem_bytes);
        for (int t = 0; t < g_nthreads; ++t) init_fiber(g_fib[t], t, block);
        run_block(rng, shuffle);
      }
  g_fn = nullptr;
}

}  // namespace emu
