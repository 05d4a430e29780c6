// Config 5 (BASELINE.json): one large GP evaluation over a P x Q process grid with a 2-D
// block-cyclic tile distribution and NCCL panel broadcasts over NVLink.
//
// Tile (I,J) of the np x np matrices (block T = 128) lives on rank (I mod P, J mod Q) at
// local tile (I/P, J/Q).  Three right-looking sweeps share one communication pattern --
// "finish panel k, make it visible to every rank, rank-T update of the local tiles":
//   1. Cholesky:   leaf(k,k) -> bcast X_kk down the process column -> L_Ik = A_Ik X_kk^T
//                  -> column panel k to everyone (all-gather in the column + bcast along
//                  rows) -> A_IJ -= L_Ik L_Jk^T on the local lower tiles with J > k.
//   2. inverse:    X_kJ = -X_kk Acc_kJ (row k), row panel k of X and column panel k of L to
//                  everyone, then Acc_IJ += L_Ik X_kJ (I > k, J <= k) and, fused in the same
//                  step, C_IJ += X_kI^T X_kJ (J <= I <= k), which is K^-1 = X^T X.
//   3. reductions: mat-vec, trace gradients and K*alpha on the local tiles, one
//                  all-reduce of a few small vectors.
// Every rank builds its own Gram tiles from the replicated X (no communication for K1).
// Local updates are table-driven launches of the same DMMA kernel as the single-GPU path.
//
// The driver loop is written over a *set* of local rank objects and an abstract Comm, so
// the identical code runs (a) one rank per process over NCCL, and (b) all ranks inside one
// process with copy-based collectives -- used to test the multi-rank algorithm on one
// device / on the SIMT emulator (tests/test_dist_sim.py), as the profiling guide asks.
#pragma once
#include <functional>
#include <memory>
#include <vector>

#include "dense.cuh"
#include "kernels.cuh"

namespace csdist {

using csdense::g_launches;
using csg::GemmProblem;

enum Group { G_ROW = 0, G_COL = 1, G_WORLD = 2 };

struct Geom {
  int n = 0, np = 0, T = 128, N = 0, P = 1, Q = 1, pr = 0, pc = 0;
  int mt = 0, nt = 0;   // local tile rows / cols (same on every rank: ceil)
  long long ldl = 0;    // local leading dimension = mt * T
  int li(int I) const { return I / P; }
  int lj(int J) const { return J / Q; }
  bool owns(int I, int J) const { return I % P == pr && J % Q == pc; }
  long long off(int I, int J) const { return (long long)li(I) * T + (long long)lj(J) * T * ldl; }
};

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
// copy `ntiles` T x T tiles: tile t from src + t*src_stride (ld = lds) to dst + t*T*T (ld = T), or back
static __global__ void pack_tiles_kernel(const double* __restrict__ src, long long lds, long long src_stride,
                                         double* __restrict__ dst, int T, int ntiles) {
  const long long tot = (long long)ntiles * T * T;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx / ((long long)T * T));
    const int e = (int)(idx - (long long)t * T * T);
    const int r = e % T, c = e / T;
    dst[idx] = src[(long long)t * src_stride + r + (long long)c * lds];
  }
}

// Gram tiles of this rank: entry e = (I, J) of `tiles` (global 128-tile coords, I >= J),
// four 64 x 64 sub-blocks each; identity padding; zeros above the diagonal.
static __global__ void __launch_bounds__(csk::ETH)
gram_dist_kernel(const double* __restrict__ X, long long ldx, const double* __restrict__ z,
                 const double* __restrict__ w, const double* __restrict__ params, csk::Layout lay, int n,
                 const int* __restrict__ tiles, int T, int P, int Q, double* __restrict__ G, long long ldg) {
  using namespace csk;
  CS_DYN_SMEM(double, sm_);
  const int p = lay.p;
  double* xr = sm_;
  double* xc = xr + p * TB;
  double* ea = xc + p * TB;
  double* eb = ea + PMAX;
  double* zr = eb + PMAX;
  double* zc = zr + TB;
  double* wr = zc + TB;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int sub = T / TB;  // sub-blocks per tile edge
  const int e = (int)blockIdx.x / (sub * sub), s = (int)blockIdx.x % (sub * sub);
  const int I = tiles[2 * e], J = tiles[2 * e + 1];
  const int sr = s % sub, sc = s / sub;
  const int r0 = I * T + sr * TB, c0 = J * T + sc * TB;
  load_x_tile(xr, X, ldx, r0, p, tid);
  load_x_tile(xc, X, ldx, c0, p, tid);
  if (tid < p) { ea[tid] = exp(-params[lay.i_Lm() + tid]); eb[tid] = exp(-params[lay.i_La() + tid]); }
  if (tid < TB) { zr[tid] = z[r0 + tid]; zc[tid] = z[c0 + tid]; wr[tid] = w[r0 + tid]; }
  __syncthreads();
  const double lam_m = params[lay.i_lm()], lam_a = params[lay.i_la()];
  const double esig = exp(params[lay.i_sigma()]);
  double sm[4][4], sa[4][4];
  sqdist_4x4(xr, xc, ea, eb, p, tx, ty, sm, sa);
  double* Gt = G + (long long)(I / P) * T + (long long)(J / Q) * T * ldg + sr * TB + (long long)sc * TB * ldg;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int cl = ty + 16 * b, gc = c0 + cl;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int rl = tx + 16 * a, gr = r0 + rl;
      double v;
      if (gr >= n || gc >= n) v = (gr == gc) ? 1.0 : 0.0;
      else if (gr < gc) v = 0.0;
      else {
        v = exp(lam_m - sm[a][b]) + exp(lam_a - sa[a][b]) * zr[rl] * zc[cl];
        if (gr == gc) v += esig / wr[rl];
      }
      Gt[rl + (long long)cl * ldg] = v;
    }
  }
}

// symmetric mat-vec on the local lower tiles: out[k][rows I] += T v_k[J];  (I != J) out[k][rows J] += T^T v_k[I]
// one CTA per tile; three vectors; fp64 atomics (distributed mode only).
static __global__ void __launch_bounds__(256)
symv_dist_kernel(const double* __restrict__ C, long long ldc, const int* __restrict__ tiles, int T, int P, int Q,
                 const double* __restrict__ v0, const double* __restrict__ shift0, const double* __restrict__ v1,
                 const double* __restrict__ v2, double* __restrict__ out, long long out_stride) {
  __shared__ double sv[3][128];
  const int e = (int)blockIdx.x;
  const int I = tiles[2 * e], J = tiles[2 * e + 1];
  const double* Ct = C + (long long)(I / P) * T + (long long)(J / Q) * T * ldc;
  const int tid = threadIdx.x;
  const double sh = shift0 ? *shift0 : 0.0;
  // rows of I:  sum_c Ct[r,c] v[J*T + c]
  for (int c = tid; c < T; c += 256) {
    sv[0][c] = v0[J * T + c] - sh; sv[1][c] = v1[J * T + c]; sv[2][c] = v2[J * T + c];
  }
  __syncthreads();
  if (tid < T) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int c = 0; c < T; ++c) {
      const double t = Ct[tid + (long long)c * ldc];
      a0 = fma(t, sv[0][c], a0); a1 = fma(t, sv[1][c], a1); a2 = fma(t, sv[2][c], a2);
    }
    atomicAdd(out + 0 * out_stride + I * T + tid, a0);
    atomicAdd(out + 1 * out_stride + I * T + tid, a1);
    atomicAdd(out + 2 * out_stride + I * T + tid, a2);
  }
  if (I != J) {
    // rows of J:  sum_r Ct[r,c] v[I*T + r]   (thread per column c, coalesced over r via loop)
    __syncthreads();
    for (int r = tid; r < T; r += 256) { sv[0][r] = v0[I * T + r] - sh; sv[1][r] = v1[I * T + r]; sv[2][r] = v2[I * T + r]; }
    __syncthreads();
    if (tid < T) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      const double* col = Ct + (long long)tid * ldc;
      for (int r = 0; r < T; ++r) {
        const double t = col[r];
        a0 = fma(t, sv[0][r], a0); a1 = fma(t, sv[1][r], a1); a2 = fma(t, sv[2][r], a2);
      }
      atomicAdd(out + 0 * out_stride + J * T + tid, a0);
      atomicAdd(out + 1 * out_stride + J * T + tid, a1);
      atomicAdd(out + 2 * out_stride + J * T + tid, a2);
    }
  }
}

// sigma-trace partial over the diagonal tiles this rank owns: sum_r (C_rr - coef a_r^2) e^sigma / w_r
static __global__ void __launch_bounds__(256)
diag_trace_dist_kernel(const double* __restrict__ C, long long ldc, const int* __restrict__ diag_tiles, int ndiag,
                       int T, int P, int Q, const double* __restrict__ alpha, const double* __restrict__ w,
                       const double* __restrict__ params, csk::Layout lay, int n,
                       const csk::EvalScalars* __restrict__ sc, double* __restrict__ out) {
  __shared__ double red[32];
  const double coef = sc->coef, esig = exp(params[lay.i_sigma()]);
  double s = 0.0;
  for (int idx = threadIdx.x; idx < ndiag * T; idx += 256) {
    const int I = diag_tiles[idx / T], r = idx % T, gr = I * T + r;
    if (gr < n) {
      const double* Ct = C + (long long)(I / P) * T + (long long)(I / Q) * T * ldc;
      s += (Ct[r + (long long)r * ldc] - coef * alpha[gr] * alpha[gr]) * (esig / w[gr]);
    }
  }
  s = csd::block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

// squared residual from the all-reduced K*alpha (replicated on every rank); rs_part = {tsig, res}
static __global__ void __launch_bounds__(256)
resid_dist_kernel(const double* __restrict__ kal, const double* __restrict__ y, const double* __restrict__ params,
                  csk::Layout lay, int n, const double* __restrict__ tsig, double* __restrict__ rs_part) {
  __shared__ double red[32];
  const double mu = params[lay.i_mu()];
  double res = 0.0;
  for (int r = threadIdx.x; r < n; r += 256) { const double e = y[r] - (kal[r] + mu); res = fma(e, e, res); }
  res = csd::block_sum(res, red);
  if (threadIdx.x == 0) { rs_part[0] = tsig[0]; rs_part[1] = res; }
}

// ---------------------------------------------------------------------------------
// communication back-ends
// ---------------------------------------------------------------------------------
struct DistRank;
typedef std::function<double*(DistRank&)> BufFn;

struct Comm {
  virtual ~Comm() {}
  // `which`: index of the participating group (row index for G_ROW, column index for G_COL), -1 = all groups
  virtual int bcast(std::vector<DistRank*>& L, int group, int which, int root, const BufFn& buf, size_t count) = 0;
  virtual int allgather(std::vector<DistRank*>& L, int group, int which, const BufFn& send, const BufFn& recv, size_t count) = 0;
  virtual int allreduce_sum(std::vector<DistRank*>& L, const BufFn& buf, size_t count) = 0;
  // Panel sharing: the ranks of one process column (group == G_COL, index `which`) or process row (G_ROW) each hold
  // `count` doubles in send(); afterwards EVERY rank of the grid holds all pieces in recv(), piece r at r * count.
  // Default: all-gather inside the owning group, then broadcast across the other dimension (two dependent
  // collectives); back-ends may fuse it into one.
  virtual int share_panel(std::vector<DistRank*>& L, int group, int which, const BufFn& send, const BufFn& recv, size_t count,
                          int pieces) {
    if (int e = allgather(L, group, which, send, recv, count)) return e;
    return bcast(L, group == G_COL ? G_ROW : G_COL, -1, which, recv, (size_t)pieces * count);
  }
};

// ---------------------------------------------------------------------------------
// one rank's state
// ---------------------------------------------------------------------------------
struct StepTab { std::vector<long long> off; std::vector<int> start; long long* d = nullptr; };
constexpr int NPK = 6;  // problems per step

struct DistRank {
  Geom g;
  csk::Layout lay{};
  cudaStream_t st{};
  bool own_stream = false;
  int gt = 64;  // CTA tile of the update GEMMs
  // replicated small data
  double *X = nullptr, *y = nullptr, *z = nullptr, *w = nullptr, *ones = nullptr, *params = nullptr, *grad = nullptr;
  // local tiles
  double *G = nullptr, *Xl = nullptr, *C = nullptr;
  // panels (tile-contiguous, T x T each)
  double *colsend = nullptr, *COL = nullptr, *rowsend = nullptr, *ROW = nullptr, *xkk = nullptr;
  // look-ahead: the bulk of step k's update runs on the side stream st2 while the main stream already does the
  // leaf / panel / communication of step k+1, so the gathered panels are double-buffered by step parity
  double *COLb[2] = {nullptr, nullptr}, *ROWb[2] = {nullptr, nullptr};
  cudaStream_t st2{};
  bool own_st2 = false;
  cudaEvent_t ev_share{}, ev_side{};
  bool ev_ok = false;
  // reductions
  double *mv = nullptr, *alpha = nullptr, *u = nullptr, *s = nullptr, *logdet_part = nullptr, *gpart = nullptr,
         *kal = nullptr, *tsig = nullptr, *rs_part = nullptr, *red = nullptr;
  int* status = nullptr;
  csk::EvalScalars* sc = nullptr;
  int ncta_grad = 1;
  // tables
  std::vector<int> lower_tiles;   // (I,J) pairs owned, I >= J
  std::vector<int> diag_tiles;    // I of owned diagonal tiles
  int *d_lower = nullptr, *d_diag = nullptr;
  int* d_k3 = nullptr; int n_k3 = 0;  // (I64, J64, coff_lo, coff_hi) per 64 x 64 lower sub-tile
  StepTab t_potrf, t_acc, t_c, t_fin;
  std::vector<int> na_potrf, na_acc;  // entries of step k's table that the NEXT step depends on (they come first)
  GemmProblem* d_probs = nullptr;   // [NPK * N] per step: potrf (next column), fin, acc (next row), c, potrf rest, acc rest
  std::vector<GemmProblem> probs;
  GemmProblem* d_panel = nullptr;   // [N] panel multiply problems
  std::vector<void*> allocs;
  int red_len = 0;

  int dalloc(void** p, size_t bytes) { int e = rt::dev_malloc(p, bytes); if (!e) allocs.push_back(*p); return e; }
  void destroy() {
    rt::stream_sync(st);
    if (own_st2) { rt::stream_sync(st2); rt::stream_destroy(st2); own_st2 = false; }
    if (ev_ok) { rt::event_destroy(ev_share); rt::event_destroy(ev_side); ev_ok = false; }
    for (void* p : allocs) rt::dev_free(p);
    allocs.clear();
    if (own_stream) rt::stream_destroy(st);
  }
  int rank() const { return g.pr * g.Q + g.pc; }
};

static int upload_tab(DistRank& R, StepTab& t) {
  const size_t b = sizeof(long long) * std::max<size_t>(t.off.size(), 1);
  if (int e = R.dalloc((void**)&t.d, b)) return e;
  return rt::h2d(t.d, t.off.data(), sizeof(long long) * t.off.size(), R.st);
}

// builds all static tables / problems of one rank
static int rank_init(DistRank& R, int n, int p, int kind, int T, int P, int Q, int pr, int pc, const double* y,
                     const double* X, const double* z, const double* w, const double* params, cudaStream_t st,
                     bool own_stream) {
  Geom& g = R.g;
  g.n = n; g.T = T; g.N = (n + T - 1) / T; g.np = g.N * T; g.P = P; g.Q = Q; g.pr = pr; g.pc = pc;
  g.mt = (g.N + P - 1) / P; g.nt = (g.N + Q - 1) / Q; g.ldl = (long long)g.mt * T;
  R.lay.kind = kind; R.lay.p = p;
  R.st = st; R.own_stream = own_stream;
  R.gt = 64;
  const int N = g.N, np = g.np, NP = R.lay.np();
  const size_t T2 = (size_t)T * T, vb = sizeof(double) * np;
  const size_t lm = sizeof(double) * (size_t)g.ldl * g.nt * T;
  int e;
#define DA(ptr, bytes) if ((e = R.dalloc((void**)&(ptr), (bytes)))) return e
  DA(R.X, vb * p); DA(R.y, vb); DA(R.z, vb); DA(R.w, vb); DA(R.ones, vb); DA(R.params, sizeof(double) * NP);
  DA(R.grad, sizeof(double) * NP);
  DA(R.G, lm); DA(R.Xl, lm); DA(R.C, lm);
  DA(R.colsend, sizeof(double) * T2 * g.mt); DA(R.COLb[0], sizeof(double) * T2 * g.mt * P); DA(R.COLb[1], sizeof(double) * T2 * g.mt * P);
  DA(R.rowsend, sizeof(double) * T2 * g.nt); DA(R.ROWb[0], sizeof(double) * T2 * g.nt * Q); DA(R.ROWb[1], sizeof(double) * T2 * g.nt * Q);
  R.COL = R.COLb[0]; R.ROW = R.ROWb[0];
  if (own_stream) {
    if ((e = rt::stream_create_low(&R.st2))) return e;
    R.own_st2 = true;
  } else {
    R.st2 = st;  // in-process grid (tests): one stream for everything, same op order
  }
  if ((e = rt::event_create_fast(&R.ev_share)) || (e = rt::event_create_fast(&R.ev_side))) return e;
  R.ev_ok = true;
  DA(R.xkk, sizeof(double) * T2);
  R.ncta_grad = rt::sm_count();
  const int nacc = 2 * p + 2;
  // one contiguous reduction buffer: mv[3 np] | logdet_part[N] | gpart[nacc] | kal[np] | tsig[1]
  R.red_len = 3 * np + N + nacc + np + 1;
  DA(R.red, sizeof(double) * R.red_len);
  R.mv = R.red; R.logdet_part = R.mv + 3 * np; R.kal = nullptr;
  DA(R.alpha, vb); DA(R.u, vb); DA(R.s, vb);
  DA(R.gpart, sizeof(double) * nacc * R.ncta_grad);
  DA(R.rs_part, sizeof(double) * 2); DA(R.sc, sizeof(csk::EvalScalars)); DA(R.status, sizeof(int));
#undef DA
  // host -> device replicated data (rows >= n zero / one padded)
  std::vector<double> hx((size_t)np * p, 0.0), hv(np, 0.0);
  for (int j = 0; j < p; ++j) for (int i = 0; i < n; ++i) hx[(size_t)j * np + i] = X[(size_t)j * n + i];
  if ((e = rt::h2d(R.X, hx.data(), vb * p, st))) return e;
  std::copy(y, y + n, hv.begin()); if ((e = rt::h2d(R.y, hv.data(), vb, st))) return e;
  if ((e = rt::stream_sync(st))) return e;
  std::fill(hv.begin(), hv.end(), 0.0); std::copy(z, z + n, hv.begin()); if ((e = rt::h2d(R.z, hv.data(), vb, st))) return e;
  if ((e = rt::stream_sync(st))) return e;
  std::fill(hv.begin(), hv.end(), 1.0); if ((e = rt::h2d(R.ones, hv.data(), vb, st))) return e;
  if ((e = rt::stream_sync(st))) return e;
  if (w) std::copy(w, w + n, hv.begin());
  if ((e = rt::h2d(R.w, hv.data(), vb, st))) return e;
  if ((e = rt::h2d(R.params, params, sizeof(double) * NP, st))) return e;
  if ((e = rt::stream_sync(st))) return e;

  // owned lower tiles / diagonal tiles / 64x64 sub-tiles for the trace kernel
  std::vector<int> k3;
  for (int J = 0; J < N; ++J)
    for (int I = J; I < N; ++I)
      if (g.owns(I, J)) {
        R.lower_tiles.push_back(I); R.lower_tiles.push_back(J);
        if (I == J) R.diag_tiles.push_back(I);
        const int sub = T / csk::TB;
        for (int sc2 = 0; sc2 < sub; ++sc2)
          for (int sr = 0; sr < sub; ++sr) {
            const int I64 = I * sub + sr, J64 = J * sub + sc2;
            if (I64 < J64) continue;
            const long long co = g.off(I, J) + sr * csk::TB + (long long)sc2 * csk::TB * g.ldl;
            k3.push_back(I64); k3.push_back(J64); k3.push_back((int)(co & 0xffffffffLL)); k3.push_back((int)(co >> 32));
          }
      }
  R.n_k3 = (int)k3.size() / 4;
  auto up_int = [&](int** d, const std::vector<int>& h) -> int {
    if ((e = R.dalloc((void**)d, sizeof(int) * std::max<size_t>(h.size(), 1)))) return e;
    return rt::h2d(*d, h.data(), sizeof(int) * h.size(), st);
  };
  if ((e = up_int(&R.d_lower, R.lower_tiles))) return e;
  if ((e = up_int(&R.d_diag, R.diag_tiles))) return e;
  if ((e = up_int(&R.d_k3, k3))) return e;

  // per-step tile tables (element offsets into A, B, C of the table-mode GEMM), 64 x 64 CTA tiles
  const int gt = R.gt, sub = T / gt;
  auto sub_tiles = [&](StepTab& tab, long long a_tile, long long lda_, bool a_kmaj, long long b_tile, long long ldb_,
                       bool b_kmaj, long long c_off, bool diag_lower_only) {
    // one T x T tile update = sub x sub CTA tiles; operand tiles are T x T with leading dimension lda_/ldb_
    for (int sj = 0; sj < sub; ++sj)
      for (int si = 0; si < sub; ++si) {
        if (diag_lower_only && si < sj) continue;
        const long long ao = a_tile + (a_kmaj ? (long long)si * gt * lda_ : (long long)si * gt);
        const long long bo = b_tile + (b_kmaj ? (long long)sj * gt * ldb_ : (long long)sj * gt);
        tab.off.push_back(ao); tab.off.push_back(bo);
        tab.off.push_back(c_off + (long long)si * gt + (long long)sj * gt * g.ldl);
      }
  };
  for (int k = 0; k < N; ++k) {
    // ---- potrf update: local lower tiles with J > k; operands from the gathered column panel
    R.t_potrf.start.push_back((int)R.t_potrf.off.size() / 3);
    {
      const int li0 = k / P, cnt = g.mt - li0;
      auto slot = [&](int I) { return ((long long)(I % P) * cnt + (I / P - li0)) * (long long)T2; };
      // lower_tiles is ordered by J: the tiles of block column k+1 (what step k+1 needs) come first
      const size_t before = R.t_potrf.off.size();
      size_t after_next = before;
      for (size_t q = 0; q < R.lower_tiles.size(); q += 2) {
        const int I = R.lower_tiles[q], J = R.lower_tiles[q + 1];
        if (J <= k) continue;
        sub_tiles(R.t_potrf, slot(I), T, false, slot(J), T, false, g.off(I, J), I == J);
        if (J == k + 1) after_next = R.t_potrf.off.size();
      }
      R.na_potrf.push_back((int)((after_next - before) / 3));
    }
    // ---- finalize row k: X_kJ = -X_kk * Acc_kJ for owned (k,J), J < k  (one CTA per 128-tile, in place)
    R.t_fin.start.push_back((int)R.t_fin.off.size() / 3);
    if (k % P == pr)
      for (int J = pc; J < k; J += Q) { R.t_fin.off.push_back(0); R.t_fin.off.push_back(g.off(k, J)); R.t_fin.off.push_back(g.off(k, J)); }
    // ---- Acc_IJ += L_Ik X_kJ  (I > k, J <= k), C_IJ += X_kI^T X_kJ  (J <= I <= k)
    R.t_acc.start.push_back((int)R.t_acc.off.size() / 3);
    R.t_c.start.push_back((int)R.t_c.off.size() / 3);
    {
      const int li0 = k / P, cnt = g.mt - li0;
      auto cslot = [&](int I) { return ((long long)(I % P) * cnt + (I / P - li0)) * (long long)T2; };
      const int cntr = k / Q + 1;
      auto rslot = [&](int J) { return ((long long)(J % Q) * cntr + J / Q) * (long long)T2; };
      // Acc: the tiles of block row k+1 (finalised by step k+1) first, then the rows below
      const size_t before = R.t_acc.off.size();
      if ((k + 1) < N && (k + 1) % P == pr)
        for (int J = pc; J <= k; J += Q) sub_tiles(R.t_acc, cslot(k + 1), T, false, rslot(J), T, true, g.off(k + 1, J), false);
      R.na_acc.push_back((int)((R.t_acc.off.size() - before) / 3));
      for (int J = pc; J <= k; J += Q) {
        for (int I = pr; I < N; I += P) {
          if (I > k + 1) sub_tiles(R.t_acc, cslot(I), T, false, rslot(J), T, true, g.off(I, J), false);
          else if (I <= k && I >= J) sub_tiles(R.t_c, rslot(I), T, true, rslot(J), T, true, g.off(I, J), false);  // diagonal tiles in full: symv reads them as symmetric
        }
      }
    }
  }
  R.t_potrf.start.push_back((int)R.t_potrf.off.size() / 3);
  R.t_fin.start.push_back((int)R.t_fin.off.size() / 3);
  R.t_acc.start.push_back((int)R.t_acc.off.size() / 3);
  R.t_c.start.push_back((int)R.t_c.off.size() / 3);
  if ((e = upload_tab(R, R.t_potrf)) || (e = upload_tab(R, R.t_fin)) || (e = upload_tab(R, R.t_acc)) || (e = upload_tab(R, R.t_c))) return e;

  // per-step problems: [0] potrf update of block column k+1, [1] finalize row, [2] Acc of block row k+1, [3] C,
  // [4] rest of the potrf update, [5] rest of Acc.  Panels are read from the buffer of the step's parity.
  R.probs.assign(NPK * (size_t)N, GemmProblem{});
  std::vector<GemmProblem> panel(N, GemmProblem{});
  for (int k = 0; k < N; ++k) {
    auto mkp = [&](const double* A, long long lda, const double* B, long long ldb, double* Cc, StepTab& tab, int first, int count,
                   double alpha, double beta) {
      GemmProblem p_{};
      p_.A = A; p_.B = B; p_.C = Cc; p_.lda = lda; p_.ldb = ldb; p_.ldc = g.ldl; p_.K = T; p_.set_scale(alpha, beta);
      p_.ntiles = count; p_.tile_base = 0; p_.table = tab.d + 3LL * (tab.start[k] + first);
      p_.mt = p_.ntiles; p_.nt = 1;
      return p_;
    };
    double* COLk = R.COLb[k & 1];
    double* ROWk = R.ROWb[k & 1];
    const int n_potrf = R.t_potrf.start[k + 1] - R.t_potrf.start[k], n_acc = R.t_acc.start[k + 1] - R.t_acc.start[k];
    R.probs[NPK * k + 0] = mkp(COLk, T, COLk, T, R.G, R.t_potrf, 0, R.na_potrf[k], -1.0, 1.0);
    R.probs[NPK * k + 4] = mkp(COLk, T, COLk, T, R.G, R.t_potrf, R.na_potrf[k], n_potrf - R.na_potrf[k], -1.0, 1.0);
    R.probs[NPK * k + 1] = mkp(R.xkk, T, R.Xl, g.ldl, R.Xl, R.t_fin, 0, R.t_fin.start[k + 1] - R.t_fin.start[k], -1.0, 0.0);
    R.probs[NPK * k + 2] = mkp(COLk, T, ROWk, T, R.Xl, R.t_acc, 0, R.na_acc[k], 1.0, 1.0);
    R.probs[NPK * k + 5] = mkp(COLk, T, ROWk, T, R.Xl, R.t_acc, R.na_acc[k], n_acc - R.na_acc[k], 1.0, 1.0);
    R.probs[NPK * k + 3] = mkp(ROWk, T, ROWk, T, R.C, R.t_c, 0, R.t_c.start[k + 1] - R.t_c.start[k], 1.0, 1.0);
    // panel multiply L_Ik = A_Ik X_kk^T for the local rows below k (process column k mod Q only)
    if (k % Q == pc) {
      int first = -1, cntp = 0;
      for (int I = pr; I < N; I += P) if (I > k) { if (first < 0) first = I; ++cntp; }
      if (cntp > 0) {
        double* pan = R.G + g.off(first, k);
        panel[k] = csdense::DenseEngine::mk(pan, g.ldl, R.xkk, T, pan, g.ldl, cntp, 1, T, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0);
      }
    }
  }
  if ((e = R.dalloc((void**)&R.d_probs, sizeof(GemmProblem) * R.probs.size()))) return e;
  if ((e = rt::h2d(R.d_probs, R.probs.data(), sizeof(GemmProblem) * R.probs.size(), st))) return e;
  if ((e = R.dalloc((void**)&R.d_panel, sizeof(GemmProblem) * N))) return e;
  if ((e = rt::h2d(R.d_panel, panel.data(), sizeof(GemmProblem) * N, st))) return e;
  R.probs.insert(R.probs.end(), panel.begin(), panel.end());  // keep host copies alive: [NPK*N .. (NPK+1)*N) = panel
  if ((e = rt::stream_sync(st))) return e;
  csdense::configure_kernels();
  { auto kf = gram_dist_kernel; CS_SET_SMEM(kf, csk::gram_smem_bytes(csk::PMAX)); }
  { auto kf = csk::grad_trace_kernel<false>; CS_SET_SMEM(kf, csk::SMEM_LIMIT); }
  return 0;
}

// ---------------------------------------------------------------------------------
// the evaluation driver (shared by the NCCL and the in-process back-ends)
// ---------------------------------------------------------------------------------
struct DistEval {
  std::vector<DistRank*> L;   // ranks living in this process
  Comm* comm = nullptr;
  int T = 128, N = 0, P = 1, Q = 1, n = 0;

  template <class F> void each(F f) { for (DistRank* r : L) f(*r); }

  int run() {
    const size_t T2 = (size_t)T * T;
    int e;
    // ---- K1: local Gram tiles; zero the accumulators ----
    each([&](DistRank& R) {
      const Geom& g = R.g;
      const size_t lm = sizeof(double) * (size_t)g.ldl * g.nt * T;
      rt::memset0(R.Xl, lm, R.st); rt::memset0(R.C, lm, R.st);
      rt::memset0(R.red, sizeof(double) * R.red_len, R.st);
      auto k0 = csdense::set_int_kernel; ++g_launches;
      CS_LAUNCH(k0, 1, 32, 0, R.st, R.status, 0x7fffffff, (const int*)nullptr, 0LL);
      const int ntl = (int)R.lower_tiles.size() / 2, sub = T / csk::TB;
      if (ntl > 0) {
        auto k = gram_dist_kernel; ++g_launches;
        CS_LAUNCH(k, ntl * sub * sub, csk::ETH, csk::gram_smem_bytes(R.lay.p), R.st, R.X, (long long)g.np, R.z, R.w, R.params,
                  R.lay, g.n, (const int*)R.d_lower, T, P, Q, R.G, g.ldl);
      }
    });
    static const bool two_sweeps = std::getenv("CS_DIST_TWO_SWEEPS") != nullptr;
    if (!two_sweeps) {
      if ((e = fused_sweep())) return e;
    } else {
    // ---- sweep 1: Cholesky, with look-ahead.  The update of step k is split: block column k+1 on the main
    // stream (the next leaf / panel multiply wait for nothing else), the rest on the side stream st2 where it
    // overlaps the leaf, the panel multiply and the communication of step k+1.
    bool side_pending = false;
    auto join_side = [&]() {  // main stream waits for the side stream's last update
      if (!side_pending) return;
      each([&](DistRank& R) { rt::stream_wait_event(R.st, R.ev_side); });
    };
    for (int k = 0; k < N; ++k) {
      const int prk = k % P, pck = k % Q;
      each([&](DistRank& R) {
        if (R.g.pr == prk && R.g.pc == pck) {
          const long long o = R.g.off(k, k);
          csdense::launch_leaf(T, R.G + o, R.g.ldl, R.Xl + o, R.g.ldl, k * T, R.logdet_part + k, R.status, n, R.st, nullptr);
          pack(R, R.Xl + o, R.g.ldl, 0, R.xkk, 1);
        }
      });
      if (k == N - 1) break;
      if ((e = comm->bcast(L, G_COL, pck, prk, [](DistRank& R) { return R.xkk; }, T2))) return e;
      const int li0 = k / P;
      each([&](DistRank& R) {
        if (R.g.pc != pck) return;
        const GemmProblem& pp = R.probs[NPK * (size_t)N + k];
        if (pp.ntiles > 0) csdense::launch_gemm(T, csdense::GK_NT, R.d_panel + k, 1, pp.ntiles, R.st);
        // publish the local part of column panel k (tiles li >= li0)
        pack(R, R.G + (long long)li0 * T + (long long)R.g.lj(k) * T * R.g.ldl, R.g.ldl, T, R.colsend, R.g.mt - li0);
      });
      if ((e = share_col_panel(k))) return e;
      join_side();  // the rest of step k-1 also updated block column k+1 (and has finished reading the other panel buffer)
      each([&](DistRank& R) {
        rt::event_record(R.ev_share, R.st);
        const GemmProblem& pa = R.probs[NPK * (size_t)k + 0];
        if (pa.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NT, R.d_probs + NPK * k + 0, 1, pa.ntiles, R.st);
        rt::stream_wait_event(R.st2, R.ev_share);
        const GemmProblem& pb = R.probs[NPK * (size_t)k + 4];
        if (pb.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NT, R.d_probs + NPK * k + 4, 1, pb.ntiles, R.st2);
        rt::event_record(R.ev_side, R.st2);
      });
      side_pending = true;
    }
    join_side();
    // ---- sweep 2: triangular inverse + K^-1 = X^T X, fused; same look-ahead (block row k+1 of Acc first) ----
    for (int k = 0; k < N; ++k) {
      const int prk = k % P, pck = k % Q;
      each([&](DistRank& R) {
        if (R.g.pr == prk && R.g.pc == pck) pack(R, R.Xl + R.g.off(k, k), R.g.ldl, 0, R.xkk, 1);
      });
      if ((e = comm->bcast(L, G_ROW, prk, pck, [](DistRank& R) { return R.xkk; }, T2))) return e;
      const int cntr = k / Q + 1;
      const int par = k & 1;
      each([&](DistRank& R) {
        if (R.g.pr != prk) return;
        const GemmProblem& pf = R.probs[NPK * (size_t)k + 1];
        if (pf.ntiles > 0) csdense::launch_gemm(T, csdense::GK_NN, R.d_probs + NPK * k + 1, 1, pf.ntiles, R.st);
        // publish the local part of row panel k of X (local column tiles 0 .. cntr-1)
        pack(R, R.Xl + (long long)R.g.li(k) * T, R.g.ldl, (long long)T * R.g.ldl, R.rowsend, cntr);
      });
      // the side stream of step k-2 has finished with this parity's buffers: the main stream joined it at step k-1
      if ((e = comm->share_panel(L, G_ROW, prk, [](DistRank& R) { return R.rowsend; }, [par](DistRank& R) { return R.ROWb[par]; }, cntr * T2, Q))) return e;
      if (k < N - 1) {
        const int li0 = k / P;
        each([&](DistRank& R) {
          if (R.g.pc != pck) return;
          pack(R, R.G + (long long)li0 * T + (long long)R.g.lj(k) * T * R.g.ldl, R.g.ldl, T, R.colsend, R.g.mt - li0);
        });
        if ((e = share_col_panel(k))) return e;
      }
      join_side();
      each([&](DistRank& R) {
        rt::event_record(R.ev_share, R.st);
        const GemmProblem& pa = R.probs[NPK * (size_t)k + 2];
        if (pa.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NN, R.d_probs + NPK * k + 2, 1, pa.ntiles, R.st);
        rt::stream_wait_event(R.st2, R.ev_share);
        const GemmProblem& pb = R.probs[NPK * (size_t)k + 5];
        if (pb.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NN, R.d_probs + NPK * k + 5, 1, pb.ntiles, R.st2);
        const GemmProblem& pcx = R.probs[NPK * (size_t)k + 3];
        if (pcx.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_TN, R.d_probs + NPK * k + 3, 1, pcx.ntiles, R.st2);
        rt::event_record(R.ev_side, R.st2);
      });
      side_pending = true;
    }
    join_side();
    }  // two sweeps
    // ---- reductions: alpha/u/s, scalars, trace gradients, stats ----
    each([&](DistRank& R) {
      const Geom& g = R.g;
      const int ntl = (int)R.lower_tiles.size() / 2;
      if (ntl > 0) {
        auto k = symv_dist_kernel; ++g_launches;
        CS_LAUNCH(k, ntl, 256, 0, R.st, R.C, g.ldl, (const int*)R.d_lower, T, P, Q, R.y, R.params + R.lay.i_mu(), R.y, R.ones,
                  R.mv, (long long)g.np);
      }
    });
    if ((e = comm->allreduce_sum(L, [](DistRank& R) { return R.red; }, 0))) return e;  // count 0 = whole buffer, phase A
    each([&](DistRank& R) {
      const Geom& g = R.g;
      const long long ld = g.np;
      auto k2 = csk::scalars_kernel; ++g_launches;
      CS_LAUNCH(k2, 1, csk::ETH, 0, R.st, R.mv, 1, ld, 3 * ld, g.n, g.np, R.y, R.params, R.lay, R.logdet_part, g.N, R.alpha, R.u,
                R.s, R.sc, (const int*)nullptr);
      // second reduction buffer reuses `red`: gpart[nacc] | kal[np] | tsig[1]
      rt::memset0(R.red, sizeof(double) * R.red_len, R.st);
      const int nacc = 2 * R.lay.p + 2;
      double* kal = R.red + nacc; double* tsig = kal + g.np;
      auto k3 = csk::grad_trace_kernel<false>; ++g_launches;
      CS_LAUNCH(k3, R.ncta_grad, csk::ETH, csk::grad_smem_bytes(R.lay.p), R.st, R.X, ld, R.z, R.params, R.lay, g.n, R.n_k3,
                R.C, g.ldl, R.alpha, R.sc, R.gpart, kal, (long long)0, (const double*)nullptr, (const double*)nullptr,
                (const double*)nullptr, (const int*)nullptr, (const int*)R.d_k3, rt::XTma{}, 0, R.lay.p);
      auto k4 = csk::reduce_gpart_kernel; ++g_launches;
      CS_LAUNCH(k4, 1, 128, 0, R.st, R.gpart, R.ncta_grad, nacc, R.red);
      if (!R.diag_tiles.empty()) {
        auto k5 = diag_trace_dist_kernel; ++g_launches;
        CS_LAUNCH(k5, 1, 256, 0, R.st, R.C, g.ldl, (const int*)R.d_diag, (int)R.diag_tiles.size(), T, P, Q, R.alpha, R.w, R.params,
                  R.lay, g.n, R.sc, tsig);
      }
    });
    if ((e = comm->allreduce_sum(L, [](DistRank& R) { return R.red; }, 1))) return e;  // phase B
    each([&](DistRank& R) {
      const Geom& g = R.g;
      const int nacc = 2 * R.lay.p + 2;
      double* kal = R.red + nacc; double* tsig = kal + g.np;
      auto k6 = resid_dist_kernel; ++g_launches;
      CS_LAUNCH(k6, 1, 256, 0, R.st, kal, R.y, R.params, R.lay, g.n, tsig, R.rs_part);
      auto k7 = csk::finalize_kernel; ++g_launches;
      CS_LAUNCH(k7, 1, csk::ETH, 0, R.st, R.red, 1, R.rs_part, 1, R.params, R.lay, g.n, R.sc, R.grad, (const int*)nullptr);
    });
    return 0;
  }

  // Round 2: ONE sweep.  Row k of X = L^-1 can be finalised as soon as block column k of L and X_kk exist (its Acc_kJ only
  // collects contributions of the steps before k), so the inverse / K^-1 step k is issued right behind the Cholesky step k
  // instead of in a second pass over the matrix: the column panel of L is shared once instead of twice, the chain of
  // leaf + collectives is walked N times instead of 2N, and the bulk work of the inverse (2/3 of all flops: Acc and C updates,
  // side stream) overlaps the latency-bound Cholesky chain of the following steps -- the same idea as the pipelined inverse of
  // the single-GPU engine.  Tables, problems and buffers are those of the two-sweep driver (CS_DIST_TWO_SWEEPS=1 keeps it).
  int fused_sweep() {
    const size_t T2 = (size_t)T * T;
    int e;
    bool side_pending = false;
    auto join_side = [&]() {
      if (!side_pending) return;
      each([&](DistRank& R) { rt::stream_wait_event(R.st, R.ev_side); });
    };
    for (int k = 0; k < N; ++k) {
      const int prk = k % P, pck = k % Q;
      const int par = k & 1;
      const bool last = (k == N - 1);
      each([&](DistRank& R) {
        if (R.g.pr == prk && R.g.pc == pck) {
          const long long o = R.g.off(k, k);
          csdense::launch_leaf(T, R.G + o, R.g.ldl, R.Xl + o, R.g.ldl, k * T, R.logdet_part + k, R.status, n, R.st, nullptr);
          pack(R, R.Xl + o, R.g.ldl, 0, R.xkk, 1);
        }
      });
      // X_kk to the process column (panel multiply) and to the process row (row k of X)
      if (!last && (e = comm->bcast(L, G_COL, pck, prk, [](DistRank& R) { return R.xkk; }, T2))) return e;
      if ((e = comm->bcast(L, G_ROW, prk, pck, [](DistRank& R) { return R.xkk; }, T2))) return e;
      const int li0 = k / P;
      const int cntr = k / Q + 1;
      if (!last)
        each([&](DistRank& R) {
          if (R.g.pc != pck) return;
          const GemmProblem& pp = R.probs[NPK * (size_t)N + k];
          if (pp.ntiles > 0) csdense::launch_gemm(T, csdense::GK_NT, R.d_panel + k, 1, pp.ntiles, R.st);
          pack(R, R.G + (long long)li0 * T + (long long)R.g.lj(k) * T * R.g.ldl, R.g.ldl, T, R.colsend, R.g.mt - li0);
        });
      each([&](DistRank& R) {
        if (R.g.pr != prk) return;
        const GemmProblem& pf = R.probs[NPK * (size_t)k + 1];
        if (pf.ntiles > 0) csdense::launch_gemm(T, csdense::GK_NN, R.d_probs + NPK * k + 1, 1, pf.ntiles, R.st);
        pack(R, R.Xl + (long long)R.g.li(k) * T, R.g.ldl, (long long)T * R.g.ldl, R.rowsend, cntr);
      });
      if (!last && (e = share_col_panel(k))) return e;
      if ((e = comm->share_panel(L, G_ROW, prk, [](DistRank& R) { return R.rowsend; }, [par](DistRank& R) { return R.ROWb[par]; }, cntr * T2, Q))) return e;
      join_side();  // the side stream has finished step k-1 (it updated block column / row k+1 and read the other buffers)
      each([&](DistRank& R) {
        rt::event_record(R.ev_share, R.st);
        if (!last) {
          const GemmProblem& pa = R.probs[NPK * (size_t)k + 0];
          if (pa.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NT, R.d_probs + NPK * k + 0, 1, pa.ntiles, R.st);
        }
        const GemmProblem& pa2 = R.probs[NPK * (size_t)k + 2];
        if (pa2.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NN, R.d_probs + NPK * k + 2, 1, pa2.ntiles, R.st);
        rt::stream_wait_event(R.st2, R.ev_share);
        if (!last) {
          const GemmProblem& pb = R.probs[NPK * (size_t)k + 4];
          if (pb.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NT, R.d_probs + NPK * k + 4, 1, pb.ntiles, R.st2);
        }
        const GemmProblem& pb2 = R.probs[NPK * (size_t)k + 5];
        if (pb2.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_NN, R.d_probs + NPK * k + 5, 1, pb2.ntiles, R.st2);
        const GemmProblem& pcx = R.probs[NPK * (size_t)k + 3];
        if (pcx.ntiles > 0) csdense::launch_gemm(R.gt, csdense::GK_TN, R.d_probs + NPK * k + 3, 1, pcx.ntiles, R.st2);
        rt::event_record(R.ev_side, R.st2);
      });
      side_pending = true;
    }
    join_side();
    return 0;
  }

  void pack(DistRank& R, const double* src, long long lds, long long stride, double* dst, int ntiles) {
    if (ntiles <= 0) return;
    auto k = pack_tiles_kernel; ++g_launches;
    const long long tot = (long long)ntiles * T * T;
    CS_LAUNCH(k, (unsigned)std::min<long long>((tot + 255) / 256, 4096), 256, 0, R.st, src, lds, stride, dst, T, ntiles);
  }

  // column panel k: all-gather inside process column k mod Q, then broadcast along every process row
  int share_col_panel(int k) {
    const size_t T2 = (size_t)T * T;
    const int pck = k % Q, li0 = k / P;
    int e;
    const size_t cnt = (size_t)(L[0]->g.mt - li0) * T2;
    const int par = k & 1;  // the updates of step k read COLb[k & 1]
    if ((e = comm->share_panel(L, G_COL, pck, [](DistRank& R) { return R.colsend; }, [par](DistRank& R) { return R.COLb[par]; }, cnt, P))) return e;
    return 0;
  }
};

// ---------------------------------------------------------------------------------
// in-process back-end: every rank of the grid lives in this process (one device or the
// emulator); collectives are stream-ordered device copies.
// ---------------------------------------------------------------------------------
struct SimComm : Comm {
  int P, Q;
  std::vector<double> tmp;
  SimComm(int P_, int Q_) : P(P_), Q(Q_) {}
  DistRank* at(std::vector<DistRank*>& L, int pr, int pc) { return L[(size_t)pr * Q + pc]; }
  void members(std::vector<DistRank*>& L, int group, int idx, std::vector<DistRank*>& out) {
    out.clear();
    if (group == G_ROW) for (int c = 0; c < Q; ++c) out.push_back(at(L, idx, c));
    else if (group == G_COL) for (int r = 0; r < P; ++r) out.push_back(at(L, r, idx));
    else for (DistRank* r : L) out.push_back(r);
  }
  int bcast(std::vector<DistRank*>& L, int group, int which, int root, const BufFn& buf, size_t count) override {
    const int ng = group == G_ROW ? P : (group == G_COL ? Q : 1);
    std::vector<DistRank*> m;
    for (int gi = 0; gi < ng; ++gi) {
      if (which >= 0 && gi != which) continue;
      members(L, group, gi, m);
      for (size_t i = 0; i < m.size(); ++i)
        if ((int)i != root) { if (int e = rt::d2d(buf(*m[i]), buf(*m[root]), sizeof(double) * count, m[i]->st)) return e; }
    }
    return 0;
  }
  int allgather(std::vector<DistRank*>& L, int group, int which, const BufFn& send, const BufFn& recv, size_t count) override {
    const int ng = group == G_ROW ? P : (group == G_COL ? Q : 1);
    std::vector<DistRank*> m;
    for (int gi = 0; gi < ng; ++gi) {
      if (which >= 0 && gi != which) continue;
      members(L, group, gi, m);
      for (size_t i = 0; i < m.size(); ++i)
        for (size_t j = 0; j < m.size(); ++j)
          if (int e = rt::d2d(recv(*m[i]) + j * count, send(*m[j]), sizeof(double) * count, m[i]->st)) return e;
    }
    return 0;
  }
  int allreduce_sum(std::vector<DistRank*>& L, const BufFn& buf, size_t) override {
    // sum into a host staging vector (test back-end only), then write back to every rank
    const size_t len = (size_t)L[0]->red_len;
    std::vector<double> acc(len, 0.0), one(len);
    for (DistRank* r : L) {
      if (int e = rt::d2h(one.data(), buf(*r), sizeof(double) * len, r->st)) return e;
      if (int e = rt::stream_sync(r->st)) return e;
      for (size_t i = 0; i < len; ++i) acc[i] += one[i];
    }
    for (DistRank* r : L) {
      if (int e = rt::h2d(buf(*r), acc.data(), sizeof(double) * len, r->st)) return e;
      if (int e = rt::stream_sync(r->st)) return e;
    }
    return 0;
  }
};

}  // namespace csdist
