// Dense engine: K_n -> L (blocked right-looking Cholesky) -> X = L^-1 (recursive
// triangular inverse, level-batched) -> K_n^-1 = X^T X (one triangular-skipping launch).
// Replaces arma::inv_sympd (LAPACK dpotrf + dpotri) of /root/reference
// src/kernel_GP_SE_cpp.cpp:56 and the log_det of :197 (log-determinant comes from diag(L)).
//
// Storage (all column-major, leading dimension np = n rounded up to TILE):
//   G : K_n lower triangle in, L out (upper triangle explicit zeros)
//   X : L^-1 (upper triangle explicit zeros)
//   C : workspace for the inverse recursion, then K_n^-1 (full symmetric)
// The padded tail (rows/cols n..np-1) is an identity block, so every tile is full and
// log det is unchanged.
#pragma once
#include <vector>

#include "gemm.cuh"

namespace csdense {

using csg::GemmProblem;

// ---------------------------------------------------------------------------------
// Leaf: Cholesky factor and inverse of one TILE x TILE diagonal block in shared memory.
// ---------------------------------------------------------------------------------
constexpr int LEAF_THREADS = 256;

template <int TILE>
constexpr int leaf_smem_bytes() { return (TILE * (TILE + 1) + 2 * TILE + 32) * (int)sizeof(double); }

template <int TILE>
__global__ void __launch_bounds__(LEAF_THREADS)
leaf_kernel(double* __restrict__ G, long long ldg, double* __restrict__ X, long long ldx, int blk,
            double* __restrict__ logdet_part, int* __restrict__ status, int n_real,
            const int* __restrict__ skip) {
  if (skip && *skip) return;
  constexpr int LD = TILE + 1;
  CS_DYN_SMEM(double, S);
  double* colbuf = S + TILE * LD;
  double* rowbuf = colbuf + TILE;  // also holds the pivots during phase 1
  double* red = rowbuf + TILE;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double* Gt = G + (long long)blk * TILE * (ldg + 1);
  double* Xt = X + (long long)blk * TILE * (ldx + 1);

  for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
    const int r = idx % TILE, c = idx / TILE;
    S[r + c * LD] = (r >= c) ? Gt[r + c * ldg] : 0.0;
  }
  __syncthreads();

  // phase 1: right-looking Cholesky, columns kept unscaled (L[r][j] = S[r][j]/sqrt(d_j))
  for (int j = 0; j < TILE; ++j) {
    double d = S[j + j * LD];
    if (!(d > 0.0)) {
      const int gidx = blk * TILE + j;
      if (tid == 0 && gidx < n_real) atomicMin(status, gidx + 1);  // LAPACK-style 1-based info
      d = 1.0;
    }
    if (tid == 0) rowbuf[j] = d;
    const double invd = 1.0 / d;
    for (int c = j + 1 + ((ty - (j + 1)) & 15); c < TILE; c += 16) {
      const double scj = S[c + j * LD] * invd;
      for (int r = c + ((tx - c) & 15); r < TILE; r += 16) S[r + c * LD] -= S[r + j * LD] * scj;
    }
    __syncthreads();
  }
  // scale columns, emit L, accumulate 0.5*log(d) = log(L_jj)
  double ld_local = 0.0;
  for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
    const int r = idx % TILE, c = idx / TILE;
    double v = 0.0;
    if (r >= c) {
      const double sq = sqrt(rowbuf[c]);
      v = (r == c) ? sq : S[r + c * LD] / sq;
      if (r == c) ld_local += log(sq);
    }
    S[r + c * LD] = v;
    Gt[r + c * ldg] = v;
  }
  const double ldsum = csd::block_sum(ld_local, red);  // contains __syncthreads
  if (tid == 0) logdet_part[blk] = ldsum;

  // phase 2: in-place inverse of the lower-triangular factor (row-by-row, rank-1 updates)
  for (int i = 0; i < TILE; ++i) {
    const double xii = 1.0 / S[i + i * LD];
    if (tid < TILE) {
      if (tid > i) colbuf[tid] = S[tid + i * LD];
      if (tid < i) rowbuf[tid] = -xii * S[i + tid * LD];
      if (tid == i) rowbuf[tid] = xii;
    }
    __syncthreads();
    if (tid <= i) S[i + tid * LD] = rowbuf[tid];
    for (int jj = ty; jj <= i; jj += 16) {
      const double xv = rowbuf[jj];
      for (int r = i + 1 + ((tx - (i + 1)) & 15); r < TILE; r += 16) {
        const double prev = (jj == i) ? 0.0 : S[r + jj * LD];
        S[r + jj * LD] = prev + colbuf[r] * xv;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
    const int r = idx % TILE, c = idx / TILE;
    Xt[r + c * ldx] = (r >= c) ? S[r + c * LD] : 0.0;
  }
}

// ---------------------------------------------------------------------------------
// Host-side plan
// ---------------------------------------------------------------------------------
enum GemmKind { GK_NT = 0, GK_NN = 1, GK_TN = 2 };

static long long g_launches = 0;  // kernels launched by this library (claimed count for bench.py)

struct Launch {
  int is_leaf;     // 1: leaf_kernel(blk)   0: gemm group
  int blk;         // leaf block index
  int kind;        // GemmKind
  int first_prob;  // index into the problem array
  int nprob;
  int ntiles;
};

inline int pick_tile(int n) { return n <= 1536 ? 64 : 128; }

template <int TILE>
inline void launch_gemm_t(int kind, const GemmProblem* dprobs, int nprob, int ntiles, cudaStream_t st,
                          const int* skip) {
  if (ntiles <= 0) return;
  if (TILE == 128) {
    if (kind == GK_NT) { auto k = csg::gemm_kernel<128, 2, 4, false, false>; CS_LAUNCH(k, ntiles, 256, (csg::gemm_smem_bytes<128, false, false>()), st, dprobs, nprob, skip); }
    else if (kind == GK_NN) { auto k = csg::gemm_kernel<128, 2, 4, false, true>; CS_LAUNCH(k, ntiles, 256, (csg::gemm_smem_bytes<128, false, true>()), st, dprobs, nprob, skip); }
    else { auto k = csg::gemm_kernel<128, 2, 4, true, true>; CS_LAUNCH(k, ntiles, 256, (csg::gemm_smem_bytes<128, true, true>()), st, dprobs, nprob, skip); }
  } else {
    if (kind == GK_NT) { auto k = csg::gemm_kernel<64, 2, 2, false, false>; CS_LAUNCH(k, ntiles, 128, (csg::gemm_smem_bytes<64, false, false>()), st, dprobs, nprob, skip); }
    else if (kind == GK_NN) { auto k = csg::gemm_kernel<64, 2, 2, false, true>; CS_LAUNCH(k, ntiles, 128, (csg::gemm_smem_bytes<64, false, true>()), st, dprobs, nprob, skip); }
    else { auto k = csg::gemm_kernel<64, 2, 2, true, true>; CS_LAUNCH(k, ntiles, 128, (csg::gemm_smem_bytes<64, true, true>()), st, dprobs, nprob, skip); }
  }
}

inline void launch_gemm(int tile, int kind, const GemmProblem* dprobs, int nprob, int ntiles, cudaStream_t st,
                        const int* skip = nullptr) {
  ++g_launches;
  if (tile == 128) launch_gemm_t<128>(kind, dprobs, nprob, ntiles, st, skip);
  else launch_gemm_t<64>(kind, dprobs, nprob, ntiles, st, skip);
}

// opt-in to > 48 KB dynamic shared memory (per device; cheap, called at every engine init)
inline void configure_kernels() {
  { auto k = csg::gemm_kernel<128, 2, 4, false, false>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<128, false, false>())); }
  { auto k = csg::gemm_kernel<128, 2, 4, false, true>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<128, false, true>())); }
  { auto k = csg::gemm_kernel<128, 2, 4, true, true>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<128, true, true>())); }
  { auto k = csg::gemm_kernel<64, 2, 2, false, false>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<64, false, false>())); }
  { auto k = csg::gemm_kernel<64, 2, 2, false, true>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<64, false, true>())); }
  { auto k = csg::gemm_kernel<64, 2, 2, true, true>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<64, true, true>())); }
  { auto k = leaf_kernel<128>; CS_SET_SMEM(k, leaf_smem_bytes<128>()); }
  { auto k = leaf_kernel<64>; CS_SET_SMEM(k, leaf_smem_bytes<64>()); }
}

struct DenseEngine {
  int n = 0, np = 0, tile = 0, N = 0;
  double *G = nullptr, *X = nullptr, *C = nullptr;  // device, np x np each
  double* logdet_part = nullptr;                     // device, N
  int* status = nullptr;                             // device, 1 (0 = OK, else 1-based failing pivot)
  GemmProblem* d_probs = nullptr;
  std::vector<GemmProblem> probs;
  std::vector<Launch> potrf, trtri, xtx;

  static GemmProblem mk(const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
                        int mt, int nt, int K, int kmode, int order, int mirror, double alpha, double beta) {
    GemmProblem p{};
    p.A = A; p.B = B; p.C = C; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.mt = mt; p.nt = nt; p.K = K; p.kmode = kmode; p.order = order; p.mirror = mirror;
    p.alpha = alpha; p.beta = beta; p.tile_base = 0;
    p.ntiles = (order == csg::ORD_LOWER) ? csg::lower_tiles(mt) : mt * nt;
    return p;
  }

  void add_group(std::vector<Launch>& seq, int kind, std::vector<GemmProblem>& group) {
    if (group.empty()) return;
    Launch L{};
    L.is_leaf = 0; L.kind = kind; L.first_prob = (int)probs.size(); L.nprob = (int)group.size();
    int base = 0;
    for (auto& p : group) { p.tile_base = base; base += p.ntiles; probs.push_back(p); }
    L.ntiles = base;
    if (L.ntiles > 0) seq.push_back(L);
    group.clear();
  }

  struct Node { int lo, mid, hi, height; };
  int build_tree(int lo, int hi, std::vector<Node>& nodes) {
    if (hi - lo <= 1) return 0;
    const int mid = lo + (hi - lo) / 2;
    const int h = 1 + std::max(build_tree(lo, mid, nodes), build_tree(mid, hi, nodes));
    nodes.push_back(Node{lo, mid, hi, h});
    return h;
  }

  // returns 0 or a runtime error code
  int init(int n_, int tile_) {
    n = n_; tile = tile_; N = (n + tile - 1) / tile; np = N * tile;
    const size_t bytes = (size_t)np * np * sizeof(double);
    int e;
    if ((e = rt::dev_malloc((void**)&G, bytes))) return e;
    if ((e = rt::dev_malloc((void**)&X, bytes))) return e;
    if ((e = rt::dev_malloc((void**)&C, bytes))) return e;
    if ((e = rt::dev_malloc((void**)&logdet_part, sizeof(double) * N))) return e;
    if ((e = rt::dev_malloc((void**)&status, sizeof(int)))) return e;
    configure_kernels();
    const long long ld = np;
    const long long T = tile;
    std::vector<GemmProblem> grp;
    // --- blocked right-looking Cholesky ---
    for (int k = 0; k < N; ++k) {
      Launch L{}; L.is_leaf = 1; L.blk = k; potrf.push_back(L);
      const int m = N - k - 1;
      if (m <= 0) continue;
      double* panel = G + (k + 1) * T + k * T * ld;
      // L21 = A21 * X_kk^T   (NT, in place)
      grp.push_back(mk(panel, ld, X + k * T * (ld + 1), ld, panel, ld, m, 1, tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
      add_group(potrf, GK_NT, grp);
      // A22 -= L21 L21^T   (NT, lower tiles only)
      grp.push_back(mk(panel, ld, panel, ld, G + (k + 1) * T * (ld + 1), ld, m, m, tile, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
      add_group(potrf, GK_NT, grp);
    }
    // --- recursive triangular inverse, batched by tree height ---
    std::vector<Node> nodes;
    const int hmax = build_tree(0, N, nodes);
    for (int h = 1; h <= hmax; ++h) {
      std::vector<GemmProblem> g1, g2;
      for (const Node& nd : nodes) {
        if (nd.height != h) continue;
        const int m = nd.hi - nd.mid, q = nd.mid - nd.lo;
        double* Tw = C + nd.mid * T + nd.lo * T * ld;
        // T = L21 * X11      (NN; X11 lower => k >= j)
        g1.push_back(mk(G + nd.mid * T + nd.lo * T * ld, ld, X + nd.lo * T * (ld + 1), ld, Tw, ld, m, q, q * tile,
                        csg::KM_GE_J, csg::ORD_COL, 0, 1.0, 0.0));
        // X21 = -X22 * T     (NN; X22 lower => k <= i)
        g2.push_back(mk(X + nd.mid * T * (ld + 1), ld, Tw, ld, X + nd.mid * T + nd.lo * T * ld, ld, m, q, m * tile,
                        csg::KM_LE_I, csg::ORD_ROW_DESC, 0, -1.0, 0.0));
      }
      add_group(trtri, GK_NN, g1);
      add_group(trtri, GK_NN, g2);
    }
    // --- K^-1 = X^T X  (TN; k >= max(i,j) = i on lower tiles; mirrored store) ---
    grp.push_back(mk(X, ld, X, ld, C, ld, N, N, np, csg::KM_GE_I, csg::ORD_LOWER, 1, 1.0, 0.0));
    add_group(xtx, GK_TN, grp);

    if ((e = rt::dev_malloc((void**)&d_probs, sizeof(GemmProblem) * probs.size()))) return e;
    if ((e = rt::h2d(d_probs, probs.data(), sizeof(GemmProblem) * probs.size(), 0))) return e;
    if ((e = rt::stream_sync(0))) return e;
    return 0;
  }

  void destroy() {
    rt::dev_free(G); rt::dev_free(X); rt::dev_free(C); rt::dev_free(logdet_part); rt::dev_free(status);
    rt::dev_free(d_probs);
    G = X = C = logdet_part = nullptr; status = nullptr; d_probs = nullptr;
  }

  void run_seq(const std::vector<Launch>& seq, cudaStream_t st, const int* skip = nullptr) {
    for (const Launch& L : seq) {
      if (L.is_leaf) {
        ++g_launches;
        if (tile == 128) { auto k = leaf_kernel<128>; CS_LAUNCH(k, 1, LEAF_THREADS, leaf_smem_bytes<128>(), st, G, (long long)np, X, (long long)np, L.blk, logdet_part, status, n, skip); }
        else { auto k = leaf_kernel<64>; CS_LAUNCH(k, 1, LEAF_THREADS, leaf_smem_bytes<64>(), st, G, (long long)np, X, (long long)np, L.blk, logdet_part, status, n, skip); }
      } else {
        launch_gemm(tile, L.kind, d_probs + L.first_prob, L.nprob, L.ntiles, st, skip);
      }
    }
  }

  // G (lower) must hold K_n; X and the upper triangle of G must be zero outside diagonal tiles.
  void factor(cudaStream_t st, const int* skip = nullptr) { set_status_max(st, skip); run_seq(potrf, st, skip); }
  void invert(cudaStream_t st, const int* skip = nullptr) { run_seq(trtri, st, skip); run_seq(xtx, st, skip); }

  void set_status_max(cudaStream_t st, const int* skip);
};

static __global__ void set_int_kernel(int* p, int v, const int* skip) {
  if (skip && *skip) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) *p = v;
}
inline void DenseEngine::set_status_max(cudaStream_t st, const int* skip) {
  auto k = set_int_kernel;
  ++g_launches;
  CS_LAUNCH(k, 1, 32, 0, st, status, 0x7fffffff, skip);
}

}  // namespace csdense
