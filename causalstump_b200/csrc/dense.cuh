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
#include <atomic>
#include <vector>

#include "gemm.cuh"

namespace csdense {

using csg::GemmProblem;

// ---------------------------------------------------------------------------------
// Leaf: Cholesky factor and inverse of one TILE x TILE diagonal block in shared memory.
// ---------------------------------------------------------------------------------
constexpr int LEAF_THREADS = 256;

template <int TILE>
constexpr int leaf_smem_bytes() { return (6 * TILE + 32) * (int)sizeof(double); }

// Register-resident leaf: the TILE x TILE block lives in the register files of a 16 x 16
// thread grid, cyclically distributed (thread (tx,ty) owns rows tx+16a, columns ty+16b), so
// the work stays balanced as the factorization shrinks.  Per elimination step only one column
// (and, for the inverse, one row) crosses shared memory; one barrier per step (double-buffered).
template <int TILE>
__global__ void __launch_bounds__(LEAF_THREADS)
leaf_kernel(double* __restrict__ Gt, long long ldg, double* __restrict__ Xt, long long ldx, int pivot_base,
            double* __restrict__ logdet_out, int* __restrict__ status, int n_real,
            const int* __restrict__ skip, long long bstride, int mode) {
  // mode 0: factor and invert (L into Gt, L^-1 into Xt); 1: factor only; 2: invert only (L is read back from Gt) -- the
  // split lets the inversion leave the critical chain of the Cholesky (plan v2: chain = factor + triangular solve)
  // Gt / Xt point at the diagonal tile; pivot_base is its first global row (for LAPACK `info`)
  skip = csd::bshift(skip, bstride);
  if (skip && *skip) return;
  Gt = csd::bshift(Gt, bstride); Xt = csd::bshift(Xt, bstride); logdet_out = csd::bshift(logdet_out, bstride);
  status = csd::bshift(status, bstride);
  constexpr int NA = TILE / 16;
  CS_DYN_SMEM(double, sm);
  double* colbuf = sm;              // [2][TILE]
  double* rowbuf = sm + 2 * TILE;   // [2][TILE]
  double* piv = sm + 4 * TILE;      // [TILE]
  double* dinv = sm + 5 * TILE;     // [TILE]
  double* red = sm + 6 * TILE;      // [32]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  double reg[NA][NA];
#pragma unroll
  for (int b = 0; b < NA; ++b)
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int r = tx + 16 * a, c = ty + 16 * b;
      reg[a][b] = (r >= c) ? Gt[r + c * ldg] : 0.0;
      if (mode == 2 && r == c) dinv[c] = 1.0 / reg[a][b];   // L_cc read back: the same 1 / sqrt(pivot) as in mode 0
    }
  if (mode != 2) {

  // ---- phase 1: right-looking Cholesky; columns stay unscaled (L[r][j] = S[r][j] / sqrt(d_j)) ----
  // Split by the 16-column block bj so every register index / loop bound is a compile-time
  // constant, and software-pipelined: at step j a thread first updates and publishes column
  // j+1 (the only data the next step waits for) and defers the bulk of its rank-1 update to
  // the next step, where it overlaps the latency chain barrier -> pivot -> reciprocal.
  if (ty == 0) {
#pragma unroll
    for (int a = 0; a < NA; ++a) colbuf[tx + 16 * a] = reg[a][0];
  }
  {
    double crp[NA], ccp[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) { crp[a] = 0.0; ccp[a] = 0.0; }
#pragma unroll
    for (int bj = 0; bj < NA; ++bj) {
      for (int tj = 0; tj < 16; ++tj) {
        const int j = 16 * bj + tj;
        const double* buf = colbuf + (j & 1) * TILE;
        double* nbuf = colbuf + ((j + 1) & 1) * TILE;
        __syncthreads();
        double d = buf[j];
        if (!(d > 0.0)) {
          const int gidx = pivot_base + j;
          if (tid == 0 && gidx < n_real) atomicMin(status, gidx + 1);  // LAPACK-style 1-based info
          d = 1.0;
        }
        if (tid == 0) piv[j] = d;
        const double invd = csd::fast_rcp(d);
        // deferred bulk of step j-1: columns c > j
        if (ty > tj) {
#pragma unroll
          for (int a = bj; a < NA; ++a) reg[a][bj] -= crp[a] * ccp[bj];
        }
#pragma unroll
        for (int b = bj + 1; b < NA; ++b)
#pragma unroll
          for (int a = b; a < NA; ++a) reg[a][b] -= crp[a] * ccp[b];
        // this step's column
#pragma unroll
        for (int a = 0; a < NA; ++a) crp[a] = (a >= bj) ? buf[tx + 16 * a] : 0.0;
#pragma unroll
        for (int b = 0; b < NA; ++b) ccp[b] = (b >= bj) ? buf[ty + 16 * b] * invd : 0.0;
        // early update + publication of column j+1
        if (tj < 15) {
          if (ty == tj + 1) {
#pragma unroll
            for (int a = bj; a < NA; ++a) { reg[a][bj] -= crp[a] * ccp[bj]; nbuf[tx + 16 * a] = reg[a][bj]; }
          }
        } else if (bj + 1 < NA) {
          if (ty == 0) {
#pragma unroll
            const int b1 = (bj + 1 < NA) ? bj + 1 : bj;  // compile-time after unrolling
#pragma unroll
            for (int a = 0; a < NA; ++a)
              if (a >= b1) { reg[a][b1] -= crp[a] * ccp[b1]; nbuf[tx + 16 * a] = reg[a][b1]; }
          }
        }
      }
    }
  }
  __syncthreads();
  // scale the columns, emit L (explicit zeros above the diagonal), 1/L_cc, log det
#pragma unroll
  for (int b = 0; b < NA; ++b) {
    const int c = ty + 16 * b;
    const double sq = sqrt(piv[c]);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int r = tx + 16 * a;
      double v = 0.0;
      if (r == c) { v = sq; dinv[c] = 1.0 / sq; }
      else if (r > c) v = reg[a][b] / sq;
      reg[a][b] = v;
      Gt[r + c * ldg] = v;
    }
  }
  const double ldsum = csd::block_sum(tid < TILE ? 0.5 * log(piv[tid < TILE ? tid : 0]) : 0.0, red);
  if (tid == 0) *logdet_out = ldsum;  // block_sum ends with a barrier: dinv is visible
  if (mode == 1) return;
  } else {
    __syncthreads();  // dinv visible
  }

  // ---- phase 2: in-place inverse X = L^-1, row by row with rank-1 updates of the rows below ----
  // Same software pipelining: at step i the owners of row i+1 apply step i's update to their
  // row, finalise and publish it (and column i+1 of L is published); the bulk update of the
  // rows below is deferred to the next step, off the barrier -> load -> fma -> store chain.
  // Slot (r,c) of `reg` holds L[r][c] until step c, then the running sum Acc[r][c], then X[r][c].
  if (tid == 0) { reg[0][0] = dinv[0]; rowbuf[0] = dinv[0]; }
  if (ty == 0) {
#pragma unroll
    for (int a = 0; a < NA; ++a) if (a > 0 || tx > 0) colbuf[tx + 16 * a] = reg[a][0];
  }
  {
    double cbp[NA], rbp[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) { cbp[a] = 0.0; rbp[a] = 0.0; }
#pragma unroll
    for (int bi = 0; bi < NA; ++bi) {
      for (int ti = 0; ti < 16; ++ti) {
        const int i = 16 * bi + ti;
        const double* cbuf = colbuf + (i & 1) * TILE;
        const double* rbuf = rowbuf + (i & 1) * TILE;
        double* ncbuf = colbuf + ((i + 1) & 1) * TILE;
        double* nrbuf = rowbuf + ((i + 1) & 1) * TILE;
        __syncthreads();
        // deferred bulk of step i-1: rows r > i, columns c <= i-1 (column i-1 starts from 0)
#pragma unroll
        for (int b = 0; b < bi; ++b) {
          const double xv = rbp[b];
          const bool first = (ti == 0) && (b == bi - 1) && (ty == 15);
          if (tx > ti) reg[bi][b] = (first ? 0.0 : reg[bi][b]) + cbp[bi] * xv;
#pragma unroll
          for (int a = bi + 1; a < NA; ++a) reg[a][b] = (first ? 0.0 : reg[a][b]) + cbp[a] * xv;
        }
        if (ti >= 1 && ty <= ti - 1) {
          const double xv = rbp[bi];
          const bool first = (ty == ti - 1);
          if (tx > ti) reg[bi][bi] = (first ? 0.0 : reg[bi][bi]) + cbp[bi] * xv;
#pragma unroll
          for (int a = bi + 1; a < NA; ++a) reg[a][bi] = (first ? 0.0 : reg[a][bi]) + cbp[a] * xv;
        }
        // this step's column of L (rows > i) and row of X (columns <= i)
#pragma unroll
        for (int a = 0; a < NA; ++a) cbp[a] = (a >= bi) ? cbuf[tx + 16 * a] : 0.0;
#pragma unroll
        for (int b = 0; b < NA; ++b) rbp[b] = (b <= bi) ? rbuf[ty + 16 * b] : 0.0;
        if (i + 1 < TILE) {
          const double x1 = dinv[i + 1];
          if (ti < 15) {
            // row i+1 = tx == ti+1 in block bi; column i+1 = ty == ti+1 in block bi
            if (tx == ti + 1) {
#pragma unroll
              for (int b = 0; b <= bi; ++b) {
                const int c = ty + 16 * b;
                if (b < bi || ty <= ti) {
                  const double acc = ((b == bi && ty == ti) ? 0.0 : reg[bi][b]) + cbp[bi] * rbp[b];
                  const double v = -x1 * acc;
                  reg[bi][b] = v; nrbuf[c] = v;
                }
              }
              if (ty == ti + 1) { reg[bi][bi] = x1; nrbuf[i + 1] = x1; }
            }
            if (ty == ti + 1) {
#pragma unroll
              for (int a = bi; a < NA; ++a) if (a > bi || tx > ti + 1) ncbuf[tx + 16 * a] = reg[a][bi];
            }
          } else {
            const int a1 = (bi + 1 < NA) ? bi + 1 : bi;  // compile-time after unrolling
            if (tx == 0) {
#pragma unroll
              for (int b = 0; b <= bi; ++b) {
                const int c = ty + 16 * b;
                const double acc = ((b == bi && ty == 15) ? 0.0 : reg[a1][b]) + cbp[a1] * rbp[b];
                const double v = -x1 * acc;
                reg[a1][b] = v; nrbuf[c] = v;
              }
              if (ty == 0) { reg[a1][a1] = x1; nrbuf[i + 1] = x1; }
            }
            if (ty == 0) {
#pragma unroll
              for (int a = 0; a < NA; ++a) if (a >= a1 && (a > a1 || tx > 0)) ncbuf[tx + 16 * a] = reg[a][a1];
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NA; ++b)
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int r = tx + 16 * a, c = ty + 16 * b;
      Xt[r + c * ldx] = (r >= c) ? reg[a][b] : 0.0;
    }
}

// ---------------------------------------------------------------------------------
// Triangular solve of the in-panel block rows: X L_kk^T = A, in place (L_ik = A_ik L_kk^-T), one warp per row of A.
// Replaces "multiply by the explicit inverse X_kk" on the critical chain, so that the inversion of the diagonal tile
// (the second half of the leaf) no longer sits between the factorization of tile k and the panel below it.
// Column-oriented forward substitution: x_j = a_j / L_jj is broadcast from its owner lane, then a_l -= x_j L_lj for l > j
// (column j of L is contiguous; every warp of the device reads the same 128 KB tile, which stays in L1 / L2).
// ---------------------------------------------------------------------------------
constexpr int TRSM_THREADS = 128;
constexpr int TRSM_ROWS = 4;   // rows of A per warp: every column of L read from shared memory serves 4 rows
// the lower triangle of the tile, packed by columns: element (l, j), l >= j, at j * TILE - j (j + 1) / 2 + l
template <int TILE> constexpr int trsm_smem_bytes() { return TILE * (TILE + 1) / 2 * (int)sizeof(double); }

// Measured on the B200 (tools/trsm_bench.cu, tools/lat_bench.cu): one substitution step is a 43-cycle dependent chain
// (SHFL + DMUL + DFMA) but about 14 issue slots per row, so the kernel is issue-bound as soon as an SM sub-partition holds more
// than one warp of it: 16 warps per CTA took 31 us whatever the row count.  Hence small CTAs (4 warps x 4 rows) spread over all
// SMs of the partition, with the packed triangle (66 KB at TILE = 128) leaving room for three of them per SM.  The tile of L is
// staged with cp.async, all copies in flight at once: reading the columns from global memory inside the loop cost one L2 round
// trip per column on every SM.
template <int TILE>
__global__ void __launch_bounds__(TRSM_THREADS)
trsm_kernel(const double* __restrict__ Lkk, long long ldl, double* __restrict__ A, long long lda, int nrows,
            const int* __restrict__ skip, long long bstride) {
  skip = csd::bshift(skip, bstride);
  if (skip && *skip) return;
  Lkk = csd::bshift(Lkk, bstride); A = csd::bshift(A, bstride);
  constexpr int NS = TILE / 32, R = TRSM_ROWS, NW = TRSM_THREADS / 32;
  CS_DYN_SMEM(double, Ls);
  __shared__ double rinv[TILE];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int c = wid; c < TILE; c += NW) {
    double* dst = Ls + c * TILE - c * (c + 1) / 2;
    const double* src = Lkk + (long long)c * ldl;
    for (int r = c + lane; r < TILE; r += 32) csd::cp_async8(dst + r, src + r);
  }
  csd::cp_async_commit();
  const int r0 = ((int)blockIdx.x * NW + wid) * R;   // nrows is a multiple of the tile, hence of R
  const bool live = r0 < nrows;
  double a[R][NS];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) a[q][s2] = live ? A[r0 + q + (long long)(lane + 32 * s2) * lda] : 0.0;
  csd::cp_async_wait<0>();
  __syncthreads();
  if (tid < TILE) rinv[tid] = 1.0 / Ls[tid * TILE - tid * (tid + 1) / 2 + tid];
  __syncthreads();
  if (!live) return;
#pragma unroll
  for (int s0 = 0; s0 < NS; ++s0) {
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
      const int j = 32 * s0 + jj;
      const double* col = Ls + j * TILE - j * (j + 1) / 2;
      double lc[NS];
#pragma unroll
      for (int s2 = 0; s2 < NS; ++s2) lc[s2] = (s2 >= s0 && lane + 32 * s2 > j) ? col[lane + 32 * s2] : 0.0;
      const double ri = rinv[j];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double xj = __shfl_sync(0xffffffffu, a[q][s0], jj) * ri;
        if (lane == jj) a[q][s0] = xj;
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2)
          if (s2 >= s0) a[q][s2] = fma(-xj, lc[s2], a[q][s2]);   // lc is zero at and above the diagonal
      }
    }
  }
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) A[r0 + q + (long long)(lane + 32 * s2) * lda] = a[q][s2];
}

// ---------------------------------------------------------------------------------
// Host-side plan
// ---------------------------------------------------------------------------------
enum GemmKind { GK_NT = 0, GK_NN = 1, GK_TN = 2 };

static std::atomic<long long> g_launches{0};  // kernels launched by this library (claimed count for bench.py)

enum OpType { OP_GEMM = 0, OP_LEAF = 1, OP_RECORD = 2, OP_WAIT = 3, OP_COPY = 4, OP_TRSM = 5 };

struct Launch {
  int is_leaf;     // OpType (1: leaf_kernel(blk), 0: gemm group, 2/3: event record / wait)
  int blk;         // leaf block index
  int kind;        // GemmKind
  int first_prob;  // index into the problem array
  int nprob;
  int ntiles;
  int stream;      // 0: main (critical path, high priority)   1: auxiliary (bulk updates)
  int ev;          // event index for OP_RECORD / OP_WAIT
  int gtile;       // CTA tile of this launch (0 = engine default gt)
  int leaf_mode;   // OP_LEAF: 0 factor + invert, 1 factor only, 2 invert only
  int r0, nb;      // OP_TRSM: block rows r0 .. r0+nb-1 of block column blk
  // OP_COPY: 2-D device copy of `rows` x `cols` doubles (leading dimension np on both sides)
  double* cp_dst; const double* cp_src; long long cp_rows, cp_cols, cp_ldd, cp_lds;
};

inline int pick_tile(int n) { return n <= 1536 ? 64 : 128; }

// CTAs per SM the 64 x 64 kernels are compiled for (__launch_bounds__ register cap: 4 -> 128 registers, 5 -> 96 with a few
// spilled descriptor fields)
constexpr int GEMM64_MINB = 4;

template <int TILE>
inline void launch_gemm_t(int kind, const GemmProblem* dprobs, int nprob, int ntiles_, cudaStream_t st,
                          const int* skip, int nbatch) {
  if (ntiles_ <= 0) return;
  const dim3 ntiles(ntiles_, 1, nbatch);
  if (TILE == 128) {
    if (kind == GK_NT) { auto k = csg::gemm_kernel<128, 2, 4, false, false>; CS_LAUNCH(k, ntiles, 256, (csg::gemm_smem_bytes<128, false, false>()), st, dprobs, nprob, skip); }
    else if (kind == GK_NN) { auto k = csg::gemm_kernel<128, 2, 4, false, true>; CS_LAUNCH(k, ntiles, 256, (csg::gemm_smem_bytes<128, false, true>()), st, dprobs, nprob, skip); }
    else { auto k = csg::gemm_kernel<128, 2, 4, true, true>; CS_LAUNCH(k, ntiles, 256, (csg::gemm_smem_bytes<128, true, true>()), st, dprobs, nprob, skip); }
  } else {
    // 64 x 64 CTA tiles, 128 registers/thread: 4 CTAs per SM hide the prologue / epilogue of one tile behind the main loop
    // of another.  Two stages everywhere: with the CTA count fixed by the registers, every extra stage was slower on the
    // B200 (K = 512 whole waves, NT: 2 stages 32.6, 3 stages 32.0, 4 stages 31.5 TFLOP/s; tools/gemm_bench)
    // (keeping three stages for the chain's one-to-twelve-tile launches made no measurable difference: n = 4096 4.5 vs 4.6 ms)
    if (kind == GK_NT) { auto k = csg::gemm_kernel<64, 2, 2, false, false, 16, 2, GEMM64_MINB>; CS_LAUNCH(k, ntiles, 128, (csg::gemm_smem_bytes<64, false, false, 16, 2>()), st, dprobs, nprob, skip); }
    else if (kind == GK_NN) { auto k = csg::gemm_kernel<64, 2, 2, false, true, 16, 2, GEMM64_MINB>; CS_LAUNCH(k, ntiles, 128, (csg::gemm_smem_bytes<64, false, true, 16, 2>()), st, dprobs, nprob, skip); }
    else { auto k = csg::gemm_kernel<64, 2, 2, true, true, 16, 2, GEMM64_MINB>; CS_LAUNCH(k, ntiles, 128, (csg::gemm_smem_bytes<64, true, true, 16, 2>()), st, dprobs, nprob, skip); }
  }
}

// 32 x 32 CTA tiles (NT only): four times the CTAs of a 64-tile launch, for the two small multiplies the Cholesky chain
// waits for at every panel boundary -- they share their SMs with resident bulk CTAs, so their latency is set by how
// thinly they spread, not by per-CTA efficiency
inline void launch_gemm_32(const GemmProblem* dprobs, int nprob, int ntiles_, cudaStream_t st, const int* skip, int nbatch) {
  if (ntiles_ <= 0) return;
  auto k = csg::gemm_kernel<32, 2, 2, false, false, 16, 3>;
  CS_LAUNCH(k, dim3(ntiles_, 1, nbatch), 128, (csg::gemm_smem_bytes<32, false, false, 16, 3>()), st, dprobs, nprob, skip);
}

inline void launch_gemm(int tile, int kind, const GemmProblem* dprobs, int nprob, int ntiles, cudaStream_t st,
                        const int* skip = nullptr, int nbatch = 1) {
  ++g_launches;
  if (tile == 32 && kind == GK_NT) launch_gemm_32(dprobs, nprob, ntiles, st, skip, nbatch);
  else if (tile == 128) launch_gemm_t<128>(kind, dprobs, nprob, ntiles, st, skip, nbatch);
  else launch_gemm_t<64>(kind, dprobs, nprob, ntiles, st, skip, nbatch);
}

// opt-in to > 48 KB dynamic shared memory (per device; cheap, called at every engine init)
inline void configure_kernels() {
  { auto k = csg::gemm_kernel<128, 2, 4, false, false>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<128, false, false>())); }
  { auto k = csg::gemm_kernel<128, 2, 4, false, true>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<128, false, true>())); }
  { auto k = csg::gemm_kernel<128, 2, 4, true, true>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<128, true, true>())); }
  { auto k = csg::gemm_kernel<64, 2, 2, false, false, 16, 2, GEMM64_MINB>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<64, false, false, 16, 2>())); }
  { auto k = csg::gemm_kernel<64, 2, 2, false, true, 16, 2, GEMM64_MINB>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<64, false, true, 16, 2>())); }
  { auto k = csg::gemm_kernel<64, 2, 2, true, true, 16, 2, GEMM64_MINB>; CS_SET_SMEM(k, (csg::gemm_smem_bytes<64, true, true, 16, 2>())); }
  { auto k = leaf_kernel<128>; CS_SET_SMEM(k, leaf_smem_bytes<128>()); }
  { auto k = leaf_kernel<64>; CS_SET_SMEM(k, leaf_smem_bytes<64>()); }
  { auto k = trsm_kernel<128>; CS_SET_SMEM(k, trsm_smem_bytes<128>()); }
}

inline void launch_leaf(int tile, double* Gt, long long ldg, double* Xt, long long ldx, int pivot_base,
                        double* logdet_out, int* status, int n_real, cudaStream_t s, const int* skip, int nbatch = 1,
                        long long bstride = 0, int mode = 0) {
  ++g_launches;
  const dim3 grid(1, 1, nbatch);
  if (tile == 128) { auto k = leaf_kernel<128>; CS_LAUNCH(k, grid, LEAF_THREADS, leaf_smem_bytes<128>(), s, Gt, ldg, Xt, ldx, pivot_base, logdet_out, status, n_real, skip, bstride, mode); }
  else { auto k = leaf_kernel<64>; CS_LAUNCH(k, grid, LEAF_THREADS, leaf_smem_bytes<64>(), s, Gt, ldg, Xt, ldx, pivot_base, logdet_out, status, n_real, skip, bstride, mode); }
}

inline void launch_trsm(int tile, const double* Lkk, long long ldl, double* A, long long lda, int nrows, cudaStream_t s,
                        const int* skip, int nbatch = 1, long long bstride = 0) {
  if (nrows <= 0) return;
  ++g_launches;
  const int per_cta = (TRSM_THREADS / 32) * TRSM_ROWS;
  const dim3 grid((nrows + per_cta - 1) / per_cta, 1, nbatch);
  if (tile == 128) { auto k = trsm_kernel<128>; CS_LAUNCH(k, grid, TRSM_THREADS, trsm_smem_bytes<128>(), s, Lkk, ldl, A, lda, nrows, skip, bstride); }
  else { auto k = trsm_kernel<64>; CS_LAUNCH(k, grid, TRSM_THREADS, trsm_smem_bytes<64>(), s, Lkk, ldl, A, lda, nrows, skip, bstride); }
}

struct DenseEngine {
  int n = 0, np = 0, tile = 0, N = 0;  // tile: block size of the leaf / panel structure
  int gt = 0, rr = 1;                   // gt: CTA tile of the GEMM kernel, rr = tile / gt
  double *G = nullptr, *X = nullptr, *C = nullptr;  // device, np x np each
  double* logdet_part = nullptr;                     // device, N
  double* Pbuf = nullptr;                            // device, np x (panel_blocks*tile): out-of-place panel multiply
  double* Rbuf = nullptr;                            // device, (panel_blocks*tile) x np: out-of-place row panel of L^-1
  int* status = nullptr;                             // device, 1 (0 = OK, else 1-based failing pivot)
  GemmProblem* d_probs = nullptr;
  std::vector<GemmProblem> probs;
  std::vector<Launch> potrf, trtri, xtx;
  std::vector<Launch> potrf_inv;    // Cholesky with the inverse pipelined behind it (inv_pipeline)
  int inv_pipeline = 0;
  int plan_version = 2;
  int tail_a_max = 0, tail_from = 0;                 // tile columns of K^-1 accumulated on the chain stream after the chain (0 = off): tail_a_max from panel tail_from on, one fewer before
  struct TailPiece { int ev_rows, ev_prev; std::vector<GemmProblem> probs; };
  std::vector<TailPiece> tail_pieces;
  bool split_u2b = false;                            // look-ahead of depth three: U2b-rest on its own stream S7
  bool top32 = false;                                // Ltop / U1top with 32 x 32 CTA tiles (tile 128 plans of single fits)
  // batched engine (cs_fit_create_batch): nbatch fits share the plan; their buffers are bstride bytes apart
  int nbatch = 1;
  long long bstride = 0;
  bool external_buffers = false;  // G, X, C, logdet_part, status, Pbuf belong to the caller's slab
  bool single_panel = false;      // one panel = the whole matrix: X = L^-1 is complete after the Cholesky plan
  bool left_looking = false;      // batched handles: left-looking Cholesky (one growing-K update per block column)
  bool reset_status = true;       // false while the optimisation loop runs: a non-PD pivot of ANY iteration must survive
  int panel_blocks = 4;             // outer panel width in blocks (maximum)
  std::vector<int> bounds;          // plan v2: panel boundaries in blocks (graded: narrow first and last panels)
  int n_events = 0;
  std::vector<cudaEvent_t> events;
  cudaStream_t aux8{};  // S8: inversion of the diagonal tiles (leaf mode 2), partition A, off the critical chain
  bool aux8_ok = false;
  bool split_leaf = false;  // plan v2: chain = factor (leaf mode 1) + triangular solve; the tile inverses run on S8
  cudaStream_t aux{}, aux2{}, aux3{}, aux4{}, aux5{}, aux6{}, aux7{};  // S7: U2b-rest (bulk trailing update behind the look-ahead);  // S6: second chain stream (X_PP), partition A;  // S4: K^-1 accumulation (lowest); S5: Acc rows below the next panel; S1: bulk trailing updates (low); S2: panel rest (high); S3: pipelined inverse (low)
  bool aux_ok = false, aux2_ok = false, aux3_ok = false, aux4_ok = false, aux5_ok = false, aux6_ok = false, aux7_ok = false;
  // SM partition (rt::partition_streams): the critical chain of the Cholesky (leaf + one-tile GEMMs, plan
  // stream 0) runs on `crit`, a stream of the small partition A; S1..S3 live on partition B.
  cudaStream_t crit{};
  bool crit_ok = false;
  int sm_crit = 0, sm_bulk = 0;
  cudaEvent_t ev_in{}, ev_out{};
  cudaEvent_t ev_gram{};  // split Gram build (api.cu): S2 starts behind the handle's stream
  bool ev_gram_ok = false;

  static GemmProblem mk(const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
                        int mt, int nt, int K, int kmode, int order, int mirror, double alpha, double beta) {
    GemmProblem p{};
    p.A = A; p.B = B; p.C = C; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.mt = mt; p.nt = nt; p.K = K; p.kmode = kmode; p.order = order; p.mirror = mirror;
    p.set_scale(alpha, beta); p.tile_base = 0;
    p.ntiles = (order == csg::ORD_LOWER) ? csg::lower_tiles(mt) : mt * nt;
    return p;
  }

  void add_group(std::vector<Launch>& seq, int kind, std::vector<GemmProblem>& group) {
    if (group.empty()) return;
    Launch L{};
    L.is_leaf = OP_GEMM; L.kind = kind; L.first_prob = (int)probs.size(); L.nprob = (int)group.size();
    int base = 0;
    for (auto& p : group) { p.tile_base = base; p.bstride = bstride; base += p.ntiles; probs.push_back(p); }
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

  static int padded(int n_, int tile_) { return (n_ + tile_ - 1) / tile_ * tile_; }
  static int panel_width() {
    int w = 4;
    if (const char* pb = std::getenv("CS_PANEL")) { const int v = std::atoi(pb); if (v >= 1 && v <= 16) w = v; }
    return w;
  }
  static size_t pbuf_bytes(int np_, int tile_) { return sizeof(double) * (size_t)np_ * panel_width() * tile_; }

  // returns 0 or a runtime error code
  int init(int n_, int tile_, int gemm_tile = 64, int inv_pipe = 0, int partition = 0) {
    inv_pipeline = inv_pipe;
    panel_blocks = panel_width();
    n = n_; tile = tile_; N = (n + tile - 1) / tile; np = N * tile;
    // A single small fit (the README / Hill n = 747 configurations) is bound by the latency of its launch chain, not by
    // throughput: one panel spanning the whole matrix removes the panel / trailing-update launches, and the chain's
    // second stream then produces the complete L^-1, so the recursive inverse is skipped (n = 672: 0.88 -> 0.70 ms per
    // evaluation).  Batched handles keep W = 4 (measured best for throughput).
    single_panel = (nbatch == 1 || std::getenv("CS_SINGLE_PANEL") != nullptr) && N <= 12 && N > 1 && std::getenv("CS_PANEL") == nullptr &&
                   std::getenv("CS_PLAN") == nullptr;
    if (single_panel) panel_blocks = N;
    gt = (gemm_tile == 128 && tile == 128) ? 128 : 64;
    rr = tile / gt;
    const size_t bytes = (size_t)np * np * sizeof(double);
    int e;
    if (!external_buffers) {
      if ((e = rt::dev_malloc((void**)&G, bytes))) return e;
      if ((e = rt::dev_malloc((void**)&X, bytes))) return e;
      if ((e = rt::dev_malloc((void**)&C, bytes))) return e;
      if ((e = rt::dev_malloc((void**)&logdet_part, sizeof(double) * N))) return e;
      if ((e = rt::dev_malloc((void**)&Pbuf, pbuf_bytes(np, tile)))) return e;
      if ((e = rt::dev_malloc((void**)&Rbuf, pbuf_bytes(np, tile)))) return e;
      if ((e = rt::dev_malloc((void**)&status, sizeof(int)))) return e;
    }
    configure_kernels();
    const long long ld = np;
    const long long T = tile;
    std::vector<GemmProblem> grp;
    // --- two-level blocked right-looking Cholesky, three streams ---
    // Outer panels of W blocks.  Inside a panel, step k is split into a *critical* part on the
    // main stream S0 -- leaf(k), the panel multiply of block row k+1 only, the update of the
    // next diagonal tile -- and the *rest* (panel rows >= k+2, remaining inner updates) on the
    // panel stream S2, which runs while S0 is already inside leaf(k+1).  After a panel the
    // trailing update (k = W*TILE) is split into U1 (next panel's columns, S0) and U2 (the
    // rest, bulk stream S1).  Dependencies are events; the whole sequence is graph-capturable.
    const int W = panel_blocks;
    int nev = 0;
    auto build_chol = [&](std::vector<Launch>& seq, bool with_inverse) {
    auto op = [&](int type, int stream, int ev) { Launch o{}; o.is_leaf = type; o.stream = stream; o.ev = ev; seq.push_back(o); };
    int npanel = 0, last_aux_ev = -1, last_inv_ev = -1, last_inv3_ev = -1, last_rest_ev = -1;
    // in-place panel multiply L = A * X_kk^T for `nb` block rows starting at block row r0 of column k
    auto push_panel = [&](int k, int r0, int nb, int stream) {
      if (nb <= 0) return;
      double* pan = G + (long long)r0 * T + (long long)k * T * ld;
      double* Xkk = X + (long long)k * T * (ld + 1);
      if (rr == 2) {
        // 64 x 64 CTA tiles: X_kk is lower triangular, so output columns 64..127 need all input
        // columns but output columns 0..63 only input columns 0..63: right half first, then the
        // left half with K = 64 (each launch is in-place safe per CTA).
        grp.push_back(mk(pan, ld, Xkk + gt, ld, pan + gt * ld, ld, nb * rr, 1, tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
        add_group(seq, GK_NT, grp); seq.back().stream = stream;
        grp.push_back(mk(pan, ld, Xkk, ld, pan, ld, nb * rr, 1, gt, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
        add_group(seq, GK_NT, grp); seq.back().stream = stream;
      } else {
        grp.push_back(mk(pan, ld, Xkk, ld, pan, ld, nb, 1, tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
        add_group(seq, GK_NT, grp); seq.back().stream = stream; seq.back().gtile = tile;
      }
    };
    for (int p0 = 0; p0 < N; p0 += W, ++npanel) {
      const int p1 = std::min(p0 + W, N);
      int ev_rest_prev = -1;
      for (int k = p0; k < p1; ++k) {
        { Launch L{}; L.is_leaf = OP_LEAF; L.blk = k; seq.push_back(L); }
        const int m = N - k - 1;          // block rows below k
        if (m <= 0) continue;
        const int ev_leaf = nev++, ev_pp = nev++, ev_rest = nev++;
        op(OP_RECORD, 0, ev_leaf);
        // ---- critical part (S0): block row k+1 ----
        if (ev_rest_prev >= 0) op(OP_WAIT, 0, ev_rest_prev);  // tile (k+1,k) carries the rest-updates of step k-1
        push_panel(k, k + 1, 1, 0);
        op(OP_RECORD, 0, ev_pp);
        const int nc = p1 - 1 - k;        // block columns k+1 .. p1-1 still inside this panel
        double* row1 = G + (k + 1) * T + k * T * ld;  // L_{k+1,k}
        if (nc > 0) {
          grp.push_back(mk(row1, ld, row1, ld, G + (k + 1) * T * (ld + 1), ld, rr, rr, tile, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
          add_group(seq, GK_NT, grp);
        }
        // ---- rest (S2): block rows >= k+2 ----
        op(OP_WAIT, 2, ev_leaf);
        push_panel(k, k + 2, m - 1, 2);
        if (nc > 0) {
          op(OP_WAIT, 2, ev_pp);
          double* row2 = G + (k + 2) * T + k * T * ld;  // L_{k+2..,k}
          const int mr = N - p1;
          if (nc >= 2) {
            // rows k+2..p1-1 x column k+1, and the lower part of rows/cols k+2..p1-1
            grp.push_back(mk(row2, ld, row1, ld, G + (k + 2) * T + (k + 1) * T * ld, ld, (nc - 1) * rr, rr, tile, csg::KM_FULL,
                             csg::ORD_COL, 0, -1.0, 1.0));
            grp.push_back(mk(row2, ld, row2, ld, G + (k + 2) * T * (ld + 1), ld, (nc - 1) * rr, (nc - 1) * rr, tile, csg::KM_FULL,
                             csg::ORD_LOWER, 0, -1.0, 1.0));
          }
          if (mr > 0)
            grp.push_back(mk(G + p1 * T + k * T * ld, ld, row1, ld, G + p1 * T + (k + 1) * T * ld, ld, mr * rr, nc * rr, tile,
                             csg::KM_FULL, csg::ORD_COL, 0, -1.0, 1.0));
          if (!grp.empty()) { add_group(seq, GK_NT, grp); seq.back().stream = 2; }
        }
        op(OP_RECORD, 2, ev_rest);
        ev_rest_prev = ev_rest;
      }
      if (ev_rest_prev >= 0) op(OP_WAIT, 0, ev_rest_prev);  // S0 rejoins S2: the panel is complete
      if (with_inverse) {
        // ---- pipelined inverse (stream S3): as soon as panel P of L is final, rows P of X = L^-1
        // are finalised and their contribution to the rows below (Acc) and to K^-1 = X^T X is
        // accumulated with K = W*TILE -- bulk work that fills the SMs the Cholesky chain leaves idle.
        const int ev_done = nev++;
        op(OP_RECORD, 0, ev_done);
        op(OP_WAIT, 3, ev_done);
        const int pw = p1 - p0;
        for (int k = p0; k < p1; ++k) {
          double* Xkk = X + (long long)k * T * (ld + 1);
          double* rowk = X + (long long)k * T;  // block row k of X, column 0
          if (k > 0) {  // X_kJ = -X_kk Acc_kJ, J < k  (in place: one CTA per whole tile)
            grp.push_back(mk(Xkk, ld, rowk, ld, rowk, ld, 1, k, tile, csg::KM_FULL, csg::ORD_COL, 0, -1.0, 0.0));
            add_group(seq, GK_NN, grp); seq.back().stream = 3; seq.back().gtile = tile;
          }
          const int nr = p1 - 1 - k;
          if (nr > 0) {  // Acc_iJ += L_ik X_kJ for the rows i in (k, p1) of this panel, J <= k
            grp.push_back(mk(G + (k + 1) * T + k * T * ld, ld, rowk, ld, X + (k + 1) * T, ld, nr * rr, (k + 1) * rr, tile,
                             csg::KM_FULL, csg::ORD_COL, 0, 1.0, 1.0));
            add_group(seq, GK_NN, grp); seq.back().stream = 3;
          }
        }
        // rows P of X are final here: their contribution to K^-1 goes to the accumulation stream S4 (lowest
        // priority, nobody waits for it before the end) so that it overlaps the row chain of the next panel
        const int ev_rows = nev++;
        op(OP_RECORD, 3, ev_rows);
        op(OP_WAIT, 4, ev_rows);
        {  // C_IJ += sum_{k in P} X_kI^T X_kJ for J <= I <= k  (mirrored store keeps C symmetric)
          double* Xp = X + p0 * T;              // rows of the panel, column 0
          double* Xd = X + p0 * T * (ld + 1);   // rows and columns of the panel
          if (p0 > 0) {
            grp.push_back(mk(Xp, ld, Xp, ld, C, ld, p0 * rr, p0 * rr, pw * tile, csg::KM_FULL, csg::ORD_LOWER, 1, 1.0, 1.0));
            GemmProblem q = mk(Xd, ld, Xp, ld, C + p0 * T, ld, pw * rr, p0 * rr, pw * tile, csg::KM_GE_I, csg::ORD_COL, 2, 1.0, 1.0);
            q.Cm = C + p0 * T * ld;  // transposed block: rows 0.., columns of the panel
            grp.push_back(q);
          }
          grp.push_back(mk(Xd, ld, Xd, ld, C + p0 * T * (ld + 1), ld, pw * rr, pw * rr, pw * tile, csg::KM_GE_I, csg::ORD_LOWER, 1, 1.0, 1.0));
          add_group(seq, GK_TN, grp); seq.back().stream = 4;
        }
        last_inv_ev = nev++;
        op(OP_RECORD, 4, last_inv_ev);
        // Acc_IJ += sum_{k in P} L_Ik X_kJ for I >= p1 (K = pw*TILE) feeds the later row chains.  Look-ahead:
        // the rows of the NEXT panel are updated on the chain stream S3 right away (AccNext), the rows
        // below them on the bulk stream S5 (AccRest); both accumulate into the same tiles of the panel
        // after next, so AccNext(P) is ordered after AccRest(P-1).
        const int mb = N - p1;
        if (mb > 0) {
          const int nn = std::min(W, mb);  // block rows of the next panel
          auto push_acc = [&](int r0, int nb_rows, int stream) {
            double* Lp = G + (long long)r0 * T + p0 * T * ld;
            if (p0 > 0)
              grp.push_back(mk(Lp, ld, X + p0 * T, ld, X + (long long)r0 * T, ld, nb_rows * rr, p0 * rr, pw * tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 1.0));
            grp.push_back(mk(Lp, ld, X + p0 * T * (ld + 1), ld, X + (long long)r0 * T + p0 * T * ld, ld, nb_rows * rr, pw * rr, pw * tile,
                             csg::KM_GE_J, csg::ORD_COL, 0, 1.0, 1.0));
            add_group(seq, GK_NN, grp); seq.back().stream = stream;
          };
          if (last_rest_ev >= 0) op(OP_WAIT, 3, last_rest_ev);
          push_acc(p1, nn, 3);
          last_inv3_ev = nev++;
          op(OP_RECORD, 3, last_inv3_ev);
          if (mb > nn) {
            op(OP_WAIT, 5, ev_rows);
            push_acc(p1 + nn, mb - nn, 5);
            last_rest_ev = nev++;
            op(OP_RECORD, 5, last_rest_ev);
          }
        }
      }
      const int mt = N - p1;  // trailing blocks
      if (mt <= 0) continue;
      const int kw = (p1 - p0) * tile;
      double* pan = G + p1 * T + p0 * T * ld;  // rows >= p1 of the finished panel
      const int n1 = std::min(W, mt);          // next panel's block columns
      const int ev_panel = nev++, ev_u2 = nev++, ev_u1 = nev++;
      op(OP_RECORD, 0, ev_panel);
      // U1: columns [p1, p1+n1): lower part + rectangle below, on the panel stream S2 (every SM of the
      // bulk partition, high priority); the critical stream only waits for it
      op(OP_WAIT, 2, ev_panel);
      if (last_aux_ev >= 0) op(OP_WAIT, 2, last_aux_ev);  // U2 of the previous panel also wrote these columns
      grp.push_back(mk(pan, ld, pan, ld, G + p1 * T * (ld + 1), ld, n1 * rr, n1 * rr, kw, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
      if (mt > n1)
        grp.push_back(mk(pan + n1 * T, ld, pan, ld, G + (p1 + n1) * T + p1 * T * ld, ld, (mt - n1) * rr, n1 * rr, kw, csg::KM_FULL,
                         csg::ORD_COL, 0, -1.0, 1.0));
      add_group(seq, GK_NT, grp); seq.back().stream = 2;
      op(OP_RECORD, 2, ev_u1);
      op(OP_WAIT, 0, ev_u1);
      // U2: everything to the right of the next panel (bulk stream S1)
      op(OP_WAIT, 1, ev_panel);
      if (mt > n1) {
        grp.push_back(mk(pan + n1 * T, ld, pan + n1 * T, ld, G + (p1 + n1) * T * (ld + 1), ld, (mt - n1) * rr, (mt - n1) * rr, kw,
                         csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
        add_group(seq, GK_NT, grp);
        seq.back().stream = 1;
      }
      op(OP_RECORD, 1, ev_u2);
      last_aux_ev = ev_u2;
    }
    // the main stream rejoins the bulk (and inverse) streams before anything reads the results
    if (last_aux_ev >= 0) op(OP_WAIT, 0, last_aux_ev);
    if (last_inv_ev >= 0) op(OP_WAIT, 0, last_inv_ev);
    if (last_inv3_ev >= 0) op(OP_WAIT, 0, last_inv3_ev);
    if (last_rest_ev >= 0) op(OP_WAIT, 0, last_rest_ev);
    };

    // ---- plan v2: the diagonal block of a panel (W x W tiles) is factored AND inverted by the chain
    // alone (streams S0 / S6 on the small partition), everything below it is ONE multiply with the
    // explicit inverse, L_IP = A_IP X_PP^T (out of place through Pbuf), so the chain never waits for
    // bulk-partition work inside a panel.  Look-ahead: the rows of the next panel are multiplied and
    // the next diagonal block is updated first (Ltop / U1top), the rest follows.
    auto build_chol2 = [&](std::vector<Launch>& seq, bool with_inverse) {
    auto op = [&](int type, int stream, int ev) { Launch o{}; o.is_leaf = type; o.stream = stream; o.ev = ev; seq.push_back(o); };
    auto copy_op = [&](int stream, double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols) {
      Launch o{}; o.is_leaf = OP_COPY; o.stream = stream; o.cp_dst = dst; o.cp_src = src; o.cp_rows = rows; o.cp_cols = cols;
      o.cp_ldd = ldd; o.cp_lds = lds; seq.push_back(o);
    };
    int last_aux_ev = -1, last_inv_ev = -1, last_inv3_ev = -1, last_rest_ev = -1, ev_u1top_prev = -1, last_s6_ev = -1, last_u2b_ev = -1, last_u2brest_ev = -1;
    int last_s8_ev = -1;
    auto push_panel = [&](int k, int r0, int nb, int stream) {  // in-place L = A X_kk^T for nb block rows of column k
      if (nb <= 0) return;
      double* pan = G + (long long)r0 * T + (long long)k * T * ld;
      double* Xkk = X + (long long)k * T * (ld + 1);
      if (rr == 2) {
        grp.push_back(mk(pan, ld, Xkk + gt, ld, pan + gt * ld, ld, nb * rr, 1, tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
        add_group(seq, GK_NT, grp); seq.back().stream = stream;
        grp.push_back(mk(pan, ld, Xkk, ld, pan, ld, nb * rr, 1, gt, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
        add_group(seq, GK_NT, grp); seq.back().stream = stream;
      } else {
        grp.push_back(mk(pan, ld, Xkk, ld, pan, ld, nb, 1, tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
        add_group(seq, GK_NT, grp); seq.back().stream = stream; seq.back().gtile = tile;
      }
    };
    for (size_t pi = 0; pi + 1 < bounds.size(); ++pi) {
      const int p0 = bounds[pi], p1 = bounds[pi + 1];
      const int pw = p1 - p0;
      const int w_next = pi + 2 < bounds.size() ? bounds[pi + 2] - bounds[pi + 1] : 0;   // width of the next panel
      const int w_next2 = pi + 3 < bounds.size() ? bounds[pi + 3] - bounds[pi + 2] : 0;  // and of the one after
      const int w_next3 = pi + 4 < bounds.size() ? bounds[pi + 4] - bounds[pi + 3] : 0;  // and of the third ahead
      double* Xpp = X + (long long)p0 * T * (ld + 1);   // diagonal block of X = inverse of the diagonal block of L
      // ---- D(P) on the chain (S0): factor the diagonal block; S6 builds the off-diagonal tiles of X_PP
      if (ev_u1top_prev >= 0) op(OP_WAIT, 0, ev_u1top_prev);
      if (last_s6_ev >= 0) op(OP_WAIT, 0, last_s6_ev);   // keeps the happens-before relation of S6 simple
      for (int k = p0; k < p1; ++k) {
        { Launch L{}; L.is_leaf = OP_LEAF; L.blk = k; L.leaf_mode = split_leaf ? 1 : 0; seq.push_back(L); }
        const int ev_leaf = nev++;
        op(OP_RECORD, 0, ev_leaf);
        int ev_inv = -1;
        if (split_leaf) {  // X_kk = L_kk^-1 on S8, beside the chain
          op(OP_WAIT, 8, ev_leaf);
          { Launch L{}; L.is_leaf = OP_LEAF; L.blk = k; L.leaf_mode = 2; L.stream = 8; seq.push_back(L); }
          ev_inv = nev++;
          op(OP_RECORD, 8, ev_inv);
          last_s8_ev = ev_inv;
          if (k == p0) op(OP_WAIT, 6, ev_inv);  // S6 (and everything ordered behind its ev_xpp) sees the first tile inverse
        }
        const int nb = p1 - 1 - k;
        if (nb > 0) {
          if (split_leaf) { Launch L{}; L.is_leaf = OP_TRSM; L.blk = k; L.r0 = k + 1; L.nb = nb; seq.push_back(L); }
          else push_panel(k, k + 1, nb, 0);
          double* col = G + (long long)(k + 1) * T + (long long)k * T * ld;  // L_{k+1..p1-1, k}
          grp.push_back(mk(col, ld, col, ld, G + (long long)(k + 1) * T * (ld + 1), ld, nb * rr, nb * rr, tile, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
          add_group(seq, GK_NT, grp);
        }
        if (k > p0) {
          // row k of X_PP: T = L_{k,p0:k} X_{p0:k,p0:k} (X lower: k-range >= j), then X_{k,p0:k} = -X_kk T in place
          const int q = k - p0;
          double* rowk = X + (long long)k * T + (long long)p0 * T * ld;
          op(OP_WAIT, 6, ev_leaf);
          grp.push_back(mk(G + (long long)k * T + (long long)p0 * T * ld, ld, Xpp, ld, rowk, ld, rr, q * rr, q * tile, csg::KM_GE_J, csg::ORD_COL, 0, 1.0, 0.0));
          add_group(seq, GK_NN, grp); seq.back().stream = 6;
          if (ev_inv >= 0) op(OP_WAIT, 6, ev_inv);
          grp.push_back(mk(X + (long long)k * T * (ld + 1), ld, rowk, ld, rowk, ld, 1, q, tile, csg::KM_FULL, csg::ORD_COL, 0, -1.0, 0.0));
          add_group(seq, GK_NN, grp); seq.back().stream = 6; seq.back().gtile = tile;
        }
      }
      const int ev_diag = nev++, ev_xpp = nev++;
      op(OP_RECORD, 0, ev_diag);
      op(OP_WAIT, 6, ev_diag);
      op(OP_RECORD, 6, ev_xpp);
      last_s6_ev = ev_xpp;
      const int mt = N - p1;  // block rows below the panel
      const int n1 = w_next;
      int ev_lpanel = -1;
      if (mt > 0) {
        // ---- S2: L_IP = A_IP X_PP^T for the rows below (k <= j), next panel's rows first
        const long long kw = (long long)pw * tile;
        double* pan = G + (long long)p1 * T + (long long)p0 * T * ld;
        op(OP_WAIT, 2, ev_xpp);
        // the two multiplies the chain waits for (Ltop, U1top) use 32 x 32 CTA tiles when the plan tile is 128
        const int qf = (top32 && tile % 32 == 0) ? tile / 32 : rr;
        auto lmul = [&](int r0, int nb_rows, bool fine) {
          double* src = G + (long long)r0 * T + (long long)p0 * T * ld;
          double* dst = Pbuf + (long long)r0 * T;
          const int q = fine ? qf : rr;
          grp.push_back(mk(src, ld, Xpp, ld, dst, ld, nb_rows * q, pw * q, (int)kw, csg::KM_LE_J, csg::ORD_COL, 0, 1.0, 0.0));
          add_group(seq, GK_NT, grp); seq.back().stream = 2;
          if (fine && qf != rr) seq.back().gtile = 32;
          copy_op(2, src, ld, dst, ld, (long long)nb_rows * tile, kw);
        };
        lmul(p1, n1, true);
        // U1top: next diagonal block (lower tiles), K = panel width
        if (last_aux_ev >= 0) op(OP_WAIT, 2, last_aux_ev);  // U2 of the previous panel also wrote these columns
        grp.push_back(mk(pan, ld, pan, ld, G + (long long)p1 * T * (ld + 1), ld, n1 * qf, n1 * qf, (int)kw, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
        add_group(seq, GK_NT, grp); seq.back().stream = 2;
        if (qf != rr) seq.back().gtile = 32;
        ev_u1top_prev = nev++;
        op(OP_RECORD, 2, ev_u1top_prev);
        if (mt > n1) {
          lmul(p1 + n1, mt - n1, false);
          // U1rest: rectangle below the next diagonal block (next panel's columns)
          grp.push_back(mk(pan + (long long)n1 * T, ld, pan, ld, G + (long long)(p1 + n1) * T + (long long)p1 * T * ld, ld, (mt - n1) * rr, n1 * rr, (int)kw,
                           csg::KM_FULL, csg::ORD_COL, 0, -1.0, 1.0));
          add_group(seq, GK_NT, grp); seq.back().stream = 2;
        }
        ev_lpanel = nev++;
        op(OP_RECORD, 2, ev_lpanel);
        // ---- S1: U2, everything to the right of the next panel.  Look-ahead of depth two: the columns of
        // the panel after next (U2a) come first and are the only part U1 of the next panel waits for.
        if (mt > n1) {
          op(OP_WAIT, 1, ev_lpanel);
          const int m2 = mt - n1;                 // block rows / columns right of the next panel
          const int n2 = w_next2;                 // columns of the panel after next
          double* pan2 = pan + (long long)n1 * T; // L rows >= p1 + n1
          double* D2 = G + (long long)(p1 + n1) * T * (ld + 1);
          grp.push_back(mk(pan2, ld, pan2, ld, D2, ld, n2 * rr, n2 * rr, (int)kw, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
          if (m2 > n2)
            grp.push_back(mk(pan2 + (long long)n2 * T, ld, pan2, ld, D2 + (long long)n2 * T, ld, (m2 - n2) * rr, n2 * rr, (int)kw, csg::KM_FULL, csg::ORD_COL, 0, -1.0, 1.0));
          add_group(seq, GK_NT, grp); seq.back().stream = 1;
          last_aux_ev = nev++;
          op(OP_RECORD, 1, last_aux_ev);
          if (m2 > n2) {
            // U2b.  Look-ahead of depth three: the columns of the third panel ahead (U2b-first) stay on S1, the rest
            // (U2b-rest) goes to S7.  U2a of the NEXT panel touches exactly the tiles of U2b-first, so it no longer
            // queues behind the whole trailing update of this panel (which cost the chain ~190 us per boundary).
            const int m3 = m2 - n2;
            const int n3 = split_u2b ? w_next3 : 0;
            double* pan3 = pan2 + (long long)n2 * T;
            double* D3 = D2 + (long long)n2 * T * (ld + 1);
            if (last_u2brest_ev >= 0) op(OP_WAIT, 1, last_u2brest_ev);  // U2b-rest of the previous panel wrote these columns
            if (n3 > 0 && m3 > n3) {
              grp.push_back(mk(pan3, ld, pan3, ld, D3, ld, n3 * rr, n3 * rr, (int)kw, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
              grp.push_back(mk(pan3 + (long long)n3 * T, ld, pan3, ld, D3 + (long long)n3 * T, ld, (m3 - n3) * rr, n3 * rr, (int)kw, csg::KM_FULL, csg::ORD_COL, 0, -1.0, 1.0));
              add_group(seq, GK_NT, grp); seq.back().stream = 1;
              last_u2b_ev = nev++;
              op(OP_RECORD, 1, last_u2b_ev);
              double* pan4 = pan3 + (long long)n3 * T;
              op(OP_WAIT, 7, ev_lpanel);
              grp.push_back(mk(pan4, ld, pan4, ld, D3 + (long long)n3 * T * (ld + 1), ld, (m3 - n3) * rr, (m3 - n3) * rr, (int)kw, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
              add_group(seq, GK_NT, grp); seq.back().stream = 7;
              last_u2brest_ev = nev++;
              op(OP_RECORD, 7, last_u2brest_ev);
            } else {
              grp.push_back(mk(pan3, ld, pan3, ld, D3, ld, m3 * rr, m3 * rr, (int)kw, csg::KM_FULL, csg::ORD_LOWER, 0, -1.0, 1.0));
              add_group(seq, GK_NT, grp); seq.back().stream = 1;
              last_u2b_ev = nev++;
              op(OP_RECORD, 1, last_u2b_ev);
            }
          }
        }
      }
      if (with_inverse) {
        // ---- S3: rows P of X left of the diagonal block, X_iJ = -sum_{l<=i} X_il Acc_lJ, bottom row first (in place)
        op(OP_WAIT, 3, ev_xpp);
        if (p0 > 0) {
          // one launch for the whole row panel, out of place through Rbuf (k <= i: X_PP is lower triangular)
          const long long ldr = (long long)panel_blocks * tile;
          grp.push_back(mk(Xpp, ld, X + (long long)p0 * T, ld, Rbuf, ldr, pw * rr, p0 * rr, pw * tile, csg::KM_LE_I, csg::ORD_COL, 0, -1.0, 0.0));
          add_group(seq, GK_NN, grp); seq.back().stream = 3;
          copy_op(3, X + (long long)p0 * T, ld, Rbuf, ldr, (long long)pw * tile, (long long)p0 * tile);
        }
        const int ev_rows = nev++;
        op(OP_RECORD, 3, ev_rows);
        op(OP_WAIT, 4, ev_rows);
        {  // S4: C_IJ += sum_{k in P} X_kI^T X_kJ  (mirrored store keeps C symmetric)
          double* Xp = X + (long long)p0 * T;
          if (p0 > 0) {
            const int m1 = p0 * rr;
            const int tail_a = (int)pi >= tail_from ? tail_a_max : tail_a_max - 1;  // nondecreasing in the panel index
            if (tail_a > 0 && m1 > tail_a) {
              // The first tail_a tile columns of the leading square belong to the chain stream (partition A, idle once
              // the Cholesky is done): their accumulations are queued behind the last leaf (emitted after the panel loop).
              const long long o = (long long)tail_a * gt;
              TailPiece tp;
              tp.ev_rows = ev_rows; tp.ev_prev = last_inv_ev;
              tp.probs.push_back(mk(Xp, ld, Xp, ld, C, ld, tail_a, tail_a, pw * tile, csg::KM_FULL, csg::ORD_LOWER, 1, 1.0, 1.0));
              GemmProblem r = mk(Xp + o * ld, ld, Xp, ld, C + o, ld, m1 - tail_a, tail_a, pw * tile, csg::KM_FULL, csg::ORD_COL, 2, 1.0, 1.0);
              r.Cm = C + o * ld;
              tp.probs.push_back(r);
              tail_pieces.push_back(tp);
              grp.push_back(mk(Xp + o * ld, ld, Xp + o * ld, ld, C + o * (ld + 1), ld, m1 - tail_a, m1 - tail_a, pw * tile, csg::KM_FULL, csg::ORD_LOWER, 1, 1.0, 1.0));
            } else {
              grp.push_back(mk(Xp, ld, Xp, ld, C, ld, m1, m1, pw * tile, csg::KM_FULL, csg::ORD_LOWER, 1, 1.0, 1.0));
            }
            // rows P of K^-1 are touched for the first time here: beta = 0 (no memset of C per evaluation)
            GemmProblem q = mk(Xpp, ld, Xp, ld, C + (long long)p0 * T, ld, pw * rr, p0 * rr, pw * tile, csg::KM_GE_I, csg::ORD_COL, 2, 1.0, 0.0);
            q.Cm = C + (long long)p0 * T * ld;
            grp.push_back(q);
          }
          grp.push_back(mk(Xpp, ld, Xpp, ld, C + (long long)p0 * T * (ld + 1), ld, pw * rr, pw * rr, pw * tile, csg::KM_GE_I, csg::ORD_LOWER, 1, 1.0, 0.0));
          add_group(seq, GK_TN, grp); seq.back().stream = 4;
        }
        last_inv_ev = nev++;
        op(OP_RECORD, 4, last_inv_ev);
        if (mt > 0) {
          // Acc_IJ += sum_{k in P} L_Ik X_kJ for I >= p1: next panel's rows on S3 (AccNext), the rest on S5
          auto push_acc = [&](int r0, int nb_rows, int stream) {
            double* Lp = G + (long long)r0 * T + (long long)p0 * T * ld;
            if (p0 > 0)
              grp.push_back(mk(Lp, ld, X + (long long)p0 * T, ld, X + (long long)r0 * T, ld, nb_rows * rr, p0 * rr, pw * tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 1.0));
            // columns P of Acc are touched for the first time by panel P: beta = 0 (no memset of X per evaluation)
            grp.push_back(mk(Lp, ld, Xpp, ld, X + (long long)r0 * T + (long long)p0 * T * ld, ld, nb_rows * rr, pw * rr, pw * tile, csg::KM_GE_J, csg::ORD_COL, 0, 1.0, 0.0));
            add_group(seq, GK_NN, grp); seq.back().stream = stream;
          };
          op(OP_WAIT, 3, ev_lpanel);
          if (last_rest_ev >= 0) op(OP_WAIT, 3, last_rest_ev);
          push_acc(p1, n1, 3);
          last_inv3_ev = nev++;
          op(OP_RECORD, 3, last_inv3_ev);
          if (mt > n1) {
            op(OP_WAIT, 5, ev_rows);
            op(OP_WAIT, 5, ev_lpanel);
            push_acc(p1 + n1, mt - n1, 5);
            last_rest_ev = nev++;
            op(OP_RECORD, 5, last_rest_ev);
          }
        }
      }
    }
    // share of the K^-1 accumulation that runs on the chain stream once the chain is done, in panel order
    for (TailPiece& tp : tail_pieces) {
      op(OP_WAIT, 0, tp.ev_rows);
      if (tp.ev_prev >= 0) op(OP_WAIT, 0, tp.ev_prev);
      for (auto& q : tp.probs) grp.push_back(q);
      add_group(seq, GK_TN, grp);
    }
    tail_pieces.clear();
    // the chain stream rejoins every other stream before anything reads the results
    for (int ev : {last_aux_ev, last_u2b_ev, last_u2brest_ev, last_inv_ev, last_inv3_ev, last_rest_ev, last_s6_ev, ev_u1top_prev, last_s8_ev})
      if (ev >= 0) op(OP_WAIT, 0, ev);
    };
    // ---- left-looking plan for batched small-n handles: block column k is first brought up to date with ONE multiply whose
    // k-extent is everything to its left, C[k:N, k] -= L[k:N, 0:k] L[k, 0:k]^T (K = k*TILE), then the diagonal tile is
    // factored and inverted (leaf) and the column below it multiplied by X_kk^T.  The flops are those of the right-looking
    // plan, but they sit in N-1 launches of growing K instead of ~N^2/2 rank-TILE updates whose pipeline fill is as long as
    // their main loop (K = 64: 5.5 TFLOP/s effective).  One stream: the members of the batch (blockIdx.z) are the
    // parallelism.  3 launches per block step.
    auto build_chol_ll = [&](std::vector<Launch>& seq) {
      for (int k = 0; k < N; ++k) {
        if (k > 0) {
          double* rows = G + (long long)k * T;                               // block rows k.., column 0
          grp.push_back(mk(rows, ld, rows, ld, G + (long long)k * T * (ld + 1), ld, (N - k) * rr, rr, k * tile, csg::KM_FULL, csg::ORD_COL, 0, -1.0, 1.0));
          add_group(seq, GK_NT, grp);
        }
        { Launch L{}; L.is_leaf = OP_LEAF; L.blk = k; seq.push_back(L); }
        const int nb = N - 1 - k;
        if (nb > 0) {
          double* pan = G + (long long)(k + 1) * T + (long long)k * T * ld;
          double* Xkk = X + (long long)k * T * (ld + 1);
          if (rr == 2) {  // see push_panel: right half first, then the left half with K = 64 (in-place safe per CTA)
            grp.push_back(mk(pan, ld, Xkk + gt, ld, pan + gt * ld, ld, nb * rr, 1, tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
            add_group(seq, GK_NT, grp);
            grp.push_back(mk(pan, ld, Xkk, ld, pan, ld, nb * rr, 1, gt, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
            add_group(seq, GK_NT, grp);
          } else {
            grp.push_back(mk(pan, ld, Xkk, ld, pan, ld, nb, 1, tile, csg::KM_FULL, csg::ORD_COL, 0, 1.0, 0.0));
            add_group(seq, GK_NT, grp); seq.back().gtile = tile;
          }
        }
      }
    };
    plan_version = 2;
    if (const char* pv = std::getenv("CS_PLAN")) plan_version = std::atoi(pv);
    left_looking = nbatch > 1 && !inv_pipeline && N > 1 && std::getenv("CS_PLAN") == nullptr && std::getenv("CS_BATCH_RIGHT") == nullptr;
    // Split leaf: below np = 8192 a single evaluation is bound by the chain of leaves, so the inversion of the diagonal
    // tiles leaves the chain (S8) and the panel below a tile comes from a triangular solve.  From np = 8192 on the run is bound
    // by the bulk streams and the chain partition's idle time is sold to the tail share, which is sized for the fused leaf
    // (measured at n = 8192: split 20.3 ms vs fused 19.5 ms; n = 4096: 4.5 vs 5.2 ms).  CS_SPLIT_LEAF=1 / CS_NO_SPLIT_LEAF=1 force it.
    split_leaf = nbatch == 1 && N > 1 && !std::getenv("CS_NO_SPLIT_LEAF") && (np < 8192 || std::getenv("CS_SPLIT_LEAF") != nullptr);
    top32 = tile == 128 && gt == 64 && nbatch == 1 && !std::getenv("CS_NO_TOP32");
    split_u2b = nbatch == 1 && !std::getenv("CS_NO_U2B_SPLIT");
    {
      // Panel widths: uniform W by default.  CS_GRADED=1 selects a narrow first panel (the bulk streams start after
      // two leaves instead of four) and narrow last panels (a shorter tail); measured on the B200 at n = 8192 it is
      // 1 % SLOWER (19.35 vs 19.13 ms): the run is bound by the throughput of the K = 512 bulk launches, not by
      // the head or the tail, and narrower panels lower it.  Kept as an option (2 = also for small N, tests;
      // 3 = narrow head only, panels 1, 1, 2, W, W, ...: 20.03 vs 20.02 ms, no gain either).
      bounds.assign(1, 0);
      bool graded = false;
      if (const char* gv = std::getenv("CS_GRADED")) {
        const int g = std::atoi(gv);
        graded = W == 4 && (g == 2 ? N >= 9 : (g == 1 && N >= 6 * W));
      }
      int head_only = 0;
      if (const char* gv = std::getenv("CS_GRADED")) head_only = (std::atoi(gv) == 3 && W == 4 && N >= 4 * W && N % W == 0);
      if (head_only) {  // 1, 1, 2, then W: the bulk streams start after a single leaf
        bounds.push_back(1); bounds.push_back(2);
        for (int pos = W; pos <= N; pos += W) bounds.push_back(pos);
      } else if (graded) {
        const int tail[4] = {2, 2, 1, 1};
        int pos = 2;
        bounds.push_back(pos);
        const int rem = (N - 2 - 6) % W;
        if (rem > 0) { pos += rem; bounds.push_back(pos); }
        while (pos < N - 6) { pos += W; bounds.push_back(pos); }
        for (int t : tail) { pos += t; bounds.push_back(pos); }
      } else {
        for (int pos = W; pos < N; pos += W) bounds.push_back(pos);
        bounds.push_back(N);
      }
    }
    {
      // Tail share of the chain partition (plan v2, single fits): once the chain of N leaves is done (about 186 us per
      // leaf step on the B200) its 8 SMs are idle until the bulk streams finish (about np^3 flops at 28 TFLOP/s), and
      // in that window they accumulate about 1.6 TFLOP/s of K^-1 tiles.  One tile column costs
      // 2 * gt * tile^2 * sum_P p0 * pw flops over the run.  n = 8192: 3 columns (measured: 2 -> 19.72, 3 -> 19.60,
      // 4 -> 21.2 ms, the chain partition becomes the last to finish); n <= 6144: chain-bound, no share.
      tail_a_max = 0;
      tail_from = ((int)bounds.size() - 1) / 2;
      if (nbatch == 1 && np >= 4096) {
        double sum = 0.0;
        for (size_t pi = 0; pi + 1 < bounds.size(); ++pi) sum += (double)bounds[pi] * (bounds[pi + 1] - bounds[pi]);
        // the three constants were measured on a 148-SM B200 at 1965 MHz (bulk 28 TFLOP/s on 140 SMs, chain partition
        // 1.6 TFLOP/s on 8 SMs, 186 us per leaf step); the bulk rate scales with the SM count of the device in use.
        // CS_TAIL_A overrides the model.
        const double bulk = 28.0e12 * std::max(rt::sm_count() - 8, 8) / 140.0;
        const double window = (double)np * np * np / bulk - (double)N * 186e-6;
        const double per_col = 2.0 * gt * (double)tile * tile * sum;
        // Round 2: the window model over-assigns at large n (measured on the B200, best a: n = 8192: 3 of 128 tile
        // columns; 12288: 3..5 of 192; 16384: 6..8 of 256, 10 already slower; 32768: 16 of 512, 24 slower -- the model said
        // 12 at 16384 and 64 at 32768 and cost 10 % / 19 % there): the pieces of the late panels carry most of the work and
        // pile up behind the chain.  So the share is capped at 1/36 of the tile columns.
        if (window > 0.0 && per_col > 0.0)
          tail_a_max = (int)std::min(std::min(window * 1.6e12 / per_col, (double)(N * rr / 8)), (double)(N * rr / 36));  // in double: no int overflow
      }
      if (const char* ta = std::getenv("CS_TAIL_A")) {
        int a = 0, f = 0;
        const int got = std::sscanf(ta, "%d,%d", &a, &f);
        tail_a_max = nbatch == 1 ? a : 0;
        tail_from = got > 1 ? f : 0;
      }
    }
    if (left_looking) {
      build_chol_ll(potrf);
    } else if (plan_version == 1) {
      build_chol(potrf, false);
      if (inv_pipeline) build_chol(potrf_inv, true);
    } else {
      build_chol2(potrf, false);
      if (inv_pipeline) build_chol2(potrf_inv, true);
    }
    n_events = nev + 1;
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
        g1.push_back(mk(G + nd.mid * T + nd.lo * T * ld, ld, X + nd.lo * T * (ld + 1), ld, Tw, ld, m * rr, q * rr, q * tile,
                        csg::KM_GE_J, csg::ORD_COL, 0, 1.0, 0.0));
        // X21 = -X22 * T     (NN; X22 lower => k <= i)
        g2.push_back(mk(X + nd.mid * T * (ld + 1), ld, Tw, ld, X + nd.mid * T + nd.lo * T * ld, ld, m * rr, q * rr, m * tile,
                        csg::KM_LE_I, csg::ORD_ROW_DESC, 0, -1.0, 0.0));
      }
      add_group(trtri, GK_NN, g1);
      add_group(trtri, GK_NN, g2);
    }
    // --- K^-1 = X^T X  (TN; k >= max(i,j) = i on lower tiles; mirrored store) ---
    grp.push_back(mk(X, ld, X, ld, C, ld, N * rr, N * rr, np, csg::KM_GE_I, csg::ORD_LOWER, 1, 1.0, 0.0));
    add_group(xtx, GK_TN, grp);

    // partition: 0 = auto (large problems only: the chain of a small fit is launch-bound, not slot-bound,
    // and many small fits share the device), 1 = on, -1 = off
    // Nsight Compute cannot attach to kernels launched into green-context streams ("Failed to prepare kernel for
    // profiling", the application dies): under its injection environment the auto mode keeps plain priority streams
    const bool under_ncu = std::getenv("NV_NSIGHT_INJECTION_PORT_BASE") != nullptr;
    const bool want_part = partition > 0 || (partition == 0 && np >= 4096 && !under_ncu && std::getenv("CS_NO_PARTITION") == nullptr);
    // priorities: S2 (panel work the Cholesky chain waits for) highest, then S3 (row chain of the inverse),
    // S1 (trailing update), then S5 (Acc of the rows below the next panel) and S4 (K^-1 accumulation) lowest
    cudaStream_t* bulk[6] = {&aux, &aux2, &aux3, &aux4, &aux5, &aux7};
    int level[6] = {2, 0, 1, 5, 5, 3};  // 0 = highest (swept on the B200: tools/gpu_sweep.sh); last: S7
    if (const char* ev = std::getenv("CS_PRIO")) std::sscanf(ev, "%d,%d,%d,%d,%d,%d", &level[0], &level[1], &level[2], &level[3], &level[4], &level[5]);
    int crit_sms = 8;   // swept on the B200 (CS_CRIT_SMS): see profiles/r02_split_leaf.md
    if (const char* ev = std::getenv("CS_CRIT_SMS")) crit_sms = std::max(4, std::min(64, std::atoi(ev)));
    if (want_part && rt::partition_streams(crit_sms, &crit, &aux6, &aux8, bulk, level, 6, &sm_crit, &sm_bulk) == 0) {
      crit_ok = aux_ok = aux2_ok = aux3_ok = aux4_ok = aux5_ok = aux6_ok = aux7_ok = aux8_ok = true;
      if ((e = rt::event_create_fast(&ev_in))) return e;
      if ((e = rt::event_create_fast(&ev_out))) return e;
    } else {
      if ((e = rt::stream_create_level(&aux, level[0]))) return e;
      aux_ok = true;
      if ((e = rt::stream_create_level(&aux2, level[1]))) return e;
      aux2_ok = true;
      if ((e = rt::stream_create_level(&aux3, level[2]))) return e;
      aux3_ok = true;
      if ((e = rt::stream_create_level(&aux4, level[3]))) return e;
      aux4_ok = true;
      if ((e = rt::stream_create_level(&aux5, level[4]))) return e;
      aux5_ok = true;
      if ((e = rt::stream_create_level(&aux6, 0))) return e;
      aux6_ok = true;
      if ((e = rt::stream_create_level(&aux7, level[5]))) return e;
      aux7_ok = true;
      if ((e = rt::stream_create_level(&aux8, 0))) return e;
      aux8_ok = true;
    }
    if ((e = rt::event_create_fast(&ev_gram))) return e;
    ev_gram_ok = true;
    events.resize(n_events);
    for (int i = 0; i < n_events; ++i)
      if ((e = rt::event_create_fast(&events[i]))) return e;
    if ((e = rt::dev_malloc((void**)&d_probs, sizeof(GemmProblem) * probs.size()))) return e;
    if ((e = rt::h2d(d_probs, probs.data(), sizeof(GemmProblem) * probs.size(), 0))) return e;
    if ((e = rt::stream_sync(0))) return e;
    return 0;
  }

  void destroy() {
    if (!external_buffers) {
      rt::dev_free(G); rt::dev_free(X); rt::dev_free(C); rt::dev_free(logdet_part); rt::dev_free(status); rt::dev_free(Pbuf);
      rt::dev_free(Rbuf);
    }
    Pbuf = nullptr; Rbuf = nullptr;
    rt::dev_free(d_probs);
    for (auto& ev : events) rt::event_destroy(ev);
    events.clear();
    if (ev_gram_ok) { rt::event_destroy(ev_gram); ev_gram_ok = false; }
    if (aux_ok) { rt::stream_sync(aux); rt::stream_destroy(aux); aux_ok = false; }
    if (aux2_ok) { rt::stream_sync(aux2); rt::stream_destroy(aux2); aux2_ok = false; }
    if (aux3_ok) { rt::stream_sync(aux3); rt::stream_destroy(aux3); aux3_ok = false; }
    if (aux4_ok) { rt::stream_sync(aux4); rt::stream_destroy(aux4); aux4_ok = false; }
    if (aux5_ok) { rt::stream_sync(aux5); rt::stream_destroy(aux5); aux5_ok = false; }
    if (aux6_ok) { rt::stream_sync(aux6); rt::stream_destroy(aux6); aux6_ok = false; }
    if (aux7_ok) { rt::stream_sync(aux7); rt::stream_destroy(aux7); aux7_ok = false; }
    if (aux8_ok) { rt::stream_sync(aux8); rt::stream_destroy(aux8); aux8_ok = false; }
    if (crit_ok) { rt::stream_sync(crit); rt::stream_destroy(crit); rt::event_destroy(ev_in); rt::event_destroy(ev_out); crit_ok = false; }
    G = X = C = logdet_part = nullptr; status = nullptr; d_probs = nullptr;
  }

  // timeline capture (cs_debug_timeline): when set, every kernel launch of run_seq is bracketed
  // by a pair of timing events on its own stream
  std::vector<cudaEvent_t>* tl = nullptr;

  // use_crit: plan stream 0 is the critical chain of the Cholesky and runs on the partition-A stream
  // (entered and left through events on `st`); plans whose stream 0 carries bulk work pass false
  void run_seq(const std::vector<Launch>& seq, cudaStream_t st, const int* skip = nullptr, bool use_crit = false) {
    const bool part = use_crit && crit_ok;
    cudaStream_t s0 = st;
    if (part) {
      rt::event_record(ev_in, st);
      rt::stream_wait_event(crit, ev_in);
      s0 = crit;
    }
    for (const Launch& L : seq) {
      cudaStream_t s = L.stream == 1 ? aux : (L.stream == 2 ? aux2 : (L.stream == 3 ? aux3 : (L.stream == 4 ? aux4 : (L.stream == 5 ? aux5 : (L.stream == 6 ? aux6 : (L.stream == 7 ? aux7 : (L.stream == 8 ? aux8 : s0)))))));
      const bool timed = tl && (L.is_leaf == OP_LEAF || L.is_leaf == OP_GEMM || L.is_leaf == OP_COPY || L.is_leaf == OP_TRSM);
      if (timed) { cudaEvent_t e0; rt::event_create(&e0); rt::event_record(e0, s); tl->push_back(e0); }
      if (L.is_leaf == OP_LEAF) {
        launch_leaf(tile, G + (long long)L.blk * tile * ((long long)np + 1), np, X + (long long)L.blk * tile * ((long long)np + 1), np,
                    L.blk * tile, logdet_part + L.blk, status, n, s, skip, nbatch, bstride, L.leaf_mode);
      } else if (L.is_leaf == OP_TRSM) {
        launch_trsm(tile, G + (long long)L.blk * tile * ((long long)np + 1), np, G + (long long)L.r0 * tile + (long long)L.blk * tile * np, np,
                    L.nb * tile, s, skip, nbatch, bstride);
      } else if (L.is_leaf == OP_COPY) {
        launch_copy(L, s, skip);
      } else if (L.is_leaf == OP_RECORD) {
        rt::event_record(events[L.ev], s);
      } else if (L.is_leaf == OP_WAIT) {
        rt::stream_wait_event(s, events[L.ev]);
      } else {
        launch_gemm(L.gtile ? L.gtile : gt, L.kind, d_probs + L.first_prob, L.nprob, L.ntiles, s, skip, nbatch);
      }
      if (timed) { cudaEvent_t e1; rt::event_create(&e1); rt::event_record(e1, s); tl->push_back(e1); }
    }
    if (part) {
      rt::event_record(ev_out, crit);
      rt::stream_wait_event(st, ev_out);
    }
  }

  // ---- plan race detector (host only) ---------------------------------------------------
  // The Cholesky plan spreads its launches over three streams ordered by events.  This
  // check rebuilds the happens-before relation of the op list (stream order + record/wait
  // edges) and verifies that every pair of launches touching overlapping regions of G / X / C
  // with at least one write is ordered.  Returns the number of unordered conflicting pairs.
  struct Rect { int buf; long long r0, r1, c0, c1; bool write; };
  void rect_of(const double* ptr, long long rows, long long cols, bool write, std::vector<Rect>& out) const {
    const long long pb = (long long)np * panel_blocks * tile;
    const double* bases[5] = {G, X, C, Pbuf, Rbuf};
    const long long sizes[5] = {(long long)np * np, (long long)np * np, (long long)np * np, pb, pb};
    const long long lds[5] = {np, np, np, np, (long long)panel_blocks * tile};
    for (int b = 0; b < 5; ++b) {
      if (!bases[b]) continue;
      const long long off = ptr - bases[b];
      if (off >= 0 && off < sizes[b]) {
        out.push_back(Rect{b, off % lds[b], off % lds[b] + rows, off / lds[b], off / lds[b] + cols, write});
        return;
      }
    }
  }
  void footprint(const Launch& L, std::vector<Rect>& out) const {
    out.clear();
    if (L.is_leaf == OP_LEAF) {
      const long long o = (long long)L.blk * tile;
      if (L.leaf_mode != 2) out.push_back(Rect{0, o, o + tile, o, o + tile, true});
      else out.push_back(Rect{0, o, o + tile, o, o + tile, false});
      if (L.leaf_mode != 1) out.push_back(Rect{1, o, o + tile, o, o + tile, true});
      return;
    }
    if (L.is_leaf == OP_TRSM) {
      const long long o = (long long)L.blk * tile, r0 = (long long)L.r0 * tile;
      out.push_back(Rect{0, o, o + tile, o, o + tile, false});
      out.push_back(Rect{0, r0, r0 + (long long)L.nb * tile, o, o + tile, true});
      return;
    }
    if (L.is_leaf == OP_COPY) {
      rect_of(L.cp_src, L.cp_rows, L.cp_cols, false, out);
      rect_of(L.cp_dst, L.cp_rows, L.cp_cols, true, out);
      return;
    }
    if (L.is_leaf != OP_GEMM) return;
    const int t = L.gtile ? L.gtile : gt;
    for (int i = 0; i < L.nprob; ++i) {
      const GemmProblem& P = probs[L.first_prob + i];
      const long long M = (long long)P.mt * t, Nn = (long long)P.nt * t, K = P.K;
      const bool ak = (L.kind == GK_TN), bk = (L.kind != GK_NT);
      if (ak) rect_of(P.A, K, M, false, out); else rect_of(P.A, M, K, false, out);
      if (bk) rect_of(P.B, K, Nn, false, out); else rect_of(P.B, Nn, K, false, out);
      rect_of(P.C, M, Nn, true, out);
      if (P.mirror) rect_of(P.Cm ? P.Cm : P.C, Nn, M, true, out);
    }
  }
  int check_plan(const std::vector<Launch>& seq, int verbose) const {
    const int n_ops = (int)seq.size();
    const int words = (n_ops + 63) / 64;
    std::vector<std::vector<unsigned long long>> anc(n_ops, std::vector<unsigned long long>(words, 0ULL));
    int last_on_stream[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
    std::vector<int> last_record(n_events + 1, -1);
    auto add_pred = [&](int j, int i) {
      if (i < 0) return;
      for (int w = 0; w < words; ++w) anc[j][w] |= anc[i][w];
      anc[j][i / 64] |= 1ULL << (i % 64);
    };
    for (int j = 0; j < n_ops; ++j) {
      const Launch& L = seq[j];
      add_pred(j, last_on_stream[L.stream]);
      if (L.is_leaf == OP_WAIT) add_pred(j, last_record[L.ev]);
      if (L.is_leaf == OP_RECORD) last_record[L.ev] = j;
      last_on_stream[L.stream] = j;
    }
    std::vector<std::vector<Rect>> fp(n_ops);
    for (int j = 0; j < n_ops; ++j) footprint(seq[j], fp[j]);
    int bad = 0;
    for (int j = 0; j < n_ops; ++j)
      for (int i = 0; i < j; ++i) {
        if (fp[i].empty() || fp[j].empty()) continue;
        if (anc[j][i / 64] >> (i % 64) & 1ULL) continue;  // ordered
        bool conflict = false;
        for (const Rect& a : fp[i]) {
          for (const Rect& b : fp[j])
            if (a.buf == b.buf && (a.write || b.write) && a.r0 < b.r1 && b.r0 < a.r1 && a.c0 < b.c1 && b.c0 < a.c1) { conflict = true; break; }
          if (conflict) break;
        }
        if (conflict) {
          if (verbose && bad < 10)
            std::fprintf(stderr, "plan race: op %d (stream %d, type %d, blk %d) vs op %d (stream %d, type %d, blk %d)\n", i,
                         seq[i].stream, seq[i].is_leaf, seq[i].blk, j, seq[j].stream, seq[j].is_leaf, seq[j].blk);
          ++bad;
        }
      }
    return bad;
  }

  // G (lower) must hold K_n; X and the upper triangle of G must be zero outside diagonal tiles.
  void factor(cudaStream_t st, const int* skip = nullptr) { set_status_max(st, skip); run_seq(potrf, st, skip, true); }
  void invert(cudaStream_t st, const int* skip = nullptr) { if (!single_panel) run_seq(trtri, st, skip); run_seq(xtx, st, skip); }
  // K_n (in G) -> L, L^-1, K_n^-1.  With inv_pipeline the inverse trails the factorization on its
  // own stream; X and C are accumulated into and must start from zero.
  void factor_and_invert(cudaStream_t st, const int* skip = nullptr) {
    if (inv_pipeline) {
      // plan v2 initialises every tile of X and C with its first contribution (beta = 0); the upper off-diagonal
      // tiles of X are zeroed once at creation and never written
      const size_t mb = sizeof(double) * (size_t)np * np;
      for (int b = 0; b < (plan_version == 2 ? 0 : nbatch); ++b) {
        rt::memset0(reinterpret_cast<char*>(X) + b * bstride, mb, st);
        rt::memset0(reinterpret_cast<char*>(C) + b * bstride, mb, st);
      }
      set_status_max(st, skip);
      run_seq(potrf_inv, st, skip, true);
    } else {
      factor(st, skip);
      invert(st, skip);
    }
  }

  void set_status_max(cudaStream_t st, const int* skip);
  void launch_copy(const Launch& L, cudaStream_t s, const int* skip);
};

// 2-D copy of a rows x cols block of doubles, one launch for every member of a batched handle (blockIdx.z)
static __global__ void __launch_bounds__(256)
copy2d_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src, long long lds, long long rows,
              long long cols, const int* __restrict__ skip, long long bstride) {
  skip = csd::bshift(skip, bstride);
  if (skip && *skip) return;
  dst = csd::bshift(dst, bstride); src = csd::bshift(src, bstride);
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (r >= rows) return;
  for (long long c = blockIdx.y; c < cols; c += gridDim.y) {
    const double* s = src + c * lds + r;
    double* d = dst + c * ldd + r;
    const double v0 = s[0], v1 = (r + 1 < rows) ? s[1] : 0.0;
    d[0] = v0;
    if (r + 1 < rows) d[1] = v1;
  }
}

static __global__ void set_int_kernel(int* p, int v, const int* skip, long long bstride) {
  skip = csd::bshift(skip, bstride);
  if (skip && *skip) return;
  p = csd::bshift(p, bstride);
  if (threadIdx.x == 0 && blockIdx.x == 0) *p = v;
}
inline void DenseEngine::launch_copy(const Launch& L, cudaStream_t s, const int* skip) {
  // rows, leading dimensions and offsets are even (multiples of the tile) and the buffers 256-byte aligned:
  // 16-byte accesses are aligned
  auto k = copy2d_kernel;
  ++g_launches;
  const unsigned gx = (unsigned)((L.cp_rows / 2 + 255) / 256);
  const unsigned gy = (unsigned)std::min<long long>(L.cp_cols, 4096);
  CS_LAUNCH(k, dim3(gx, gy, nbatch), 256, 0, s, L.cp_dst, L.cp_ldd, L.cp_src, L.cp_lds, L.cp_rows, L.cp_cols, skip, bstride);
}

inline void DenseEngine::set_status_max(cudaStream_t st, const int* skip) {
  if (!reset_status) return;  // inside cs_fit_run the failing pivot is sticky (reset once per run)
  auto k = set_int_kernel;
  ++g_launches;
  CS_LAUNCH(k, dim3(1, 1, nbatch), 32, 0, st, status, 0x7fffffff, skip, bstride);
}

}  // namespace csdense
