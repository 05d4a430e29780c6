// Structured FP64 tile GEMM on the DMMA.8x8x4 tensor path (sm_100a), used for every
// dense contraction of the fit: Cholesky trailing updates, the panel multiply by the
// inverted diagonal block, the recursive triangular inverse, K^-1 = X^T X, and the
// predict / treatment cross products.  One launch executes a *group* of problems; each
// CTA owns one TILE x TILE output tile and walks its own k-range (triangular operands
// are skipped at tile granularity; the opposite triangle is stored as explicit zeros so
// correctness never depends on the skipping).
//
// Operand staging: cp.async (LDGSTS) 16-byte copies into a NSTAGE-deep padded
// shared-memory ring; pads (+4 doubles) make every fragment LDS.64 conflict-free.
// All matrices are column-major fp64 (the R / Armadillo layout of the reference).
#pragma once
#include "cs_dev.cuh"

namespace csg {

enum KMode { KM_FULL = 0, KM_GE_J = 1, KM_LE_I = 2, KM_GE_I = 3, KM_LE_J = 4 };
enum TileOrder { ORD_COL = 0, ORD_ROW_DESC = 1, ORD_LOWER = 2 };

struct GemmProblem {
  const double* A;  // element (i,k): AK ? A[k + i*lda] : A[i + k*lda]
  const double* B;  // element (j,k): BK ? B[k + j*ldb] : B[j + k*ldb]
  double* C;        // element (i,j): C[i + j*ldc]
  long long lda, ldb, ldc;
  int mt, nt;       // tiles along m / n
  int K;            // k extent in elements (multiple of GEMM_BK)
  int kmode;        // KMode, in units of TILE
  int order;        // TileOrder; ORD_LOWER enumerates only I >= J (mt == nt)
  int mirror;       // 1: also store C(j,i) = C(i,j) for I != J (square, diagonal origin); 2: always (rectangular)
  int tile_base;    // first linear tile id of this problem inside the launch
  int ntiles;
  double alpha, beta;  // C = beta*C + alpha*acc   (alpha in {+1,-1}, beta in {0,1}: every contraction of the fit)
  // What the kernel reads: sign masks / flags derived from (alpha, beta) by set_scale().  The prologue and epilogue of a
  // tile must not touch the FP64 pipe -- it is saturated by the DMMAs of the co-resident CTAs, and every dependent
  // DMUL / DFMA / fp64 divide issued there queues behind them (ncu, K = 512 launches: 14 % of all warp time sat in a
  // `beta / alpha` divide and an fp64 sqrt of the tile decode).  +-1 scaling is an integer XOR on the sign bit.
  unsigned neg_mask;   // 0x80000000 if alpha < 0
  int beta_one;        // 1: accumulate into C
  __host__ __device__ void set_scale(double a, double b) { alpha = a; beta = b; neg_mask = a < 0.0 ? 0x80000000u : 0u; beta_one = b != 0.0 ? 1 : 0; }
  // Table mode (distributed block-cyclic sweeps): tile t uses element offsets
  // table[3t] into A, table[3t+1] into B, table[3t+2] into C and the full k-range.
  const long long* table;
  double* Cm;       // base of the transposed block for mirror == 2 (nullptr: same as C)
  long long bstride; // batched launch: byte offset between consecutive fits (blockIdx.z), 0 = single problem set
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_PAD = 4;
constexpr int GEMM_STAGES = 4;

__host__ __device__ inline int lower_tiles(int n) { return n * (n + 1) / 2; }

// row i of the lower-triangular enumeration t = i(i+1)/2 + j: single-precision estimate, exact integer correction
// (no fp64 arithmetic: see GemmProblem::neg_mask)
__host__ __device__ inline int lower_row(int t) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((long long)(i + 1) * (i + 2) / 2 <= t) ++i;
  while ((long long)i * (i + 1) / 2 > t) --i;
  return i;
}

__host__ __device__ inline void decode_tile(int order, int mt, int nt, int t, int& I, int& J) {
  if (order == ORD_COL) {
    J = t / mt; I = t - J * mt;
  } else if (order == ORD_ROW_DESC) {
    int r = t / nt; I = mt - 1 - r; J = t - r * nt;
  } else {
    const int i = lower_row(t);
    I = i; J = t - i * (i + 1) / 2;
  }
}

// x or -x without the FP64 pipe
__device__ __forceinline__ double flip_sign(double x, unsigned mask) {
#ifdef CS_EMU
  return mask ? -x : x;
#else
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
#endif
}

__host__ __device__ inline void tile_krange(int kmode, int K, int TILE, int I, int J, int& kbeg, int& kend) {
  kbeg = 0; kend = K;
  if (kmode == KM_GE_J) kbeg = J * TILE;
  else if (kmode == KM_GE_I) kbeg = I * TILE;
  else if (kmode == KM_LE_I) kend = (I + 1) * TILE;
  else if (kmode == KM_LE_J) kend = (J + 1) * TILE;
  if (kend > K) kend = K;
  if (kbeg > kend) kbeg = kend;
}

template <int TILE, bool KMAJ, int KC = GEMM_BK>
struct OperandSmem {
  // m-major: [KC][TILE+PAD]   k-major: [TILE][KC+PAD]
  static constexpr int LD = KMAJ ? (KC + GEMM_PAD) : (TILE + GEMM_PAD);
  static constexpr int SIZE = KMAJ ? TILE * LD : KC * LD;
};

template <int TILE, int WM, int WN>
struct GemmCfg {
  static constexpr int THREADS = WM * WN * 32;
  static constexpr int WTM = TILE / WM, WTN = TILE / WN;
  static constexpr int FM = WTM / 8, FN = WTN / 8;
};

template <int TILE, bool AK, bool BK_, int KC = GEMM_BK, int NS = GEMM_STAGES>
constexpr int gemm_smem_bytes() {
  return NS * (OperandSmem<TILE, AK, KC>::SIZE + OperandSmem<TILE, BK_, KC>::SIZE) * (int)sizeof(double);
}

// copy one TILE x KC operand chunk into a stage buffer; g points at the chunk's first element (tile row 0, first k)
template <int TILE, bool KMAJ, int THREADS, int KC>
__device__ __forceinline__ void load_operand(double* __restrict__ s, const double* __restrict__ g, long long ld, int tid) {
  using S = OperandSmem<TILE, KMAJ, KC>;
  constexpr int NCOPY = TILE * KC / 2;
  static_assert(NCOPY % THREADS == 0, "copy count must divide");
#pragma unroll
  for (int it = 0; it < NCOPY / THREADS; ++it) {
    const int idx = tid + it * THREADS;
    if (KMAJ) {
      const int i = idx / (KC / 2), kc = idx - i * (KC / 2);
      csd::cp_async16(s + i * S::LD + 2 * kc, g + (long long)i * ld + 2 * kc);
    } else {
      const int kk = idx / (TILE / 2), ic = idx - kk * (TILE / 2);
      csd::cp_async16(s + kk * S::LD + 2 * ic, g + (long long)kk * ld + 2 * ic);
    }
  }
}

// The same copies with the per-thread part of the addresses computed ONCE per tile: thread tid moves the 16-byte pieces
// idx = tid + it * THREADS, whose row (m-major: k row, k-major: tile row) advances by a constant ROWS per `it` while the position
// inside the row stays, so piece `it` sits at a fixed stride behind piece 0 in global and in shared memory.  (Recomputing
// row, column and the 64-bit product per copy was 60 of the ~250 instructions a thread issues per chunk: ncu source page of a
// K = 512 launch, profiles/r02_ncu_nt512.md.)
template <int TILE, bool KMAJ, int THREADS, int KC>
struct OperandCopy {
  using S = OperandSmem<TILE, KMAJ, KC>;
  static constexpr int PER_ROW = KMAJ ? KC / 2 : TILE / 2;      // 16-byte pieces per row
  static constexpr int NCOPY = TILE * KC / 2 / THREADS;         // pieces per thread and chunk
  static constexpr int ROWS = THREADS / PER_ROW;                // rows between consecutive pieces of a thread
  static_assert(THREADS % PER_ROW == 0 && (TILE * KC / 2) % THREADS == 0, "copy layout must divide");
  __device__ __forceinline__ static int soff0(int tid) { return (tid / PER_ROW) * S::LD + 2 * (tid % PER_ROW); }
  __device__ __forceinline__ static long long goff0(int tid, long long ld) { return (long long)(tid / PER_ROW) * ld + 2 * (tid % PER_ROW); }
  // s, g: stage buffer / chunk origin already advanced by soff0 / goff0; gstep = ROWS * ld
  __device__ __forceinline__ static void copy(double* __restrict__ s, const double* __restrict__ g, long long gstep) {
#pragma unroll
    for (int it = 0; it < NCOPY; ++it) csd::cp_async16(s + it * ROWS * S::LD, g + it * gstep);
  }
};

// What one CTA needs to know about its tile.  Decoded by ONE thread (descriptor search, tile decode, k-range, operand
// bases) and handed to the others through shared memory: 128 threads each walking the descriptor list in global memory
// and decoding the same tile cost 3 % at K = 512 and 12-18 % at K = 64-128 (tools/gemm_bench, profiles/r02_gemm_experiments.md §6).
struct TileCtx {
  const double* A;   // operand bases already at (row0, kbeg)
  const double* B;
  double* C;         // at (i0, j0)
  double* Cm;        // mirrored tile base at (j0, i0), or nullptr
  long long lda, ldb, ldc;
  int nchunks;
  unsigned neg;
  int beta_one;
  int pad;
};

template <int TILE, bool AK, bool BK_, int KC>
__device__ inline void tile_setup(const GemmProblem* __restrict__ probs, int nprob, int bt, int& pi, TileCtx& c) {
  while (pi + 1 < nprob && bt >= probs[pi + 1].tile_base) ++pi;
  const GemmProblem& P = probs[pi];
  const double* A = csd::bshift(P.A, P.bstride);
  const double* B = csd::bshift(P.B, P.bstride);
  double* C = csd::bshift(P.C, P.bstride);
  double* Cm = csd::bshift(P.Cm, P.bstride);
  int I = 0, J = 0, kbeg = 0, kend = P.K;
  if (P.table) {
    const long long* te = P.table + 3LL * (bt - P.tile_base);
    A += te[0]; B += te[1]; C += te[2];
  } else {
    decode_tile(P.order, P.mt, P.nt, bt - P.tile_base, I, J);
    tile_krange(P.kmode, P.K, TILE, I, J, kbeg, kend);
  }
  const long long i0 = (long long)I * TILE, j0 = (long long)J * TILE;
  c.lda = P.lda; c.ldb = P.ldb; c.ldc = P.ldc;
  c.A = AK ? A + i0 * P.lda + kbeg : A + (long long)kbeg * P.lda + i0;
  c.B = BK_ ? B + j0 * P.ldb + kbeg : B + (long long)kbeg * P.ldb + j0;
  c.C = C + i0 + j0 * P.ldc;
  c.Cm = (P.mirror == 2 || (P.mirror == 1 && I != J)) ? (Cm ? Cm : C) + j0 + i0 * P.ldc : nullptr;
  c.nchunks = (kend - kbeg) / KC;
  c.neg = P.neg_mask;
  c.beta_one = P.beta_one;
  c.pad = 0;
}

template <int TILE, int WM, int WN, bool AK, bool BK_, int KC = GEMM_BK, int NS = GEMM_STAGES, int MINB = 1>
__global__ void __launch_bounds__(WM * WN * 32, MINB)
gemm_kernel(const GemmProblem* __restrict__ probs, int nprob, const int* __restrict__ skip) {
  skip = csd::bshift(skip, probs[0].bstride);
  if (skip && *skip) return;
  using Cfg = GemmCfg<TILE, WM, WN>;
  using SA = OperandSmem<TILE, AK, KC>;
  using SB = OperandSmem<TILE, BK_, KC>;
  constexpr int FM = Cfg::FM, FN = Cfg::FN;
  CS_DYN_SMEM(double, smem);
  double* sA = smem;
  double* sB = smem + NS * SA::SIZE;
  __shared__ TileCtx sctx;

  const int tid = threadIdx.x;
  if (tid == 0) { int pi = 0; tile_setup<TILE, AK, BK_, KC>(probs, nprob, (int)blockIdx.x, pi, sctx); }
  __syncthreads();
  const TileCtx cur = sctx;
  const int nchunks = cur.nchunks;
  const long long achunk = AK ? (long long)KC : (long long)KC * cur.lda;   // per-chunk advance of the operand bases
  const long long bchunk = BK_ ? (long long)KC : (long long)KC * cur.ldb;
  // measured (tools/gemm_bench, 8880 tiles, K = 512): NT 33.5 -> 34.5, NN 33.5 -> 34.8 TFLOP/s with the precomputed offsets, which
  // closes the gap of the m-major operands; with both operands k-major the plain form stays ahead (34.7 vs 34.4)
  constexpr bool PLAN = !(AK && BK_);
  using CA = OperandCopy<TILE, AK, Cfg::THREADS, KC>;
  using CB = OperandCopy<TILE, BK_, Cfg::THREADS, KC>;
  const double* gA = cur.A + CA::goff0(tid, cur.lda);    // this thread's first piece of chunk 0
  const double* gB = cur.B + CB::goff0(tid, cur.ldb);
  const long long gsA = CA::ROWS * cur.lda, gsB = CB::ROWS * cur.ldb;
  double* sAt = sA + CA::soff0(tid);
  double* sBt = sB + CB::soff0(tid);

  const int lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (wid % WM) * Cfg::WTM, wn = (wid / WM) * Cfg::WTN;

  // Accumulators start from (beta/alpha)*C so the read of C overlaps the pipeline fill and
  // the epilogue is store-only:  C = alpha * ((beta/alpha) C + sum_k A B^T).
  double acc[FM][FN][2];
  const unsigned neg = cur.neg;
  if (cur.beta_one) {  // acc = (beta / alpha) C = +-C
#pragma unroll
    for (int a = 0; a < FM; ++a) {
      const int r = wm + a * 8 + g;
#pragma unroll
      for (int b = 0; b < FN; ++b) {
        const double* p0 = cur.C + r + (long long)(wn + b * 8 + 2 * t) * cur.ldc;
        acc[a][b][0] = flip_sign(p0[0], neg);
        acc[a][b][1] = flip_sign(p0[cur.ldc], neg);
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < FM; ++a)
#pragma unroll
      for (int b = 0; b < FN; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
  }

  // prologue
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) {
    if (s < nchunks) {
      if (PLAN) {
        CA::copy(sAt + s * SA::SIZE, gA + s * achunk, gsA);
        CB::copy(sBt + s * SB::SIZE, gB + s * bchunk, gsB);
      } else {
        load_operand<TILE, AK, Cfg::THREADS, KC>(sA + s * SA::SIZE, cur.A + s * achunk, cur.lda, tid);
        load_operand<TILE, BK_, Cfg::THREADS, KC>(sB + s * SB::SIZE, cur.B + s * bchunk, cur.ldb, tid);
      }
    }
    csd::cp_async_commit();
  }

  for (int c = 0; c < nchunks; ++c) {
    csd::cp_async_wait<NS - 2>();
    __syncthreads();
    {
      const int cn = c + NS - 1;
      if (cn < nchunks) {
        const int sn = cn % NS;
        if (PLAN) {
          CA::copy(sAt + sn * SA::SIZE, gA + cn * achunk, gsA);
          CB::copy(sBt + sn * SB::SIZE, gB + cn * bchunk, gsB);
        } else {
          load_operand<TILE, AK, Cfg::THREADS, KC>(sA + sn * SA::SIZE, cur.A + cn * achunk, cur.lda, tid);
          load_operand<TILE, BK_, Cfg::THREADS, KC>(sB + sn * SB::SIZE, cur.B + cn * bchunk, cur.ldb, tid);
        }
      }
      csd::cp_async_commit();
    }
    const double* a_s = sA + (c % NS) * SA::SIZE;
    const double* b_s = sB + (c % NS) * SB::SIZE;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      double af[FM], bf[FN];
#pragma unroll
      for (int a = 0; a < FM; ++a)
        af[a] = AK ? a_s[(wm + a * 8 + g) * SA::LD + kk + t] : a_s[(kk + t) * SA::LD + wm + a * 8 + g];
#pragma unroll
      for (int b = 0; b < FN; ++b)
        bf[b] = BK_ ? b_s[(wn + b * 8 + g) * SB::LD + kk + t] : b_s[(kk + t) * SB::LD + wn + b * 8 + g];
#pragma unroll
      for (int a = 0; a < FM; ++a)
#pragma unroll
        for (int b = 0; b < FN; ++b) csd::dmma8x8x4(acc[a][b], af[a], bf[b]);
    }
  }
  csd::cp_async_wait<0>();

  // epilogue: C = alpha*acc = +-acc (+ mirrored store)
#pragma unroll
  for (int a = 0; a < FM; ++a) {
    const int r = wm + a * 8 + g;
#pragma unroll
    for (int b = 0; b < FN; ++b) {
      const int cc = wn + b * 8 + 2 * t;
      double* p0 = cur.C + r + (long long)cc * cur.ldc;
      const double v0 = flip_sign(acc[a][b][0], neg), v1 = flip_sign(acc[a][b][1], neg);
      p0[0] = v0; p0[cur.ldc] = v1;
      if (cur.Cm) {
        double* q = cur.Cm + cc + (long long)r * cur.ldc;  // (cc, r) and (cc+1, r) are adjacent
        q[0] = v0; q[1] = v1;
      }
    }
  }
}

}  // namespace csg
