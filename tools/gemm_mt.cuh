// Experiment, not shipped: several consecutive tiles per CTA with the cp.async pipeline running across the tile boundary
// (profiles/r02_gemm_experiments.md §6: no gain over one tile per CTA at any K -- the hardware already overlaps one CTA's
// prologue / epilogue with the main loops of its three neighbours).  Used by tools/gemm_bench.cu only.
#pragma once
#include "gemm.cuh"
namespace csg {
// Several consecutive tiles per CTA with the operand pipeline running across the tile boundary.
//
// gemm_kernel pays, per tile, a CTA launch, three dependent descriptor loads, the first operand round trip to
// HBM and the drain of the epilogue: about 10 % of a 64 x 64 x 512 tile (K = 4096 tiles run at 35, K = 512 tiles at 31.9
// TFLOP/s).  Here a CTA owns `tpc` consecutive tiles of the launch.  While it multiplies the last NS-1 chunks of a tile its
// cp.async copies already fetch the first chunks of the next one, the next tile's descriptor is decoded by one thread
// during the main loop (handed over through shared memory) and its C tile is pulled into L2 by prefetches, so a
// tile switch costs the epilogue stores plus an L2 hit.  Per-tile arithmetic (k order, fragments) is that of gemm_kernel:
// results are bitwise identical.
template <int TILE, int WM, int WN, bool AK, bool BK_, int KC = GEMM_BK, int NS = GEMM_STAGES>
__global__ void __launch_bounds__(WM * WN * 32)
gemm_kernel_mt(const GemmProblem* __restrict__ probs, int nprob, const int* __restrict__ skip, int tpc, int total_tiles) {
  skip = csd::bshift(skip, probs[0].bstride);
  if (skip && *skip) return;
  using Cfg = GemmCfg<TILE, WM, WN>;
  using SA = OperandSmem<TILE, AK, KC>;
  using SB = OperandSmem<TILE, BK_, KC>;
  constexpr int FM = Cfg::FM, FN = Cfg::FN;
  CS_DYN_SMEM(double, smem);
  double* sA = smem;
  double* sB = smem + NS * SA::SIZE;
  __shared__ TileCtx sctx[2];

  const int tid = threadIdx.x;
  const int bt0 = (int)blockIdx.x * tpc;
  const int nt_cta = min(tpc, total_tiles - bt0);
  int pi = 0;   // thread 0 only: problem index of the tile decoded last (tiles of a CTA are consecutive, so it only grows)
  if (tid == 0) tile_setup<TILE, AK, BK_, KC>(probs, nprob, bt0, pi, sctx[0]);
  __syncthreads();

  const int lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (wid % WM) * Cfg::WTM, wn = (wid / WM) * Cfg::WTN;

  // pipeline prologue: the first NS-1 chunks of the first tile.  Every tile has at least NS chunks (K >= TILE >= NS * KC,
  // checked by the launcher), so the first copy that crosses into the next tile is issued behind the SECOND barrier of a tile.
  {
    const TileCtx& c0 = sctx[0];
#pragma unroll
    for (int s = 0; s < NS - 1; ++s) {
      load_operand<TILE, AK, Cfg::THREADS, KC>(sA + s * SA::SIZE, c0.A + s * (AK ? (long long)KC : (long long)KC * c0.lda), c0.lda, tid);
      load_operand<TILE, BK_, Cfg::THREADS, KC>(sB + s * SB::SIZE, c0.B + s * (BK_ ? (long long)KC : (long long)KC * c0.ldb), c0.ldb, tid);
      csd::cp_async_commit();
    }
  }
  int sc = 0;   // chunks consumed so far by this CTA: ring position = sc % NS
  for (int ti = 0; ti < nt_cta; ++ti) {
    const TileCtx cur = sctx[ti & 1];
    const bool has_next = ti + 1 < nt_cta;

    double acc[FM][FN][2];
    if (cur.beta_one) {
#pragma unroll
      for (int a = 0; a < FM; ++a) {
        const int r = wm + a * 8 + g;
#pragma unroll
        for (int b = 0; b < FN; ++b) {
          const double* p0 = cur.C + r + (long long)(wn + b * 8 + 2 * t) * cur.ldc;
          acc[a][b][0] = flip_sign(p0[0], cur.neg);
          acc[a][b][1] = flip_sign(p0[cur.ldc], cur.neg);
        }
      }
    } else {
#pragma unroll
      for (int a = 0; a < FM; ++a)
#pragma unroll
        for (int b = 0; b < FN; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    }
    const long long achunk = AK ? (long long)KC : (long long)KC * cur.lda;
    const long long bchunk = BK_ ? (long long)KC : (long long)KC * cur.ldb;

    for (int c = 0; c < cur.nchunks; ++c, ++sc) {
      csd::cp_async_wait<NS - 2>();
      __syncthreads();
      // the other slot of sctx held the previous tile, whose epilogue every thread has left behind at this barrier;
      // its readers (the cross-tile copies) come behind the next one
      if (c == 0 && has_next && tid == 0) tile_setup<TILE, AK, BK_, KC>(probs, nprob, bt0 + ti + 1, pi, sctx[(ti + 1) & 1]);
      {
        const int cn = c + NS - 1;
        const int sn = (sc + NS - 1) % NS;
        if (cn < cur.nchunks) {
          load_operand<TILE, AK, Cfg::THREADS, KC>(sA + sn * SA::SIZE, cur.A + cn * achunk, cur.lda, tid);
          load_operand<TILE, BK_, Cfg::THREADS, KC>(sB + sn * SB::SIZE, cur.B + cn * bchunk, cur.ldb, tid);
        } else if (has_next) {
          const TileCtx& nx = sctx[(ti + 1) & 1];
          const int cq = cn - cur.nchunks;
          if (cq < nx.nchunks) {
            load_operand<TILE, AK, Cfg::THREADS, KC>(sA + sn * SA::SIZE, nx.A + cq * (AK ? (long long)KC : (long long)KC * nx.lda), nx.lda, tid);
            load_operand<TILE, BK_, Cfg::THREADS, KC>(sB + sn * SB::SIZE, nx.B + cq * (BK_ ? (long long)KC : (long long)KC * nx.ldb), nx.ldb, tid);
          }
          if (cq == 0 && nx.beta_one && tid < TILE) {   // pull the next C tile into L2: one column (TILE doubles) per thread
            const double* pc = nx.C + (long long)tid * nx.ldc;
#pragma unroll
            for (int q = 0; q < TILE; q += 16) csd::prefetch_l2(pc + q);
          }
        }
        csd::cp_async_commit();
      }
      const double* a_s = sA + (sc % NS) * SA::SIZE;
      const double* b_s = sB + (sc % NS) * SB::SIZE;
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

    // epilogue: C = +-acc (+ mirrored store)
#pragma unroll
    for (int a = 0; a < FM; ++a) {
      const int r = wm + a * 8 + g;
#pragma unroll
      for (int b = 0; b < FN; ++b) {
        const int cc = wn + b * 8 + 2 * t;
        double* p0 = cur.C + r + (long long)cc * cur.ldc;
        const double v0 = flip_sign(acc[a][b][0], cur.neg), v1 = flip_sign(acc[a][b][1], cur.neg);
        p0[0] = v0; p0[cur.ldc] = v1;
        if (cur.Cm) {
          double* q = cur.Cm + cc + (long long)r * cur.ldc;  // (cc, r) and (cc+1, r) are adjacent
          q[0] = v0; q[1] = v1;
        }
      }
    }
  }
  csd::cp_async_wait<0>();
}

}  // namespace csg
