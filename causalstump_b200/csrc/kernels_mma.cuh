// Gram (K1) and fused trace-gradient (K3/K4) kernels with their O(p) inner loops on the FP64 tensor path.
// Same arithmetic as kernels.cuh (reference: kernmat_*_SE_cpp, grad_*_SE_cpp + evid_scale_gradients), reorganised so that
// the two O(p n^2) passes are small matrix products issued as DMMA.8x8x4 instead of per-entry subtract / square / FMA chains:
//
//   pass 1 (exponents)   E[r,c] = sum_i a_i (x_ri - x_ci)^2 = u_r + v_c - 2 sum_i (a_i x_ri) x_ci            (64 x 64 x p product)
//   pass 2 (traces)      sum_{r,c} W[r,c] (x_ri - x_ci)^2 = sum_r x_ri (x_ri R[r] - 2 P[r,i]) + sum_c x_ci^2 C[c],
//                        P = W Xc (64 x p x 64 product), R / C = row / column sums of W = T o Km (and T o Ka)
//
// Vector and tensor FP64 share one pipe on sm_100a (tools/pipe_probe.cu), so the gain is not a second pipe: a DMMA retires
// 512 flops per issue against 64 for a warp-wide DFMA, the subtract / square of every (entry, dimension) pair disappears,
// and the per-dimension accumulators become 16 registers per thread instead of 2p+2 shared-memory slots per thread (106 KB).
//
// Thread layout of a 64 x 64 tile (8 warps): warp w owns rows 8w..8w+7; lane (g = lane>>2, t = lane&3) owns row 8w+g and the
// 16 columns 8cb + 2t + i (cb = 0..7, i = 0,1) -- exactly the accumulator fragment of DMMA.8x8x4, so the pass-1 result is used
// in place and W = T o K feeds pass 2 as the A fragment without leaving the registers.
// Covariate tiles are staged as [32 dims][XS = 68] (rows >= p are zero): fragment loads of pass 1 are bank-conflict-free.
// They arrive by bulk asynchronous copies (cp.async.bulk + mbarrier, SASS UBLKCP), the tile after next being prefetched while
// the current one is processed.  p <= 32; larger p uses the kernels of kernels.cuh.
#pragma once
#include "kernels.cuh"

namespace csk {

constexpr int XS = 68;         // doubles per staged dimension row
constexpr int XD = 32;         // staged dimensions
constexpr int XROWS = XD + 4;  // staged rows of a tile: the dimensions, then z, the observation weights and the two scaled norms
constexpr int XT = XROWS * XS; // doubles per staged tile
constexpr int ROW_Z = XD, ROW_W = XD + 1, ROW_UM = XD + 2, ROW_UA = XD + 3;
// per-evaluation constants of the two kernels, computed once by the pack kernel instead of by every CTA:
// exp(-Lm_k) | exp(-La_k) | 2^(j/64) (csd::exp_tab) | lambda_m, lambda_a, exp(sigma), 5 x padding
constexpr int CST_EA = 0, CST_EB = XD, CST_TAB = 2 * XD, CST_SC = 2 * XD + csd::EXP_TAB, CST_N = CST_SC + 8;

inline bool mma_kernels_ok(int p) { return p <= XD; }

__device__ __forceinline__ void stage_x_coop(double* __restrict__ xs, const double* __restrict__ X, long long ldx, int row0, int p,
                                             int tid) {
  for (int idx = tid; idx < p * TB; idx += ETH) {
    const int d = idx / TB, r = idx - d * TB;
    xs[d * XS + r] = X[(long long)d * ldx + row0 + r];
  }
}

// Packed tiles: Xt[blk][XROWS][XS] holds, for the 64 observations of block blk, the p covariate rows (zero up to XD), z, w
// and the scaled norms u_m[i] = sum_k exp(-Lm_k) x_ik^2, u_a[i] (same with La), already in the padded shared-memory layout, so
// that a kernel stages a whole operand tile with ONE bulk copy; behind the nblk tiles sit the CST_N per-evaluation constants.
// (A tile used to arrive as 2p bulk copies of 512 bytes; the copy engine takes them one at a time, about 100 cycles each, and
// the seven other warps waited for the issuing warp at the next barrier: 18 % of all stall samples of the gradient kernel.
// The constants -- 2p + 1 library exp, the 64-entry exp2 table -- and the norms were recomputed by every CTA: with one tile
// per CTA that was a quarter of the Gram kernel, profiles/r02_ncu_grad_source.md.)
// Rebuilt at the start of every evaluation: the parameters change with every iteration, X, z, w may have been rewritten.
__device__ __forceinline__ void eval_constants(const double* __restrict__ params, const Layout& lay, int tid, double* __restrict__ cst) {
  const int p = lay.p;
  if (tid < XD) {
    cst[CST_EA + tid] = tid < p ? exp(-params[lay.i_Lm() + tid]) : 0.0;
    cst[CST_EB + tid] = tid < p ? exp(-params[lay.i_La() + tid]) : 0.0;
  }
  csd::exp_tab_fill(cst + CST_TAB, tid);
  if (tid == 0) {
    cst[CST_SC + 0] = params[lay.i_lm()];
    cst[CST_SC + 1] = params[lay.i_la()];
    cst[CST_SC + 2] = exp(params[lay.i_sigma()]);
#pragma unroll
    for (int q = 3; q < 8; ++q) cst[CST_SC + q] = 0.0;
  }
}

__global__ void __launch_bounds__(ETH)
pack_x_tiles_kernel(const double* __restrict__ X, long long ldx, const double* __restrict__ z, const double* __restrict__ w,
                    const double* __restrict__ params, Layout lay, int nblk, double* __restrict__ Xt, const int* __restrict__ skip) {
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(X); CS_BS(z); CS_BS(w); CS_BS(params); CS_BS(Xt);
  __shared__ double cst[CST_N];
  const int p = lay.p, blk = (int)blockIdx.x, tid = (int)threadIdx.x;
  eval_constants(params, lay, tid, cst);
  __syncthreads();
  double* T = Xt + (long long)blk * XT;
  for (int idx = tid; idx < ROW_UM * XS; idx += ETH) {
    const int row = idx / XS, c = idx - row * XS;
    double v = 0.0;
    if (c < TB) {
      const long long gi = (long long)blk * TB + c;
      if (row < p) v = X[(long long)row * ldx + gi];
      else if (row == ROW_Z) v = z[gi];
      else if (row == ROW_W) v = w[gi];
    }
    T[idx] = v;
  }
  if (tid < 2 * XS) {   // norms: thread = (matrix, observation)
    const int which = tid / XS, c = tid - which * XS;
    double sacc = 0.0;
    if (c < TB) {
      const double* e = cst + (which ? CST_EB : CST_EA);
      const long long gi = (long long)blk * TB + c;
      for (int k = 0; k < p; ++k) {
        const double x = X[(long long)k * ldx + gi];
        sacc = fma(e[k] * x, x, sacc);
      }
    }
    T[(ROW_UM + which) * XS + c] = sacc;
  }
  if (blk == 0 && tid < CST_N) Xt[(long long)nblk * XT + tid] = cst[tid];
}

// cooperative staging of the same layout (emulator, table-driven tiles of the distributed engine, CS_NO_TMA)
__device__ __forceinline__ void stage_tile_coop(double* __restrict__ xs, const double* __restrict__ X, long long ldx,
                                                const double* __restrict__ z, const double* __restrict__ w, int row0, int p, int tid) {
  stage_x_coop(xs, X, ldx, row0, p, tid);
  if (tid < TB) { xs[ROW_Z * XS + tid] = z[row0 + tid]; xs[ROW_W * XS + tid] = w ? w[row0 + tid] : 0.0; }
}

#ifndef CS_EMU
// one thread: one bulk copy of `bytes` completing on `bar` (the byte count has been posted by mbar_expect_tx)
__device__ __forceinline__ void bulk_copy(double* dst, const double* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(csd::smem_u32(dst)), "l"(src), "r"(bytes), "r"(csd::smem_u32(bar)) : "memory");
}
#endif

// cooperative staging only: the norm rows of the row tile and of the column tile (packed tiles bring them along), one value
// per thread; a barrier separates the staging before and the readers behind
__device__ __forceinline__ void tile_norms(double* __restrict__ xr, double* __restrict__ xc, const double* __restrict__ ea,
                                           const double* __restrict__ eb, int p, int tid) {
  const int which = tid >> 6, idx = tid & 63;
  double* xs = which < 2 ? xr : xc;
  const double* e = (which & 1) ? eb : ea;
  double s = 0.0;
  for (int k = 0; k < p; ++k) {
    const double x = xs[k * XS + idx];
    s = fma(e[k] * x, x, s);
  }
  xs[(ROW_UM + (which & 1)) * XS + idx] = s;
}

// D[cb][i] = sum_k e_k x_rk x_ck for the thread's row and 16 columns, both scalings at once
__device__ __forceinline__ void pass1_mma(const double* __restrict__ xr, const double* __restrict__ xc, const double* __restrict__ ea,
                                          const double* __restrict__ eb, int p, int r, int g, int t, double (&Dm)[8][2],
                                          double (&Da)[8][2]) {
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) { Dm[cb][0] = Dm[cb][1] = 0.0; Da[cb][0] = Da[cb][1] = 0.0; }
  const int ks = (p + 3) >> 2;
  for (int s = 0; s < ks; ++s) {
    const int k = 4 * s + t;
    const double xrk = xr[k * XS + r];
    const double am = ea[k] * xrk, aa = eb[k] * xrk;
    const double* xck = xc + k * XS + g;
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
      const double b = xck[8 * cb];
      csd::dmma8x8x4(Dm[cb], am, b);
      csd::dmma8x8x4(Da[cb], aa, b);
    }
  }
}

// ---------------------------------------------------------------------------------
// K1 on the tensor path
// ---------------------------------------------------------------------------------
inline int gram_mma_smem_bytes() { return (2 * XT + CST_N) * (int)sizeof(double); }

__global__ void __launch_bounds__(ETH, 3)
gram_train_mma_kernel(const double* __restrict__ X, long long ldx, const double* __restrict__ z, const double* __restrict__ w,
                      const double* __restrict__ params, Layout lay, int n, int nblk, double* __restrict__ G, long long ldg,
                      const int* __restrict__ skip, int tile0, const double* __restrict__ Xt) {
  // Xt != nullptr: packed tiles + constants (pack_x_tiles_kernel), three bulk copies per CTA; nullptr: cooperative staging
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(X); CS_BS(z); CS_BS(w); CS_BS(params); CS_BS(G); CS_BS(Xt);
  CS_DYN_SMEM(double, sm_);
  const int p = lay.p;
  double* xr = sm_;              // row tile: dimensions | z | w | u_m | u_a
  double* xc = xr + XT;          // column tile
  double* cst = xc + XT;         // exp(-Lm) | exp(-La) | exp table | lambda_m, lambda_a, exp(sigma)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int tl = (int)blockIdx.x + tile0;
  const int I = csg_lower_row(tl), J = tl - I * (I + 1) / 2;
  const int r0 = I * TB, c0 = J * TB;
#ifndef CS_EMU
  __shared__ unsigned long long xbar;
  const bool bulk = Xt != nullptr;
  if (bulk) {
    if (tid == 0) {
      csd::mbar_init(&xbar, 1);
      csd::mbar_fence_init();
      csd::mbar_expect_tx(&xbar, (2u * XT + CST_N) * (unsigned)sizeof(double));
      bulk_copy(xr, Xt + (long long)I * XT, XT * (unsigned)sizeof(double), &xbar);
      bulk_copy(xc, Xt + (long long)J * XT, XT * (unsigned)sizeof(double), &xbar);
      bulk_copy(cst, Xt + (long long)nblk * XT, CST_N * (unsigned)sizeof(double), &xbar);
    }
    __syncthreads();   // the barrier object is initialised
    csd::mbar_wait(&xbar, 0);
  } else
#else
  if (Xt != nullptr) {   // emulator: the three bulk copies as plain copies, so that the CPU tests consume the packed layout too
    if (tid == 0) {
      std::memcpy(xr, Xt + (long long)I * XT, XT * sizeof(double));
      std::memcpy(xc, Xt + (long long)J * XT, XT * sizeof(double));
      std::memcpy(cst, Xt + (long long)nblk * XT, CST_N * sizeof(double));
    }
    __syncthreads();
  } else
#endif
  {
    for (int idx = tid; idx < (XD - p) * XS; idx += ETH) { xr[p * XS + idx] = 0.0; xc[p * XS + idx] = 0.0; }
    stage_tile_coop(xr, X, ldx, z, w, r0, p, tid);
    stage_tile_coop(xc, X, ldx, z, w, c0, p, tid);
    eval_constants(params, lay, tid, cst);
    __syncthreads();
    tile_norms(xr, xc, cst + CST_EA, cst + CST_EB, p, tid);
    __syncthreads();
  }
  const double* zr = xr + ROW_Z * XS;
  const double* zc = xc + ROW_Z * XS;
  const double* wr = xr + ROW_W * XS;
  const double* vm = xc + ROW_UM * XS;
  const double* va = xc + ROW_UA * XS;
  const double* etab = cst + CST_TAB;
  const int r = 8 * wid + g;
  double Dm[8][2], Da[8][2];
  pass1_mma(xr, xc, cst + CST_EA, cst + CST_EB, p, r, g, t, Dm, Da);
  const double lam_m = cst[CST_SC], lam_a = cst[CST_SC + 1], esig = cst[CST_SC + 2];
  const int gr = r0 + r;
  const double umr = xr[ROW_UM * XS + r], uar = xr[ROW_UA * XS + r], zrr = zr[r];
#pragma unroll
  for (int cb = 0; cb < 8; ++cb)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int cl = 8 * cb + 2 * t + i, gc = c0 + cl;
      double v;
      if (gr >= n || gc >= n) v = (gr == gc) ? 1.0 : 0.0;
      else if (gr < gc) v = 0.0;
      else {
        const double em = fma(-2.0, Dm[cb][i], umr + vm[cl]);
        const double ea_ = fma(-2.0, Da[cb][i], uar + va[cl]);
        v = csd::exp_tab(lam_m - em, etab) + csd::exp_tab(lam_a - ea_, etab) * zrr * zc[cl];
        if (gr == gc) v += esig / wr[r];
      }
      G[(long long)gc * ldg + gr] = v;
    }
}

// ---------------------------------------------------------------------------------
// K3 / K4 on the tensor path
// ---------------------------------------------------------------------------------
inline int grad_mma_smem_bytes() {
  return (4 * XT + CST_N + 4 * TB + 24 * TB + 3 * TB + 16 * XD + 32) * (int)sizeof(double);
}

// column sums over the 8 rows of a warp (lanes that differ in g) of a 16-vector indexed e = 2cb + i: recursive halving, after
// which lane (g, t) holds the sums of the two columns 8g + 2t + {0, 1}.  42 shuffles instead of 48 x 3.  Deterministic.
__device__ __forceinline__ void col_butterfly(const double (&v)[16], int g, double& o0, double& o1) {
  const bool b2 = (g & 4) != 0, b1 = (g & 2) != 0, b0 = (g & 1) != 0;
  double a[8], b[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double keep = b2 ? v[j + 8] : v[j], send = b2 ? v[j] : v[j + 8];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double keep = b1 ? a[j + 4] : a[j], send = b1 ? a[j] : a[j + 4];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const double k0 = b0 ? b[2] : b[0], s0 = b0 ? b[0] : b[2];
    const double k1 = b0 ? b[3] : b[1], s1 = b0 ? b[1] : b[3];
    o0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
    o1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
  }
}

template <bool FROM_MEM>
__global__ void __launch_bounds__(ETH, 1)
grad_trace_mma_kernel(const double* __restrict__ X, long long ldx, const double* __restrict__ z,
                      const double* __restrict__ params, Layout lay, int n, int nblk,
                      const double* __restrict__ Cinv, long long ldc, const double* __restrict__ alpha,
                      const EvalScalars* __restrict__ sc, double* __restrict__ gpart, double* __restrict__ kal_part,
                      long long kal_stride, const double* __restrict__ Kmat_in, const double* __restrict__ Km_in,
                      const double* __restrict__ Ka_in, const int* __restrict__ skip,
                      const int* __restrict__ tile_tab, const double* __restrict__ Xt) {
  // outputs and the tile_tab (distributed) mode exactly as grad_trace_kernel (kernels.cuh)
  // Xt != nullptr: packed tiles (pack_x_tiles_kernel): every operand tile is one bulk copy, prefetched one tile ahead
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(X); CS_BS(z); CS_BS(params); CS_BS(Cinv); CS_BS(alpha); CS_BS(sc); CS_BS(gpart); CS_BS(kal_part); CS_BS(Xt);
  CS_DYN_SMEM(double, sm_);
  const int p = lay.p;
  const int nacc = 2 * p + 2;
  double* xbuf = sm_;                  // [2 buffers][row tile | column tile][XT]
  double* cst = xbuf + 4 * XT;         // exp(-Lm) | exp(-La) | exp table | scalars (pack_x_tiles_kernel / eval_constants)
  double* abuf = cst + CST_N;          // [2 buffers][alpha rows | alpha cols][TB]
  double* colred = abuf + 4 * TB;      // [8 warps][3][TB]: column sums of T o Km, T o Ka, K alpha_r per warp
  double* colsum = colred + 24 * TB;   // [3][TB]
  double* fin = colsum + 3 * TB;       // [8][2][XD] final reduction scratch
  double* red = fin + 16 * XD;         // [32]
  const double* ea = cst + CST_EA;
  const double* eb = cst + CST_EB;
  const double* etab = cst + CST_TAB;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const bool bulk = Xt != nullptr && !tile_tab;
  if (!bulk) {
    for (int idx = tid; idx < 4 * XT; idx += ETH) xbuf[idx] = 0.0;   // cooperative staging: rows >= p stay zero for the whole kernel
    eval_constants(params, lay, tid, cst);
  }
#ifdef CS_EMU
  else if (tid == 0) std::memcpy(cst, Xt + (long long)nblk * XT, CST_N * sizeof(double));   // emulator: copies instead of bulk copies
#endif
  const double lam_m = params[lay.i_lm()], lam_a = params[lay.i_la()];
  const double coef = sc->coef;
  const int ntiles = tile_tab ? nblk : nblk * (nblk + 1) / 2;
  __syncthreads();
#ifndef CS_EMU
  __shared__ unsigned long long xbar[2];
  constexpr unsigned TILE_TX = (2u * XT + 2u * TB) * (unsigned)sizeof(double);   // two operand tiles + alpha rows / cols
  if (bulk && tid == 0) {
    csd::mbar_init(&xbar[0], 1);
    csd::mbar_init(&xbar[1], 1);
    csd::mbar_fence_init();
    const int t0 = (int)blockIdx.x;
    if (t0 < ntiles) {
      int I0, J0;
      lower_tile_of(t0, I0, J0);
      csd::mbar_expect_tx(&xbar[0], TILE_TX + CST_N * (unsigned)sizeof(double));
      bulk_copy(cst, Xt + (long long)nblk * XT, CST_N * (unsigned)sizeof(double), &xbar[0]);
      bulk_copy(xbuf, Xt + (long long)I0 * XT, XT * (unsigned)sizeof(double), &xbar[0]);
      bulk_copy(xbuf + XT, Xt + (long long)J0 * XT, XT * (unsigned)sizeof(double), &xbar[0]);
      bulk_copy(abuf, alpha + I0 * TB, TB * (unsigned)sizeof(double), &xbar[0]);
      bulk_copy(abuf + TB, alpha + J0 * TB, TB * (unsigned)sizeof(double), &xbar[0]);
    }
  }
  int local_tile = 0;
#endif
  // per-thread accumulators over all tiles of this CTA: dims 8nb + 2t + ii of the row-side terms, one (matrix, dim,
  // column quarter) of the column-side term, and the two plain traces
  double acc_m[4][2], acc_a[4][2];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) { acc_m[nb][0] = acc_m[nb][1] = 0.0; acc_a[nb][0] = acc_a[nb][1] = 0.0; }
  double acc_col = 0.0, tm_acc = 0.0, ta_acc = 0.0;
  const int r = 8 * wid + g;
  const int nbk = (p + 7) >> 3;

  for (int tl = (int)blockIdx.x; tl < ntiles; tl += (int)gridDim.x) {
    int I, J;
    const double* Ct;
    if (tile_tab) {
      I = tile_tab[4 * tl]; J = tile_tab[4 * tl + 1];
      const long long coff = (long long)(unsigned)tile_tab[4 * tl + 2] | ((long long)tile_tab[4 * tl + 3] << 32);
      Ct = Cinv + coff;
    } else {
      lower_tile_of(tl, I, J);
      Ct = Cinv + (long long)J * TB * ldc + (long long)I * TB;
    }
    const int r0 = I * TB, c0 = J * TB;
    const double wgt = (I == J) ? 1.0 : 2.0;
    __syncthreads();  // previous tile fully consumed
    double* xr = xbuf;
    double* xc = xbuf + XT;
    double* ar_ = abuf;                  // alpha rows | alpha cols of this tile
#ifndef CS_EMU
    if (bulk) {
      const int b = local_tile & 1;
      if (tid == 0) {
        const int tn = tl + (int)gridDim.x;
        if (tn < ntiles) {
          int In, Jn;
          lower_tile_of(tn, In, Jn);
          double* xn = xbuf + (b ^ 1) * 2 * XT;
          double* an = abuf + (b ^ 1) * 2 * TB;
          csd::fence_proxy_async();   // the generic-proxy reads of the tile before last (barrier above) precede the refill
          csd::mbar_expect_tx(&xbar[b ^ 1], TILE_TX);
          bulk_copy(xn, Xt + (long long)In * XT, XT * (unsigned)sizeof(double), &xbar[b ^ 1]);
          bulk_copy(xn + XT, Xt + (long long)Jn * XT, XT * (unsigned)sizeof(double), &xbar[b ^ 1]);
          bulk_copy(an, alpha + In * TB, TB * (unsigned)sizeof(double), &xbar[b ^ 1]);
          bulk_copy(an + TB, alpha + Jn * TB, TB * (unsigned)sizeof(double), &xbar[b ^ 1]);
        }
      }
      xr = xbuf + b * 2 * XT; xc = xr + XT;
      ar_ = abuf + b * 2 * TB;
    } else
#else
    if (bulk) {   // emulator: the tile of THIS iteration, copied by one thread (no prefetch)
      if (tid == 0) {
        std::memcpy(xr, Xt + (long long)I * XT, XT * sizeof(double));
        std::memcpy(xc, Xt + (long long)J * XT, XT * sizeof(double));
        std::memcpy(ar_, alpha + r0, TB * sizeof(double));
        std::memcpy(ar_ + TB, alpha + c0, TB * sizeof(double));
      }
    } else
#endif
    {
      stage_tile_coop(xr, X, ldx, z, nullptr, r0, p, tid);
      stage_tile_coop(xc, X, ldx, z, nullptr, c0, p, tid);
      if (tid < TB) { ar_[tid] = alpha[r0 + tid]; ar_[TB + tid] = alpha[c0 + tid]; }
    }
    const double* zr = xr + ROW_Z * XS;
    const double* zc = xc + ROW_Z * XS;
    const double* ar = ar_;
    const double* ac = ar_ + TB;
    // the K^-1 entries of this tile are requested before anything waits (HBM latency hides behind the staging and pass 1)
    const int gr = r0 + r;
    double tv[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int cl = 8 * (e >> 1) + 2 * t + (e & 1);
      tv[e] = (gr < n && c0 + cl < n) ? Ct[(long long)cl * ldc + r] : 0.0;
    }
    __syncthreads();
#ifndef CS_EMU
    if (bulk) { csd::mbar_wait(&xbar[local_tile & 1], (unsigned)((local_tile >> 1) & 1)); ++local_tile; }
#endif
    double Wm[16], Wa[16];   // pass-1 products, then T o Km and T o Ka
    if (!FROM_MEM) {
      if (!bulk) {   // packed tiles bring their norm rows along
        tile_norms(xr, xc, ea, eb, p, tid);
        __syncthreads();
      }
      double Dm[8][2], Da[8][2];
      pass1_mma(xr, xc, ea, eb, p, r, g, t, Dm, Da);
#pragma unroll
      for (int e = 0; e < 16; ++e) { Wm[e] = Dm[e >> 1][e & 1]; Wa[e] = Da[e >> 1][e & 1]; }
    }
    const double* vm = xc + ROW_UM * XS;
    const double* va = xc + ROW_UA * XS;
    // ---- per entry: Km, Ka, T = K^-1 - coef alpha alpha^T, W = T o K; row / column pieces of K alpha
    double kc[16];           // K[r,c] alpha_r: column partials of K alpha
    double rm = 0.0, ra = 0.0, rk = 0.0;
    {
      const double umr = xr[ROW_UM * XS + r], uar = xr[ROW_UA * XS + r], zrr = zr[r], arr = ar[r];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int cl = 8 * (e >> 1) + 2 * t + (e & 1), gc = c0 + cl;
        double vkm = 0.0, vka = 0.0, T = 0.0, kfull = 0.0;
        if (gr < n && gc < n) {
          if (FROM_MEM) {
            vkm = Km_in[(long long)gc * ldc + gr];
            vka = Ka_in[(long long)gc * ldc + gr];
            kfull = Kmat_in[(long long)gc * ldc + gr];
          } else {
            vkm = csd::exp_tab(lam_m - fma(-2.0, Wm[e], umr + vm[cl]), etab);
            vka = csd::exp_tab(lam_a - fma(-2.0, Wa[e], uar + va[cl]), etab) * zrr * zc[cl];
            kfull = vkm + vka;
          }
          T = tv[e] - coef * arr * ac[cl];
        }
        rk = fma(kfull, ac[cl], rk);
        kc[e] = kfull * arr;
        Wm[e] = T * vkm;
        Wa[e] = T * vka;
        rm += Wm[e];
        ra += Wa[e];
      }
    }
    tm_acc = fma(wgt, rm, tm_acc);
    ta_acc = fma(wgt, ra, ta_acc);
    // full row sums over the 64 columns: the four lanes of a quad hold the same row
    rm += __shfl_xor_sync(0xffffffffu, rm, 1); rm += __shfl_xor_sync(0xffffffffu, rm, 2);
    ra += __shfl_xor_sync(0xffffffffu, ra, 1); ra += __shfl_xor_sync(0xffffffffu, ra, 2);
    rk += __shfl_xor_sync(0xffffffffu, rk, 1); rk += __shfl_xor_sync(0xffffffffu, rk, 2);
    if (t == 0) {  // K alpha, row part: every row of the tile belongs to exactly one quad
      if (tile_tab) atomicAdd(kal_part + r0 + r, rk);
      else kal_part[(long long)J * kal_stride + r0 + r] = rk;
    }
    // column sums over this warp's 8 rows, then across the warps through shared memory
    {
      double o0, o1;
      double* cr = colred + wid * 3 * TB + 8 * g + 2 * t;
      col_butterfly(Wm, g, o0, o1); cr[0] = o0; cr[1] = o1;
      col_butterfly(Wa, g, o0, o1); cr[TB] = o0; cr[TB + 1] = o1;
      col_butterfly(kc, g, o0, o1); cr[2 * TB] = o0; cr[2 * TB + 1] = o1;
    }
    // ---- pass 2: P = W Xc on the tensor path (A fragments straight from the registers), then the row-side terms
    {
      double Pm[4][2], Pa[4][2];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) { Pm[nb][0] = Pm[nb][1] = 0.0; Pa[nb][0] = Pa[nb][1] = 0.0; }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const double* xce = xc + g * XS + 8 * (e >> 1) + 2 * t + (e & 1);
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          if (nb < nbk) {
            const double b = xce[8 * nb * XS];
            csd::dmma8x8x4(Pm[nb], Wm[e], b);
            csd::dmma8x8x4(Pa[nb], Wa[e], b);
          }
        }
      }
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
          const int d = 8 * nb + 2 * t + ii;
          const double x = xr[d * XS + r];          // zero for d >= p
          acc_m[nb][ii] = fma(wgt * x, fma(x, rm, -2.0 * Pm[nb][ii]), acc_m[nb][ii]);
          acc_a[nb][ii] = fma(wgt * x, fma(x, ra, -2.0 * Pa[nb][ii]), acc_a[nb][ii]);
        }
    }
    __syncthreads();  // colred complete
    if (tid < 3 * TB) {
      const int q = tid >> 6, c = tid & 63;
      double s = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += colred[w8 * 3 * TB + q * TB + c];
      colsum[q * TB + c] = s;
      if (q == 2 && I != J) {  // K alpha, column part
        if (tile_tab) atomicAdd(kal_part + c0 + c, s);
        else kal_part[(long long)I * kal_stride + c0 + c] = s;
      }
    }
    __syncthreads();
    {  // column-side terms: thread = (matrix m, dim d, column quarter q4): sum_c x_cd^2 C_m[c] over the columns 4j + q4
      const int m = tid >> 7, d = (tid >> 2) & 31, q4 = tid & 3;
      const double* cs_ = colsum + m * TB;
      const double* xd = xc + d * XS;
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const double x = xd[4 * j + q4];
        s = fma(x * x, cs_[4 * j + q4], s);
      }
      acc_col = fma(wgt, s, acc_col);
    }
  }
  // ---- reductions over the CTA: gpart[cta][0] = sum T o Km, [1] = sum T o Ka, [2+i] / [2+p+i] = sums weighted by d_i^2
  __syncthreads();
#pragma unroll
  for (int nb = 0; nb < 4; ++nb)
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      double vm = acc_m[nb][ii], va = acc_a[nb][ii];
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) { vm += __shfl_xor_sync(0xffffffffu, vm, o); va += __shfl_xor_sync(0xffffffffu, va, o); }
      if (g == 0) {
        const int d = 8 * nb + 2 * t + ii;
        fin[(wid * 2 + 0) * XD + d] = vm;
        fin[(wid * 2 + 1) * XD + d] = va;
      }
    }
  acc_col += __shfl_xor_sync(0xffffffffu, acc_col, 1);
  acc_col += __shfl_xor_sync(0xffffffffu, acc_col, 2);
  if ((tid & 3) == 0) colsum[(tid >> 7) * XD + ((tid >> 2) & 31)] = acc_col;   // [matrix][dim]
  __syncthreads();
  if (tid < 2 * XD) {
    const int m = tid >> 5, d = tid & 31;
    if (d < p) {
      double s = colsum[m * XD + d];
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += fin[(w8 * 2 + m) * XD + d];
      gpart[(long long)blockIdx.x * nacc + 2 + m * p + d] = s;
    }
  }
  const double tms = csd::block_sum(tm_acc, red);
  const double tas = csd::block_sum(ta_acc, red);
  if (tid == 0) { gpart[(long long)blockIdx.x * nacc] = tms; gpart[(long long)blockIdx.x * nacc + 1] = tas; }
}

}  // namespace csk
