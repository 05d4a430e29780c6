// Gram (K1), K^-1 mat-vec, fused trace-gradient (K3/K4), finalize and optimizer (K5)
// kernels of the GP / Student-t-process fit.  Reference arithmetic:
//   kernmat_{GP,TP}_SE_cpp   /root/reference/src/kernel_GP_SE_cpp.cpp:72-140, kernel_TP_SE_cpp.cpp:46-90
//   grad_{GP,TP}_SE_cpp      kernel_GP_SE_cpp.cpp:143-202 (+ :8-44), kernel_TP_SE_cpp.cpp:100-163
//   mu_solution_cpp          kernel_GP_SE_cpp.cpp:61-70
//   Adam/Nadam/Nesterov_cpp  optimizer_cpp.cpp:9-63
// All matrices column-major fp64.  Vectors are padded to np (multiple of 64).
#pragma once
#include "cs_dev.cuh"

#define CS_BS(ptr) ptr = csd::bshift(ptr, lay.bstride)

namespace csk {

// row of the lower-triangular tile enumeration (same as csg::lower_row in gemm.cuh, which includes this header's users later)
__host__ __device__ inline int csg_lower_row(int t) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((long long)(i + 1) * (i + 2) / 2 <= t) ++i;
  while ((long long)i * (i + 1) / 2 > t) --i;
  return i;
}

constexpr int TB = 64;        // tile edge of the element-wise kernels
constexpr int ETH = 256;      // threads per CTA (16 x 16, 4 x 4 entries each)
constexpr int PMAX = 128;     // max covariates (== CS_MAX_P): sizes of the per-dimension shared-memory vectors
constexpr int SMEM_LIMIT = 226 * 1024;  // opt-in dynamic shared memory per CTA on sm_100a (227 KB) minus room for the static part

enum Kind { KIND_GP = 0, KIND_TP = 1 };
enum Optim { OPT_ADAM = 0, OPT_NADAM = 1, OPT_NESTEROV = 2 };

// flat parameter vector layout = reference list order
//   GP: sigma, lambdam, lambdaa, Lm[p], La[p], mu          (R/kernel_GP_SE_class.R:13-19)
//   TP: sigma, lambdam, nu, lambda0, Lm[p], La[p], mu      (R/kernel_TP_SE_class.R:14-21)
struct Layout {
  int kind, p;
  long long bstride = 0;  // batched launches: byte distance between the slabs of consecutive fits (0 = single fit)
  __host__ __device__ int base() const { return kind == KIND_TP ? 4 : 3; }
  __host__ __device__ int i_sigma() const { return 0; }
  __host__ __device__ int i_lm() const { return 1; }
  __host__ __device__ int i_nu() const { return 2; }                       // TP only
  __host__ __device__ int i_la() const { return kind == KIND_TP ? 3 : 2; }  // lambdaa / lambda0
  __host__ __device__ int i_Lm() const { return base(); }
  __host__ __device__ int i_La() const { return base() + p; }
  __host__ __device__ int i_mu() const { return base() + 2 * p; }
  __host__ __device__ int np() const { return base() + 2 * p + 1; }
};

// device-resident scalars of one evaluation
struct EvalScalars {
  double M;        // ybar' alpha
  double sum_alpha, sum_u, sum_s;
  double coef;     // 1 (GP) or TP_adjust (kernel_TP_SE_cpp.cpp:122)
  double logdet;   // log det K_n
  double mu_sol;   // 0.5 * sum(K^-1 y) / sum(K^-1)      (quirk Q1)
  double stats[2]; // squared residual, LML
};

// device-resident optimisation-loop state
struct LoopState {
  int it;          // next iteration number (1-based)
  int done;        // 1 once converged or maxiter reached
  int iters;       // iterations performed when done
  int maxiter;
  double tol;
  double prev_lml; // stats[2, iter] of the previous iteration (0 before the first)
  int err;         // 0, or why this member stopped: > 0 = failing pivot of a non positive definite K_n (arma::inv_sympd
                   // throws, src/kernel_GP_SE_cpp.cpp:56), LOOP_ERR_NONFINITE = |dLML| is NaN (R stops at R/main.R:82)
  int pad;
};
constexpr int LOOP_ERR_NONFINITE = -6;  // == CS_ERR_NONFINITE (include/cs_b200.h)

// ---------------------------------------------------------------------------------
// shared micro-tile: thread (tx,ty) owns rows tx+16a, cols ty+16b (a,b = 0..3) of a 64x64 tile
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void load_x_tile(double* __restrict__ xs, const double* __restrict__ X, long long ldx,
                                            int row0, int p, int tid) {
  for (int idx = tid; idx < p * TB; idx += ETH) {
    const int i = idx / TB, r = idx - i * TB;
    xs[idx] = X[(long long)i * ldx + row0 + r];
  }
}

// exponents sm = sum_i a_i d_i^2, sa = sum_i b_i d_i^2 for the 4x4 micro-tile
__device__ __forceinline__ void sqdist_4x4(const double* __restrict__ xr, const double* __restrict__ xc,
                                           const double* __restrict__ ea, const double* __restrict__ eb, int p,
                                           int tx, int ty, double (&sm)[4][4], double (&sa)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { sm[a][b] = 0.0; sa[a][b] = 0.0; }
  for (int i = 0; i < p; ++i) {
    double xa[4], xb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xa[a] = xr[i * TB + tx + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) xb[b] = xc[i * TB + ty + 16 * b];
    const double ai = ea[i], bi = eb[i];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double d = xa[a] - xb[b];
        const double d2 = d * d;
        sm[a][b] = fma(d2, ai, sm[a][b]);
        sa[a][b] = fma(d2, bi, sa[a][b]);
      }
  }
}

// ---------------------------------------------------------------------------------
// K1: training Gram K_n = Km + Ka + diag(exp(sigma)/w), lower tiles, identity padding.
// ---------------------------------------------------------------------------------
inline int gram_smem_bytes(int p) { return (2 * p * TB + 2 * PMAX + 3 * TB) * (int)sizeof(double); }

__global__ void __launch_bounds__(ETH)
gram_train_kernel(const double* __restrict__ X, long long ldx, const double* __restrict__ z,
                  const double* __restrict__ w, const double* __restrict__ params, Layout lay, int n,
                  double* __restrict__ G, long long ldg, const int* __restrict__ skip, int tile0,
                  const __grid_constant__ rt::XTma xt) {
  // tile0: first lower-triangular tile of this launch (the Gram build is split so that the first diagonal
  // block of the Cholesky can start while the rest of the matrix is still being built)
  // xt: TMA descriptor of X (rt::make_x_tensormap): the two [p][64] covariate tiles arrive by one bulk tensor copy each
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(X); CS_BS(z); CS_BS(w); CS_BS(params); CS_BS(G);
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
  // lower-triangular tile enumeration
  const int t = (int)blockIdx.x + tile0;
  const int I = csg_lower_row(t);   // integer / single-precision decode: the FP64 pipe belongs to the tile arithmetic
  const int J = t - I * (I + 1) / 2;
  const int r0 = I * TB, c0 = J * TB;
#ifndef CS_EMU
  __shared__ unsigned long long xbar;
  if (xt.on) {
    if (tid == 0) {
      csd::mbar_init(&xbar, 1);
      csd::mbar_fence_init();
      csd::mbar_expect_tx(&xbar, 2u * (unsigned)p * TB * (unsigned)sizeof(double));
      csd::tma_load_3d(xr, &xt.map, &xbar, r0, 0, (int)blockIdx.z);
      csd::tma_load_3d(xc, &xt.map, &xbar, c0, 0, (int)blockIdx.z);
    }
  } else
#endif
  {
    load_x_tile(xr, X, ldx, r0, p, tid);
    load_x_tile(xc, X, ldx, c0, p, tid);
  }
  if (tid < p) { ea[tid] = exp(-params[lay.i_Lm() + tid]); eb[tid] = exp(-params[lay.i_La() + tid]); }
  if (tid < TB) { zr[tid] = z[r0 + tid]; zc[tid] = z[c0 + tid]; wr[tid] = w[r0 + tid]; }
  __syncthreads();
#ifndef CS_EMU
  if (xt.on) csd::mbar_wait(&xbar, 0);
#endif
  const double lam_m = params[lay.i_lm()], lam_a = params[lay.i_la()];
  const double esig = exp(params[lay.i_sigma()]);
  double sm[4][4], sa[4][4];
  sqdist_4x4(xr, xc, ea, eb, p, tx, ty, sm, sa);
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
        const double km = exp(lam_m - sm[a][b]);
        const double ka = exp(lam_a - sa[a][b]) * zr[rl] * zc[cl];
        v = km + ka;
        if (gr == gc) v += esig / wr[rl];
      }
      G[(long long)gc * ldg + gr] = v;
    }
  }
}

// ---------------------------------------------------------------------------------
// cross Gram (kernmat_*_SE_cpp with X1 != X2): K, Km, Ka (any may be null), n1 x n2,
// rows padded to ld; entries outside (n1,n2) are written as 0.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(ETH)
gram_cross_kernel(const double* __restrict__ X1, long long ldx1, const double* __restrict__ z1, int n1,
                  const double* __restrict__ X2, long long ldx2, const double* __restrict__ z2, int n2,
                  const double* __restrict__ params, Layout lay, double* __restrict__ K, double* __restrict__ Km,
                  double* __restrict__ Ka, long long ldk) {
  CS_DYN_SMEM(double, sm_);
  const int p = lay.p;
  double* xr = sm_;
  double* xc = xr + p * TB;
  double* ea = xc + p * TB;
  double* eb = ea + PMAX;
  double* zr = eb + PMAX;
  double* zc = zr + TB;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = (int)blockIdx.x * TB, c0 = (int)blockIdx.y * TB;
  load_x_tile(xr, X1, ldx1, r0, p, tid);
  load_x_tile(xc, X2, ldx2, c0, p, tid);
  if (tid < p) { ea[tid] = exp(-params[lay.i_Lm() + tid]); eb[tid] = exp(-params[lay.i_La() + tid]); }
  if (tid < TB) { zr[tid] = z1[r0 + tid]; zc[tid] = z2[c0 + tid]; }
  __syncthreads();
  const double lam_m = params[lay.i_lm()], lam_a = params[lay.i_la()];
  double sm[4][4], sa[4][4];
  sqdist_4x4(xr, xc, ea, eb, p, tx, ty, sm, sa);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int cl = ty + 16 * b, gc = c0 + cl;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int rl = tx + 16 * a, gr = r0 + rl;
      double km = 0.0, ka = 0.0;
      if (gr < n1 && gc < n2) {
        km = exp(lam_m - sm[a][b]);
        ka = exp(lam_a - sa[a][b]) * zr[rl] * zc[cl];
      }
      const long long o = (long long)gc * ldk + gr;
      if (K) K[o] = km + ka;
      if (Km) Km[o] = km;
      if (Ka) Ka[o] = ka;
    }
  }
}

// ---------------------------------------------------------------------------------
// Mat-vec panel: out_part[split][k][r] = sum_{c in split} A[r,c] * v_k[c], k < NV.
// A is rows x cols column-major (ld), rows/cols multiples of 64, 16-byte aligned; CTA = 64 rows as 32 row pairs
// (one 16-byte load each) x 8 column lanes.
// ---------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ETH)
matvec_kernel(const double* __restrict__ A, long long ld, int cols_per_split, int cols,
              const double* __restrict__ v0, const double* __restrict__ v1, const double* __restrict__ v2,
              const double* __restrict__ shift0, double* __restrict__ out_part, long long out_stride_k,
              long long out_stride_split, const int* __restrict__ skip, long long bstride) {
  skip = csd::bshift(skip, bstride);
  if (skip && *skip) return;
  A = csd::bshift(A, bstride); v0 = csd::bshift(v0, bstride); v1 = csd::bshift(v1, bstride); v2 = csd::bshift(v2, bstride);
  shift0 = csd::bshift(shift0, bstride); out_part = csd::bshift(out_part, bstride);
  constexpr int CL = 8;  // column lanes
  __shared__ double red[3][CL][TB];
  const double sh = shift0 ? *shift0 : 0.0;  // v0 is used as (v0 - *shift0): ybar = y - mu
  const int tid = threadIdx.x, rp = tid & 31, cl = tid >> 5;
  const double* Ar = A + (int)blockIdx.x * TB + 2 * rp;
  const int cbeg = (int)blockIdx.y * cols_per_split;
  int cend = cbeg + cols_per_split;
  if (cend > cols) cend = cols;
  double s0x = 0.0, s0y = 0.0, s1x = 0.0, s1y = 0.0, s2x = 0.0, s2y = 0.0;
  auto use = [&](double2 a, int cc) {
    const double w0 = v0[cc] - sh;
    s0x = fma(a.x, w0, s0x); s0y = fma(a.y, w0, s0y);
    if (NV > 1) { const double w1 = v1[cc]; s1x = fma(a.x, w1, s1x); s1y = fma(a.y, w1, s1y); }
    if (NV > 2) { const double w2 = v2[cc]; s2x = fma(a.x, w2, s2x); s2y = fma(a.y, w2, s2y); }
  };
  int c = cbeg + cl;
  // four 16-byte matrix loads are requested before the first one is used: with one load in flight per thread the
  // pass is bound by HBM latency, not bandwidth
  for (; c + 3 * CL < cend; c += 4 * CL) {
    double2 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = csd::ld_stream2(Ar + (long long)(c + CL * u) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) use(a[u], c + CL * u);
  }
  for (; c < cend; c += CL) use(csd::ld_stream2(Ar + (long long)c * ld), c);
  red[0][cl][2 * rp] = s0x; red[0][cl][2 * rp + 1] = s0y;
  if (NV > 1) { red[1][cl][2 * rp] = s1x; red[1][cl][2 * rp + 1] = s1y; }
  if (NV > 2) { red[2][cl][2 * rp] = s2x; red[2][cl][2 * rp + 1] = s2y; }
  __syncthreads();
  if (tid < TB) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double s = ((red[k][0][tid] + red[k][1][tid]) + (red[k][2][tid] + red[k][3][tid])) +
                       ((red[k][4][tid] + red[k][5][tid]) + (red[k][6][tid] + red[k][7][tid]));
      out_part[(long long)blockIdx.y * out_stride_split + k * out_stride_k + (int)blockIdx.x * TB + tid] = s;
    }
  }
}

// ---------------------------------------------------------------------------------
// After the K^-1 mat-vec: alpha = K^-1 (y - mu), u = K^-1 y, s = K^-1 1 (reduced over
// splits), then the scalars M, sums, TP coefficient, log det and the closed-form mu.
// One CTA.
// ---------------------------------------------------------------------------------
constexpr int SCALARS_THREADS = 1024;  // one CTA; the loop is latency-bound (strided partials), so use every warp slot
__global__ void __launch_bounds__(SCALARS_THREADS)
scalars_kernel(const double* __restrict__ mv_part, int nsplit, long long stride_k, long long stride_split, int n,
               int npad, const double* __restrict__ y, const double* __restrict__ params, Layout lay,
               const double* __restrict__ logdet_part, int nblk, double* __restrict__ alpha,
               double* __restrict__ u, double* __restrict__ s, EvalScalars* __restrict__ sc,
               const int* __restrict__ skip) {
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(mv_part); CS_BS(y); CS_BS(params); CS_BS(logdet_part); CS_BS(alpha); CS_BS(u); CS_BS(s); CS_BS(sc);
  __shared__ double red[32];
  const int tid = threadIdx.x;
  const double mu = params[lay.i_mu()];
  double M = 0.0, sa = 0.0, su = 0.0, ss = 0.0;
  for (int r = tid; r < npad; r += (int)blockDim.x) {
    double a = 0.0, uu = 0.0, s1 = 0.0;
    if (r < n) {
      for (int k = 0; k < nsplit; ++k) {
        const double* base = mv_part + (long long)k * stride_split + r;
        a += base[0];
        uu += base[stride_k];
        s1 += base[2 * stride_k];
      }
      M = fma(y[r] - mu, a, M);
      sa += a; su += uu; ss += s1;
    }
    alpha[r] = a; u[r] = uu; s[r] = s1;
  }
  M = csd::block_sum(M, red);
  sa = csd::block_sum(sa, red);
  su = csd::block_sum(su, red);
  ss = csd::block_sum(ss, red);
  double ld = 0.0;
  for (int b = tid; b < nblk; b += (int)blockDim.x) ld += logdet_part[b];
  ld = csd::block_sum(ld, red);
  if (tid == 0) {
    sc->M = M; sc->sum_alpha = sa; sc->sum_u = su; sc->sum_s = ss;
    sc->logdet = 2.0 * ld;
    sc->mu_sol = 0.5 * su / ss;  // kernel_GP_SE_cpp.cpp:64
    double coef = 1.0;
    if (lay.kind == KIND_TP) {
      const double nu = exp(params[lay.i_nu()]);
      const double lambda0 = exp(params[lay.i_la()]);
      coef = (nu + (double)n) / ((nu + M) * lambda0);  // kernel_TP_SE_cpp.cpp:122
    }
    sc->coef = coef;
  }
}

// ---------------------------------------------------------------------------------
// K3/K4: fused trace gradients.  T = K^-1 - coef * alpha alpha^T is formed per entry,
// Km / Ka / the per-dimension d^2 are recomputed from X, nothing n x n is materialised.
// Persistent CTAs walk the lower 64x64 tiles (off-diagonal tiles weigh 2).  Per CTA:
//   gpart[cta][0]        = sum T .* Km
//   gpart[cta][1]        = sum T .* Ka
//   gpart[cta][2+i]      = sum T .* Km .* d_i^2
//   gpart[cta][2+p+i]    = sum T .* Ka .* d_i^2
// and K*alpha row partials go to kal_part[slot][row] (each slot/row written exactly once).
// ---------------------------------------------------------------------------------
// prefetch: a second pair of X tiles (TMA double buffering); dw: dimensions whose trace sums this launch accumulates
// (the per-thread accumulators are the large term: 2 KB per accumulated quantity)
inline int grad_smem_bytes(int p, bool prefetch = false, int dw = -1) {
  if (dw < 0) dw = p;
  return ((2 * dw + 2) * ETH + (prefetch ? 4 : 2) * p * TB + 2 * PMAX + 5 * TB + 8 * TB + 32) * (int)sizeof(double);
}
inline bool grad_prefetch_fits(int p) { return grad_smem_bytes(p, true) <= SMEM_LIMIT; }
// largest dimension window that fits beside the X tiles of all p dimensions (p > ~43 needs several launches: every launch
// recomputes the exponents over all p dimensions and accumulates the 2 dw trace sums of its window)
inline int grad_window(int p) {
  int dw = p;
  while (dw > 1 && grad_smem_bytes(p, false, dw) > SMEM_LIMIT) --dw;
  return dw;
}

__host__ __device__ inline void lower_tile_of(int t, int& I, int& J) {
  I = csg_lower_row(t);
  J = t - I * (I + 1) / 2;
}

template <bool FROM_MEM>
__global__ void __launch_bounds__(ETH)
grad_trace_kernel(const double* __restrict__ X, long long ldx, const double* __restrict__ z,
                  const double* __restrict__ params, Layout lay, int n, int nblk,
                  const double* __restrict__ Cinv, long long ldc, const double* __restrict__ alpha,
                  const EvalScalars* __restrict__ sc, double* __restrict__ gpart, double* __restrict__ kal_part,
                  long long kal_stride, const double* __restrict__ Kmat_in, const double* __restrict__ Km_in,
                  const double* __restrict__ Ka_in, const int* __restrict__ skip,
                  const int* __restrict__ tile_tab, const __grid_constant__ rt::XTma xt, int d0, int dw) {
  // [d0, d0 + dw): the dimensions whose trace sums this launch accumulates (gpart keeps the global 2p+2 layout; entries 0, 1
  // and the K alpha partials are written by every launch with identical values)
  // xt.on (never together with tile_tab): the X tiles of the NEXT tile of this persistent CTA are fetched by TMA into a
  // second pair of buffers while the current tile is being processed (grad_smem_bytes(p, true))
  // tile_tab != null (distributed mode): nblk is the number of table entries
  // (I64, J64, coff_lo, coff_hi); Cinv + coff is the local 64 x 64 tile and K*alpha is
  // accumulated into the vector kal_part with atomics (kal_stride is ignored).
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(X); CS_BS(z); CS_BS(params); CS_BS(Cinv); CS_BS(alpha); CS_BS(sc); CS_BS(gpart); CS_BS(kal_part);
  CS_DYN_SMEM(double, sm_);
  const int p = lay.p;
  const int nacc = 2 * dw + 2;          // accumulated here; the global layout has 2p + 2 entries per CTA
  const int nacc_all = 2 * p + 2;
  double* acc = sm_;                    // [nacc][ETH]
  double* xr = acc + nacc * ETH;        // [p][TB]
  double* xc = xr + p * TB;
  double* ea = xc + p * TB;
  double* eb = ea + PMAX;
  double* zr = eb + PMAX;
  double* zc = zr + TB;
  double* ar = zc + TB;                 // alpha rows
  double* ac = ar + TB;                 // alpha cols
  double* csum = ac + TB;               // [TB] column sums of the current tile
  double* rred = csum + TB;             // [8][TB]
  double* red = rred + 8 * TB;          // [32]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, wid = tid >> 5;
#ifndef CS_EMU
  double* const xbase = xr;                                  // buffer b: row tile at xbase + b * xoff, column tile p*TB behind it
  const long long xoff = (red + 32) - xr;                   // (all 128-byte aligned)
  __shared__ unsigned long long xbar[2];
  const bool tma = xt.on && !tile_tab;
  if (tma && tid == 0) {
    csd::mbar_init(&xbar[0], 1);
    csd::mbar_init(&xbar[1], 1);
    csd::mbar_fence_init();
    const int t0 = (int)blockIdx.x;
    if (t0 < nblk * (nblk + 1) / 2) {
      int I0, J0;
      lower_tile_of(t0, I0, J0);
      csd::mbar_expect_tx(&xbar[0], 2u * (unsigned)p * TB * (unsigned)sizeof(double));
      csd::tma_load_3d(xbase, &xt.map, &xbar[0], I0 * TB, 0, (int)blockIdx.z);
      csd::tma_load_3d(xbase + p * TB, &xt.map, &xbar[0], J0 * TB, 0, (int)blockIdx.z);
    }
  }
  int local_tile = 0;
#endif
  for (int k = 0; k < nacc; ++k) acc[k * ETH + tid] = 0.0;
  if (tid < p) { ea[tid] = exp(-params[lay.i_Lm() + tid]); eb[tid] = exp(-params[lay.i_La() + tid]); }
  const double lam_m = params[lay.i_lm()], lam_a = params[lay.i_la()];
  const double coef = sc->coef;
  const int ntiles = tile_tab ? nblk : nblk * (nblk + 1) / 2;

  for (int t = (int)blockIdx.x; t < ntiles; t += (int)gridDim.x) {
    int I, J;
    const double* Ct;
    if (tile_tab) {
      I = tile_tab[4 * t]; J = tile_tab[4 * t + 1];
      const long long coff = (long long)(unsigned)tile_tab[4 * t + 2] | ((long long)tile_tab[4 * t + 3] << 32);
      Ct = Cinv + coff;
    } else {
      lower_tile_of(t, I, J);
      Ct = Cinv + (long long)J * TB * ldc + (long long)I * TB;
    }
    const int r0 = I * TB, c0 = J * TB;
    const double wgt = (I == J) ? 1.0 : 2.0;
    __syncthreads();  // previous tile fully consumed
#ifndef CS_EMU
    if (tma) {
      const int b = local_tile & 1;
      if (tid == 0) {  // the buffers of the previous tile are free now: refill them with the next tile of this CTA
        const int tn = t + (int)gridDim.x;
        if (tn < ntiles) {
          int In, Jn;
          lower_tile_of(tn, In, Jn);
          csd::fence_proxy_async();
          csd::mbar_expect_tx(&xbar[b ^ 1], 2u * (unsigned)p * TB * (unsigned)sizeof(double));
          double* nxt = xbase + (b ^ 1) * xoff;
          csd::tma_load_3d(nxt, &xt.map, &xbar[b ^ 1], In * TB, 0, (int)blockIdx.z);
          csd::tma_load_3d(nxt + p * TB, &xt.map, &xbar[b ^ 1], Jn * TB, 0, (int)blockIdx.z);
        }
      }
      xr = xbase + b * xoff; xc = xr + p * TB;
    } else
#endif
    {
      load_x_tile(xr, X, ldx, r0, p, tid);
      load_x_tile(xc, X, ldx, c0, p, tid);
    }
    if (tid < TB) { zr[tid] = z[r0 + tid]; zc[tid] = z[c0 + tid]; ar[tid] = alpha[r0 + tid]; ac[tid] = alpha[c0 + tid]; }
    __syncthreads();
#ifndef CS_EMU
    if (tma) { csd::mbar_wait(&xbar[local_tile & 1], (unsigned)((local_tile >> 1) & 1)); ++local_tile; }
#endif
    // the K^-1 entries of this tile are requested first so that their HBM latency hides behind the
    // distance / exponent arithmetic
    double tv[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int rl = tx + 16 * a, cl = ty + 16 * b;
        tv[a][b] = (r0 + rl < n && c0 + cl < n) ? Ct[(long long)cl * ldc + rl] : 0.0;
      }
    double km[4][4], ka[4][4];
    if (!FROM_MEM) sqdist_4x4(xr, xc, ea, eb, p, tx, ty, km, ka);
    double rs[4] = {0.0, 0.0, 0.0, 0.0}, cs[4] = {0.0, 0.0, 0.0, 0.0};
    double tm = 0.0, ta = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int cl = ty + 16 * b, gc = c0 + cl;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int rl = tx + 16 * a, gr = r0 + rl;
        double vkm = 0.0, vka = 0.0, T = 0.0, kfull = 0.0;
        if (gr < n && gc < n) {
          if (FROM_MEM) {  // matrices exactly as passed to grad_*_SE_cpp
            vkm = Km_in[(long long)gc * ldc + gr];
            vka = Ka_in[(long long)gc * ldc + gr];
            kfull = Kmat_in[(long long)gc * ldc + gr];
          } else {
            vkm = exp(lam_m - km[a][b]);
            vka = exp(lam_a - ka[a][b]) * zr[rl] * zc[cl];
            kfull = vkm + vka;
          }
          T = tv[a][b] - coef * ar[rl] * ac[cl];
        }
        rs[a] = fma(kfull, ac[cl], rs[a]);
        cs[b] = fma(kfull, ar[rl], cs[b]);
        km[a][b] = T * vkm;  // T .* Km
        ka[a][b] = T * vka;  // T .* Ka
        tm += km[a][b];
        ta += ka[a][b];
      }
    }
    acc[0 * ETH + tid] += wgt * tm;
    acc[1 * ETH + tid] += wgt * ta;
    // two dimensions per trip: with two warps per scheduler the FP64 latency is only hidden by independent work
    // inside the warp (eight partial sums per kernel in flight)
    auto one_dim = [&](int i, double& om, double& oa) {
      double xa[4], xb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = xr[i * TB + tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) xb[b] = xc[i * TB + ty + 16 * b];
      double gm[4] = {0.0, 0.0, 0.0, 0.0}, ga[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double d = xa[a] - xb[b];
          const double d2 = d * d;
          gm[b] = fma(km[a][b], d2, gm[b]);
          ga[b] = fma(ka[a][b], d2, ga[b]);
        }
      om = (gm[0] + gm[1]) + (gm[2] + gm[3]);
      oa = (ga[0] + ga[1]) + (ga[2] + ga[3]);
    };
    int i = 0;
    for (; i + 1 < dw; i += 2) {
      double m0, a0, m1, a1;
      one_dim(d0 + i, m0, a0);
      one_dim(d0 + i + 1, m1, a1);
      acc[(2 + i) * ETH + tid] += wgt * m0;
      acc[(2 + dw + i) * ETH + tid] += wgt * a0;
      acc[(3 + i) * ETH + tid] += wgt * m1;
      acc[(3 + dw + i) * ETH + tid] += wgt * a1;
    }
    if (i < dw) {
      double m0, a0;
      one_dim(d0 + i, m0, a0);
      acc[(2 + i) * ETH + tid] += wgt * m0;
      acc[(2 + dw + i) * ETH + tid] += wgt * a0;
    }
    // K*alpha partials: rows (reduce over ty) and columns (reduce over tx)
#pragma unroll
    for (int a = 0; a < 4; ++a) rs[a] += __shfl_xor_sync(0xffffffffu, rs[a], 16);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double v = cs[b];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      cs[b] = v;
    }
    if (lane < 16) {
#pragma unroll
      for (int a = 0; a < 4; ++a) rred[wid * TB + tx + 16 * a] = rs[a];
    }
    if (tx == 0) {
#pragma unroll
      for (int b = 0; b < 4; ++b) csum[ty + 16 * b] = cs[b];
    }
    __syncthreads();
    if (tid < TB) {
      double sacc = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) sacc += rred[w8 * TB + tid];
      if (tile_tab) {
        atomicAdd(kal_part + r0 + tid, sacc);
        if (I != J) atomicAdd(kal_part + c0 + tid, csum[tid]);
      } else {
        kal_part[(long long)J * kal_stride + r0 + tid] = sacc;
        if (I != J) kal_part[(long long)I * kal_stride + c0 + tid] = csum[tid];
      }
    }
  }
  __syncthreads();
  for (int k = 0; k < nacc; ++k) {
    const double v = csd::block_sum(acc[k * ETH + tid], red);
    const int kg = k < 2 ? k : (k < 2 + dw ? 2 + d0 + (k - 2) : 2 + p + d0 + (k - 2 - dw));
    if (tid == 0) gpart[(long long)blockIdx.x * nacc_all + kg] = v;
  }
}

// ---------------------------------------------------------------------------------
// Row statistics (multi-CTA): K*alpha from its per-tile partials, squared residual
// ||y - (K alpha + mu)||^2 (:192) and the sigma trace sum_i T_ii exp(sigma)/w_i (:162-163),
// as per-CTA partials rs_part[cta][2] (deterministic two-stage reduction).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(ETH)
rowstats_kernel(const double* __restrict__ kal_part, long long kal_stride, int nblk,
                const double* __restrict__ Cinv, long long ldc, const double* __restrict__ alpha,
                const double* __restrict__ y, const double* __restrict__ w, const double* __restrict__ params,
                Layout lay, int n, const EvalScalars* __restrict__ sc, double* __restrict__ rs_part,
                const int* __restrict__ skip) {
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(kal_part); CS_BS(Cinv); CS_BS(alpha); CS_BS(y); CS_BS(w); CS_BS(params); CS_BS(sc); CS_BS(rs_part);
  __shared__ double red[32];
  const int r = (int)blockIdx.x * ETH + (int)threadIdx.x;
  const double coef = sc->coef;
  const double mu = params[lay.i_mu()];
  const double esig = exp(params[lay.i_sigma()]);
  double tsig = 0.0, res = 0.0;
  if (r < n) {
    const double T = Cinv[(long long)r * ldc + r] - coef * alpha[r] * alpha[r];
    tsig = T * (esig / w[r]);
    double ka = 0.0;
    for (int b = 0; b < nblk; ++b) ka += kal_part[(long long)b * kal_stride + r];
    const double e = y[r] - (ka + mu);
    res = e * e;
  }
  tsig = csd::block_sum(tsig, red);
  res = csd::block_sum(res, red);
  if (threadIdx.x == 0) { rs_part[2 * blockIdx.x] = tsig; rs_part[2 * blockIdx.x + 1] = res; }
}

// ---------------------------------------------------------------------------------
// Finalize: gradients in parameter order, stats = (squared residual, LML), mu solution.
// One CTA.  grad_*_SE_cpp :161-198 / TP :127-159.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(ETH)
finalize_kernel(const double* __restrict__ gpart, int ncta, const double* __restrict__ rs_part, int nrs,
                const double* __restrict__ params, Layout lay, int n, EvalScalars* __restrict__ sc,
                double* __restrict__ grad, const int* __restrict__ skip) {
  CS_BS(skip);
  if (skip && *skip) return;
  CS_BS(gpart); CS_BS(rs_part); CS_BS(params); CS_BS(sc); CS_BS(grad);
  __shared__ double red[32];
  __shared__ double gsum[2 * PMAX + 2];
  const int tid = threadIdx.x;
  const int p = lay.p, nacc = 2 * p + 2;
  // deterministic reduction over CTAs
  for (int k = tid; k < nacc; k += ETH) {
    double s = 0.0;
    for (int c = 0; c < ncta; ++c) s += gpart[(long long)c * nacc + k];
    gsum[k] = s;
  }
  const double coef = sc->coef;
  double tsig = 0.0, res = 0.0;
  for (int c = tid; c < nrs; c += ETH) { tsig += rs_part[2 * c]; res += rs_part[2 * c + 1]; }
  tsig = csd::block_sum(tsig, red);
  res = csd::block_sum(res, red);
  __syncthreads();
  if (tid < p) {
    grad[lay.i_Lm() + tid] = -0.5 * exp(-params[lay.i_Lm() + tid]) * gsum[2 + tid];
    grad[lay.i_La() + tid] = -0.5 * exp(-params[lay.i_La() + tid]) * gsum[2 + p + tid];
  }
  if (tid == 0) {
    grad[lay.i_sigma()] = -0.5 * tsig;
    grad[lay.i_lm()] = -0.5 * gsum[0];
    grad[lay.i_la()] = -0.5 * gsum[1];
    grad[lay.i_mu()] = coef * sc->sum_alpha;
    const double dn = (double)n;
    double lml;
    if (lay.kind == KIND_TP) {
      grad[lay.i_nu()] = 0.0;
      const double nu = exp(params[lay.i_nu()]);
      lml = lgamma(0.5 * (nu + dn)) - lgamma(0.5 * nu) -
            0.5 * (dn * log(nu * 3.14159265358979323846) + sc->logdet + (nu + dn) * log(1.0 + sc->M / nu));
    } else {
      lml = -0.5 * (dn * log(2.0 * 3.14159265358979323846) + sc->logdet + sc->M);
    }
    sc->stats[0] = res;
    sc->stats[1] = lml;
  }
}

// ---------------------------------------------------------------------------------
// K5: optimizer step on the flat vectors (one warp), closed-form mu overwrite, trace
// append and the device-side stopping rule of R/main.R:79-84.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void opt_update(int optim, int i, double iter, double lr, double b1, double b2,
                                           double eps, double momentum, double* m, double* v,
                                           const double* grad, double* para) {
  const double g = grad[i];
  if (optim == OPT_NESTEROV) {
    const double nu_old = m[i];
    m[i] = momentum * nu_old + lr * g;  // optimizer_cpp.cpp:19
    para[i] = para[i] + nu_old;         // :20 (old velocity, quirk Q3)
  } else {
    const double mn = b1 * m[i] + (1.0 - b1) * g;      // :39 / :57
    const double vn = b2 * v[i] + (1.0 - b2) * g * g;  // :40 / :58
    m[i] = mn; v[i] = vn;
    const double num = (optim == OPT_NADAM) ? (b1 * mn + (1.0 - b1) * g) : mn;
    para[i] = para[i] + lr * (num / (1.0 - pow(b1, iter))) / (sqrt(vn / (1.0 - pow(b2, iter))) + eps);  // :41 / :59
  }
}

__global__ void opt_step_kernel(int optim, int nparam, double iter, double lr, double b1, double b2, double eps,
                                double momentum, double* m, double* v, const double* grad, double* para) {
  for (int i = threadIdx.x; i < nparam; i += blockDim.x)
    opt_update(optim, i, iter, lr, b1, b2, eps, momentum, m, v, grad, para);
}

__global__ void loop_step_kernel(int optim, Layout lay, double lr, double b1, double b2, double eps, double momentum,
                                 double* m, double* v, const double* grad, double* para, EvalScalars* sc,
                                 LoopState* ls, double* trace, const int* status) {
  CS_BS(ls);
  if (ls->done) return;
  CS_BS(m); CS_BS(v); CS_BS(grad); CS_BS(para); CS_BS(sc); CS_BS(trace); CS_BS(status);
  const int it = ls->it;
  // A member whose K_n was not positive definite stops HERE, before the optimizer step, with the failing pivot (the
  // reference throws inside getinv_kernel, R/kernel_GP_SE_class.R:42); the other members of the batch keep running.
  // |dLML| = NaN stops it too: `if (change < tol && iter > 3)` is an R error on NA (R/main.R:82).
  const int stt = status ? *status : 0x7fffffff;
  const double change_now = fabs(sc->stats[1] - ls->prev_lml);
  if ((stt != 0x7fffffff && stt > 0) || change_now != change_now) {
    if (threadIdx.x == 0) {
      ls->err = (stt != 0x7fffffff && stt > 0) ? stt : LOOP_ERR_NONFINITE;
      ls->iters = it;
      ls->done = 1;
    }
    return;
  }
  const int nparam = lay.np();
  for (int i = threadIdx.x; i < nparam; i += blockDim.x)
    opt_update(optim, i, (double)it, lr, b1, b2, eps, momentum, m, v, grad, para);
  __syncthreads();
  if (threadIdx.x == 0) {
    para[lay.i_mu()] = sc->mu_sol;  // mean_solution overrides the optimizer's mu (quirk Q1)
    trace[2 * it + 0] = sc->stats[0];
    trace[2 * it + 1] = sc->stats[1];
    const double change = fabs(sc->stats[1] - ls->prev_lml);
    ls->prev_lml = sc->stats[1];
    int done = 0;
    if (change < ls->tol && it > 3) done = 1;  // R/main.R:82-83
    if (it >= ls->maxiter) done = 1;
    if (done) { ls->iters = it; ls->done = 1; }
    ls->it = it + 1;
  }
}

// Last node of the WHILE-loop body (rt::graph_while_build): the loop goes on while any member of the handle is still
// running.  With it one cs_fit_run is ONE graph launch; the host only waits for the result (R/main.R:79-84 on the device).
#ifndef CS_EMU
__global__ void loop_cond_kernel(const LoopState* __restrict__ ls, long long bstride, int nbatch, cudaGraphConditionalHandle h) {
  int active = 0;
  for (int b = threadIdx.x; b < nbatch; b += 32) {
    const LoopState* l = reinterpret_cast<const LoopState*>(reinterpret_cast<const char*>(ls) + (long long)b * bstride);
    active |= l->done ? 0 : 1;
  }
  active = __any_sync(0xffffffffu, active);
  if (threadIdx.x == 0) cudaGraphSetConditional(h, active ? 1u : 0u);
}
#endif

// out[k] = sum over CTAs of gpart[cta][k]  (distributed mode: feeds the all-reduce)
__global__ void reduce_gpart_kernel(const double* __restrict__ gpart, int ncta, int nacc, double* __restrict__ out) {
  for (int k = threadIdx.x; k < nacc; k += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < ncta; ++c) s += gpart[(long long)c * nacc + k];
    out[k] = s;
  }
}

__global__ void negate_kernel(double* v, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = -v[i];
}

// stats_*_SE (kernel_GP_SE_cpp.cpp:204-227, kernel_TP_SE_cpp.cpp:165-192) from alpha = invK*ybar
// and kal = Kmat*alpha; logdet_part holds log det(invK)/2 per block.  One CTA.
__global__ void __launch_bounds__(ETH)
stats_only_kernel(int kind, int n, const double* __restrict__ y, const double* __restrict__ alpha,
                  const double* __restrict__ kal, const double* __restrict__ logdet_part, int nblk, double mu,
                  double nu, double* __restrict__ stats) {
  __shared__ double red[32];
  double M = 0.0, res = 0.0, ld = 0.0;
  for (int r = threadIdx.x; r < n; r += ETH) {
    M = fma(y[r] - mu, alpha[r], M);
    const double e = y[r] - (kal[r] + mu);
    res = fma(e, e, res);
  }
  for (int b = threadIdx.x; b < nblk; b += ETH) ld += logdet_part[b];
  M = csd::block_sum(M, red);
  res = csd::block_sum(res, red);
  ld = csd::block_sum(ld, red);
  if (threadIdx.x == 0) {
    const double val = 2.0 * ld;  // log det of the INVERSE (:223)
    const double dn = (double)n;
    stats[0] = res;
    if (kind == KIND_TP)
      stats[1] = lgamma(0.5 * (nu + dn)) - lgamma(0.5 * nu) -
                 0.5 * (dn * log(nu * 3.14159265358979323846) - val + (nu + dn) * log(1.0 + M / nu));
    else
      stats[1] = -0.5 * (dn * log(2.0 * 3.14159265358979323846) - val + M);
  }
}

// mu <- closed-form solution (used by the stateless mirror of mu_solution_cpp)
__global__ void mu_solution_kernel(const double* __restrict__ mv_part, int nsplit, long long stride_k,
                                   long long stride_split, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double su = 0.0, ss = 0.0;
  for (int r = threadIdx.x; r < n; r += blockDim.x)
    for (int k = 0; k < nsplit; ++k) {
      su += mv_part[(long long)k * stride_split + r];
      ss += mv_part[(long long)k * stride_split + stride_k + r];
    }
  su = csd::block_sum(su, red);
  ss = csd::block_sum(ss, red);
  if (threadIdx.x == 0) out[0] = 0.5 * su / ss;
}

}  // namespace csk
