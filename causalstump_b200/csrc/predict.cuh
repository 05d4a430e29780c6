// Predict / treatment cross-kernel path (K6) and the Student-t posterior sampler (K7).
// Reference: R/kernel_GP_SE_class.R:74-105, R/kernel_TP_SE_class.R:72-125 (R `%*%` algebra
// and mvnfast::rmvt + apply(...,sort)).
#pragma once
#include "kernels.cuh"

namespace csp {

using csk::ETH;
using csk::TB;

// out[r] = sum_split part[split][r] (+ *shift)
__global__ void reduce_parts_kernel(const double* __restrict__ part, int nsplit, long long stride, int n,
                                    const double* __restrict__ shift, double* __restrict__ out) {
  const int r = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (r >= n) return;
  double s = 0.0;
  for (int k = 0; k < nsplit; ++k) s += part[(long long)k * stride + r];
  if (shift) s += *shift;
  out[r] = s;
}

// part[split][r] = sum_{c in split} V[r,c] * Kc[r,c]   (rows x cols, both ld)
__global__ void __launch_bounds__(ETH)
rowdot_kernel(const double* __restrict__ V, const double* __restrict__ Kc, long long ld, int cols_per_split,
              int cols, double* __restrict__ part, long long part_stride) {
  __shared__ double red[4][TB];
  const int tid = threadIdx.x, rl = tid & 63, cl = tid >> 6;
  const int r = (int)blockIdx.x * TB + rl;
  const int cbeg = (int)blockIdx.y * cols_per_split;
  int cend = cbeg + cols_per_split;
  if (cend > cols) cend = cols;
  double s = 0.0;
  for (int c = cbeg + cl; c < cend; c += 4) s = fma(V[(long long)c * ld + r], Kc[(long long)c * ld + r], s);
  red[cl][rl] = s;
  __syncthreads();
  if (tid < TB)
    part[(long long)blockIdx.y * part_stride + (int)blockIdx.x * TB + tid] =
        (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
}

// var[i] = kxx_i - q[i] + add, kxx_i = exp(lam_m)*use_m + exp(lam_a) * z2_i^2
__global__ void var_diag_kernel(const double* __restrict__ q, const double* __restrict__ z2,
                                const double* __restrict__ params, csk::Layout lay, int use_m, double add_sigma,
                                double add_const, int n2, double* __restrict__ var) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n2) return;
  const double km = use_m ? exp(params[lay.i_lm()]) : 0.0;
  const double ka = exp(params[lay.i_la()]) * z2[i] * z2[i];
  const double noise = add_sigma != 0.0 ? add_sigma * exp(params[lay.i_sigma()]) : 0.0;
  var[i] = km + ka - q[i] + noise + add_const;
}

// cov[i,i] += add_sigma*exp(sigma) + add_const for i < n2; padded diagonal := 1
__global__ void cov_diag_kernel(double* __restrict__ cov, long long ld, int n2, int n2p,
                                const double* __restrict__ params, csk::Layout lay, double add_sigma,
                                double add_const) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n2p) return;
  if (i < n2) {
    const double noise = add_sigma != 0.0 ? add_sigma * exp(params[lay.i_sigma()]) : 0.0;
    cov[(long long)i * ld + i] += noise + add_const;
  } else {
    cov[(long long)i * ld + i] = 1.0;
  }
}

// c[j] = sum_{i < rows} A[i,j]; one warp per column
__global__ void colsum_kernel(const double* __restrict__ A, long long ld, int rows, int cols,
                              double* __restrict__ out) {
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= cols) return;
  double s = 0.0;
  for (int i = lane; i < rows; i += 32) s += A[(long long)warp * ld + i];
  s = csd::warp_sum(s);
  if (lane == 0) out[warp] = s;
}

// single-CTA helpers
__global__ void __launch_bounds__(ETH)
dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += ETH) s = fma(a[i], b ? b[i] : 1.0, s);
  s = csd::block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

// sum over i,j < n2 of Ka(x_i, x_j) with z = 1 (treatment block), per-CTA partials
__global__ void __launch_bounds__(ETH)
gram_sum_kernel(const double* __restrict__ X2, long long ldx, int n2, const double* __restrict__ params,
                csk::Layout lay, double* __restrict__ part) {
  CS_DYN_SMEM(double, sm_);
  __shared__ double red[32];
  const int p = lay.p;
  double* xr = sm_;
  double* xc = xr + p * TB;
  double* ea = xc + p * TB;
  double* eb = ea + csk::PMAX;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = (int)blockIdx.x * TB, c0 = (int)blockIdx.y * TB;
  csk::load_x_tile(xr, X2, ldx, r0, p, tid);
  csk::load_x_tile(xc, X2, ldx, c0, p, tid);
  if (tid < p) { ea[tid] = exp(-params[lay.i_Lm() + tid]); eb[tid] = exp(-params[lay.i_La() + tid]); }
  __syncthreads();
  const double lam_a = params[lay.i_la()];
  double sm[4][4], sa[4][4];
  csk::sqdist_4x4(xr, xc, ea, eb, p, tx, ty, sm, sa);
  double s = 0.0;
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (r0 + tx + 16 * a < n2 && c0 + ty + 16 * b < n2) s += exp(lam_a - sa[a][b]);
  s = csd::block_sum(s, red);
  if (tid == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = s;
}

// ---------------------------------------------------------------------------------
// K7 sampler pieces: Philox-4x32-10 counter RNG, N(0,1) by Box-Muller, chi^2 by
// Marsaglia-Tsang, bitonic order statistic.
// ---------------------------------------------------------------------------------
struct Philox {
  unsigned k0, k1;
  __host__ __device__ static void mulhilo(unsigned a, unsigned b, unsigned& hi, unsigned& lo) {
    const unsigned long long pr = (unsigned long long)a * b;
    hi = (unsigned)(pr >> 32); lo = (unsigned)pr;
  }
  __host__ __device__ void block(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned (&out)[4]) const {
    unsigned a = k0, b = k1;
    for (int r = 0; r < 10; ++r) {
      unsigned h0, l0, h1, l1;
      mulhilo(0xD2511F53u, c0, h0, l0);
      mulhilo(0xCD9E8D57u, c2, h1, l1);
      const unsigned n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

__host__ __device__ inline double u01(unsigned hi, unsigned lo) {
  // (0,1): 53 random bits, never exactly 0
  const unsigned long long bits = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

// Z[s + k*ldz], s < ldz rows (samples), k < cols: two normals per Philox block
__global__ void normal_fill_kernel(double* __restrict__ Z, long long ldz, int rows, int cols,
                                   unsigned long long seed, unsigned round) {
  const long long npair = ((long long)rows * cols + 1) / 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npair) return;
  Philox ph{(unsigned)seed, (unsigned)(seed >> 32)};
  unsigned o[4];
  ph.block((unsigned)i, (unsigned)(i >> 32), round, 0x5eedu, o);
  const double u1 = u01(o[0], o[1]), u2 = u01(o[2], o[3]);
  const double rad = sqrt(-2.0 * log(u1));
  const double ang = 6.283185307179586476925 * u2;
  const long long e0 = 2 * i, e1 = 2 * i + 1;
  Z[(e0 / rows) * ldz + (e0 % rows)] = rad * cos(ang);
  if (e1 < (long long)rows * cols) Z[(e1 / rows) * ldz + (e1 % rows)] = rad * sin(ang);
}

// rowscale[s] = 1/sqrt(chi2_df/df) ; Marsaglia-Tsang gamma(df/2, 2)
__global__ void chi2_scale_kernel(double* __restrict__ rowscale, int rows, double df, unsigned long long seed,
                                  unsigned round) {
  const int s = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (s >= rows) return;
  Philox ph{(unsigned)seed ^ 0xA5A5A5A5u, (unsigned)(seed >> 32) + 0x1234567u};
  double shape = 0.5 * df, boost = 1.0;
  unsigned ctr = 0;
  unsigned o[4];
  if (shape < 1.0) {
    ph.block((unsigned)s, ctr++, round, 0xC41u, o);
    boost = pow(u01(o[0], o[1]), 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double g = d;
  for (int it = 0; it < 64; ++it) {
    ph.block((unsigned)s, ctr++, round, 0xC41u, o);
    const double u1 = u01(o[0], o[1]), u2 = u01(o[2], o[3]);
    const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
    const double vv = 1.0 + c * x;
    if (vv <= 0.0) continue;
    const double v3 = vv * vv * vv;
    ph.block((unsigned)s, ctr++, round, 0xC41u, o);
    const double uu = u01(o[0], o[1]);
    if (log(uu) < 0.5 * x * x + d - d * v3 + d * log(v3)) { g = d * v3; break; }
  }
  const double chi2 = 2.0 * g * boost;
  rowscale[s] = 1.0 / sqrt(chi2 / df);
}

// One CTA per column (or a single CTA for the row means when `vals` is a vector):
// sorts NS (<= 1024) scaled values ascending and adds weight * value[rank] to out[col].
constexpr int SORT_N = 1024;
__global__ void __launch_bounds__(512)
order_stat_kernel(const double* __restrict__ Y, long long ldy, const double* __restrict__ rowscale, int ns,
                  int rank, double weight, double* __restrict__ out) {
  __shared__ double v[SORT_N];
  const int col = (int)blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < SORT_N; i += 512)
    v[i] = (i < ns) ? Y[(long long)col * ldy + i] * (rowscale ? rowscale[i] : 1.0) : 1.0e300;
  __syncthreads();
  for (int k = 2; k <= SORT_N; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < SORT_N; i += 512) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = ((i & k) == 0);
          const double a = v[i], b = v[ixj];
          if ((a > b) == up) { v[i] = b; v[ixj] = a; }
        }
      }
      __syncthreads();
    }
  if (tid == 0) out[col] += weight * v[rank];
}

// rowmean[s] = mean_{k < cols} Y[s,k]
__global__ void rowmean_kernel(const double* __restrict__ Y, long long ldy, int rows, int cols,
                               double* __restrict__ out) {
  const int s = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (s >= rows) return;
  double acc = 0.0;
  for (int k = 0; k < cols; ++k) acc += Y[(long long)k * ldy + s];
  out[s] = acc / (double)cols;
}

}  // namespace csp
