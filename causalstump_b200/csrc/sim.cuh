// Simulation driver on the device: the replicate loop of the reference's simulation scripts,
//   /root/reference/test/sec5-3_Hill2011.R:49-66   Hill11_surfaceB (response surface B)
//   :157-176                                        run(i): seed, sample, 10 % test split
//   R/utilities.R:1-33                              norm_variables on the training split
//   R/kernel_GP_SE_class.R:10-20                    parainit (the two runif draws per restart)
//   :104-152                                        update.results (22 metrics per replicate)
// as two kernels over a whole batch of (replicate, restart) units: `sim_generate_kernel` writes the training data of
// every unit straight into the slab of its member of a batched fit handle (nothing crosses PCIe), `sim_metrics_kernel`
// turns the predict / treatment outputs (left on the device) into the metric block.
//
// Random numbers: counter-based Philox-4x32-10 keyed by the run's seed, counter = (index, stream, replicate, tag), so every
// draw is addressable; causalstump_b200/hill.py holds the host mirror of the same counters (R's Mersenne-Twister stream is
// not available either way).  Discrete draws (beta categories, Bernoulli covariates, treatment, the split) are identical
// on host and device; normals and sums agree to rounding (libm vs CUDA exp/log/cos, reduction order).
#pragma once
#include "predict.cuh"

namespace css {

using csk::Layout;

constexpr int SIM_THREADS = 256;
constexpr int NMETRIC = 22;
enum Stream { S_BETA = 0, S_XNORM = 1, S_XBIN = 2, S_Z = 3, S_N0 = 4, S_N1 = 5, S_SPLIT = 6, S_RESTART = 7 };
constexpr unsigned SIM_TAG = 0x48494c4cu;  // "HILL"

struct SimArgs {
  int n, p, ntr, nte, np_fit;
  unsigned long long seed;
  const int* units;                 // [B][2] (replicate, restart)
  const double* Xshared;            // n x p raw covariates shared by every replicate (the reference's fixed IHDP X), or null
  const double* zshared;            // n, or null
  double* Xraw;                     // [B][n*p]   raw covariates of the unit (copy of Xshared, or generated)
  double *zraw, *y, *y1, *y0;       // [B][n]
  int *is_test, *train_idx, *test_idx;  // [B][n], [B][ntr], [B][nte]
  double* Xall;                     // [B][n*p]   all rows normalised with the training moments (ld = n)
  double* Xte;                      // [B][nte*p] test rows normalised (ld = nte)
  double* moments;                  // [B][2p+2]  meanX[p], varX[p], meanY, varY
  // member 0 of the batched fit handle; member b is bstride bytes further
  double *fX, *fy, *fz, *fw, *fones, *fparams;
  long long bstride;
  Layout lay;
};

// marginals of the 19 binary IHDP covariates (causalstump_b200/hill.py: _BIN_MEANS; index 7 is `first`, coded {1,2})
__device__ __forceinline__ double bin_mean(int b) {
  const double m[19] = {.51, .09, .52, .36, .27, .22, .36, .46, .14, .96, .59, .96, .14, .14, .16, .08, .07, .13, .16};
  return m[b];
}

__device__ __forceinline__ double sim_uniform(const csp::Philox& ph, unsigned index, unsigned stream, unsigned rep) {
  unsigned o[4];
  ph.block(index, stream, rep, SIM_TAG, o);
  return csp::u01(o[0], o[1]);
}
__device__ __forceinline__ double sim_normal(const csp::Philox& ph, unsigned index, unsigned stream, unsigned rep) {
  unsigned o[4];
  ph.block(index, stream, rep, SIM_TAG, o);
  const double u1 = csp::u01(o[0], o[1]), u2 = csp::u01(o[2], o[3]);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}
// a*b + c with two roundings (the host mirror is NumPy, which does not fuse)
__device__ __forceinline__ double mul_add_unfused(double a, double b, double c) {
#ifdef CS_EMU
  volatile double m = a * b;
  return m + c;
#else
  return __dadd_rn(__dmul_rn(a, b), c);
#endif
}

template <class T>
__device__ __forceinline__ T* unit_ptr(T* base, long long stride) { return base + (long long)blockIdx.x * stride; }

inline int sim_smem_bytes(int n, int p) { return (int)(sizeof(double) * (2 * (size_t)n + 3 * (size_t)p + 8 + 32) + sizeof(int) * (size_t)n); }

__global__ void __launch_bounds__(SIM_THREADS)
sim_generate_kernel(SimArgs a) {
  CS_DYN_SMEM(double, sm);
  const int n = a.n, p = a.p, tid = threadIdx.x;
  double* key = sm;                 // [n]   split keys, later scratch
  double* lin = key + n;            // [n]
  double* beta = lin + n;           // [p+1]
  double* meanX = beta + p + 1;     // [p]
  double* sdX = meanX + p;          // [p]
  double* red = sdX + p + 7;        // [32]
  int* tflag = reinterpret_cast<int*>(red + 32);  // [n]
  const int rep = a.units[2 * blockIdx.x], rs = a.units[2 * blockIdx.x + 1];
  const csp::Philox ph{(unsigned)a.seed, (unsigned)(a.seed >> 32)};
  double* X = unit_ptr(a.Xraw, (long long)n * p);
  double* z = unit_ptr(a.zraw, n);
  double *y = unit_ptr(a.y, n), *y1 = unit_ptr(a.y1, n), *y0 = unit_ptr(a.y0, n);

  // ---- covariates and treatment: the reference's are fixed (obs / Trt of the workspace); synthetic ones are drawn
  // with the IHDP marginals, 6 N(0,1) + 19 Bernoulli columns, z ~ Bernoulli(0.186)
  const int ncont = p < 6 ? p : 6;
  for (int idx = tid; idx < n * p; idx += SIM_THREADS) {
    double v;
    if (a.Xshared) {
      v = a.Xshared[idx];
    } else {
      const int j = idx / n;
      if (j < ncont) {
        v = sim_normal(ph, (unsigned)idx, S_XNORM, (unsigned)rep);
      } else {
        const int b = (j - ncont) % 19;
        v = sim_uniform(ph, (unsigned)idx, S_XBIN, (unsigned)rep) < bin_mean(b) ? 1.0 : 0.0;
        if (j - ncont == 7) v += 1.0;
      }
    }
    X[idx] = v;
  }
  double nz = 0.0;
  for (int i = tid; i < n; i += SIM_THREADS) {
    const double v = a.zshared ? a.zshared[i] : (sim_uniform(ph, (unsigned)i, S_Z, (unsigned)rep) < 0.186 ? 1.0 : 0.0);
    z[i] = v;
    nz += v;
  }
  nz = csd::block_sum(nz, red);
  if (nz == 0.0 && tid == 0) z[0] = 1.0;
  // ---- betaB = sample(c(0,.1,.2,.3,.4), p+1, replace=TRUE, prob=c(.6,.1,.1,.1,.1))          (:55)
  for (int j = tid; j <= p; j += SIM_THREADS) {
    const double u = sim_uniform(ph, (unsigned)j, S_BETA, (unsigned)rep);
    beta[j] = u < 0.6 ? 0.0 : (u < 0.7 ? 0.1 : (u < 0.8 ? 0.2 : (u < 0.9 ? 0.3 : 0.4)));
  }
  __syncthreads();
  // ---- yb0hat = exp([1, X+.5] beta), yb1hat = [1, X+.5] beta - offset, offset = mean((yb1hat - yb0hat)[z==1]) - 4   (:56-59)
  double dsum = 0.0, n1 = 0.0;
  for (int i = tid; i < n; i += SIM_THREADS) {
    double l = beta[0];
    for (int j = 0; j < p; ++j) l = mul_add_unfused(X[(long long)j * n + i] + 0.5, beta[j + 1], l);
    lin[i] = l;
    const double e0 = exp(l);
    y0[i] = e0;
    if (z[i] == 1.0) { dsum += l - e0; n1 += 1.0; }
  }
  dsum = csd::block_sum(dsum, red);
  n1 = csd::block_sum(n1, red);
  const double offset = dsum / n1 - 4.0;
  // ---- YB0 = rnorm(N, yb0hat, 1), YB1 = rnorm(N, yb1hat, 1), YB = z ? YB1 : YB0            (:60-63)
  for (int i = tid; i < n; i += SIM_THREADS) {
    const double m1 = lin[i] - offset;
    y1[i] = m1;
    const double e0 = sim_normal(ph, (unsigned)i, S_N0, (unsigned)rep), e1 = sim_normal(ph, (unsigned)i, S_N1, (unsigned)rep);
    y[i] = z[i] == 1.0 ? m1 + e1 : y0[i] + e0;
    key[i] = sim_uniform(ph, (unsigned)i, S_SPLIT, (unsigned)rep);
  }
  __syncthreads();
  // ---- idx.test = sample.int(n, ceiling(n*0.1)): the nte rows with the smallest keys (a uniform random subset)   (:171)
  int* is_test = unit_ptr(a.is_test, n);
  for (int i = tid; i < n; i += SIM_THREADS) {
    const double k = key[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (key[j] < k || (key[j] == k && j < i)) ? 1 : 0;
    tflag[i] = rank < a.nte ? 1 : 0;
    is_test[i] = tflag[i];
  }
  __syncthreads();
  int* train_idx = unit_ptr(a.train_idx, a.ntr);
  int* test_idx = unit_ptr(a.test_idx, a.nte);
  int* pos = reinterpret_cast<int*>(key);  // position of row i inside its subset (key is free now: 2 ints per double)
  for (int i = tid; i < n; i += SIM_THREADS) {
    int before = 0;
    for (int j = 0; j < i; ++j) before += tflag[j];
    const int ps = tflag[i] ? before : i - before;
    pos[i] = ps;
    if (tflag[i]) test_idx[ps] = i; else train_idx[ps] = i;
  }
  __syncthreads();
  // ---- norm_variables on the training split (R/utilities.R:1-33): a column that is binary or inside [0,1] is left
  // alone, any other is centred by its mean and divided by max|x| (uncentred, quirk Q9); y is divided by sd(y), not centred
  const int lane = tid & 31, wid = tid >> 5;
  for (int j = wid; j < p; j += SIM_THREADS / 32) {
    double s = 0.0, mx = 0.0;
    int bin = 1, unit = 1;
    for (int k = lane; k < a.ntr; k += 32) {
      const double v = X[(long long)j * n + train_idx[k]];
      s += v;
      mx = fmax(mx, fabs(v));
      bin &= (v == 0.0 || v == 1.0) ? 1 : 0;
      unit &= (v <= 1.0 && v >= 0.0) ? 1 : 0;
    }
    s = csd::warp_sum(s);
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      bin &= __shfl_xor_sync(0xffffffffu, bin, o);
      unit &= __shfl_xor_sync(0xffffffffu, unit, o);
    }
    if (lane == 0) {
      const bool scale = !bin && !unit;
      meanX[j] = scale ? s / (double)a.ntr : 0.0;
      sdX[j] = scale ? mx : 1.0;   // sqrt(varX), varX = max|x|^2
    }
  }
  double ys = 0.0;
  for (int k = tid; k < a.ntr; k += SIM_THREADS) ys += y[train_idx[k]];
  const double ymean = csd::block_sum(ys, red) / (double)a.ntr;
  double yq = 0.0;
  for (int k = tid; k < a.ntr; k += SIM_THREADS) { const double d = y[train_idx[k]] - ymean; yq += d * d; }
  const double varY = csd::block_sum(yq, red) / (double)(a.ntr - 1);   // var(y), R's n-1 denominator
  const double sdY = sqrt(varY);
  __syncthreads();
  double* mom = unit_ptr(a.moments, 2 * p + 2);
  for (int j = tid; j < p; j += SIM_THREADS) { mom[j] = meanX[j]; mom[p + j] = sdX[j] * sdX[j]; }
  if (tid == 0) { mom[2 * p] = 0.0; mom[2 * p + 1] = varY; }
  // ---- outputs: normalised covariates of all rows / the test rows, and the member's training data inside the fit slab
  const long long sh = (long long)blockIdx.x * a.bstride;
  double* fX = reinterpret_cast<double*>(reinterpret_cast<char*>(a.fX) + sh);
  double* fy = reinterpret_cast<double*>(reinterpret_cast<char*>(a.fy) + sh);
  double* fz = reinterpret_cast<double*>(reinterpret_cast<char*>(a.fz) + sh);
  double* fw = reinterpret_cast<double*>(reinterpret_cast<char*>(a.fw) + sh);
  double* fones = reinterpret_cast<double*>(reinterpret_cast<char*>(a.fones) + sh);
  double* fpar = reinterpret_cast<double*>(reinterpret_cast<char*>(a.fparams) + sh);
  double* Xall = unit_ptr(a.Xall, (long long)n * p);
  double* Xte = unit_ptr(a.Xte, (long long)a.nte * p);
  for (int idx = tid; idx < n * p; idx += SIM_THREADS) {
    const int j = idx / n, i = idx - j * n;
    const double v = (X[idx] - meanX[j]) / sdX[j];
    Xall[idx] = v;
    if (tflag[i]) Xte[(long long)j * a.nte + pos[i]] = v;
    else fX[(long long)j * a.np_fit + pos[i]] = v;
  }
  double s1 = 0.0;
  for (int i = tid; i < n; i += SIM_THREADS)
    if (!tflag[i]) {
      const double yn = y[i] / sdY;
      fy[pos[i]] = yn;
      fz[pos[i]] = z[i];
      s1 += yn;
    }
  for (int i = tid; i < a.np_fit; i += SIM_THREADS) { fw[i] = 1.0; fones[i] = 1.0; if (i >= a.ntr) { fy[i] = 0.0; fz[i] = 0.0; } }
  // ---- parainit (R/kernel_GP_SE_class.R:10-20): sigma = log(var(y)), lambdam / lambdaa = log(runif(1, .2, 1)),
  // Lm = -0.1, La = 0.1, mu = mean(y) -- on the normalised training y
  const double mu = csd::block_sum(s1, red) / (double)a.ntr;
  __syncthreads();
  double q1 = 0.0;
  for (int k = tid; k < a.ntr; k += SIM_THREADS) { const double d = fy[k] - mu; q1 += d * d; }
  const double varn = csd::block_sum(q1, red) / (double)(a.ntr - 1);
  if (tid == 0) {
    const Layout& L = a.lay;
    fpar[L.i_sigma()] = log(varn);
    fpar[L.i_lm()] = log(0.2 + 0.8 * sim_uniform(ph, (unsigned)(2 * rs), S_RESTART, (unsigned)rep));
    fpar[L.i_la()] = log(0.2 + 0.8 * sim_uniform(ph, (unsigned)(2 * rs + 1), S_RESTART, (unsigned)rep));
    fpar[L.i_mu()] = mu;
  }
  for (int j = tid; j < p; j += SIM_THREADS) { fpar[a.lay.i_Lm() + j] = -0.1; fpar[a.lay.i_La() + j] = 0.1; }
}

// ---------------------------------------------------------------------------------
// update.results (test/sec5-3_Hill2011.R:104-152) for every unit of a batch.  Inputs on the normalised scale as the
// device predict / treatment calls leave them; de-normalisation as R/predict.CausalStump.R:33-36 and R/treatment.R:28-31
// (map * sd(y) + mean(y); intervals varY * (+-1.96 * variance) + map, quirk Q8).  Rows reproduce the script as written,
// including `[z.train=1]` (:122,:124), which selects the FIRST element: PEHE.tt.* = |error of the subset's first unit|.
// ---------------------------------------------------------------------------------
struct MetricArgs {
  int n, p, ntr, nte;
  const double *y, *z, *y1, *y0;          // [B][n]
  const int *train_idx, *test_idx;        // [B][ntr], [B][nte]
  const double* moments;                  // [B][2p+2]
  const double* pred_map;                 // [B][n]    predict(all rows): K_xX alpha + mu
  const double *tr_map, *tr_var;          // [B][ntr]  treatment on the training rows: map, diag(cov)
  const double *te_map, *te_var;          // [B][nte]
  const double* scal;                     // [B][4]    ate_train, ate_var_train, ate_test, ate_var_test
  double* metrics;                        // [B][22]
};

__global__ void __launch_bounds__(SIM_THREADS)
sim_metrics_kernel(MetricArgs a) {
  __shared__ double red[32];
  const int tid = threadIdx.x;
  const int n = a.n;
  const double* y = unit_ptr(a.y, n);
  const double* z = unit_ptr(a.z, n);
  const double* y1 = unit_ptr(a.y1, n);
  const double* y0 = unit_ptr(a.y0, n);
  const double* mom = unit_ptr(a.moments, 2 * a.p + 2);
  const double* pm = unit_ptr(a.pred_map, n);
  const double* scal = unit_ptr(a.scal, 4);
  double* out = unit_ptr(a.metrics, NMETRIC);
  const double meanY = mom[2 * a.p], vy = mom[2 * a.p + 1], sd = sqrt(vy);
  for (int sub = 0; sub < 2; ++sub) {
    const int m = sub == 0 ? a.ntr : a.nte;
    const int* idx = sub == 0 ? unit_ptr(a.train_idx, a.ntr) : unit_ptr(a.test_idx, a.nte);
    const double* tmap = sub == 0 ? unit_ptr(a.tr_map, a.ntr) : unit_ptr(a.te_map, a.nte);
    const double* tvar = sub == 0 ? unit_ptr(a.tr_var, a.ntr) : unit_ptr(a.te_var, a.nte);
    double se = 0.0, st = 0.0, sh = 0.0, st1 = 0.0, sh1 = 0.0, c1 = 0.0, sy = 0.0, sy1 = 0.0, cov = 0.0;
    for (int k = tid; k < m; k += SIM_THREADS) {
      const int i = idx[k];
      const double tau = y1[i] - y0[i];
      const double th = sd * tmap[k];
      const double lo = vy * (-1.96 * tvar[k]) + th, hi = vy * (1.96 * tvar[k]) + th;
      const double e = tau - th;
      const double ry = y[i] - (meanY + sd * pm[i]);
      se += e * e; st += tau; sh += th; sy += ry * ry;
      if (z[i] == 1.0) { st1 += tau; sh1 += th; sy1 += ry * ry; c1 += 1.0; }
      cov += (lo < tau && hi > tau) ? 1.0 : 0.0;
    }
    se = csd::block_sum(se, red); st = csd::block_sum(st, red); sh = csd::block_sum(sh, red);
    st1 = csd::block_sum(st1, red); sh1 = csd::block_sum(sh1, red); c1 = csd::block_sum(c1, red);
    sy = csd::block_sum(sy, red); sy1 = csd::block_sum(sy1, red); cov = csd::block_sum(cov, red);
    if (tid == 0) {
      double* o = out + 11 * sub;
      const double dm = (double)m;
      const int i0 = idx[0];
      const double nanv = nan("");
      o[0] = sqrt(se / dm);                                                   // PEHE.all
      o[1] = fabs((y1[i0] - y0[i0]) - sd * tmap[0]);                          // PEHE.tt  (quirk: first element)
      o[2] = sqrt(sy / dm);                                                   // RMSE.all
      o[3] = c1 > 0.0 ? sqrt(sy1 / c1) : nanv;                                // RMSE.tt
      const double e_ate = st / dm - sh / dm;
      const double e_att = c1 > 0.0 ? st1 / c1 - sh1 / c1 : nanv;
      o[4] = e_ate; o[5] = e_att; o[6] = e_ate * e_ate; o[7] = e_att * e_att; // E.ate, E.att, SE.ate, SE.att
      o[8] = cov / dm;                                                        // coverage.ite
      const double ate = sd * scal[2 * sub], half = 1.96 * sqrt(scal[2 * sub + 1]);   // sqrt(negative) = NaN, as in R
      const double alo = vy * (-half) + ate, ahi = vy * half + ate;
      const double mt = st / dm;
      o[9] = (alo < mt && ahi > mt) ? 1.0 : 0.0;                              // coverage.ate
      o[10] = ahi - alo;                                                      // range.ate
    }
    __syncthreads();
  }
}

// ate = mean(map), ate_var = (sum Ka_xx - c' K^-1 c + jitter * n2) / n_train^2   (R/kernel_GP_SE_class.R:96,101)
__global__ void __launch_bounds__(csk::ETH)
treat_scalars_kernel(const double* __restrict__ map, int n2, const double* __restrict__ h2, double jitter, int n_train,
                     double* __restrict__ out2) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n2; i += csk::ETH) s += map[i];
  s = csd::block_sum(s, red);
  if (threadIdx.x == 0) {
    out2[0] = s / (double)n2;
    const double dn = (double)n_train;
    out2[1] = (h2[1] - h2[0] + jitter * (double)n2) / (dn * dn);
  }
}

}  // namespace css
