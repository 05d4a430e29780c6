// oracle/glue_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// Flat C entry points over the R-side glue of the drop-in (causalstump_b200/r/src/b200_glue.cpp), compiled against
// oracle/shim/Rcpp.h, with the SAME names and argument lists as oracle/ref_driver.cpp (prefix glue_ instead of ref_).
// tests/test_r_glue.py calls both libraries with identical inputs: the reference's own kernels on one side, the glue ->
// C ABI (include/cs_b200.h) -> CUDA kernels (SIMT-emulated on CPU, or the real library on the GPU box) on the other.
#include <Rcpp.h>

#include <cstring>

using namespace Rcpp;

// the glue's exported routines (same signatures as the reference's, with NumericMatrix for arma::mat)
NumericMatrix invkernel_cpp(NumericVector z, NumericVector w, NumericMatrix mymat, List parameters);
double mu_solution_cpp(NumericVector y, NumericVector z, NumericMatrix invKmat, List parameters);
List kernmat_GP_SE_cpp(NumericMatrix X1, NumericMatrix X2, NumericVector Z1, NumericVector Z2, List para);
List kernmat_TP_SE_cpp(NumericMatrix X1, NumericMatrix X2, NumericVector Z1, NumericVector Z2, List para);
List grad_GP_SE_cpp(NumericVector y, NumericMatrix X, NumericVector z, NumericVector w, NumericMatrix Kmat, NumericMatrix Km, NumericMatrix Ka, NumericMatrix invKmatn, List parameters);
List grad_TP_SE_cpp(NumericVector y, NumericMatrix X, NumericVector z, NumericVector w, NumericMatrix Smat, NumericMatrix Sm, NumericMatrix Sa, NumericMatrix invSmatn, List parameters);
NumericVector stats_GP_SE(NumericVector y, NumericMatrix Kmat, NumericMatrix invKmatn, List parameters);
NumericVector stats_TP_SE(NumericVector y, NumericMatrix Kmat, NumericMatrix invKmatn, List parameters);
List Nesterov_cpp(double learn_rate, double momentum, List nu, List grad, List para);
List Nadam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, List m, List v, List grad, List para);
List Adam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, List m, List v, List grad, List para);
SEXP b200_fit_create(int kind, int optim, double lr, double beta1, double beta2, double momentum, NumericVector y, NumericMatrix X,
                     NumericVector z, NumericVector w, List parameters);
List b200_fit_run(SEXP handle, int maxiter, double tol, List parameters);
List b200_treat(SEXP handle, NumericMatrix X2, double jitter, bool want_cov);
List b200_predict(SEXP handle, NumericMatrix X2, NumericVector z2, bool want_cov);
SEXP b200_fit_create_batch(int kind, int optim, double lr, double beta1, double beta2, double momentum, List ys, List Xs, List zs,
                           Nullable<List> ws, List inits);
List b200_fit_run_batch(SEXP handle, int maxiter, double tol);
void b200_fit_select(SEXP handle, int member);
List b200_fit_get_params(SEXP handle, List parameters);

namespace {

NumericVector mknv(const double* p, int n) { return NumericVector(std::vector<double>(p, p + n)); }
NumericMatrix mknm(const double* p, int r, int c) { NumericMatrix m(r, c); std::memcpy(m.d.data(), p, sizeof(double) * (size_t)r * c); return m; }
List mklist(int kind, int p, const double* v) {  // parainit order, as oracle/ref_driver.cpp
  List l;
  int i = 0;
  l["sigma"] = v[i++];
  l["lambdam"] = v[i++];
  if (kind == 0) l["lambdaa"] = v[i++];
  else { l["nu"] = v[i++]; l["lambda0"] = v[i++]; }
  l["Lm"] = std::vector<double>(v + i, v + i + p); i += p;
  l["La"] = std::vector<double>(v + i, v + i + p); i += p;
  l["mu"] = v[i++];
  return l;
}
void flatten(const List& l, double* out) { size_t k = 0; for (const auto& it : l.items) for (double x : it.second.num) out[k++] = x; }
const Value& get(const List& l, const char* name) {
  for (const auto& it : l.items) if (it.first == name) return it.second;
  throw std::runtime_error(std::string("missing list entry ") + name);
}
thread_local std::string g_msg;

}  // namespace

extern "C" {

const char* glue_last_error() { return g_msg.c_str(); }

int glue_kernmat(int kind, int n1, int n2, int p, const double* X1, const double* X2, const double* Z1, const double* Z2,
                 const double* Lm, const double* La, double lambdam, double lambda_a, double* K, double* Km, double* Ka) {
  try {
    List para;
    para["Lm"] = std::vector<double>(Lm, Lm + p);
    para["La"] = std::vector<double>(La, La + p);
    para["lambdam"] = lambdam;
    para[kind == 0 ? "lambdaa" : "lambda0"] = lambda_a;
    List out = kind == 0 ? kernmat_GP_SE_cpp(mknm(X1, n1, p), mknm(X2, n2, p), mknv(Z1, n1), mknv(Z2, n2), para)
                         : kernmat_TP_SE_cpp(mknm(X1, n1, p), mknm(X2, n2, p), mknv(Z1, n1), mknv(Z2, n2), para);
    const size_t nb = sizeof(double) * (size_t)n1 * n2;
    std::memcpy(K, get(out, kind == 0 ? "Kmat" : "Smat").num.data(), nb);
    std::memcpy(Km, get(out, kind == 0 ? "Km" : "Sm").num.data(), nb);
    std::memcpy(Ka, get(out, kind == 0 ? "Ka" : "Sa").num.data(), nb);
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

int glue_invkernel(int n, const double* z, const double* w, const double* K, double sigma, double* invK) {
  try {
    List para; para["sigma"] = sigma;
    NumericMatrix out = invkernel_cpp(mknv(z, n), mknv(w, n), mknm(K, n, n), para);
    std::memcpy(invK, out.d.data(), sizeof(double) * (size_t)n * n);
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return 1; }  // the R condition of a non-PD matrix
}

int glue_mu_solution(int n, const double* y, const double* z, const double* invK, double* mu) {
  try { List para; *mu = mu_solution_cpp(mknv(y, n), mknv(z, n), mknm(invK, n, n), para); return 0; }
  catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

int glue_grad(int kind, int n, int p, const double* y, const double* X, const double* z, const double* w, const double* K,
              const double* Km, const double* Ka, const double* invK, const double* params, double* grad, double* stats) {
  try {
    List para = mklist(kind, p, params);
    List out = kind == 0 ? grad_GP_SE_cpp(mknv(y, n), mknm(X, n, p), mknv(z, n), mknv(w, n), mknm(K, n, n), mknm(Km, n, n), mknm(Ka, n, n), mknm(invK, n, n), para)
                         : grad_TP_SE_cpp(mknv(y, n), mknm(X, n, p), mknv(z, n), mknv(w, n), mknm(K, n, n), mknm(Km, n, n), mknm(Ka, n, n), mknm(invK, n, n), para);
    flatten(*get(out, "gradients").lst, grad);
    const Value& st = get(out, "stats");
    stats[0] = st.num[0]; stats[1] = st.num[1];
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

int glue_stats(int kind, int n, int p, const double* y, const double* K, const double* invK, const double* params, double* stats) {
  try {
    List para = mklist(kind, p, params);
    NumericVector st = kind == 0 ? stats_GP_SE(mknv(y, n), mknm(K, n, n), mknm(invK, n, n), para)
                                 : stats_TP_SE(mknv(y, n), mknm(K, n, n), mknm(invK, n, n), para);
    stats[0] = st[0]; stats[1] = st[1];
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

int glue_opt_step(int optim, int kind, int p, double iter, double lr, double beta1, double beta2, double eps, double* m,
                  double* v, const double* grad, double* para) {
  try {
    List lm = mklist(kind, p, m), lg = mklist(kind, p, grad), lp = mklist(kind, p, para);
    if (optim == 2) {
      List out = Nesterov_cpp(lr, beta1, lm, lg, lp);
      flatten(*get(out, "nu").lst, m);
      flatten(*get(out, "parameters").lst, para);
    } else {
      List lv = mklist(kind, p, v);
      List out = optim == 0 ? Adam_cpp(iter, lr, beta1, beta2, eps, lm, lv, lg, lp) : Nadam_cpp(iter, lr, beta1, beta2, eps, lm, lv, lg, lp);
      flatten(*get(out, "m").lst, m);
      flatten(*get(out, "v").lst, v);
      flatten(*get(out, "parameters").lst, para);
    }
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

// the fast path of INTEGRATION.md step 3: b200_fit_create + b200_fit_run (+ b200_treat) as the R side would call them
int glue_fit_run(int kind, int optim, int n, int p, const double* y, const double* X, const double* z, const double* w,
                 const double* params, double lr, double momentum, int maxiter, double tol, int* iters, double* stats_trace,
                 double* params_out, int n2, const double* X2, double* treat_map, double* ate) {
  try {
    List par = mklist(kind, p, params);
    SEXP h = b200_fit_create(kind, optim, lr, 0.9, 0.999, momentum, mknv(y, n), mknm(X, n, p), mknv(z, n), mknv(w, n), par);
    List res = b200_fit_run(h, maxiter, tol, par);
    *iters = (int)get(res, "iter").num[0];
    std::memcpy(stats_trace, get(res, "stats").num.data(), sizeof(double) * 2 * (maxiter + 2));
    flatten(*get(res, "parameters").lst, params_out);
    if (n2 > 0) {
      List tr = b200_treat(h, mknm(X2, n2, p), kind == 1 ? 1e-6 : 0.0, false);
      std::memcpy(treat_map, get(tr, "map").num.data(), sizeof(double) * n2);
      *ate = get(tr, "ate").num[0];
    }
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

// batched handle as the R side would drive it: B members stored back to back (y: B x n, X: B x n x p column-major, ...)
int glue_fit_run_batch(int kind, int optim, int B, int n, int p, const double* ys, const double* Xs, const double* zs,
                       const double* params, double lr, int maxiter, double tol, int* iters, double* params_out) {
  try {
    const int NP = 2 * p + (kind == 0 ? 4 : 5);
    List ly(B), lX(B), lz(B), li(B);
    for (int b = 0; b < B; ++b) {
      ly[b] = mknv(ys + (size_t)b * n, n);
      lX[b] = mknm(Xs + (size_t)b * n * p, n, p);
      lz[b] = mknv(zs + (size_t)b * n, n);
      li[b] = mklist(kind, p, params + (size_t)b * NP);
    }
    SEXP h = b200_fit_create_batch(kind, optim, lr, 0.9, 0.999, 0.0, ly, lX, lz, Nullable<List>(), li);
    List res = b200_fit_run_batch(h, maxiter, tol);
    const Value& it = get(res, "iter");
    for (int b = 0; b < B; ++b) {
      iters[b] = (int)it.num[b];
      b200_fit_select(h, b);
      flatten(b200_fit_get_params(h, mklist(kind, p, params)), params_out + (size_t)b * NP);
    }
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

}  // extern "C"
