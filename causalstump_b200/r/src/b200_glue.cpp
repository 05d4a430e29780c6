// Rcpp glue that keeps the reference's 11 registered .Call routines
// (/root/reference/src/RcppExports.cpp:187-200) and routes them to the C ABI of
// libcausalstump_b200.so (include/cs_b200.h).  Replaces src/kernel_GP_SE_cpp.cpp,
// src/kernel_TP_SE_cpp.cpp and src/optimizer_cpp.cpp of the reference package; run
// Rcpp::compileAttributes() to regenerate RcppExports.{R,cpp} (names and arities are
// unchanged, the b200_* routines are added).  The build image has no R / Rcpp headers: this file is
// compiled there against oracle/shim/Rcpp.h and run next to the reference's own routines (tests/test_r_glue.py) --
// see INTEGRATION.md.  Only Rcpp is needed (RcppArmadillo is dropped).
#include <Rcpp.h>
#include "cs_b200.h"
using namespace Rcpp;

static void cs_check(int st) {
  if (st == 0) return;
  // same condition class the reference raises through BEGIN_RCPP/END_RCPP when
  // arma::inv_sympd throws (src/RcppExports.cpp:12,21)
  if (st > 0) stop("inv_sympd(): matrix is singular or not positive definite (pivot %d)", st);
  stop("causalstump_b200: %s", cs_last_error());
}

static bool is_tp(const List& par) { return par.containsElementNamed("lambda0"); }

// flat vector in the reference's list order (scalars, Lm[p], La[p], mu)
static NumericVector pack(const List& par) {
  std::vector<double> v;
  for (int i = 0; i < par.size(); ++i) { NumericVector e = as<NumericVector>(par[i]); v.insert(v.end(), e.begin(), e.end()); }
  return wrap(v);
}
static List unpack(const NumericVector& v, const List& shape) {
  List out = clone(shape);
  int k = 0;
  for (int i = 0; i < shape.size(); ++i) {
    NumericVector e = as<NumericVector>(shape[i]);
    NumericVector r(e.size());
    for (int j = 0; j < e.size(); ++j) r[j] = v[k++];
    out[i] = r;
  }
  return out;
}

static List kernmat(NumericMatrix X1, NumericMatrix X2, NumericVector Z1, NumericVector Z2, List para, bool tp) {
  const int n1 = X1.nrow(), n2 = X2.nrow(), p = X2.ncol();
  NumericMatrix K(n1, n2), Km(n1, n2), Ka(n1, n2);
  NumericVector Lm = para["Lm"], La = para["La"];
  cs_check(cs_kernmat(n1, n2, p, X1.begin(), X2.begin(), Z1.begin(), Z2.begin(), Lm.begin(), La.begin(),
                      as<double>(para["lambdam"]), as<double>(para[tp ? "lambda0" : "lambdaa"]), K.begin(),
                      Km.begin(), Ka.begin()));
  List out;
  if (tp) { out["Smat"] = K; out["Sm"] = Km; out["Sa"] = Ka; } else { out["Kmat"] = K; out["Km"] = Km; out["Ka"] = Ka; }
  return out;
}

// [[Rcpp::export]]
List kernmat_GP_SE_cpp(NumericMatrix X1, NumericMatrix X2, NumericVector Z1, NumericVector Z2, List para) {
  return kernmat(X1, X2, Z1, Z2, para, false);
}
// [[Rcpp::export]]
List kernmat_TP_SE_cpp(NumericMatrix X1, NumericMatrix X2, NumericVector Z1, NumericVector Z2, List para) {
  return kernmat(X1, X2, Z1, Z2, para, true);
}
// [[Rcpp::export]]
NumericMatrix invkernel_cpp(NumericVector z, NumericVector w, NumericMatrix mymat, List parameters) {
  const int n = mymat.nrow();
  NumericMatrix out(n, n);
  cs_check(cs_invkernel(n, w.begin(), mymat.begin(), as<double>(parameters["sigma"]), out.begin(), NULL));
  return out;
}
// [[Rcpp::export]]
double mu_solution_cpp(NumericVector y, NumericVector z, NumericMatrix invKmat, List parameters) {
  double mu = 0.0;
  cs_check(cs_mu_solution(y.size(), y.begin(), invKmat.begin(), &mu));
  return mu;
}
static List grad_common(int kind, NumericVector y, NumericMatrix X, NumericVector z, NumericVector w, NumericMatrix K,
                        NumericMatrix Km, NumericMatrix Ka, NumericMatrix invK, List parameters) {
  NumericVector pv = pack(parameters), g(pv.size()), stats(2);
  cs_check(cs_grad(kind, X.nrow(), X.ncol(), y.begin(), X.begin(), z.begin(), w.begin(), K.begin(), Km.begin(),
                   Ka.begin(), invK.begin(), pv.begin(), g.begin(), stats.begin()));
  return List::create(Named("gradients") = unpack(g, parameters), Named("stats") = stats);
}
// [[Rcpp::export]]
List grad_GP_SE_cpp(NumericVector y, NumericMatrix X, NumericVector z, NumericVector w, NumericMatrix Kmat,
                    NumericMatrix Km, NumericMatrix Ka, NumericMatrix invKmatn, List parameters) {
  return grad_common(CS_GP, y, X, z, w, Kmat, Km, Ka, invKmatn, parameters);
}
// [[Rcpp::export]]
List grad_TP_SE_cpp(NumericVector y, NumericMatrix X, NumericVector z, NumericVector w, NumericMatrix Smat,
                    NumericMatrix Sm, NumericMatrix Sa, NumericMatrix invSmatn, List parameters) {
  return grad_common(CS_TP, y, X, z, w, Smat, Sm, Sa, invSmatn, parameters);
}
// [[Rcpp::export]]
NumericVector stats_GP_SE(NumericVector y, NumericMatrix Kmat, NumericMatrix invKmatn, List parameters) {
  NumericVector st(2);
  cs_check(cs_stats(CS_GP, y.size(), y.begin(), Kmat.begin(), invKmatn.begin(), as<double>(parameters["mu"]), 1.0, st.begin()));
  return st;
}
// [[Rcpp::export]]
NumericVector stats_TP_SE(NumericVector y, NumericMatrix Kmat, NumericMatrix invKmatn, List parameters) {
  NumericVector st(2);
  cs_check(cs_stats(CS_TP, y.size(), y.begin(), Kmat.begin(), invKmatn.begin(), as<double>(parameters["mu"]),
                    std::exp(as<double>(parameters["nu"])), st.begin()));
  return st;
}
static List opt_common(int optim, double iter, double lr, double b1, double b2, double eps, List m, List v, List grad,
                       List para, bool has_v) {
  // positional over the first m.size() entries, like the reference loops (optimizer_cpp.cpp:16,36,54)
  const int k = m.size();
  List psub(k), gsub(k);
  for (int i = 0; i < k; ++i) { psub[i] = para[i]; gsub[i] = grad[i]; }
  NumericVector mv = pack(m), vv = has_v ? pack(v) : NumericVector(mv.size()), gv = pack(gsub), pv = pack(psub);
  cs_check(cs_opt_step(optim, pv.size(), iter, lr, b1, b2, eps, mv.begin(), has_v ? vv.begin() : NULL, gv.begin(), pv.begin()));
  List newp = clone(para), pu = unpack(pv, psub);
  for (int i = 0; i < k; ++i) newp[i] = pu[i];
  if (!has_v) return List::create(Named("nu") = unpack(mv, m), Named("parameters") = newp);
  return List::create(Named("m") = unpack(mv, m), Named("v") = unpack(vv, v), Named("parameters") = newp);
}
// [[Rcpp::export]]
List Nesterov_cpp(double learn_rate, double momentum, List nu, List grad, List para) {
  return opt_common(CS_OPT_NESTEROV, 1.0, learn_rate, momentum, 0.0, 0.0, nu, nu, grad, para, false);
}
// [[Rcpp::export]]
List Nadam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, List m, List v, List grad, List para) {
  return opt_common(CS_OPT_NADAM, iter, learn_rate, beta1, beta2, eps, m, v, grad, para, true);
}
// [[Rcpp::export]]
List Adam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, List m, List v, List grad, List para) {
  return opt_common(CS_OPT_ADAM, iter, learn_rate, beta1, beta2, eps, m, v, grad, para, true);
}

// ---- stateful fast path: the device-resident handle as an R external pointer --------------
static void fit_finalizer(cs_fit* f) { cs_fit_destroy(f); }
typedef XPtr<cs_fit, PreserveStorage, fit_finalizer> FitPtr;

// [[Rcpp::export]]
SEXP b200_fit_create(int kind, int optim, double lr, double beta1, double beta2, double momentum, NumericVector y,
                     NumericMatrix X, NumericVector z, NumericVector w, List parameters) {
  cs_fit_config cfg = {kind, optim, lr, beta1, beta2, 1e-8, momentum, 0, 0, 0, 0, 0, 0};
  NumericVector pv = pack(parameters);
  cs_fit* f = NULL;
  cs_check(cs_fit_create(&cfg, X.nrow(), X.ncol(), y.begin(), X.begin(), z.begin(), w.begin(), pv.begin(), &f));
  return FitPtr(f, true);
}
// [[Rcpp::export]]
List b200_fit_run(SEXP handle, int maxiter, double tol, List parameters) {
  FitPtr f(handle);
  int iters = 0;
  NumericMatrix stats(2, maxiter + 2);
  cs_check(cs_fit_run(f.get(), maxiter, tol, &iters, stats.begin()));
  NumericVector pv(cs_fit_nparam(f.get()));
  cs_check(cs_fit_get_params(f.get(), pv.begin()));
  return List::create(Named("iter") = iters, Named("stats") = stats, Named("parameters") = unpack(pv, parameters));
}
// Batched handle: the replicates x restarts of test/sec5-3_Hill2011.R:611-616 behind one device handle
// (cs_fit_create_batch); ys / Xs / zs / ws / inits are lists of equally shaped members, ws may be NULL.
// [[Rcpp::export]]
SEXP b200_fit_create_batch(int kind, int optim, double lr, double beta1, double beta2, double momentum, List ys, List Xs,
                           List zs, Nullable<List> ws, List inits) {
  const int B = ys.size();
  cs_fit_config cfg = {kind, optim, lr, beta1, beta2, 1e-8, momentum, 0, 0, 0, 0, 0, 0};
  std::vector<NumericVector> yv(B), zv(B), wv(B), pv(B);
  std::vector<NumericMatrix> Xv(B);
  std::vector<const double*> yp(B), Xp(B), zp(B), wp(B, (const double*)NULL), pp(B);
  List wl = ws.isNotNull() ? List(ws) : List();
  for (int b = 0; b < B; ++b) {
    yv[b] = as<NumericVector>(ys[b]); Xv[b] = as<NumericMatrix>(Xs[b]); zv[b] = as<NumericVector>(zs[b]);
    pv[b] = pack(as<List>(inits[b]));
    yp[b] = yv[b].begin(); Xp[b] = Xv[b].begin(); zp[b] = zv[b].begin(); pp[b] = pv[b].begin();
    if (ws.isNotNull()) { wv[b] = as<NumericVector>(wl[b]); wp[b] = wv[b].begin(); }
  }
  cs_fit* f = NULL;
  cs_check(cs_fit_create_batch(&cfg, B, Xv[0].nrow(), Xv[0].ncol(), yp.data(), Xp.data(), zp.data(), wp.data(), pp.data(), &f));
  return FitPtr(f, true);
}
// [[Rcpp::export]]
List b200_fit_run_batch(SEXP handle, int maxiter, double tol) {
  FitPtr f(handle);
  const int B = cs_fit_batch_size(f.get());
  IntegerVector iters(B);
  NumericVector stats(2 * (maxiter + 2) * B);
  cs_check(cs_fit_run(f.get(), maxiter, tol, iters.begin(), stats.begin()));
  stats.attr("dim") = IntegerVector::create(2, maxiter + 2, B);   // stats[, , b] = the `stats` matrix of R/main.R:77
  return List::create(Named("iter") = iters, Named("stats") = stats);
}
// [[Rcpp::export]]
void b200_fit_select(SEXP handle, int member) {   // 0-based; predict / treat / get_params then act on this member
  FitPtr f(handle);
  cs_check(cs_fit_select(f.get(), member));
}
// [[Rcpp::export]]
List b200_fit_get_params(SEXP handle, List parameters) {
  FitPtr f(handle);
  NumericVector pv(cs_fit_nparam(f.get()));
  cs_check(cs_fit_get_params(f.get(), pv.begin()));
  return unpack(pv, parameters);
}
// [[Rcpp::export]]
List b200_predict(SEXP handle, NumericMatrix X2, NumericVector z2, bool want_cov) {
  FitPtr f(handle);
  const int n2 = X2.nrow();
  NumericVector map(n2), var(n2);
  NumericMatrix cov(want_cov ? n2 : 0, want_cov ? n2 : 0);
  cs_check(cs_predict(f.get(), n2, X2.begin(), z2.begin(), map.begin(), var.begin(), want_cov ? cov.begin() : NULL));
  return List::create(Named("map") = map, Named("var") = var, Named("cov") = cov);
}
// [[Rcpp::export]]
List b200_treat(SEXP handle, NumericMatrix X2, double jitter, bool want_cov) {
  FitPtr f(handle);
  const int n2 = X2.nrow();
  NumericVector map(n2), var(n2);
  double ate = 0, ate_var = 0;
  NumericMatrix cov(want_cov ? n2 : 0, want_cov ? n2 : 0);
  cs_check(cs_treat(f.get(), n2, X2.begin(), jitter, map.begin(), var.begin(), &ate, &ate_var, want_cov ? cov.begin() : NULL));
  return List::create(Named("map") = map, Named("var") = var, Named("ate") = ate, Named("ate_var") = ate_var, Named("cov") = cov);
}
// [[Rcpp::export]]
List b200_tp_sample(NumericMatrix cov, double df, int nsampling, double seed) {
  const int n2 = cov.nrow();
  NumericVector U(n2);
  double ate_U = 0;
  cs_check(cs_tp_sample(n2, cov.begin(), df, nsampling, (unsigned long long)seed, U.begin(), &ate_U));
  return List::create(Named("U") = U, Named("ate_U") = ate_U);
}
