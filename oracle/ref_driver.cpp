// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// Flat C entry points over the reference's OWN functions (compiled unmodified from
// /root/reference/src/*.cpp against oracle/shim/RcppArmadillo.h, see oracle/Makefile), so
// tests/ and tests/golden/make_golden_ref.py can call them through ctypes.  The prototypes
// below are the ones src/RcppExports.cpp declares (:10, :24, :38, :53, :72, :86, :101, :120,
// :134, :149, :168).  Parameter vectors use the reference's list order
// (GP: sigma, lambdam, lambdaa, Lm[p], La[p], mu; TP: sigma, lambdam, nu, lambda0, Lm[p], La[p], mu).
#include <RcppArmadillo.h>

#include <cstring>

arma::mat invkernel_cpp(arma::colvec z, arma::colvec w, arma::mat mymat, Rcpp::List parameters);
double mu_solution_cpp(arma::colvec y, arma::colvec z, arma::mat invKmat, Rcpp::List parameters);
Rcpp::List kernmat_GP_SE_cpp(Rcpp::NumericMatrix X1, Rcpp::NumericMatrix X2, Rcpp::NumericVector Z1, Rcpp::NumericVector Z2, Rcpp::List para);
Rcpp::List grad_GP_SE_cpp(arma::colvec y, arma::mat X, arma::colvec z, arma::colvec w, arma::mat Kmat, arma::mat Km, arma::mat Ka, arma::mat invKmatn, Rcpp::List parameters);
arma::rowvec stats_GP_SE(arma::colvec y, arma::mat Kmat, arma::mat invKmatn, Rcpp::List parameters);
Rcpp::List kernmat_TP_SE_cpp(Rcpp::NumericMatrix X1, Rcpp::NumericMatrix X2, Rcpp::NumericVector Z1, Rcpp::NumericVector Z2, Rcpp::List para);
Rcpp::List grad_TP_SE_cpp(arma::colvec y, arma::mat X, arma::colvec z, arma::colvec w, arma::mat Smat, arma::mat Sm, arma::mat Sa, arma::mat invSmatn, Rcpp::List parameters);
arma::rowvec stats_TP_SE(arma::colvec y, arma::mat Kmat, arma::mat invKmatn, Rcpp::List parameters);
Rcpp::List Nesterov_cpp(double learn_rate, double momentum, Rcpp::List nu, Rcpp::List grad, Rcpp::List para);
Rcpp::List Nadam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, Rcpp::List m, Rcpp::List v, Rcpp::List grad, Rcpp::List para);
Rcpp::List Adam_cpp(double iter, double learn_rate, double beta1, double beta2, double eps, Rcpp::List m, Rcpp::List v, Rcpp::List grad, Rcpp::List para);

namespace {

arma::vec mkvec(const double* p, int n) { arma::vec v((arma::uword)n); std::memcpy(v.d.data(), p, sizeof(double) * n); return v; }
arma::mat mkmat(const double* p, int r, int c) { arma::mat m((arma::uword)r, (arma::uword)c); std::memcpy(m.d.data(), p, sizeof(double) * (size_t)r * c); return m; }
Rcpp::NumericVector mknv(const double* p, int n) { return Rcpp::NumericVector(std::vector<double>(p, p + n)); }
Rcpp::NumericMatrix mknm(const double* p, int r, int c) { Rcpp::NumericMatrix m(r, c); std::memcpy(m.d.data(), p, sizeof(double) * (size_t)r * c); return m; }

// flat vector -> the parameter list of parainit (R/kernel_GP_SE_class.R:13-19, R/kernel_TP_SE_class.R:14-21)
Rcpp::List mklist(int kind, int p, const double* v) {
  Rcpp::List l;
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
void flatten(const Rcpp::List& l, double* out) {
  size_t k = 0;
  for (const auto& it : l.items) for (double x : it.second.num) out[k++] = x;
}
const Rcpp::Value& get(const Rcpp::List& l, const char* name) {
  for (const auto& it : l.items) if (it.first == name) return it.second;
  throw std::runtime_error(std::string("missing list entry ") + name);
}

}  // namespace

extern "C" {

// Optional: host LAPACK for arma::inv_sympd / arma::log_det (LP64 Fortran dpotrf_, dpotri_, dgetrf_); nulls switch
// back to the shim's plain loops.
void ref_set_lapack(void* dpotrf, void* dpotri, void* dgetrf) {
  arma::lapack_hook::potrf = (arma::lapack_hook::potrf_t)dpotrf;
  arma::lapack_hook::potri = (arma::lapack_hook::potrf_t)dpotri;
  arma::lapack_hook::getrf = (arma::lapack_hook::getrf_t)dgetrf;
}

// arma::log_det alone (the LU that grad_*_SE_cpp / stats_*_SE run on K^-1), for the bounded-sample timing of bench.py
int ref_log_det(int n, const double* A, double* val, double* sign) {
  try { arma::log_det(*val, *sign, mkmat(A, n, n)); return 0; } catch (const std::exception&) { return -1; }
}

int ref_kernmat(int kind, int n1, int n2, int p, const double* X1, const double* X2, const double* Z1, const double* Z2,
                const double* Lm, const double* La, double lambdam, double lambda_a, double* K, double* Km, double* Ka) {
  try {
    Rcpp::List para;
    para["Lm"] = std::vector<double>(Lm, Lm + p);
    para["La"] = std::vector<double>(La, La + p);
    para["lambdam"] = lambdam;
    para[kind == 0 ? "lambdaa" : "lambda0"] = lambda_a;
    Rcpp::List out = kind == 0 ? kernmat_GP_SE_cpp(mknm(X1, n1, p), mknm(X2, n2, p), mknv(Z1, n1), mknv(Z2, n2), para)
                               : kernmat_TP_SE_cpp(mknm(X1, n1, p), mknm(X2, n2, p), mknv(Z1, n1), mknv(Z2, n2), para);
    const size_t nb = sizeof(double) * (size_t)n1 * n2;
    std::memcpy(K, get(out, kind == 0 ? "Kmat" : "Smat").num.data(), nb);
    std::memcpy(Km, get(out, kind == 0 ? "Km" : "Sm").num.data(), nb);
    std::memcpy(Ka, get(out, kind == 0 ? "Ka" : "Sa").num.data(), nb);
    return 0;
  } catch (const std::exception&) { return -1; }
}

int ref_invkernel(int n, const double* z, const double* w, const double* K, double sigma, double* invK) {
  try {
    Rcpp::List para; para["sigma"] = sigma;
    arma::mat out = invkernel_cpp(mkvec(z, n), mkvec(w, n), mkmat(K, n, n), para);
    std::memcpy(invK, out.d.data(), sizeof(double) * (size_t)n * n);
    return 0;
  } catch (const std::exception&) { return 1; }  // inv_sympd threw
}

int ref_mu_solution(int n, const double* y, const double* z, const double* invK, double* mu) {
  Rcpp::List para;
  *mu = mu_solution_cpp(mkvec(y, n), mkvec(z, n), mkmat(invK, n, n), para);
  return 0;
}

int ref_grad(int kind, int n, int p, const double* y, const double* X, const double* z, const double* w, const double* K,
             const double* Km, const double* Ka, const double* invK, const double* params, double* grad, double* stats) {
  try {
    Rcpp::List para = mklist(kind, p, params);
    Rcpp::List out = kind == 0 ? grad_GP_SE_cpp(mkvec(y, n), mkmat(X, n, p), mkvec(z, n), mkvec(w, n), mkmat(K, n, n), mkmat(Km, n, n), mkmat(Ka, n, n), mkmat(invK, n, n), para)
                               : grad_TP_SE_cpp(mkvec(y, n), mkmat(X, n, p), mkvec(z, n), mkvec(w, n), mkmat(K, n, n), mkmat(Km, n, n), mkmat(Ka, n, n), mkmat(invK, n, n), para);
    flatten(*get(out, "gradients").lst, grad);
    const Rcpp::Value& st = get(out, "stats");
    stats[0] = st.num[0]; stats[1] = st.num[1];
    return 0;
  } catch (const std::exception&) { return -1; }
}

int ref_stats(int kind, int n, int p, const double* y, const double* K, const double* invK, const double* params, double* stats) {
  try {
    Rcpp::List para = mklist(kind, p, params);
    arma::rowvec st = kind == 0 ? stats_GP_SE(mkvec(y, n), mkmat(K, n, n), mkmat(invK, n, n), para)
                                : stats_TP_SE(mkvec(y, n), mkmat(K, n, n), mkmat(invK, n, n), para);
    stats[0] = st(0); stats[1] = st(1);
    return 0;
  } catch (const std::exception&) { return -1; }
}

// optim: 0 Adam, 1 Nadam, 2 Nesterov (m = velocity `nu`, beta1 = momentum).  The moment lists hold the
// same entries as the parameter list INCLUDING mu (list2zero, R/optimizer_classes.R:31), so mu moves too;
// m, v, para are updated in place.
int ref_opt_step(int optim, int kind, int p, double iter, double lr, double beta1, double beta2, double eps, double* m,
                 double* v, const double* grad, double* para) {
  try {
    Rcpp::List lm = mklist(kind, p, m), lg = mklist(kind, p, grad), lp = mklist(kind, p, para);
    if (optim == 2) {
      Rcpp::List out = Nesterov_cpp(lr, beta1, lm, lg, lp);
      flatten(*get(out, "nu").lst, m);
      flatten(*get(out, "parameters").lst, para);
    } else {
      Rcpp::List lv = mklist(kind, p, v);
      Rcpp::List out = optim == 0 ? Adam_cpp(iter, lr, beta1, beta2, eps, lm, lv, lg, lp) : Nadam_cpp(iter, lr, beta1, beta2, eps, lm, lv, lg, lp);
      flatten(*get(out, "m").lst, m);
      flatten(*get(out, "v").lst, v);
      flatten(*get(out, "parameters").lst, para);
    }
    return 0;
  } catch (const std::exception&) { return -1; }
}

}  // extern "C"
