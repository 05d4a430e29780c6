// oracle/shim/RcppArmadillo.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal, eager stand-in for the slice of the Rcpp / RcppArmadillo API that the reference's
// three source files use (src/kernel_GP_SE_cpp.cpp, src/kernel_TP_SE_cpp.cpp,
// src/optimizer_cpp.cpp of mazphilip/CausalStump).  R, Rcpp and Armadillo are not installed in
// this image, so oracle/Makefile compiles those files UNMODIFIED, from where they lie under
// /root/reference, against this header (their `#include <RcppArmadillo.h>` resolves here) into
// oracle/_ref/libcausalstump_ref.so.  Every loop, formula and quirk of the reference's own C++
// therefore runs verbatim; only the third-party container / linear-algebra calls are restated:
//   arma::inv_sympd   Cholesky (lower) -> triangular inverse -> L^-T L^-1, symmetric result;
//                     throws std::runtime_error on a non-positive pivot
//   arma::log_det     LU with partial pivoting (dgetf2 order), val = sum log|u_ii|, sign
//   arma::trace(A*B)  sum_k sum_i A(k,i) B(i,k) without forming the product (Armadillo
//                     special-cases the same way)
//   Rcpp::List        ordered (name, value) pairs, value = numeric vector (+dims) or list
// inv_sympd / log_det use the host's LAPACK (dpotrf + dpotri, dgetrf -- the calls Armadillo makes for them)
// when arma::lapack_hook holds the three entry points (ref_set_lapack in oracle/ref_driver.cpp; the timing
// legs of bench.py set it to SciPy's OpenBLAS), else the plain loops below.
// Nothing here is shipped or linked into the product library.
#pragma once
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <initializer_list>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------
namespace Rcpp {
struct NumericVector;
struct NumericMatrix;
}  // namespace Rcpp

namespace arma {

typedef unsigned int uword;
struct mat;
struct vec;
struct rowvec;

namespace datum { static const double pi = 3.14159265358979323846264338327950288; }

struct vec {  // column vector (arma::Col<double>)
  std::vector<double> d;
  uword n_elem = 0, n_rows = 0, n_cols = 1;
  vec() {}
  explicit vec(uword n) : d(n, 0.0), n_elem(n), n_rows(n) {}
  vec(const std::vector<double>& v) : d(v), n_elem((uword)v.size()), n_rows((uword)v.size()) {}
  uword size() const { return n_elem; }
  double& operator()(uword i) { return d[i]; }
  double operator()(uword i) const { return d[i]; }
  double& operator[](uword i) { return d[i]; }
  double operator[](uword i) const { return d[i]; }
  rowvec t() const;
};
typedef vec colvec;

struct rowvec {  // row vector (arma::Row<double>)
  std::vector<double> d;
  uword n_elem = 0, n_rows = 1, n_cols = 0;
  rowvec() {}
  explicit rowvec(uword n) : d(n, 0.0), n_elem(n), n_cols(n) {}
  uword size() const { return n_elem; }
  double& operator()(uword i) { return d[i]; }
  double operator()(uword i) const { return d[i]; }
  rowvec& operator=(const Rcpp::NumericVector& v);  // RcppArmadillo sugar -> Row conversion
};
inline rowvec vec::t() const { rowvec r(n_elem); r.d = d; return r; }

struct colview;
struct rowview;
struct diagview;
struct matprod;

struct mat {  // column-major dense matrix (arma::Mat<double>)
  std::vector<double> d;
  uword n_rows = 0, n_cols = 0;
  mat() {}
  mat(uword r, uword c) : d((size_t)r * c, 0.0), n_rows(r), n_cols(c) {}
  mat(const matprod& p);
  mat& operator=(const matprod& p);
  void zeros() { std::fill(d.begin(), d.end(), 0.0); }
  double& operator()(uword r, uword c) { return d[(size_t)c * n_rows + r]; }
  double operator()(uword r, uword c) const { return d[(size_t)c * n_rows + r]; }
  colview col(uword c);
  rowview row(uword r);
  diagview diag();
  mat t() const {
    mat o(n_cols, n_rows);
    for (uword c = 0; c < n_cols; ++c) for (uword r = 0; r < n_rows; ++r) o(c, r) = (*this)(r, c);
    return o;
  }
};

struct colview {
  mat& m; uword c;
  void operator=(const vec& v) { for (uword r = 0; r < m.n_rows; ++r) m(r, c) = v.d[r]; }
  operator vec() const { vec v(m.n_rows); for (uword r = 0; r < m.n_rows; ++r) v.d[r] = m(r, c); return v; }
};
struct rowview {
  mat& m; uword r;
  void operator+=(const rowvec& v) { for (uword c = 0; c < m.n_cols; ++c) m(r, c) += v.d[c]; }
  void operator=(const rowvec& v) { for (uword c = 0; c < m.n_cols; ++c) m(r, c) = v.d[c]; }
};
struct diagview {
  mat& m;
  void operator+=(const vec& v) { for (uword i = 0; i < m.n_rows && i < m.n_cols; ++i) m(i, i) += v.d[i]; }
};
inline colview mat::col(uword c) { return colview{*this, c}; }
inline rowview mat::row(uword r) { return rowview{*this, r}; }
inline diagview mat::diag() { return diagview{*this}; }

// ---- vec / rowvec arithmetic ---------------------------------------------------------------
inline vec operator-(double a, const colview& v) { vec o(v.m.n_rows); for (uword r = 0; r < o.n_elem; ++r) o.d[r] = a - v.m(r, v.c); return o; }
inline vec operator-(const vec& a, double b) { vec o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.d[i] = a.d[i] - b; return o; }
inline vec operator+(const vec& a, double b) { vec o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.d[i] = a.d[i] + b; return o; }
inline vec operator-(const vec& a, const vec& b) { vec o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.d[i] = a.d[i] - b.d[i]; return o; }
inline vec operator/(double a, const vec& b) { vec o(b.n_elem); for (uword i = 0; i < b.n_elem; ++i) o.d[i] = a / b.d[i]; return o; }
inline vec operator*(double a, const vec& b) { vec o(b.n_elem); for (uword i = 0; i < b.n_elem; ++i) o.d[i] = a * b.d[i]; return o; }
inline vec pow(const vec& a, int e) {
  vec o(a.n_elem);
  for (uword i = 0; i < a.n_elem; ++i) o.d[i] = (e == 2) ? a.d[i] * a.d[i] : std::pow(a.d[i], (double)e);
  return o;
}
inline rowvec operator*(const rowvec& a, double b) { rowvec o(a.n_elem); for (uword i = 0; i < a.n_elem; ++i) o.d[i] = a.d[i] * b; return o; }
inline mat operator*(const vec& a, const rowvec& b) {  // outer product
  mat o(a.n_elem, b.n_elem);
  for (uword c = 0; c < b.n_elem; ++c) for (uword r = 0; r < a.n_elem; ++r) o(r, c) = a.d[r] * b.d[c];
  return o;
}
inline double dot(const vec& a, const vec& b) { double s = 0.0; for (uword i = 0; i < a.n_elem; ++i) s += a.d[i] * b.d[i]; return s; }
inline double norm(const vec& a) { return std::sqrt(dot(a, a)); }
inline double sum(const vec& a) { double s = 0.0; for (double x : a.d) s += x; return s; }
inline double sum(const rowvec& a) { double s = 0.0; for (double x : a.d) s += x; return s; }

// ---- mat arithmetic ------------------------------------------------------------------------
inline rowvec sum(const mat& A) {  // column sums
  rowvec o(A.n_cols);
  for (uword c = 0; c < A.n_cols; ++c) { double s = 0.0; for (uword r = 0; r < A.n_rows; ++r) s += A(r, c); o.d[c] = s; }
  return o;
}
inline vec operator*(const mat& A, const vec& x) {
  vec o(A.n_rows);
  for (uword c = 0; c < A.n_cols; ++c) { const double xc = x.d[c]; for (uword r = 0; r < A.n_rows; ++r) o.d[r] += A(r, c) * xc; }
  return o;
}
inline mat operator%(const mat& A, const mat& B) { mat o(A.n_rows, A.n_cols); for (size_t i = 0; i < o.d.size(); ++i) o.d[i] = A.d[i] * B.d[i]; return o; }
inline mat operator*(const mat& A, double b) { mat o(A.n_rows, A.n_cols); for (size_t i = 0; i < o.d.size(); ++i) o.d[i] = A.d[i] * b; return o; }
inline mat operator-(const mat& A, const mat& B) { mat o(A.n_rows, A.n_cols); for (size_t i = 0; i < o.d.size(); ++i) o.d[i] = A.d[i] - B.d[i]; return o; }
inline mat diagmat(const vec& v) { mat o(v.n_elem, v.n_elem); for (uword i = 0; i < v.n_elem; ++i) o(i, i) = v.d[i]; return o; }

// A * B stays lazy so that trace(A * B) is O(n^2), as in Armadillo
struct matprod { const mat& A; const mat& B; };
inline matprod operator*(const mat& A, const mat& B) { return matprod{A, B}; }
inline void eval_prod(const matprod& p, mat& o) {
  o = mat(p.A.n_rows, p.B.n_cols);
  for (uword c = 0; c < p.B.n_cols; ++c)
    for (uword k = 0; k < p.A.n_cols; ++k) { const double b = p.B(k, c); for (uword r = 0; r < p.A.n_rows; ++r) o(r, c) += p.A(r, k) * b; }
}
inline mat::mat(const matprod& p) { eval_prod(p, *this); }
inline mat& mat::operator=(const matprod& p) { mat o; eval_prod(p, o); *this = o; return *this; }
inline double trace(const matprod& p) {
  double s = 0.0;
  for (uword k = 0; k < p.A.n_rows; ++k) for (uword i = 0; i < p.A.n_cols; ++i) s += p.A(k, i) * p.B(i, k);
  return s;
}
inline double trace(const mat& A) { double s = 0.0; for (uword i = 0; i < A.n_rows && i < A.n_cols; ++i) s += A(i, i); return s; }

// LP64 Fortran entry points of the host LAPACK (optional)
namespace lapack_hook {
typedef void (*potrf_t)(const char* uplo, const int* n, double* a, const int* lda, int* info);
typedef void (*getrf_t)(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
inline potrf_t potrf = nullptr, potri = nullptr;
inline getrf_t getrf = nullptr;
}  // namespace lapack_hook

// inv_sympd: A = L L^T (lower, unblocked), X = L^-1, A^-1 = X^T X
inline mat inv_sympd(const mat& Ain) {
  const uword n = Ain.n_rows;
  if (lapack_hook::potrf && lapack_hook::potri) {
    mat A = Ain;
    const int N = (int)n; int info = 0;
    lapack_hook::potrf("L", &N, A.d.data(), &N, &info);
    if (info != 0) throw std::runtime_error("inv_sympd(): matrix is singular or not positive definite");
    lapack_hook::potri("L", &N, A.d.data(), &N, &info);
    if (info != 0) throw std::runtime_error("inv_sympd(): matrix is singular or not positive definite");
    for (uword c = 1; c < n; ++c) for (uword r = 0; r < c; ++r) A(r, c) = A(c, r);  // lower -> full
    return A;
  }
  mat L = Ain;
  for (uword j = 0; j < n; ++j) {
    double djj = L(j, j);
    for (uword k = 0; k < j; ++k) djj -= L(j, k) * L(j, k);
    if (!(djj > 0.0)) throw std::runtime_error("inv_sympd(): matrix is singular or not positive definite");
    djj = std::sqrt(djj);
    L(j, j) = djj;
    for (uword i = j + 1; i < n; ++i) {
      double s = L(i, j);
      for (uword k = 0; k < j; ++k) s -= L(i, k) * L(j, k);
      L(i, j) = s / djj;
    }
  }
  mat X(n, n);  // lower triangular inverse, column by column (forward substitution on e_j)
  for (uword j = 0; j < n; ++j) {
    X(j, j) = 1.0 / L(j, j);
    for (uword i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (uword k = j; k < i; ++k) s -= L(i, k) * X(k, j);
      X(i, j) = s / L(i, i);
    }
  }
  mat out(n, n);
  for (uword j = 0; j < n; ++j)
    for (uword i = j; i < n; ++i) {
      double s = 0.0;
      for (uword k = i; k < n; ++k) s += X(k, i) * X(k, j);
      out(i, j) = s; out(j, i) = s;
    }
  return out;
}

// log_det via LU with partial pivoting
inline void log_det(double& val, double& sign, const mat& Ain) {
  const uword n = Ain.n_rows;
  mat A = Ain;
  val = 0.0; sign = 1.0;
  if (lapack_hook::getrf) {
    const int N = (int)n; int info = 0;
    std::vector<int> ipiv(n);
    lapack_hook::getrf(&N, &N, A.d.data(), &N, ipiv.data(), &info);
    if (info < 0) throw std::runtime_error("log_det(): dgetrf failed");
    for (uword j = 0; j < n; ++j) {
      const double u = A(j, j);
      if (u == 0.0) { val = -INFINITY; sign = 0.0; return; }
      if (u < 0.0) sign = -sign;
      if ((uword)ipiv[j] != j + 1) sign = -sign;
      val += std::log(std::fabs(u));
    }
    return;
  }
  for (uword j = 0; j < n; ++j) {
    uword piv = j; double best = std::fabs(A(j, j));
    for (uword i = j + 1; i < n; ++i) if (std::fabs(A(i, j)) > best) { best = std::fabs(A(i, j)); piv = i; }
    if (piv != j) { for (uword c = 0; c < n; ++c) std::swap(A(j, c), A(piv, c)); sign = -sign; }
    const double u = A(j, j);
    if (u == 0.0) { val = -INFINITY; sign = 0.0; return; }
    if (u < 0.0) sign = -sign;
    val += std::log(std::fabs(u));
    for (uword i = j + 1; i < n; ++i) A(i, j) /= u;
    for (uword c = j + 1; c < n; ++c) { const double a = A(j, c); if (a != 0.0) for (uword i = j + 1; i < n; ++i) A(i, c) -= A(i, j) * a; }
  }
}

}  // namespace arma

// ------------------------------------------------------------------------------------------
namespace Rcpp {

namespace internal { struct NamedPlaceHolder {}; }
static const internal::NamedPlaceHolder _ = internal::NamedPlaceHolder();

struct NumericVector {
  std::vector<double> d;
  NumericVector() {}
  explicit NumericVector(size_t n) : d(n, 0.0) {}
  NumericVector(const std::vector<double>& v) : d(v) {}
  size_t size() const { return d.size(); }
  double& operator()(size_t i) { return d[i]; }
  double operator()(size_t i) const { return d[i]; }
  double& operator[](size_t i) { return d[i]; }
  double operator[](size_t i) const { return d[i]; }
  double* begin() { return d.data(); }
  double* end() { return d.data() + d.size(); }
  const double* begin() const { return d.data(); }
  const double* end() const { return d.data() + d.size(); }
  // attributes (only "dim" is ever set, by the R glue)
  struct AttrProxy { std::vector<int>* dims; template <class T> void operator=(const T& v) { dims->assign(v.d.begin(), v.d.end()); } };
  std::vector<int> dims;
  AttrProxy attr(const char*) { return AttrProxy{&dims}; }
};
#define CS_SHIM_VV(op) \
  inline NumericVector operator op(const NumericVector& a, const NumericVector& b) { NumericVector o(a.size()); for (size_t i = 0; i < a.size(); ++i) o.d[i] = a.d[i] op b.d[i]; return o; }
#define CS_SHIM_VS(op) \
  inline NumericVector operator op(const NumericVector& a, double b) { NumericVector o(a.size()); for (size_t i = 0; i < a.size(); ++i) o.d[i] = a.d[i] op b; return o; } \
  inline NumericVector operator op(double a, const NumericVector& b) { NumericVector o(b.size()); for (size_t i = 0; i < b.size(); ++i) o.d[i] = a op b.d[i]; return o; }
CS_SHIM_VV(+) CS_SHIM_VV(-) CS_SHIM_VV(*) CS_SHIM_VV(/)
CS_SHIM_VS(+) CS_SHIM_VS(-) CS_SHIM_VS(*) CS_SHIM_VS(/)
#undef CS_SHIM_VV
#undef CS_SHIM_VS
inline NumericVector pow(const NumericVector& a, int e) {
  NumericVector o(a.size());
  for (size_t i = 0; i < a.size(); ++i) o.d[i] = std::pow(a.d[i], (double)e);  // Rcpp sugar pow calls ::pow
  return o;
}
inline NumericVector sqrt(const NumericVector& a) { NumericVector o(a.size()); for (size_t i = 0; i < a.size(); ++i) o.d[i] = std::sqrt(a.d[i]); return o; }

struct MatrixColumn { const struct NumericMatrix& m; int c; };
struct NumericMatrix {
  std::vector<double> d;
  int nr = 0, nc = 0;
  NumericMatrix() {}
  NumericMatrix(int r, int c) : d((size_t)r * c, 0.0), nr(r), nc(c) {}
  int nrow() const { return nr; }
  int ncol() const { return nc; }
  double* begin() { return d.data(); }
  double& operator()(int r, int c) { return d[(size_t)c * nr + r]; }
  double operator()(int r, int c) const { return d[(size_t)c * nr + r]; }
  MatrixColumn operator()(internal::NamedPlaceHolder, int c) const { return MatrixColumn{*this, c}; }
};
inline NumericVector operator-(double a, const MatrixColumn& v) {
  NumericVector o((size_t)v.m.nr);
  for (int r = 0; r < v.m.nr; ++r) o.d[r] = a - v.m(r, v.c);
  return o;
}

// ---- values and lists ----------------------------------------------------------------------
struct List;
struct Value {  // an R object: numeric vector (optionally with dim) or a list
  std::vector<double> num;
  int nrow = -1, ncol = -1;
  std::shared_ptr<List> lst;
  Value() {}
  Value(double x) : num(1, x) {}
  Value(int x) : num(1, (double)x) {}
  template <class IV, class = decltype(std::declval<IV>().d), class = typename std::enable_if<std::is_same<decltype(std::declval<IV>().d), std::vector<int>>::value>::type>
  Value(const IV& v) : num(v.d.begin(), v.d.end()) {}
  Value(const std::vector<double>& v) : num(v) {}
  Value(const NumericVector& v) : num(v.d) {}
  Value(const NumericMatrix& m) : num(m.d), nrow(m.nr), ncol(m.nc) {}
  Value(const arma::vec& v) : num(v.d) {}
  Value(const arma::rowvec& v) : num(v.d) {}
  Value(const arma::mat& m) : num(m.d), nrow((int)m.n_rows), ncol((int)m.n_cols) {}
  Value(const List& l);
};

struct Named {
  std::string name; Value val;
  explicit Named(const char* n) : name(n) {}
  template <class T> Named operator=(const T& v) const { Named o(name.c_str()); o.val = Value(v); return o; }
};

struct ListProxy;
struct ConstListProxy { const Value& v; };
struct List {
  std::vector<std::pair<std::string, Value>> items;
  List() {}
  explicit List(int k) : items((size_t)k) {}
  int size() const { return (int)items.size(); }
  bool containsElementNamed(const char* name) const { for (const auto& it : items) if (it.first == name) return true; return false; }
  ConstListProxy operator[](int i) const { return ConstListProxy{items.at((size_t)i).second}; }
  ConstListProxy operator[](const char* name) const {
    for (const auto& it : items) if (it.first == name) return ConstListProxy{it.second};
    throw std::runtime_error(std::string("Index out of bounds: [index='") + name + "']");
  }
  template <class... Ns> static List create(const Named& a, const Named& b, const Named& c, const Named& d, const Ns&... rest) {
    List l = create(a, b, c);
    l.items.push_back({d.name, d.val});
    (void)std::initializer_list<int>{(l.items.push_back({rest.name, rest.val}), 0)...};
    return l;
  }
  ListProxy operator[](const char* name);
  ListProxy operator[](const std::string& name);
  ListProxy operator[](int i);
  ListProxy operator[](unsigned int i);
  static List create(const Named& a, const Named& b) { List l; l.items.push_back({a.name, a.val}); l.items.push_back({b.name, b.val}); return l; }
  static List create(const Named& a, const Named& b, const Named& c) { List l = create(a, b); l.items.push_back({c.name, c.val}); return l; }
  Value& slot(const std::string& name) {
    for (auto& it : items) if (it.first == name) return it.second;
    items.push_back({name, Value()});
    return items.back().second;
  }
};
inline Value::Value(const List& l) : lst(std::make_shared<List>(l)) {}

struct ListProxy {
  Value& v;
  template <class T> ListProxy& operator=(const T& x) { v = Value(x); return *this; }
  ListProxy& operator=(const ListProxy& o) { v = o.v; return *this; }
  ListProxy& operator=(const ConstListProxy& o) { v = o.v; return *this; }
  operator NumericVector() const { return NumericVector(v.num); }
};
inline ListProxy List::operator[](const char* name) { return ListProxy{slot(name)}; }
inline ListProxy List::operator[](const std::string& name) { return ListProxy{slot(name)}; }
inline ListProxy List::operator[](int i) { return ListProxy{items.at((size_t)i).second}; }
inline ListProxy List::operator[](unsigned int i) { return ListProxy{items.at((size_t)i).second}; }

inline List clone(const List& l) {  // deep copy (nested lists included)
  List o;
  for (const auto& it : l.items) {
    Value v = it.second;
    if (v.lst) v.lst = std::make_shared<List>(clone(*v.lst));
    o.items.push_back({it.first, v});
  }
  return o;
}

template <class T> struct AsImpl;
template <> struct AsImpl<double> { static double get(const Value& v) { if (v.num.size() != 1) throw std::runtime_error("Expecting a single value"); return v.num[0]; } };
template <> struct AsImpl<NumericVector> { static NumericVector get(const Value& v) { return NumericVector(v.num); } };
template <> struct AsImpl<std::vector<double>> { static std::vector<double> get(const Value& v) { return v.num; } };
template <> struct AsImpl<arma::vec> { static arma::vec get(const Value& v) { return arma::vec(v.num); } };
template <class T> T as(const ListProxy& p) { return AsImpl<T>::get(p.v); }
template <class T> T as(const ConstListProxy& p) { return AsImpl<T>::get(p.v); }
template <class T> T as(const Value& v) { return AsImpl<T>::get(v); }

}  // namespace Rcpp

inline arma::rowvec& arma::rowvec::operator=(const Rcpp::NumericVector& v) {
  d = v.d; n_elem = (uword)v.d.size(); n_cols = n_elem;
  return *this;
}
