// oracle/shim/Rcpp.h -- TEST INFRASTRUCTURE ONLY.
// The slice of plain Rcpp that causalstump_b200/r/src/b200_glue.cpp (the R-side glue of the drop-in, which
// `#include <Rcpp.h>`) uses on top of oracle/shim/RcppArmadillo.h: stop(), wrap(), XPtr, Nullable, IntegerVector,
// attributes, const list access.  It lets the glue be COMPILED and RUN here without R: tests/test_r_glue.py calls
// the glue's 11 routines (same names and argument lists as src/RcppExports.cpp:10-168) next to the reference's own
// compiled routines.  Eager, minimal, not a general Rcpp replacement.
#pragma once
#include "RcppArmadillo.h"

#include <cstdarg>
#include <cstdio>

typedef struct SEXPREC { void* ptr; } * SEXP;

namespace Rcpp {

inline void stop(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw std::runtime_error(buf);
}

struct IntegerVector {
  std::vector<int> d;
  IntegerVector() {}
  explicit IntegerVector(size_t n) : d(n, 0) {}
  int* begin() { return d.data(); }
  size_t size() const { return d.size(); }
  int& operator[](size_t i) { return d[i]; }
  static IntegerVector create(int a, int b, int c) { IntegerVector v(3); v.d[0] = a; v.d[1] = b; v.d[2] = c; return v; }
};

inline NumericVector wrap(const std::vector<double>& v) { return NumericVector(v); }
inline double* begin(NumericVector& v) { return v.d.data(); }

// external pointer with finalizer (the device handle held by R)
struct PreserveStorage {};
template <class T, class Storage, void Finalizer(T*)>
struct XPtr {
  SEXP s;
  XPtr(T* p, bool) : s(new SEXPREC{p}) {}
  XPtr(SEXP h) : s(h) {}
  T* get() const { return static_cast<T*>(s->ptr); }
  operator SEXP() const { return s; }
  static void release(SEXP h) { if (h && h->ptr) { Finalizer(static_cast<T*>(h->ptr)); h->ptr = nullptr; } }
};

template <class T>
struct Nullable {
  bool set = false;
  T val;
  Nullable() {}
  Nullable(const T& v) : set(true), val(v) {}
  bool isNotNull() const { return set; }
  operator T() const { return val; }
};

template <> struct AsImpl<NumericMatrix> {
  static NumericMatrix get(const Value& v) {
    NumericMatrix m(v.nrow < 0 ? (int)v.num.size() : v.nrow, v.ncol < 0 ? 1 : v.ncol);
    m.d = v.num;
    return m;
  }
};
template <> struct AsImpl<List> { static List get(const Value& v) { return v.lst ? *v.lst : List(); } };

}  // namespace Rcpp
