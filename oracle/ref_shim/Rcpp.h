// Rcpp.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A minimal stand-in for the Rcpp API, just large enough to compile the reference's UNMODIFIED C++ sources
// (/root/reference/src/*.cpp, compiled where they lie by oracle/ref_shim/Makefile into oracle/_ref/) without R:
// vectors / matrices with R's column-major layout and handle (aliasing) semantics, List / Named, Range and `_`,
// the Rcpp sugar the sources use (runif, rnorm, sum, abs, sqrt, elementwise arithmetic), RNGScope.
// Random draws come from the tape-fed R-math stand-ins of Rmath.h (unif_rand / norm_rand / exp_rand).
// Nothing here is derived from Rcpp's implementation; it restates the documented behaviour of the calls in use.
#ifndef UBM_REF_SHIM_RCPP_H
#define UBM_REF_SHIM_RCPP_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "R.h"
#include "Rmath.h"

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

namespace Rcpp {

struct NamePlaceholder {};
static const NamePlaceholder _ = NamePlaceholder();

struct Range {
  int lo, hi;   // inclusive, 0-based
  Range(int a, int b) : lo(a), hi(b) {}
};

// ---- vectors: a handle to shared storage, as an R vector seen through Rcpp -------------------------------------
template <class T> class Matrix;
template <class T>
class Vector {
 public:
  std::shared_ptr<std::vector<T>> p;
  Vector() : p(std::make_shared<std::vector<T>>()) {}
  Vector(const Matrix<T>& m);                    // an R matrix is a vector with a dim attribute (NumericVector(wrap(matrix)))
  Vector(int n) : p(std::make_shared<std::vector<T>>((size_t)n, T())) {}
  Vector(int n, T v) : p(std::make_shared<std::vector<T>>((size_t)n, v)) {}
  Vector(const T* src, int n) : p(std::make_shared<std::vector<T>>(src, src + n)) {}
  int size() const { return (int)p->size(); }
  int length() const { return size(); }
  T& operator()(int i) { return (*p)[(size_t)i]; }
  const T& operator()(int i) const { return (*p)[(size_t)i]; }
  T& operator[](int i) { return (*p)[(size_t)i]; }
  const T& operator[](int i) const { return (*p)[(size_t)i]; }
  typename std::vector<T>::iterator begin() { return p->begin(); }
  typename std::vector<T>::iterator end() { return p->end(); }
  typename std::vector<T>::const_iterator begin() const { return p->begin(); }
  typename std::vector<T>::const_iterator end() const { return p->end(); }
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;
typedef Vector<int> LogicalVector;

// ---- matrices: column-major, handle semantics; rows and columns as assignable views -----------------------------
template <class T>
class Matrix {
 public:
  std::shared_ptr<std::vector<T>> p;
  int nr = 0, nc = 0;
  Matrix() : p(std::make_shared<std::vector<T>>()) {}
  Matrix(int r, int c) : p(std::make_shared<std::vector<T>>((size_t)r * c, T())), nr(r), nc(c) {}
  Matrix(const T* src, int r, int c) : p(std::make_shared<std::vector<T>>(src, src + (size_t)r * c)), nr(r), nc(c) {}
  int nrow() const { return nr; }
  int ncol() const { return nc; }
  int rows() const { return nr; }
  int cols() const { return nc; }
  int size() const { return nr * nc; }
  T& operator()(int i, int j) { return (*p)[(size_t)j * nr + i]; }
  const T& operator()(int i, int j) const { return (*p)[(size_t)j * nr + i]; }
  typename std::vector<T>::iterator begin() { return p->begin(); }
  typename std::vector<T>::iterator end() { return p->end(); }
  typename std::vector<T>::const_iterator begin() const { return p->begin(); }
  typename std::vector<T>::const_iterator end() const { return p->end(); }

  // a strided view of one row or column
  struct View {
    T* base; int stride, n;
    T& operator()(int i) const { return base[(size_t)i * stride]; }
    T& operator[](int i) const { return base[(size_t)i * stride]; }
    int size() const { return n; }
    const View& operator=(const View& o) const { for (int i = 0; i < n; ++i) (*this)(i) = o(i); return *this; }
    operator Vector<T>() const { Vector<T> v(n); for (int i = 0; i < n; ++i) v(i) = (*this)(i); return v; }
  };
  View operator()(int i, NamePlaceholder) const { return View{const_cast<T*>(p->data()) + i, nr, nc}; }
  View operator()(NamePlaceholder, int j) const { return View{const_cast<T*>(p->data()) + (size_t)j * nr, 1, nr}; }
  View row(int i) const { return (*this)(i, _); }
  View column(int j) const { return (*this)(_, j); }
  // rows lo..hi as a new matrix
  Matrix operator()(const Range& r, NamePlaceholder) const {
    Matrix out(r.hi - r.lo + 1, nc);
    for (int j = 0; j < nc; ++j) for (int i = r.lo; i <= r.hi; ++i) out(i - r.lo, j) = (*this)(i, j);
    return out;
  }
};
typedef Matrix<double> NumericMatrix;
template <class T> Vector<T>::Vector(const Matrix<T>& m) : p(std::make_shared<std::vector<T>>(*m.p)) {}
typedef Matrix<int> IntegerMatrix;

// ---- List: named heterogeneous values --------------------------------------------------------------------------
struct Any {
  enum Kind { NONE, INT, DBL, NVEC, IVEC, NMAT, IMAT } kind = NONE;
  int i = 0; double d = 0;
  NumericVector nv; IntegerVector iv; NumericMatrix nm; IntegerMatrix im;
  Any() {}
  Any(int v) : kind(INT), i(v) {}
  Any(bool v) : kind(INT), i(v ? 1 : 0) {}
  Any(double v) : kind(DBL), d(v) {}
  Any(const NumericVector& v) : kind(NVEC), nv(v) {}
  Any(const IntegerVector& v) : kind(IVEC), iv(v) {}
  Any(const NumericMatrix& v) : kind(NMAT), nm(v) {}
  Any(const IntegerMatrix& v) : kind(IMAT), im(v) {}
  Any(const NumericMatrix::View& v) : kind(NVEC), nv(v) {}
  template <class T> T as() const;
};
template <> inline int Any::as<int>() const { if (kind == INT) return i; if (kind == DBL) return (int)d; throw std::runtime_error("List: not a scalar"); }
template <> inline double Any::as<double>() const { if (kind == DBL) return d; if (kind == INT) return (double)i; throw std::runtime_error("List: not a scalar"); }
template <> inline bool Any::as<bool>() const { return as<int>() != 0; }
template <> inline NumericVector Any::as<NumericVector>() const { if (kind != NVEC) throw std::runtime_error("List: not a numeric vector"); return nv; }
template <> inline IntegerVector Any::as<IntegerVector>() const { if (kind != IVEC) throw std::runtime_error("List: not an integer vector"); return iv; }
template <> inline NumericMatrix Any::as<NumericMatrix>() const { if (kind != NMAT) throw std::runtime_error("List: not a numeric matrix"); return nm; }
template <> inline IntegerMatrix Any::as<IntegerMatrix>() const { if (kind != IMAT) throw std::runtime_error("List: not an integer matrix"); return im; }

// conversion of a value into a List element; other headers (RcppEigen.h) add overloads found by argument-dependent lookup
inline Any to_any(int v) { return Any(v); }
inline Any to_any(bool v) { return Any(v); }
inline Any to_any(double v) { return Any(v); }
inline Any to_any(const NumericVector& v) { return Any(v); }
inline Any to_any(const IntegerVector& v) { return Any(v); }
inline Any to_any(const NumericMatrix& v) { return Any(v); }
inline Any to_any(const IntegerMatrix& v) { return Any(v); }
inline Any to_any(const NumericMatrix::View& v) { return Any(v); }

struct NamedObject { std::string name; Any value; };
struct Named {
  std::string name;
  explicit Named(const std::string& n) : name(n) {}
  template <class T> NamedObject operator=(const T& v) const { return NamedObject{name, to_any(v)}; }
};

class List {
 public:
  std::vector<NamedObject> items;
  struct Proxy {
    const Any* a;
    template <class T> operator T() const { return a->as<T>(); }
  };
  Proxy operator[](const std::string& name) const {
    for (const NamedObject& o : items) if (o.name == name) return Proxy{&o.value};
    throw std::runtime_error("List: no element named " + name);
  }
  void push(const NamedObject& o) { items.push_back(o); }
  static List create() { return List(); }
  template <class... Rest>
  static List create(const NamedObject& first, const Rest&... rest) {
    List l;
    l.push(first);
    List tail = create(rest...);
    for (const NamedObject& o : tail.items) l.push(o);
    return l;
  }
};

// ---- random draws (Rcpp sugar): n calls of the scalar R generators, in order -----------------------------------
struct RNGScope { RNGScope() {} ~RNGScope() {} };

inline NumericVector runif(int n, double a = 0.0, double b = 1.0) {
  NumericVector v(n);
  for (int i = 0; i < n; ++i) {
    double u;
    do { u = ::unif_rand(); } while (u <= 0 || u >= 1);
    v(i) = a + (b - a) * u;
  }
  return v;
}
inline NumericVector rnorm(int n, double mean = 0.0, double sd = 1.0) {
  NumericVector v(n);
  const bool std01 = (mean == 0.0 && sd == 1.0);
  for (int i = 0; i < n; ++i) v(i) = std01 ? ::norm_rand() : mean + sd * ::norm_rand();
  return v;
}

// ---- the sugar used by the sources (eager, elementwise, same operation order per element) -------------------------
inline double sum(const NumericVector& v) { double s = 0; for (int i = 0; i < v.size(); ++i) s += v(i); return s; }
inline NumericVector abs(const NumericVector& v) { NumericVector r(v.size()); for (int i = 0; i < v.size(); ++i) r(i) = std::fabs(v(i)); return r; }
inline NumericVector sqrt(const NumericVector& v) { NumericVector r(v.size()); for (int i = 0; i < v.size(); ++i) r(i) = std::sqrt(v(i)); return r; }
#define UBM_SHIM_BINOP(OP)                                                                                              \
  inline NumericVector operator OP(const NumericVector& a, const NumericVector& b) {                                    \
    NumericVector r(a.size()); for (int i = 0; i < a.size(); ++i) r(i) = a(i) OP b(i); return r; }                    \
  inline NumericVector operator OP(const NumericVector& a, double b) {                                                  \
    NumericVector r(a.size()); for (int i = 0; i < a.size(); ++i) r(i) = a(i) OP b; return r; }                       \
  inline NumericVector operator OP(double a, const NumericVector& b) {                                                  \
    NumericVector r(b.size()); for (int i = 0; i < b.size(); ++i) r(i) = a OP b(i); return r; }
UBM_SHIM_BINOP(+)
UBM_SHIM_BINOP(-)
UBM_SHIM_BINOP(*)
UBM_SHIM_BINOP(/)
#undef UBM_SHIM_BINOP

}  // namespace Rcpp

#endif
