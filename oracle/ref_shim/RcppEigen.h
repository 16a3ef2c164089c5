// RcppEigen.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Stand-in for RcppEigen: the small slice of the Eigen dense API that the reference's sources use (src/mvnorm.cpp,
// src/logisticregression.cpp, src/vs_mlikelihood.cpp), evaluated eagerly with plain loops.  Eigen itself is not in the
// reference tree (RcppEigen, version unpinned) and leaves the summation order of products, reductions and decompositions
// to its vectorised kernels; here every product / reduction is a sequential sum in index order, LLT is the textbook
// left-looking Cholesky, triangular solves are forward / back substitution and inverse() is Gauss-Jordan with partial
// pivoting.  So the reference's CONTROL FLOW, draw order and formulas run unmodified, while results that depend on Eigen's
// summation order agree with an R build to rounding only (tests compare continuous outputs to 1e-10 of their scale).
#ifndef UBM_REF_SHIM_RCPPEIGEN_H
#define UBM_REF_SHIM_RCPPEIGEN_H

#include "Rcpp.h"

namespace Eigen {

enum { Lower = 1, Upper = 2 };
template <class T> class ArrT;
template <class T> class VecT;
template <class M> class LLT;

template <class T>
class MatT {
 public:
  typedef T Scalar;
  int r = 0, c = 0;
  std::vector<T> a;
  MatT() {}
  MatT(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, T()) {}
  int rows() const { return r; }
  int cols() const { return c; }
  int size() const { return r * c; }
  T& operator()(int i, int j) { return a[(size_t)j * r + i]; }
  const T& operator()(int i, int j) const { return a[(size_t)j * r + i]; }
  T& operator()(int i) { return a[(size_t)i]; }
  const T& operator()(int i) const { return a[(size_t)i]; }
  MatT transpose() const {
    MatT t(c, r);
    for (int j = 0; j < c; ++j) for (int i = 0; i < r; ++i) t(j, i) = (*this)(i, j);
    return t;
  }
  struct ColView {
    T* base; int n;
    const ColView& operator=(const ColView& o) const { for (int i = 0; i < n; ++i) base[i] = o.base[i]; return *this; }
    const ColView& operator=(const ArrT<T>& o) const { for (int i = 0; i < n; ++i) base[i] = o.a[(size_t)i]; return *this; }
    const ColView& operator=(const MatT& o) const { for (int i = 0; i < n; ++i) base[i] = o.a[(size_t)i]; return *this; }
  };
  ColView col(int j) const { return ColView{const_cast<T*>(a.data()) + (size_t)j * r, r}; }
  VecT<T> diagonal() const;
  ArrT<T> array() const;
  T sum() const { T s = T(); for (size_t i = 0; i < a.size(); ++i) s += a[i]; return s; }
  T squaredNorm() const { T s = T(); for (size_t i = 0; i < a.size(); ++i) s += a[i] * a[i]; return s; }
  T norm() const { return std::sqrt(squaredNorm()); }
  void setIdentity() { for (int j = 0; j < c; ++j) for (int i = 0; i < r; ++i) (*this)(i, j) = (i == j) ? T(1) : T(0); }
  LLT<MatT> llt() const;
  MatT inverse() const {            // Gauss-Jordan with partial pivoting
    const int n = r;
    MatT A(*this), I(n, n);
    I.setIdentity();
    for (int k = 0; k < n; ++k) {
      int piv = k;
      for (int i = k + 1; i < n; ++i) if (std::fabs(A(i, k)) > std::fabs(A(piv, k))) piv = i;
      if (piv != k) for (int j = 0; j < n; ++j) { std::swap(A(k, j), A(piv, j)); std::swap(I(k, j), I(piv, j)); }
      const T d = A(k, k);
      for (int j = 0; j < n; ++j) { A(k, j) /= d; I(k, j) /= d; }
      for (int i = 0; i < n; ++i) {
        if (i == k) continue;
        const T f = A(i, k);
        if (f == T(0)) continue;
        for (int j = 0; j < n; ++j) { A(i, j) -= f * A(k, j); I(i, j) -= f * I(k, j); }
      }
    }
    return I;
  }
  template <int Mode> struct TriView {
    const MatT* m;
    MatT solve(const MatT& B) const {           // Mode == Lower: forward substitution, column by column of B
      const int n = m->r;
      MatT X(B);
      for (int col = 0; col < B.c; ++col)
        for (int i = 0; i < n; ++i) {
          T s = X(i, col);
          for (int k = 0; k < i; ++k) s -= (*m)(i, k) * X(k, col);
          X(i, col) = s / (*m)(i, i);
        }
      return X;
    }
  };
  template <int Mode> TriView<Mode> triangularView() const { return TriView<Mode>{this}; }
  struct Colwise {
    const MatT* m;
    MatT squaredNorm() const {                  // 1 x cols
      MatT out(1, m->c);
      for (int j = 0; j < m->c; ++j) { T s = T(); for (int i = 0; i < m->r; ++i) s += (*m)(i, j) * (*m)(i, j); out(0, j) = s; }
      return out;
    }
  };
  Colwise colwise() const { return Colwise{this}; }
};

template <class T>
class VecT : public MatT<T> {
 public:
  VecT() {}
  VecT(int n) : MatT<T>(n, 1) {}
  VecT(const MatT<T>& m) : MatT<T>(m.c == 1 ? m : m.transpose()) {       // vectors transpose implicitly, as in Eigen
    if (m.c != 1 && m.r != 1) throw std::runtime_error("VecT: not a vector");
  }
};

template <class T>
class ArrT {
 public:
  int r = 0, c = 0;
  std::vector<T> a;
  ArrT() {}
  ArrT(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, T()) {}
  int size() const { return r * c; }
  T& operator()(int i) { return a[(size_t)i]; }
  const T& operator()(int i) const { return a[(size_t)i]; }
  T sum() const { T s = T(); for (size_t i = 0; i < a.size(); ++i) s += a[i]; return s; }
  ArrT log() const { ArrT o(r, c); for (size_t i = 0; i < a.size(); ++i) o.a[i] = std::log(a[i]); return o; }
  ArrT square() const { ArrT o(r, c); for (size_t i = 0; i < a.size(); ++i) o.a[i] = a[i] * a[i]; return o; }
  MatT<T> matrix() const { MatT<T> m(r, c); m.a = a; return m; }
};
template <class T> VecT<T> MatT<T>::diagonal() const { VecT<T> d(r < c ? r : c); for (int i = 0; i < d.r; ++i) d(i) = (*this)(i, i); return d; }
template <class T> ArrT<T> MatT<T>::array() const { ArrT<T> o(r, c); o.a = a; return o; }

#define UBM_EIG_ELT(CLS, OP)                                                                                   \
  template <class T> CLS<T> operator OP(const CLS<T>& x, const CLS<T>& y) {                                    \
    CLS<T> o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = x.a[i] OP y.a[i]; return o; }
UBM_EIG_ELT(MatT, +)
UBM_EIG_ELT(MatT, -)
UBM_EIG_ELT(ArrT, +)
UBM_EIG_ELT(ArrT, -)
UBM_EIG_ELT(ArrT, *)
UBM_EIG_ELT(ArrT, /)
#undef UBM_EIG_ELT
#define UBM_EIG_SCAL(CLS)                                                                                      \
  template <class T> CLS<T> operator*(double s, const CLS<T>& x) {                                             \
    CLS<T> o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = (T)s * x.a[i]; return o; }            \
  template <class T> CLS<T> operator*(const CLS<T>& x, double s) {                                             \
    CLS<T> o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = x.a[i] * (T)s; return o; }            \
  template <class T> CLS<T> operator/(const CLS<T>& x, double s) {                                             \
    CLS<T> o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = x.a[i] / (T)s; return o; }            \
  template <class T> CLS<T> operator-(const CLS<T>& x) {                                                       \
    CLS<T> o(x.r, x.c); for (size_t i = 0; i < x.a.size(); ++i) o.a[i] = -x.a[i]; return o; }
UBM_EIG_SCAL(MatT)
UBM_EIG_SCAL(ArrT)
#undef UBM_EIG_SCAL
// matrix product: sequential sum over the inner index
template <class T> MatT<T> operator*(const MatT<T>& x, const MatT<T>& y) {
  if (x.c != y.r) throw std::runtime_error("MatT: product dimension mismatch");
  MatT<T> o(x.r, y.c);
  for (int j = 0; j < y.c; ++j)
    for (int i = 0; i < x.r; ++i) {
      T s = T();
      for (int k = 0; k < x.c; ++k) s += x(i, k) * y(k, j);
      o(i, j) = s;
    }
  return o;
}

template <class M>
class LLT {
 public:
  typedef typename M::Scalar T;
  M L;
  LLT(const M& A) : L(A.r, A.c) {               // left-looking Cholesky, A = L L^T
    const int n = A.r;
    for (int j = 0; j < n; ++j) {
      T s = A(j, j);
      for (int k = 0; k < j; ++k) s -= L(j, k) * L(j, k);
      const T ljj = std::sqrt(s);
      L(j, j) = ljj;
      for (int i = j + 1; i < n; ++i) {
        T t = A(i, j);
        for (int k = 0; k < j; ++k) t -= L(i, k) * L(j, k);
        L(i, j) = t / ljj;
      }
    }
  }
  M matrixL() const { return L; }
  M matrixU() const { return L.transpose(); }
  M solve(const M& b) const {                    // L y = b, L^T x = y
    const int n = L.r;
    M x(b);
    for (int col = 0; col < b.c; ++col) {
      for (int i = 0; i < n; ++i) { T s = x(i, col); for (int k = 0; k < i; ++k) s -= L(i, k) * x(k, col); x(i, col) = s / L(i, i); }
      for (int i = n - 1; i >= 0; --i) { T s = x(i, col); for (int k = i + 1; k < n; ++k) s -= L(k, i) * x(k, col); x(i, col) = s / L(i, i); }
    }
    return x;
  }
};
template <class T> LLT<MatT<T>> MatT<T>::llt() const { return LLT<MatT<T>>(*this); }

template <class M>
class Map : public M {
 public:
  Map() {}
  Map(const M& m) : M(m) {}
};

typedef MatT<double> MatrixXd;
typedef VecT<double> VectorXd;
typedef ArrT<double> ArrayXd;
typedef MatT<float> MatrixXf;
typedef VecT<float> VectorXf;

inline Rcpp::Any to_any(const MatT<double>& m) { return Rcpp::Any(Rcpp::NumericMatrix(m.a.data(), m.r, m.c)); }
inline Rcpp::Any to_any(const VecT<double>& v) { return Rcpp::Any(Rcpp::NumericVector(v.a.data(), v.r)); }

}  // namespace Eigen

namespace Rcpp {
template <class T> struct AsImpl;
template <> struct AsImpl<Eigen::Map<Eigen::MatrixXd>> {
  static Eigen::Map<Eigen::MatrixXd> get(const NumericMatrix& m) { Eigen::MatrixXd o(m.nrow(), m.ncol()); o.a = *m.p; return Eigen::Map<Eigen::MatrixXd>(o); }
};
template <> struct AsImpl<Eigen::MatrixXd> {
  static Eigen::MatrixXd get(const NumericMatrix& m) { Eigen::MatrixXd o(m.nrow(), m.ncol()); o.a = *m.p; return o; }
};
template <> struct AsImpl<Eigen::ArrayXd> {
  static Eigen::ArrayXd get(const NumericVector& v) { Eigen::ArrayXd o(v.size(), 1); o.a = *v.p; return o; }
};
template <> struct AsImpl<Eigen::VectorXd> {
  static Eigen::VectorXd get(const NumericVector& v) { Eigen::VectorXd o(v.size()); o.a = *v.p; return o; }
  static Eigen::VectorXd get(const NumericMatrix& m) { Eigen::VectorXd o(m.size()); o.a = *m.p; return o; }
};
template <class T, class S> T as(const S& s) { return AsImpl<T>::get(s); }
inline NumericMatrix wrap(const Eigen::MatT<double>& m) { return NumericMatrix(m.a.data(), m.r, m.c); }
inline NumericVector wrap(const Eigen::VecT<double>& v) { return NumericVector(v.a.data(), v.r); }
inline NumericMatrix wrap(const Eigen::ArrT<double>& m) { return NumericMatrix(m.a.data(), m.r, m.c); }
}  // namespace Rcpp

#endif
