// ref_eigen_capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" doors into the reference's unmodified Eigen-using functions (src/mvnorm.cpp, src/logisticregression.cpp,
// src/vs_mlikelihood.cpp), built against the dense-algebra stand-in of RcppEigen.h.  Matrices cross this boundary
// COLUMN-major (R layout), as in include/ubm.h.
#include <RcppEigen.h>
#include "ref_runtime.h"
#include "mvnorm.h"   /* the reference's own header, found through -I$(REF)/src */
using namespace Rcpp;

NumericMatrix sigma_(const NumericMatrix& X, const NumericVector& w);                                            // src/logisticregression.cpp:5
List m_sigma_function_(const Eigen::Map<Eigen::MatrixXd>& omega, const Eigen::Map<Eigen::MatrixXd>& X,
                       const Eigen::Map<Eigen::MatrixXd>& invB, const Eigen::Map<Eigen::VectorXd>& KTkappaplusinvBtimesb);   // :32
double marginal_likelihood_c_2(Eigen::VectorXf selection, const Eigen::MatrixXf& X, const Eigen::VectorXf& Y, double Y2, double g);   // src/vs_mlikelihood.cpp:5

List blassoconditional(const Eigen::VectorXd& Y, const Eigen::MatrixXd& X, const Eigen::VectorXd& XtY, const Eigen::MatrixXd& XtX,
                       const NumericVector tau2, const double sigma2);                                           // src/blassoutil.cpp:17

namespace {
Eigen::MatrixXd emat(const double* a, int r, int c) { Eigen::MatrixXd m(r, c); m.a.assign(a, a + (size_t)r * c); return m; }
Eigen::VectorXd evec(const double* a, int n) { Eigen::VectorXd v(n); v.a.assign(a, a + n); return v; }
void put(const NumericMatrix& m, double* out) { std::copy(m.begin(), m.end(), out); }
}  // namespace

extern "C" {

// blassoconditional (src/blassoutil.cpp:17-41): beta [p], norm, betaDbeta
void ref_blassoconditional(int32_t n, int32_t p, const double* Y, const double* X, const double* XtY, const double* XtX,
                           const double* tau2, double sigma2, double* beta, double* norm_bdb) {
  List r = blassoconditional(evec(Y, n), emat(X, n, p), evec(XtY, p), emat(XtX, p, p), NumericVector(tau2, p), sigma2);
  NumericVector b = r["beta"];
  for (int j = 0; j < p; ++j) beta[j] = b(j);
  norm_bdb[0] = (double)r["norm"]; norm_bdb[1] = (double)r["betaDbeta"];
}

// out: nsamples x d column-major
void ref_fast_rmvnorm(int32_t nsamples, int32_t d, const double* mean, const double* covariance, double* out) {
  put(fast_rmvnorm_(nsamples, NumericVector(mean, d), NumericMatrix(covariance, d, d)), out);
}
void ref_fast_rmvnorm_cholesky(int32_t nsamples, int32_t d, const double* mean, const double* cholesky, double* out) {
  put(fast_rmvnorm_cholesky_(nsamples, NumericVector(mean, d), emat(cholesky, d, d)), out);
}
// x: n x d column-major
void ref_fast_dmvnorm(int32_t n, int32_t d, const double* x, const double* mean, const double* covariance, double* out) {
  NumericVector r = fast_dmvnorm_(NumericMatrix(x, n, d), NumericVector(mean, d), NumericMatrix(covariance, d, d));
  for (int i = 0; i < n; ++i) out[i] = r(i);
}
void ref_fast_dmvnorm_cholesky_inverse(int32_t n, int32_t d, const double* x, const double* mean, const double* chol_inv, double* out) {
  NumericVector r = fast_dmvnorm_cholesky_inverse_(NumericMatrix(x, n, d), NumericVector(mean, d), emat(chol_inv, d, d));
  for (int i = 0; i < n; ++i) out[i] = r(i);
}
// xy: d x 2 column-major (x then y)
void ref_rmvnorm_max_coupling(int32_t d, const double* mu1, const double* mu2, const double* Sigma1, const double* Sigma2,
                              double* xy, int32_t* identical, int32_t* nattempts) {
  List r = rmvnorm_max_coupling_(NumericVector(mu1, d), NumericVector(mu2, d), NumericMatrix(Sigma1, d, d), NumericMatrix(Sigma2, d, d));
  NumericMatrix m = r["xy"];
  put(m, xy);
  *identical = (int)r["identical"];
  *nattempts = (int)r["nattempts"];
}
void ref_rmvnorm_max_coupling_cholesky(int32_t d, const double* mu1, const double* mu2, const double* chol1, const double* chol2,
                                       const double* chol_inv1, const double* chol_inv2, double* xy, int32_t* identical) {
  List r = rmvnorm_max_coupling_cholesky(NumericVector(mu1, d), NumericVector(mu2, d), emat(chol1, d, d), emat(chol2, d, d),
                                         emat(chol_inv1, d, d), emat(chol_inv2, d, d));
  NumericMatrix m = r["xy"];
  put(m, xy);
  *identical = (int)r["identical"];
}
void ref_rmvnorm_reflection_max_coupling(int32_t d, const double* mu1, const double* mu2, const double* chol, const double* chol_inv,
                                         double* xy, int32_t* identical) {
  List r = rmvnorm_reflection_max_coupling_(evec(mu1, d), evec(mu2, d), emat(chol, d, d), emat(chol_inv, d, d));
  NumericMatrix m = r["xy"];
  put(m, xy);
  *identical = (int)r["identical"];
}
// X: n x p column-major; out: p x p
void ref_sigma(int32_t n, int32_t p, const double* X, const double* w, double* out) {
  put(sigma_(NumericMatrix(X, n, p), NumericVector(w, n)), out);
}
// outputs: m [p], Sigma_inverse, Cholesky_inverse (= L), Cholesky (= L^{-1}): p x p column-major each
void ref_m_sigma_function(int32_t n, int32_t p, const double* omega, const double* X, const double* invB, const double* ktk,
                          double* m, double* Sigma_inverse, double* Cholesky_inverse, double* Cholesky) {
  Eigen::Map<Eigen::MatrixXd> om(emat(omega, n, 1)), Xm(emat(X, n, p)), iB(emat(invB, p, p));
  Eigen::Map<Eigen::VectorXd> kt(evec(ktk, p));
  List r = m_sigma_function_(om, Xm, iB, kt);
  NumericVector mv = r["m"];
  for (int i = 0; i < p; ++i) m[i] = mv(i);
  NumericMatrix a = r["Sigma_inverse"], b = r["Cholesky_inverse"], c = r["Cholesky"];
  put(a, Sigma_inverse); put(b, Cholesky_inverse); put(c, Cholesky);
}
// selection [p] of 0/1, X n x p column-major (doubles; converted to float32 as the R layer's call does)
double ref_marginal_likelihood_c_2(int32_t n, int32_t p, const double* selection, const double* X, const double* Y, double Y2, double g) {
  Eigen::VectorXf sel(p), Yf(n);
  Eigen::MatrixXf Xf(n, p);
  for (int j = 0; j < p; ++j) sel(j) = (float)selection[j];
  for (int i = 0; i < n; ++i) Yf(i) = (float)Y[i];
  for (size_t q = 0; q < (size_t)n * p; ++q) Xf.a[q] = (float)X[q];
  return marginal_likelihood_c_2(sel, Xf, Yf, Y2, g);
}

}  // extern "C"
