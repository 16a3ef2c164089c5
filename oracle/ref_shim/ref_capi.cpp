// ref_capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" doors into the reference's own, unmodified C++ functions (compiled from /root/reference/src by
// oracle/ref_shim/Makefile against the stand-in headers of this directory).  tests/ call these through ctypes
// (oracle/ref_py.py) to check the oracle restatement against the reference itself on identical draws.
// Layout convention of this file: matrices cross the C boundary ROW-major ([row][col], as the oracle stores chains);
// they are transposed into the column-major Rcpp stand-ins here.  Lattices are passed column-major as in R.
#include <cstring>
#include <Rcpp.h>
#include "ref_runtime.h"
using namespace Rcpp;

// ---- declarations of the reference functions (signatures as in /root/reference/src/*.cpp) ------------------------
int ising_sum_(const IntegerMatrix& state);                                                            // src/ising.cpp:11
IntegerMatrix ising_gibbs_sweep_(IntegerMatrix state, NumericVector proba_beta);                       // :43
List ising_coupled_gibbs_sweep_(IntegerMatrix state1, IntegerMatrix state2, NumericVector proba_beta); // :67
double estimator_bin_(List c_chains, int component, double lower, double upper, int k, int m, int lag);  // src/estimator_bin.cpp:10
IntegerVector sample_pair01(const NumericVector& selection);                                           // src/sample_pairs01.cpp:12
List c_chains_to_measure_as_list_(const List& c_chains, int k, int m);                                 // src/c_chains_to_measure.cpp:38
NumericMatrix prune_measure_(const NumericMatrix& df);                                                 // src/prune.cpp:6
double logcosh(double x);                                                                              // src/logisticregressioncoupling.cpp:9
NumericVector xbeta_(const NumericMatrix& X, const NumericVector& beta);                               // :22
double rpg_devroye(int n, double z);                                                                   // :34
NumericMatrix w_rejsamplerC(const NumericVector& beta1, const NumericVector& beta2, const NumericMatrix& X);   // :42
NumericMatrix w_max_couplingC(const NumericVector& beta1, const NumericVector& beta2, const NumericMatrix& X); // :79
NumericVector rinvgaussian_c(int n, double mu, double lambda);                                         // src/inversegaussian.cpp:6
NumericVector rinvgaussian_coupled_c(double mu1, double mu2, double lambda1, double lambda2);          // :32

namespace {
NumericMatrix from_rowmajor(const double* a, int nrow, int ncol) {
  NumericMatrix m(nrow, ncol);
  for (int i = 0; i < nrow; ++i) for (int j = 0; j < ncol; ++j) m(i, j) = a[(size_t)i * ncol + j];
  return m;
}
List make_c_chains(const double* samples1, int nrow1, const double* samples2, int nrow2, int dim, double meetingtime) {
  return List::create(Named("samples1") = from_rowmajor(samples1, nrow1, dim),
                      Named("samples2") = from_rowmajor(samples2, nrow2, dim), Named("meetingtime") = meetingtime);
}
}  // namespace

extern "C" {

const char* ref_flavour() {
#ifdef UBM_REF_UBMNUM
  return "ubmnum";
#else
  return "libm";
#endif
}

int32_t ref_ising_sum(const int32_t* state, int32_t size) { return ising_sum_(IntegerMatrix(state, size, size)); }

void ref_ising_gibbs_sweep(int32_t* state, int32_t size, const double* proba_beta) {
  IntegerMatrix s(state, size, size);
  IntegerMatrix out = ising_gibbs_sweep_(s, NumericVector(proba_beta, 5));
  memcpy(state, out.p->data(), sizeof(int32_t) * size * size);
}

void ref_ising_coupled_gibbs_sweep(int32_t* state1, int32_t* state2, int32_t size, const double* proba_beta) {
  IntegerMatrix s1(state1, size, size), s2(state2, size, size);
  List out = ising_coupled_gibbs_sweep_(s1, s2, NumericVector(proba_beta, 5));
  IntegerMatrix o1 = out["state1"], o2 = out["state2"];
  memcpy(state1, o1.p->data(), sizeof(int32_t) * size * size);
  memcpy(state2, o2.p->data(), sizeof(int32_t) * size * size);
}

// component is 1-based as in R
double ref_estimator_bin(const double* samples1, int32_t nrow1, const double* samples2, int32_t nrow2, int32_t dim,
                         double meetingtime, int32_t component, double lower, double upper, int32_t k, int32_t m,
                         int32_t lag) {
  return estimator_bin_(make_c_chains(samples1, nrow1, samples2, nrow2, dim, meetingtime), component, lower, upper, k, m, lag);
}

void ref_sample_pair01(const double* selection, int32_t l, int32_t* out2) {
  IntegerVector r = sample_pair01(NumericVector(selection, l));
  out2[0] = r(0); out2[1] = r(1);
}

// returns the number of atoms; atoms [size][dim] row-major, weights [size], mcmc [size] (caller sizes them with cap rows)
int32_t ref_c_chains_to_measure(const double* samples1, int32_t nrow1, const double* samples2, int32_t nrow2, int32_t dim,
                                double meetingtime, int32_t k, int32_t m, int32_t cap, double* atoms, double* weights,
                                int32_t* mcmc) {
  List r = c_chains_to_measure_as_list_(make_c_chains(samples1, nrow1, samples2, nrow2, dim, meetingtime), k, m);
  NumericMatrix a = r["atoms"];
  NumericVector w = r["weights"];
  LogicalVector part = r["MCMC"];
  const int size = a.nrow();
  if (size > cap) return -size;
  for (int i = 0; i < size; ++i) {
    for (int j = 0; j < dim; ++j) atoms[(size_t)i * dim + j] = a(i, j);
    weights[i] = w(i); mcmc[i] = part(i);
  }
  return size;
}

// df, out: [nrow][ncol] row-major; returns the number of rows kept
int32_t ref_prune_measure(const double* df, int32_t nrow, int32_t ncol, double* out) {
  NumericMatrix r = prune_measure_(from_rowmajor(df, nrow, ncol));
  for (int i = 0; i < r.nrow(); ++i) for (int j = 0; j < ncol; ++j) out[(size_t)i * ncol + j] = r(i, j);
  return r.nrow();
}

double ref_logcosh(double x) { return logcosh(x); }

// X: n x p COLUMN-major (R layout, as in ubm_model_pg_logistic)
void ref_xbeta(const double* X, int32_t n, int32_t p, const double* beta, double* out) {
  NumericVector r = xbeta_(NumericMatrix(X, n, p), NumericVector(beta, p));
  for (int i = 0; i < n; ++i) out[i] = r(i);
}
void ref_rpg_devroye(const double* z, int64_t count, double* out) {
  for (int64_t i = 0; i < count; ++i) out[i] = rpg_devroye(1, z[i]);
}
// w: [n][2] row-major (column 0 = chain 1)
void ref_w_max_coupling(const double* beta1, const double* beta2, const double* X, int32_t n, int32_t p, double* w) {
  NumericMatrix r = w_max_couplingC(NumericVector(beta1, p), NumericVector(beta2, p), NumericMatrix(X, n, p));
  for (int i = 0; i < n; ++i) { w[2 * i] = r(i, 0); w[2 * i + 1] = r(i, 1); }
}
void ref_w_rejsampler(const double* beta1, const double* beta2, const double* X, int32_t n, int32_t p, double* w) {
  NumericMatrix r = w_rejsamplerC(NumericVector(beta1, p), NumericVector(beta2, p), NumericMatrix(X, n, p));
  for (int i = 0; i < n; ++i) { w[2 * i] = r(i, 0); w[2 * i + 1] = r(i, 1); }
}

void ref_rinvgaussian(int32_t n, double mu, double lambda, double* out) {
  NumericVector r = rinvgaussian_c(n, mu, lambda);
  for (int i = 0; i < n; ++i) out[i] = r(i);
}
void ref_rinvgaussian_coupled(double mu1, double mu2, double lambda1, double lambda2, double* out2) {
  NumericVector r = rinvgaussian_coupled_c(mu1, mu2, lambda1, lambda2);
  out2[0] = r(0); out2[1] = r(1);
}

}  // extern "C"
