// oracle_models.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header for the parity status).
// CPU restatement of the reference's kernels for C1 (Normal mixture RW-MH) and C2 (MVN RW-MH).
#pragma once
#include "oracle_core.hpp"

namespace orc {

static const double LN_SQRT_2PI = 0.918938533204672741780329736406;  // R's M_LN_SQRT_2PI (nmath dnorm)

// R dnorm(x, mu, sigma, log = TRUE) = -(M_LN_SQRT_2PI + 0.5 * z * z + log(sigma)), z = (x - mu) / sigma
inline double dnorm_log(double x, double mu, double sigma, double log_sigma) {
  double z = (x - mu) / sigma;
  return -((LN_SQRT_2PI + 0.5 * z * z) + log_sigma);
}

// ---------------------------------------------------------------------------------------------------
// C1: mixture target + RW-MH kernels
// ---------------------------------------------------------------------------------------------------
struct MixtureModel {
  struct State { double x; double pdf; };
  int ncomp = 0;
  int coupling = UBM_COUPLING_REFLECTION_MAX;
  double logw[8], mean[8], sd[8], logsd[8];
  double prop_sd = 1, log_prop_sd = 0, init_mean = 0, init_sd = 1;

  explicit MixtureModel(const ubm_model_normal_mixture* p) {
    ncomp = p->ncomp; coupling = p->coupling;
    for (int i = 0; i < ncomp; ++i) {
      logw[i] = ubm_log(p->weights[i]); mean[i] = p->means[i]; sd[i] = p->sds[i]; logsd[i] = ubm_log(p->sds[i]);
    }
    prop_sd = p->proposal_sd; log_prop_sd = ubm_log(p->proposal_sd);
    init_mean = p->init_mean; init_sd = p->init_sd;
  }
  int dim() const { return 1; }
  const double* chain_state(const State& s) const { return &s.x; }
  void copy(State& dst, const State& src) const { dst = src; }

  // README.Rmd:69-72  evals <- log(w) + dnorm(x, mean, sd, log=TRUE); max(evals) + log(sum(exp(evals - max)))
  double target(double x) const {
    double ev[8];
    double mx = -INFINITY;
    for (int i = 0; i < ncomp; ++i) {
      ev[i] = logw[i] + dnorm_log(x, mean[i], sd[i], logsd[i]);
      if (ev[i] > mx) mx = ev[i];
    }
    double s = 0.0;
    for (int i = 0; i < ncomp; ++i) s = s + ubm_exp(ev[i] - mx);
    return mx + ubm_log(s);
  }
  // README.Rmd:80-84
  void rinit(State& s, Draws& D) const {
    s.x = init_mean + init_sd * D.n();
    s.pdf = target(s.x);
  }
  // R/get_mh_kernels.R:29-40 (fast_rmvnorm_chol(1, x, chol): z * chol + x, src/mvnorm.cpp:39-44)
  void single(State& s, Draws& D) const {
    double prop = D.n() * prop_sd + s.x;
    double ppdf = target(prop);
    double logu = ubm_log(D.u());
    if (logu < ppdf - s.pdf) { s.x = prop; s.pdf = ppdf; }
  }
  // R/rnorm_reflectionmax.R:18-39
  bool reflmax(double mu1, double mu2, double sigma, Draws& D, double* x, double* y) const {
    double xdot = D.n();
    double z = (mu1 - mu2) / sigma;
    double normz = sqrt(z * z);
    double e = z / normz;
    double utilde = D.u();
    double a = xdot + z;
    bool accept = ubm_log(utilde) < (-((LN_SQRT_2PI + 0.5 * a * a) + 0.0)) - (-((LN_SQRT_2PI + 0.5 * xdot * xdot) + 0.0));
    double ydot;
    if (accept) ydot = xdot + z;
    else ydot = xdot - 2 * (e * xdot) * e;
    *x = mu1 + sigma * xdot;
    *y = mu2 + sigma * ydot;
    return accept;
  }
  // R/get_max_coupling.R:24-39 with the four closures of R/rnorm_max_coupling.R:17-23
  bool maxcoupling(double mu1, double mu2, double s1, double ls1, double s2, double ls2, Draws& D,
                   double* x, double* y) const {
    double xv = mu1 + s1 * D.n();
    *x = xv;
    if (dnorm_log(xv, mu1, s1, ls1) + ubm_log(D.u()) < dnorm_log(xv, mu2, s2, ls2)) { *y = xv; return true; }
    bool reject = true;
    double yv = 0;
    while (reject && !D.exhausted) {
      yv = mu2 + s2 * D.n();
      reject = dnorm_log(yv, mu2, s2, ls2) + ubm_log(D.u()) < dnorm_log(yv, mu1, s1, ls1);
    }
    *y = yv;
    return false;
  }
  // R/get_mh_kernels.R:42-77 (d = 1 branch) ; inst/reproducebimodal/bimodal.run.R:30-57 for max coupling
  bool coupled(State& s1, State& s2, Draws& D) const {
    double p1, p2;
    bool ident;
    if (coupling == UBM_COUPLING_REFLECTION_MAX) ident = reflmax(s1.x, s2.x, prop_sd, D, &p1, &p2);
    else ident = maxcoupling(s1.x, s2.x, prop_sd, log_prop_sd, prop_sd, log_prop_sd, D, &p1, &p2);
    double pdf1 = target(p1), pdf2;
    if (ident) { p2 = p1; pdf2 = pdf1; } else pdf2 = target(p2);
    double logu = ubm_log(D.u());
    bool acc1 = false, acc2 = false;
    if (std::isfinite(pdf1)) acc1 = logu < pdf1 - s1.pdf;
    if (std::isfinite(pdf2)) acc2 = logu < pdf2 - s2.pdf;
    if (acc1) { s1.x = p1; s1.pdf = pdf1; }
    if (acc2) { s2.x = p2; s2.pdf = pdf2; }
    return ident && acc1 && acc2;
  }
};

// ---------------------------------------------------------------------------------------------------
// C2: MVN target + RW-MH with reflection-max coupled proposals.  Upper-triangular column-major factors.
// y = v^T U  <=>  y_j = sum_{i<=j} v_i U[i,j]   (sequential in i, fma)
// ---------------------------------------------------------------------------------------------------
struct MvnModel {
  struct State { std::vector<double> x; double pdf; };
  int d = 0;
  std::vector<double> mean_pi, cinv_pi, c_prop, cinv_prop, init_mean, init_chol;
  double cst = 0;  // src/mvnorm.cpp:74-75

  explicit MvnModel(const ubm_model_mvn* p) {
    d = p->d;
    auto cp = [&](const double* src, size_t n) { return std::vector<double>(src, src + n); };
    mean_pi = cp(p->mean_pi, d); cinv_pi = cp(p->chol_inv_pi, (size_t)d * d);
    c_prop = cp(p->chol_prop, (size_t)d * d); cinv_prop = cp(p->chol_inv_prop, (size_t)d * d);
    init_mean = cp(p->init_mean, d); init_chol = cp(p->init_chol, (size_t)d * d);
    double halflogdet = 0.0;                                  // :74 diagonal().array().log().sum()
    for (int j = 0; j < d; ++j) halflogdet = halflogdet + ubm_log(cinv_pi[(size_t)j * d + j]);
    cst = -(-halflogdet) - (d * 0.9189385);                   // :75 (truncated constant, verbatim)
  }
  int dim() const { return d; }
  const double* chain_state(const State& s) const { return s.x.data(); }
  void copy(State& dst, const State& src) const { dst = src; }

  // Summation order of y_j = sum_{i<=j} v_i U[i,j] (the reference leaves it to Eigen): one fma chain, sequential in i from
  // 0.0.  This is also, bit for bit, what a chain of FP64 tensor-core MMAs computes (DMMA m8n8k4 accumulates its four
  // products in k order with one rounding each — tools/dmma_probe.cu), which is how the device evaluates it (ubm_mvn.cuh).
  void vec_times_upper(const double* v, const std::vector<double>& U, double* out) const {
    for (int j = 0; j < d; ++j) {
      double acc = 0.0;
      for (int i = 0; i <= j; ++i) acc = ubm_fma(v[i], U[(size_t)j * d + i], acc);
      out[j] = acc;
    }
  }
  // fast_dmvnorm_cholesky_inverse_ — src/mvnorm.cpp:71-86
  double target(const double* x) const {
    std::vector<double> xc(d), w(d);
    for (int j = 0; j < d; ++j) xc[j] = x[j] - mean_pi[j];              // :76-80
    vec_times_upper(xc.data(), cinv_pi, w.data());                      // :81 Cinv^T xc
    return -0.5 * sum_g4_sq(w.data(), d) + cst;                          // :81-84
  }
  // fast_rmvnorm_cholesky_ — src/mvnorm.cpp:30-46 (one rnorm call per column, nsamples = 1)
  void rmvnorm_chol(const double* mean, const std::vector<double>& chol, Draws& D, double* out) const {
    std::vector<double> z(d), t(d);
    for (int j = 0; j < d; ++j) z[j] = D.n();                           // :34-38
    vec_times_upper(z.data(), chol, t.data());                          // :39
    for (int j = 0; j < d; ++j) out[j] = t[j] + mean[j];                // :40-44
  }
  void rinit(State& s, Draws& D) const {                                // scaling...reflectionmaxcoupling.R:88-99
    s.x.assign(d, 0.0);
    rmvnorm_chol(init_mean.data(), init_chol, D, s.x.data());
    s.pdf = target(s.x.data());
  }
  void single(State& s, Draws& D) const {                               // R/get_mh_kernels.R:29-40
    std::vector<double> prop(d);
    rmvnorm_chol(s.x.data(), c_prop, D, prop.data());
    double ppdf = target(prop.data());
    double logu = ubm_log(D.u());
    if (logu < ppdf - s.pdf) { s.x = prop; s.pdf = ppdf; }
  }
  // rmvnorm_reflection_max_coupling_ — src/mvnorm.cpp:185-236
  bool reflmax(const double* mu1, const double* mu2, Draws& D, double* x, double* y) const {
    std::vector<double> diff(d), z(d), xi(d), eta(d), e(d), t(d);
    for (int j = 0; j < d; ++j) diff[j] = mu2[j] - mu1[j];
    vec_times_upper(diff.data(), cinv_prop, z.data());                  // :191 scaled_diff
    for (int j = 0; j < d; ++j) xi[j] = D.n();                          // :193-195
    for (int j = 0; j < d; ++j) z[j] = -z[j];                           // :198
    double normz = sqrt(sum_g4_sq(z.data(), d));                         // :199
    if (normz < 1e-15) {                                                // :200-210
      vec_times_upper(xi.data(), c_prop, t.data());
      for (int j = 0; j < d; ++j) { x[j] = mu1[j] + t[j]; y[j] = x[j]; }
      return true;
    }
    for (int j = 0; j < d; ++j) e[j] = z[j] / normz;                    // :212
    double utilde = D.u();                                              // :213-215
    double edotxi = sum_g4_dot(e.data(), xi.data(), d);                  // :216
    bool ident = false;
    double a = edotxi + normz;
    if (ubm_log(utilde) < (-0.5 * a * a + 0.5 * edotxi * edotxi)) {     // :217
      for (int j = 0; j < d; ++j) eta[j] = xi[j] + z[j];                // :218
      ident = true;
    } else {
      for (int j = 0; j < d; ++j) eta[j] = xi[j] - 2. * edotxi * e[j];  // :221
    }
    vec_times_upper(xi.data(), c_prop, t.data());                       // :224
    for (int j = 0; j < d; ++j) x[j] = mu1[j] + t[j];
    if (ident) {
      for (int j = 0; j < d; ++j) y[j] = x[j];                          // :228-229
    } else {
      vec_times_upper(eta.data(), c_prop, t.data());                    // :225
      for (int j = 0; j < d; ++j) y[j] = mu2[j] + t[j];
    }
    return ident;
  }
  bool coupled(State& s1, State& s2, Draws& D) const {                  // R/get_mh_kernels.R:42-77
    std::vector<double> p1(d), p2(d);
    bool ident = reflmax(s1.x.data(), s2.x.data(), D, p1.data(), p2.data());
    double pdf1 = target(p1.data()), pdf2;
    if (ident) { p2 = p1; pdf2 = pdf1; } else pdf2 = target(p2.data());
    double logu = ubm_log(D.u());
    bool acc1 = false, acc2 = false;
    if (std::isfinite(pdf1)) acc1 = logu < pdf1 - s1.pdf;
    if (std::isfinite(pdf2)) acc2 = logu < pdf2 - s2.pdf;
    if (acc1) { s1.x = p1; s1.pdf = pdf1; }
    if (acc2) { s2.x = p2; s2.pdf = pdf2; }
    return ident && acc1 && acc2;
  }
};

}  // namespace orc
