// oracle_blasso.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header for the parity status).
// CPU restatement of the reference's Bayesian-lasso Gibbs sampler and its coupling:
//   R/bayesianlasso.R:22-98 (get_blasso: rinit, gibbs_kernel, coupled_gibbs_kernel)
//   src/blassoutil.cpp:17-41 (blassoconditional), :45-88 (blassoconditional_coupled)
//   with fast_rmvnorm_ / rmvnorm_max_coupling_ (src/mvnorm.cpp:9-26, :91-129), rinversegamma[_coupled]
//   (R/invgamma_couplings.R:17-36) and rinvgaussian[_coupled] (src/inversegaussian.cpp:6-63) from oracle_scalar.hpp.
// Chain state = c(beta[p], tau2[p], sigma2) (R/bayesianlasso.R:32,56).
// Canonical arithmetic (the engine's contract, matched bit for bit by ubm_blasso.cuh): A^{-1} through the Cholesky factor
// (A = L L', Linv by forward substitution with reciprocal diagonal, A^{-1} = Linv' Linv; the reference calls Eigen's
// .inverse()); A^{-1} XtY and X beta as fma chains in index order; the residual sum of squares as 512 interleaved chains
// (observation i in class i mod 512), a 32-wide xor-butterfly inside each group of 32 classes, then the 16 group sums added
// in order; Gamma draws by Marsaglia-Tsang (oracle_scalar.hpp).
#pragma once
#include <cmath>
#include "oracle_core.hpp"
#include "oracle_pg.hpp"
#include "oracle_scalar.hpp"

namespace orc {

struct BlassoModel {
  struct State { std::vector<double> x; };     // beta[p], tau2[p], sigma2
  int n = 0, p = 0;
  double lambda2 = 0, alpha1 = 0, lgam = 0;
  std::vector<double> X, Y, XtX, XtY;

  explicit BlassoModel(const ubm_model_blasso* s) {
    n = s->n; p = s->p;
    lambda2 = s->lambda * s->lambda;                                   // R/bayesianlasso.R:28
    alpha1 = (n - 1) / 2.0 + p / 2.0;                                  // :27
    lgam = std::lgamma(alpha1);
    X.assign(s->X, s->X + (size_t)n * p);
    Y.assign(s->Y, s->Y + n);
    XtX.assign((size_t)p * p, 0.0); XtY.assign(p, 0.0);                // :25-26
    for (int a = 0; a < p; ++a) {
      for (int b = 0; b <= a; ++b) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc = ubm_fma(X[(size_t)a * n + i], X[(size_t)b * n + i], acc);
        XtX[(size_t)a * p + b] = acc; XtX[(size_t)b * p + a] = acc;
      }
      double t = 0.0;
      for (int i = 0; i < n; ++i) t = ubm_fma(X[(size_t)a * n + i], Y[i], t);
      XtY[a] = t;
    }
  }
  int dim() const { return 2 * p + 1; }
  const double* chain_state(const State& s) const { return s.x.data(); }
  void copy(State& dst, const State& src) const { dst = src; }
  void rinit(State& s, Draws&) const {                                  // :30-32 (no draws)
    s.x.assign(2 * p + 1, 0.0);
    for (int j = 0; j < p; ++j) s.x[p + j] = 1.0;
    s.x[2 * p] = 1.0;
  }

  struct Cond { std::vector<double> mean, Ls; };                        // mean of beta, lower Cholesky factor of sigma2 A^{-1}
  Cond conditional(const double* tau2, double sigma2) const {           // src/blassoutil.cpp:21-30
    std::vector<double> A(XtX);
    for (int j = 0; j < p; ++j) A[(size_t)j * p + j] = XtX[(size_t)j * p + j] + 1.0 / tau2[j];
    bool ok = true;
    PgLogisticModel::chol_lower(A, p, &ok);
    std::vector<double> Ainv = PgLogisticModel::ltl(PgLogisticModel::tri_inverse(A, p), p);
    Cond c;
    c.mean.assign(p, 0.0);
    for (int j = 0; j < p; ++j) {
      double acc = 0.0;
      for (int k = 0; k < p; ++k) acc = ubm_fma(Ainv[(size_t)k * p + j], XtY[k], acc);
      c.mean[j] = acc;
    }
    c.Ls = Ainv;
    for (auto& v : c.Ls) v = v * sigma2;
    PgLogisticModel::chol_lower(c.Ls, p, &ok);                          // LLT inside fast_rmvnorm_ / rmvnorm_max_coupling_
    return c;
  }
  // x = mean + z U, U = Ls' (src/mvnorm.cpp:18-24): x_j = sum_{i <= j} z_i Ls(j, i) + mean_j
  void draw_beta(const Cond& c, Draws& D, double* out) const {
    std::vector<double> z(p);
    for (int j = 0; j < p; ++j) z[j] = D.n();
    for (int j = 0; j < p; ++j) {
      double acc = 0.0;
      for (int i = 0; i <= j; ++i) acc = ubm_fma(z[i], c.Ls[(size_t)i * p + j], acc);
      out[j] = acc + c.mean[j];
    }
  }
  double dmvnorm(const double* x, const Cond& c) const {               // fast_dmvnorm_ (src/mvnorm.cpp:49-67)
    double halflogdet = 0.0;
    for (int j = 0; j < p; ++j) halflogdet = halflogdet + ubm_log(c.Ls[(size_t)j * p + j]);
    const double cst = -(halflogdet) - (p * 0.9189385);
    std::vector<double> y(p);
    for (int i = 0; i < p; ++i) {
      double s = x[i] - c.mean[i];
      for (int k = 0; k < i; ++k) s = ubm_fma(-c.Ls[(size_t)k * p + i], y[k], s);
      y[i] = s * (1.0 / c.Ls[(size_t)i * p + i]);
    }
    double sq = 0.0;
    for (int i = 0; i < p; ++i) sq = ubm_fma(y[i], y[i], sq);
    return -0.5 * sq + cst;
  }
  double rss(const double* beta) const {                               // sum((Y - X beta)^2), src/blassoutil.cpp:33
    double part[512];
    for (int q = 0; q < 512; ++q) part[q] = 0.0;
    for (int i = 0; i < n; ++i) {
      double xb = 0.0;
      for (int j = 0; j < p; ++j) xb = ubm_fma(X[(size_t)j * n + i], beta[j], xb);
      const double r = Y[i] - xb;
      part[i & 511] = ubm_fma(r, r, part[i & 511]);
    }
    double tot = 0.0;
    for (int w = 0; w < 16; ++w) tot = tot + ubm_butterfly32(part + 32 * w);
    return tot;
  }
  double bdb(const double* beta, const double* tau2) const {            // :34-37
    double acc = 0.0;
    for (int j = 0; j < p; ++j) acc = acc + beta[j] * beta[j] / tau2[j];
    return acc;
  }
  double cst_of(double rate) const { return alpha1 * ubm_log(rate) - lgam; }

  void single(State& s, Draws& D) const {                               // gibbs_kernel, R/bayesianlasso.R:34-50
    double* beta = s.x.data(); double* tau2 = beta + p; double& sigma2 = s.x[2 * p];
    const Cond c = conditional(tau2, sigma2);
    draw_beta(c, D, beta);
    const double rate = 0.5 * (rss(beta) + bdb(beta, tau2));
    sigma2 = 1.0 / rgamma_mt(alpha1, rate, D);                          // rinversegamma(1, alpha1, rate)
    const double sl = sqrt(lambda2 * sigma2);
    for (int j = 0; j < p; ++j) tau2[j] = 1.0 / rinvgaussian1(sl / fabs(beta[j]), lambda2, D);
  }
  bool coupled(State& s1, State& s2, Draws& D) const {                  // coupled_gibbs_kernel, :52-98
    double* b1 = s1.x.data(); double* t1 = b1 + p; double& g1 = s1.x[2 * p];
    double* b2 = s2.x.data(); double* t2 = b2 + p; double& g2 = s2.x[2 * p];
    bool same = (g1 == g2);
    for (int j = 0; j < p; ++j) if (t1[j] != t2[j]) same = false;
    double rate1, rate2;
    if (same) {                                                         // :60-67
      const Cond c = conditional(t1, g1);
      draw_beta(c, D, b1);
      for (int j = 0; j < p; ++j) b2[j] = b1[j];
      rate1 = 0.5 * (rss(b1) + bdb(b1, t1));
      rate2 = rate1;
    } else {                                                            // :68-76, src/blassoutil.cpp:45-88
      const Cond c1 = conditional(t1, g1), c2 = conditional(t2, g2);
      std::vector<double> x1(p), x2(p);
      draw_beta(c1, D, x1.data());                                      // src/mvnorm.cpp:96
      const double logu = ubm_log(D.u());                               // :103-106
      if (dmvnorm(x1.data(), c1) + logu < dmvnorm(x1.data(), c2)) x2 = x1;   // :107
      else {
        bool accept = false;
        while (!accept && !D.exhausted) {                               // :113-123
          draw_beta(c2, D, x2.data());
          const double lu = ubm_log(D.u());
          if (dmvnorm(x2.data(), c2) + lu > dmvnorm(x2.data(), c1)) accept = true;
        }
      }
      for (int j = 0; j < p; ++j) { b1[j] = x1[j]; b2[j] = x2[j]; }
      rate1 = 0.5 * (rss(b1) + bdb(b1, t1));
      rate2 = 0.5 * (rss(b2) + bdb(b2, t2));
    }
    double ig1, ig2;                                                    // :77-81 rinversegamma_coupled
    positive_max_coupling(1, alpha1, alpha1, rate1, rate2, cst_of(rate1), cst_of(rate2), D, &ig1, &ig2);
    g1 = ig1; g2 = ig2;
    const double sl1 = sqrt(lambda2 * g1), sl2 = sqrt(lambda2 * g2);
    for (int j = 0; j < p; ++j) {                                       // :85-96
      if (g1 == g2 && b1[j] == b2[j]) {
        const double it = rinvgaussian1(sl1 / fabs(b1[j]), lambda2, D);
        t1[j] = 1.0 / it; t2[j] = t1[j];
      } else {
        double u1, u2;
        positive_max_coupling(2, sl1 / fabs(b1[j]), sl2 / fabs(b2[j]), lambda2, lambda2, 0.0, 0.0, D, &u1, &u2);
        t1[j] = 1.0 / u1; t2[j] = 1.0 / u2;
      }
    }
    return s1.x == s2.x;                                                // :99
  }
};

}  // namespace orc
