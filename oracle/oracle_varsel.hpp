// oracle_varsel.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header for the parity status).
// CPU restatement of the reference's spike-and-slab variable selection sampler (C5):
//   R/variableselection.R:3-16 (prior, rinit), :17-55 (single kernel), :57-148 (coupled kernel),
//   :215-279 (unbiasedestimator_vs; same estimator as the generic lag-1 driver used here)
//   src/vs_mlikelihood.cpp:5-28 (marginal_likelihood_c_2), src/sample_pairs01.cpp:12-42, R/coupled_pairs01.R:3-40
//
// Arithmetic pinned where the reference leaves it to Eigen / R internals (SURVEY.md §7 "hard parts", Appendix A):
//   * marginal likelihood in FP64 (the reference converts to float32 and calls Eigen's .inverse(), which is not
//     reproducible without Eigen): G = X'X and v = X'Y are formed once with sequential fma sums; for a selection
//     gamma, Y'X_g (X_g'X_g)^-1 X_g'Y = |L^-1 v_g|^2 with L the left-looking Cholesky factor of G[g,g];
//   * sample.int(p, 1) = floor(U p); sample(prob = w) = inversion of the cumulative weights in natural index order
//     with one uniform, the cumulative weight after cA positions of weight wA and cB of weight wB being
//     cA * wA + cB * wB (products, not the reference's running sums — same distribution, no sequential dependence);
//   * rinit: R's no-replacement sample (x[j] = x[--n]) with floor(U n) indices; the Bernoulli uniforms come first.
#pragma once
#include <algorithm>
#include "oracle_core.hpp"

namespace orc {

struct VarselModel {
  struct State {
    std::vector<uint8_t> g;       // gamma
    std::vector<double> gd;       // gamma as doubles (what h sees)
    double pdf = 0;
    int s = 0;
  };
  int n = 0, p = 0, s0 = 0;
  double g_ = 0, kappa = 0, prop_flip = 0.5, Y2 = 0, lg1 = 0, gg = 0, logp = 0;
  std::vector<double> G, v;       // G p x p (row-major symmetric), v [p]

  explicit VarselModel(const ubm_model_varsel* m) {
    n = m->n; p = m->p; s0 = m->s0; g_ = m->g; kappa = m->kappa; prop_flip = m->proportion_singleflip;
    G.assign((size_t)p * p, 0.0); v.assign(p, 0.0);
    const double* X = m->X;       // n x p column-major
    for (int a = 0; a < p; ++a) {
      for (int b = 0; b <= a; ++b) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc = ubm_fma(X[(size_t)a * n + i], X[(size_t)b * n + i], acc);
        G[(size_t)a * p + b] = acc; G[(size_t)b * p + a] = acc;
      }
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc = ubm_fma(X[(size_t)a * n + i], m->Y[i], acc);
      v[a] = acc;
    }
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = ubm_fma(m->Y[i], m->Y[i], acc);
    Y2 = acc;                                            // R/variableselection.R:6
    lg1 = ubm_log(g_ + 1.); gg = g_ / (g_ + 1.); logp = ubm_log((double)p);
  }
  int dim() const { return p; }
  const double* chain_state(const State& s) const { return s.gd.data(); }
  void copy(State& dst, const State& src) const { dst = src; }

  double prior(int s) const {                            // R/variableselection.R:8-11
    if (s > s0) return -INFINITY;
    return -kappa * (double)s * logp;
  }
  // marginal_likelihood_c_2 — src/vs_mlikelihood.cpp:5-28
  double ml(const std::vector<uint8_t>& g) const {
    std::vector<int> idx;
    for (int j = 0; j < p; ++j) if (g[j]) idx.push_back(j);
    const int s = (int)idx.size();
    double l = 0.0;
    if (s > 0) {
      std::vector<double> L((size_t)s * s, 0.0), y(s), rinv(s);
      for (int a = 0; a < s; ++a) for (int b = 0; b <= a; ++b) L[(size_t)a * s + b] = G[(size_t)idx[a] * p + idx[b]];
      // Cholesky, row-major lower; columns are scaled by the reciprocal of the pivot (as LAPACK's dpotf2 does)
      for (int j = 0; j < s; ++j) {
        double d = L[(size_t)j * s + j];
        for (int k = 0; k < j; ++k) d = ubm_fma(-L[(size_t)j * s + k], L[(size_t)j * s + k], d);
        const double ljj = sqrt(d);
        L[(size_t)j * s + j] = ljj;
        rinv[j] = 1.0 / ljj;
        for (int i = j + 1; i < s; ++i) {
          double t = L[(size_t)i * s + j];
          for (int k = 0; k < j; ++k) t = ubm_fma(-L[(size_t)i * s + k], L[(size_t)j * s + k], t);
          L[(size_t)i * s + j] = t * rinv[j];
        }
      }
      for (int i = 0; i < s; ++i) {                      // L y = v_g
        double t = v[idx[i]];
        for (int k = 0; k < i; ++k) t = ubm_fma(-y[k], L[(size_t)i * s + k], t);
        y[i] = t * rinv[i];
      }
      for (int i = 0; i < s; ++i) l = ubm_fma(y[i], y[i], l);
    }
    return -((double)s + 1.) / 2. * lg1 - (double)n / 2. * ubm_log(Y2 - gg * l);   // :26
  }
  void sync(State& st) const {
    st.s = 0;
    st.gd.assign(p, 0.0);
    for (int j = 0; j < p; ++j) { st.s += st.g[j]; st.gd[j] = (double)st.g[j]; }
  }
  // weighted index by inversion in natural order: positions of class A (weight wA) and class B (weight wB)
  int inv2(const std::vector<uint8_t>& A, double wA, const std::vector<uint8_t>* B, double wB, double u) const {
    long cA = 0, cB = 0;
    int last = -1;
    for (int k = 0; k < p; ++k) {
      const bool a = A[k] != 0, b = B && (*B)[k] != 0;
      if (!a && !b) continue;
      if (a) ++cA; else ++cB;
      if ((a && wA > 0) || (b && wB > 0)) last = k;
      if ((double)cA * wA + (double)cB * wB > u) return k;
    }
    return last;
  }
  // rinit — R/variableselection.R:12-16
  void rinit(State& st, Draws& D) const {
    const int k = std::min(s0, p);
    std::vector<double> ub(k);
    for (int t = 0; t < k; ++t) ub[t] = D.u();
    std::vector<int> x(p);
    for (int j = 0; j < p; ++j) x[j] = j;
    st.g.assign(p, 0);
    int nn = p;
    for (int t = 0; t < k; ++t) {
      int j = (int)((double)nn * D.u());
      if (j >= nn) j = nn - 1;
      st.g[x[j]] = (ub[t] < 0.5) ? 1 : 0;
      x[j] = x[--nn];
    }
    sync(st);
    st.pdf = ml(st.g) + prior(st.s);                     // R/variableselection.R:154-155
  }
  // sample_pair01 — src/sample_pairs01.cpp:12-42: (zero index, one index)
  void sample_pair01(const State& st, Draws& D, int* i0, int* i1) const {
    const double u0 = D.u(), u1 = D.u();
    const double w0 = 1. / (double)(p - st.s), w1 = 1. / (double)st.s;
    std::vector<uint8_t> zeros(p), ones(p);
    for (int j = 0; j < p; ++j) { zeros[j] = !st.g[j]; ones[j] = st.g[j]; }
    *i0 = inv2(zeros, w0, nullptr, 0.0, u0);
    *i1 = inv2(ones, w1, nullptr, 0.0, u1);
  }
  void try_move(State& st, const std::vector<uint8_t>& prop, int sprop, double logu_or_nan, Draws& D, bool own_u) const {
    const double pp = prior(sprop);
    if (std::isfinite(pp)) {
      const double ppdf = ml(prop) + pp;
      const double logu = own_u ? ubm_log(D.u()) : logu_or_nan;
      if (logu < ppdf - st.pdf) { st.g = prop; st.pdf = ppdf; sync(st); }
    }
  }
  // single_kernel — R/variableselection.R:17-55
  void single(State& st, Draws& D) const {
    if (D.u() < prop_flip) {
      int c = (int)((double)p * D.u());
      if (c >= p) c = p - 1;
      std::vector<uint8_t> prop = st.g;
      prop[c] = 1 - prop[c];
      try_move(st, prop, st.s + (prop[c] ? 1 : -1), 0.0, D, true);
    } else if (st.s > 0 && st.s < p) {
      int i0, i1;
      sample_pair01(st, D, &i0, &i1);
      std::vector<uint8_t> prop = st.g;
      prop[i0] = 1; prop[i1] = 0;
      try_move(st, prop, st.s, 0.0, D, true);
    }
  }
  // one half of coupled_pairs01 — R/coupled_pairs01.R:5-21 (ones) / :23-38 (zeros): `m1`, `m2` mark the candidate
  // positions of each chain, c1, c2 their counts
  void coupled_index(const std::vector<uint8_t>& m1, const std::vector<uint8_t>& m2, int c1, int c2, Draws& D,
                     int* x1, int* x2) const {
    const double a = 1. / (double)c1, b = 1. / (double)c2;
    const double mn = a < b ? a : b;
    std::vector<uint8_t> common(p), only1(p), only2(p);
    int C = 0;
    for (int j = 0; j < p; ++j) {
      common[j] = m1[j] && m2[j]; only1[j] = m1[j] && !m2[j]; only2[j] = m2[j] && !m1[j];
      C += common[j];
    }
    const double alpha = (double)C * mn;
    if (D.u() < alpha) {
      const int x = inv2(common, mn / alpha, nullptr, 0.0, D.u());
      *x1 = x; *x2 = x;
    } else {
      *x1 = inv2(common, (a - mn) / (1. - alpha), &only1, a / (1. - alpha), D.u());
      *x2 = inv2(common, (b - mn) / (1. - alpha), &only2, b / (1. - alpha), D.u());
    }
  }
  // coupled_kernel — R/variableselection.R:57-148; identical <=> states equal (:187, :258)
  bool coupled(State& s1, State& s2, Draws& D) const {
    if (D.u() < prop_flip) {
      int c = (int)((double)p * D.u());
      if (c >= p) c = p - 1;
      std::vector<uint8_t> p1 = s1.g, p2 = s2.g;
      p1[c] = 1 - p1[c]; p2[c] = 1 - p2[c];
      const int sp1 = s1.s + (p1[c] ? 1 : -1), sp2 = s2.s + (p2[c] ? 1 : -1);
      const double logu = ubm_log(D.u());                // :67 always drawn
      try_move(s1, p1, sp1, logu, D, false);
      try_move(s2, p2, sp2, logu, D, false);
    } else {
      const bool sw1 = s1.s > 0 && s1.s < p, sw2 = s2.s > 0 && s2.s < p;
      if (sw1 && sw2) {
        std::vector<uint8_t> z1(p), z2(p);
        for (int j = 0; j < p; ++j) { z1[j] = !s1.g[j]; z2[j] = !s2.g[j]; }
        int o1, o2, e1, e2;
        coupled_index(s1.g, s2.g, s1.s, s2.s, D, &o1, &o2);          // ones first (R/coupled_pairs01.R:5-21)
        coupled_index(z1, z2, p - s1.s, p - s2.s, D, &e1, &e2);      // then zeros (:23-38)
        std::vector<uint8_t> p1 = s1.g, p2 = s2.g;
        p1[e1] = 1; p1[o1] = 0;                                      // R/variableselection.R:92-97
        p2[e2] = 1; p2[o2] = 0;
        const double logu = ubm_log(D.u());                          // :100 always drawn
        try_move(s1, p1, s1.s, logu, D, false);
        try_move(s2, p2, s2.s, logu, D, false);
      } else {
        if (sw1) {                                                   // :118-131
          int i0, i1;
          sample_pair01(s1, D, &i0, &i1);
          std::vector<uint8_t> prop = s1.g;
          prop[i0] = 1; prop[i1] = 0;
          try_move(s1, prop, s1.s, 0.0, D, true);
        }
        if (sw2) {                                                   // :133-145
          int i0, i1;
          sample_pair01(s2, D, &i0, &i1);
          std::vector<uint8_t> prop = s2.g;
          prop[i0] = 1; prop[i1] = 0;
          try_move(s2, prop, s2.s, 0.0, D, true);
        }
      }
    }
    return s1.g == s2.g;
  }
};

}  // namespace orc
