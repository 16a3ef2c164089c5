// oracle_pg.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header for the parity status).
// CPU restatement of the reference's Polya-Gamma Gibbs sampler for logistic regression (C3):
//   src/PolyaGamma.cpp:65-79 (a), :89-104 (mass_texpon), :106-139 (rtigauss), :175-226 (draw_like_devroye),
//   src/PolyaGamma.h:34-38 (constants), src/RRNG.cpp:50-95 (RNG facade: unif / expon_rate / norm / p_norm)
//   src/logisticregressioncoupling.cpp:9-17 (logcosh), :22-32 (xbeta_), :79-115 (w_max_couplingC)
//   src/logisticregression.cpp:32-56 (m_sigma_function_)
//   src/mvnorm.cpp:9-26 (fast_rmvnorm_), :49-67 (fast_dmvnorm_), :91-129 (rmvnorm_max_coupling_)
//   R/logistic_regression.R:180-209 (sample_beta), inst/reproducegermancredit/germancredit.run.R:23-49 (kernels, rinit)
//
// Draw addressing.  The reference consumes one sequential stream, observation after observation; a device cannot
// (the n Polya-Gamma draws of a step run in parallel).  The engine therefore addresses the PG draws of a step by
// CELL: cell c = (sweep * n + observation) owns its own blocks of the replicate's typed streams (struct CellDraws
// below gives the exact address); `sweep` = index of the kernel call (= time).
// The multivariate-normal draws of the beta update (and rinit) use the ordinary sequential region from index 0.  Same distribution as the reference's order; this is the contract the CUDA kernel matches.
// REPLAY (UBM_DRAWS_TAPE, and the oracle-only R-session stream): the reference's own order.  Every draw comes from the
// replicate's typed tapes in the sequence the reference consumes them — observation after observation through
// w_max_couplingC / w_rejsamplerC (src/logisticregressioncoupling.cpp:42-113), rejection loops included, then the
// normals and uniforms of the beta update (struct SeqDraws).  The single kernel draws its n PG variables with the same
// vendored Devroye sampler, in observation order (the reference calls BayesLogit::rpg there, a package outside the tree).
//
// Arithmetic pinned where the reference leaves it to the compiler / Eigen / BLAS: a + b * c is one fma;
// Cholesky = left-looking column algorithm with sequential fma sums; triangular inverse by forward substitution;
// Sigma = Linv^T Linv by sequential sums over the non-zero range.
#pragma once
#include "oracle_core.hpp"
#include "../include/ubm_pg.h"

namespace orc {

static const double PG_PI = 3.141592653589793238462643383279502884197;   // src/PolyaGamma.h:34
static const double PG_TRUNC = 0.64;                                      // :37
static const double PG_TRUNC_RECIP = 1.0 / 0.64;                          // :38


// Cell draws.  A cell = (sweep, observation); it is divided into ATTEMPTS of 128 uniforms / 64 normals / 64
// exponentials (32 Philox blocks of each typed stream): attempt 0 holds the first PG draw and its coupling test,
// attempt a >= 1 the a-th candidate of the max-coupling rejection loop.  Attempts are independent of one another,
// which lets the device evaluate 32 of them at once and keep the first accepted one — the same answer as this
// sequential loop.  Block address inside the replicate's typed stream:
//     blk = 2^59 | (cell mod 2^26) << 33 | (attempt mod 2^28) << 5 | local,   local in [0, 32)
// (the ordinary sequential region of the streams, used by the beta update, lives below 2^59), and attempts beyond
// 2^28 move to replicate index rep + (attempt >> 28) << 44, so the loop never revisits a draw.
// Draws inside an attempt are sequential and wrap inside the attempt's 32 blocks (never an error).
struct CellDraws {
  uint64_t seed, rep, cell, base = 0, rep_eff = 0;
  uint32_t ku = 0, kn = 0, ke = 0;
  uint64_t cbu = ~0ull, cbn = ~0ull, cbe = ~0ull, cru = ~0ull, crn = ~0ull, cre = ~0ull;
  ubm_u32x4 bu{}, bn{}, be{};
  bool overflow = false;   // never set
  CellDraws(Draws* d, uint64_t cell_) : seed(d->seed), rep(d->rep), cell(cell_) { set_attempt(0); }
  void set_attempt(uint64_t a) {
    base = (1ull << 59) | ((cell & ((1ull << 26) - 1)) << 33) | ((a & ((1ull << 28) - 1)) << 5);
    rep_eff = rep + ((a >> 28) << 44);
    ku = kn = ke = 0;
  }
  double u() {
    const uint32_t k = (ku++) & 127u;
    const uint64_t blk = base | (k >> 2);
    if (blk != cbu || rep_eff != cru) { cbu = blk; cru = rep_eff; bu = ubm_stream_block(seed, rep_eff, UBM_STREAM_U, blk); }
    return ubm_u01_32(bu.w[k & 3]);
  }
  double n() {
    const uint32_t k = (kn++) & 63u;
    const uint64_t blk = base | (k >> 1);
    if (blk != cbn || rep_eff != crn) { cbn = blk; crn = rep_eff; bn = ubm_stream_block(seed, rep_eff, UBM_STREAM_N, blk); }
    return ubm_qnorm(ubm_block_u53(bn, k & 1));
  }
  double e() {
    const uint32_t k = (ke++) & 63u;
    const uint64_t blk = base | (k >> 1);
    if (blk != cbe || rep_eff != cre) { cbe = blk; cre = rep_eff; be = ubm_stream_block(seed, rep_eff, UBM_STREAM_E, blk); }
    return -ubm_log(ubm_block_u53(be, k & 1));
  }
};

// Sequential draws (replay): the replicate's own typed streams, in the reference's order.
struct SeqDraws {
  Draws* d;
  bool overflow = false;        // tape exhausted: the samplers' loops stop
  explicit SeqDraws(Draws* d_) : d(d_) {}
  void set_attempt(uint64_t) {}
  double u() { double v = d->u(); overflow = d->exhausted; return v; }
  double n() { double v = d->n(); overflow = d->exhausted; return v; }
  double e() { double v = d->e(); overflow = d->exhausted; return v; }
};

// PolyaGamma::a — src/PolyaGamma.cpp:65-79
inline double pg_a(int n, double x) {
  double K = (n + 0.5) * PG_PI;
  double y = 0;
  if (x > PG_TRUNC) {
    y = K * ubm_exp(-0.5 * K * K * x);
  } else if (x > 0) {
    double expnt = -1.5 * (ubm_log(0.5 * PG_PI) + ubm_log(x)) + ubm_log(K) - 2.0 * (n + 0.5) * (n + 0.5) / x;
    y = ubm_exp(expnt);
  }
  return y;
}
// PolyaGamma::mass_texpon — src/PolyaGamma.cpp:89-104
inline double pg_mass_texpon(double Z) {
  double t = PG_TRUNC;
  double fz = 0.125 * PG_PI * PG_PI + 0.5 * Z * Z;
  double b = sqrt(1.0 / t) * (t * Z - 1);
  double a = sqrt(1.0 / t) * (t * Z + 1) * -1.0;
  double x0 = ubm_log(fz) + fz * t;
  double xb = x0 - Z + ubm_pnorm_log(b);
  double xa = x0 + Z + ubm_pnorm_log(a);
  double qdivp = 4 / PG_PI * (ubm_exp(xb) + ubm_exp(xa));
  return 1.0 / (1.0 + qdivp);
}
// PolyaGamma::rtigauss — src/PolyaGamma.cpp:106-139
template <class R>
inline double pg_rtigauss(double Z, R& r) {
  Z = fabs(Z);
  double t = PG_TRUNC;
  double X = t + 1.0;
  if (PG_TRUNC_RECIP > Z) {
    double alpha = 0.0;
    while (r.u() > alpha) {
      double E1 = r.e(); double E2 = r.e();
      while (E1 * E1 > 2 * E2 / t) { E1 = r.e(); E2 = r.e(); if (r.overflow) break; }
      X = 1 + E1 * t;
      X = t / (X * X);
      alpha = ubm_exp(-0.5 * Z * Z * X);
      if (r.overflow) break;
    }
  } else {
    double mu = 1.0 / Z;
    while (X > t) {
      double Y = r.n(); Y *= Y;
      double half_mu = 0.5 * mu;
      double mu_Y = mu * Y;
      X = mu + half_mu * mu_Y - half_mu * sqrt(4 * mu_Y + mu_Y * mu_Y);
      if (r.u() > mu / (mu + X)) X = mu * mu / X;
      if (r.overflow) break;
    }
  }
  return X;
}
// PolyaGamma::draw_like_devroye — src/PolyaGamma.cpp:175-226 (rpg_devroye(1, z), logisticregressioncoupling.cpp:34-39)
template <class R>
inline double pg_draw(double Zin, R& r) {
  double Z = fabs(Zin) * 0.5;
  double fz = 0.125 * PG_PI * PG_PI + 0.5 * Z * Z;
  double X = 0.0, S = 1.0, Y = 0.0;
  const double mass = pg_mass_texpon(Z);        // loop-invariant in the reference too (recomputed there)
  while (true) {
    if (r.u() < mass) X = PG_TRUNC + r.e() / fz;
    else X = pg_rtigauss(Z, r);
    S = pg_a(0, X);
    Y = r.u() * S;
    int n = 0;
    bool go = true;
    while (go) {
      ++n;
      if (n % 2 == 1) {
        S = S - pg_a(n, X);
        if (Y <= S) return 0.25 * X;
      } else {
        S = S + pg_a(n, X);
        if (Y > S) go = false;
      }
    }
    if (r.overflow) return 0.25 * X;
  }
}
// logcosh — src/logisticregressioncoupling.cpp:9-17 (truncated constant kept verbatim)
inline double pg_logcosh(double x) {
  if (x > 0.) return x + ubm_log(1.0 + ubm_exp(-2.0 * x)) - 0.6931472;
  return -x + ubm_log(1.0 + ubm_exp(2.0 * x)) - 0.6931472;
}

struct PgLogisticModel {
  struct State { std::vector<double> beta; uint64_t sweep = 0; };
  int n = 0, p = 0, w_coupling = 0;   // w_coupling: sample_w(mode=): 0 'max', 1 'rej_samp'
  int beta_coupling = 0;              // coupled beta-step: 0 maximal coupling (sample_beta), 1 common random numbers (vignette)
  std::vector<double> X;        // n x p column-major
  std::vector<double> invB;     // p x p column-major  (logisticregression_precomputation, R/logistic_regression.R:34-50)
  std::vector<double> ktk;      // KTkappaplusinvBtimesb = X^T (Y - 1/2) + invB b
  std::vector<double> b, cholB; // prior mean, upper Cholesky factor of B (rinit: fast_rmvnorm(1, b, B))

  static void chol_lower(std::vector<double>& A, int p, bool* ok) {   // in place: lower triangle becomes L, A = L L^T
    for (int j = 0; j < p; ++j) {
      double s = A[(size_t)j * p + j];
      for (int k = 0; k < j; ++k) s = ubm_fma(-A[(size_t)k * p + j], A[(size_t)k * p + j], s);
      if (!(s > 0)) *ok = false;
      double ljj = sqrt(s);
      A[(size_t)j * p + j] = ljj;
      for (int i = j + 1; i < p; ++i) {
        double t = A[(size_t)j * p + i];
        for (int k = 0; k < j; ++k) t = ubm_fma(-A[(size_t)k * p + i], A[(size_t)k * p + j], t);
        A[(size_t)j * p + i] = t / ljj;
      }
    }
    for (int j = 0; j < p; ++j) for (int i = 0; i < j; ++i) A[(size_t)j * p + i] = 0.0;
  }

  explicit PgLogisticModel(const ubm_model_pg_logistic* s) {
    n = s->n; p = s->p; w_coupling = s->w_coupling; beta_coupling = s->beta_coupling;
    X.assign(s->X, s->X + (size_t)n * p);
    b.assign(s->b, s->b + p);
    // invB = solve(B); ktk = t(X) %*% (Y - 0.5) + invB %*% b  (host-side precomputation, as in R)
    std::vector<double> Bm(s->B, s->B + (size_t)p * p);
    bool ok = true;
    std::vector<double> LB = Bm; chol_lower(LB, p, &ok);
    cholB.assign((size_t)p * p, 0.0);                         // upper factor U = L^T, column-major
    for (int j = 0; j < p; ++j) for (int i = 0; i <= j; ++i) cholB[(size_t)j * p + i] = LB[(size_t)i * p + j];
    std::vector<double> Li = tri_inverse(LB, p);
    invB = ltl(Li, p);
    ktk.assign(p, 0.0);
    for (int j = 0; j < p; ++j) {
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc = ubm_fma(X[(size_t)j * n + i], s->Y[i] - 0.5, acc);
      for (int k = 0; k < p; ++k) acc = ubm_fma(invB[(size_t)k * p + j], b[k], acc);
      ktk[j] = acc;
    }
  }
  int dim() const { return p; }
  const double* chain_state(const State& s) const { return s.beta.data(); }
  void copy(State& dst, const State& src) const { dst = src; }

  // inverse of a lower-triangular matrix (column-major), forward substitution column by column.  Canonical arithmetic
  // of all the triangular solves of this model: the diagonal enters through its reciprocal (1 / L_ii once, then one
  // multiplication per unknown) — on the device a solve is a chain of dependent steps, and a division per step would
  // triple its length.
  static std::vector<double> tri_inverse(const std::vector<double>& L, int p) {
    std::vector<double> Li((size_t)p * p, 0.0);
    for (int j = 0; j < p; ++j) {
      Li[(size_t)j * p + j] = 1.0 / L[(size_t)j * p + j];
      for (int i = j + 1; i < p; ++i) {
        double s = 0.0;
        for (int k = j; k < i; ++k) s = ubm_fma(L[(size_t)k * p + i], Li[(size_t)j * p + k], s);
        Li[(size_t)j * p + i] = -s * (1.0 / L[(size_t)i * p + i]);
      }
    }
    return Li;
  }
  // t(Li) %*% Li for lower-triangular Li: S[a,b] = sum_{i >= max(a,b)} Li[i,a] Li[i,b]
  static std::vector<double> ltl(const std::vector<double>& Li, int p) {
    std::vector<double> S((size_t)p * p, 0.0);
    for (int a = 0; a < p; ++a)
      for (int c = a; c < p; ++c) {
        double s = 0.0;
        for (int i = c; i < p; ++i) s = ubm_fma(Li[(size_t)a * p + i], Li[(size_t)c * p + i], s);
        S[(size_t)c * p + a] = s; S[(size_t)a * p + c] = s;
      }
    return S;
  }
  // xbeta_ — src/logisticregressioncoupling.cpp:22-32
  void xbeta(const double* beta, std::vector<double>& out) const {
    out.assign(n, 0.0);
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int j = 0; j < p; ++j) acc = ubm_fma(X[(size_t)j * n + i], beta[j], acc);
      out[i] = acc;
    }
  }
  struct MS { std::vector<double> m, L, Linv; };
  // m_sigma_function_ — src/logisticregression.cpp:32-56
  MS m_sigma(const std::vector<double>& omega) const {
    MS r;
    r.L.assign((size_t)p * p, 0.0);
    for (int j1 = 0; j1 < p; ++j1)
      for (int j2 = j1; j2 < p; ++j2) {
        // :42-45.  Canonical order: four chains over the groups of four observations (class (i / 4) mod 4), each sequential
        // in i, every term one fma of the rounded X(i,j1) omega(i) with X(i,j2); added to invB in turn.  (What a chain
        // of FP64 tensor-core m8n8k4 operations computes; the reference sums (X X) omega sequentially.)
        double part[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < n; ++i) {
          const int c = (i >> 2) & 3;
          part[c] = ubm_fma(X[(size_t)j1 * n + i] * omega[i], X[(size_t)j2 * n + i], part[c]);
        }
        double a = invB[(size_t)j2 * p + j1];
        for (int c = 0; c < 4; ++c) a = a + part[c];
        r.L[(size_t)j1 * p + j2] = a;                                             // lower triangle (column j1, row j2)
        r.L[(size_t)j2 * p + j1] = a;
      }
    bool ok = true;
    chol_lower(r.L, p, &ok);                                                      // :49-50
    r.Linv = tri_inverse(r.L, p);                                                 // :55  lower.inverse()
    // x = A^{-1} b (:51) through the inverse factor that :55 needs anyway: y = Linv b, x = Linv' y — two sets of independent
    // dot products (terms in increasing index order) instead of two substitutions of p dependent steps each
    std::vector<double> y(p);
    for (int i = 0; i < p; ++i) {
      double s = 0.0;
      for (int k = 0; k <= i; ++k) s = ubm_fma(r.Linv[(size_t)k * p + i], ktk[k], s);
      y[i] = s;
    }
    r.m.assign(p, 0.0);
    for (int j = 0; j < p; ++j) {
      double s = 0.0;
      for (int i = j; i < p; ++i) s = ubm_fma(r.Linv[(size_t)j * p + i], y[i], s);
      r.m[j] = s;
    }
    return r;
  }
  // fast_rmvnorm_ with a precomputed upper factor U (src/mvnorm.cpp:9-26): one normal per column, y = z U + mean
  void rmvnorm_upper(const double* mean, const std::vector<double>& U, Draws& D, double* out) const {
    std::vector<double> z(p);
    for (int j = 0; j < p; ++j) z[j] = D.n();
    for (int j = 0; j < p; ++j) {
      double acc = 0.0;
      for (int i = 0; i <= j; ++i) acc = ubm_fma(z[i], U[(size_t)j * p + i], acc);
      out[j] = acc + mean[j];
    }
  }
  // fast_dmvnorm_ (src/mvnorm.cpp:49-67) given the LOWER Cholesky factor Ls of the covariance (column-major)
  double dmvnorm_lower(const double* x, const double* mean, const std::vector<double>& Ls) const {
    double halflogdet = 0.0;
    for (int j = 0; j < p; ++j) halflogdet = halflogdet + ubm_log(Ls[(size_t)j * p + j]);     // :55
    double cst = -(halflogdet) - (p * 0.9189385);                                             // :56
    std::vector<double> y(p);
    for (int i = 0; i < p; ++i) {                                                             // :62 lower.solve(xc^T)
      double sacc = x[i] - mean[i];
      for (int k = 0; k < i; ++k) sacc = ubm_fma(-Ls[(size_t)k * p + i], y[k], sacc);
      y[i] = sacc * (1.0 / Ls[(size_t)i * p + i]);
    }
    double sq = 0.0;
    for (int i = 0; i < p; ++i) sq = ubm_fma(y[i], y[i], sq);
    return -0.5 * sq + cst;
  }
  // rinit — germancredit.run.R:47-49: c(fast_rmvnorm(1, b, B)[1,], rep(0, n))
  void rinit(State& s, Draws& D) const {
    s.beta.assign(p, 0.0); s.sweep = 0;
    rmvnorm_upper(b.data(), cholB, D, s.beta.data());
  }
  // single_kernel — germancredit.run.R:23-30 (BayesLogit::rpg restated with the vendored Devroye sampler)
  void single(State& s, Draws& D) const {
    s.sweep += 1;
    std::vector<double> zs, w(n);
    xbeta(s.beta.data(), zs);
    if (D.mode != UBM_DRAWS_PHILOX) {                  // replay: one sequential stream
      SeqDraws c(&D);
      for (int i = 0; i < n && !c.overflow; ++i) w[i] = pg_draw(fabs(zs[i]), c);
    } else {
      for (int i = 0; i < n; ++i) {
        CellDraws c(&D, s.sweep * (uint64_t)n + i);
        w[i] = pg_draw(fabs(zs[i]), c);
        if (c.overflow) D.exhausted = true;
      }
    }
    MS r = m_sigma(w);
    // fast_rmvnorm_chol(1, m, Cholesky): y = z Cholesky + m with Cholesky = L^{-1} (lower): y_j = sum_{i >= j} z_i Linv[i,j]
    std::vector<double> z(p);
    for (int j = 0; j < p; ++j) z[j] = D.n();
    for (int j = 0; j < p; ++j) {
      double acc = 0.0;
      for (int i = j; i < p; ++i) acc = ubm_fma(z[i], r.Linv[(size_t)j * p + i], acc);
      s.beta[j] = acc + r.m[j];
    }
  }
  // one observation of sample_w: w_max_couplingC (src/logisticregressioncoupling.cpp:87-113) or, for mode 'rej_samp',
  // w_rejsamplerC (:50-74).  `c` is the draw source: the observation's cell (set_attempt moves to the cell's next
  // attempt) or the sequential replay stream (set_attempt is a no-op).
  template <class R>
  void w_pair(double z1, double z2, R& c, double* w1, double* w2) const {
    if (w_coupling == 1) {                                         // w_rejsamplerC
      const double z_min = z1 < z2 ? z1 : z2, z_max = z1 < z2 ? z2 : z1;   // std::min / std::max (:54-55)
      const double w_min = pg_draw(z_min, c);                      // :57
      const double log_u = ubm_log(c.u());                         // :59
      const double log_ratio = -0.5 * w_min * (z_max * z_max - z_min * z_min);   // :61 (pow(z, 2.) is exact)
      double w_max;
      if (log_u < log_ratio) w_max = w_min;                        // :62-63
      else { c.set_attempt(1); w_max = pg_draw(z_max, c); }        // :65, second draw = attempt 1 of the cell
      if (z1 < z2) { *w1 = w_min; *w2 = w_max; } else { *w1 = w_max; *w2 = w_min; }   // :67-73
      return;
    }
    const double a1 = pg_draw(z1, c);                              // :90
    double a2;
    const double u1 = c.u();                                       // :94
    const double logaccept1 = pg_logcosh(z2 / 2.) - 0.5 * z2 * z2 * a1 - (pg_logcosh(z1 / 2.) - 0.5 * z1 * z1 * a1);   // :96
    if (ubm_log(u1) <= logaccept1) a2 = a1;                        // :97-98
    else {
      bool accept = false;
      a2 = a1;
      uint64_t attempt = 0;
      while (!accept && !c.overflow) {                             // :100-110
        c.set_attempt(++attempt);
        a2 = pg_draw(z2, c);
        const double u2 = c.u();
        const double logaccept2 = pg_logcosh(z1 / 2.) - 0.5 * z1 * z1 * a2 - (pg_logcosh(z2 / 2.) - 0.5 * z2 * z2 * a2);
        if (ubm_log(u2) > logaccept2) accept = true;
      }
    }
    *w1 = a1; *w2 = a2;
  }
  // coupled_kernel — germancredit.run.R:32-45: sample_w (w_max_couplingC) + sample_beta (rmvnorm_max)
  bool coupled(State& s1, State& s2, Draws& D) const {
    s1.sweep += 1;
    std::vector<double> z1s, z2s, w1(n), w2(n);
    xbeta(s1.beta.data(), z1s);
    xbeta(s2.beta.data(), z2s);
    bool all_equal = true;
    if (D.mode != UBM_DRAWS_PHILOX) {                                // replay: the reference's sequential order
      SeqDraws c(&D);
      for (int i = 0; i < n && !c.overflow; ++i) {
        w_pair(fabs(z1s[i]), fabs(z2s[i]), c, &w1[i], &w2[i]);
        if (w1[i] != w2[i]) all_equal = false;
      }
    } else {
      for (int i = 0; i < n; ++i) {                                  // src/logisticregressioncoupling.cpp:87-113
        CellDraws c(&D, s1.sweep * (uint64_t)n + i);
        w_pair(fabs(z1s[i]), fabs(z2s[i]), c, &w1[i], &w2[i]);
        if (w1[i] != w2[i]) all_equal = false;
        if (c.overflow) D.exhausted = true;
      }
    }
    // sample_beta — R/logistic_regression.R:180-209: rmvnorm_max(m1, m2, t(C1) %*% C1, t(C2) %*% C2)
    MS r1 = m_sigma(w1), r2 = m_sigma(w2);
    if (beta_coupling == 1) {
      // pgg_coupled_kernel, vignettes/polyagammagibbs.Rmd:203-211: increment <- rnorm(p); beta_c = m_c + increment %*% Cholesky_c
      std::vector<double> z(p);
      for (int j = 0; j < p; ++j) z[j] = D.n();
      for (int j = 0; j < p; ++j) {
        double a1 = 0.0, a2 = 0.0;
        for (int i = j; i < p; ++i) { a1 = ubm_fma(z[i], r1.Linv[(size_t)j * p + i], a1); a2 = ubm_fma(z[i], r2.Linv[(size_t)j * p + i], a2); }
        s1.beta[j] = a1 + r1.m[j];
        s2.beta[j] = a2 + r2.m[j];
      }
      if (all_equal) s2.beta = s1.beta;
      return all_equal;
    }
    std::vector<double> S1 = ltl(r1.Linv, p), S2 = ltl(r2.Linv, p);
    bool ok = true;
    chol_lower(S1, p, &ok); chol_lower(S2, p, &ok);                  // LLT of Sigma1, Sigma2 (src/mvnorm.cpp:12,52)
    std::vector<double> U1((size_t)p * p, 0.0), U2((size_t)p * p, 0.0);
    for (int j = 0; j < p; ++j) for (int i = 0; i <= j; ++i) { U1[(size_t)j * p + i] = S1[(size_t)i * p + j]; U2[(size_t)j * p + i] = S2[(size_t)i * p + j]; }
    std::vector<double> x1(p), x2(p);
    rmvnorm_upper(r1.m.data(), U1, D, x1.data());                    // src/mvnorm.cpp:96
    const double logu = ubm_log(D.u());                              // :103-106
    bool ident_prop;
    if (dmvnorm_lower(x1.data(), r1.m.data(), S1) + logu < dmvnorm_lower(x1.data(), r2.m.data(), S2)) {   // :107
      x2 = x1; ident_prop = true;
    } else {
      ident_prop = false;
      bool accept = false;
      while (!accept && !D.exhausted) {                              // :113-123
        rmvnorm_upper(r2.m.data(), U2, D, x2.data());
        const double lu = ubm_log(D.u());
        if (dmvnorm_lower(x2.data(), r2.m.data(), S2) + lu > dmvnorm_lower(x2.data(), r1.m.data(), S1)) accept = true;
      }
    }
    (void)ident_prop;
    s1.beta = x1; s2.beta = x2;
    bool identical = false;
    if (all_equal) { s2.beta = s1.beta; identical = true; }          // germancredit.run.R:38-41
    return identical;
  }
};

}  // namespace orc
