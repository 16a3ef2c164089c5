// oracle_capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header for the parity status).
// extern "C" surface of the CPU oracle, loaded with ctypes by tests/, smoke() and bench.py's CPU legs.
// Mirrors the product's C ABI (include/ubm.h) argument for argument so that a test can hand the same
// structs to both sides.  `nthreads` > 1 runs replicates on a std::thread pool (the cpu_baseline leg).
#include <atomic>
#include <cstring>
#include <functional>
#include <thread>
#include "oracle_models.hpp"
#include "oracle_more.hpp"
#include "oracle_scalar.hpp"

using namespace orc;

namespace {

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n <= 1) { for (int64_t i = 0; i < n; ++i) f(i); return; }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&]() { for (;;) { int64_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
  for (auto& th : pool) th.join();
}

// canonical reduction over replicates: 1024 interleaved partials (ascending r), then xor-butterfly.
double sum1024(int64_t n, const std::function<double(int64_t)>& term, bool square) {
  std::vector<double> s(1024, 0.0), t(1024);
  for (int64_t r = 0; r < n; ++r) {
    double v = term(r);
    int q = (int)(r & 1023);
    s[q] = square ? ubm_fma(v, v, s[q]) : (s[q] + v);
  }
  for (int off = 512; off >= 1; off >>= 1) {
    for (int q = 0; q < 1024; ++q) t[q] = s[q] + s[q ^ off];
    s.swap(t);
  }
  return s[0];
}

template <class M>
int run_meeting(const M& model, const ubm_draws* draws, int64_t lag, int64_t max_it, int64_t nrep,
                int64_t rep_offset, double* meetingtime, int nthreads) {
  std::atomic<int> bad(0);
  parallel_for(nrep, nthreads, [&](int64_t r) {
    Draws D = make_draws(draws, r, rep_offset);
    RepResult res = sample_meetingtime(model, D, lag, max_it);
    meetingtime[r] = res.meetingtime;
    if (res.tape_exhausted) bad = 1;
  });
  return bad ? UBM_ERR_TAPE : UBM_OK;
}

template <class M>
int run_estimator(const M& model, const ubm_draws* draws, int h_kind, const ubm_bins* bins, int64_t k, int64_t m,
                  int64_t lag, int64_t max_it, int64_t nrep, int64_t rep_offset, const ubm_estimator_out* out,
                  int nthreads) {
  const int dimh = h_dim(h_kind, model.dim());
  const int nbins = (bins && bins->nbreaks >= 2) ? bins->nbreaks - 1 : 0;
  std::vector<double> ue((size_t)nrep * dimh), mc((size_t)nrep * dimh), co((size_t)nrep * dimh),
      bv((size_t)nrep * std::max(nbins, 1));
  std::vector<int64_t> bn((size_t)nrep * std::max(nbins, 1));
  std::vector<RepResult> rr(nrep);
  std::atomic<int> bad(0);
  parallel_for(nrep, nthreads, [&](int64_t r) {
    Draws D = make_draws(draws, r, rep_offset);
    rr[r] = sample_unbiasedestimator(model, D, h_kind, bins, k, m, lag, max_it, &mc[(size_t)r * dimh],
                                     &co[(size_t)r * dimh], &ue[(size_t)r * dimh],
                                     nbins ? &bv[(size_t)r * nbins] : nullptr, nbins ? &bn[(size_t)r * nbins] : nullptr);
    if (rr[r].tape_exhausted) bad = 1;
  });
  if (out->uestimator) memcpy(out->uestimator, ue.data(), ue.size() * 8);
  if (out->mcmcestimator) memcpy(out->mcmcestimator, mc.data(), mc.size() * 8);
  if (out->correction) memcpy(out->correction, co.data(), co.size() * 8);
  if (out->bins && nbins) memcpy(out->bins, bv.data(), (size_t)nrep * nbins * 8);
  for (int64_t r = 0; r < nrep; ++r) {
    if (out->meetingtime) out->meetingtime[r] = rr[r].meetingtime;
    if (out->iteration) out->iteration[r] = rr[r].iteration;
    if (out->cost) out->cost[r] = rr[r].cost;
  }
  if (out->summary) {
    double* S = out->summary;
    auto ok = [&](int64_t r) { return std::isfinite(rr[r].meetingtime); };
    double nd = 0, nc = 0;
    for (int64_t r = 0; r < nrep; ++r) { if (ok(r)) nd += 1; else nc += 1; }
    S[0] = nd; S[1] = nc;
    S[2] = sum1024(nrep, [&](int64_t r) { return ok(r) ? rr[r].meetingtime : 0.0; }, false);
    S[3] = sum1024(nrep, [&](int64_t r) { return ok(r) ? rr[r].meetingtime : 0.0; }, true);
    S[4] = sum1024(nrep, [&](int64_t r) { return ok(r) ? rr[r].cost : 0.0; }, false);
    S[5] = sum1024(nrep, [&](int64_t r) { return ok(r) ? (double)rr[r].iteration : 0.0; }, false);
    for (int i = 0; i < dimh; ++i) {
      S[6 + i] = sum1024(nrep, [&](int64_t r) { return ok(r) ? ue[(size_t)r * dimh + i] : 0.0; }, false);
      S[6 + dimh + i] = sum1024(nrep, [&](int64_t r) { return ok(r) ? ue[(size_t)r * dimh + i] : 0.0; }, true);
    }
    // bins: exact integer sums of the numerators (and of their squares), one rounding each
    const double denom = (double)(m - k + 1);
    for (int b = 0; b < nbins; ++b) {
      long long c1 = 0, c2 = 0;
      for (int64_t r = 0; r < nrep; ++r) if (ok(r)) { long long c = bn[(size_t)r * nbins + b]; c1 += c; c2 += c * c; }
      S[6 + 2 * dimh + b] = (double)c1 / denom;
      S[6 + 2 * dimh + nbins + b] = (double)c2 / (denom * denom);
    }
  }
  return bad ? UBM_ERR_TAPE : UBM_OK;
}

template <class M>
int run_chains(const M& model, const ubm_draws* draws, int64_t m, int64_t lag, int64_t max_it, int64_t stride,
               int64_t nrep, int64_t rep_offset, double* samples1, double* samples2, double* meetingtime,
               int64_t* iteration, double* cost, int nthreads) {
  const int d = model.dim();
  std::atomic<int> bad(0);
  int64_t cap = stride - 1;
  int64_t eff_max = (max_it <= 0 || max_it > cap) ? cap : max_it;
  parallel_for(nrep, nthreads, [&](int64_t r) {
    Draws D = make_draws(draws, r, rep_offset);
    std::vector<double> s1, s2;
    RepResult res = sample_coupled_chains(model, D, m, lag, eff_max, s1, s2);
    memcpy(samples1 + (size_t)r * stride * d, s1.data(), s1.size() * 8);
    memcpy(samples2 + (size_t)r * stride * d, s2.data(), s2.size() * 8);
    if (meetingtime) meetingtime[r] = res.meetingtime;
    if (iteration) iteration[r] = res.iteration;
    if (cost) cost[r] = res.cost;
    if (res.tape_exhausted) bad = 1;
  });
  return bad ? UBM_ERR_TAPE : UBM_OK;
}

}  // namespace

#define DISPATCH(CALL)                                                                              \
  switch (model_kind) {                                                                             \
    case UBM_MODEL_NORMAL_MIXTURE: { MixtureModel M((const ubm_model_normal_mixture*)model); return CALL; } \
    case UBM_MODEL_MVN: { MvnModel M((const ubm_model_mvn*)model); return CALL; }                   \
    ORC_MORE_DISPATCH(CALL)                                                                         \
    default: return UBM_ERR_UNSUPPORTED;                                                            \
  }

extern "C" {

// R-emulated session stream (r_stream.hpp): create with set.seed(seed); draw typed values (runif / rnorm / rexp order)
void* orc_rsession_create(uint32_t seed) { return new RStream(seed); }
void orc_rsession_destroy(void* s) { delete reinterpret_cast<RStream*>(s); }
void orc_rsession_draw(void* s, int32_t type, int64_t n, double* out) {
  RStream* R = reinterpret_cast<RStream*>(s);
  for (int64_t i = 0; i < n; ++i) out[i] = (type == 0) ? R->unif_rand() : (type == 1) ? R->norm_rand() : R->exp_rand();
}

int orc_sample_meetingtime(int32_t model_kind, const void* model, const ubm_draws* draws, int64_t lag,
                           int64_t max_iterations, int64_t nrep, int64_t rep_offset, double* meetingtime,
                           int32_t nthreads) {
  if (lag < 1) return UBM_ERR_INVALID;
  if (draws->mode == ORC_DRAWS_RSTREAM) nthreads = 1;   // one sequential stream
  DISPATCH(run_meeting(M, draws, lag, max_iterations, nrep, rep_offset, meetingtime, nthreads))
}

int orc_sample_unbiasedestimator(int32_t model_kind, const void* model, const ubm_draws* draws, int32_t h_kind,
                                 const ubm_bins* bins, int64_t k, int64_t m, int64_t lag, int64_t max_iterations,
                                 int64_t nrep, int64_t rep_offset, const ubm_estimator_out* out, int32_t nthreads) {
  if (lag < 1 || k > m || lag > m || k < 0) return UBM_ERR_INVALID;   // R/unbiasedestimator.R:44-51
  if (draws->mode == ORC_DRAWS_RSTREAM) nthreads = 1;
  DISPATCH(run_estimator(M, draws, h_kind, bins, k, m, lag, max_iterations, nrep, rep_offset, out, nthreads))
}

int orc_sample_coupled_chains(int32_t model_kind, const void* model, const ubm_draws* draws, int64_t m, int64_t lag,
                              int64_t max_iterations, int64_t stride, int64_t nrep, int64_t rep_offset,
                              double* samples1, double* samples2, double* meetingtime, int64_t* iteration,
                              double* cost, int32_t nthreads) {
  if (lag < 1) return UBM_ERR_INVALID;
  if (draws->mode == ORC_DRAWS_RSTREAM) nthreads = 1;
  DISPATCH(run_chains(M, draws, m, lag, max_iterations, stride, nrep, rep_offset, samples1, samples2, meetingtime,
                      iteration, cost, nthreads))
}

int orc_h_bar(int32_t dim, int32_t h_kind, int64_t k, int64_t m, int64_t lag, int64_t stride, int64_t nrep,
              const double* samples1, const double* samples2, const double* meetingtime, const int64_t* iteration,
              double* hbar) {
  const int dimh = h_dim(h_kind, dim);
  for (int64_t r = 0; r < nrep; ++r) {
    if (k > iteration[r] || m > iteration[r]) return UBM_ERR_INVALID;   // R/H_bar.R:21-28
    H_bar(samples1 + (size_t)r * stride * dim, samples2 + (size_t)r * stride * dim, dim, h_kind, k, m, lag,
          meetingtime[r], hbar + (size_t)r * dimh);
  }
  return UBM_OK;
}

// c_chains_to_measure_as_list_ — src/c_chains_to_measure.cpp:38-84 (R/c_chains_to_measure_as_list.R:19-27), per replicate
int orc_c_chains_to_measure(int32_t dim, int64_t k, int64_t m, int64_t lag, int64_t stride, int64_t nrep,
                            const double* samples1, const double* samples2, const double* meetingtime,
                            const int64_t* iteration, int64_t cap, double* atoms, double* weights, int32_t* mcmc_part,
                            int64_t* sizes) {
  if (k > m) return UBM_ERR_INVALID;                                                   // R :20-22
  for (int64_t r = 0; r < nrep; ++r) {
    if (m > iteration[r]) return UBM_ERR_INVALID;                                      // R :23-25
    if (!(meetingtime[r] < 9.0e18)) return UBM_ERR_INVALID;
    const int64_t meeting = (int64_t)meetingtime[r];
    int64_t size;                                                                      // :45-49
    if (((meeting - 1) - (k + lag) + 1) > 0) size = (m - k + 1) + 2 * ((meeting - 1) - (k + lag) + 1);
    else size = (m - k + 1);
    if (size > cap) return UBM_ERR_INVALID;
    sizes[r] = size;
    const double eqweight = 1. / (double)(m - k + 1);                                  // :51
    const double* s1 = samples1 + (size_t)r * stride * dim;
    const double* s2 = samples2 + (size_t)r * stride * dim;
    double* A = atoms + (size_t)r * cap * dim;
    double* W = weights + (size_t)r * cap;
    int32_t* P = mcmc_part + (size_t)r * cap;
    for (int64_t i = 0; i < cap; ++i) { W[i] = 0.0; P[i] = 0; for (int c = 0; c < dim; ++c) A[i * dim + c] = 0.0; }
    for (int64_t i = 0; i < (m - k + 1); ++i) {                                        // :62-66
      for (int c = 0; c < dim; ++c) A[i * dim + c] = s1[(k + i) * dim + c];
      W[i] = eqweight; P[i] = 1;
    }
    int64_t index = m - k + 1;
    if (((meeting - 1) - (k + lag) + 1) > 0) {                                         // :70-82
      for (int64_t time = k + lag; time <= meeting - 1; ++time) {
        for (int c = 0; c < dim; ++c) { A[index * dim + c] = s1[time * dim + c]; A[(index + 1) * dim + c] = s2[(time - lag) * dim + c]; }
        double w = floor(((double)time - k) / (double)lag) - ceil(std::max<double>((double)lag, (double)time - m) / (double)lag) + 1.;
        w = w * eqweight;
        W[index] = w; W[index + 1] = -w;
        index += 2;
      }
    }
  }
  return UBM_OK;
}

// tv_upper_bound — vignettes/polyagammagibbs.Rmd:238-244: mean over meeting times of pmax(0, ceiling((tau - lag - t) / lag))
int orc_tv_upper_bounds(const double* meetingtime, int64_t nrep, int64_t lag, const int64_t* t, int32_t nt, double* bound) {
  for (int32_t j = 0; j < nt; ++j) {
    double s = 0.0;
    for (int64_t r = 0; r < nrep; ++r) {
      if (!(meetingtime[r] < 9.0e18)) return UBM_ERR_INVALID;
      s += std::max(0.0, ceil((meetingtime[r] - (double)lag - (double)t[j]) / (double)lag));
    }
    bound[j] = s / (double)nrep;
  }
  return UBM_OK;
}

int orc_estimator_bins(int32_t dim, const ubm_bins* bins, int64_t k, int64_t m, int64_t lag, int64_t stride,
                       int64_t nrep, const double* samples1, const double* samples2, const double* meetingtime,
                       double* binvals) {
  const int nbins = bins->nbreaks - 1;
  for (int64_t r = 0; r < nrep; ++r)
    for (int b = 0; b < nbins; ++b)
      binvals[(size_t)r * nbins + b] =
          estimator_bin(samples1 + (size_t)r * stride * dim, samples2 + (size_t)r * stride * dim, dim,
                        bins->component, bins->breaks[b], bins->breaks[b + 1], k, m, lag, meetingtime[r]);
  return UBM_OK;
}

// stand-alone primitives (same argument meaning as the product's ubm_* entry points)
int orc_rnorm_coupling(int32_t mode, const ubm_draws* draws, int64_t n, int64_t rep_offset, double mu1, double mu2,
                       double s1, double s2, double* xy, int32_t* identical) {
  ubm_model_normal_mixture dummy{};
  double one = 1.0, zero = 0.0;
  dummy.ncomp = 1; dummy.weights = &one; dummy.means = &zero; dummy.sds = &one; dummy.proposal_sd = 1.0; dummy.init_sd = 1.0;
  MixtureModel M(&dummy);
  int bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    Draws D = make_draws(draws, i, rep_offset);
    bool id;
    if (mode == 0) id = M.reflmax(mu1, mu2, s1, D, &xy[2 * i], &xy[2 * i + 1]);
    else id = M.maxcoupling(mu1, mu2, s1, ubm_log(s1), s2, ubm_log(s2), D, &xy[2 * i], &xy[2 * i + 1]);
    identical[i] = id ? 1 : 0;
    if (D.exhausted) bad = 1;
  }
  return bad ? UBM_ERR_TAPE : UBM_OK;
}

// covariance forms (src/mvnorm.cpp:9-26, :49-67): LLT by PgLogisticModel::chol_lower, then the factor forms.  `which` 0: the
// upper factor U (column-major) of the covariance, 1: its inverse
int orc_cov_factor(int32_t which, int32_t d, const double* covariance, double* out) {
  std::vector<double> L(covariance, covariance + (size_t)d * d);
  bool ok = true;
  PgLogisticModel::chol_lower(L, d, &ok);
  if (!ok) return UBM_ERR_INVALID;
  std::vector<double> T = which ? PgLogisticModel::tri_inverse(L, d) : L;
  for (int j = 0; j < d; ++j) for (int i = 0; i < d; ++i) out[(size_t)j * d + i] = (i <= j) ? T[(size_t)i * d + j] : 0.0;
  return UBM_OK;
}

// Bayesian lasso probe: one blassoconditional step (src/blassoutil.cpp:17-41) from a tape of normals: beta [p], {norm, betaDbeta}
int orc_blasso_conditional(const ubm_model_blasso* mdl, const ubm_draws* draws, const double* tau2, double sigma2, double* beta,
                           double* norm_bdb) {
  BlassoModel M(mdl);
  Draws D = make_draws(draws, 0, 0);
  const BlassoModel::Cond c = M.conditional(tau2, sigma2);
  M.draw_beta(c, D, beta);
  norm_bdb[0] = M.rss(beta); norm_bdb[1] = M.bdb(beta, tau2);
  return D.exhausted ? UBM_ERR_TAPE : UBM_OK;
}

void orc_ising_swap_counts(int32_t reset, int64_t* out63) {
  for (int i = 0; i < 63; ++i) { if (out63) out63[i] = g_ising_pt[i].load(); if (reset) g_ising_pt[i] = 0; }
}

// kind 0 Gamma, 1 inverse Gamma, 2 inverse Gaussian; mode 0: n plain draws of the first distribution, 1: coupled pairs
int orc_positive_coupling(int32_t kind, int32_t mode, const ubm_draws* draws, int64_t n, int64_t rep_offset, double a1,
                          double a2, double b1, double b2, double* out, int32_t* identical) {
  const double c1 = (kind == 2) ? 0.0 : a1 * ubm_log(b1) - std::lgamma(a1), c2 = (kind == 2) ? 0.0 : a2 * ubm_log(b2) - std::lgamma(a2);
  int bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    Draws D = make_draws(draws, i, rep_offset);
    if (mode == 0) {
      out[i] = (kind == 0) ? rgamma_mt(a1, b1, D) : (kind == 1) ? 1.0 / rgamma_mt(a1, b1, D) : rinvgaussian1(a1, b1, D);
    } else {
      identical[i] = positive_max_coupling(kind, a1, a2, b1, b2, c1, c2, D, &out[2 * i], &out[2 * i + 1]) ? 1 : 0;
    }
    if (D.exhausted) bad = 1;
  }
  return bad ? UBM_ERR_TAPE : UBM_OK;
}

int orc_mvn_sample(int32_t mode, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d, const double* mu1,
                   const double* mu2, const double* chol, const double* chol_inv, double* out, int32_t* identical) {
  std::vector<double> eye((size_t)d * d, 0.0), zero(d, 0.0);
  for (int j = 0; j < d; ++j) eye[(size_t)j * d + j] = 1.0;
  ubm_model_mvn s{};
  s.d = d; s.mean_pi = zero.data(); s.chol_inv_pi = eye.data(); s.chol_prop = chol;
  s.chol_inv_prop = chol_inv ? chol_inv : eye.data(); s.init_mean = zero.data(); s.init_chol = eye.data();
  MvnModel M(&s);
  int bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    Draws D = make_draws(draws, i, rep_offset);
    if (mode == 0) M.rmvnorm_chol(mu1, M.c_prop, D, out + (size_t)i * d);
    else identical[i] = M.reflmax(mu1, mu2, D, out + (size_t)i * 2 * d, out + (size_t)i * 2 * d + d) ? 1 : 0;
    if (D.exhausted) bad = 1;
  }
  return bad ? UBM_ERR_TAPE : UBM_OK;
}

// rmvnorm_max_coupling_ (src/mvnorm.cpp:91-129, variant 1: A = covariances) and rmvnorm_max_coupling_cholesky
// (:133-174, variant 0: A = chol, B = chol_inverse, both upper).  Variant 1 takes the LLT once (PgLogisticModel::chol_lower,
// the same code path as sample_beta's use of this coupling) and then draws / evaluates with sequential sums; variant 0
// goes through MvnModel's fast_rmvnorm_cholesky_ / fast_dmvnorm_cholesky_inverse_ restatements.
int orc_mvn_max(int32_t variant, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d, const double* mu1,
                const double* mu2, const double* A1, const double* A2, const double* B1, const double* B2, double* xy,
                int32_t* identical, int32_t* nattempts) {
  int bad = 0;
  if (variant == 0) {
    std::vector<double> eye((size_t)d * d, 0.0);
    for (int j = 0; j < d; ++j) eye[(size_t)j * d + j] = 1.0;
    ubm_model_mvn s1{}, s2{};
    s1.d = d; s1.mean_pi = mu1; s1.chol_inv_pi = B1; s1.chol_prop = A1; s1.chol_inv_prop = eye.data(); s1.init_mean = mu1; s1.init_chol = eye.data();
    s2.d = d; s2.mean_pi = mu2; s2.chol_inv_pi = B2; s2.chol_prop = A2; s2.chol_inv_prop = eye.data(); s2.init_mean = mu2; s2.init_chol = eye.data();
    MvnModel M1(&s1), M2(&s2);
    for (int64_t i = 0; i < n; ++i) {
      Draws D = make_draws(draws, i, rep_offset);
      double* X = xy + (size_t)i * 2 * d;
      double* Y = X + d;
      std::vector<double> x(d);
      M1.rmvnorm_chol(mu1, M1.c_prop, D, x.data());                                     // :144
      for (int j = 0; j < d; ++j) X[j] = x[j];
      const double u1 = ubm_log(D.u());                                                 // :149
      int na = 0;
      const bool same = (M1.target(x.data()) + u1) <= M2.target(x.data());              // :152
      if (!same) {
        bool accept = false;
        while (!accept && !D.exhausted) {                                               // :159-168
          ++na;
          M2.rmvnorm_chol(mu2, M2.c_prop, D, x.data());
          const double u2 = ubm_log(D.u());
          if ((M2.target(x.data()) + u2) > M1.target(x.data())) accept = true;
        }
      }
      for (int j = 0; j < d; ++j) Y[j] = x[j];
      identical[i] = same ? 1 : 0;
      if (nattempts) nattempts[i] = na;
      if (D.exhausted) bad = 1;
    }
    return bad ? UBM_ERR_TAPE : UBM_OK;
  }
  std::vector<double> L1(A1, A1 + (size_t)d * d), L2(A2, A2 + (size_t)d * d);
  bool ok = true;
  PgLogisticModel::chol_lower(L1, d, &ok);
  PgLogisticModel::chol_lower(L2, d, &ok);
  auto draw = [&](const double* mu, const std::vector<double>& L, Draws& D, double* out) {   // fast_rmvnorm_ :9-26
    std::vector<double> z(d);
    for (int j = 0; j < d; ++j) z[j] = D.n();
    for (int j = 0; j < d; ++j) {
      double a = 0.0;
      for (int q = 0; q <= j; ++q) a = ubm_fma(z[q], L[(size_t)q * d + j], a);
      out[j] = a + mu[j];
    }
  };
  auto dens = [&](const double* x, const double* mu, const std::vector<double>& L) {         // fast_dmvnorm_ :49-67
    double halflogdet = 0.0;
    for (int j = 0; j < d; ++j) halflogdet = halflogdet + ubm_log(L[(size_t)j * d + j]);
    const double cst = -(halflogdet) - (d * 0.9189385);
    std::vector<double> y(d);
    for (int r = 0; r < d; ++r) {
      double a = x[r] - mu[r];
      for (int k = 0; k < r; ++k) a = ubm_fma(-L[(size_t)k * d + r], y[k], a);
      y[r] = a / L[(size_t)r * d + r];
    }
    double sq = 0.0;
    for (int r = 0; r < d; ++r) sq = ubm_fma(y[r], y[r], sq);
    return -0.5 * sq + cst;
  };
  for (int64_t i = 0; i < n; ++i) {
    Draws D = make_draws(draws, i, rep_offset);
    double* X = xy + (size_t)i * 2 * d;
    double* Y = X + d;
    std::vector<double> x(d);
    draw(mu1, L1, D, x.data());                                                         // :96
    for (int j = 0; j < d; ++j) X[j] = x[j];
    const double logu = ubm_log(D.u());                                                 // :103-106
    int na = 0;
    const bool same = (dens(x.data(), mu1, L1) + logu) < dens(x.data(), mu2, L2);       // :107
    if (!same) {
      bool accept = false;
      while (!accept && !D.exhausted) {                                                 // :113-123
        ++na;
        draw(mu2, L2, D, x.data());
        const double lu = ubm_log(D.u());
        if ((dens(x.data(), mu2, L2) + lu) > dens(x.data(), mu1, L1)) accept = true;
      }
    }
    for (int j = 0; j < d; ++j) Y[j] = x[j];
    identical[i] = same ? 1 : 0;
    if (nattempts) nattempts[i] = na;
    if (D.exhausted) bad = 1;
  }
  return bad ? UBM_ERR_TAPE : UBM_OK;
}

int orc_dmvnorm_chol_inverse(int64_t n, int32_t d, const double* x, const double* mean, const double* chol_inv, double* out) {
  std::vector<double> eye((size_t)d * d, 0.0);
  for (int j = 0; j < d; ++j) eye[(size_t)j * d + j] = 1.0;
  ubm_model_mvn s{};
  s.d = d; s.mean_pi = mean; s.chol_inv_pi = chol_inv; s.chol_prop = eye.data(); s.chol_inv_prop = eye.data();
  s.init_mean = mean; s.init_chol = eye.data();
  MvnModel M(&s);
  for (int64_t i = 0; i < n; ++i) out[i] = M.target(x + (size_t)i * d);
  return UBM_OK;
}

// Polya-Gamma probes: n draws of PG(1, z) (cell i of a fresh replicate stream), log Phi
void orc_pg_draws(uint64_t seed, double z, int64_t n, double* out) {
  Draws D; D.seed = seed; D.rep = 0;
  for (int64_t i = 0; i < n; ++i) { CellDraws c(&D, (uint64_t)i); out[i] = pg_draw(z, c); }
}
void orc_vec_pnorm_log(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ubm_pnorm_log(x[i]); }

// ---- probes of single reference functions, for the comparison with the reference's own compiled code (oracle/_ref) ----
// lattices: 32 x 32 column-major int32 (R's IntegerMatrix), as passed to src/ising.cpp
static void lattice_in(const IsingModel& M, IsingModel::State& s, const int32_t* a) {
  s.x.assign((size_t)32 * 32, 0); s.sums.assign(1, 0.0);
  for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) M.at(s, 0, i, j) = (int8_t)a[j * 32 + i];
}
static void lattice_out(const IsingModel& M, const IsingModel::State& s, int32_t* a) {
  for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) a[j * 32 + i] = M.at(s, 0, i, j);
}
int32_t orc_ising_sum(const int32_t* state) {
  double beta = 0.4; ubm_model_ising_pt m{32, 1, &beta, 0.0, 0, 0};
  IsingModel M(&m); IsingModel::State s; lattice_in(M, s, state);
  return M.ising_sum(s, 0);
}
// one (coupled, if state2 != NULL) sweep at inverse temperature beta with the uniforms of `u`; probas[5] out
void orc_ising_sweep(double beta, int32_t* state1, int32_t* state2, const double* u, int64_t lu, double* probas) {
  ubm_model_ising_pt m{32, 1, &beta, 0.0, 0, 0};
  IsingModel M(&m); IsingModel::State s1, s2;
  for (int q = 0; q < 5; ++q) probas[q] = M.probas[q];
  Draws D; D.mode = UBM_DRAWS_TAPE; D.tu = u; D.lu = lu;
  lattice_in(M, s1, state1);
  if (state2) { lattice_in(M, s2, state2); M.coupled_sweep(s1, s2, 0, D); lattice_out(M, s2, state2); }
  else M.sweep(s1, 0, D);
  lattice_out(M, s1, state1);
}
// sample_pair01 (0-based indices: zero index, one index)
void orc_sample_pair01(const ubm_model_varsel* mdl, const double* selection, double u0, double u1, int32_t* out2) {
  VarselModel M(mdl);
  VarselModel::State st; st.g.assign(M.p, 0);
  for (int j = 0; j < M.p; ++j) st.g[j] = selection[j] != 0.0;
  M.sync(st);
  double tu[2] = {u0, u1};
  Draws D; D.mode = UBM_DRAWS_TAPE; D.tu = tu; D.lu = 2;
  int i0, i1; M.sample_pair01(st, D, &i0, &i1);
  out2[0] = i0; out2[1] = i1;
}
double orc_varsel_ml(const ubm_model_varsel* mdl, const double* selection) {
  VarselModel M(mdl);
  std::vector<uint8_t> g(M.p);
  for (int j = 0; j < M.p; ++j) g[j] = selection[j] != 0.0;
  return M.ml(g);
}
// Polya-Gamma pieces under SEQUENTIAL (replay) draws of replicate 0 of `draws`; used[3] = draws consumed (u, n, e)
static void used_out(const Draws& D, int64_t* used) { if (used) { used[0] = (int64_t)D.iu; used[1] = (int64_t)D.in; used[2] = (int64_t)D.ie; } }
int orc_pg_rpg_seq(const ubm_draws* draws, const double* z, int64_t count, double* out, int64_t* used) {
  Draws D = make_draws(draws, 0, 0); SeqDraws c(&D);
  for (int64_t i = 0; i < count; ++i) out[i] = pg_draw(z[i], c);
  used_out(D, used);
  return D.exhausted ? UBM_ERR_TAPE : UBM_OK;
}
int orc_pg_w_seq(const ubm_model_pg_logistic* mdl, const ubm_draws* draws, const double* beta1, const double* beta2,
                 double* w, int64_t* used) {
  PgLogisticModel M(mdl);
  std::vector<double> z1, z2;
  M.xbeta(beta1, z1); M.xbeta(beta2, z2);
  Draws D = make_draws(draws, 0, 0); SeqDraws c(&D);
  for (int i = 0; i < M.n; ++i) M.w_pair(fabs(z1[i]), fabs(z2[i]), c, &w[2 * i], &w[2 * i + 1]);
  used_out(D, used);
  return D.exhausted ? UBM_ERR_TAPE : UBM_OK;
}
double orc_pg_logcosh(double x) { return pg_logcosh(x); }
void orc_pg_xbeta(const ubm_model_pg_logistic* mdl, const double* beta, double* out) {
  PgLogisticModel M(mdl); std::vector<double> z; M.xbeta(beta, z);
  for (int i = 0; i < M.n; ++i) out[i] = z[i];
}
// m_sigma_function_: m [p], L (= Cholesky_inverse) and Linv (= Cholesky), p x p column-major; also invB and ktk
void orc_pg_m_sigma(const ubm_model_pg_logistic* mdl, const double* omega, double* m, double* L, double* Linv,
                    double* invB, double* ktk) {
  PgLogisticModel M(mdl);
  PgLogisticModel::MS r = M.m_sigma(std::vector<double>(omega, omega + M.n));
  const int p = M.p;
  for (int i = 0; i < p; ++i) { m[i] = r.m[i]; ktk[i] = M.ktk[i]; }
  for (int q = 0; q < p * p; ++q) { L[q] = r.L[q]; Linv[q] = r.Linv[q]; invB[q] = M.invB[q]; }
}

// numerics probes (tests/test_numerics.py)
void orc_vec_log(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ubm_log(x[i]); }
void orc_vec_exp(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ubm_exp(x[i]); }
void orc_vec_log1p(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ubm_log1p(x[i]); }
void orc_vec_qnorm(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ubm_qnorm(x[i]); }
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  ubm_u32x4 b = ubm_philox4x32_10(c0, c1, c2, c3, k0, k1);
  for (int i = 0; i < 4; ++i) out[i] = b.w[i];
}
// the first n draws of each typed stream of one replicate (to build tapes that replay a Philox run)
void orc_stream(uint64_t seed, uint64_t rep, int32_t type, int64_t n, double* out) {
  Draws D; D.seed = seed; D.rep = rep;
  for (int64_t i = 0; i < n; ++i) out[i] = (type == 0) ? D.u() : (type == 1) ? D.n() : D.e();
}

}  // extern "C"
