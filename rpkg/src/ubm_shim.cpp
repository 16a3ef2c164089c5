// ubm_shim.cpp — the Rcpp side of the drop-in: R objects -> the structs of include/ubm.h -> ONE batched call per driver.
// Replaces the per-iteration .Call()s of the reference (R/RcppExports.R; registrations src/RcppExports.cpp:388-423).
// NOT BUILT in the development image (no R there); the identical C ABI is exercised through ctypes by the test-suite.
// Kernels / rinit closures returned by the R constructors carry attr(, "ubm_spec") = list(kind = <model>, ...): a plain
// R closure cannot run on a GPU, so models come from the device registry.
#include <Rcpp.h>
#include <cstdint>
#include <string>
#include <vector>
#include "ubm.h"
using namespace Rcpp;

static ubm_handle* handle() {                      // one handle per R process (R is single-threaded)
  static ubm_handle* h = nullptr;
  if (!h && ubm_create(0, &h) != UBM_OK) stop("ubm: no CUDA device (there is no CPU fallback)");
  return h;
}
static void check(int st) { if (st != UBM_OK) stop(std::string("ubm: ") + ubm_last_error(handle())); }

// seed: drawn once per call from R's stream, so that set.seed() still governs reproducibility
static uint64_t r_seed() {
  GetRNGstate(); double a = unif_rand(), b = unif_rand(); PutRNGstate();
  return ((uint64_t)(a * 4294967296.0) << 32) | (uint64_t)(b * 4294967296.0);
}
static ubm_draws philox() { ubm_draws d = {UBM_DRAWS_PHILOX, 0, r_seed(), nullptr, nullptr, nullptr, 0, 0, 0}; return d; }

// ---- the five registry structs.  `Model` keeps the R vectors alive while the struct points into them -----------------
struct Model {
  int32_t kind = 0;
  ubm_model_normal_mixture mix{}; ubm_model_mvn mvn{}; ubm_model_ising_pt ising{}; ubm_model_pg_logistic pg{}; ubm_model_varsel vs{};
  ubm_model_blasso bl{};
  std::vector<NumericVector> keep;
  const void* ptr() const {
    switch (kind) {
      case UBM_MODEL_NORMAL_MIXTURE: return &mix;
      case UBM_MODEL_MVN: return &mvn;
      case UBM_MODEL_ISING_PT: return &ising;
      case UBM_MODEL_PG_LOGISTIC: return &pg;
      case UBM_MODEL_BLASSO: return &bl;
      default: return &vs;
    }
  }
  double* vec(SEXP x) { keep.push_back(NumericVector(x)); return REAL(keep.back()); }
};

static void model_from(List spec, Model& M) {
  const std::string kind = as<std::string>(spec["kind"]);
  if (kind == "normal_mixture") {                       // README.Rmd:69-84, R/get_mh_kernels.R:20-79, bimodal.run.R:14-63
    M.kind = UBM_MODEL_NORMAL_MIXTURE;
    NumericVector w = spec["weights"];
    M.mix.ncomp = w.size(); M.mix.coupling = as<int>(spec["coupling"]);
    M.mix.weights = M.vec(spec["weights"]); M.mix.means = M.vec(spec["means"]); M.mix.sds = M.vec(spec["sds"]);
    M.mix.proposal_sd = as<double>(spec["proposal_sd"]);
    M.mix.init_mean = as<double>(spec["init_mean"]); M.mix.init_sd = as<double>(spec["init_sd"]);
  } else if (kind == "mvn") {                           // inst/reproducemvnorm/*.R: fast_dmvnorm_chol_inverse target, reflection-max RW-MH
    M.kind = UBM_MODEL_MVN;
    NumericVector mean = spec["mean_pi"];
    M.mvn.d = mean.size();
    M.mvn.mean_pi = M.vec(spec["mean_pi"]); M.mvn.chol_inv_pi = M.vec(spec["chol_inv_pi"]);
    M.mvn.chol_prop = M.vec(spec["chol_prop"]); M.mvn.chol_inv_prop = M.vec(spec["chol_inv_prop"]);
    M.mvn.init_mean = M.vec(spec["init_mean"]); M.mvn.init_chol = M.vec(spec["init_chol"]);
  } else if (kind == "ising_pt") {                      // R/ising.R:30-141
    M.kind = UBM_MODEL_ISING_PT;
    NumericVector betas = spec["betas"];
    M.ising.size = 32; M.ising.nchains = betas.size(); M.ising.betas = M.vec(spec["betas"]);
    M.ising.proba_swapmove = as<double>(spec["proba_swapmove"]); M.ising.scan = UBM_ISING_SCAN_SEQUENTIAL;
  } else if (kind == "pg_logistic") {                   // R/logistic_regression.R, germancredit.run.R:23-49
    M.kind = UBM_MODEL_PG_LOGISTIC;
    NumericMatrix X = spec["X"];
    M.pg.n = X.nrow(); M.pg.p = X.ncol();
    M.pg.X = M.vec(spec["X"]); M.pg.Y = M.vec(spec["Y"]); M.pg.b = M.vec(spec["b"]); M.pg.B = M.vec(spec["B"]);
    M.pg.w_coupling = as<std::string>(spec["mode"]) == "rej_samp" ? 1 : 0;      // sample_w(mode = ), R/logistic_regression.R:166-175
    M.pg.beta_coupling = spec.containsElementNamed("beta_coupling") && as<std::string>(spec["beta_coupling"]) == "crn" ? 1 : 0;
  } else if (kind == "varsel") {                        // R/variableselection.R:1-148
    M.kind = UBM_MODEL_VARSEL;
    NumericMatrix X = spec["X"];
    M.vs.n = X.nrow(); M.vs.p = X.ncol();
    M.vs.X = M.vec(spec["X"]); M.vs.Y = M.vec(spec["Y"]);
    M.vs.g = as<double>(spec["g"]); M.vs.kappa = as<double>(spec["kappa"]); M.vs.s0 = as<int>(spec["s0"]);
    M.vs.proportion_singleflip = as<double>(spec["proportion_singleflip"]);
  } else if (kind == "blasso") {                        // R/bayesianlasso.R:22-98
    M.kind = UBM_MODEL_BLASSO;
    NumericMatrix X = spec["X"];
    M.bl.n = X.nrow(); M.bl.p = X.ncol();
    M.bl.X = M.vec(spec["X"]); M.bl.Y = M.vec(spec["Y"]); M.bl.lambda = as<double>(spec["lambda"]);
  } else {
    stop("ubm: unknown model kind '" + kind + "'");
  }
}
static int64_t maxit(double max_iterations) { return R_finite(max_iterations) ? (int64_t)max_iterations : 0; }

// [[Rcpp::export]]
List ubm_sample_meetingtime_(List spec, double lag, double max_iterations, double nrep) {   // R/meetingtime.R:24-48
  Model M; model_from(spec, M);
  const int64_t n = (int64_t)nrep;
  ubm_draws dr = philox();
  NumericVector tau(n);
  check(ubm_sample_meetingtime(handle(), M.kind, M.ptr(), &dr, (int64_t)lag, maxit(max_iterations), n, 0, REAL(tau)));
  return List::create(_["meetingtime"] = tau);
}

// [[Rcpp::export]]
List ubm_sample_unbiasedestimator_(List spec, int h_kind, Nullable<NumericVector> breaks, int component, double k, double m,
                                   double lag, double max_iterations, double nrep) {          // R/unbiasedestimator.R:43-104
  Model M; model_from(spec, M);
  int dimh = 0; check(ubm_h_dim(M.kind, M.ptr(), h_kind, &dimh));
  ubm_draws dr = philox();
  ubm_bins bins = {component - 1, 0, nullptr};                        // R is 1-based, the ABI 0-based
  NumericVector br; if (breaks.isNotNull()) { br = breaks.get(); bins.nbreaks = br.size(); bins.breaks = REAL(br); }
  const int64_t n = (int64_t)nrep; const int nbins = bins.nbreaks > 1 ? bins.nbreaks - 1 : 0;
  NumericMatrix uest(dimh, n), mcmc(dimh, n), corr(dimh, n), hist(nbins > 0 ? nbins : 1, n);   // column r = replicate r
  NumericVector tau(n), cost(n), summary(ubm_summary_len(dimh, nbins));
  std::vector<int64_t> iter(n);
  ubm_estimator_out out = {REAL(uest), REAL(mcmc), REAL(corr), REAL(tau), iter.data(), REAL(cost), nbins ? REAL(hist) : nullptr, REAL(summary)};
  check(ubm_sample_unbiasedestimator(handle(), M.kind, M.ptr(), &dr, h_kind, nbins ? &bins : nullptr, (int64_t)k, (int64_t)m,
                                     (int64_t)lag, maxit(max_iterations), n, 0, &out));
  return List::create(_["uestimator"] = uest, _["mcmcestimator"] = mcmc, _["correction"] = corr, _["meetingtime"] = tau,
                      _["iteration"] = NumericVector(iter.begin(), iter.end()), _["cost"] = cost, _["bins"] = hist,
                      _["summary"] = summary);
}

// [[Rcpp::export]]
List ubm_sample_coupled_chains_(List spec, double m, double lag, double max_iterations, double stride, double nrep) {   // R/coupled_chains.R:39-87
  Model M; model_from(spec, M);
  int dim = 0; check(ubm_model_dim(M.kind, M.ptr(), &dim));
  const int64_t n = (int64_t)nrep, rows = (int64_t)stride;
  ubm_draws dr = philox();
  NumericVector s1((double)n * rows * dim), s2((double)n * rows * dim), tau(n), cost(n);   // [nrep][stride][dim], rows never reached are NaN
  std::vector<int64_t> iter(n);
  check(ubm_sample_coupled_chains(handle(), M.kind, M.ptr(), &dr, (int64_t)m, (int64_t)lag, maxit(max_iterations), rows, n, 0,
                                  REAL(s1), REAL(s2), REAL(tau), iter.data(), REAL(cost)));
  return List::create(_["samples1"] = s1, _["samples2"] = s2, _["meetingtime"] = tau, _["iteration"] = NumericVector(iter.begin(), iter.end()),
                      _["cost"] = cost, _["dim"] = dim, _["stride"] = stride, _["lag"] = lag);
}

// [[Rcpp::export]]
NumericMatrix ubm_h_bar_(List cc, int h_kind, double k, double m) {                            // R/H_bar.R:33-71
  const int dim = as<int>(cc["dim"]); const int64_t rows = (int64_t)as<double>(cc["stride"]);
  NumericVector s1 = cc["samples1"], s2 = cc["samples2"], tau = cc["meetingtime"], it = cc["iteration"];
  const int64_t n = tau.size();
  std::vector<int64_t> iter(it.begin(), it.end());
  const int dimh = (h_kind == UBM_H_BOTH) ? 2 * dim : dim;
  NumericMatrix out(dimh, n);
  check(ubm_h_bar(handle(), dim, h_kind, (int64_t)k, (int64_t)m, (int64_t)as<double>(cc["lag"]), rows, n, REAL(s1), REAL(s2), REAL(tau),
                  iter.data(), REAL(out)));
  return out;
}

// [[Rcpp::export]]
NumericMatrix ubm_estimator_bins_(List cc, int component, NumericVector breaks, double k, double m) {   // src/estimator_bin.cpp:10-42
  const int dim = as<int>(cc["dim"]); const int64_t rows = (int64_t)as<double>(cc["stride"]);
  NumericVector s1 = cc["samples1"], s2 = cc["samples2"], tau = cc["meetingtime"];
  const int64_t n = tau.size();
  ubm_bins bins = {component - 1, (int32_t)breaks.size(), REAL(breaks)};
  NumericMatrix out(breaks.size() - 1, n);
  check(ubm_estimator_bins(handle(), dim, &bins, (int64_t)k, (int64_t)m, (int64_t)as<double>(cc["lag"]), rows, n, REAL(s1), REAL(s2),
                           REAL(tau), REAL(out)));
  return out;
}

// ---- coupling primitives: same R signatures, n = 1 for a drop-in call ---------------------------------------------------
static List pair_out(NumericVector xy, int identical) { return List::create(_["xy"] = xy, _["identical"] = (bool)identical); }

// [[Rcpp::export]]
List rnorm_max_coupling_(double mu1, double mu2, double sigma1, double sigma2) {              // R/rnorm_max_coupling.R:17-23
  ubm_draws dr = philox(); NumericVector xy(2); int32_t id = 0;
  check(ubm_rnorm_max_coupling(handle(), &dr, 1, 0, mu1, mu2, sigma1, sigma2, REAL(xy), &id));
  return pair_out(xy, id);
}
// [[Rcpp::export]]
List rnorm_reflectionmax_(double mu1, double mu2, double sigma) {                             // R/rnorm_reflectionmax.R:18-39
  ubm_draws dr = philox(); NumericVector xy(2); int32_t id = 0;
  check(ubm_rnorm_reflectionmax(handle(), &dr, 1, 0, mu1, mu2, sigma, REAL(xy), &id));
  return pair_out(xy, id);
}
// [[Rcpp::export]]
List rgamma_coupled_(double alpha1, double alpha2, double beta1, double beta2) {              // R/gamma_couplings.R:14-21
  ubm_draws dr = philox(); NumericVector xy(2); int32_t id = 0;
  check(ubm_rgamma_coupled(handle(), &dr, 1, 0, alpha1, alpha2, beta1, beta2, REAL(xy), &id));
  return pair_out(xy, id);
}
// [[Rcpp::export]]
List rinversegamma_coupled_(double alpha1, double alpha2, double beta1, double beta2) {       // R/invgamma_couplings.R:29-36
  ubm_draws dr = philox(); NumericVector xy(2); int32_t id = 0;
  check(ubm_rinversegamma_coupled(handle(), &dr, 1, 0, alpha1, alpha2, beta1, beta2, REAL(xy), &id));
  return pair_out(xy, id);
}
// [[Rcpp::export]]
NumericVector rinvgaussian_coupled_(double mu1, double mu2, double lambda1, double lambda2) { // src/inversegaussian.cpp:33-63
  ubm_draws dr = philox(); NumericVector xy(2); int32_t id = 0;
  check(ubm_rinvgaussian_coupled(handle(), &dr, 1, 0, mu1, mu2, lambda1, lambda2, REAL(xy), &id));
  return xy;
}
// [[Rcpp::export]]
NumericVector rpositive_(int kind, int n, double a, double b) {                               // rgamma / rinversegamma / rinvgaussian_c
  ubm_draws dr = philox(); NumericVector out(n);
  check(ubm_rpositive(handle(), &dr, kind, n, 0, a, b, REAL(out)));
  return out;
}
// [[Rcpp::export]]
NumericMatrix fast_rmvnorm_(int nparticles, NumericVector mean, NumericMatrix covariance) {   // src/mvnorm.cpp:9-26
  ubm_draws dr = philox(); const int d = mean.size();
  NumericMatrix out(d, nparticles);                                                         // [n][d] row-major = d x n column-major
  check(ubm_fast_rmvnorm(handle(), &dr, nparticles, 0, d, REAL(mean), REAL(covariance), REAL(out)));
  return transpose(out);
}
// [[Rcpp::export]]
NumericVector fast_dmvnorm_(NumericMatrix x, NumericVector mean, NumericMatrix covariance) {  // src/mvnorm.cpp:49-67
  NumericMatrix xt = transpose(x); NumericVector out(x.nrow());
  check(ubm_fast_dmvnorm(handle(), x.nrow(), x.ncol(), REAL(xt), REAL(mean), REAL(covariance), REAL(out)));
  return out;
}
// [[Rcpp::export]]
List ising_swap_counts_(int nchains) {                                                        // R/ising.R:76,137-140
  int64_t att = 0; std::vector<int64_t> a1(nchains), a2(nchains);
  check(ubm_ising_swap_counts(handle(), &att, a1.data(), a2.data()));
  return List::create(_["nswap_attempts"] = (double)att, _["nswap_accepts1"] = NumericVector(a1.begin(), a1.end() - 1),
                      _["nswap_accepts2"] = NumericVector(a2.begin(), a2.end() - 1));
}
