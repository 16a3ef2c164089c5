/*
 * ubm.h — C ABI of libubm.so, the B200-native engine for the coupled-chain hot path of
 * pierrejacob/unbiasedmcmc.
 *
 * This is the drop-in boundary.  In the reference the boundary is R `.Call` into the 28
 * `RcppExport SEXP _unbiasedmcmc_*` wrappers registered in src/RcppExports.cpp:388-423, one coupled
 * pair per call, driven by the interpreted loops in R/coupled_chains.R:39-87, R/meetingtime.R:24-48 and
 * R/unbiasedestimator.R:43-104.  A GPU cannot run R closures, so the replacement boundary is BATCHED:
 * one call runs `nrep` independent replicates of a registered model to completion on the device.
 *
 * Conventions
 *   - plain C, no C++/torch/SEXP types; every function returns an int status (UBM_OK == 0);
 *     ubm_last_error(handle) gives the message for the last non-zero status on that handle.
 *   - the caller owns every buffer it passes; the library keeps no caller pointer after return.
 *     Unless a name ends in `_dev`, pointers are HOST pointers and the call is synchronous.
 *   - matrices are column-major (R layout) unless stated; indices are 0-based (the R shim converts).
 *   - there is NO CPU fallback: without a CUDA device ubm_create fails with UBM_ERR_CUDA.
 *   - replicate r of a call uses global index rep_offset + r; draws are keyed by (seed, global index),
 *     so results do not depend on how replicates are split over calls, GPUs or thread blocks.
 */
#ifndef UBM_H
#define UBM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UBM_OK 0
#define UBM_ERR_INVALID 1      /* bad argument (the R layer's print("error...") / stop() cases) */
#define UBM_ERR_CUDA 2         /* CUDA runtime error or no device */
#define UBM_ERR_UNSUPPORTED 3  /* model / option not in the device registry (no CPU fallback) */
#define UBM_ERR_TAPE 4         /* replay tape exhausted */
#define UBM_ERR_NOMEM 5

typedef struct ubm_handle ubm_handle;

/* ---------------------------------------------------------------------------------------------
 * handle
 * ------------------------------------------------------------------------------------------- */
int ubm_create(int device, ubm_handle** out);
int ubm_destroy(ubm_handle* h);
const char* ubm_last_error(const ubm_handle* h);
/* device name, SM count, and the library's build arch, for logs */
int ubm_device_info(const ubm_handle* h, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches) */
int64_t ubm_launch_count(const ubm_handle* h);
/* live roofline denominators that MEASURED_PEAKS.json lacks: FP64 FMA TFLOP/s (2 flop per fma) and INT32
 * multiply-add Top/s, from saturating micro-kernels on this device */
int ubm_measure_peaks(ubm_handle* h, double* fp64_tflops, double* int32_tops);
/* Ising parallel tempering: the swap-move counters the reference's kernels return (R/ising.R:48-49,52,60,76 and :89-91,95,
 * 108,118,137-140), summed over every replicate and iteration of the LAST run of that model on this handle:
 * swap-move iterations, accepted swaps of chain 1 and of chain 2 per adjacent pair [nchains - 1] (any pointer may be NULL) */
int ubm_ising_swap_counts(ubm_handle* h, int64_t* nswap_attempts, int64_t* nswap_accepts1, int64_t* nswap_accepts2);
/* the same for the FP64 tensor pipe (mma.m8n8k4.f64, 2 flop per fma): the pipe of C2's triangular products and C3's Gram */
int ubm_measure_fp64_tensor_peak(ubm_handle* h, double* fp64_tensor_tflops);

/* ---------------------------------------------------------------------------------------------
 * draw sources.  Replaces R's global RNG stream (GetRNGstate/PutRNGstate brackets, e.g.
 * src/mvnorm.cpp:15-17).  PHILOX: typed counter-based streams of include/ubm_numerics.h.
 * TAPE: replay mode — three typed tapes per replicate (uniforms, standard normals, Exp(1) draws)
 * consumed in the order of SURVEY.md Appendix A; running off a tape is UBM_ERR_TAPE.
 * ------------------------------------------------------------------------------------------- */
#define UBM_DRAWS_PHILOX 0
#define UBM_DRAWS_TAPE 1

typedef struct {
  int32_t mode;
  int32_t reserved;
  uint64_t seed;          /* PHILOX */
  const double* tape_u;   /* TAPE: [nrep][len_u] row-major */
  const double* tape_n;   /*       [nrep][len_n] */
  const double* tape_e;   /*       [nrep][len_e] */
  int64_t len_u, len_n, len_e;
} ubm_draws;

/* ---------------------------------------------------------------------------------------------
 * model registry (north_star: "targets are chosen from a device-side registry").
 * A model bundles what the reference passes as the three closures single_kernel / coupled_kernel /
 * rinit (R/coupled_chains.R:39).
 * ------------------------------------------------------------------------------------------- */
#define UBM_COUPLING_REFLECTION_MAX 0  /* R/rnorm_reflectionmax.R:18, src/mvnorm.cpp:185 */
#define UBM_COUPLING_MAX 1             /* R/get_max_coupling.R:24, R/rnorm_max_coupling.R:17 */

/* C1: univariate Normal-mixture target, random-walk MH.
 * target: README.Rmd:69-72; kernels: R/get_mh_kernels.R:20-79 (reflection-max) or
 * inst/reproducebimodal/bimodal.run.R:14-63 (max coupling); rinit: README.Rmd:80-84. */
typedef struct {
  int32_t ncomp;          /* <= 8 */
  int32_t coupling;
  const double* weights;  /* [ncomp] mixture weights (log taken with ubm_log at registration) */
  const double* means;    /* [ncomp] */
  const double* sds;      /* [ncomp] */
  double proposal_sd;     /* chol(Sigma_proposal)[1,1] */
  double init_mean, init_sd;
} ubm_model_normal_mixture;

/* C2: multivariate Normal target N(mean_pi, Sigma_pi), random-walk MH with reflection-max coupled
 * proposals.  Kernels: R/get_mh_kernels.R:20-79 with src/mvnorm.cpp:30-46 (proposal),
 * src/mvnorm.cpp:71-86 (target density via Cholesky inverse), src/mvnorm.cpp:185-236 (coupling);
 * shape: inst/reproducescalingdimension/scaling.rwmh.reflectionmaxcoupling.R:12-103.
 * All matrices d x d column-major, UPPER triangular as produced by R's chol() / solve(chol()). */
typedef struct {
  int32_t d;              /* 2 <= d <= 128 */
  int32_t reserved;
  const double* mean_pi;            /* [d] */
  const double* chol_inv_pi;        /* solve(chol(Sigma_pi)) */
  const double* chol_prop;          /* chol(Sigma_proposal) */
  const double* chol_inv_prop;      /* solve(chol(Sigma_proposal)) */
  const double* init_mean;          /* [d]  rinit: fast_rmvnorm_chol(1, init_mean, init_chol) */
  const double* init_chol;          /* d x d upper */
} ubm_model_mvn;

/* C4: 32 x 32 periodic Ising model, parallel tempering over `nchains` inverse temperatures, single-site
 * systematic-scan Gibbs sweeps.  Kernels: R/ising.R:45-77 (single), :87-141 (coupled: shared uniforms per
 * site and per swap), rinit R/ising.R:30-36; sweeps src/ising.cpp:43-92, natural statistic src/ising.cpp:11-35.
 * The chain state exposed to h / trajectories is the vector of natural statistics (dim = nchains), as in
 * ising_unbiased_estimator (R/ising.R:148-240). */
#define UBM_ISING_SCAN_SEQUENTIAL 0   /* the reference's in-place row-major scan, reproduced exactly */
typedef struct {
  int32_t size;           /* must be 32 */
  int32_t nchains;        /* 1 .. 32 */
  const double* betas;    /* [nchains] */
  double proba_swapmove;
  int32_t scan;
  int32_t reserved;
} ubm_model_ising_pt;

/* C3: Bayesian logistic regression, Polya-Gamma Gibbs sampler (Polson, Scott & Windle), coupled as in
 * inst/reproducegermancredit/germancredit.run.R:23-49: PG draws max-coupled per observation
 * (src/logisticregressioncoupling.cpp:79-115 over src/PolyaGamma.cpp:175-226), beta from a max-coupled pair of
 * multivariate Normals (R/logistic_regression.R:180-209, src/mvnorm.cpp:91-129) whose parameters come from
 * m_sigma_function_ (src/logisticregression.cpp:32-56).  Prior beta ~ N(b, B); rinit: beta ~ N(b, B), w = 0.
 * The chain state exposed to h / trajectories is beta (dim = p), as the reference's scripts store it. */
typedef struct {
  int32_t n, p;           /* n <= 4096 observations, p <= 60 covariates */
  const double* X;        /* n x p column-major design matrix */
  const double* Y;        /* [n] responses in {0, 1} */
  const double* b;        /* [p] prior mean */
  const double* B;        /* p x p prior covariance (column-major, symmetric positive definite) */
  int32_t w_coupling;     /* sample_w(mode=) of R/logistic_regression.R:166-175: 0 = 'max' (w_max_couplingC, the scripts'
                             choice), 1 = 'rej_samp' (w_rejsamplerC, src/logisticregressioncoupling.cpp:42-75) */
  int32_t beta_coupling;  /* coupled beta-step: 0 = maximal coupling of the two Normals (sample_beta, R/logistic_regression.R:180-209,
                             germancredit.run.R:36), 1 = common random numbers: one increment for both chains
                             (pgg_coupled_kernel, vignettes/polyagammagibbs.Rmd:189-214) */
} ubm_model_pg_logistic;

/* C5: Bayesian variable selection, g-prior spike-and-slab, flip / swap Metropolis-Hastings on gamma in {0,1}^p,
 * coupled as in R/variableselection.R:57-148 (shared flip coordinate, max-coupled swap indices
 * R/coupled_pairs01.R:3-40, shared accept uniform).  Marginal likelihood: src/vs_mlikelihood.cpp:5-28, evaluated in
 * FP64 from G = X'X and X'Y (formed once on the device) with a Cholesky factorisation of G[gamma, gamma].
 * Prior: -kappa |gamma| log p, -Inf when |gamma| > s0 (R/variableselection.R:8-11).  The chain state exposed to
 * h / trajectories is gamma as 0/1 doubles (dim = p): H estimates the posterior inclusion probabilities. */
typedef struct {
  int32_t n, p;                 /* p <= 1024 */
  const double* X;              /* n x p column-major */
  const double* Y;              /* [n] */
  double g, kappa;
  int32_t s0;                   /* <= 128 */
  int32_t reserved;
  double proportion_singleflip;
} ubm_model_varsel;

/* Bayesian lasso: the Gibbs sampler of get_blasso and its coupling, R/bayesianlasso.R:22-98 with
 * src/blassoutil.cpp:17-88 (beta | tau2, sigma2 through rmvnorm_max_coupling_), rinversegamma_coupled for sigma2 and
 * rinvgaussian_coupled per component of 1 / tau2.  Chain state = c(beta[p], tau2[p], sigma2): dim = 2 p + 1.
 * rinit is the reference's deterministic c(0.., 1.., 1). */
typedef struct {
  int32_t n, p;           /* n <= 4096 observations, p <= 60 covariates */
  const double* X;        /* n x p column-major */
  const double* Y;        /* [n] */
  double lambda;          /* the lasso penalty (lambda2 = lambda^2, R/bayesianlasso.R:28) */
} ubm_model_blasso;

/* ---------------------------------------------------------------------------------------------
 * test functions h and histogram bins
 * ------------------------------------------------------------------------------------------- */
#define UBM_H_IDENTITY 0   /* h(x) = x (the R default, R/unbiasedestimator.R:43); dimh = dim */
#define UBM_H_SQUARE 1     /* h(x) = x^2 componentwise (inst/check/check_lags.R:60) */
#define UBM_H_BOTH 2       /* h(x) = (x, x^2); dimh = 2 dim */

/* estimator_bin_ (src/estimator_bin.cpp:10-42) for every bin [breaks[b], breaks[b+1]) of one
 * component, membership strict on both sides as in the reference. nbreaks == 0 disables bins. */
typedef struct {
  int32_t component;      /* 0-based */
  int32_t nbreaks;        /* <= 257 */
  const double* breaks;   /* [nbreaks] increasing */
} ubm_bins;

/* Aggregates over the uncensored replicates of one call, in one flat double buffer so that a
 * multi-GPU caller can all-reduce it:
 *   [0] n_done  [1] n_censored  [2] sum tau  [3] sum tau^2  [4] sum cost  [5] sum iteration
 *   [6 .. 6+dimh)            sum_r uestimator_r
 *   [6+dimh .. 6+2dimh)      sum_r uestimator_r^2
 *   [6+2dimh .. +nbins)      sum_r bin_r        (bin_r = estimator_bin value of replicate r)
 *   [6+2dimh+nbins .. +nbins) sum_r bin_r^2
 * Sums over replicates run in replicate-index order with a fixed tree, so they are reproducible. */
#define UBM_SUMMARY_HEAD 6
int64_t ubm_summary_len(int32_t dimh, int32_t nbins);

typedef struct {
  /* per replicate; any pointer may be NULL */
  double* uestimator;     /* [nrep][dimh]  R/unbiasedestimator.R:102 */
  double* mcmcestimator;  /* [nrep][dimh] */
  double* correction;     /* [nrep][dimh] */
  double* meetingtime;    /* [nrep]  +Inf when max_iterations was hit (R/coupled_chains.R:57) */
  int64_t* iteration;     /* [nrep] */
  double* cost;           /* [nrep]  lag + 2 (tau - lag) + max(0, time - tau), +Inf if censored */
  double* bins;           /* [nrep][nbins] */
  /* aggregate */
  double* summary;        /* [ubm_summary_len(dimh, nbins)] */
} ubm_estimator_out;

/* ---------------------------------------------------------------------------------------------
 * the three drivers, batched.
 * `model_kind` + `model` select the registry entry: model points at the matching ubm_model_* struct.
 * ------------------------------------------------------------------------------------------- */
#define UBM_MODEL_NORMAL_MIXTURE 1
#define UBM_MODEL_MVN 2
#define UBM_MODEL_PG_LOGISTIC 3
#define UBM_MODEL_ISING_PT 4
#define UBM_MODEL_VARSEL 5
#define UBM_MODEL_BLASSO 6

/* dimension of the chain state / of h for a model (so callers can size buffers) */
int ubm_model_dim(int32_t model_kind, const void* model, int32_t* dim);
int ubm_h_dim(int32_t model_kind, const void* model, int32_t h_kind, int32_t* dimh);

/* sample_meetingtime (R/meetingtime.R:24-48): lag >= 1; max_iterations <= 0 means Inf.
 * meetingtime[r] = +Inf if the replicate was stopped by max_iterations. */
int ubm_sample_meetingtime(ubm_handle* h, int32_t model_kind, const void* model, const ubm_draws* draws,
                           int64_t lag, int64_t max_iterations, int64_t nrep, int64_t rep_offset,
                           double* meetingtime);

/* sample_unbiasedestimator (R/unbiasedestimator.R:43-104) + estimator_bin_ for all bins, on the fly.
 * Requires k <= m and lag <= m (the reference prints an error and returns NULL: UBM_ERR_INVALID). */
int ubm_sample_unbiasedestimator(ubm_handle* h, int32_t model_kind, const void* model, const ubm_draws* draws,
                                 int32_t h_kind, const ubm_bins* bins,
                                 int64_t k, int64_t m, int64_t lag, int64_t max_iterations,
                                 int64_t nrep, int64_t rep_offset, const ubm_estimator_out* out);

/* sample_coupled_chains (R/coupled_chains.R:39-87): trajectories to HBM and back (parity / debug path).
 * Each replicate gets `stride` rows of storage: samples1[r] is [stride][dim] row-major holding
 * X_0..X_iteration, samples2[r] holds Y_0..Y_{iteration-lag}.  Replicates that would need more than
 * `stride` rows are stopped as if max_iterations = stride - 1. */
int ubm_sample_coupled_chains(ubm_handle* h, int32_t model_kind, const void* model, const ubm_draws* draws,
                              int64_t m, int64_t lag, int64_t max_iterations, int64_t stride,
                              int64_t nrep, int64_t rep_offset,
                              double* samples1, double* samples2,
                              double* meetingtime, int64_t* iteration, double* cost);

/* H_bar (R/H_bar.R:19-54) and estimator_bin_ (src/estimator_bin.cpp:10-42) from stored chains, batched
 * over replicates: the post-processing half of the path, run on the device over the layout produced by
 * ubm_sample_coupled_chains. */
int ubm_h_bar(ubm_handle* h, int32_t dim, int32_t h_kind, int64_t k, int64_t m, int64_t lag, int64_t stride,
              int64_t nrep, const double* samples1, const double* samples2,
              const double* meetingtime, const int64_t* iteration, double* hbar /* [nrep][dimh] */);
int ubm_estimator_bins(ubm_handle* h, int32_t dim, const ubm_bins* bins, int64_t k, int64_t m, int64_t lag,
                       int64_t stride, int64_t nrep, const double* samples1, const double* samples2,
                       const double* meetingtime, double* binvals /* [nrep][nbins] */);

/* Signed empirical measure of stored coupled chains: c_chains_to_measure_as_list_ (src/c_chains_to_measure.cpp:38-84,
 * R/c_chains_to_measure_as_list.R:19-27), batched over replicates.  Replicate r yields
 *   size_r = (m - k + 1) + 2 max(0, tau_r - (k + lag))   atoms,
 * first the MCMC part (X_t, t = k..m, weight 1 / (m-k+1)), then for t = k+lag .. tau-1 the pair X_t (+w_t) and
 * Y_{t-lag} (-w_t) with w_t the estimator's weight / (m-k+1).  Outputs are padded to `cap` rows per replicate
 * (rows >= size_r are zeroed); UBM_ERR_INVALID if some size_r > cap, if k > m or m > iteration_r (the R stops), or if a
 * replicate is censored (tau = Inf).  mcmc_part: 1 = MCMC atom, 0 = bias-correction atom (the reference's `MCMC`). */
int64_t ubm_measure_size(int64_t k, int64_t m, int64_t lag, double meetingtime);
int ubm_c_chains_to_measure(ubm_handle* h, int32_t dim, int64_t k, int64_t m, int64_t lag, int64_t stride, int64_t nrep,
                            const double* samples1, const double* samples2, const double* meetingtime,
                            const int64_t* iteration, int64_t cap, double* atoms /* [nrep][cap][dim] */,
                            double* weights /* [nrep][cap] */, int32_t* mcmc_part /* [nrep][cap] */,
                            int64_t* sizes /* [nrep] */);

/* H_bar for a GRID of (k, m) pairs from one upload of the stored chains (the tuning sweeps of
 * inst/reproducegermancredit/germancredit.run.R:171-186 and inst/reproducebimodal/bimodal.plots.R:43-46 call H_bar once
 * per k).  hbar: [npairs][nrep][dim h]; every entry equals ubm_h_bar with that (k, m) bit for bit. */
int ubm_h_bar_grid(ubm_handle* h, int32_t dim, int32_t h_kind, int32_t npairs, const int64_t* ks, const int64_t* ms,
                   int64_t lag, int64_t stride, int64_t nrep, const double* samples1, const double* samples2,
                   const double* meetingtime, const int64_t* iteration, double* hbar);

/* L-lag total-variation upper bounds from meeting times (Biswas, Jacob & Vanetti 2019), as computed by the callers
 * directly above the path (vignettes/polyagammagibbs.Rmd:238-244, inst/reproducegermancredit/germancredit.run.R:171-186):
 *   bound[j] = mean_r max(0, ceil((tau_r - lag - t[j]) / lag)).
 * Integer sums, fixed order: the result does not depend on the launch geometry.  UBM_ERR_INVALID for a censored tau. */
int ubm_tv_upper_bounds(ubm_handle* h, const double* meetingtime, int64_t nrep, int64_t lag, const int64_t* t, int32_t nt,
                        double* bound /* [nt] */);

/* ---------------------------------------------------------------------------------------------
 * stand-alone coupling / MVN primitives of the R API, batched over n samples.  Sample i draws from the typed
 * streams of replicate index rep_offset + i (from draw 0), or from row i of the tapes.
 *   ubm_rnorm_max_coupling        R/rnorm_max_coupling.R:17-23 (R/get_max_coupling.R:24-39)   xy [n][2]
 *   ubm_rnorm_reflectionmax       R/rnorm_reflectionmax.R:18-39                                xy [n][2]
 *   ubm_rmvnorm_reflectionmax     R/mvnorm_couplings.R:94-100 -> src/mvnorm.cpp:185-236        xy [n][2][d]
 *                                 (d = 1 is UBM_ERR_INVALID with the reference's stop() message)
 *   ubm_rmvnorm_max               R/mvnorm_couplings.R:25 -> src/mvnorm.cpp:91-129 (rmvnorm_max_coupling_)
 *                                 Sigma1, Sigma2: d x d covariances (LLT taken once on the host)  xy [n][2][d]
 *   ubm_rmvnorm_max_chol          R/mvnorm_couplings.R:57 -> src/mvnorm.cpp:133-174 (the `<=` of :152 kept) xy [n][2][d]
 *                                 nattempts [n]: rejection-loop trips (the reference returns it for the cov version)
 *   ubm_fast_rmvnorm_chol         R/mvnorm.R:32 -> src/mvnorm.cpp:30-46                        out [n][d]
 *   ubm_fast_dmvnorm_chol_inverse R/mvnorm.R:69 -> src/mvnorm.cpp:71-86    x [n][d] row-major   out [n]
 * chol / chol_inv: d x d column-major upper-triangular, as R's chol() / solve(chol()).
 * ------------------------------------------------------------------------------------------- */
int ubm_rnorm_max_coupling(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double mu1, double mu2,
                           double sigma1, double sigma2, double* xy, int32_t* identical);
int ubm_rnorm_reflectionmax(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double mu1, double mu2,
                            double sigma, double* xy, int32_t* identical);
/* the other scalar maximal couplings, all through R/get_max_coupling.R:24-39; xy [n][2], identical [n]:
 *   ubm_rgamma_coupled           R/gamma_couplings.R:14-21     Gamma(alpha, rate beta)
 *   ubm_rinversegamma_coupled    R/invgamma_couplings.R:29-36  (dinversegamma :7-9, rinversegamma :17-19)
 *   ubm_rinvgaussian_coupled     src/inversegaussian.cpp:33-63 (rinvgaussian_c :6-25, dinvgaussian_c :27-29)
 *   ubm_rpositive                n plain draws: kind 0 rgamma(shape a, rate b), 1 rinversegamma, 2 rinvgaussian(mu a, lambda b)
 * The inverse-Gaussian sampler is the reference's, draw for draw.  R's rgamma belongs to the R runtime: Gamma draws are
 * Marsaglia-Tsang on the typed streams (a replay tape reproduces this engine's draws, not R's). */
int ubm_rgamma_coupled(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double alpha1, double alpha2,
                       double beta1, double beta2, double* xy, int32_t* identical);
int ubm_rinversegamma_coupled(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double alpha1, double alpha2,
                              double beta1, double beta2, double* xy, int32_t* identical);
int ubm_rinvgaussian_coupled(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double mu1, double mu2,
                             double lambda1, double lambda2, double* xy, int32_t* identical);
int ubm_rpositive(ubm_handle* h, const ubm_draws* draws, int32_t kind, int64_t n, int64_t rep_offset, double a, double b, double* out);
int ubm_rmvnorm_reflectionmax(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d,
                              const double* mu1, const double* mu2, const double* chol, const double* chol_inv,
                              double* xy, int32_t* identical);
int ubm_rmvnorm_max(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d, const double* mu1,
                    const double* mu2, const double* Sigma1, const double* Sigma2, double* xy, int32_t* identical,
                    int32_t* nattempts);
int ubm_rmvnorm_max_chol(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d,
                         const double* mu1, const double* mu2, const double* chol1, const double* chol2,
                         const double* chol_inv1, const double* chol_inv2, double* xy, int32_t* identical,
                         int32_t* nattempts);
int ubm_fast_rmvnorm_chol(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d,
                          const double* mean, const double* chol, double* out);
int ubm_fast_dmvnorm_chol_inverse(ubm_handle* h, int64_t n, int32_t d, const double* x, const double* mean,
                                  const double* chol_inv, double* out);
/* covariance forms: fast_rmvnorm_ (src/mvnorm.cpp:9-26; R/mvnorm.R:12) and fast_dmvnorm_ (:49-67; R/mvnorm.R:48).  The LLT of
 * the covariance (column-major, d x d) is taken on the host, then the factor forms above run; not positive definite: UBM_ERR_INVALID */
int ubm_fast_rmvnorm(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d, const double* mean,
                     const double* covariance, double* out);
int ubm_fast_dmvnorm(ubm_handle* h, int64_t n, int32_t d, const double* x, const double* mean, const double* covariance,
                     double* out);

/* ---------------------------------------------------------------------------------------------
 * device-resident variant used for kernel-only timing (bench.py `value`): the model is uploaded once,
 * the run leaves per-replicate results and the summary in device memory owned by the handle.
 * ------------------------------------------------------------------------------------------- */
typedef struct ubm_plan ubm_plan;
int ubm_plan_create(ubm_handle* h, int32_t model_kind, const void* model, int32_t h_kind, const ubm_bins* bins,
                    int64_t k, int64_t m, int64_t lag, int64_t max_iterations, int64_t max_nrep,
                    ubm_plan** out);
/* enqueue one batch on the handle's stream; returns without synchronising */
int ubm_plan_launch(ubm_plan* p, uint64_t seed, int64_t nrep, int64_t rep_offset);
/* wait for the stream and copy the summary (and optionally per-replicate arrays) to the host */
int ubm_plan_fetch(ubm_plan* p, const ubm_estimator_out* out);
/* device pointer of the summary buffer (for NCCL all-reduce through torch) and the raw stream */
int ubm_plan_summary_dev(ubm_plan* p, void** dev_ptr, int64_t* len);
int ubm_plan_stream(ubm_plan* p, void** cuda_stream);
/* time of the last launched batch in ms, measured with CUDA events on the plan's stream */
int ubm_plan_last_ms(ubm_plan* p, float* ms_main_kernel, float* ms_total);
int ubm_plan_destroy(ubm_plan* p);

#ifdef __cplusplus
}
#endif
#endif /* UBM_H */
