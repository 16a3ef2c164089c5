# R layer of the drop-in: the reference's names, argument order, defaults and return fields (SURVEY.md Appendix C) over
# the batched shims of src/ubm_shim.cpp.  `nrep` is the only new argument (default 1 = the reference's behaviour).

h_kind <- function(h) {                                   # h must be one of the device registry's test functions
  if (identical(h, "square") || identical(body(h), body(function(x) x^2))) return(1L)
  if (identical(h, "both")) return(2L)
  0L                                                      # function(x) x, the default of R/unbiasedestimator.R:43
}
spec_of <- function(single_kernel, coupled_kernel, rinit) {
  spec <- attr(coupled_kernel, "ubm_spec")
  if (is.null(spec)) stop("ubm: kernels must come from the device registry (get_mh_kernels, ising_pt_kernels, ...)")
  utils::modifyList(spec, as.list(attr(rinit, "ubm_spec")))
}
registry_kernels <- function(spec) {
  k <- list(single_kernel = function(state) stop("runs on the device"),
            coupled_kernel = function(state1, state2) stop("runs on the device"),
            rinit = function() stop("runs on the device"))
  for (n in names(k)) attr(k[[n]], "ubm_spec") <- spec
  k
}

normal_mixture_target <- function(weights, means, sds) {  # README.Rmd:69-72
  f <- function(x) stop("runs on the device")
  attr(f, "ubm_spec") <- list(kind = "normal_mixture", weights = weights, means = means, sds = sds)
  f
}
get_mh_kernels <- function(target, Sigma_proposal, coupling = "reflectionmax") {   # R/get_mh_kernels.R:20
  spec <- attr(target, "ubm_spec")
  if (is.null(spec)) stop("ubm: target must come from the device registry (normal_mixture_target, mvn_target)")
  if (spec$kind == "normal_mixture") {
    spec$proposal_sd <- sqrt(as.numeric(Sigma_proposal)); spec$coupling <- if (coupling == "max") 1L else 0L
  } else {
    U <- chol(Sigma_proposal); spec$chol_prop <- U; spec$chol_inv_prop <- solve(U)
  }
  registry_kernels(spec)[c("single_kernel", "coupled_kernel")]
}
normal_rinit <- function(mean, sd_or_Sigma) {             # README.Rmd:80-84 / mvnorm scripts
  f <- function() stop("runs on the device")
  attr(f, "ubm_spec") <- if (length(mean) == 1) list(init_mean = mean, init_sd = sd_or_Sigma)
                         else list(init_mean = mean, init_chol = chol(sd_or_Sigma))
  f
}
mvn_target <- function(mean, Sigma) {
  f <- function(x) stop("runs on the device")
  attr(f, "ubm_spec") <- list(kind = "mvn", mean_pi = mean, chol_inv_pi = solve(chol(Sigma)))
  f
}
ising_pt_kernels <- function(betas, proba_swapmove)       # R/ising.R:45-141
  registry_kernels(list(kind = "ising_pt", betas = betas, proba_swapmove = proba_swapmove))
logisticregression_kernels <- function(Y, X, b, B, mode = "max")   # germancredit.run.R:23-49
  registry_kernels(list(kind = "pg_logistic", X = X, Y = as.numeric(Y), b = b, B = B, mode = mode))
get_blasso <- function(Y, X, lambda) {                      # R/bayesianlasso.R:22
  k <- registry_kernels(list(kind = "blasso", X = X, Y = as.numeric(Y), lambda = lambda))
  list(rinit = k$rinit, gibbs_kernel = k$single_kernel, coupled_gibbs_kernel = k$coupled_kernel)
}
get_variableselection <- function(Y, X, g, kappa, s0, proportion_singleflip)   # R/variableselection.R:1
  registry_kernels(list(kind = "varsel", X = X, Y = as.numeric(Y), g = g, kappa = kappa, s0 = s0,
                        proportion_singleflip = proportion_singleflip))

sample_meetingtime <- function(single_kernel, coupled_kernel, rinit, lag = 1, max_iterations = Inf, nrep = 1) {
  ubm_sample_meetingtime_(spec_of(single_kernel, coupled_kernel, rinit), lag, max_iterations, nrep)   # R/meetingtime.R:24
}
sample_coupled_chains <- function(single_kernel, coupled_kernel, rinit, m = 1, lag = 1, max_iterations = Inf,
                                  preallocate = 10, nrep = 1, stride = NULL) {                        # R/coupled_chains.R:39
  if (is.null(stride)) stride <- if (is.finite(max_iterations)) max_iterations + 2 else 10 * (m + preallocate + lag) + 2
  ubm_sample_coupled_chains_(spec_of(single_kernel, coupled_kernel, rinit), m, lag, max_iterations, stride, nrep)
}
sample_unbiasedestimator <- function(single_kernel, coupled_kernel, rinit, h = function(x) x, k = 0, m = 1, lag = 1,
                                     max_iterations = Inf, nrep = 1) {
  if (k > m) { print("error: k has to be less than m"); return(NULL) }      # R/unbiasedestimator.R:44-47
  if (lag > m) { print("error: lag has to be less than m"); return(NULL) }  # :48-51
  ubm_sample_unbiasedestimator_(spec_of(single_kernel, coupled_kernel, rinit), h_kind(h), NULL, 1L, k, m, lag, max_iterations, nrep)
}
H_bar <- function(c_chains, h = function(x) x, k = 0, m = 1) {              # R/H_bar.R:33
  if (k > m) { print("error: k has to be less than m"); return(NULL) }
  ubm_h_bar_(c_chains, h_kind(h), k, m)
}
histogram_c_chains <- function(c_chains, component, k, m, breaks = NULL, nclass = 30, dopar = FALSE) {   # R/histogram_c_chains.R
  if (is.null(breaks)) stop("ubm: pass breaks (find_breaks of the reference works on the returned samples)")
  prop <- rowMeans(ubm_estimator_bins_(c_chains, component, breaks, k, m))
  list(mids = (breaks[-1] + breaks[-length(breaks)]) / 2, breaks = breaks, proportions = prop, width = diff(breaks))
}

rnorm_max_coupling <- function(mu1, mu2, sigma1, sigma2) rnorm_max_coupling_(mu1, mu2, sigma1, sigma2)
rnorm_reflectionmax <- function(mu1, mu2, sigma) rnorm_reflectionmax_(mu1, mu2, sigma)
rgamma_coupled <- function(alpha1, alpha2, beta1, beta2) rgamma_coupled_(alpha1, alpha2, beta1, beta2)
rinversegamma <- function(n, alpha, beta) rpositive_(1L, n, alpha, beta)
dinversegamma <- function(x, alpha, beta) alpha * log(beta) - lgamma(alpha) - (alpha + 1) * log(x) - beta / x
rinversegamma_coupled <- function(alpha1, alpha2, beta1, beta2) rinversegamma_coupled_(alpha1, alpha2, beta1, beta2)
rinvgaussian <- function(n, mu, lambda) rpositive_(2L, n, mu, lambda)
dinvgaussian <- function(x, mu, lambda) 0.5 * log(lambda / (2 * pi)) - 1.5 * log(x) - lambda * (x - mu)^2 / (2 * mu^2 * x)
rinvgaussian_coupled <- function(mu1, mu2, lambda1, lambda2) rinvgaussian_coupled_(mu1, mu2, lambda1, lambda2)
fast_rmvnorm <- function(nparticles, mean, covariance) fast_rmvnorm_(nparticles, mean, covariance)          # R/mvnorm.R:12
fast_dmvnorm <- function(x, mean, covariance) fast_dmvnorm_(x, mean, covariance)                            # R/mvnorm.R:48
