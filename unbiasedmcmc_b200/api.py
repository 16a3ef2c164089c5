"""The reference's public R API, mirrored name for name over the C ABI.

R users write `kernels <- get_mh_kernels(target, Sigma_proposal)` and pass the closures
`single_kernel`, `coupled_kernel`, `rinit` to `sample_meetingtime` / `sample_coupled_chains` /
`sample_unbiasedestimator` (README.Rmd:69-99).  Closures cannot run on a GPU, so here — exactly as the Rcpp shim of
INTEGRATION.md does — the "kernels" are small objects that carry a registry spec; the drivers look the spec up and run
ALL replicates in one batched device call (`nrep`, default 1 as in the reference).  Return values keep the
reference's field names; with nrep > 1 every field gains a leading replicate axis.

Differences, all documented in DESIGN.md: a plain Python callable as target/kernel is an error (no CPU fallback);
`elapsedtime` is batch wall time / nrep; randomness is Philox keyed by (`seed`, replicate) — pass `seed=` (the R shim
draws it from R's stream so that set.seed() still governs reproducibility)."""
import time as _time

import numpy as np

from . import _abi as A
from . import registry as R
from .engine import Engine

_engine = None


def _eng():
    global _engine
    if _engine is None:
        _engine = Engine(0)
    return _engine


# ---- targets and kernels --------------------------------------------------------------------------------------
class NormalMixtureTarget:
    """target <- function(x) log-density of sum_i w_i N(mean_i, sd_i^2) (README.Rmd:69-72)."""

    def __init__(self, weights, means, sds):
        self.weights, self.means, self.sds = map(lambda a: np.asarray(a, dtype=np.float64), (weights, means, sds))


class MvnTarget:
    """target <- function(x) fast_dmvnorm_chol_inverse(x, mean, solve(chol(Sigma)))
    (scaling.rwmh.reflectionmaxcoupling.R:34)."""

    def __init__(self, mean, Sigma):
        self.mean, self.Sigma = np.asarray(mean, dtype=np.float64), np.asarray(Sigma, dtype=np.float64)


class _Kernel:
    def __init__(self, spec, role):
        self.ubm_spec, self.role = spec, role

    def __call__(self, *a, **k):
        raise RuntimeError("this kernel runs on the device: pass it to sample_meetingtime / sample_coupled_chains / "
                           "sample_unbiasedestimator (there is no CPU fallback)")


class _Rinit(_Kernel):
    pass


def get_mh_kernels(target, Sigma_proposal, coupling="reflectionmax"):
    """R/get_mh_kernels.R:20 — returns {'single_kernel', 'coupled_kernel'}.  `coupling="max"` gives the kernels of
    inst/reproducebimodal/bimodal.run.R:14-57 (d = 1 only)."""
    if callable(target) and not isinstance(target, (NormalMixtureTarget, MvnTarget)):
        raise TypeError("target must come from the device registry (NormalMixtureTarget, MvnTarget): arbitrary "
                        "closures cannot run on the GPU and there is no CPU fallback")
    Sp = np.atleast_2d(np.asarray(Sigma_proposal, dtype=np.float64))          # :21-23 scalar -> 1 x 1
    spec = {"target": target, "Sigma_proposal": Sp,
            "coupling": A.UBM_COUPLING_MAX if coupling == "max" else A.UBM_COUPLING_REFLECTION_MAX}
    return {"single_kernel": _Kernel(spec, "single"), "coupled_kernel": _Kernel(spec, "coupled")}


def normal_rinit(mean, sd_or_Sigma):
    """rinit <- function() list(chain_state = rnorm / fast_rmvnorm_chol(...), current_pdf = target(chain_state))
    (README.Rmd:80-84; scaling.rwmh.reflectionmaxcoupling.R:88-99)."""
    return _Rinit({"init_mean": np.atleast_1d(np.asarray(mean, dtype=np.float64)),
                   "init_cov": np.atleast_2d(np.asarray(sd_or_Sigma, dtype=np.float64))}, "rinit")


def _model(single_kernel, coupled_kernel, rinit):
    if isinstance(coupled_kernel, R.ModelSpec):          # registry models that own their three closures
        return coupled_kernel
    for k in (single_kernel, coupled_kernel, rinit):
        if not isinstance(k, _Kernel):
            raise TypeError("kernels and rinit must come from get_mh_kernels()/normal_rinit() or a registry model")
    spec, init = coupled_kernel.ubm_spec, rinit.ubm_spec
    tgt = spec["target"]
    if isinstance(tgt, NormalMixtureTarget):
        if spec["Sigma_proposal"].shape != (1, 1):
            raise ValueError("a univariate target needs a scalar proposal variance")
        return R.NormalMixtureRWMH(tgt.weights, tgt.means, tgt.sds, np.sqrt(spec["Sigma_proposal"][0, 0]),
                                   init["init_mean"][0], init["init_cov"][0, 0], spec["coupling"])
    if spec["coupling"] != A.UBM_COUPLING_REFLECTION_MAX:
        raise ValueError("multivariate RW-MH kernels use the reflection-maximal coupling (R/get_mh_kernels.R:46-50)")
    return R.MvnRWMH(tgt.mean, tgt.Sigma, spec["Sigma_proposal"], init["init_mean"], init["init_cov"])


def ising_pt_kernels(betas, proba_swapmove):
    """ising_pt_single_kernel / ising_pt_coupled_kernel / ising_pt_rinit (R/ising.R:30-141) as one registry model."""
    m = R.IsingPT(betas, proba_swapmove)
    return {"single_kernel": m, "coupled_kernel": m, "rinit": m}


def logisticregression_kernels(Y, X, b, B, mode="max", beta_coupling="max"):
    """single_kernel / coupled_kernel / rinit of inst/reproducegermancredit/germancredit.run.R:23-49; `mode` is
    sample_w's (R/logistic_regression.R:166): 'max' (w_max_couplingC) or 'rej_samp' (w_rejsamplerC); `beta_coupling` 'max'
    (sample_beta) or 'crn' (the common-random-number step of pgg_coupled_kernel, vignettes/polyagammagibbs.Rmd:189-214)."""
    m = R.PgLogistic(Y, X, b, B, mode=mode, beta_coupling=beta_coupling)
    return {"single_kernel": m, "coupled_kernel": m, "rinit": m}


def get_blasso(Y, X, lam):
    """get_blasso(Y, X, lambda) (R/bayesianlasso.R:22): list(rinit, gibbs_kernel, coupled_gibbs_kernel)"""
    m = R.BayesianLasso(Y, X, lam)
    return {"rinit": m, "gibbs_kernel": m, "coupled_gibbs_kernel": m}


def get_variableselection(Y, X, g, kappa, s0, proportion_singleflip):
    """R/variableselection.R:3 — returns the kernels and rinit (marginal likelihood / prior run on the device)."""
    m = R.VariableSelection(Y, X, g, kappa, s0, proportion_singleflip)
    return {"single_kernel": m, "coupled_kernel": m, "rinit": m}


# ---- drivers ----------------------------------------------------------------------------------------------------
_H = {"identity": A.UBM_H_IDENTITY, "square": A.UBM_H_SQUARE, "both": A.UBM_H_BOTH}


def _maxit(max_iterations):
    return 0 if max_iterations is None or np.isinf(max_iterations) else int(max_iterations)


def _squeeze(d, nrep):
    return {k: (v[0] if (nrep == 1 and isinstance(v, np.ndarray) and v.ndim >= 1 and k != "summary") else v)
            for k, v in d.items()}


def sample_meetingtime(single_kernel, coupled_kernel, rinit, lag=1, max_iterations=np.inf, nrep=1, seed=1):
    """R/meetingtime.R:24 — list(meetingtime, elapsedtime)."""
    mdl = _model(single_kernel, coupled_kernel, rinit)
    t0 = _time.perf_counter()
    res = _eng().sample_meetingtime(mdl, R.Draws(seed=seed), lag=lag, max_iterations=_maxit(max_iterations), nrep=nrep)
    res["elapsedtime"] = (_time.perf_counter() - t0) / nrep
    return _squeeze(res, nrep)


def sample_coupled_chains(single_kernel, coupled_kernel, rinit, m=1, lag=1, max_iterations=np.inf, preallocate=10,
                          nrep=1, seed=1, stride=None):
    """R/coupled_chains.R:39 — list(samples1, samples2, meetingtime, iteration, elapsedtime, cost).  The device needs
    a row budget per replicate (`stride`, default 4 (m + lag + preallocate)): a replicate that would outgrow it is
    stopped as if max_iterations = stride - 1 (meetingtime Inf), where R would keep doubling its matrices."""
    mdl = _model(single_kernel, coupled_kernel, rinit)
    if stride is None:
        stride = 4 * (m + lag + preallocate)
    t0 = _time.perf_counter()
    res = _eng().sample_coupled_chains(mdl, R.Draws(seed=seed), m=m, lag=lag, max_iterations=_maxit(max_iterations),
                                       stride=stride, nrep=nrep)
    res["elapsedtime"] = (_time.perf_counter() - t0) / nrep
    out = []
    for r in range(nrep):
        it = int(res["iteration"][r])
        out.append({"samples1": res["samples1"][r, :it + 1], "samples2": res["samples2"][r, :it - lag + 1],
                    "meetingtime": res["meetingtime"][r], "iteration": it, "elapsedtime": res["elapsedtime"],
                    "cost": res["cost"][r], "_batch": (res, r)})
    return out[0] if nrep == 1 else out


def sample_unbiasedestimator(single_kernel, coupled_kernel, rinit, h="identity", k=0, m=1, lag=1,
                             max_iterations=np.inf, nrep=1, seed=1):
    """R/unbiasedestimator.R:43 — list(mcmcestimator, correction, uestimator, meetingtime, iteration, elapsedtime,
    cost); returns None after printing the reference's message when k > m or lag > m (:44-51)."""
    if k > m:
        print("error: k has to be less than m")
        return None
    if lag > m:
        print("error: lag has to be less than m")
        return None
    mdl = _model(single_kernel, coupled_kernel, rinit)
    t0 = _time.perf_counter()
    res = _eng().sample_unbiasedestimator(mdl, R.Draws(seed=seed), h_kind=_H[h], k=k, m=m, lag=lag,
                                          max_iterations=_maxit(max_iterations), nrep=nrep)
    res["elapsedtime"] = (_time.perf_counter() - t0) / nrep
    res.pop("bins", None)
    return _squeeze(res, nrep)


# ---- estimators from stored chains --------------------------------------------------------------------------------
def _batch_of(c_chains):
    lst = c_chains if isinstance(c_chains, list) else [c_chains]
    lag = lst[0]["samples1"].shape[0] - lst[0]["samples2"].shape[0]          # R/H_bar.R:30
    stride = max(c["samples1"].shape[0] for c in lst)
    d = lst[0]["samples1"].shape[1]
    s1 = np.zeros((len(lst), stride, d))
    s2 = np.zeros((len(lst), stride, d))
    for r, c in enumerate(lst):
        s1[r, :c["samples1"].shape[0]] = c["samples1"]
        s2[r, :c["samples2"].shape[0]] = c["samples2"]
    return {"samples1": s1, "samples2": s2, "lag": lag,
            "meetingtime": np.array([c["meetingtime"] for c in lst], dtype=np.float64),
            "iteration": np.array([c["iteration"] for c in lst], dtype=np.int64)}, len(lst)


def H_bar(c_chains, h="identity", k=0, m=1):
    """R/H_bar.R:19 — unbiased estimator(s) from stored coupled chains (one per element when given a list)."""
    lst = c_chains if isinstance(c_chains, list) else [c_chains]
    if any(k > c["iteration"] for c in lst):
        print("error: k has to be less than the horizon of the coupled chains")
        return None
    if any(m > c["iteration"] for c in lst):
        print("error: m has to be less than the horizon of the coupled chains")
        return None
    batch, n = _batch_of(c_chains)
    out = _eng().h_bar(batch, _H[h], k=k, m=m)
    return out if isinstance(c_chains, list) else out[0]


def c_chains_to_measure_as_list(c_chains, k, m):
    """R/c_chains_to_measure_as_list.R:19 — list(atoms, weights, MCMC): the signed empirical measure
    (m-k+1)^-1 [sum_{t=k}^m delta_{X_t} + sum_{t=k+lag}^{tau-1} w_t (delta_{X_t} - delta_{Y_{t-lag}})] of one coupled chain
    (a list of chains gives a list of measures)."""
    if k > m:
        raise ValueError("k must be <= than m")                                               # :20-22
    lst = c_chains if isinstance(c_chains, list) else [c_chains]
    if any(m > c["iteration"] for c in lst):
        raise ValueError("m has to be less than the time horizon of the coupled chains")    # :23-25
    batch, n = _batch_of(c_chains)
    out = _eng().c_chains_to_measure(batch, k=k, m=m)
    return out if isinstance(c_chains, list) else out[0]


def prune_measure_(df):
    """src/prune.cpp:6-45 on a matrix with columns (rep, MCMC, weight, atom.1, ...) sorted lexicographically: consecutive
    rows of the same repeat whose atoms differ by less than 1e-20 in l1 are merged, their weights added (vectorised:
    run boundaries + segmented sums; the first row of a run is kept, as in the reference)."""
    df = np.asarray(df, dtype=np.float64)
    if df.shape[0] == 0:
        return df.copy()
    same_rep = df[1:, 0] == df[:-1, 0]
    close = np.abs(df[1:, 3:] - df[:-1, 3:]).sum(axis=1) < 1e-20
    new_run = np.concatenate([[True], ~(same_rep & close)])
    run_id = np.cumsum(new_run) - 1
    out = df[new_run].copy()
    out[:, 2] = np.bincount(run_id, weights=df[:, 2], minlength=out.shape[0])
    return out


def c_chains_to_dataframe(c_chains, k, m, dopar=False, prune=True):
    """R/c_chains_to_dataframe.R:22-53 — the signed measure of one coupled chain or of a list of them as a table with
    columns rep (1-based), MCMC, weight (divided by the number of chains), atom.1..atom.d; identical atoms of a chain
    pruned.  Returned as a dict of columns (`atom` is the (rows, d) matrix); `dopar` is accepted and ignored."""
    lst = [c_chains] if isinstance(c_chains, dict) else list(c_chains)
    measures = c_chains_to_measure_as_list(lst, k, m)
    nsamples = len(lst)
    rows = [np.column_stack([np.full(len(M["weights"]), r + 1.0), M["MCMC"].astype(np.float64), M["weights"] / nsamples,
                             M["atoms"]]) for r, M in enumerate(measures)]
    df = np.concatenate(rows, axis=0)
    if prune:
        keys = [df[:, j] for j in range(df.shape[1] - 1, 2, -1)] + [df[:, 1], df[:, 0]]      # order(rep, MCMC, atom.1, ...)
        df = prune_measure_(df[np.lexsort(keys)])
    df = df[np.lexsort([df[:, 1], df[:, 0]])]                                                # order(rep, MCMC), stable
    return {"rep": df[:, 0].astype(np.int64), "MCMC": df[:, 1] != 0, "weight": df[:, 2], "atom": df[:, 3:]}


def H_bar_grid(c_chains, ks, ms, h="identity"):
    """H_bar(c_chains, h, k, m) for a grid of k (and m, scalar or one per k) from one upload of the chains — the k / m tuning
    sweeps of germancredit.run.R:171-186 and bimodal.plots.R:43-46; returns [len(ks)][nrep][dim h]."""
    batch, n = _batch_of(c_chains)
    return _eng().h_bar_grid(batch, ks, ms, _H[h])


def tv_upper_bound(meetingtimes, lag, t):
    """vignettes/polyagammagibbs.Rmd:238-240 — pmax(0, ceiling((meetingtimes - lag - t) / lag)), per meeting time."""
    mt = np.asarray(meetingtimes, dtype=np.float64)
    return np.maximum(0.0, np.ceil((mt - lag - t) / lag))


def tv_upper_bounds(meetingtimes, lag, t):
    """The curve the callers plot (polyagammagibbs.Rmd:243): mean(tv_upper_bound(meetingtimes, lag, t)) for each t of a
    grid, reduced on the device in integer arithmetic."""
    return _eng().tv_upper_bounds(meetingtimes, lag, t)


def create_mids(breaks):
    """R/histogram_c_chains.R:77-83."""
    breaks = np.asarray(breaks, dtype=np.float64)
    return breaks[:-1] + (breaks[1:] - breaks[:-1]) / 2


def find_breaks(c_chains, component, nclass, k, m, lag):
    """R/histogram_c_chains.R:69-74: breaks of hist(all X_k..X_m and Y samples, nclass); R's pretty() breaks are
    approximated by numpy's 'nice' bin edges over the pooled range."""
    xs = np.concatenate([c["samples1"][k:m + 1, component - 1] for c in c_chains] +
                        [c["samples2"][max(k - 1, 0):m + 1 - lag, component - 1] for c in c_chains])
    lo, hi = xs.min(), xs.max()
    raw = (hi - lo) / nclass
    mag = 10.0 ** np.floor(np.log10(raw))
    step = min((1, 2, 5, 10), key=lambda q: abs(q * mag - raw)) * mag
    return np.arange(np.floor(lo / step) * step, np.ceil(hi / step) * step + step / 2, step)


def estimator_bin(c_chains, component, lower, upper, k, m, lag):
    """R/histogram_c_chains.R:86 -> src/estimator_bin.cpp:10-42 (component is 1-based, as in R)."""
    batch, _ = _batch_of(c_chains)
    return _eng().estimator_bins(batch, R.Bins(component - 1, [lower, upper]), k=k, m=m)[0, 0]


def histogram_c_chains(c_chains, component, k, m, breaks=None, nclass=30, dopar=False):
    """R/histogram_c_chains.R:13 — list(mids, breaks, proportions, sd, width)."""
    batch, n = _batch_of(c_chains)
    if breaks is None:
        breaks = find_breaks(c_chains, component, nclass, k, m, batch["lag"])
    breaks = np.asarray(breaks, dtype=np.float64)
    vals = _eng().estimator_bins(batch, R.Bins(component - 1, breaks), k=k, m=m)      # [nrep][nbins]
    return {"mids": create_mids(breaks), "breaks": breaks, "proportions": vals.mean(0),
            "sd": vals.std(0, ddof=1) / np.sqrt(n), "width": np.diff(breaks)[0]}


# ---- couplings and MVN helpers ------------------------------------------------------------------------------------
def rnorm_max_coupling(mu1, mu2, sigma1, sigma2, n=1, seed=1):
    """R/rnorm_max_coupling.R:17 — list(xy, identical)."""
    return _squeeze(_eng().rnorm_max_coupling(mu1, mu2, sigma1, sigma2, R.Draws(seed=seed), n=n), n)


def rnorm_reflectionmax(mu1, mu2, sigma, n=1, seed=1):
    """R/rnorm_reflectionmax.R:18 — list(xy, identical)."""
    return _squeeze(_eng().rnorm_reflectionmax(mu1, mu2, sigma, R.Draws(seed=seed), n=n), n)


def rmvnorm_reflectionmax(mu1, mu2, Cholesky, Cholesky_inverse, n=1, seed=1):
    """R/mvnorm_couplings.R:94 — list(xy [d x 2 in R; here (2, d) per sample], identical)."""
    if np.asarray(Cholesky).ndim < 2 or np.asarray(Cholesky).shape[0] == 1:
        raise ValueError("function 'rmvnorm_reflectionmax' is meant for multivariate Normals; for univariate Normals, "
                         "use 'rnorm_reflectionmax'")                          # :96-98
    return _squeeze(_eng().rmvnorm_reflectionmax(mu1, mu2, Cholesky, Cholesky_inverse, R.Draws(seed=seed), n=n), n)


def rmvnorm_max(mu1, mu2, Sigma1, Sigma2, n=1, seed=1):
    """R/mvnorm_couplings.R:25 — list(xy, identical, nattempts): maximal coupling of N(mu1, Sigma1) and N(mu2, Sigma2)."""
    return _squeeze(_eng().rmvnorm_max(mu1, mu2, Sigma1, Sigma2, R.Draws(seed=seed), n=n), n)


def rmvnorm_max_chol(mu1, mu2, Cholesky1, Cholesky2, Cholesky_inverse1, Cholesky_inverse2, n=1, seed=1):
    """R/mvnorm_couplings.R:57 — the same coupling from chol(Sigma) and solve(chol(Sigma)) of both Normals."""
    for Ch in (Cholesky1, Cholesky2):                                          # :59-61 stopifnot(valid_Cholesky)
        if np.any(np.tril(np.asarray(Ch, dtype=np.float64), -1) != 0):
            raise ValueError("valid_Cholesky1, valid_Cholesky2 are not all TRUE: pass chol(Sigma), not its transpose")
    return _squeeze(_eng().rmvnorm_max_chol(mu1, mu2, Cholesky1, Cholesky2, Cholesky_inverse1, Cholesky_inverse2,
                                            R.Draws(seed=seed), n=n), n)


def fast_rmvnorm(nparticles, mean, covariance, seed=1):
    """R/mvnorm.R:12"""
    return _eng().fast_rmvnorm(nparticles, mean, covariance, R.Draws(seed=seed))


def fast_dmvnorm(x, mean, covariance):
    """R/mvnorm.R:48"""
    return _eng().fast_dmvnorm(x, mean, covariance)


def rgamma_coupled(alpha1, alpha2, beta1, beta2, n=1, seed=1):
    """R/gamma_couplings.R:14-21"""
    return _eng().rgamma_coupled(alpha1, alpha2, beta1, beta2, R.Draws(seed=seed), n=n)


def rinversegamma_coupled(alpha1, alpha2, beta1, beta2, n=1, seed=1):
    """R/invgamma_couplings.R:29-36"""
    return _eng().rinversegamma_coupled(alpha1, alpha2, beta1, beta2, R.Draws(seed=seed), n=n)


def rinversegamma(n, alpha, beta, seed=1):
    """R/invgamma_couplings.R:17-19"""
    return _eng().rinversegamma(n, alpha, beta, R.Draws(seed=seed))


def dinversegamma(x, alpha, beta):
    """R/invgamma_couplings.R:7-9 (host arithmetic: a closed form, no device work)"""
    from math import lgamma
    x = np.asarray(x, dtype=np.float64)
    return alpha * np.log(beta) - lgamma(alpha) - (alpha + 1) * np.log(x) - beta / x


def rinvgaussian(n, mu, lam, seed=1):
    """R/invgaussian_couplings.R:27-29 / src/inversegaussian.cpp:6-25"""
    return _eng().rinvgaussian(n, mu, lam, R.Draws(seed=seed))


def dinvgaussian(x, mu, lam):
    """R/invgaussian_couplings.R:7-9"""
    x = np.asarray(x, dtype=np.float64)
    return 0.5 * np.log(lam / (2 * np.pi)) - 1.5 * np.log(x) - lam * (x - mu) ** 2 / (2 * mu ** 2 * x)


def rinvgaussian_coupled(mu1, mu2, lambda1, lambda2, n=1, seed=1):
    """R/invgaussian_couplings.R:36-38 / src/inversegaussian.cpp:33-63"""
    return _eng().rinvgaussian_coupled(mu1, mu2, lambda1, lambda2, R.Draws(seed=seed), n=n)


def fast_rmvnorm_chol(nparticles, mean, chol, seed=1):
    """R/mvnorm.R:32."""
    return _eng().fast_rmvnorm_chol(nparticles, mean, chol, R.Draws(seed=seed))


def fast_dmvnorm_chol_inverse(x, mean, chol_inverse):
    """R/mvnorm.R:69."""
    return _eng().fast_dmvnorm_chol_inverse(x, mean, chol_inverse)
