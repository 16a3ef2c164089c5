"""Device-side model registry, host view.

Each class holds the numpy arrays of one registered model and exposes `.kind` and `.struct()` — the
`ubm_model_*` struct of include/ubm.h.  These play the role of the closure triples
(single_kernel, coupled_kernel, rinit) that the reference's drivers take (R/coupled_chains.R:39):
R closures cannot run on a GPU, so kernels returned by `get_mh_kernels` & co. carry one of these specs.
"""
import ctypes as C

import numpy as np

from . import _abi as A


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _colmajor(a):
    """Flat column-major (R layout) copy of a 2-D array."""
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(a.T).reshape(-1)


class ModelSpec:
    kind = 0
    dim = 0

    def struct(self):
        raise NotImplementedError

    def ref(self):
        s = self.struct()
        self._keep = s
        return C.cast(C.pointer(s), C.c_void_p)

    def h_dim(self, h_kind):
        return 2 * self.dim if h_kind == A.UBM_H_BOTH else self.dim


class NormalMixtureRWMH(ModelSpec):
    """C1 — README.Rmd:69-84 target / rinit; R/get_mh_kernels.R:20-79 (reflection-max) or
    inst/reproducebimodal/bimodal.run.R:14-63 (max coupling)."""
    kind = A.UBM_MODEL_NORMAL_MIXTURE
    dim = 1

    def __init__(self, weights, means, sds, proposal_sd, init_mean, init_sd,
                 coupling=A.UBM_COUPLING_REFLECTION_MAX):
        self.weights, self.means, self.sds = _f64(weights), _f64(means), _f64(sds)
        if not (len(self.weights) == len(self.means) == len(self.sds)) or not 1 <= len(self.weights) <= 8:
            raise ValueError("mixture needs 1..8 components with matching weights/means/sds")
        self.proposal_sd, self.init_mean, self.init_sd = float(proposal_sd), float(init_mean), float(init_sd)
        self.coupling = int(coupling)

    def struct(self):
        return A.ubm_model_normal_mixture(len(self.weights), self.coupling, A.dptr(self.weights), A.dptr(self.means),
                                          A.dptr(self.sds), self.proposal_sd, self.init_mean, self.init_sd)


class MvnRWMH(ModelSpec):
    """C2 — inst/reproducescalingdimension/scaling.rwmh.reflectionmaxcoupling.R:12-103.
    Takes covariance matrices (like the R script) and factorises on the host with numpy, as R's
    chol()/solve(chol()) do before the kernels are built (R/get_mh_kernels.R:25-26)."""
    kind = A.UBM_MODEL_MVN

    def __init__(self, mean_pi, Sigma_pi, Sigma_proposal, init_mean, init_Sigma):
        self.mean_pi = _f64(mean_pi)
        d = self.dim = len(self.mean_pi)
        if not 2 <= d <= 128:
            raise ValueError("MVN model supports 2 <= d <= 128 (d = 1: use NormalMixtureRWMH, as R does)")
        up = lambda S: np.linalg.cholesky(np.asarray(S, dtype=np.float64)).T  # R's chol(): upper factor
        self.chol_pi = up(Sigma_pi)
        self.chol_inv_pi = np.triu(np.linalg.inv(self.chol_pi))
        self.chol_prop = up(Sigma_proposal)
        self.chol_inv_prop = np.triu(np.linalg.inv(self.chol_prop))
        self.init_mean = _f64(init_mean)
        self.init_chol = up(init_Sigma)
        self._flat = [_colmajor(np.triu(m)) for m in (self.chol_inv_pi, self.chol_prop, self.chol_inv_prop,
                                                      self.init_chol)]

    def struct(self):
        return A.ubm_model_mvn(self.dim, 0, A.dptr(self.mean_pi), A.dptr(self._flat[0]), A.dptr(self._flat[1]),
                               A.dptr(self._flat[2]), A.dptr(self.init_mean), A.dptr(self._flat[3]))


class Draws:
    """Draw source: Philox (seed) or replay tapes (arrays [nrep, len])."""

    def __init__(self, seed=None, tape_u=None, tape_n=None, tape_e=None):
        if seed is None and tape_u is None and tape_n is None and tape_e is None:
            raise ValueError("give a seed (Philox) or tapes (replay)")
        self.seed = seed
        self.tapes = [None if t is None else np.ascontiguousarray(np.asarray(t, dtype=np.float64))
                      for t in (tape_u, tape_n, tape_e)]

    @property
    def is_tape(self):
        return self.seed is None

    def struct(self, nrep=None):
        if not self.is_tape:
            return A.ubm_draws(A.UBM_DRAWS_PHILOX, 0, int(self.seed) & (2**64 - 1), A.dptr(None), A.dptr(None),
                               A.dptr(None), 0, 0, 0)
        lens = []
        for t in self.tapes:
            if t is None:
                lens.append(0)
            else:
                if t.ndim != 2 or (nrep is not None and t.shape[0] < nrep):
                    raise ValueError("tapes must be [nrep, len] with at least nrep rows")
                lens.append(t.shape[1])
        return A.ubm_draws(A.UBM_DRAWS_TAPE, 0, 0, A.dptr(self.tapes[0]), A.dptr(self.tapes[1]),
                           A.dptr(self.tapes[2]), lens[0], lens[1], lens[2])


class Bins:
    """estimator_bin_ over all bins of one component (src/estimator_bin.cpp:10-42); component 0-based."""

    def __init__(self, component, breaks):
        self.component = int(component)
        self.breaks = _f64(breaks)
        if len(self.breaks) < 2 or len(self.breaks) > 257 or np.any(np.diff(self.breaks) <= 0):
            raise ValueError("breaks must be 2..257 strictly increasing values")

    @property
    def nbins(self):
        return len(self.breaks) - 1

    def struct(self):
        return A.ubm_bins(self.component, len(self.breaks), A.dptr(self.breaks))


class IsingPT(ModelSpec):
    """C4 — 32x32 periodic Ising, parallel tempering (R/ising.R:30-141; run.batch.ising.R:169-173 shape).
    The chain state seen by h is the vector of natural statistics, one per inverse temperature."""
    kind = A.UBM_MODEL_ISING_PT

    def __init__(self, betas, proba_swapmove, size=32):
        self.betas = _f64(betas)
        if size != 32:
            raise ValueError("the device kernel is specialised for the reference's 32x32 lattice")
        if not 1 <= len(self.betas) <= 32:
            raise ValueError("1..32 temperatures")
        self.dim = len(self.betas)
        self.proba_swapmove = float(proba_swapmove)

    def struct(self):
        return A.ubm_model_ising_pt(32, self.dim, A.dptr(self.betas), self.proba_swapmove, 0, 0)


class PgLogistic(ModelSpec):
    """C3 — Polya-Gamma Gibbs sampler for Bayesian logistic regression, coupled as in
    inst/reproducegermancredit/germancredit.run.R:23-49.  Prior beta ~ N(b, B)."""
    kind = A.UBM_MODEL_PG_LOGISTIC

    def __init__(self, Y, X, b, B, mode="max", beta_coupling="max"):
        if mode not in ("max", "rej_samp"):                       # sample_w(mode=), R/logistic_regression.R:166-175
            raise ValueError("PG logistic: mode must be 'max' or 'rej_samp'")
        if beta_coupling not in ("max", "crn"):                   # sample_beta vs the vignette's common random numbers
            raise ValueError("PG logistic: beta_coupling must be 'max' or 'crn'")
        self.w_coupling = 0 if mode == "max" else 1
        self.beta_coupling = 0 if beta_coupling == "max" else 1
        X = np.asarray(X, dtype=np.float64)
        self.n, self.p = X.shape
        if not (1 <= self.p <= 60 and 1 <= self.n <= 4096):
            raise ValueError("PG logistic: n <= 4096, p <= 60")
        self.dim = self.p
        self.X = _colmajor(X)
        self.Y = _f64(Y)
        self.b = _f64(b).reshape(-1)
        self.B = _colmajor(B)

    def struct(self):
        return A.ubm_model_pg_logistic(self.n, self.p, A.dptr(self.X), A.dptr(self.Y), A.dptr(self.b), A.dptr(self.B),
                                       self.w_coupling, self.beta_coupling)


class VariableSelection(ModelSpec):
    """C5 — get_variableselection(Y, X, g, kappa, s0, proportion_singleflip) (R/variableselection.R:3)."""
    kind = A.UBM_MODEL_VARSEL

    def __init__(self, Y, X, g, kappa, s0, proportion_singleflip):
        X = np.asarray(X, dtype=np.float64)
        self.n, self.p = X.shape
        if not (2 <= self.p <= 1024 and 1 <= int(s0) <= 128):
            raise ValueError("variable selection: 2 <= p <= 1024, 1 <= s0 <= 128")
        self.dim = self.p
        self.X = _colmajor(X)
        self.Y = _f64(Y).reshape(-1)
        self.g, self.kappa, self.s0 = float(g), float(kappa), int(s0)
        self.proportion_singleflip = float(proportion_singleflip)

    def struct(self):
        return A.ubm_model_varsel(self.n, self.p, A.dptr(self.X), A.dptr(self.Y), self.g, self.kappa, self.s0, 0,
                                  self.proportion_singleflip)


class BayesianLasso(ModelSpec):
    """get_blasso(Y, X, lambda) (R/bayesianlasso.R:22): Gibbs sampler on (beta, tau2, sigma2) and its coupling
    (src/blassoutil.cpp:17-88, rinversegamma_coupled, rinvgaussian_coupled).  Chain state c(beta, tau2, sigma2)."""
    kind = A.UBM_MODEL_BLASSO

    def __init__(self, Y, X, lam):
        X = np.asarray(X, dtype=np.float64)
        self.n, self.p = X.shape
        if not (1 <= self.p <= 60 and 3 <= self.n <= 4096):
            raise ValueError("Bayesian lasso: 3 <= n <= 4096, p <= 60")
        self.dim = 2 * self.p + 1
        self.X = _colmajor(X)
        self.Y = _f64(Y).reshape(-1)
        self.lam = float(lam)
        if not self.lam > 0:
            raise ValueError("Bayesian lasso: lambda > 0")

    def struct(self):
        return A.ubm_model_blasso(self.n, self.p, A.dptr(self.X), A.dptr(self.Y), self.lam)
