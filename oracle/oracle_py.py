"""ctypes loader of the CPU oracle (oracle/_build/libubm_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
(cpu_baseline, --impl reference) may import this module.  Parity status: see oracle_core.hpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from unbiasedmcmc_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libubm_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.orc_sample_meetingtime.argtypes = [C.c_int32, vp, C.POINTER(A.ubm_draws), C.c_int64, C.c_int64, C.c_int64,
                                             C.c_int64, A.c_double_p, C.c_int32]
        L.orc_sample_unbiasedestimator.argtypes = [C.c_int32, vp, C.POINTER(A.ubm_draws), C.c_int32,
                                                   C.POINTER(A.ubm_bins), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                                   C.c_int64, C.c_int64, C.POINTER(A.ubm_estimator_out), C.c_int32]
        L.orc_sample_coupled_chains.argtypes = [C.c_int32, vp, C.POINTER(A.ubm_draws), C.c_int64, C.c_int64,
                                                C.c_int64, C.c_int64, C.c_int64, C.c_int64, A.c_double_p,
                                                A.c_double_p, A.c_double_p, A.c_int64_p, A.c_double_p, C.c_int32]
        L.orc_h_bar.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                A.c_double_p, A.c_double_p, A.c_double_p, A.c_int64_p, A.c_double_p]
        L.orc_estimator_bins.argtypes = [C.c_int32, C.POINTER(A.ubm_bins), C.c_int64, C.c_int64, C.c_int64,
                                         C.c_int64, C.c_int64, A.c_double_p, A.c_double_p, A.c_double_p,
                                         A.c_double_p]
        L.orc_tv_upper_bounds.argtypes = [A.c_double_p, C.c_int64, C.c_int64, A.c_int64_p, C.c_int32, A.c_double_p]
        L.orc_c_chains_to_measure.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                              A.c_double_p, A.c_double_p, A.c_double_p, A.c_int64_p, C.c_int64,
                                              A.c_double_p, A.c_double_p, A.c_int32_p, A.c_int64_p]
        L.orc_pg_draws.argtypes = [C.c_uint64, C.c_double, C.c_int64, A.c_double_p]
        L.orc_pg_draws.restype = None
        for f in ("orc_vec_log", "orc_vec_exp", "orc_vec_log1p", "orc_vec_qnorm", "orc_vec_pnorm_log"):
            getattr(L, f).argtypes = [A.c_double_p, A.c_double_p, C.c_int64]
            getattr(L, f).restype = None
        L.orc_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.orc_philox.restype = None
        L.orc_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, A.c_double_p]
        L.orc_stream.restype = None
        L.orc_rsession_create.restype = C.c_void_p
        L.orc_rsession_create.argtypes = [C.c_uint32]
        L.orc_rsession_destroy.argtypes = [C.c_void_p]
        L.orc_rsession_destroy.restype = None
        L.orc_rsession_draw.argtypes = [C.c_void_p, C.c_int32, C.c_int64, A.c_double_p]
        L.orc_rsession_draw.restype = None
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def _check(st):
    if st != 0:
        raise OracleError(st)


def vec(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    getattr(lib(), "orc_vec_" + fn)(A.dptr(x), A.dptr(y), x.size)
    return y


def pg_draws(seed, z, n):
    out = np.empty(n, dtype=np.float64)
    lib().orc_pg_draws(int(seed), float(z), n, A.dptr(out))
    return out


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(*[int(c) for c in ctr], *[int(k) for k in key], out)
    return [int(v) for v in out]


def stream(seed, rep, type_, n):
    out = np.empty(n, dtype=np.float64)
    lib().orc_stream(int(seed), int(rep), int(type_), n, A.dptr(out))
    return out


class RSession:
    """One emulated R session stream (oracle/r_stream.hpp): `RSession(1)` is `set.seed(1)`.  Passed as `draws`, the
    oracle's drivers run consecutive replicates on this single sequential stream, exactly as an R loop would (Mersenne-
    Twister, inversion normals).  With `record=(len_u, len_n, len_e)` the typed draws each replicate consumed are
    written to `tapes` ([nrep, len] arrays, NaN beyond the consumed prefix) for replay through UBM_DRAWS_TAPE."""
    ORC_DRAWS_RSTREAM = 2

    def __init__(self, seed, record=None):
        self._s = lib().orc_rsession_create(int(seed))
        self.record = record
        self.tapes = None

    def close(self):
        if self._s:
            lib().orc_rsession_destroy(self._s)
            self._s = None

    def _draw(self, type_, n):
        out = np.empty(n, dtype=np.float64)
        lib().orc_rsession_draw(self._s, type_, n, A.dptr(out))
        return out

    def runif(self, n):
        return self._draw(0, n)

    def rnorm(self, n):
        return self._draw(1, n)

    def rexp(self, n):
        return self._draw(2, n)

    def struct(self, nrep=None):
        if self.record is None:
            return A.ubm_draws(self.ORC_DRAWS_RSTREAM, 0, self._s, A.dptr(None), A.dptr(None), A.dptr(None), 0, 0, 0)
        self.tapes = [np.full((nrep, max(int(n), 1)), np.nan) for n in self.record]
        return A.ubm_draws(self.ORC_DRAWS_RSTREAM, 0, self._s, A.dptr(self.tapes[0]), A.dptr(self.tapes[1]),
                           A.dptr(self.tapes[2]), self.tapes[0].shape[1], self.tapes[1].shape[1], self.tapes[2].shape[1])


def sample_meetingtime(model, draws, lag=1, max_iterations=0, nrep=1, rep_offset=0, nthreads=1):
    mt = np.empty(nrep, dtype=np.float64)
    ds = draws.struct(nrep)
    _check(lib().orc_sample_meetingtime(model.kind, model.ref(), C.byref(ds), lag, max_iterations, nrep, rep_offset,
                                        A.dptr(mt), nthreads))
    return {"meetingtime": mt}


def alloc_estimator_out(nrep, dimh, nbins, per_replicate=True):
    res = {
        "uestimator": np.zeros((nrep, dimh)), "mcmcestimator": np.zeros((nrep, dimh)),
        "correction": np.zeros((nrep, dimh)), "meetingtime": np.zeros(nrep),
        "iteration": np.zeros(nrep, dtype=np.int64), "cost": np.zeros(nrep),
        "bins": np.zeros((nrep, nbins)) if nbins else None,
        "summary": np.zeros(A.summary_len(dimh, nbins)),
    }
    if not per_replicate:
        for k in ("uestimator", "mcmcestimator", "correction", "meetingtime", "iteration", "cost", "bins"):
            res[k] = None
    out = A.ubm_estimator_out(A.dptr(res["uestimator"]), A.dptr(res["mcmcestimator"]), A.dptr(res["correction"]),
                              A.dptr(res["meetingtime"]), A.i64ptr(res["iteration"]), A.dptr(res["cost"]),
                              A.dptr(res["bins"]), A.dptr(res["summary"]))
    return res, out


def sample_unbiasedestimator(model, draws, h_kind=A.UBM_H_IDENTITY, bins=None, k=0, m=1, lag=1, max_iterations=0,
                             nrep=1, rep_offset=0, nthreads=1):
    dimh = model.h_dim(h_kind)
    nbins = bins.nbins if bins is not None else 0
    res, out = alloc_estimator_out(nrep, dimh, nbins)
    ds = draws.struct(nrep)
    bs = bins.struct() if bins is not None else None
    _check(lib().orc_sample_unbiasedestimator(model.kind, model.ref(), C.byref(ds), h_kind,
                                              C.byref(bs) if bs is not None else None, k, m, lag, max_iterations,
                                              nrep, rep_offset, C.byref(out), nthreads))
    return res


def sample_coupled_chains(model, draws, m=1, lag=1, max_iterations=0, stride=None, nrep=1, rep_offset=0, nthreads=1):
    d = model.dim
    if stride is None:
        raise ValueError("stride (rows of storage per replicate) is required")
    s1 = np.full((nrep, stride, d), np.nan)
    s2 = np.full((nrep, stride, d), np.nan)
    mt = np.zeros(nrep)
    it = np.zeros(nrep, dtype=np.int64)
    cost = np.zeros(nrep)
    ds = draws.struct(nrep)
    _check(lib().orc_sample_coupled_chains(model.kind, model.ref(), C.byref(ds), m, lag, max_iterations, stride, nrep,
                                           rep_offset, A.dptr(s1), A.dptr(s2), A.dptr(mt), A.i64ptr(it), A.dptr(cost),
                                           nthreads))
    return {"samples1": s1, "samples2": s2, "meetingtime": mt, "iteration": it, "cost": cost, "lag": lag}


def tv_upper_bounds(meetingtimes, lag, t):
    mt = np.ascontiguousarray(meetingtimes, dtype=np.float64).reshape(-1)
    tg = np.ascontiguousarray(np.atleast_1d(t), dtype=np.int64)
    out = np.zeros(len(tg))
    _check(lib().orc_tv_upper_bounds(A.dptr(mt), len(mt), int(lag), A.i64ptr(tg), len(tg), A.dptr(out)))
    return out


def c_chains_to_measure(chains, k=0, m=1):
    s1, s2 = chains["samples1"], chains["samples2"]
    nrep, stride, d = s1.shape
    lag = chains["lag"]
    tau = chains["meetingtime"]
    sizes = [(m - k + 1) + 2 * max(0, int(t) - (k + lag)) for t in tau]
    cap = max(max(sizes), 1)
    atoms, weights = np.zeros((nrep, cap, d)), np.zeros((nrep, cap))
    part, sz = np.zeros((nrep, cap), dtype=np.int32), np.zeros(nrep, dtype=np.int64)
    _check(lib().orc_c_chains_to_measure(d, k, m, lag, stride, nrep, A.dptr(s1), A.dptr(s2), A.dptr(tau),
                                         A.i64ptr(chains["iteration"]), cap, A.dptr(atoms), A.dptr(weights),
                                         A.i32ptr(part), A.i64ptr(sz)))
    return [{"atoms": atoms[r, :sz[r]], "weights": weights[r, :sz[r]], "MCMC": part[r, :sz[r]].astype(bool)}
            for r in range(nrep)]


def h_bar(chains, h_kind=A.UBM_H_IDENTITY, k=0, m=1):
    s1, s2 = chains["samples1"], chains["samples2"]
    nrep, stride, d = s1.shape
    dimh = 2 * d if h_kind == A.UBM_H_BOTH else d
    out = np.zeros((nrep, dimh))
    _check(lib().orc_h_bar(d, h_kind, k, m, chains["lag"], stride, nrep, A.dptr(s1), A.dptr(s2),
                           A.dptr(chains["meetingtime"]), A.i64ptr(chains["iteration"]), A.dptr(out)))
    return out


def estimator_bins(chains, bins, k=0, m=1):
    s1, s2 = chains["samples1"], chains["samples2"]
    nrep, stride, d = s1.shape
    out = np.zeros((nrep, bins.nbins))
    bs = bins.struct()
    _check(lib().orc_estimator_bins(d, C.byref(bs), k, m, chains["lag"], stride, nrep, A.dptr(s1), A.dptr(s2),
                                    A.dptr(chains["meetingtime"]), A.dptr(out)))
    return out


# ---- stand-alone primitives (oracle side) ---------------------------------------------------------------------
def _prim_lib():
    L = lib()
    if not getattr(L, "_prims_bound", False):
        dp = C.POINTER(A.ubm_draws)
        L.orc_rnorm_coupling.argtypes = [C.c_int32, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                         C.c_double, A.c_double_p, A.c_int32_p]
        L.orc_mvn_sample.argtypes = [C.c_int32, dp, C.c_int64, C.c_int64, C.c_int32, A.c_double_p, A.c_double_p,
                                     A.c_double_p, A.c_double_p, A.c_double_p, A.c_int32_p]
        L.orc_dmvnorm_chol_inverse.argtypes = [C.c_int64, C.c_int32, A.c_double_p, A.c_double_p, A.c_double_p,
                                               A.c_double_p]
        L._prims_bound = True
    return L


def _cm(m):
    return np.ascontiguousarray(np.asarray(m, dtype=np.float64).T).reshape(-1)


def rnorm_max_coupling(mu1, mu2, sigma1, sigma2, draws, n=1, rep_offset=0):
    xy, ident = np.zeros((n, 2)), np.zeros(n, dtype=np.int32)
    ds = draws.struct(n)
    _check(_prim_lib().orc_rnorm_coupling(1, C.byref(ds), n, rep_offset, mu1, mu2, sigma1, sigma2, A.dptr(xy), A.i32ptr(ident)))
    return {"xy": xy, "identical": ident.astype(bool)}


def blasso_conditional(model, draws, tau2, sigma2):
    L = lib()
    L.orc_blasso_conditional.argtypes = [C.c_void_p, C.POINTER(A.ubm_draws), A.c_double_p, C.c_double, A.c_double_p, A.c_double_p]
    L.orc_blasso_conditional.restype = C.c_int
    beta, nb = np.zeros(model.p), np.zeros(2)
    ds = draws.struct(1)
    tau2 = np.ascontiguousarray(tau2, dtype=np.float64)
    _check(L.orc_blasso_conditional(model.ref(), C.byref(ds), A.dptr(tau2), sigma2, A.dptr(beta), A.dptr(nb)))
    return beta, nb[0], nb[1]


def ising_swap_counts(nchains, reset=True):
    out = np.zeros(63, dtype=np.int64)
    L = lib()
    L.orc_ising_swap_counts.argtypes = [C.c_int32, C.POINTER(C.c_int64)]
    L.orc_ising_swap_counts.restype = None
    L.orc_ising_swap_counts(1 if reset else 0, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return {"nswap_attempts": int(out[0]), "nswap_accepts1": out[1:nchains], "nswap_accepts2": out[32:31 + nchains]}


def _positive(kind, mode, a1, a2, b1, b2, draws, n, rep_offset):
    L = _prim_lib()
    L.orc_positive_coupling.argtypes = [C.c_int32, C.c_int32, C.POINTER(A.ubm_draws), C.c_int64, C.c_int64,
                                        C.c_double, C.c_double, C.c_double, C.c_double, A.c_double_p, A.c_int32_p]
    L.orc_positive_coupling.restype = C.c_int
    out = np.zeros((n, 2)) if mode == 1 else np.zeros(n)
    ident = np.zeros(n, dtype=np.int32)
    ds = draws.struct(n)
    _check(L.orc_positive_coupling(kind, mode, C.byref(ds), n, rep_offset, a1, a2, b1, b2, A.dptr(out), A.i32ptr(ident)))
    return {"xy": out, "identical": ident.astype(bool)} if mode == 1 else out


def rgamma_coupled(alpha1, alpha2, beta1, beta2, draws, n=1, rep_offset=0):
    return _positive(0, 1, alpha1, alpha2, beta1, beta2, draws, n, rep_offset)


def rinversegamma_coupled(alpha1, alpha2, beta1, beta2, draws, n=1, rep_offset=0):
    return _positive(1, 1, alpha1, alpha2, beta1, beta2, draws, n, rep_offset)


def rinvgaussian_coupled(mu1, mu2, lambda1, lambda2, draws, n=1, rep_offset=0):
    return _positive(2, 1, mu1, mu2, lambda1, lambda2, draws, n, rep_offset)


def rgamma(n, shape, rate, draws, rep_offset=0):
    return _positive(0, 0, shape, shape, rate, rate, draws, n, rep_offset)


def rinversegamma(n, alpha, beta, draws, rep_offset=0):
    return _positive(1, 0, alpha, alpha, beta, beta, draws, n, rep_offset)


def rinvgaussian(n, mu, lam, draws, rep_offset=0):
    return _positive(2, 0, mu, mu, lam, lam, draws, n, rep_offset)


def rnorm_reflectionmax(mu1, mu2, sigma, draws, n=1, rep_offset=0):
    xy, ident = np.zeros((n, 2)), np.zeros(n, dtype=np.int32)
    ds = draws.struct(n)
    _check(_prim_lib().orc_rnorm_coupling(0, C.byref(ds), n, rep_offset, mu1, mu2, sigma, sigma, A.dptr(xy), A.i32ptr(ident)))
    return {"xy": xy, "identical": ident.astype(bool)}


def rmvnorm_reflectionmax(mu1, mu2, Cholesky, Cholesky_inverse, draws, n=1, rep_offset=0):
    mu1, mu2 = np.ascontiguousarray(mu1, dtype=np.float64), np.ascontiguousarray(mu2, dtype=np.float64)
    d = len(mu1)
    ch, ci = _cm(Cholesky), _cm(Cholesky_inverse)
    xy, ident = np.zeros((n, 2, d)), np.zeros(n, dtype=np.int32)
    ds = draws.struct(n)
    _check(_prim_lib().orc_mvn_sample(1, C.byref(ds), n, rep_offset, d, A.dptr(mu1), A.dptr(mu2), A.dptr(ch), A.dptr(ci),
                                      A.dptr(xy), A.i32ptr(ident)))
    return {"xy": xy, "identical": ident.astype(bool)}


def rmvnorm_max(mu1, mu2, Sigma1, Sigma2, draws, n=1, rep_offset=0):
    mu1, mu2 = np.ascontiguousarray(mu1, dtype=np.float64), np.ascontiguousarray(mu2, dtype=np.float64)
    d = len(mu1)
    s1, s2 = _cm(Sigma1), _cm(Sigma2)
    xy, ident, na = np.zeros((n, 2, d)), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    ds = draws.struct(n)
    _check(_prim_lib().orc_mvn_max(1, C.byref(ds), n, rep_offset, d, A.dptr(mu1), A.dptr(mu2), A.dptr(s1), A.dptr(s2),
                                   None, None, A.dptr(xy), A.i32ptr(ident), A.i32ptr(na)))
    return {"xy": xy, "identical": ident.astype(bool), "nattempts": na}


def rmvnorm_max_chol(mu1, mu2, Cholesky1, Cholesky2, Cholesky_inverse1, Cholesky_inverse2, draws, n=1, rep_offset=0):
    mu1, mu2 = np.ascontiguousarray(mu1, dtype=np.float64), np.ascontiguousarray(mu2, dtype=np.float64)
    d = len(mu1)
    c1, c2, i1, i2 = _cm(Cholesky1), _cm(Cholesky2), _cm(Cholesky_inverse1), _cm(Cholesky_inverse2)
    xy, ident, na = np.zeros((n, 2, d)), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    ds = draws.struct(n)
    _check(_prim_lib().orc_mvn_max(0, C.byref(ds), n, rep_offset, d, A.dptr(mu1), A.dptr(mu2), A.dptr(c1), A.dptr(c2),
                                   A.dptr(i1), A.dptr(i2), A.dptr(xy), A.i32ptr(ident), A.i32ptr(na)))
    return {"xy": xy, "identical": ident.astype(bool), "nattempts": na}


def fast_rmvnorm_chol(nparticles, mean, chol, draws, rep_offset=0):
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    d = len(mean)
    ch = _cm(chol)
    out = np.zeros((nparticles, d))
    ds = draws.struct(nparticles)
    _check(_prim_lib().orc_mvn_sample(0, C.byref(ds), nparticles, rep_offset, d, A.dptr(mean), A.dptr(None), A.dptr(ch),
                                      A.dptr(None), A.dptr(out), A.i32ptr(None)))
    return out


def _cov_factor(which, covariance):
    cov = np.asarray(covariance, dtype=np.float64)
    d = cov.shape[0]
    flat = _cm(cov)
    out = np.zeros(d * d)
    L = lib()
    L.orc_cov_factor.argtypes = [C.c_int32, C.c_int32, A.c_double_p, A.c_double_p]
    L.orc_cov_factor.restype = C.c_int
    _check(L.orc_cov_factor(which, d, A.dptr(flat), A.dptr(out)))
    return out.reshape(d, d).T           # column-major -> the matrix


def fast_rmvnorm(nparticles, mean, covariance, draws, rep_offset=0):
    return fast_rmvnorm_chol(nparticles, mean, _cov_factor(0, covariance), draws, rep_offset)


def fast_dmvnorm(x, mean, covariance):
    return fast_dmvnorm_chol_inverse(x, mean, _cov_factor(1, covariance))


def fast_dmvnorm_chol_inverse(x, mean, chol_inverse):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    n, d = x.shape
    ci = _cm(chol_inverse)
    out = np.zeros(n)
    _check(_prim_lib().orc_dmvnorm_chol_inverse(n, d, A.dptr(x), A.dptr(mean), A.dptr(ci), A.dptr(out)))
    return out


# ---- probes of single reference functions (compared with oracle/_ref in tests/test_ref_parity.py) -----------------
def _probe_lib():
    L = lib()
    if not getattr(L, "_probes_bound", False):
        vp, dpp = C.c_void_p, C.POINTER(A.ubm_draws)
        L.orc_ising_sum.argtypes = [A.c_int32_p]
        L.orc_ising_sum.restype = C.c_int32
        L.orc_ising_sweep.argtypes = [C.c_double, A.c_int32_p, A.c_int32_p, A.c_double_p, C.c_int64, A.c_double_p]
        L.orc_ising_sweep.restype = None
        L.orc_sample_pair01.argtypes = [vp, A.c_double_p, C.c_double, C.c_double, A.c_int32_p]
        L.orc_sample_pair01.restype = None
        L.orc_varsel_ml.argtypes = [vp, A.c_double_p]
        L.orc_varsel_ml.restype = C.c_double
        L.orc_pg_rpg_seq.argtypes = [dpp, A.c_double_p, C.c_int64, A.c_double_p, A.c_int64_p]
        L.orc_pg_w_seq.argtypes = [vp, dpp, A.c_double_p, A.c_double_p, A.c_double_p, A.c_int64_p]
        L.orc_pg_logcosh.argtypes = [C.c_double]
        L.orc_pg_logcosh.restype = C.c_double
        L.orc_pg_xbeta.argtypes = [vp, A.c_double_p, A.c_double_p]
        L.orc_pg_xbeta.restype = None
        L.orc_pg_m_sigma.argtypes = [vp] + [A.c_double_p] * 6
        L.orc_pg_m_sigma.restype = None
        L._probes_bound = True
    return L


def ising_sum(state):
    st = np.ascontiguousarray(state, dtype=np.int32)
    return int(_probe_lib().orc_ising_sum(A.i32ptr(st)))


def ising_sweep(beta, state1, u, state2=None):
    """state*: flat column-major 32 x 32 int32 lattices (modified copies are returned)"""
    s1 = np.array(state1, dtype=np.int32)
    s2 = None if state2 is None else np.array(state2, dtype=np.int32)
    u = np.ascontiguousarray(u, dtype=np.float64)
    probas = np.zeros(5)
    _probe_lib().orc_ising_sweep(float(beta), A.i32ptr(s1), A.i32ptr(s2), A.dptr(u), u.size, A.dptr(probas))
    return s1, s2, probas


def sample_pair01(model, selection, u0, u1):
    sel = np.ascontiguousarray(selection, dtype=np.float64)
    out = np.zeros(2, dtype=np.int32)
    _probe_lib().orc_sample_pair01(model.ref(), A.dptr(sel), float(u0), float(u1), A.i32ptr(out))
    return out


def varsel_ml(model, selection):
    sel = np.ascontiguousarray(selection, dtype=np.float64)
    return float(_probe_lib().orc_varsel_ml(model.ref(), A.dptr(sel)))


def pg_rpg_seq(draws, z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    out, used = np.zeros(z.size), np.zeros(3, dtype=np.int64)
    ds = draws.struct(1)
    _check(_probe_lib().orc_pg_rpg_seq(C.byref(ds), A.dptr(z), z.size, A.dptr(out), A.i64ptr(used)))
    return out, used


def pg_w_seq(model, draws, beta1, beta2):
    b1, b2 = np.ascontiguousarray(beta1, dtype=np.float64), np.ascontiguousarray(beta2, dtype=np.float64)
    w, used = np.zeros((model.n, 2)), np.zeros(3, dtype=np.int64)
    ds = draws.struct(1)
    _check(_probe_lib().orc_pg_w_seq(model.ref(), C.byref(ds), A.dptr(b1), A.dptr(b2), A.dptr(w), A.i64ptr(used)))
    return w, used


def pg_logcosh(x):
    return float(_probe_lib().orc_pg_logcosh(float(x)))


def pg_xbeta(model, beta):
    b = np.ascontiguousarray(beta, dtype=np.float64)
    out = np.zeros(model.n)
    _probe_lib().orc_pg_xbeta(model.ref(), A.dptr(b), A.dptr(out))
    return out


def pg_m_sigma(model, omega):
    om = np.ascontiguousarray(omega, dtype=np.float64)
    p = model.p
    m, L, Li, iB, ktk = np.zeros(p), np.zeros(p * p), np.zeros(p * p), np.zeros(p * p), np.zeros(p)
    _probe_lib().orc_pg_m_sigma(model.ref(), A.dptr(om), A.dptr(m), A.dptr(L), A.dptr(Li), A.dptr(iB), A.dptr(ktk))
    cmj = lambda a: a.reshape(p, p).T.copy()
    return {"m": m, "Cholesky_inverse": cmj(L), "Cholesky": cmj(Li), "invB": cmj(iB), "ktk": ktk}
