"""Thin Python host over the C ABI of libubm.so (include/ubm.h).

Everything here is argument marshalling: numpy arrays in, numpy arrays out, one C call per method.
All computation happens in the CUDA kernels behind the C ABI; there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _abi as A


class UbmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"ubm status {status}: {message}")
        self.status = status


def alloc_estimator_out(nrep, dimh, nbins, per_replicate=True, want_bins=True):
    res = {
        "uestimator": np.zeros((nrep, dimh)), "mcmcestimator": np.zeros((nrep, dimh)),
        "correction": np.zeros((nrep, dimh)), "meetingtime": np.zeros(nrep),
        "iteration": np.zeros(nrep, dtype=np.int64), "cost": np.zeros(nrep),
        "bins": np.zeros((nrep, nbins)) if (nbins and want_bins) else None,
        "summary": np.zeros(A.summary_len(dimh, nbins)),
    }
    if not per_replicate:
        for k in ("uestimator", "mcmcestimator", "correction", "meetingtime", "iteration", "cost", "bins"):
            res[k] = None
    out = A.ubm_estimator_out(A.dptr(res["uestimator"]), A.dptr(res["mcmcestimator"]), A.dptr(res["correction"]),
                              A.dptr(res["meetingtime"]), A.i64ptr(res["iteration"]), A.dptr(res["cost"]),
                              A.dptr(res["bins"]), A.dptr(res["summary"]))
    return res, out


class Engine:
    """One handle = one device + one CUDA stream (include/ubm.h: ubm_create)."""

    def __init__(self, device=0):
        self.lib = A.load_lib()
        self.h = C.c_void_p()
        st = self.lib.ubm_create(int(device), C.byref(self.h))
        if st != A.UBM_OK:
            raise UbmError(st, "ubm_create failed: no usable CUDA device (the engine has no CPU fallback)")
        self.device = int(device)

    def close(self):
        if self.h:
            self.lib.ubm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != A.UBM_OK:
            raise UbmError(st, self.lib.ubm_last_error(self.h).decode())

    def device_info(self):
        name = C.create_string_buffer(128)
        sm, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self.lib.ubm_device_info(self.h, name, 128, C.byref(sm), C.byref(maj), C.byref(mnr)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value)}

    def measure_peaks(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.ubm_measure_peaks(self.h, C.byref(a), C.byref(b)))
        t = C.c_double()
        self._check(self.lib.ubm_measure_fp64_tensor_peak(self.h, C.byref(t)))
        return {"fp64_tflops": a.value, "int32_tops": b.value, "fp64_tensor_tflops": t.value}

    def ising_swap_counts(self, nchains):
        """nswap_attempts, nswap_accepts1, nswap_accepts2 (R/ising.R:76,137-140) summed over the last Ising PT run"""
        att = C.c_int64()
        a1, a2 = np.zeros(max(nchains - 1, 1), dtype=np.int64), np.zeros(max(nchains - 1, 1), dtype=np.int64)
        self._check(self.lib.ubm_ising_swap_counts(self.h, C.byref(att), A.i64ptr(a1), A.i64ptr(a2)))
        return {"nswap_attempts": att.value, "nswap_accepts1": a1[:nchains - 1], "nswap_accepts2": a2[:nchains - 1]}

    def launch_count(self):
        return int(self.lib.ubm_launch_count(self.h))

    # ---- the three drivers -------------------------------------------------------------------------
    def sample_meetingtime(self, model, draws, lag=1, max_iterations=0, nrep=1, rep_offset=0):
        mt = np.empty(nrep, dtype=np.float64)
        ds = draws.struct(nrep)
        self._check(self.lib.ubm_sample_meetingtime(self.h, model.kind, model.ref(), C.byref(ds), lag, max_iterations,
                                                    nrep, rep_offset, A.dptr(mt)))
        return {"meetingtime": mt}

    def sample_unbiasedestimator(self, model, draws, h_kind=A.UBM_H_IDENTITY, bins=None, k=0, m=1, lag=1,
                                 max_iterations=0, nrep=1, rep_offset=0, per_replicate=True, want_bins=True):
        dimh = model.h_dim(h_kind)
        nbins = bins.nbins if bins is not None else 0
        res, out = alloc_estimator_out(nrep, dimh, nbins, per_replicate, want_bins)
        ds = draws.struct(nrep)
        bs = bins.struct() if bins is not None else None
        self._check(self.lib.ubm_sample_unbiasedestimator(
            self.h, model.kind, model.ref(), C.byref(ds), h_kind, C.byref(bs) if bs is not None else None, k, m, lag,
            max_iterations, nrep, rep_offset, C.byref(out)))
        return res

    def sample_coupled_chains(self, model, draws, m=1, lag=1, max_iterations=0, stride=None, nrep=1, rep_offset=0):
        d = model.dim
        if stride is None:
            raise ValueError("stride (rows of storage per replicate) is required")
        s1 = np.empty((nrep, stride, d))
        s2 = np.empty((nrep, stride, d))
        mt = np.zeros(nrep)
        it = np.zeros(nrep, dtype=np.int64)
        cost = np.zeros(nrep)
        ds = draws.struct(nrep)
        self._check(self.lib.ubm_sample_coupled_chains(self.h, model.kind, model.ref(), C.byref(ds), m, lag,
                                                       max_iterations, stride, nrep, rep_offset, A.dptr(s1),
                                                       A.dptr(s2), A.dptr(mt), A.i64ptr(it), A.dptr(cost)))
        return {"samples1": s1, "samples2": s2, "meetingtime": mt, "iteration": it, "cost": cost, "lag": lag}

    # ---- estimators from stored chains -------------------------------------------------------------
    def h_bar(self, chains, h_kind=A.UBM_H_IDENTITY, k=0, m=1):
        s1, s2 = chains["samples1"], chains["samples2"]
        nrep, stride, d = s1.shape
        dimh = 2 * d if h_kind == A.UBM_H_BOTH else d
        out = np.zeros((nrep, dimh))
        self._check(self.lib.ubm_h_bar(self.h, d, h_kind, k, m, chains["lag"], stride, nrep, A.dptr(s1), A.dptr(s2),
                                       A.dptr(chains["meetingtime"]), A.i64ptr(chains["iteration"]), A.dptr(out)))
        return out

    def estimator_bins(self, chains, bins, k=0, m=1):
        s1, s2 = chains["samples1"], chains["samples2"]
        nrep, stride, d = s1.shape
        out = np.zeros((nrep, bins.nbins))
        bs = bins.struct()
        self._check(self.lib.ubm_estimator_bins(self.h, d, C.byref(bs), k, m, chains["lag"], stride, nrep, A.dptr(s1),
                                                A.dptr(s2), A.dptr(chains["meetingtime"]), A.dptr(out)))
        return out

    def h_bar_grid(self, chains, ks, ms, h_kind=A.UBM_H_IDENTITY):
        """H_bar for every (k, m) pair from one upload of the stored chains: [npairs][nrep][dim h]."""
        s1, s2 = chains["samples1"], chains["samples2"]
        nrep, stride, d = s1.shape
        dimh = 2 * d if h_kind == A.UBM_H_BOTH else d
        ks = np.ascontiguousarray(np.atleast_1d(ks), dtype=np.int64)
        ms = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(ms), ks.shape), dtype=np.int64)
        out = np.zeros((len(ks), nrep, dimh))
        self._check(self.lib.ubm_h_bar_grid(self.h, d, h_kind, len(ks), A.i64ptr(ks), A.i64ptr(ms), chains["lag"], stride,
                                            nrep, A.dptr(s1), A.dptr(s2), A.dptr(chains["meetingtime"]),
                                            A.i64ptr(chains["iteration"]), A.dptr(out)))
        return out

    def tv_upper_bounds(self, meetingtimes, lag, t):
        """mean over the meeting times of max(0, ceil((tau - lag - t) / lag)) for every t of the grid."""
        mt = np.ascontiguousarray(meetingtimes, dtype=np.float64).reshape(-1)
        tg = np.ascontiguousarray(np.atleast_1d(t), dtype=np.int64)
        out = np.zeros(len(tg))
        self._check(self.lib.ubm_tv_upper_bounds(self.h, A.dptr(mt), len(mt), int(lag), A.i64ptr(tg), len(tg), A.dptr(out)))
        return out

    def c_chains_to_measure(self, chains, k=0, m=1):
        """src/c_chains_to_measure.cpp:38-84 for every replicate of a stored batch: list of dicts(atoms, weights, MCMC)."""
        s1, s2 = chains["samples1"], chains["samples2"]
        nrep, stride, d = s1.shape
        lag = chains["lag"]
        sizes = [self.lib.ubm_measure_size(k, m, lag, float(t)) for t in chains["meetingtime"]]
        cap = max(max(sizes), 1)
        atoms, weights = np.zeros((nrep, cap, d)), np.zeros((nrep, cap))
        part, sz = np.zeros((nrep, cap), dtype=np.int32), np.zeros(nrep, dtype=np.int64)
        self._check(self.lib.ubm_c_chains_to_measure(self.h, d, k, m, lag, stride, nrep, A.dptr(s1), A.dptr(s2),
                                                     A.dptr(chains["meetingtime"]), A.i64ptr(chains["iteration"]), cap,
                                                     A.dptr(atoms), A.dptr(weights), A.i32ptr(part), A.i64ptr(sz)))
        return [{"atoms": atoms[r, :sz[r]], "weights": weights[r, :sz[r]], "MCMC": part[r, :sz[r]].astype(bool)}
                for r in range(nrep)]

    # ---- stand-alone primitives of the R API, batched over n samples ----------------------------------
    def rnorm_max_coupling(self, mu1, mu2, sigma1, sigma2, draws, n=1, rep_offset=0):
        xy, ident = np.zeros((n, 2)), np.zeros(n, dtype=np.int32)
        ds = draws.struct(n)
        self._check(self.lib.ubm_rnorm_max_coupling(self.h, C.byref(ds), n, rep_offset, mu1, mu2, sigma1, sigma2,
                                                    A.dptr(xy), A.i32ptr(ident)))
        return {"xy": xy, "identical": ident.astype(bool)}

    # ---- the other scalar maximal couplings (R/gamma_couplings.R, R/invgamma_couplings.R, src/inversegaussian.cpp)
    def _positive_coupled(self, fn, p1, p2, q1, q2, draws, n, rep_offset):
        xy, ident = np.zeros((n, 2)), np.zeros(n, dtype=np.int32)
        ds = draws.struct(n)
        self._check(fn(self.h, C.byref(ds), n, rep_offset, p1, p2, q1, q2, A.dptr(xy), A.i32ptr(ident)))
        return {"xy": xy, "identical": ident.astype(bool)}

    def rgamma_coupled(self, alpha1, alpha2, beta1, beta2, draws, n=1, rep_offset=0):
        return self._positive_coupled(self.lib.ubm_rgamma_coupled, alpha1, alpha2, beta1, beta2, draws, n, rep_offset)

    def rinversegamma_coupled(self, alpha1, alpha2, beta1, beta2, draws, n=1, rep_offset=0):
        return self._positive_coupled(self.lib.ubm_rinversegamma_coupled, alpha1, alpha2, beta1, beta2, draws, n, rep_offset)

    def rinvgaussian_coupled(self, mu1, mu2, lambda1, lambda2, draws, n=1, rep_offset=0):
        return self._positive_coupled(self.lib.ubm_rinvgaussian_coupled, mu1, mu2, lambda1, lambda2, draws, n, rep_offset)

    def _rpositive(self, kind, a, b, draws, n, rep_offset):
        out = np.zeros(n)
        ds = draws.struct(n)
        self._check(self.lib.ubm_rpositive(self.h, C.byref(ds), kind, n, rep_offset, a, b, A.dptr(out)))
        return out

    def rgamma(self, n, shape, rate, draws, rep_offset=0):
        return self._rpositive(0, shape, rate, draws, n, rep_offset)

    def rinversegamma(self, n, alpha, beta, draws, rep_offset=0):
        return self._rpositive(1, alpha, beta, draws, n, rep_offset)

    def rinvgaussian(self, n, mu, lam, draws, rep_offset=0):
        return self._rpositive(2, mu, lam, draws, n, rep_offset)

    def rnorm_reflectionmax(self, mu1, mu2, sigma, draws, n=1, rep_offset=0):
        xy, ident = np.zeros((n, 2)), np.zeros(n, dtype=np.int32)
        ds = draws.struct(n)
        self._check(self.lib.ubm_rnorm_reflectionmax(self.h, C.byref(ds), n, rep_offset, mu1, mu2, sigma, A.dptr(xy),
                                                     A.i32ptr(ident)))
        return {"xy": xy, "identical": ident.astype(bool)}

    def rmvnorm_reflectionmax(self, mu1, mu2, Cholesky, Cholesky_inverse, draws, n=1, rep_offset=0):
        mu1, mu2 = np.ascontiguousarray(mu1, dtype=np.float64), np.ascontiguousarray(mu2, dtype=np.float64)
        d = len(mu1)
        ch = np.ascontiguousarray(np.asarray(Cholesky, dtype=np.float64).T).reshape(-1)       # column-major
        ci = np.ascontiguousarray(np.asarray(Cholesky_inverse, dtype=np.float64).T).reshape(-1)
        xy, ident = np.zeros((n, 2, d)), np.zeros(n, dtype=np.int32)
        ds = draws.struct(n)
        self._check(self.lib.ubm_rmvnorm_reflectionmax(self.h, C.byref(ds), n, rep_offset, d, A.dptr(mu1), A.dptr(mu2),
                                                       A.dptr(ch), A.dptr(ci), A.dptr(xy), A.i32ptr(ident)))
        return {"xy": xy, "identical": ident.astype(bool)}

    def rmvnorm_max(self, mu1, mu2, Sigma1, Sigma2, draws, n=1, rep_offset=0):
        """R/mvnorm_couplings.R:25 -> src/mvnorm.cpp:91-129."""
        mu1, mu2 = np.ascontiguousarray(mu1, dtype=np.float64), np.ascontiguousarray(mu2, dtype=np.float64)
        d = len(mu1)
        cm = lambda M: np.ascontiguousarray(np.asarray(M, dtype=np.float64).T).reshape(-1)       # column-major
        s1, s2 = cm(Sigma1), cm(Sigma2)
        xy, ident, na = np.zeros((n, 2, d)), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        ds = draws.struct(n)
        self._check(self.lib.ubm_rmvnorm_max(self.h, C.byref(ds), n, rep_offset, d, A.dptr(mu1), A.dptr(mu2), A.dptr(s1),
                                             A.dptr(s2), A.dptr(xy), A.i32ptr(ident), A.i32ptr(na)))
        return {"xy": xy, "identical": ident.astype(bool), "nattempts": na}

    def rmvnorm_max_chol(self, mu1, mu2, Cholesky1, Cholesky2, Cholesky_inverse1, Cholesky_inverse2, draws, n=1,
                         rep_offset=0):
        """R/mvnorm_couplings.R:57 -> src/mvnorm.cpp:133-174."""
        mu1, mu2 = np.ascontiguousarray(mu1, dtype=np.float64), np.ascontiguousarray(mu2, dtype=np.float64)
        d = len(mu1)
        cm = lambda M: np.ascontiguousarray(np.asarray(M, dtype=np.float64).T).reshape(-1)
        c1, c2, i1, i2 = cm(Cholesky1), cm(Cholesky2), cm(Cholesky_inverse1), cm(Cholesky_inverse2)
        xy, ident, na = np.zeros((n, 2, d)), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        ds = draws.struct(n)
        self._check(self.lib.ubm_rmvnorm_max_chol(self.h, C.byref(ds), n, rep_offset, d, A.dptr(mu1), A.dptr(mu2),
                                                  A.dptr(c1), A.dptr(c2), A.dptr(i1), A.dptr(i2), A.dptr(xy),
                                                  A.i32ptr(ident), A.i32ptr(na)))
        return {"xy": xy, "identical": ident.astype(bool), "nattempts": na}

    def fast_rmvnorm_chol(self, nparticles, mean, chol, draws, rep_offset=0):
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        d = len(mean)
        ch = np.ascontiguousarray(np.asarray(chol, dtype=np.float64).T).reshape(-1)
        out = np.zeros((nparticles, d))
        ds = draws.struct(nparticles)
        self._check(self.lib.ubm_fast_rmvnorm_chol(self.h, C.byref(ds), nparticles, rep_offset, d, A.dptr(mean),
                                                   A.dptr(ch), A.dptr(out)))
        return out

    def fast_rmvnorm(self, nparticles, mean, covariance, draws, rep_offset=0):
        """R/mvnorm.R:12 / src/mvnorm.cpp:9-26"""
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        d = len(mean)
        cov = np.ascontiguousarray(np.asarray(covariance, dtype=np.float64).T).reshape(-1)
        out = np.zeros((nparticles, d))
        ds = draws.struct(nparticles)
        self._check(self.lib.ubm_fast_rmvnorm(self.h, C.byref(ds), nparticles, rep_offset, d, A.dptr(mean), A.dptr(cov), A.dptr(out)))
        return out

    def fast_dmvnorm(self, x, mean, covariance):
        """R/mvnorm.R:48 / src/mvnorm.cpp:49-67"""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        n, d = x.shape
        cov = np.ascontiguousarray(np.asarray(covariance, dtype=np.float64).T).reshape(-1)
        out = np.zeros(n)
        self._check(self.lib.ubm_fast_dmvnorm(self.h, n, d, A.dptr(x), A.dptr(mean), A.dptr(cov), A.dptr(out)))
        return out

    def fast_dmvnorm_chol_inverse(self, x, mean, chol_inverse):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        n, d = x.shape
        ci = np.ascontiguousarray(np.asarray(chol_inverse, dtype=np.float64).T).reshape(-1)
        out = np.zeros(n)
        self._check(self.lib.ubm_fast_dmvnorm_chol_inverse(self.h, n, d, A.dptr(x), A.dptr(mean), A.dptr(ci), A.dptr(out)))
        return out

    # ---- device-resident plan (kernel-only timing, multi-GPU summaries) -----------------------------
    def plan(self, model, h_kind=A.UBM_H_IDENTITY, bins=None, k=0, m=1, lag=1, max_iterations=0, max_nrep=1):
        return Plan(self, model, h_kind, bins, k, m, lag, max_iterations, max_nrep)


class Plan:
    def __init__(self, eng, model, h_kind, bins, k, m, lag, max_iterations, max_nrep):
        self.eng, self.model, self.bins = eng, model, bins
        self.dimh = model.h_dim(h_kind)
        self.nbins = bins.nbins if bins is not None else 0
        self.max_nrep = max_nrep
        self.p = C.c_void_p()
        bs = bins.struct() if bins is not None else None
        eng._check(eng.lib.ubm_plan_create(eng.h, model.kind, model.ref(), h_kind,
                                           C.byref(bs) if bs is not None else None, k, m, lag, max_iterations,
                                           max_nrep, C.byref(self.p)))

    def launch(self, seed, nrep, rep_offset=0):
        self.eng._check(self.eng.lib.ubm_plan_launch(self.p, int(seed) & (2**64 - 1), nrep, rep_offset))
        self.last_nrep = nrep

    def fetch(self, per_replicate=False):
        res, out = alloc_estimator_out(self.last_nrep, self.dimh, self.nbins, per_replicate, want_bins=False)
        self.eng._check(self.eng.lib.ubm_plan_fetch(self.p, C.byref(out)))
        return res

    def last_ms(self):
        a, b = C.c_float(), C.c_float()
        self.eng._check(self.eng.lib.ubm_plan_last_ms(self.p, C.byref(a), C.byref(b)))
        return a.value, b.value

    def summary_dev(self):
        ptr, n = C.c_void_p(), C.c_int64()
        self.eng._check(self.eng.lib.ubm_plan_summary_dev(self.p, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def stream(self):
        s = C.c_void_p()
        self.eng._check(self.eng.lib.ubm_plan_stream(self.p, C.byref(s)))
        return s.value

    def close(self):
        if self.p:
            self.eng.lib.ubm_plan_destroy(self.p)
            self.p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
