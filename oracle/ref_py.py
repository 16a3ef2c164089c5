"""ctypes loader of oracle/_ref: the reference's OWN, UNMODIFIED C++ sources (src/*.cpp), compiled where they lie under
/root/reference against the stand-in R / Rcpp / Eigen headers of oracle/ref_shim (recipe: oracle/ref_shim/Makefile).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/ use it, to check the oracle restatement against the reference itself on
identical draws.  Two flavours: 'libm' (log / exp from libm, as an R build has them) and 'ubmnum' (the same sources with
log / exp redirected to include/ubm_numerics.h, so that continuous results can be compared bit for bit)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"
_libs = {}

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(ip)


def available():
    return os.path.exists(os.path.join(REF_DIR, "libubm_ref.so")) or os.path.isdir(os.path.join(REF_SRC, "src"))


def build():
    """Runs the committed recipe; needs the reference tree (present in the build container, absent on the GPU box, where
    the prebuilt files travel with the snapshot)."""
    if os.path.isdir(os.path.join(REF_SRC, "src")):
        subprocess.run(["make", "-C", os.path.join(HERE, "ref_shim"), f"REF={REF_SRC}"], check=True,
                       stdout=subprocess.DEVNULL)


def lib(flavour="libm"):
    if flavour not in _libs:
        name = "libubm_ref.so" if flavour == "libm" else "libubm_ref_ubmnum.so"
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path, mode=os.RTLD_NOW)
        L.ref_flavour.restype = C.c_char_p
        assert L.ref_flavour().decode() == flavour
        L.ref_tape_set.argtypes = [dp, C.c_int64, dp, C.c_int64, dp, C.c_int64]
        L.ref_tape_set.restype = None
        L.ref_tape_used.argtypes = [C.POINTER(C.c_int64)] * 3 + [ip]
        L.ref_tape_used.restype = None
        L.ref_use_rsession.argtypes = [C.c_void_p]
        L.ref_use_rsession.restype = None
        L.ref_ising_sum.argtypes = [ip, C.c_int32]
        L.ref_ising_sum.restype = C.c_int32
        L.ref_ising_gibbs_sweep.argtypes = [ip, C.c_int32, dp]
        L.ref_ising_gibbs_sweep.restype = None
        L.ref_ising_coupled_gibbs_sweep.argtypes = [ip, ip, C.c_int32, dp]
        L.ref_ising_coupled_gibbs_sweep.restype = None
        L.ref_estimator_bin.argtypes = [dp, C.c_int32, dp, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double,
                                        C.c_double, C.c_int32, C.c_int32, C.c_int32]
        L.ref_estimator_bin.restype = C.c_double
        L.ref_sample_pair01.argtypes = [dp, C.c_int32, ip]
        L.ref_sample_pair01.restype = None
        L.ref_c_chains_to_measure.argtypes = [dp, C.c_int32, dp, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                              C.c_int32, dp, dp, ip]
        L.ref_c_chains_to_measure.restype = C.c_int32
        L.ref_prune_measure.argtypes = [dp, C.c_int32, C.c_int32, dp]
        L.ref_prune_measure.restype = C.c_int32
        L.ref_logcosh.argtypes = [C.c_double]
        L.ref_logcosh.restype = C.c_double
        L.ref_xbeta.argtypes = [dp, C.c_int32, C.c_int32, dp, dp]
        L.ref_xbeta.restype = None
        L.ref_rpg_devroye.argtypes = [dp, C.c_int64, dp]
        L.ref_rpg_devroye.restype = None
        for f in ("ref_w_max_coupling", "ref_w_rejsampler"):
            getattr(L, f).argtypes = [dp, dp, dp, C.c_int32, C.c_int32, dp]
            getattr(L, f).restype = None
        L.ref_rinvgaussian.argtypes = [C.c_int32, C.c_double, C.c_double, dp]
        L.ref_rinvgaussian.restype = None
        L.ref_rinvgaussian_coupled.argtypes = [C.c_double] * 4 + [dp]
        L.ref_rinvgaussian_coupled.restype = None
        L.ref_fast_rmvnorm.argtypes = [C.c_int32, C.c_int32, dp, dp, dp]
        L.ref_fast_rmvnorm_cholesky.argtypes = [C.c_int32, C.c_int32, dp, dp, dp]
        L.ref_fast_dmvnorm.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, dp]
        L.ref_fast_dmvnorm_cholesky_inverse.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, dp]
        L.ref_rmvnorm_max_coupling.argtypes = [C.c_int32, dp, dp, dp, dp, dp, ip, ip]
        L.ref_rmvnorm_max_coupling_cholesky.argtypes = [C.c_int32, dp, dp, dp, dp, dp, dp, dp, ip]
        L.ref_rmvnorm_reflection_max_coupling.argtypes = [C.c_int32, dp, dp, dp, dp, dp, ip]
        L.ref_sigma.argtypes = [C.c_int32, C.c_int32, dp, dp, dp]
        L.ref_m_sigma_function.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, dp, dp, dp, dp, dp]
        for f in ("ref_fast_rmvnorm", "ref_fast_rmvnorm_cholesky", "ref_fast_dmvnorm", "ref_fast_dmvnorm_cholesky_inverse",
                  "ref_rmvnorm_max_coupling", "ref_rmvnorm_max_coupling_cholesky", "ref_rmvnorm_reflection_max_coupling",
                  "ref_sigma", "ref_m_sigma_function"):
            getattr(L, f).restype = None
        L.ref_blassoconditional.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, dp, dp, C.c_double, dp, dp]
        L.ref_blassoconditional.restype = None
        L.ref_marginal_likelihood_c_2.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, C.c_double, C.c_double]
        L.ref_marginal_likelihood_c_2.restype = C.c_double
        _libs[flavour] = L
    return _libs[flavour]


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def cm(a):
    """flat column-major copy of a 2-D array (R layout)"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1)


class Tape:
    """Installs typed replay tapes for the reference's unif_rand / norm_rand / exp_rand; `used()` = draws consumed."""

    def __init__(self, L, u=None, n=None, e=None):
        self.L = L
        self.keep = [None if t is None else f64(t).reshape(-1) for t in (u, n, e)]
        lens = [0 if t is None else t.size for t in self.keep]
        L.ref_tape_set(_d(self.keep[0]), lens[0], _d(self.keep[1]), lens[1], _d(self.keep[2]), lens[2])

    def used(self):
        a, b, c, ex = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        self.L.ref_tape_used(C.byref(a), C.byref(b), C.byref(c), C.byref(ex))
        return (a.value, b.value, c.value), bool(ex.value)
