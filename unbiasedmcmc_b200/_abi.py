"""ctypes mirror of include/ubm.h (struct layouts, constants) and the loader of libubm.so.

The shared library is the product: if it is missing or cannot be loaded, importing the engine fails
loudly — there is no CPU or PyTorch fallback (north_star: "there is no CPU fallback").
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UBM_LIB_PATH") or os.path.join(HERE, "_lib", "libubm.so")   # override: debug builds only

UBM_OK, UBM_ERR_INVALID, UBM_ERR_CUDA, UBM_ERR_UNSUPPORTED, UBM_ERR_TAPE, UBM_ERR_NOMEM = range(6)
UBM_DRAWS_PHILOX, UBM_DRAWS_TAPE = 0, 1
UBM_COUPLING_REFLECTION_MAX, UBM_COUPLING_MAX = 0, 1
UBM_H_IDENTITY, UBM_H_SQUARE, UBM_H_BOTH = 0, 1, 2
UBM_MODEL_NORMAL_MIXTURE, UBM_MODEL_MVN, UBM_MODEL_PG_LOGISTIC, UBM_MODEL_ISING_PT, UBM_MODEL_VARSEL = 1, 2, 3, 4, 5
UBM_MODEL_BLASSO = 6
UBM_SUMMARY_HEAD = 6

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)


def dptr(a):
    """double* of a C-contiguous float64 numpy array (or NULL for None)."""
    if a is None:
        return C.cast(None, c_double_p)
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def i64ptr(a):
    if a is None:
        return C.cast(None, c_int64_p)
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int64_p)


def i32ptr(a):
    if a is None:
        return C.cast(None, c_int32_p)
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


class ubm_draws(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64),
                ("tape_u", c_double_p), ("tape_n", c_double_p), ("tape_e", c_double_p),
                ("len_u", C.c_int64), ("len_n", C.c_int64), ("len_e", C.c_int64)]


class ubm_model_normal_mixture(C.Structure):
    _fields_ = [("ncomp", C.c_int32), ("coupling", C.c_int32),
                ("weights", c_double_p), ("means", c_double_p), ("sds", c_double_p),
                ("proposal_sd", C.c_double), ("init_mean", C.c_double), ("init_sd", C.c_double)]


class ubm_model_mvn(C.Structure):
    _fields_ = [("d", C.c_int32), ("reserved", C.c_int32),
                ("mean_pi", c_double_p), ("chol_inv_pi", c_double_p), ("chol_prop", c_double_p),
                ("chol_inv_prop", c_double_p), ("init_mean", c_double_p), ("init_chol", c_double_p)]


class ubm_model_ising_pt(C.Structure):
    _fields_ = [("size", C.c_int32), ("nchains", C.c_int32), ("betas", c_double_p),
                ("proba_swapmove", C.c_double), ("scan", C.c_int32), ("reserved", C.c_int32)]


class ubm_model_pg_logistic(C.Structure):
    _fields_ = [("n", C.c_int32), ("p", C.c_int32), ("X", c_double_p), ("Y", c_double_p),
                ("b", c_double_p), ("B", c_double_p), ("w_coupling", C.c_int32), ("beta_coupling", C.c_int32)]


class ubm_model_blasso(C.Structure):
    _fields_ = [("n", C.c_int32), ("p", C.c_int32), ("X", c_double_p), ("Y", c_double_p), ("lambda_", C.c_double)]


class ubm_model_varsel(C.Structure):
    _fields_ = [("n", C.c_int32), ("p", C.c_int32), ("X", c_double_p), ("Y", c_double_p),
                ("g", C.c_double), ("kappa", C.c_double), ("s0", C.c_int32), ("reserved", C.c_int32),
                ("proportion_singleflip", C.c_double)]


class ubm_bins(C.Structure):
    _fields_ = [("component", C.c_int32), ("nbreaks", C.c_int32), ("breaks", c_double_p)]


class ubm_estimator_out(C.Structure):
    _fields_ = [("uestimator", c_double_p), ("mcmcestimator", c_double_p), ("correction", c_double_p),
                ("meetingtime", c_double_p), ("iteration", c_int64_p), ("cost", c_double_p),
                ("bins", c_double_p), ("summary", c_double_p)]


def summary_len(dimh, nbins):
    return UBM_SUMMARY_HEAD + 2 * dimh + 2 * nbins


_lib = None


def load_lib():
    """Load libubm.so (built in-tree by `make -C unbiasedmcmc_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UBM_LIB_PATH", LIB_PATH)      # debug builds (per-phase cycle counters) live beside the product library
    if not os.path.exists(path):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the engine has no CPU fallback)")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.ubm_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.ubm_destroy.argtypes = [vp]
    lib.ubm_last_error.argtypes = [vp]
    lib.ubm_last_error.restype = C.c_char_p
    lib.ubm_device_info.argtypes = [vp, C.c_char_p, C.c_int, c_int32_p, c_int32_p, c_int32_p]
    lib.ubm_launch_count.argtypes = [vp]
    lib.ubm_launch_count.restype = C.c_int64
    lib.ubm_measure_peaks.argtypes = [vp, c_double_p, c_double_p]
    lib.ubm_measure_peaks.restype = C.c_int
    lib.ubm_fast_rmvnorm.argtypes = [vp, C.POINTER(ubm_draws), C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p]
    lib.ubm_fast_rmvnorm.restype = C.c_int
    lib.ubm_fast_dmvnorm.argtypes = [vp, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.ubm_fast_dmvnorm.restype = C.c_int
    lib.ubm_ising_swap_counts.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.ubm_ising_swap_counts.restype = C.c_int
    lib.ubm_measure_fp64_tensor_peak.argtypes = [vp, c_double_p]
    lib.ubm_measure_fp64_tensor_peak.restype = C.c_int
    lib.ubm_summary_len.argtypes = [C.c_int32, C.c_int32]
    lib.ubm_summary_len.restype = C.c_int64
    lib.ubm_model_dim.argtypes = [C.c_int32, vp, c_int32_p]
    lib.ubm_h_dim.argtypes = [C.c_int32, vp, C.c_int32, c_int32_p]
    lib.ubm_sample_meetingtime.argtypes = [vp, C.c_int32, vp, C.POINTER(ubm_draws), C.c_int64, C.c_int64,
                                           C.c_int64, C.c_int64, c_double_p]
    lib.ubm_sample_unbiasedestimator.argtypes = [vp, C.c_int32, vp, C.POINTER(ubm_draws), C.c_int32,
                                                 C.POINTER(ubm_bins), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                                 C.c_int64, C.c_int64, C.POINTER(ubm_estimator_out)]
    lib.ubm_sample_coupled_chains.argtypes = [vp, C.c_int32, vp, C.POINTER(ubm_draws), C.c_int64, C.c_int64,
                                              C.c_int64, C.c_int64, C.c_int64, C.c_int64, c_double_p, c_double_p,
                                              c_double_p, c_int64_p, c_double_p]
    lib.ubm_h_bar.argtypes = [vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                              c_double_p, c_double_p, c_double_p, c_int64_p, c_double_p]
    lib.ubm_estimator_bins.argtypes = [vp, C.c_int32, C.POINTER(ubm_bins), C.c_int64, C.c_int64, C.c_int64,
                                       C.c_int64, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.ubm_h_bar_grid.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, c_int64_p, c_int64_p, C.c_int64, C.c_int64,
                                   C.c_int64, c_double_p, c_double_p, c_double_p, c_int64_p, c_double_p]
    lib.ubm_h_bar_grid.restype = C.c_int
    lib.ubm_tv_upper_bounds.argtypes = [vp, c_double_p, C.c_int64, C.c_int64, c_int64_p, C.c_int32, c_double_p]
    lib.ubm_tv_upper_bounds.restype = C.c_int
    lib.ubm_measure_size.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_double]
    lib.ubm_measure_size.restype = C.c_int64
    lib.ubm_c_chains_to_measure.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                            c_double_p, c_double_p, c_double_p, c_int64_p, C.c_int64, c_double_p,
                                            c_double_p, c_int32_p, c_int64_p]
    lib.ubm_c_chains_to_measure.restype = C.c_int
    dp = C.POINTER(ubm_draws)
    for name in ("ubm_rgamma_coupled", "ubm_rinversegamma_coupled", "ubm_rinvgaussian_coupled"):
        getattr(lib, name).argtypes = [vp, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                       c_double_p, C.POINTER(C.c_int32)]
        getattr(lib, name).restype = C.c_int
    lib.ubm_rpositive.argtypes = [vp, dp, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_double, c_double_p]
    lib.ubm_rpositive.restype = C.c_int
    lib.ubm_rnorm_max_coupling.argtypes = [vp, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                           C.c_double, c_double_p, c_int32_p]
    lib.ubm_rnorm_reflectionmax.argtypes = [vp, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                            c_double_p, c_int32_p]
    lib.ubm_rmvnorm_reflectionmax.argtypes = [vp, dp, C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p,
                                              c_double_p, c_double_p, c_double_p, c_int32_p]
    lib.ubm_rmvnorm_max.argtypes = [vp, dp, C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p,
                                    c_double_p, c_double_p, c_int32_p, c_int32_p]
    lib.ubm_rmvnorm_max_chol.argtypes = [vp, dp, C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p,
                                         c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p]
    lib.ubm_fast_rmvnorm_chol.argtypes = [vp, dp, C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p]
    lib.ubm_fast_dmvnorm_chol_inverse.argtypes = [vp, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p,
                                                  c_double_p]
    for name in ("ubm_rnorm_max_coupling", "ubm_rnorm_reflectionmax", "ubm_rmvnorm_reflectionmax",
                 "ubm_rmvnorm_max", "ubm_rmvnorm_max_chol",
                 "ubm_fast_rmvnorm_chol", "ubm_fast_dmvnorm_chol_inverse"):
        getattr(lib, name).restype = C.c_int
    lib.ubm_plan_create.argtypes = [vp, C.c_int32, vp, C.c_int32, C.POINTER(ubm_bins), C.c_int64, C.c_int64,
                                    C.c_int64, C.c_int64, C.c_int64, C.POINTER(vp)]
    lib.ubm_plan_launch.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64]
    lib.ubm_plan_fetch.argtypes = [vp, C.POINTER(ubm_estimator_out)]
    lib.ubm_plan_summary_dev.argtypes = [vp, C.POINTER(vp), c_int64_p]
    lib.ubm_plan_stream.argtypes = [vp, C.POINTER(vp)]
    lib.ubm_plan_last_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.ubm_plan_destroy.argtypes = [vp]
    for name in ("ubm_create", "ubm_destroy", "ubm_device_info", "ubm_model_dim", "ubm_h_dim",
                 "ubm_sample_meetingtime", "ubm_sample_unbiasedestimator", "ubm_sample_coupled_chains",
                 "ubm_h_bar", "ubm_estimator_bins", "ubm_plan_create", "ubm_plan_launch", "ubm_plan_fetch",
                 "ubm_plan_summary_dev", "ubm_plan_stream", "ubm_plan_last_ms", "ubm_plan_destroy"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
