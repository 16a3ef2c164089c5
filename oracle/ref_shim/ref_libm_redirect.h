// ref_libm_redirect.h — TEST INFRASTRUCTURE.  Force-included (-include) when building the "ubmnum" flavour of oracle/_ref:
// the unmodified reference sources then take log / exp from the engine's numerics contract (include/ubm_numerics.h) instead
// of libm, exactly as the oracle restatement and the CUDA kernels do.  With the transcendentals equal, every remaining
// difference between the reference's code and the restatement would show as a bit difference.
#ifndef UBM_REF_LIBM_REDIRECT_H
#define UBM_REF_LIBM_REDIRECT_H
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <math.h>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <valarray>
#include <vector>
#include "../../include/ubm_numerics.h"
namespace ubm_redirect {
inline double r_log(double x) { return ubm_log(x); }
inline double r_exp(double x) { return ubm_exp(x); }
}  // namespace ubm_redirect
namespace std { using ubm_redirect::r_log; using ubm_redirect::r_exp; }
using ubm_redirect::r_log;
using ubm_redirect::r_exp;
#define log r_log
#define exp r_exp
#endif
