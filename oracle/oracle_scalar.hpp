// oracle_scalar.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header for the parity status).
// CPU restatement of the remaining scalar maximal couplings of the reference:
//   R/gamma_couplings.R:14-21 (rgamma_coupled), R/invgamma_couplings.R:7-36 (dinversegamma, rinversegamma,
//   rinversegamma_coupled), R/invgaussian_couplings.R + src/inversegaussian.cpp:6-63 (rinvgaussian_c, dinvgaussian_c,
//   rinvgaussian_coupled_c), all through R/get_max_coupling.R:24-39.
// The inverse-Gaussian sampler is the reference's own C++ and is reproduced draw for draw (pinned against the compiled
// reference, tests/test_ref_parity.py).  R's rgamma lives in the R runtime, not in the reference tree: Gamma draws here
// are Marsaglia & Tsang (2000) on the typed streams (shape < 1 by the U^(1/shape) boost) — the coupling logic and the
// densities follow the reference, the marginal sampler is this engine's contract; lgamma enters through host-computed
// constants shared with the device.
#pragma once
#include <cmath>
#include "oracle_core.hpp"

namespace orc {

inline double rgamma_mt(double shape, double rate, Draws& D) {
  const double a = shape < 1.0 ? shape + 1.0 : shape;
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double g = 0.0;
  for (;;) {
    const double x = D.n();
    const double t = 1.0 + c * x;
    const double u = D.u();                                    // always drawn: a fixed two draws per attempt
    if (D.exhausted) return 1.0;
    if (t <= 0.0) continue;
    const double v = t * t * t;
    if (ubm_log(u) < 0.5 * x * x + d - d * v + d * ubm_log(v)) { g = d * v; break; }
  }
  if (shape < 1.0) g = g * ubm_exp(ubm_log(D.u()) / shape);
  return g / rate;
}
// log densities with the normalising constant cst = alpha log(beta) - lgamma(alpha) passed in
inline double dgamma_log(double x, double alpha, double beta, double cst) { return cst + (alpha - 1.0) * ubm_log(x) - beta * x; }
inline double dinvgamma_log(double x, double alpha, double beta, double cst) { return cst - (alpha + 1.0) * ubm_log(x) - beta / x; }   // R/invgamma_couplings.R:7-9

// src/inversegaussian.cpp:6-25 with n = 1: one normal, then one uniform
inline double rinvgaussian1(double mu, double lambda, Draws& D) {
  const double nu = D.n();
  const double z = D.u();
  double x = nu * nu;
  x = mu + mu * mu * x / (2. * lambda) - mu / (2. * lambda) * sqrt(4. * mu * lambda * x + mu * mu * x * x);
  return (z <= mu / (mu + x)) ? x : mu * mu / x;
}
inline double dinvgaussian_log(double x, double mu, double lambda) {   // :27-29 (the reference's 6.283185)
  return 0.5 * ubm_log(lambda / 6.283185) - 1.5 * ubm_log(x) - lambda * (x - mu) * (x - mu) / (2 * mu * mu * x);
}

// kind 0: Gamma(a, b) rate parametrisation; 1: inverse Gamma; 2: inverse Gaussian (a = mu, b = lambda)
inline bool positive_max_coupling(int kind, double a1, double a2, double b1, double b2, double c1, double c2, Draws& D,
                                  double* xo, double* yo) {
  auto rp = [&](double a, double b) {
    if (kind == 0) return rgamma_mt(a, b, D);
    if (kind == 1) return 1.0 / rgamma_mt(a, b, D);
    double v = rinvgaussian1(a, b, D);
    return v < 1e-20 ? 1e-20 : v;                              // src/inversegaussian.cpp:40-42,52-54
  };
  auto dl = [&](double x, double a, double b, double c) {
    if (kind == 0) return dgamma_log(x, a, b, c);
    if (kind == 1) return dinvgamma_log(x, a, b, c);
    return dinvgaussian_log(x, a, b);
  };
  const double x = rp(a1, b1);
  *xo = x;
  if (dl(x, a1, b1, c1) + ubm_log(D.u()) < dl(x, a2, b2, c2)) { *yo = x; return true; }      // R/get_max_coupling.R:27-29
  double y = 0.0;
  bool reject = true;
  while (reject && !D.exhausted) {                                                            // :31-35
    y = rp(a2, b2);
    reject = dl(y, a2, b2, c2) + ubm_log(D.u()) < dl(y, a1, b1, c1);
  }
  *yo = y;
  return false;
}

}  // namespace orc
