// ubm_prims.cuh — the reference's stand-alone coupling / MVN primitives, batched: one THREAD per sample, sample i
// drawing from the typed streams of replicate index rep_offset + i (from draw 0).  These are the exported R
// functions of the path that are not drivers (NAMESPACE: rnorm_max_coupling, rnorm_reflectionmax,
// rmvnorm_reflectionmax, fast_rmvnorm_chol, fast_dmvnorm_chol_inverse); the driver kernels inline the same
// arithmetic.  Utility kernels: matrices are read from global memory, vectors live in local memory (d <= 128).
#pragma once
#include "ubm_common.cuh"
#include "ubm_mix.cuh"

namespace ubm {

#define PRIM_MAXD 128

// R/rnorm_max_coupling.R:17-23 over R/get_max_coupling.R:24-39 (mode 1) ; R/rnorm_reflectionmax.R:18-39 (mode 0)
template <bool TAPE>
__global__ void prim_scalar_coupling_kernel(const RunParams P, int mode, double mu1, double mu2, double s1, double s2,
                                            double ls1, double ls2, double* __restrict__ xy, int* __restrict__ ident) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nrep) return;
  ScalarDraws<TAPE> D;
  D.init(P, i);
  double x, y;
  bool id;
  if (mode == 0) {
    const double xdot = D.n();
    const double z = (mu1 - mu2) / s1;
    const double normz = sqrt(z * z);
    const double e = z / normz;
    const double utilde = D.u();
    const double a = xdot + z;
    id = ubm_log(utilde) < (-((UBM_LN_SQRT_2PI + 0.5 * a * a) + 0.0)) - (-((UBM_LN_SQRT_2PI + 0.5 * xdot * xdot) + 0.0));
    const double ydot = id ? (xdot + z) : (xdot - 2 * (e * xdot) * e);
    x = mu1 + s1 * xdot;
    y = mu2 + s1 * ydot;
  } else {
    x = mu1 + s1 * D.n();
    if (MixModel::dnorm_log(x, mu1, s1, ls1) + ubm_log(D.u()) < MixModel::dnorm_log(x, mu2, s2, ls2)) { id = true; y = x; }
    else {
      id = false;
      bool reject = true;
      y = 0.0;
      while (reject && !D.bad) {
        y = mu2 + s2 * D.n();
        reject = MixModel::dnorm_log(y, mu2, s2, ls2) + ubm_log(D.u()) < MixModel::dnorm_log(y, mu1, s1, ls1);
      }
    }
  }
  xy[2 * i] = x; xy[2 * i + 1] = y; ident[i] = id ? 1 : 0;
  if (D.bad) atomicExch(P.status, UBM_ERR_TAPE);
}

// ---- the other scalar maximal couplings: R/gamma_couplings.R:14-21, R/invgamma_couplings.R:7-36,
// src/inversegaussian.cpp:6-63, through R/get_max_coupling.R:24-39.  kind 0: Gamma(a, b) (rate b), 1: inverse Gamma,
// 2: inverse Gaussian (a = mu, b = lambda).  R's rgamma is part of the R runtime, not of the reference: Gamma draws are
// Marsaglia & Tsang (2000) on the typed streams (two draws per attempt; shape < 1 by the U^(1/shape) boost); lgamma
// enters through the host-computed constants c = a log(b) - lgamma(a).
template <class DS>
__device__ inline double prim_rgamma(double shape, double rate, DS& D) {
  const double a = shape < 1.0 ? shape + 1.0 : shape;
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double g = 0.0;
  for (;;) {
    const double x = D.n();
    const double t = 1.0 + c * x;
    const double u = D.u();
    if (D.bad) return 1.0;
    if (t <= 0.0) continue;
    const double v = t * t * t;
    if (ubm_log(u) < 0.5 * x * x + d - d * v + d * ubm_log(v)) { g = d * v; break; }
  }
  if (shape < 1.0) g = g * ubm_exp(ubm_log(D.u()) / shape);
  return g / rate;
}
template <class DS>
__device__ inline double prim_rinvgaussian(double mu, double lambda, DS& D) {   // src/inversegaussian.cpp:6-25
  const double nu = D.n();
  const double z = D.u();
  double x = nu * nu;
  x = mu + mu * mu * x / (2. * lambda) - mu / (2. * lambda) * sqrt(4. * mu * lambda * x + mu * mu * x * x);
  return (z <= mu / (mu + x)) ? x : mu * mu / x;
}
template <class DS>
__device__ inline double prim_positive_draw(int kind, double a, double b, DS& D, bool clamp) {
  if (kind == 0) return prim_rgamma(a, b, D);
  if (kind == 1) return 1.0 / prim_rgamma(a, b, D);
  const double v = prim_rinvgaussian(a, b, D);
  return (clamp && v < 1e-20) ? 1e-20 : v;                      // :40-42,52-54
}
__device__ inline double prim_positive_logpdf(int kind, double x, double a, double b, double c) {
  if (kind == 0) return c + (a - 1.0) * ubm_log(x) - b * x;
  if (kind == 1) return c - (a + 1.0) * ubm_log(x) - b / x;   // R/invgamma_couplings.R:7-9
  return 0.5 * ubm_log(b / 6.283185) - 1.5 * ubm_log(x) - b * (x - a) * (x - a) / (2 * a * a * x);   // src/inversegaussian.cpp:27-29
}
// R/get_max_coupling.R:24-39 for the three families; returns `identical`
template <class DS>
__device__ inline bool prim_positive_max_coupling(int kind, double a1, double a2, double b1, double b2, double c1, double c2, DS& D,
                                                  double* xo, double* yo) {
  const double x = prim_positive_draw(kind, a1, b1, D, true);
  *xo = x;
  if (prim_positive_logpdf(kind, x, a1, b1, c1) + ubm_log(D.u()) < prim_positive_logpdf(kind, x, a2, b2, c2)) { *yo = x; return true; }
  bool reject = true;
  double y = 0.0;
  while (reject && !D.bad) {
    y = prim_positive_draw(kind, a2, b2, D, true);
    reject = prim_positive_logpdf(kind, y, a2, b2, c2) + ubm_log(D.u()) < prim_positive_logpdf(kind, y, a1, b1, c1);
  }
  *yo = y;
  return false;
}
// mode 0: out [n] plain draws of the first distribution; mode 1: out [n][2] maximally coupled pairs, ident [n]
template <bool TAPE>
__global__ void prim_positive_coupling_kernel(const RunParams P, int kind, int mode, double a1, double a2, double b1, double b2,
                                              double c1, double c2, double* __restrict__ out, int* __restrict__ ident) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nrep) return;
  ScalarDraws<TAPE> D;
  D.init(P, i);
  if (mode == 0) {
    out[i] = prim_positive_draw(kind, a1, b1, D, false);
  } else {
    double x, y;
    const bool id = prim_positive_max_coupling(kind, a1, a2, b1, b2, c1, c2, D, &x, &y);
    out[2 * i] = x; out[2 * i + 1] = y; ident[i] = id ? 1 : 0;
  }
  if (D.bad) atomicExch(P.status, UBM_ERR_TAPE);
}

// y_j = sum_{i <= j} v_i U[i, j], U column-major upper triangular: one fma chain, sequential in i (the engine's order;
// see ubm_mvn.cuh / oracle_models.hpp: vec_times_upper)
__device__ inline void prim_vec_times_upper(const double* v, const double* __restrict__ U, int d, double* out) {
  for (int j = 0; j < d; ++j) {
    double acc = 0.0;
    for (int i = 0; i <= j; ++i) acc = ubm_fma(v[i], U[(size_t)j * d + i], acc);
    out[j] = acc;
  }
}
__device__ inline double prim_sum_g4(const double* a, const double* b, int d) {   // canonical reduction order
  double s[32];
  for (int q = 0; q < 32; ++q) s[q] = 0.0;
  for (int j = 0; j < d; ++j) s[j >> 2] = ubm_fma(a[j], b[j], s[j >> 2]);
  return ubm_butterfly32(s);
}

// mode 0: fast_rmvnorm_cholesky_ (src/mvnorm.cpp:30-46) -> out [n][d]
// mode 1: rmvnorm_reflection_max_coupling_ (src/mvnorm.cpp:185-236) -> out [n][2][d], ident [n]
template <bool TAPE>
__global__ void prim_mvn_sample_kernel(const RunParams P, int mode, int d, const double* __restrict__ mu1,
                                       const double* __restrict__ mu2, const double* __restrict__ chol,
                                       const double* __restrict__ chol_inv, double* __restrict__ out,
                                       int* __restrict__ ident) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nrep) return;
  ScalarDraws<TAPE> D;
  D.init(P, i);
  double xi[PRIM_MAXD], z[PRIM_MAXD], t[PRIM_MAXD];
  if (mode == 0) {
    for (int j = 0; j < d; ++j) xi[j] = D.n();
    prim_vec_times_upper(xi, chol, d, t);
    for (int j = 0; j < d; ++j) out[i * d + j] = t[j] + mu1[j];
  } else {
    for (int j = 0; j < d; ++j) t[j] = mu2[j] - mu1[j];
    prim_vec_times_upper(t, chol_inv, d, z);                               // :191
    for (int j = 0; j < d; ++j) xi[j] = D.n();                             // :193-195
    for (int j = 0; j < d; ++j) z[j] = -z[j];                              // :198
    const double normz = sqrt(prim_sum_g4(z, z, d));                       // :199
    double* X = out + (size_t)i * 2 * d;
    double* Y = X + d;
    bool id;
    if (normz < 1e-15) {                                                   // :200-210
      id = true;
      prim_vec_times_upper(xi, chol, d, t);
      for (int j = 0; j < d; ++j) { X[j] = mu1[j] + t[j]; Y[j] = X[j]; }
    } else {
      for (int j = 0; j < d; ++j) z[j] = z[j];
      double e[PRIM_MAXD];
      for (int j = 0; j < d; ++j) e[j] = z[j] / normz;                     // :212
      const double utilde = D.u();
      const double edx = prim_sum_g4(e, xi, d);                            // :216
      const double a = edx + normz;
      id = ubm_log(utilde) < (-0.5 * a * a + 0.5 * edx * edx);             // :217
      prim_vec_times_upper(xi, chol, d, t);                                // :224
      for (int j = 0; j < d; ++j) X[j] = mu1[j] + t[j];
      if (id) { for (int j = 0; j < d; ++j) Y[j] = X[j]; }                 // :228-229
      else {
        for (int j = 0; j < d; ++j) e[j] = xi[j] - 2. * edx * e[j];        // :221 (eta)
        prim_vec_times_upper(e, chol, d, t);                               // :225
        for (int j = 0; j < d; ++j) Y[j] = mu2[j] + t[j];
      }
    }
    ident[i] = id ? 1 : 0;
  }
  if (D.bad) atomicExch(P.status, UBM_ERR_TAPE);
}

// Maximal coupling of two multivariate Normals by rejection, one thread per sample.
//   variant 0: rmvnorm_max_coupling_cholesky (src/mvnorm.cpp:133-174): F?a = chol (upper), F?b = chol_inverse (upper);
//              draws through fast_rmvnorm_cholesky_ (:30-46), densities through fast_dmvnorm_cholesky_inverse_ (:71-86),
//              first test `<=` (:152), loop test `>` (:163)
//   variant 1: rmvnorm_max_coupling_ (src/mvnorm.cpp:91-129): F?a = lower Cholesky factor L of the covariance (the
//              LLT of fast_rmvnorm_ :12 / fast_dmvnorm_ :52, taken once on the host); y = z L' + mean, density by
//              forward substitution; first test `<` (:107), loop test `>` (:118)
// cst? = the density constants (sum of log-diagonals and -d * 0.9189385), computed on the host in the oracle's order.
template <bool TAPE>
__global__ void prim_mvn_max_kernel(const RunParams P, int variant, int d, const double* __restrict__ mu1,
                                    const double* __restrict__ mu2, const double* __restrict__ F1a,
                                    const double* __restrict__ F2a, const double* __restrict__ F1b,
                                    const double* __restrict__ F2b, double cst1, double cst2, double* __restrict__ xy,
                                    int* __restrict__ ident, int* __restrict__ natt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nrep) return;
  ScalarDraws<TAPE> D;
  D.init(P, i);
  double z[PRIM_MAXD], x[PRIM_MAXD], t[PRIM_MAXD];
  auto draw = [&](const double* mu, const double* Fa, double* out) {
    for (int j = 0; j < d; ++j) z[j] = D.n();
    if (variant == 0) {
      prim_vec_times_upper(z, Fa, d, t);
      for (int j = 0; j < d; ++j) out[j] = t[j] + mu[j];
    } else {
      for (int j = 0; j < d; ++j) {                                   // y_j = sum_{i <= j} z_i L(j, i)
        double a = 0.0;
        for (int q = 0; q <= j; ++q) a = ubm_fma(z[q], Fa[(size_t)q * d + j], a);
        out[j] = a + mu[j];
      }
    }
  };
  auto dens = [&](const double* pt, const double* mu, const double* Fa, const double* Fb, double cst) -> double {
    if (variant == 0) {
      for (int j = 0; j < d; ++j) z[j] = pt[j] - mu[j];
      prim_vec_times_upper(z, Fb, d, t);
      return -0.5 * prim_sum_g4(t, t, d) + cst;
    }
    for (int r = 0; r < d; ++r) {                                     // L y = pt - mu
      double a = pt[r] - mu[r];
      for (int k = 0; k < r; ++k) a = ubm_fma(-Fa[(size_t)k * d + r], t[k], a);
      t[r] = a / Fa[(size_t)r * d + r];
    }
    double sq = 0.0;
    for (int r = 0; r < d; ++r) sq = ubm_fma(t[r], t[r], sq);
    return -0.5 * sq + cst;
  };
  double* X = xy + (size_t)i * 2 * d;
  double* Y = X + d;
  draw(mu1, F1a, x);
  for (int j = 0; j < d; ++j) X[j] = x[j];
  const double logu = ubm_log(D.u());
  const double d11 = dens(x, mu1, F1a, F1b, cst1), d12 = dens(x, mu2, F2a, F2b, cst2);
  const bool same = (variant == 0) ? (d11 + logu <= d12) : (d11 + logu < d12);
  int n = 0;
  if (same) { for (int j = 0; j < d; ++j) Y[j] = x[j]; }
  else {
    bool accept = false;
    while (!accept && !D.bad) {
      ++n;
      draw(mu2, F2a, x);
      const double lu = ubm_log(D.u());
      const double d22 = dens(x, mu2, F2a, F2b, cst2), d21 = dens(x, mu1, F1a, F1b, cst1);
      if (d22 + lu > d21) accept = true;
    }
    for (int j = 0; j < d; ++j) Y[j] = x[j];
  }
  ident[i] = same ? 1 : 0;
  natt[i] = n;
  if (D.bad) atomicExch(P.status, UBM_ERR_TAPE);
}

// fast_dmvnorm_cholesky_inverse_ (src/mvnorm.cpp:71-86): x [n][d] -> out [n]
__global__ void prim_dmvnorm_kernel(int64_t n, int d, const double* __restrict__ x, const double* __restrict__ mean,
                                    const double* __restrict__ chol_inv, double cst, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xc[PRIM_MAXD], w[PRIM_MAXD];
  for (int j = 0; j < d; ++j) xc[j] = x[i * d + j] - mean[j];
  prim_vec_times_upper(xc, chol_inv, d, w);
  out[i] = -0.5 * prim_sum_g4(w, w, d) + cst;
}

}  // namespace ubm
