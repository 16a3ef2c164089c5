// ubm_post.cuh — estimators from STORED coupled chains, batched over replicates:
//   hbar_kernel            R/H_bar.R:19-54
//   estimator_bin_kernel   src/estimator_bin.cpp:10-42
// Layout of the stored chains: samples1[r][t][dim], samples2[r][t][dim] (stride rows per replicate), as
// written by the MODE_CHAINS drivers.  One thread per (replicate, output); each thread walks time in the
// reference's order, so results equal the oracle's bit for bit.  These are the only HBM-streaming kernels
// on the path (algorithmic bytes: 8 (m-k+1 + 2 (tau-k-lag)) per output).
#pragma once
#include "ubm_common.cuh"

namespace ubm {

__device__ inline double h_component(int h_kind, const double* x, int dim, int i) {
  if (h_kind == UBM_H_IDENTITY) return x[i];
  if (h_kind == UBM_H_SQUARE) return x[i] * x[i];
  return (i < dim) ? x[i] : x[i - dim] * x[i - dim];
}

__global__ void hbar_kernel(const double* __restrict__ s1, const double* __restrict__ s2,
                            const double* __restrict__ mt, int dim, int h_kind, int64_t k, int64_t m, int64_t lag,
                            int64_t stride, int64_t nrep, double* __restrict__ out) {
  const int dimh = (h_kind == UBM_H_BOTH) ? 2 * dim : dim;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrep * dimh) return;
  int64_t r = idx / dimh;
  int i = (int)(idx - r * dimh);
  const double* a = s1 + r * stride * dim;
  const double* b = s2 + r * stride * dim;
  double acc = 0.0;
  for (int64_t t = k; t <= m; ++t) acc = acc + h_component(h_kind, a + t * dim, dim, i);       // :33-39
  double tau = mt[r];
  if (!(tau <= (double)(k + lag))) {                                                            // :42-51
    int64_t T = (int64_t)tau;
    for (int64_t t = k + lag; t <= T - 1; ++t) {
      double w = (double)corr_weight(t, k, m, lag);
      acc = acc + w * (h_component(h_kind, a + t * dim, dim, i) - h_component(h_kind, b + (t - lag) * dim, dim, i));
    }
  }
  out[idx] = acc / (double)(m - k + 1);                                                         // :53
}

__global__ void estimator_bin_kernel(const double* __restrict__ s1, const double* __restrict__ s2,
                                     const double* __restrict__ mt, int dim, int component,
                                     const double* __restrict__ breaks, int nbins, int64_t k, int64_t m, int64_t lag,
                                     int64_t stride, int64_t nrep, double* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrep * nbins) return;
  int64_t r = idx / nbins;
  int b = (int)(idx - r * nbins);
  const double lower = breaks[b], upper = breaks[b + 1];
  const double* a = s1 + r * stride * dim + component;
  const double* c = s2 + r * stride * dim + component;
  double estimator = 0;
  for (int64_t t = k; t <= m; ++t) {                                                            // :15-19
    double x = a[t * dim];
    if (x > lower && x < upper) estimator += 1;
  }
  double tau = mt[r];
  if (tau > (double)(k + lag)) {                                                                // :21
    int64_t T = (int64_t)tau;
    for (int64_t t = k + lag; t <= T - 1; ++t) {
      double coefficient = (double)corr_weight(t, k, m, lag);                                   // :26
      double increment = 0.;
      double x = a[t * dim], y = c[(t - lag) * dim];
      if (x > lower && x < upper) increment += coefficient;
      if (y > lower && y < upper) increment -= coefficient;
      estimator += increment;
    }
  }
  out[idx] = estimator / ((double)(m - k) + 1.);                                                // :41
}

// c_chains_to_measure_as_list_ (src/c_chains_to_measure.cpp:38-84): one thread per (replicate, atom row, coordinate);
// coordinate 0 also writes the weight and the MCMC flag.  HBM-bound copy: 8 dim bytes read and written per atom.
__global__ void measure_kernel(const double* __restrict__ s1, const double* __restrict__ s2, const double* __restrict__ mt,
                               int dim, int64_t k, int64_t m, int64_t lag, int64_t stride, int64_t nrep, int64_t cap,
                               double* __restrict__ atoms, double* __restrict__ weights, int* __restrict__ part) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrep * cap * dim) return;
  const int64_t row_all = idx / dim;
  const int c = (int)(idx - row_all * dim);
  const int64_t r = row_all / cap, row = row_all - r * cap;
  const int64_t nm = m - k + 1;
  const int64_t T = (int64_t)mt[r];
  const int64_t ncorr = (T - (k + lag) > 0) ? T - (k + lag) : 0;          // :45-49
  const double eq = 1. / (double)nm;                                      // :51
  const double* a = s1 + r * stride * dim;
  const double* b = s2 + r * stride * dim;
  double v = 0.0, w = 0.0;
  int flag = 0;
  if (row < nm) {                                                         // :62-66
    v = a[(k + row) * dim + c]; w = eq; flag = 1;
  } else if (row < nm + 2 * ncorr) {                                      // :70-82
    const int64_t q = row - nm, t = k + lag + (q >> 1);
    const double wt = (double)corr_weight(t, k, m, lag) * eq;             // :78-79
    if (q & 1) { v = b[(t - lag) * dim + c]; w = -wt; }
    else { v = a[t * dim + c]; w = wt; }
  }
  atoms[idx] = v;
  if (c == 0) { weights[row_all] = w; part[row_all] = flag; }
}

// tv_upper_bound of vignettes/polyagammagibbs.Rmd:238-240, averaged over the meeting times: one block per t
__global__ void tv_bound_kernel(const double* __restrict__ mt, int64_t nrep, int64_t lag, const int64_t* __restrict__ t,
                                double* __restrict__ out) {
  __shared__ long long part[256];
  const int64_t tj = t[blockIdx.x];
  long long acc = 0;
  for (int64_t r = threadIdx.x; r < nrep; r += blockDim.x) {
    const long long x = (long long)mt[r] - lag - tj;
    if (x > 0) acc += (x + lag - 1) / lag;                       // ceiling of a positive ratio of integers
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (double)part[0] / (double)nrep;
}

}  // namespace ubm
