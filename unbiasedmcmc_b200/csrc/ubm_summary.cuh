// ubm_summary.cuh — deterministic aggregation of per-replicate results into the flat summary buffer
// (include/ubm.h: ubm_summary_len).  One block of 1024 threads per summary column: thread q adds the
// replicates r = q, q+1024, ... in ascending order, then the 1024 partials are combined by an xor-butterfly
// (512, 256, ..., 1).  The order is fixed, so the sums are reproducible and match the oracle's sum1024.
#pragma once
#include "ubm_common.cuh"

namespace ubm {

struct SummaryArgs {
  int64_t nrep;
  int dimh, nbins;
  const double *tau, *cost, *uest;
  const int64_t* iter;
  const long long* bin_acc;
  double denom;
  double* summary;
};

__global__ void __launch_bounds__(1024) summary_kernel(const SummaryArgs A) {
  __shared__ double sa[1024], sb[1024];
  const int col = blockIdx.x, q = threadIdx.x;
  const int nfloat_cols = UBM_SUMMARY_HEAD + 2 * A.dimh;
  if (col >= nfloat_cols) {   // bins: exact integers, one rounding
    if (q == 0) {
      int b = col - nfloat_cols;   // 0 .. 2 nbins
      double c = (double)A.bin_acc[b];
      A.summary[col] = (b < A.nbins) ? c / A.denom : c / (A.denom * A.denom);
    }
    return;
  }
  double s = 0.0;
  for (int64_t r = q; r < A.nrep; r += 1024) {
    double t = A.tau[r];
    bool ok = isfinite(t);
    double v;
    bool sq = false;
    if (col == 0) v = ok ? 1.0 : 0.0;
    else if (col == 1) v = ok ? 0.0 : 1.0;
    else if (col == 2) v = ok ? t : 0.0;
    else if (col == 3) { v = ok ? t : 0.0; sq = true; }
    else if (col == 4) v = ok ? A.cost[r] : 0.0;
    else if (col == 5) v = ok ? (double)A.iter[r] : 0.0;
    else {
      int c = col - UBM_SUMMARY_HEAD;
      if (c >= A.dimh) { c -= A.dimh; sq = true; }
      v = (ok && A.uest) ? A.uest[r * A.dimh + c] : 0.0;
    }
    s = sq ? ubm_fma(v, v, s) : (s + v);
  }
  sa[q] = s;
  __syncthreads();
  double* cur = sa;
  double* nxt = sb;
  for (int off = 512; off >= 1; off >>= 1) {
    nxt[q] = cur[q] + cur[q ^ off];
    __syncthreads();
    double* t = cur; cur = nxt; nxt = t;
  }
  if (q == 0) A.summary[col] = cur[0];
}

// estimator_bin_ values of every replicate and their sums over the uncensored replicates, from the integer numerators the
// driver kernels accumulated with global reductions (P.bin_num).  One block per bin; integer sums: order-independent.
__global__ void __launch_bounds__(256) bins_finalize_kernel(int64_t nrep, int nbins, const double* __restrict__ tau,
                                                            const long long* __restrict__ bin_num, double denom,
                                                            double* __restrict__ bins_out, long long* __restrict__ bin_acc) {
  __shared__ long long s1[256], s2[256];
  const int b = blockIdx.x, q = threadIdx.x;
  long long c1 = 0, c2 = 0;
  for (int64_t r = q; r < nrep; r += 256) {
    const long long c = bin_num[r * nbins + b];
    if (bins_out) bins_out[r * nbins + b] = (double)c / denom;
    if (isfinite(tau[r])) { c1 += c; c2 += c * c; }
  }
  s1[q] = c1; s2[q] = c2;
  __syncthreads();
  for (int off = 128; off >= 1; off >>= 1) {
    if (q < off) { s1[q] += s1[q + off]; s2[q] += s2[q + off]; }
    __syncthreads();
  }
  if (q == 0) { bin_acc[b] = s1[0]; bin_acc[nbins + b] = s2[0]; }
}

}  // namespace ubm
