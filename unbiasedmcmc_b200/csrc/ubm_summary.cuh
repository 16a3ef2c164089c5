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

}  // namespace ubm
