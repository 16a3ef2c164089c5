// ubm_common.cuh — shared device-side pieces of libubm.so: run parameters, draw sources, estimator weights.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ubm.h"
#include "../../include/ubm_numerics.h"

namespace ubm {

enum RunMode : int { MODE_MEETING = 0, MODE_ESTIMATOR = 1, MODE_CHAINS = 2 };

// Everything a driver kernel needs besides the model itself.  Passed by value as a kernel argument.
struct RunParams {
  int mode;
  int h_kind;
  int64_t k, m, lag, max_it;   // max_it <= 0: unbounded
  int64_t nrep, rep_offset;
  // draws
  int draws_mode;
  uint64_t seed;
  const double *tape_u, *tape_n, *tape_e;
  int64_t len_u, len_n, len_e;
  // bins
  int component, nbreaks;
  const double* breaks;          // device
  // per-replicate outputs (device; may be null)
  double *uest, *mcmc, *corr;    // [nrep][dimh]
  double *tau, *cost;            // [nrep]
  int64_t* iter;                 // [nrep]
  double* bins_out;              // [nrep][nbins]
  long long* bin_acc;            // [2][nbins] integer sums over uncensored replicates: numerators, squares
  long long* bin_num;            // [nrep][nbins] per-replicate integer numerators (block-per-replicate kernels: global
                                 // reductions, turned into bins_out / bin_acc by bins_finalize_kernel); null for C1
  // chains mode
  double *samples1, *samples2;   // [nrep][stride][dim]
  int64_t stride;
  // control
  unsigned long long* work_counter;   // replicate hand-out
  int* status;                        // sticky: UBM_ERR_TAPE
  long long* pt_counts;               // Ising PT: [0] swap-move iterations, [1 .. N-1] accepted swaps of chain 1, [N .. 2N-2] of chain 2
  // C2 estimator runs in two passes (ubm_mvn.cuh: coupled phase, ubm_mvn_single.cuh: single-chain continuation).  The hand-off
  // record of a replicate that has met but not reached its horizon lives in its own, not yet final, output rows: state in
  // uest[rep], mcmc / correction numerators in mcmc[rep] / corr[rep], log-density in cost[rep], time in iter[rep], tau in
  // tau[rep]; plus the two draw counters below.  Null pointers = one pass.
  long long* ho_int;                  // [nrep][2]  next uniform, next normal
  int* ho_list;                       // [nrep]     replicates to continue, in hand-off order
  unsigned long long* ho_count;       // number of entries of ho_list
};

// weight of Delta_t in H_{k:m}: floor((t-k)/lag) - ceiling(max(lag, t-m)/lag) + 1
// (R/unbiasedestimator.R:71,94; R/H_bar.R:47; src/estimator_bin.cpp:26), t >= k + lag.
__host__ __device__ inline int64_t corr_weight(int64_t t, int64_t k, int64_t m, int64_t lag) {
  int64_t fl = (t - k) / lag;
  int64_t a = (t - m > lag) ? (t - m) : lag;
  int64_t ce = (a + lag - 1) / lag;
  return fl - ce + 1;
}

// bin b with breaks[b] < x < breaks[b+1] (strict, src/estimator_bin.cpp:16) or -1.
// `inv_w` / `b0` give a first guess for (near-)uniform breaks; the fix-up loops make it exact for any
// increasing breaks.
__device__ inline int bin_of(double x, const double* __restrict__ breaks, int nbreaks, double b0, double inv_w) {
  if (!(x > breaks[0]) || !(x < breaks[nbreaks - 1])) return -1;
  int g = (int)((x - b0) * inv_w);
  g = g < 0 ? 0 : (g > nbreaks - 2 ? nbreaks - 2 : g);
  while (x < breaks[g]) --g;          // cannot underflow: x > breaks[0]
  while (x > breaks[g + 1]) ++g;      // cannot overflow: x < breaks[nbreaks-1]
  // now breaks[g] <= x <= breaks[g+1]
  if (x == breaks[g] || x == breaks[g + 1]) return -1;
  return g;
}

// Sequential typed draws for one replicate held by ONE thread (C1 and the scalar parts of other models).
// Out-of-line copies of the numerics used at many call sites of the thread-per-replicate kernels: inlined everywhere
// they made those kernels 13-27 K instructions long, and ncu showed `no_inst` (instruction fetch) as the top stall
// (42 % of the samples of mix_driver_kernel).  Same arithmetic, one copy.
__device__ __noinline__ double ool_log(double x) { return ubm_log(x); }
__device__ __noinline__ double ool_exp(double x) { return ubm_exp(x); }
__device__ __noinline__ double ool_qnorm(double p) { return ubm_qnorm(p); }
__device__ __noinline__ ubm_u32x4 ool_stream_block(uint64_t seed, uint64_t rep, uint32_t type, uint64_t blk) {
  return ubm_stream_block(seed, rep, type, blk);
}

template <bool TAPE, bool OOL = false>      // OOL: call the out-of-line copies (kernels with many draw sites)
struct ScalarDraws {
  uint64_t seed, rep;
  uint64_t iu, in, ie;
  ubm_u32x4 bu, bn, be;
  const double *tu, *tn, *te;
  int64_t lu, ln, le;
  bool bad;

  __device__ void init(const RunParams& P, int64_t local_rep) {
    seed = P.seed; rep = (uint64_t)(P.rep_offset + local_rep);
    iu = in = ie = 0; bad = false;
    if (TAPE) {
      lu = P.tape_u ? P.len_u : 0; ln = P.tape_n ? P.len_n : 0; le = P.tape_e ? P.len_e : 0;
      tu = P.tape_u + local_rep * P.len_u; tn = P.tape_n + local_rep * P.len_n; te = P.tape_e + local_rep * P.len_e;
    }
  }
  __device__ double u() {
    uint64_t i = iu++;
    if (TAPE) { if ((int64_t)i >= lu) { bad = true; return 0.5; } return tu[i]; }
    uint32_t s = (uint32_t)i & 3u;
    if (s == 0) bu = OOL ? ool_stream_block(seed, rep, UBM_STREAM_U, i >> 2) : ubm_stream_block(seed, rep, UBM_STREAM_U, i >> 2);
    uint32_t w = (s == 0) ? bu.w[0] : (s == 1) ? bu.w[1] : (s == 2) ? bu.w[2] : bu.w[3];
    return ubm_u01_32(w);
  }
  __device__ double n() {
    uint64_t i = in++;
    if (TAPE) { if ((int64_t)i >= ln) { bad = true; return 0.0; } return tn[i]; }
    uint32_t s = (uint32_t)i & 1u;
    if (s == 0) bn = OOL ? ool_stream_block(seed, rep, UBM_STREAM_N, i >> 1) : ubm_stream_block(seed, rep, UBM_STREAM_N, i >> 1);
    double p = s ? ubm_u01_53(bn.w[2], bn.w[3]) : ubm_u01_53(bn.w[0], bn.w[1]);
    return OOL ? ool_qnorm(p) : ubm_qnorm(p);
  }
  __device__ double e() {
    uint64_t i = ie++;
    if (TAPE) { if ((int64_t)i >= le) { bad = true; return 1.0; } return te[i]; }
    uint32_t s = (uint32_t)i & 1u;
    if (s == 0) be = OOL ? ool_stream_block(seed, rep, UBM_STREAM_E, i >> 1) : ubm_stream_block(seed, rep, UBM_STREAM_E, i >> 1);
    double p = s ? ubm_u01_53(be.w[2], be.w[3]) : ubm_u01_53(be.w[0], be.w[1]);
    return OOL ? -ool_log(p) : -ubm_log(p);
  }
};

// random-access typed draws (any thread, any index) — block-parallel models
template <bool TAPE>
__device__ inline double draw_n_at(const RunParams& P, int64_t local_rep, uint64_t i, bool* bad) {
  if (TAPE) {
    if (!P.tape_n || (int64_t)i >= P.len_n) { *bad = true; return 0.0; }
    return P.tape_n[local_rep * P.len_n + (int64_t)i];
  }
  return ubm_philox_n(P.seed, (uint64_t)(P.rep_offset + local_rep), i);
}
template <bool TAPE>
__device__ inline double draw_u_at(const RunParams& P, int64_t local_rep, uint64_t i, bool* bad) {
  if (TAPE) {
    if (!P.tape_u || (int64_t)i >= P.len_u) { *bad = true; return 0.5; }
    return P.tape_u[local_rep * P.len_u + (int64_t)i];
  }
  return ubm_philox_u(P.seed, (uint64_t)(P.rep_offset + local_rep), i);
}

__device__ inline double dinf() { return __longlong_as_double(0x7FF0000000000000LL); }

}  // namespace ubm
