// oracle_core.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's three driver loops and two estimator routines.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build or call this.
//
// PARITY STATUS: pinned two ways.
//  (1) Against R's own printed output: on the emulated R session stream of r_stream.hpp (set.seed(1)), the restatement of
//      the R drivers and kernels reproduces the three literals of the reference's README (README.md:146, :171) digit for
//      digit (tests/test_r_stream.py).  These are the only literal outputs in the reference tree.
//  (2) Against the reference's own compiled code: oracle/_ref holds the UNMODIFIED src/*.cpp of the reference built against
//      stand-in R / Rcpp / Eigen headers (oracle/ref_shim); fed identical draws, the functions restated here agree with it
//      bit for bit on every integer / index / control-flow result and, where the reference fixes the operation order,
//      on the FP64 values too (tests/test_ref_parity.py).
//  What stays a restatement without a compiled counterpart: the interpreted R code that is not exercised by the README
//  (R/ising.R, R/variableselection.R, R/coupled_pairs01.R, R/unbiasedestimator.R) — pinned by the identities of
//  tests/test_oracle_*.py — and the arithmetic of R's nmath / Eigen, which are outside the reference tree.
//
// Randomness and transcendentals come from include/ubm_numerics.h (the engine's numerics contract, which
// replaces the R runtime layer L0 that is absent from /root/reference).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include "../include/ubm.h"
#include "../include/ubm_numerics.h"
#include "r_stream.hpp"

namespace orc {

// oracle-only third draw source: ONE sequential R-emulated stream (r_stream.hpp) shared by consecutive replicates, as in
// an R session.  ubm_draws.seed carries the RStream pointer; if tapes are given they are RECORDING buffers of capacity
// len_* per replicate: the typed draws each replicate consumed are written there (NaN beyond), so that the same draws
// can be replayed through UBM_DRAWS_TAPE on the device.
#define ORC_DRAWS_RSTREAM 2

// ---------------------------------------------------------------------------------------------------
// Draw source: typed sequential streams (U, N, E) of one replicate — Philox or replay tape.
// Order of consumption is part of the semantics.
// ---------------------------------------------------------------------------------------------------
struct Draws {
  int mode = UBM_DRAWS_PHILOX;
  uint64_t seed = 0, rep = 0;
  const double *tu = nullptr, *tn = nullptr, *te = nullptr;  // this replicate's tapes
  int64_t lu = 0, ln = 0, le = 0;
  uint64_t iu = 0, in = 0, ie = 0;
  bool exhausted = false;
  // block caches
  uint64_t cbu = ~0ull, cbn = ~0ull, cbe = ~0ull;
  ubm_u32x4 bu{}, bn{}, be{};
  RStream* rs = nullptr;                       // ORC_DRAWS_RSTREAM
  double *ru = nullptr, *rn = nullptr, *re = nullptr;   // recording buffers (capacity lu / ln / le)

  double u() {
    uint64_t i = iu++;
    if (mode == ORC_DRAWS_RSTREAM) { double v = rs->unif_rand(); if (ru && (int64_t)i < lu) ru[i] = v; return v; }
    if (mode == UBM_DRAWS_TAPE) {
      if ((int64_t)i >= lu) { exhausted = true; return 0.5; }
      return tu[i];
    }
    if ((i >> 2) != cbu) { cbu = i >> 2; bu = ubm_stream_block(seed, rep, UBM_STREAM_U, cbu); }
    return ubm_u01_32(bu.w[i & 3]);
  }
  // Philox streams only: skip to the next block boundary of the U stream (the Ising sweeps start on one, so that a
  // row's 32 uniforms are 8 whole blocks — include/ubm_numerics.h); tapes and the R stream are dense
  void align_u4() { if (mode == UBM_DRAWS_PHILOX) iu = (iu + 3ull) & ~3ull; }
  double n() {
    uint64_t i = in++;
    if (mode == ORC_DRAWS_RSTREAM) { double v = rs->norm_rand(); if (rn && (int64_t)i < ln) rn[i] = v; return v; }
    if (mode == UBM_DRAWS_TAPE) {
      if ((int64_t)i >= ln) { exhausted = true; return 0.0; }
      return tn[i];
    }
    if ((i >> 1) != cbn) { cbn = i >> 1; bn = ubm_stream_block(seed, rep, UBM_STREAM_N, cbn); }
    return ubm_qnorm(ubm_block_u53(bn, (uint32_t)(i & 1)));
  }
  double e() {
    uint64_t i = ie++;
    if (mode == ORC_DRAWS_RSTREAM) { double v = rs->exp_rand(); if (re && (int64_t)i < le) re[i] = v; return v; }
    if (mode == UBM_DRAWS_TAPE) {
      if ((int64_t)i >= le) { exhausted = true; return 1.0; }
      return te[i];
    }
    if ((i >> 1) != cbe) { cbe = i >> 1; be = ubm_stream_block(seed, rep, UBM_STREAM_E, cbe); }
    return -ubm_log(ubm_block_u53(be, (uint32_t)(i & 1)));
  }
};

inline Draws make_draws(const ubm_draws* d, int64_t local_rep, int64_t rep_offset) {
  Draws r;
  r.mode = d->mode;
  r.seed = d->seed;
  r.rep = (uint64_t)(rep_offset + local_rep);
  if (d->mode == UBM_DRAWS_TAPE) {
    r.lu = d->len_u; r.ln = d->len_n; r.le = d->len_e;
    r.tu = d->tape_u ? d->tape_u + local_rep * d->len_u : nullptr;
    r.tn = d->tape_n ? d->tape_n + local_rep * d->len_n : nullptr;
    r.te = d->tape_e ? d->tape_e + local_rep * d->len_e : nullptr;
    if (!r.tu) r.lu = 0;
    if (!r.tn) r.ln = 0;
    if (!r.te) r.le = 0;
  }
  if (d->mode == ORC_DRAWS_RSTREAM) {
    r.rs = reinterpret_cast<RStream*>((uintptr_t)d->seed);
    r.lu = d->len_u; r.ln = d->len_n; r.le = d->len_e;
    r.ru = d->tape_u ? const_cast<double*>(d->tape_u) + local_rep * d->len_u : nullptr;
    r.rn = d->tape_n ? const_cast<double*>(d->tape_n) + local_rep * d->len_n : nullptr;
    r.re = d->tape_e ? const_cast<double*>(d->tape_e) + local_rep * d->len_e : nullptr;
  }
  return r;
}

// canonical reduction over a d-vector (d <= 128): fma chain inside each group of 4 consecutive elements,
// then the xor-butterfly of include/ubm_numerics.h over the <= 32 group partials.
inline double sum_g4_sq(const double* v, int n) {
  double s[32];
  for (int q = 0; q < 32; ++q) s[q] = 0.0;
  for (int j = 0; j < n; ++j) s[j >> 2] = ubm_fma(v[j], v[j], s[j >> 2]);
  return ubm_butterfly32(s);
}
inline double sum_g4_dot(const double* a, const double* b, int n) {
  double s[32];
  for (int q = 0; q < 32; ++q) s[q] = 0.0;
  for (int j = 0; j < n; ++j) s[j >> 2] = ubm_fma(a[j], b[j], s[j >> 2]);
  return ubm_butterfly32(s);
}

// ---------------------------------------------------------------------------------------------------
// Test functions h (R closures in the reference; a registry here)
// ---------------------------------------------------------------------------------------------------
inline int h_dim(int h_kind, int dim) { return h_kind == UBM_H_BOTH ? 2 * dim : dim; }
inline void h_eval(int h_kind, const double* x, int dim, double* out) {
  if (h_kind == UBM_H_IDENTITY) { for (int i = 0; i < dim; ++i) out[i] = x[i]; }
  else if (h_kind == UBM_H_SQUARE) { for (int i = 0; i < dim; ++i) out[i] = x[i] * x[i]; }
  else { for (int i = 0; i < dim; ++i) { out[i] = x[i]; out[dim + i] = x[i] * x[i]; } }
}

// weight of Delta_t in H_{k:m}:  floor((t-k)/lag) - ceiling(max(lag, t-m)/lag) + 1
// R/unbiasedestimator.R:71,94 ; R/H_bar.R:47 ; src/estimator_bin.cpp:24   (integers; t >= k + lag)
inline int64_t corr_weight(int64_t t, int64_t k, int64_t m, int64_t lag) {
  int64_t fl = (t - k) / lag;                    // t - k >= lag > 0: floor == truncation
  int64_t a = (t - m > lag) ? (t - m) : lag;     // max(lag, t - m) > 0
  int64_t ce = (a + lag - 1) / lag;
  return fl - ce + 1;
}

// position of x among breaks: returns b with breaks[b] < x < breaks[b+1], or -1 (strict both sides,
// src/estimator_bin.cpp:16,32,35)
inline int bin_of(double x, const double* breaks, int nbreaks) {
  if (nbreaks < 2) return -1;
  if (!(x > breaks[0]) || !(x < breaks[nbreaks - 1])) return -1;
  int lo = 0, hi = nbreaks - 1;  // invariant breaks[lo] < x < breaks[hi]... refine
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (x > breaks[mid]) lo = mid; else if (x < breaks[mid]) hi = mid; else return -1;  // x == break: no bin
  }
  return lo;
}

struct RepResult {
  double meetingtime = std::numeric_limits<double>::infinity();
  int64_t iteration = 0;
  double cost = std::numeric_limits<double>::infinity();
  bool tape_exhausted = false;
};

// A Model provides:
//   typedef State;  int dim() const;
//   void rinit(State&, Draws&) const;  void single(State&, Draws&) const;
//   bool coupled(State& s1, State& s2, Draws&) const;   // returns `identical`
//   void copy(State& dst, const State& src) const;      // state2 <- state1 after meeting
//   const double* chain_state(const State&) const;      // dim doubles, what h sees and what is stored

// ---------------------------------------------------------------------------------------------------
// sample_meetingtime — R/meetingtime.R:24-48
// ---------------------------------------------------------------------------------------------------
template <class M>
RepResult sample_meetingtime(const M& model, Draws& D, int64_t lag, int64_t max_iterations) {
  typename M::State s1, s2;
  model.rinit(s1, D);                  // :27  state1 <- rinit(); state2 <- rinit()
  model.rinit(s2, D);
  int64_t time = 0;
  for (int64_t t = 0; t < lag; ++t) {  // :30-33
    time += 1;
    model.single(s1, D);
  }
  RepResult r;
  bool met = false;
  while (!met && (max_iterations <= 0 || time < max_iterations)) {   // :36
    time += 1;
    bool ident = model.coupled(s1, s2, D);                            // :39-41
    if (ident) { met = true; r.meetingtime = (double)time; }          // :43
  }
  r.iteration = time;
  if (met) r.cost = (double)(lag + 2 * (time - lag));
  r.tape_exhausted = D.exhausted;
  return r;
}

// ---------------------------------------------------------------------------------------------------
// sample_coupled_chains — R/coupled_chains.R:39-87.  Trajectories appended to samples1 / samples2
// (row t of samples1 = X_t, row t of samples2 = Y_t), each row `dim` doubles.
// ---------------------------------------------------------------------------------------------------
template <class M>
RepResult sample_coupled_chains(const M& model, Draws& D, int64_t m, int64_t lag, int64_t max_iterations,
                                std::vector<double>& samples1, std::vector<double>& samples2) {
  const int d = model.dim();
  typename M::State s1, s2;
  model.rinit(s1, D);                                                    // :41
  model.rinit(s2, D);
  auto push = [&](std::vector<double>& v, const typename M::State& s) {
    const double* x = model.chain_state(s);
    v.insert(v.end(), x, x + d);
  };
  samples1.clear(); samples2.clear();
  push(samples1, s1);                                                    // :46-47
  push(samples2, s2);
  int64_t time = 0;
  for (int64_t t = 0; t < lag; ++t) {                                    // :50-54
    time += 1;
    model.single(s1, D);
    push(samples1, s1);
  }
  RepResult r;
  bool met = false;
  int64_t tau = 0;
  // :58  while ((time < max(meetingtime, m)) && (time < max_iterations))
  while ((!met || time < std::max(tau, m)) && (max_iterations <= 0 || time < max_iterations)) {
    time += 1;
    if (met) {                                                           // :60-62
      model.single(s1, D);
      model.copy(s2, s1);
    } else {                                                             // :63-69
      bool ident = model.coupled(s1, s2, D);
      if (ident) { met = true; tau = time; }
    }
    push(samples1, s1);                                                  // :77-78
    push(samples2, s2);
  }
  r.iteration = time;
  if (met) {
    r.meetingtime = (double)tau;
    r.cost = (double)(lag + 2 * (tau - lag) + std::max<int64_t>(0, time - tau));   // :82
  }
  r.tape_exhausted = D.exhausted;
  return r;
}

// ---------------------------------------------------------------------------------------------------
// sample_unbiasedestimator — R/unbiasedestimator.R:43-104, with estimator_bin_ (src/estimator_bin.cpp:10-42)
// accumulated on the fly for every bin of `bins` (integer numerators; divided by m-k+1 at the end).
// mcmc / corr / uest: dimh doubles each.  binvals: nbins doubles (may be null).
// ---------------------------------------------------------------------------------------------------
template <class M>
RepResult sample_unbiasedestimator(const M& model, Draws& D, int h_kind, const ubm_bins* bins,
                                   int64_t k, int64_t m, int64_t lag, int64_t max_iterations,
                                   double* mcmc, double* corr, double* uest, double* binvals,
                                   int64_t* binnum = nullptr) {
  const int d = model.dim();
  const int dimh = h_dim(h_kind, d);
  const int nbins = (bins && bins->nbreaks >= 2) ? bins->nbreaks - 1 : 0;
  std::vector<double> h1(dimh), h2(dimh);
  std::vector<int64_t> bnum(nbins, 0);
  typename M::State s1, s2;
  model.rinit(s1, D);                                                    // :53
  model.rinit(s2, D);
  auto bin_add = [&](const typename M::State& s, int64_t w) {
    if (!nbins) return;
    int b = bin_of(model.chain_state(s)[bins->component], bins->breaks, bins->nbreaks);
    if (b >= 0) bnum[b] += w;
  };
  // :55-59  mcmcestimator <- h(state1) ; zeroed if k > 0
  h_eval(h_kind, model.chain_state(s1), d, h1.data());
  for (int i = 0; i < dimh; ++i) { mcmc[i] = (k > 0) ? 0.0 : h1[i]; corr[i] = 0.0; }
  if (k == 0) bin_add(s1, 1);
  int64_t time = 0;
  for (int64_t t = 0; t < lag; ++t) {                                    // :63-69
    time += 1;
    model.single(s1, D);
    if (time >= k) {
      h_eval(h_kind, model.chain_state(s1), d, h1.data());
      for (int i = 0; i < dimh; ++i) mcmc[i] = mcmc[i] + h1[i];
      bin_add(s1, 1);
    }
  }
  auto add_correction = [&]() {                                          // :70-72, :93-95
    double w = (double)corr_weight(time, k, m, lag);
    h_eval(h_kind, model.chain_state(s1), d, h1.data());
    h_eval(h_kind, model.chain_state(s2), d, h2.data());
    for (int i = 0; i < dimh; ++i) corr[i] = corr[i] + w * (h1[i] - h2[i]);
    bin_add(s1, corr_weight(time, k, m, lag));
    bin_add(s2, -corr_weight(time, k, m, lag));
  };
  if (time >= k + lag) add_correction();
  RepResult r;
  bool met = false;
  int64_t tau = 0;
  while ((!met || time < std::max(tau, m)) && (max_iterations <= 0 || time < max_iterations)) {   // :75
    time += 1;
    if (met) {                                                           // :77-83
      model.single(s1, D);
      model.copy(s2, s1);
      if (k <= time && time <= m) {
        h_eval(h_kind, model.chain_state(s1), d, h1.data());
        for (int i = 0; i < dimh; ++i) mcmc[i] = mcmc[i] + h1[i];
        bin_add(s1, 1);
      }
    } else {                                                             // :84-96
      bool ident = model.coupled(s1, s2, D);
      if (ident) { met = true; tau = time; }
      if (k <= time && time <= m) {
        h_eval(h_kind, model.chain_state(s1), d, h1.data());
        for (int i = 0; i < dimh; ++i) mcmc[i] = mcmc[i] + h1[i];
        bin_add(s1, 1);
      }
      if (time >= k + lag) add_correction();
    }
  }
  const double denom = (double)(m - k + 1);                              // :102
  for (int i = 0; i < dimh; ++i) {
    double ue = mcmc[i] + corr[i];                                       // :98
    uest[i] = ue / denom;
    mcmc[i] = mcmc[i] / denom;
    corr[i] = corr[i] / denom;
  }
  if (binvals) for (int b = 0; b < nbins; ++b) binvals[b] = (double)bnum[b] / denom;
  if (binnum) for (int b = 0; b < nbins; ++b) binnum[b] = bnum[b];
  r.iteration = time;
  if (met) {
    r.meetingtime = (double)tau;
    r.cost = (double)(lag + 2 * (tau - lag) + std::max<int64_t>(0, time - tau));   // :99
  }
  r.tape_exhausted = D.exhausted;
  return r;
}

// ---------------------------------------------------------------------------------------------------
// H_bar — R/H_bar.R:19-54 — from stored chains (samples1: (iteration+1) x dim row-major, samples2:
// (iteration-lag+1) x dim).  Sum over rows k..m first, then corrections in time order, then / (m-k+1).
// ---------------------------------------------------------------------------------------------------
inline void H_bar(const double* samples1, const double* samples2, int dim, int h_kind, int64_t k, int64_t m,
                  int64_t lag, double meetingtime, double* out) {
  const int dimh = h_dim(h_kind, dim);
  std::vector<double> h1(dimh), h2(dimh);
  for (int i = 0; i < dimh; ++i) out[i] = 0.0;
  for (int64_t t = k; t <= m; ++t) {                                     // :33-39
    h_eval(h_kind, samples1 + t * dim, dim, h1.data());
    for (int i = 0; i < dimh; ++i) out[i] = out[i] + h1[i];
  }
  if (!(meetingtime <= (double)(k + lag))) {                             // :42-51
    int64_t tau = (int64_t)meetingtime;
    for (int64_t t = k + lag; t <= tau - 1; ++t) {
      double w = (double)corr_weight(t, k, m, lag);
      h_eval(h_kind, samples1 + t * dim, dim, h1.data());
      h_eval(h_kind, samples2 + (t - lag) * dim, dim, h2.data());
      for (int i = 0; i < dimh; ++i) out[i] = out[i] + w * (h1[i] - h2[i]);
    }
  }
  for (int i = 0; i < dimh; ++i) out[i] = out[i] / (double)(m - k + 1);  // :53
}

// estimator_bin_ — src/estimator_bin.cpp:10-42 (component 0-based here)
inline double estimator_bin(const double* samples1, const double* samples2, int dim, int component,
                            double lower, double upper, int64_t k, int64_t m, int64_t lag, double meetingtime) {
  double estimator = 0;
  for (int64_t t = k; t <= m; ++t) {                                     // :15-19
    double x = samples1[t * dim + component];
    if (x > lower && x < upper) estimator += 1;
  }
  if (meetingtime > (double)(k + lag)) {                                 // :21
    int64_t tau = (int64_t)meetingtime;
    for (int64_t time = k + lag; time <= tau - 1; ++time) {              // :24
      double increment = 0.;
      double coefficient = std::floor(((double)time - k) / (double)lag) -
                           std::ceil(std::max<double>((double)lag, (double)time - m) / (double)lag) + 1.;   // :26
      double x = samples1[time * dim + component];
      double y = samples2[(time - lag) * dim + component];
      if (x > lower && x < upper) increment += coefficient;              // :32
      if (y > lower && y < upper) increment -= coefficient;              // :35
      estimator += increment;
    }
  }
  return estimator / (m - k + 1.);                                       // :41
}

}  // namespace orc
