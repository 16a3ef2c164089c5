// oracle_ising.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header for the parity status).
// CPU restatement of the reference's Ising parallel-tempering kernels (C4):
//   src/ising.cpp:11-35 (ising_sum_), :43-62 (ising_gibbs_sweep_), :67-92 (ising_coupled_gibbs_sweep_)
//   R/ising.R:4-8,30-36 (rinit), :45-77 (PT single kernel), :87-141 (PT coupled kernel)
// driven by the generic drivers of oracle_core.hpp.  The chain state seen by h() and by trajectory storage is
// the vector of natural statistics (one per temperature), as in ising_unbiased_estimator (R/ising.R:148-240),
// whose lag-1 weights min(1,(t-k)/(m-k+1)) equal the generic driver's.
#pragma once
#include <atomic>
#include <cstring>
#include "oracle_core.hpp"

namespace orc {

// swap-move counters of the reference's kernels (R/ising.R:48-49,76 / :89-91,137-140), summed over everything run since the
// last reset: [0] swap-move iterations, [1 .. 31] accepted swaps of chain 1 per pair, [32 .. 62] of chain 2
inline std::atomic<long long> g_ising_pt[64];

struct IsingModel {
  static const int SZ = 32;
  struct State {
    std::vector<int8_t> x;       // [nchains][32][32], x[c][i][j]
    std::vector<double> sums;    // natural statistics
  };
  int N = 0;
  std::vector<double> betas;
  std::vector<double> probas;    // [nchains][5]
  double pswap = 0;

  explicit IsingModel(const ubm_model_ising_pt* p) {
    N = p->nchains; pswap = p->proba_swapmove;
    betas.assign(p->betas, p->betas + N);
    probas.resize((size_t)N * 5);
    for (int c = 0; c < N; ++c)
      for (int q = 0; q < 5; ++q) {               // R/ising.R:151-152  exp(ss*beta) / (exp(ss*beta) + exp(-ss*beta))
        double ss = -4.0 + 2.0 * q;
        double a = ubm_exp(ss * betas[c]), b = ubm_exp(-ss * betas[c]);
        probas[(size_t)c * 5 + q] = a / (a + b);
      }
  }
  int dim() const { return N; }
  const double* chain_state(const State& s) const { return s.sums.data(); }
  void copy(State& dst, const State& src) const { dst = src; }
  int8_t& at(State& s, int c, int i, int j) const { return s.x[((size_t)c * SZ + i) * SZ + j]; }
  int8_t at(const State& s, int c, int i, int j) const { return s.x[((size_t)c * SZ + i) * SZ + j]; }

  // ising_sum_ — src/ising.cpp:11-35
  int ising_sum(const State& s, int c) const {
    int tot = 0;
    for (int i = 0; i < SZ; ++i) {
      int i1 = (i == SZ - 1) ? 0 : i + 1;
      for (int j = 0; j < SZ; ++j) {
        int j1 = (j == SZ - 1) ? 0 : j + 1;
        tot += at(s, c, i, j) * (at(s, c, i, j1) + at(s, c, i1, j));
      }
    }
    return tot;
  }
  void refresh_sums(State& s) const {
    for (int c = 0; c < N; ++c) s.sums[c] = (double)ising_sum(s, c);
  }
  // ising_rinit / ising_pt_rinit — R/ising.R:4-8,30-36: 2 * rbinom(1024, 1, 0.5) - 1, filled column-major.
  // rbinom(n=1, p=0.5) by inversion returns 1{U >= 0.5} (R nmath rbinom.c, n*p < 30 branch).
  void rinit(State& s, Draws& D) const {
    s.x.assign((size_t)N * SZ * SZ, 0);
    s.sums.assign(N, 0.0);
    for (int c = 0; c < N; ++c)
      for (int idx = 0; idx < SZ * SZ; ++idx) {
        int v = (D.u() >= 0.5) ? 1 : 0;
        at(s, c, idx % SZ, idx / SZ) = (int8_t)(2 * v - 1);
      }
    refresh_sums(s);                               // R/ising.R:156-157
  }
  // ising_gibbs_sweep_ — src/ising.cpp:43-62
  void sweep(State& s, int c, Draws& D) const {
    for (int i = 0; i < SZ; ++i)
      for (int j = 0; j < SZ; ++j) {
        int itop = (i + 1) % SZ, ibottom = (i + SZ - 1) % SZ, jright = (j + 1) % SZ, jleft = (j + SZ - 1) % SZ;
        int sn = at(s, c, itop, j) + at(s, c, ibottom, j) + at(s, c, i, jright) + at(s, c, i, jleft);
        at(s, c, i, j) = (int8_t)(2 * (D.u() < probas[(size_t)c * 5 + (sn + 4) / 2]) - 1);
      }
  }
  // ising_coupled_gibbs_sweep_ — src/ising.cpp:67-92: one uniform per site shared by both lattices
  void coupled_sweep(State& s1, State& s2, int c, Draws& D) const {
    for (int i = 0; i < SZ; ++i)
      for (int j = 0; j < SZ; ++j) {
        int itop = (i + 1) % SZ, ibottom = (i + SZ - 1) % SZ, jright = (j + 1) % SZ, jleft = (j + SZ - 1) % SZ;
        int n1 = at(s1, c, itop, j) + at(s1, c, ibottom, j) + at(s1, c, i, jright) + at(s1, c, i, jleft);
        int n2 = at(s2, c, itop, j) + at(s2, c, ibottom, j) + at(s2, c, i, jright) + at(s2, c, i, jleft);
        double u = D.u();
        at(s1, c, i, j) = (int8_t)(2 * (u < probas[(size_t)c * 5 + (n1 + 4) / 2]) - 1);
        at(s2, c, i, j) = (int8_t)(2 * (u < probas[(size_t)c * 5 + (n2 + 4) / 2]) - 1);
      }
  }
  void swap_lattices(State& s, int c) const {
    for (int q = 0; q < SZ * SZ; ++q) std::swap(s.x[(size_t)c * SZ * SZ + q], s.x[(size_t)(c + 1) * SZ * SZ + q]);
    std::swap(s.sums[c], s.sums[c + 1]);
  }
  // ising_pt_single_kernel — R/ising.R:45-77
  void single(State& s, Draws& D) const {
    double u_iteration = D.u();
    if (u_iteration < pswap) {
      g_ising_pt[0] += 1;
      for (int c = 0; c + 1 < N; ++c) {
        double logprob = (betas[c] - betas[c + 1]) * (s.sums[c + 1] - s.sums[c]);   // :55-56
        double u = D.u();
        if (ubm_log(u) < logprob) { swap_lattices(s, c); g_ising_pt[1 + c] += 1; }
      }
    } else {
      D.align_u4();
      for (int c = 0; c < N; ++c) sweep(s, c, D);
    }
    refresh_sums(s);
  }
  // ising_pt_coupled_kernel — R/ising.R:87-141; identical <=> all lattice pairs equal (R/ising.R:215)
  bool coupled(State& s1, State& s2, Draws& D) const {
    double u_iteration = D.u();
    if (u_iteration < pswap) {
      g_ising_pt[0] += 1;
      for (int c = 0; c + 1 < N; ++c) {
        double deltabeta = betas[c] - betas[c + 1];
        double lp1 = deltabeta * (s1.sums[c + 1] - s1.sums[c]);
        double lp2 = deltabeta * (s2.sums[c + 1] - s2.sums[c]);
        double logu = ubm_log(D.u());
        if (logu < lp1) { swap_lattices(s1, c); g_ising_pt[1 + c] += 1; }
        if (logu < lp2) { swap_lattices(s2, c); g_ising_pt[32 + c] += 1; }
      }
    } else {
      D.align_u4();
      for (int c = 0; c < N; ++c) coupled_sweep(s1, s2, c, D);
    }
    refresh_sums(s1);
    refresh_sums(s2);
    return s1.x == s2.x;
  }
};

}  // namespace orc
