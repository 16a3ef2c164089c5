// ubm_varsel.cuh — C5: spike-and-slab Bayesian variable selection, flip / swap MH on gamma in {0,1}^p, coupled.
//
// Replaces, fused into one persistent kernel per launch:
//   drivers   R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//             (= coupled_chains_vs / unbiasedestimator_vs, R/variableselection.R:151-279: the generic lag-1 estimator)
//   kernels   R/variableselection.R:8-16 (prior, rinit), :17-55 (single), :57-148 (coupled)
//   numerics  src/vs_mlikelihood.cpp:5-28, src/sample_pairs01.cpp:12-42, R/coupled_pairs01.R:3-40
//
// Design (B200): one WARP per replicate (two per CTA, 14 per SM).  gamma is a bitmap: lane w keeps bits 32w..32w+31 of both chains in two
// registers, so sums, equality tests and all index sampling are popcounts + warp scans (no O(p) loops, no memory).
// G = X'X and X'Y are formed ONCE on the device (varsel_gram_kernel) and stay L2-resident (8 MB for p = 1000); a
// marginal likelihood gathers the s x s block G[gamma, gamma] into shared memory (one L2 round trip), runs a
// right-looking reciprocal-scaled Cholesky with the right-hand side appended as an extra row (so the solve comes free)
// and returns in a few thousand cycles — the reference instead re-reads the n x p design matrix, converts it to
// float32 and inverts with Eigen on every call.  The estimator of the p inclusion probabilities is accumulated
// lazily in exact integers: a coordinate's counter is only touched when that coordinate flips.
#pragma once
#include <string>
#include <vector>
#include "ubm_common.cuh"

namespace ubm {

#define VS_MAXW 32
#define FULLMASK 0xffffffffu
#ifndef VS_SFAST
#define VS_SFAST 59      // selections of at most this many variables are factorised in shared memory (16 KB per warp), larger ones in an L2-resident per-warp buffer
                         // (measured on C5: 63 -> 4.65 K est/s at 12 warps/SM, 59 -> 5.58 K at 14, 55 -> 5.3 K at 16, 51 -> 4.2 K: the fallback is slow)
#endif
#ifndef VS_WPB
#define VS_WPB 2         // replicates (warps) per CTA
#endif
#ifndef VS_MINB
#define VS_MINB 7        // CTAs per SM asked of the register allocator (x VS_WPB warps)
#endif

struct VarselDevice {
  int n, p, s0, nwords;
  double g, kappa, prop_flip, Y2, lg1, gg, logp;
  const double *G, *v;     // device: G p x p (symmetric), v [p]
};

// G[a][b] = sum_i X(i,a) X(i,b), v[a] = sum_i X(i,a) Y(i): sequential fma over observations (the oracle's order)
__global__ void varsel_gram_kernel(const double* __restrict__ X, const double* __restrict__ Y, int n, int p,
                                   double* __restrict__ G, double* __restrict__ v) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)p * p) return;
  const int a = (int)(e / p), b = (int)(e % p);
  if (b > a) return;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc = ubm_fma(X[(size_t)a * n + i], X[(size_t)b * n + i], acc);
  G[(size_t)a * p + b] = acc; G[(size_t)b * p + a] = acc;
  if (b == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t = ubm_fma(X[(size_t)a * n + i], Y[i], t);
    v[a] = t;
  }
}

__device__ __forceinline__ int warp_excl_scan(int c, int lane) {
  int incl = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int nb = __shfl_up_sync(FULLMASK, incl, off);
    if (lane >= off) incl += nb;
  }
  return incl - c;
}
__device__ __forceinline__ int warp_total(int c) { return (int)__reduce_add_sync(FULLMASK, (unsigned)c); }

// weighted index by inversion in natural order (oracle: VarselModel::inv2): class A bits weigh wA, class B bits wB;
// smallest position k with cA(k) wA + cB(k) wB > u, else the last position of positive weight.  Lane w holds word w.
__device__ inline int vs_inv2(unsigned wordA, double wA, unsigned wordB, double wB, double u, int lane) {
  const int cA = __popc(wordA), cB = __popc(wordB);
  const int pA = warp_excl_scan(cA, lane), pB = warp_excl_scan(cB, lane);
  const bool any = (wordA | wordB) != 0u;
  const double cum_end = (double)(pA + cA) * wA + (double)(pB + cB) * wB;
  const unsigned hit = __ballot_sync(FULLMASK, any && cum_end > u);
  if (hit) {
    const int L = __ffs(hit) - 1;
    const unsigned a = __shfl_sync(FULLMASK, wordA, L), b = __shfl_sync(FULLMASK, wordB, L);
    const int qa = __shfl_sync(FULLMASK, pA, L), qb = __shfl_sync(FULLMASK, pB, L);
    const unsigned upto = (lane == 31) ? 0xffffffffu : ((2u << lane) - 1u);
    const bool here = ((a | b) >> lane) & 1u;
    const double cum = (double)(qa + __popc(a & upto)) * wA + (double)(qb + __popc(b & upto)) * wB;
    const unsigned h2 = __ballot_sync(FULLMASK, here && cum > u);
    return 32 * L + (__ffs(h2) - 1);
  }
  // fallback: last position whose class has positive weight
  const unsigned pos = ((wA > 0) ? wordA : 0u) | ((wB > 0) ? wordB : 0u);
  const unsigned nz = __ballot_sync(FULLMASK, pos != 0u);
  if (!nz) return -1;
  const int L = 31 - __clz(nz);
  const unsigned w = __shfl_sync(FULLMASK, pos, L);
  return 32 * L + (31 - __clz(w));
}

// doubles of per-warp work space: the packed Cholesky rows, or (during rinit, which aliases it) p ints + s0 doubles
__host__ __device__ inline int varsel_work_doubles(int s0, int p) {
  const int work = (s0 + 2) * (s0 + 3) / 2, init = (p + 1) / 2 + s0 + 2;
  return work > init ? work : init;
}

struct VsWork {           // per-warp work space
  double* Lbs;            // shared: packed lower triangle for s <= VS_SFAST, rows 0..s (row s = right-hand side / solution)
  double* Lbg;            // global (L2): the same for larger s, (s0+2)(s0+3)/2 doubles; also the rinit scratch
  double* colk;           // shared [VS_SFAST + 1]: the scaled column of the current elimination step
  int* idx;               // shared [s0 + 2]
};
__host__ __device__ inline int varsel_fast_doubles() { return (VS_SFAST + 2) * (VS_SFAST + 3) / 2; }

#ifdef UBM_VS_TIMING
static __device__ long long g_vs_t[8];      // debug builds: cycles of warp (0,0): gather, cholesky, calls, sum of s, steps, total
#define VSMARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long n_ = clock64(); g_vs_t[i] += n_ - vt_; vt_ = n_; } } while (0)
#else
#define VSMARK(i) do { } while (0)
#endif
// Cholesky of G[idx, idx] with the right-hand side as row s, in the packed buffer Lb; returns |L^{-1} v_g|^2
template <bool RIGHT>
__device__ __forceinline__ double vs_chol_core(const VarselDevice& M, double* __restrict__ Lb, double* __restrict__ colk,
                                               const int* __restrict__ idx, int s, int lane) {
#ifdef UBM_VS_TIMING
  long long vt_ = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) { g_vs_t[2] += 1; g_vs_t[3] += s; }
#endif
  // gather G[idx, idx] (lower) and the right-hand side as row s: four independent L2 loads in flight per lane
  const int tot = (s + 1) * (s + 2) / 2 - 1;           // last row has s entries (no diagonal)
  for (int e0 = lane; e0 < tot; e0 += 128) {
    double val[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = e0 + 32 * q;
      val[q] = 0.0;
      if (e < tot) {
        int a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while ((a + 1) * (a + 2) / 2 <= e) ++a;
        while (a * (a + 1) / 2 > e) --a;
        const int b = e - a * (a + 1) / 2;
        val[q] = (a < s) ? __ldg(M.G + (size_t)idx[a] * M.p + idx[b]) : __ldg(M.v + idx[b]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) if (e0 + 32 * q < tot) Lb[e0 + 32 * q] = val[q];
  }
  __syncwarp();
  VSMARK(0);
  double l = 0.0;
  if (RIGHT) {
    // Cholesky, right-looking, reciprocal-scaled (dpotf2): at step k every remaining entry (i, c), k < c <= i <= s,
    // takes its k-th update A[i][c] = fma(-L[i][k], L[c][k], A[i][c]) with L[x][k] = A[x][k] * (1 / L[k][k]).  Each
    // entry therefore sees exactly the fma sequence k = 0, 1, ... of the textbook left-looking loop (the oracle's), but
    // the updates of one step are independent, so a lane keeps many in flight instead of walking one dependent chain.
    // Per step: the scaled column goes to a small vector `colk` (one entry per lane), one barrier, then lane r updates
    // rows k+1+r, k+33+r; the next pivot is handed on by shuffle from lane 0, which always owns row k+1.
    // Row s carries the forward substitution L y = v_g: |y|^2 is accumulated by lane 0 as the y_k appear.
    double piv = Lb[0];
    for (int k = 0; k < s; ++k) {
      const double rinv = 1.0 / sqrt(piv);
      for (int i = k + 1 + lane; i <= s; i += 32) colk[i] = Lb[i * (i + 1) / 2 + k] * rinv;
      __syncwarp();
      if (lane == 0) { const double yk = colk[s]; l = ubm_fma(yk, yk, l); }
      double newpiv = 0.0;
      for (int i = k + 1 + lane; i <= s; i += 32) {
        double* __restrict__ row = Lb + i * (i + 1) / 2;
        const double nl = -colk[i];
        const int cmax = (i == s) ? s - 1 : i;           // the right-hand-side row has no diagonal entry
        // eight independent entries per trip, guarded (a scalar remainder loop of dependent load-fma-store trips cost
        // as much as the unrolled part: ncu, profiles/r01_other_kernels.md)
        for (int c0 = k + 1; c0 <= cmax; c0 += 8) {
          double a[8], b[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const bool ok = c0 + q <= cmax;
            a[q] = ok ? row[c0 + q] : 0.0;
            b[q] = ok ? colk[c0 + q] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) if (c0 + q <= cmax) row[c0 + q] = ubm_fma(nl, b[q], a[q]);
        }
        if (i == k + 1) newpiv = row[k + 1];             // entry (k+1, k+1): the next pivot (lane 0)
      }
      piv = __shfl_sync(FULLMASK, newpiv, 0);
      __syncwarp();
    }
  } else {
    // left-looking, column by column (global-memory buffer: finished columns are only read, so they stay in L1)
    for (int j = 0; j < s; ++j) {
      const int rj = j * (j + 1) / 2;
      double t[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int i = lane + 32 * q;
        t[q] = 0.0;
        if (i >= j && i <= s) {
          const int ri = i * (i + 1) / 2;
          double d = Lb[ri + j];
          for (int k = 0; k < j; ++k) d = ubm_fma(-Lb[ri + k], Lb[rj + k], d);
          t[q] = d;
        }
      }
      const int ql = j >> 5;
      const double sjj = __shfl_sync(FULLMASK, (ql == 0) ? t[0] : (ql == 1) ? t[1] : (ql == 2) ? t[2] : (ql == 3) ? t[3] : t[4], j & 31);
      const double ljj = sqrt(sjj);
      const double rinv = 1.0 / ljj;
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int i = lane + 32 * q;
        if (i == j) Lb[rj + j] = ljj;
        else if (i > j && i <= s) Lb[i * (i + 1) / 2 + j] = t[q] * rinv;
      }
      __syncwarp();
    }
  }
  if (!RIGHT && lane == 0) {
    const int rs = s * (s + 1) / 2;
    for (int i = 0; i < s; ++i) l = ubm_fma(Lb[rs + i], Lb[rs + i], l);
  }
  l = __shfl_sync(FULLMASK, l, 0);
  __syncwarp();
  VSMARK(1);
  return l;
}

// marginal_likelihood_c_2 (src/vs_mlikelihood.cpp:5-28) of the selection held one word per lane
__device__ __forceinline__ double vs_ml(const VarselDevice& M, const VsWork& W, unsigned word, int lane) {
  const int c = __popc(word);
  const int off = warp_excl_scan(c, lane);
  const int s = warp_total(c);
  double l = 0.0;
  if (s > 0) {
    {
      unsigned w = word;
      int t = off;
      while (w) { const int b = __ffs(w) - 1; w &= w - 1; W.idx[t++] = 32 * lane + b; }
    }
    __syncwarp();
    // two copies of the same code so that the small case compiles to LDS/STS and the large one to LDG/STG
    if (s <= VS_SFAST) l = vs_chol_core<true>(M, W.Lbs, W.colk, W.idx, s, lane);
    else l = vs_chol_core<false>(M, W.Lbg, W.colk, W.idx, s, lane);
  }
  return -((double)s + 1.) / 2. * M.lg1 - (double)M.n / 2. * ubm_log(M.Y2 - M.gg * l);   // :26
}

__device__ __forceinline__ double vs_prior(const VarselDevice& M, int s) {   // R/variableselection.R:8-11
  if (s > M.s0) return -dinf();
  return -M.kappa * (double)s * M.logp;
}
__device__ __forceinline__ unsigned vs_bit(int pos, int lane) { return ((pos >> 5) == lane) ? (1u << (pos & 31)) : 0u; }

// exact lazy bookkeeping of sum_{t in [k, m]} gamma1_t[j]: counters live in global memory, touched only on flips
struct VsEst {
  int* cnt; int* last; long long* corr;   // [p] each (per resident warp)
  long long k, m;
  __device__ __forceinline__ long long window(long long a, long long b) const {   // |[a, b] ∩ [k, m]|
    const long long lo = a > k ? a : k, hi = b < m ? b : m;
    return hi >= lo ? hi - lo + 1 : 0;
  }
};

template <bool TAPE>
__global__ void __launch_bounds__(32 * VS_WPB, VS_MINB) varsel_driver_kernel(const VarselDevice M, const RunParams P, int* scratch_i,
                                                                            long long* scratch_l, double* scratch_d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * VS_WPB + wib;              // resident-warp index: owns one slice of every scratch array
  const int p = M.p;
  const int idx_ints = (M.s0 + 2 > 34 ? M.s0 + 2 : 34) + ((M.s0 + 2 > 34 ? M.s0 + 2 : 34) & 1);
  const size_t per_warp = (size_t)(varsel_fast_doubles() + VS_SFAST + 1) * sizeof(double) + (size_t)idx_ints * sizeof(int);
  VsWork W;
  W.Lbs = reinterpret_cast<double*>(smem_raw + wib * per_warp);
  W.colk = W.Lbs + varsel_fast_doubles();
  W.idx = reinterpret_cast<int*>(W.colk + VS_SFAST + 1);
  W.Lbg = scratch_d + (size_t)gw * varsel_work_doubles(M.s0, p);
  const unsigned vmask = (32 * lane + 32 <= p) ? 0xffffffffu : ((32 * lane >= p) ? 0u : ((1u << (p - 32 * lane)) - 1u));
  VsEst E;
  E.cnt = scratch_i + (size_t)gw * 2 * p; E.last = E.cnt + p; E.corr = scratch_l + (size_t)gw * p;
  E.k = P.k; E.m = P.m;
  const double denom = (double)(P.m - P.k + 1);

  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(P.work_counter, 1ull);
    got = __shfl_sync(FULLMASK, got, 0);
    if (got >= (unsigned long long)P.nrep) break;
    const long long rep = (long long)got;
    ScalarDraws<TAPE> D;          // every lane runs the same sequential U stream
    D.init(P, rep);
    // ---- rinit (R/variableselection.R:12-16), twice; then pdf = ml + prior (:154-155)
    unsigned g1 = 0, g2 = 0;
    int s1 = 0, s2 = 0;
    double pdf1 = 0, pdf2 = 0;
    for (int chain = 0; chain < 2; ++chain) {
      const int kk = M.s0 < p ? M.s0 : p;
      int* x = reinterpret_cast<int*>(W.Lbg);           // p ints: remaining indices (aliases the global work buffer)
      unsigned* gws = reinterpret_cast<unsigned*>(W.idx);   // not used yet: idx has s0 + 2 >= 32 ints? ensured by launch
      double* ub = W.Lbg + (p + 1) / 2;                 // kk doubles after the int array
      for (int j = lane; j < p; j += 32) x[j] = j;
      if (lane < VS_MAXW) gws[lane] = 0u;
      __syncwarp();
      for (int t = 0; t < kk; ++t) { const double u = D.u(); if (lane == 0) ub[t] = u; }
      int nn = p;
      for (int t = 0; t < kk; ++t) {
        int j = (int)((double)nn * D.u());
        if (j >= nn) j = nn - 1;
        if (lane == 0) {
          const int pos = x[j];
          if (ub[t] < 0.5) gws[pos >> 5] |= 1u << (pos & 31);
          x[j] = x[nn - 1];
        }
        --nn;
        __syncwarp();
      }
      const unsigned w = gws[lane];
      __syncwarp();
      const int s = warp_total(__popc(w));
      const double pdf = vs_ml(M, W, w, lane) + vs_prior(M, s);
      if (chain == 0) { g1 = w; s1 = s; pdf1 = pdf; } else { g2 = w; s2 = s; pdf2 = pdf; }
    }
    long long time = 0, tau = 0;
    bool met = false;
    if (P.mode == MODE_ESTIMATOR) {
      for (int j = lane; j < p; j += 32) { E.cnt[j] = 0; E.last[j] = 0; E.corr[j] = 0; }
    } else if (P.mode == MODE_CHAINS) {
      for (int b = 0; b < 32; ++b) {
        const int j = 32 * lane + b;
        if (j < p) {
          P.samples1[(rep * P.stride) * p + j] = (double)((g1 >> b) & 1u);
          P.samples2[(rep * P.stride) * p + j] = (double)((g2 >> b) & 1u);
        }
      }
    }
    __syncwarp();
    // ---- time loop (generic drivers)
    for (;;) {
      bool go;
      if (time < P.lag) go = true;
      else {
        const bool capped = (P.max_it > 0 && time >= P.max_it) || D.bad;
        if (P.mode == MODE_MEETING) go = !met && !capped;
        else go = (!met || time < (tau > P.m ? tau : P.m)) && !capped;
      }
      if (!go) break;
      time += 1;
#ifdef UBM_VS_TIMING
      if (blockIdx.x == 0 && threadIdx.x == 0) { const long long n_ = clock64(); if (g_vs_t[6]) g_vs_t[5] += n_ - g_vs_t[6]; g_vs_t[6] = n_; g_vs_t[4] += 1; }
#endif
      const bool coupled_step = (time > P.lag) && !met;
      const unsigned old1 = g1;
      // The branches below only PROPOSE (in the reference's draw order); the marginal likelihoods, which consume no
      // draws, are evaluated in one place afterwards: two call sites instead of nine, and one evaluation when both
      // chains propose the same selection.
      bool need1 = false, need2 = false;
      unsigned prop1 = 0u, prop2 = 0u;
      int sp1 = 0, sp2 = 0;
      double logu1 = 0.0, logu2 = 0.0;
      if (!coupled_step) {
        // ---- single_kernel (R/variableselection.R:17-55) on chain 1
        if (D.u() < M.prop_flip) {
          int c = (int)((double)p * D.u());
          if (c >= p) c = p - 1;
          const unsigned bit = vs_bit(c, lane);
          const int was = (int)__reduce_or_sync(FULLMASK, (g1 & bit) ? 1u : 0u);
          sp1 = s1 + (was ? -1 : 1);
          if (isfinite(vs_prior(M, sp1))) {
            need1 = true; prop1 = g1 ^ bit;
            logu1 = ubm_log(D.u());                                   // drawn only if the prior is finite (:24-26)
          }
        } else if (s1 > 0 && s1 < p) {
          const double u0 = D.u(), u1 = D.u();                        // src/sample_pairs01.cpp:20-24
          const int i0 = vs_inv2(~g1 & vmask, 1. / (double)(p - s1), 0u, 0.0, u0, lane);
          const int i1 = vs_inv2(g1, 1. / (double)s1, 0u, 0.0, u1, lane);
          sp1 = s1;
          if (isfinite(vs_prior(M, sp1))) {
            need1 = true; prop1 = (g1 | vs_bit(i0, lane)) & ~vs_bit(i1, lane);
            logu1 = ubm_log(D.u());
          }
        }
      } else {
        // ---- coupled_kernel (R/variableselection.R:57-148)
        if (D.u() < M.prop_flip) {
          int c = (int)((double)p * D.u());
          if (c >= p) c = p - 1;
          const unsigned bit = vs_bit(c, lane);
          const int was1 = (int)__reduce_or_sync(FULLMASK, (g1 & bit) ? 1u : 0u);
          const int was2 = (int)__reduce_or_sync(FULLMASK, (g2 & bit) ? 1u : 0u);
          logu1 = logu2 = ubm_log(D.u());                             // :67 always drawn
          sp1 = s1 + (was1 ? -1 : 1); sp2 = s2 + (was2 ? -1 : 1);
          prop1 = g1 ^ bit; prop2 = g2 ^ bit;
          need1 = isfinite(vs_prior(M, sp1)); need2 = isfinite(vs_prior(M, sp2));
        } else {
          const bool sw1 = s1 > 0 && s1 < p, sw2 = s2 > 0 && s2 < p;
          if (sw1 && sw2) {
            int o1, o2, e1, e2;
            // coupled_pairs01: ones first (R/coupled_pairs01.R:5-21), then zeros (:23-38)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const unsigned m1 = half ? (~g1 & vmask) : g1, m2 = half ? (~g2 & vmask) : g2;
              const int c1 = half ? p - s1 : s1, c2 = half ? p - s2 : s2;
              const double a = 1. / (double)c1, b = 1. / (double)c2;
              const double mn = a < b ? a : b;
              const unsigned common = m1 & m2;
              const int C = warp_total(__popc(common));
              const double alpha = (double)C * mn;
              int x1, x2;
              if (D.u() < alpha) {
                x1 = vs_inv2(common, mn / alpha, 0u, 0.0, D.u(), lane);
                x2 = x1;
              } else {
                x1 = vs_inv2(common, (a - mn) / (1. - alpha), m1 & ~m2, a / (1. - alpha), D.u(), lane);
                x2 = vs_inv2(common, (b - mn) / (1. - alpha), m2 & ~m1, b / (1. - alpha), D.u(), lane);
              }
              if (half) { e1 = x1; e2 = x2; } else { o1 = x1; o2 = x2; }
            }
            prop1 = (g1 | vs_bit(e1, lane)) & ~vs_bit(o1, lane);      // R/variableselection.R:92-97
            prop2 = (g2 | vs_bit(e2, lane)) & ~vs_bit(o2, lane);
            logu1 = logu2 = ubm_log(D.u());                           // :100 always drawn
            sp1 = s1; sp2 = s2;
            need1 = isfinite(vs_prior(M, sp1)); need2 = isfinite(vs_prior(M, sp2));
          } else {
            if (sw1) {                                                // :118-131
              const double u0 = D.u(), u1 = D.u();
              const int i0 = vs_inv2(~g1 & vmask, 1. / (double)(p - s1), 0u, 0.0, u0, lane);
              const int i1 = vs_inv2(g1, 1. / (double)s1, 0u, 0.0, u1, lane);
              sp1 = s1;
              if (isfinite(vs_prior(M, sp1))) {
                need1 = true; prop1 = (g1 | vs_bit(i0, lane)) & ~vs_bit(i1, lane);
                logu1 = ubm_log(D.u());
              }
            }
            if (sw2) {                                                // :133-145
              const double u0 = D.u(), u1 = D.u();
              const int i0 = vs_inv2(~g2 & vmask, 1. / (double)(p - s2), 0u, 0.0, u0, lane);
              const int i1 = vs_inv2(g2, 1. / (double)s2, 0u, 0.0, u1, lane);
              sp2 = s2;
              if (isfinite(vs_prior(M, sp2))) {
                need2 = true; prop2 = (g2 | vs_bit(i0, lane)) & ~vs_bit(i1, lane);
                logu2 = ubm_log(D.u());
              }
            }
          }
        }
      }
      // ---- Metropolis-Hastings accept steps (R/variableselection.R:27-30, :68-79, :101-112)
      double ml1 = 0.0;
      if (need1) {
        ml1 = vs_ml(M, W, prop1, lane);
        const double ppdf = ml1 + vs_prior(M, sp1);
        if (logu1 < ppdf - pdf1) { g1 = prop1; s1 = sp1; pdf1 = ppdf; }
      }
      if (need2) {
        const bool same = need1 && __all_sync(FULLMASK, prop1 == prop2);
        const double ppdf = (same ? ml1 : vs_ml(M, W, prop2, lane)) + vs_prior(M, sp2);
        if (logu2 < ppdf - pdf2) { g2 = prop2; s2 = sp2; pdf2 = ppdf; }
      }
      if (coupled_step && __all_sync(FULLMASK, g1 == g2)) { met = true; tau = time; }   // :187, :258
      // ---- estimator / trajectories
      if (P.mode == MODE_ESTIMATOR) {
        unsigned ch = old1 ^ g1;                                      // chain-1 coordinates that flipped at `time`
        while (ch) {
          const int b = __ffs(ch) - 1; ch &= ch - 1;
          const int j = 32 * lane + b;
          if ((g1 >> b) & 1u) E.last[j] = (int)time;                  // 0 -> 1: present from X_time on
          else E.cnt[j] += (int)E.window((long long)E.last[j], time - 1);   // 1 -> 0: present over [last, time-1]
        }
        if ((time >= P.k + P.lag) && (coupled_step || time == P.lag)) {
          const long long w = corr_weight(time, P.k, P.m, P.lag);
          unsigned df = g1 ^ g2;
          while (df) {
            const int b = __ffs(df) - 1; df &= df - 1;
            E.corr[32 * lane + b] += ((g1 >> b) & 1u) ? w : -w;
          }
        }
      } else if (P.mode == MODE_CHAINS) {
        const unsigned y = (time > P.lag && !coupled_step) ? g1 : g2;
        for (int b = 0; b < 32; ++b) {
          const int j = 32 * lane + b;
          if (j < p) {
            P.samples1[(rep * P.stride + time) * p + j] = (double)((g1 >> b) & 1u);
            if (time > P.lag) P.samples2[(rep * P.stride + (time - P.lag)) * p + j] = (double)((y >> b) & 1u);
          }
        }
      }
    }
    // ---- results
    if (lane == 0) {
      double cost = dinf();
      if (met) {
        if (P.mode == MODE_MEETING) cost = (double)(P.lag + 2 * (tau - P.lag));
        else cost = (double)(P.lag + 2 * (tau - P.lag) + ((time - tau) > 0 ? (time - tau) : 0));
      }
      if (P.tau) P.tau[rep] = met ? (double)tau : dinf();
      if (P.cost) P.cost[rep] = cost;
      if (P.iter) P.iter[rep] = time;
      if (D.bad) atomicExch(P.status, UBM_ERR_TAPE);
    }
    if (P.mode == MODE_ESTIMATOR) {
      for (int b = 0; b < 32; ++b) {
        const int j = 32 * lane + b;
        if (j < p) {
          long long c = E.cnt[j];
          if ((g1 >> b) & 1u) c += E.window((long long)E.last[j], time);
          const double mc = (double)c, co = (double)E.corr[j];
          const double ue = mc + co;
          P.uest[rep * p + j] = ue / denom;
          if (P.mcmc) P.mcmc[rep * p + j] = mc / denom;
          if (P.corr) P.corr[rep * p + j] = co / denom;
          if (P.nbreaks >= 2 && j == P.component) {
            // estimator_bin_ numerators (src/estimator_bin.cpp:15-38) of a 0/1 coordinate, in closed form: it was 1 at c of the
            // window's `terms` times and 0 at the others; every correction term moves its weight between the two values
            const int nb = P.nbreaks - 1;
            const double b0 = P.breaks[0], iw = (double)nb / (P.breaks[nb] - P.breaks[0]);
            const long long terms = E.window(0, time);
            const int bin1 = bin_of(1.0, P.breaks, P.nbreaks, b0, iw), bin0 = bin_of(0.0, P.breaks, P.nbreaks, b0, iw);
            if (bin1 >= 0) P.bin_num[rep * nb + bin1] += c + E.corr[j];
            if (bin0 >= 0) P.bin_num[rep * nb + bin0] += (terms - c) - E.corr[j];
          }
        }
      }
    }
    __syncwarp();
#ifdef UBM_VS_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      printf("varsel timing (warp 0 of block 0): steps %lld cycles/step %.0f | ml calls/step %.2f mean s %.1f gather %.0f chol %.0f cycles/call\n",
             g_vs_t[4], (double)g_vs_t[5] / g_vs_t[4], (double)g_vs_t[2] / g_vs_t[4], (double)g_vs_t[3] / g_vs_t[2],
             (double)g_vs_t[0] / g_vs_t[2], (double)g_vs_t[1] / g_vs_t[2]);
      g_vs_t[6] = 0;
    }
#endif
  }
}

inline size_t varsel_smem_bytes(int s0) {      // per CTA
  int idx_ints = std::max(s0 + 2, 34);
  idx_ints += idx_ints & 1;
  return (size_t)VS_WPB * ((size_t)(varsel_fast_doubles() + VS_SFAST + 1) * sizeof(double) + (size_t)idx_ints * sizeof(int));
}

template <class Buf>
inline int varsel_launch(const VarselDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st, Buf& scratch,
                         int64_t* launches, std::string* err) {
  if (P.mode == MODE_ESTIMATOR && P.h_kind != UBM_H_IDENTITY) { *err = "variable selection: h must be the identity (inclusion indicators)"; return UBM_ERR_UNSUPPORTED; }
  const size_t smem = varsel_smem_bytes(M.s0);
  auto kern = tape ? varsel_driver_kernel<true> : varsel_driver_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = std::string("variable selection: smem attribute: ") + cudaGetErrorString(e); return UBM_ERR_CUDA; }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * VS_WPB, smem);
  if (e != cudaSuccess || per_sm < 1) { *err = "variable selection: occupancy query failed"; return UBM_ERR_CUDA; }
  const int grid = (int)std::min<int64_t>((int64_t)nsm * per_sm, (P.nrep + VS_WPB - 1) / VS_WPB);
  const size_t nw = (size_t)grid * VS_WPB;                 // resident warps: one slice of each scratch array per warp
  const size_t bd = nw * (size_t)varsel_work_doubles(M.s0, M.p) * sizeof(double);
  const size_t bi = (P.mode == MODE_ESTIMATOR) ? nw * 2 * M.p * sizeof(int) : 0;
  const size_t bl = (P.mode == MODE_ESTIMATOR) ? nw * M.p * sizeof(long long) : 0;
  const size_t bi_off = (bd + 15) & ~(size_t)15, bl_off = (bi_off + bi + 15) & ~(size_t)15;
  if (scratch.bytes < bl_off + bl) { if (scratch.alloc(bl_off + bl) != cudaSuccess) { *err = "variable selection: scratch alloc failed"; return UBM_ERR_NOMEM; } }
  unsigned char* base = scratch.template as<unsigned char>();
  kern<<<grid, 32 * VS_WPB, smem, st>>>(M, P, reinterpret_cast<int*>(base + bi_off), reinterpret_cast<long long*>(base + bl_off),
                                        reinterpret_cast<double*>(base));
  *launches += 1;
  return UBM_OK;
}

}  // namespace ubm
