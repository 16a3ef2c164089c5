// ubm_ising.cuh — C4: 32 x 32 periodic Ising model, parallel tempering, coupled single-site Gibbs sweeps.
//
// Replaces, fused into one persistent kernel per launch:
//   drivers   R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//             (= ising_unbiased_estimator, R/ising.R:148-240, whose lag-1 weights are the generic ones)
//   kernels   R/ising.R:45-77 (PT single), :87-141 (PT coupled), :30-36 (rinit)
//   numerics  src/ising.cpp:43-62 / :67-92 (systematic-scan sweeps), :11-35 (natural statistic)
//
// Design (B200): ONE LANE PER LATTICE PAIR.  Lane l of a warp owns temperature l % N of replicate l / N of the warp's
// group of 32 / N replicates; its two lattices (chain 1, chain 2) are 2 x 32 row words in shared memory, laid out
// [row][lane] so that every access is conflict free and no lane ever touches another lane's words.  The reference
// sweep is an in-place sequential scan, and it is reproduced EXACTLY, but a whole row at a time:
//   * the 32 uniforms of a row are 8 Philox blocks; comparing each word with the 5 acceptance thresholds of the
//     lane's temperature (as the carry out of word + (2^32 - threshold)) gives five bit planes G_q (bit j: "site j becomes +1 if q of its neighbours are +1") that
//     both chains share;
//   * the vertical neighbours (new row above, old row below) are fixed during the row, so P_m = G_{up + down + m}
//     (m = left + right) are three more planes by bitwise multiplexing; the right neighbour is old (except for
//     the last column), which folds P into M0 / M1 = the new spin when the left neighbour is -1 / +1;
//   * what remains, new_j = new_{j-1} ? M1_j : M0_j, is a carry chain: with M0 a subset of M1 (thresholds increase
//     with the neighbour count, i.e. beta >= 0) the 32 new spins are the carries of ONE integer addition M1 + M0 +
//     carry-in.  (Negative beta: a 32-step serial recurrence on the same two planes.)
// A sweep of a lattice pair is therefore 32 x (8 Philox blocks + 160 compares + about 60 logic operations) with
// no shuffles, no idle lanes and no sequential dependence inside a row; the cost is the random numbers.
// Natural statistics are popcounts; estimator sums are exact integers; swap moves permute shared-memory columns.
#pragma once
#include <string>
#include <vector>
#include "ubm_common.cuh"

namespace ubm {

#define ISING_MAXN 32
#ifndef ISING_WPB
#define ISING_WPB 2          // warps per CTA (independent of each other)
#endif
#ifndef ISING_MINB
#define ISING_MINB 8         // CTAs per SM asked of the register allocator
#endif

struct IsingDevice {
  int N, monotone;                         // monotone: thresholds / probabilities nondecreasing in the neighbour count
  double pswap;
  double betas[ISING_MAXN];
  double probas[ISING_MAXN][5];            // tape path (FP64 compare, as the reference)
  unsigned long long thr[ISING_MAXN][5];   // Philox path: U < p  <=>  word < thr
};

struct IsingKeys { uint32_t k0[10], k1[10]; };   // Philox round keys of the run's seed (kernel-uniform: constant-bank operands)

inline void ising_pack_host(const ubm_model_ising_pt* s, IsingDevice* M) {
  M->N = s->nchains; M->pswap = s->proba_swapmove; M->monotone = 1;
  for (int c = 0; c < s->nchains; ++c) {
    M->betas[c] = s->betas[c];
    for (int q = 0; q < 5; ++q) {                       // R/ising.R:151-152
      double ss = -4.0 + 2.0 * q;
      double a = ubm_exp(ss * s->betas[c]), b = ubm_exp(-ss * s->betas[c]);
      double p = a / (a + b);
      M->probas[c][q] = p;
      // (w + 1/2) 2^-32 < p  <=>  w < p 2^32 - 1/2 (both sides exact in binary64)  <=>  w < ceil(p 2^32 - 1/2)
      double t = ceil(p * 4294967296.0 - 0.5);
      M->thr[c][q] = (t <= 0.0) ? 0ull : (unsigned long long)t;
      if (q > 0 && (M->probas[c][q] < M->probas[c][q - 1] || M->thr[c][q] < M->thr[c][q - 1])) M->monotone = 0;
    }
  }
}
inline void ising_keys_host(uint64_t seed, IsingKeys* K) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { K->k0[r] = k0; K->k1[r] = k1; k0 += UBM_PHILOX_W0; k1 += UBM_PHILOX_W1; }
}

// ubm_philox4x32_10 (include/ubm_numerics.h) with the key schedule precomputed
__device__ __forceinline__ void ising_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const IsingKeys& K, uint32_t (&w)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)UBM_PHILOX_M0 * c0, p1 = (unsigned long long)UBM_PHILOX_M1 * c2;   // IMAD.WIDE
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r]; c1 = (uint32_t)p1; c2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r]; c3 = (uint32_t)p0;
  }
  w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

template <bool TAPE>
struct IsingU {   // the replicate's U stream, addressed by absolute index
  const RunParams& P;
  long long rep;
  uint64_t grep;
  bool bad, mute;     // mute: this lane is only shadowing the warp's control flow (tape reads are skipped)
  __device__ IsingU(const RunParams& p, long long r) : P(p), rep(r), grep((uint64_t)(p.rep_offset + r)), bad(false), mute(false) {}
  __device__ __forceinline__ double u(unsigned long long i) {
    if (TAPE) {
      if (mute) return 0.5;
      if (!P.tape_u || (long long)i >= P.len_u) { bad = true; return 0.5; }
      return P.tape_u[rep * P.len_u + (long long)i];
    }
    return ubm_philox_u(P.seed, grep, i);
  }
};

struct IsingLane {          // per-lane constants of a replicate group
  unsigned C[5], never[5];  // Philox path: the carry out of word + C, C = 2^32 - thr (mod 2^32), is [word >= thr]; thr = 0: never +1
  double prob[5];           // tape path
};

// acc = 2 acc + [w + c carries]: one alu-pipe add for the carry, one fma-pipe multiply-add (IMAD.X) to shift it in.  The two
// pipes issue one warp instruction every other cycle each, and the alu pipe is this kernel's bound (ncu: alu 64%, fma 26% with a
// compare + predicated OR), so the plane building is split across them (measured +15%).  A compare done entirely on the fma
// pipe (the high word of IMAD.WIDE w * 1 + c) is undone by ptxas, which turns it back into the add.
__device__ __forceinline__ void ising_shift_in_carry(unsigned& acc, uint32_t w, unsigned c) {
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\tmadc.lo.u32 %0, %0, 2, 0;\n\t}" : "+r"(acc) : "r"(w), "r"(c));
}
__device__ __forceinline__ unsigned ising_mux(unsigned s, unsigned a, unsigned b) { return (s & a) | (~s & b); }   // s ? a : b, bitwise

// The five threshold planes of one row: bit j of G[q] = [U_j < p_q].  `idx0` = index of the row's first uniform in the
// replicate's U stream (a multiple of 4 on the Philox path: sweeps start on a block boundary).
template <bool TAPE>
__device__ __forceinline__ void ising_row_planes(unsigned (&G)[5], const IsingLane& LN, const IsingKeys& K, IsingU<TAPE>& U,
                                                 unsigned long long idx0) {
#pragma unroll
  for (int q = 0; q < 5; ++q) G[q] = 0u;
  if (TAPE) {
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const double ud = U.u(idx0 + (unsigned)j);
#pragma unroll
      for (int q = 0; q < 5; ++q) G[q] |= (ud < LN.prob[q]) ? (1u << j) : 0u;
    }
  } else {
    const unsigned long long b0 = idx0 >> 2;
    const uint32_t r0 = (uint32_t)U.grep, r1 = (uint32_t)(U.grep >> 32);
    // columns 31 .. 0, so that the planes built by shifting (2 acc + bit) end with column j in bit j
#pragma unroll
    for (int b = 7; b >= 0; --b) {
      const unsigned long long blk = b0 + (unsigned)b;
      uint32_t w[4];
      ising_philox((uint32_t)blk, (UBM_STREAM_U << 28) | ((uint32_t)(blk >> 32) & 0x0FFFFFFFu), r0, r1, K, w);
#pragma unroll
      for (int k = 3; k >= 0; --k) {
#pragma unroll
        for (int q = 0; q < 5; ++q) ising_shift_in_carry(G[q], w[k], LN.C[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) G[q] = ~(G[q] | LN.never[q]);     // [w >= thr] -> [w < thr]
  }
}

// The in-place sequential scan of one row (src/ising.cpp:47-60): `cur` old row, `up` the row above (already new, or
// the old last row for row 0), `dn` the row below (old, or the new row 0 for the last row).  Returns the new row.
template <bool MONO>
__device__ __forceinline__ unsigned ising_row_scan(unsigned cur, unsigned up, unsigned dn, const unsigned (&G)[5]) {
  const unsigned H0 = ising_mux(dn, G[1], G[0]), H1 = ising_mux(dn, G[2], G[1]);
  const unsigned H2 = ising_mux(dn, G[3], G[2]), H3 = ising_mux(dn, G[4], G[3]);
  const unsigned P0 = ising_mux(up, H1, H0), P1 = ising_mux(up, H2, H1), P2 = ising_mux(up, H3, H2);   // by left + right
  const unsigned cin = cur >> 31;                        // left neighbour of column 0: the old column 31
  unsigned R = __funnelshift_r(cur, cur, 1);             // bit j = old column j + 1 (bit 31 fixed below)
  unsigned M0 = ising_mux(R, P1, P0), M1 = ising_mux(R, P2, P1);
  const unsigned new0 = (cin ? M1 : M0) & 1u;            // column 0 first: it is the right neighbour of column 31
  R = (R & 0x7fffffffu) | (new0 << 31);
  M0 = ising_mux(R, P1, P0); M1 = ising_mux(R, P2, P1);
  if (MONO) {
    // new_j = M0_j | (M1_j & new_{j-1}): generate M0, propagate M1 & ~M0  ->  the carries of M1 + M0 + cin
    const unsigned S = M1 + M0 + cin;
    const unsigned Cin = S ^ M1 ^ M0;                    // carry INTO each bit = new_{j-1}
    return (M1 & M0) | (Cin & (M1 ^ M0));                // carry OUT of each bit = new_j
  } else {
    unsigned nw = 0u, prev = cin;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      prev = ((prev ? M1 : M0) >> j) & 1u;
      nw |= prev << j;
    }
    return nw;
  }
}

// One systematic-scan sweep of this lane's lattice (pair).  `base` = index of the lattice's first uniform.  Rows are
// written back only where `st1` / `st2` (lanes of other replicates of the warp may be doing something else).
// Returns through dis1 / dis2 the number of disagreeing neighbour pairs, through diff the OR of chain1 ^ chain2.
template <bool TAPE, bool COUPLED, bool MONO>
__device__ __forceinline__ void ising_sweep_lane(unsigned* __restrict__ L1, unsigned* __restrict__ L2, const IsingLane& LN,
                                                 const IsingKeys& K, IsingU<TAPE>& U, unsigned long long base, bool st1, bool st2,
                                                 int& dis1, int& dis2, unsigned& diff) {
  unsigned up1 = L1[31 * 32], up2 = COUPLED ? L2[31 * 32] : 0u;
  unsigned first1 = 0u, first2 = 0u;
  int d1 = 0, d2 = 0;
  unsigned df = 0u;
#pragma unroll 1
  for (int i = 0; i < 32; ++i) {
    unsigned G[5];
    ising_row_planes<TAPE>(G, LN, K, U, base + 32ull * (unsigned)i);
    const int in = ((i + 1) & 31) * 32;
    const unsigned n1 = ising_row_scan<MONO>(L1[i * 32], up1, L1[in], G);
    if (st1) L1[i * 32] = n1;
    d1 += __popc(n1 ^ __funnelshift_r(n1, n1, 1));
    if (i == 0) first1 = n1; else d1 += __popc(n1 ^ up1);
    up1 = n1;
    if (COUPLED) {
      const unsigned n2 = ising_row_scan<MONO>(L2[i * 32], up2, L2[in], G);
      if (st2) L2[i * 32] = n2;
      d2 += __popc(n2 ^ __funnelshift_r(n2, n2, 1));
      if (i == 0) first2 = n2; else d2 += __popc(n2 ^ up2);
      up2 = n2;
      df |= n1 ^ n2;
    }
  }
  dis1 = d1 + __popc(up1 ^ first1);
  if (COUPLED) { dis2 = d2 + __popc(up2 ^ first2); diff = df; }
}

// natural statistic sum_{i~j} x_i x_j over right and down neighbours (src/ising.cpp:11-35); x y = 1 - 2 (bit_x xor bit_y)
__device__ __forceinline__ int ising_stat_lane(const unsigned* __restrict__ L) {
  int dis = 0;
  unsigned prev = L[31 * 32];
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    const unsigned r = L[i * 32];
    dis += __popc(r ^ __funnelshift_r(r, r, 1)) + __popc(r ^ prev);
    prev = r;
  }
  return 2048 - 2 * dis;
}

// rinit (R/ising.R:4-8,30-36): the lattice from uniforms [first, first + 1024), filled column-major (draw idx -> row
// idx % 32, column idx / 32), spin +1 iff U >= 0.5 (the top bit of the Philox word)
template <bool TAPE>
__device__ __forceinline__ void ising_rinit_lane(unsigned* __restrict__ L, const IsingKeys& K, IsingU<TAPE>& U, unsigned long long first) {
  unsigned rows[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) rows[i] = 0u;
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    if (TAPE) {
#pragma unroll
      for (int i = 0; i < 32; ++i) rows[i] |= (U.u(first + 32ull * (unsigned)j + (unsigned)i) >= 0.5 ? 1u : 0u) << j;
    } else {
      const unsigned long long b0 = (first >> 2) + 8ull * (unsigned)j;
      const uint32_t r0 = (uint32_t)U.grep, r1 = (uint32_t)(U.grep >> 32);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const unsigned long long blk = b0 + (unsigned)b;
        uint32_t w[4];
        ising_philox((uint32_t)blk, (UBM_STREAM_U << 28) | ((uint32_t)(blk >> 32) & 0x0FFFFFFFu), r0, r1, K, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) rows[4 * b + k] |= (w[k] >> 31) << j;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) L[i * 32] = rows[i];
}

struct IsingShared {
  unsigned lat[ISING_WPB][2][32 * 32];     // [warp][chain][row * 32 + lane]
  int sums[ISING_WPB][2][32];
  int perm[ISING_WPB][2][32];
  unsigned long long thr[ISING_MAXN * 5];
  double prob[ISING_MAXN * 5];
  double betas[ISING_MAXN];
};

template <bool TAPE, bool MONO>
__global__ void __launch_bounds__(32 * ISING_WPB, ISING_MINB) ising_driver_kernel(const IsingDevice M, const RunParams P, const IsingKeys K) {
  __shared__ IsingShared sh;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int N = M.N, R = 32 / N;
  const int slot = lane / N, c = lane - slot * N;
  const bool lane_ok = slot < R;
  const unsigned slotmask = lane_ok ? ((N == 32) ? FULL : (((1u << N) - 1u) << (slot * N))) : (1u << lane);
  for (int t = threadIdx.x; t < N * 5; t += blockDim.x) { sh.thr[t] = M.thr[t / 5][t % 5]; sh.prob[t] = M.probas[t / 5][t % 5]; }
  for (int t = threadIdx.x; t < N; t += blockDim.x) sh.betas[t] = M.betas[t];
  __syncthreads();
  IsingLane LN;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const unsigned long long t = sh.thr[c * 5 + q];
    LN.C[q] = (unsigned)(4294967296ull - t);      // thr = 2^32: C = 0, never a carry, always +1
    LN.never[q] = (t == 0ull) ? FULL : 0u;
    LN.prob[q] = sh.prob[c * 5 + q];
  }
  unsigned* L1 = &sh.lat[wib][0][lane];
  unsigned* L2 = &sh.lat[wib][1][lane];
  int* sums1 = sh.sums[wib][0]; int* sums2 = sh.sums[wib][1];
  int* perm1 = sh.perm[wib][0]; int* perm2 = sh.perm[wib][1];
  const double denom = (double)(P.m - P.k + 1);
  const int nbins = (P.mode == MODE_ESTIMATOR && P.nbreaks >= 2) ? P.nbreaks - 1 : 0;
  double bin_b0 = 0.0, bin_iw = 0.0;
  if (nbins) { bin_b0 = P.breaks[0]; bin_iw = (double)nbins / (P.breaks[nbins] - P.breaks[0]); }

  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(P.work_counter, (unsigned long long)R);
    got = __shfl_sync(FULL, got, 0);
    if (got >= (unsigned long long)P.nrep) break;
    const bool valid = lane_ok && (long long)got + slot < P.nrep;
    const long long rep = valid ? (long long)got + slot : (long long)got;   // idle lanes shadow replicate `got`, without writing
    IsingU<TAPE> U(P, rep);
    ising_rinit_lane<TAPE>(L1, K, U, 1024ull * (unsigned)c);
    ising_rinit_lane<TAPE>(L2, K, U, 1024ull * (unsigned)(N + c));
    unsigned long long iu = 2048ull * N;
    int s1 = ising_stat_lane(L1), s2 = ising_stat_lane(L2);
    long long mc = 0, co = 0;           // exact integer estimator sums of this temperature
    long long time = 0, tau = 0;
    bool met = false, done = !valid;
    long long* bn = (nbins && valid && c == P.component) ? P.bin_num + rep * nbins : nullptr;
    if (P.mode == MODE_ESTIMATOR && P.k == 0) {
      mc = s1;
      if (bn) { const int b = bin_of((double)s1, P.breaks, P.nbreaks, bin_b0, bin_iw); if (b >= 0) bn[b] += 1; }
    }
    if (P.mode == MODE_CHAINS && valid) {
      P.samples1[(rep * P.stride) * N + c] = (double)s1;
      P.samples2[(rep * P.stride) * N + c] = (double)s2;
    }
    // ---- time loop (generic drivers); the replicates of a warp advance together, each with its own stopping rule
    for (;;) {
      if (TAPE) { if (__ballot_sync(FULL, U.bad) & slotmask) U.bad = true; }
      bool go = false;
      if (!done) {
        if (time < P.lag) go = true;
        else {
          const bool capped = (P.max_it > 0 && time >= P.max_it) || U.bad;
          if (P.mode == MODE_MEETING) go = !met && !capped;
          else go = (!met || time < (tau > P.m ? tau : P.m)) && !capped;
        }
        if (!go) done = true;
      }
      if (!__any_sync(FULL, go)) break;
      if (go) time += 1;
      const bool coupled_step = go && (time > P.lag) && !met;
      double u_iter = 1.0;
      if (go) { u_iter = U.u(iu); iu += 1; }                        // R/ising.R:46,92
      const bool do_swap = go && (u_iter < M.pswap);
      const bool do_sweep = go && !do_swap;
      unsigned diff = 0u;
      if (__any_sync(FULL, do_swap)) {
        // ---- swap move (R/ising.R:48-68 / :97-126): sequential over adjacent pairs, one shared uniform per pair
        sums1[lane] = s1; sums2[lane] = s2; perm1[lane] = lane; perm2[lane] = lane;
        __syncwarp();
        if (do_swap && c == 0) {
          const int o = slot * N;
          unsigned long long* cnt = reinterpret_cast<unsigned long long*>(P.pt_counts);   // R/ising.R:48-49,52,60 / :89-91,95,108,118
          if (cnt) atomicAdd(cnt, 1ull);
          for (int cc = 0; cc + 1 < N; ++cc) {
            const double deltabeta = sh.betas[cc] - sh.betas[cc + 1];
            const double logu = ubm_log(U.u(iu + (unsigned)cc));
            const double lp1 = deltabeta * ((double)sums1[o + cc + 1] - (double)sums1[o + cc]);
            if (logu < lp1) {
              int t = sums1[o + cc]; sums1[o + cc] = sums1[o + cc + 1]; sums1[o + cc + 1] = t;
              t = perm1[o + cc]; perm1[o + cc] = perm1[o + cc + 1]; perm1[o + cc + 1] = t;
              if (cnt) atomicAdd(cnt + 1 + cc, 1ull);
            }
            if (coupled_step) {
              const double lp2 = deltabeta * ((double)sums2[o + cc + 1] - (double)sums2[o + cc]);
              if (logu < lp2) {
                int t = sums2[o + cc]; sums2[o + cc] = sums2[o + cc + 1]; sums2[o + cc + 1] = t;
                t = perm2[o + cc]; perm2[o + cc] = perm2[o + cc + 1]; perm2[o + cc + 1] = t;
                if (cnt) atomicAdd(cnt + N + cc, 1ull);
              }
            }
          }
        }
        __syncwarp();
        const int src1 = perm1[lane], src2 = perm2[lane];          // identity for lanes that do not swap
        const unsigned* B1 = &sh.lat[wib][0][0];
        const unsigned* B2 = &sh.lat[wib][1][0];
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          const unsigned v1 = B1[i * 32 + src1], v2 = B2[i * 32 + src2];
          __syncwarp();
          L1[i * 32] = v1; L2[i * 32] = v2;
          if (do_swap) diff |= v1 ^ v2;
          __syncwarp();
        }
        if (do_swap) { s1 = sums1[lane]; s2 = sums2[lane]; iu += (unsigned long long)(N - 1); }
        __syncwarp();
      }
      if (__any_sync(FULL, do_sweep)) {
        // Philox path: a sweep starts on a block boundary of the U stream (the numerics contract; tapes are dense)
        const unsigned long long sweep0 = TAPE ? iu : ((iu + 3ull) & ~3ull);
        const unsigned long long base = sweep0 + 1024ull * (unsigned)c;
        int dis1 = 0, dis2 = 0;
        unsigned df = 0u;
        U.mute = !do_sweep;
        if (__any_sync(FULL, do_sweep && coupled_step)) {
          ising_sweep_lane<TAPE, true, MONO>(L1, L2, LN, K, U, base, do_sweep, do_sweep && coupled_step, dis1, dis2, df);
          if (do_sweep && coupled_step) { s2 = 2048 - 2 * dis2; diff = df; }
        } else {
          ising_sweep_lane<TAPE, false, MONO>(L1, L2, LN, K, U, base, do_sweep, false, dis1, dis2, df);
        }
        U.mute = false;
        if (do_sweep) { s1 = 2048 - 2 * dis1; iu = sweep0 + 1024ull * (unsigned)N; }
      }
      if (__any_sync(FULL, coupled_step)) {
        const unsigned eq = __ballot_sync(FULL, diff == 0u);
        if (coupled_step && (eq & slotmask) == slotmask) { met = true; tau = time; }     // R/ising.R:215
      }
      if (go) {
        const int y = (time > P.lag && !coupled_step) ? s1 : s2;     // after meeting chain 2 copies chain 1
        if (P.mode == MODE_ESTIMATOR) {
          const bool in_window = (time <= P.lag) ? (time >= P.k) : (P.k <= time && time <= P.m);
          const bool add_corr = (time >= P.k + P.lag) && (coupled_step || time == P.lag);
          if (in_window) mc += s1;
          long long w = 0;
          if (add_corr) { w = corr_weight(time, P.k, P.m, P.lag); co += w * (long long)(s1 - y); }
          if (bn) {                                                   // estimator_bin_ numerators (src/estimator_bin.cpp:15-38)
            const int b1 = bin_of((double)s1, P.breaks, P.nbreaks, bin_b0, bin_iw);
            if (in_window && b1 >= 0) bn[b1] += 1;
            if (add_corr) {
              const int b2 = bin_of((double)y, P.breaks, P.nbreaks, bin_b0, bin_iw);
              if (b1 >= 0) bn[b1] += w;
              if (b2 >= 0) bn[b2] -= w;
            }
          }
        } else if (P.mode == MODE_CHAINS) {
          P.samples1[(rep * P.stride + time) * N + c] = (double)s1;
          if (time > P.lag) P.samples2[(rep * P.stride + (time - P.lag)) * N + c] = (double)y;
        }
      }
    }
    // ---- results
    if (valid) {
      if (c == 0) {
        double cost = dinf();
        if (met) {
          if (P.mode == MODE_MEETING) cost = (double)(P.lag + 2 * (tau - P.lag));
          else cost = (double)(P.lag + 2 * (tau - P.lag) + ((time - tau) > 0 ? (time - tau) : 0));
        }
        if (P.tau) P.tau[rep] = met ? (double)tau : dinf();
        if (P.cost) P.cost[rep] = cost;
        if (P.iter) P.iter[rep] = time;
        if (U.bad) atomicExch(P.status, UBM_ERR_TAPE);
      }
      if (P.mode == MODE_ESTIMATOR) {
        const double dmc = (double)mc, dco = (double)co;
        const double ue = dmc + dco;
        P.uest[rep * N + c] = ue / denom;
        if (P.mcmc) P.mcmc[rep * N + c] = dmc / denom;
        if (P.corr) P.corr[rep * N + c] = dco / denom;
      }
    }
    __syncwarp();
  }
}

inline int ising_launch(const IsingDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st,
                        int64_t* launches, std::string* err) {
  if (P.mode == MODE_ESTIMATOR && P.h_kind != UBM_H_IDENTITY) { *err = "ising: h must be the identity on the natural statistics"; return UBM_ERR_UNSUPPORTED; }
  IsingKeys K;
  ising_keys_host(P.seed, &K);
  void (*kern)(const IsingDevice, const RunParams, const IsingKeys);
  if (tape) kern = M.monotone ? ising_driver_kernel<true, true> : ising_driver_kernel<true, false>;
  else kern = M.monotone ? ising_driver_kernel<false, true> : ising_driver_kernel<false, false>;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * ISING_WPB, 0);
  if (e != cudaSuccess || per_sm < 1) { *err = "ising: occupancy query failed"; return UBM_ERR_CUDA; }
  const int per_group = 32 / M.N;                                   // replicates of one warp
  const int64_t groups = (P.nrep + per_group - 1) / per_group;
  const int grid = (int)std::min<int64_t>((int64_t)nsm * per_sm, (groups + ISING_WPB - 1) / ISING_WPB);
  kern<<<grid, 32 * ISING_WPB, 0, st>>>(M, P, K);
  *launches += 1;
  return UBM_OK;
}

}  // namespace ubm
