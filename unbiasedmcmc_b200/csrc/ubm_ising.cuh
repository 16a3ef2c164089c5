// ubm_ising.cuh — C4: 32 x 32 periodic Ising model, parallel tempering, coupled single-site Gibbs sweeps.
//
// Replaces, fused into one persistent kernel per launch:
//   drivers   R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//             (= ising_unbiased_estimator, R/ising.R:148-240, whose lag-1 weights are the generic ones)
//   kernels   R/ising.R:45-77 (PT single), :87-141 (PT coupled), :30-36 (rinit)
//   numerics  src/ising.cpp:43-62 / :67-92 (systematic-scan sweeps), :11-35 (natural statistic)
//
// Design (B200): one CTA per replicate, ONE WARP PER TEMPERATURE.  A lattice is 32 x 32 spins = 32 words: lane i
// keeps row i of chain 1 and of chain 2 in two registers for the whole replicate (no shared memory, no HBM).
// The reference sweep is an in-place sequential scan, not a checkerboard; it is reproduced EXACTLY by an
// anti-diagonal wavefront: site (i, j) is updated at step i + j, when rows i-1 / column j-1 already hold new
// values and row i+1 / column j+1 still hold old ones (periodic wrap-arounds fall on the right side too) —
// 63 warp steps per sweep, neighbour rows through warp shuffles.  The uniform of site (i, j) is draw number
// 32 i + j of the sweep, shared by both chains; `U < p` is evaluated as an exact integer compare of the 32-bit
// Philox word against ceil(p 2^32 - 1/2).  Natural statistics are popcounts; estimator sums are exact integers.
// The 16 warps only meet at a CTA barrier for the meeting test, and for the (rare) swap moves.
#pragma once
#include <string>
#include <vector>
#include "ubm_common.cuh"

namespace ubm {

#define ISING_MAXN 32

struct IsingDevice {
  int N;
  double pswap;
  double betas[ISING_MAXN];
  double probas[ISING_MAXN][5];            // for the tape path (FP64 compare, as the reference)
  unsigned long long thr[ISING_MAXN][5];   // Philox path: U < p  <=>  word < thr
};

inline void ising_pack_host(const ubm_model_ising_pt* s, IsingDevice* M) {
  M->N = s->nchains; M->pswap = s->proba_swapmove;
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
    }
  }
}

// natural statistic of the lattice whose row `lane` is `r`: sum_{i~j} x_i x_j over right and down neighbours
// (src/ising.cpp:11-35); x y = 1 - 2 (bit_x xor bit_y)
__device__ __forceinline__ int ising_sum_rows(unsigned r, int lane) {
  const unsigned right = (r >> 1) | (r << 31);                       // bit j <- bit j+1 (periodic)
  const unsigned down = __shfl_sync(0xffffffffu, r, (lane + 1) & 31);   // row i+1
  const int dis = __popc(r ^ right) + __popc(r ^ down);
  return 2048 - 2 * (int)__reduce_add_sync(0xffffffffu, (unsigned)dis);
}

template <bool TAPE>
struct IsingU {   // the replicate's U stream, addressed by absolute index; every warp tracks the same counter
  const RunParams& P;
  long long rep;
  uint64_t grep;
  bool bad;
  __device__ IsingU(const RunParams& p, long long r) : P(p), rep(r), grep((uint64_t)(p.rep_offset + r)), bad(false) {}
  __device__ __forceinline__ ubm_u32x4 block(unsigned long long b) const { return ubm_stream_block(P.seed, grep, UBM_STREAM_U, b); }
  __device__ __forceinline__ double u(unsigned long long i) {
    if (TAPE) {
      if (!P.tape_u || (long long)i >= P.len_u) { bad = true; return 0.5; }
      return P.tape_u[rep * P.len_u + (long long)i];
    }
    return ubm_philox_u(P.seed, grep, i);
  }
};

__device__ __forceinline__ unsigned pick_word(const ubm_u32x4& b, unsigned s) {
  return (s == 0) ? b.w[0] : (s == 1) ? b.w[1] : (s == 2) ? b.w[2] : b.w[3];
}

// One systematic-scan sweep of this warp's lattice (pair): wavefront over anti-diagonals.  `base` = index of the
// sweep's first uniform in the replicate's U stream.  COUPLED: both rows, same uniform (src/ising.cpp:67-92).
template <bool TAPE, bool COUPLED>
__device__ __forceinline__ void ising_sweep(unsigned& r1, unsigned& r2, int lane, unsigned long long base,
                                            const unsigned long long (&thr)[5], const double (&prob)[5], IsingU<TAPE>& U) {
  const unsigned long long B = base + 32ull * lane;     // index of site (lane, 0)
  ubm_u32x4 cur, nxt;
  if (!TAPE) { cur = U.block(B >> 2); nxt = cur; }
#pragma unroll 1
  for (int d = 0; d < 63; ++d) {
    const int j = d - lane;
    const unsigned up1 = __shfl_sync(0xffffffffu, r1, (lane + 1) & 31);
    const unsigned dn1 = __shfl_sync(0xffffffffu, r1, (lane + 31) & 31);
    unsigned up2 = 0, dn2 = 0;
    if (COUPLED) {
      up2 = __shfl_sync(0xffffffffu, r2, (lane + 1) & 31);
      dn2 = __shfl_sync(0xffffffffu, r2, (lane + 31) & 31);
    }
    const int jc = j < 0 ? 0 : (j > 31 ? 31 : j);
    const unsigned long long idx = B + (unsigned)jc;
    unsigned w = 0;
    double ud = 0.0;
    if (TAPE) {
      if (j >= 0 && j < 32) ud = U.u(idx);
    } else {
      // block bookkeeping, executed by all lanes together: a lane's next block is fetched at d % 4 == 0
      if (j > 0 && j < 32 && (idx & 3ull) == 0) cur = nxt;
      if ((d & 3) == 0) nxt = U.block((idx >> 2) + 1);
      w = pick_word(cur, (unsigned)(idx & 3ull));
    }
    if (j >= 0 && j < 32) {
      const int jr = (j + 1) & 31, jl = (j + 31) & 31;
      const int c1 = (int)((up1 >> j) & 1u) + (int)((dn1 >> j) & 1u) + (int)((r1 >> jr) & 1u) + (int)((r1 >> jl) & 1u);
      // neighbour sum s = 2 c - 4, table index (s + 4) / 2 = c   (src/ising.cpp:57)
      bool b1;
      if (TAPE) b1 = ud < ((c1 == 0) ? prob[0] : (c1 == 1) ? prob[1] : (c1 == 2) ? prob[2] : (c1 == 3) ? prob[3] : prob[4]);
      else b1 = (unsigned long long)w < ((c1 == 0) ? thr[0] : (c1 == 1) ? thr[1] : (c1 == 2) ? thr[2] : (c1 == 3) ? thr[3] : thr[4]);
      r1 = (r1 & ~(1u << j)) | ((b1 ? 1u : 0u) << j);
      if (COUPLED) {
        const int c2 = (int)((up2 >> j) & 1u) + (int)((dn2 >> j) & 1u) + (int)((r2 >> jr) & 1u) + (int)((r2 >> jl) & 1u);
        bool b2;
        if (TAPE) b2 = ud < ((c2 == 0) ? prob[0] : (c2 == 1) ? prob[1] : (c2 == 2) ? prob[2] : (c2 == 3) ? prob[3] : prob[4]);
        else b2 = (unsigned long long)w < ((c2 == 0) ? thr[0] : (c2 == 1) ? thr[1] : (c2 == 2) ? thr[2] : (c2 == 3) ? thr[3] : thr[4]);
        r2 = (r2 & ~(1u << j)) | ((b2 ? 1u : 0u) << j);
      }
    }
  }
}

struct IsingShared {
  int sums1[ISING_MAXN], sums2[ISING_MAXN];
  int perm1[ISING_MAXN], perm2[ISING_MAXN];
  unsigned lat[2][ISING_MAXN][32];
  long long rep;
};

// swap move (R/ising.R:48-68 / :97-126): sequential over adjacent pairs, one shared uniform per pair
template <bool TAPE>
__device__ __forceinline__ void ising_swap_move(const IsingDevice& M, IsingShared& sh, unsigned& r1, unsigned& r2,
                                                int lane, int warp, unsigned long long base, bool do2, IsingU<TAPE>& U) {
  const int N = M.N;
  sh.lat[0][warp][lane] = r1;
  sh.lat[1][warp][lane] = r2;
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c < N; ++c) { sh.perm1[c] = c; sh.perm2[c] = c; }
      for (int c = 0; c + 1 < N; ++c) {
        const double deltabeta = M.betas[c] - M.betas[c + 1];
        const double logu = ubm_log(U.u(base + c));
        const double lp1 = deltabeta * ((double)sh.sums1[c + 1] - (double)sh.sums1[c]);
        if (logu < lp1) {
          int t = sh.sums1[c]; sh.sums1[c] = sh.sums1[c + 1]; sh.sums1[c + 1] = t;
          t = sh.perm1[c]; sh.perm1[c] = sh.perm1[c + 1]; sh.perm1[c + 1] = t;
        }
        if (do2) {
          const double lp2 = deltabeta * ((double)sh.sums2[c + 1] - (double)sh.sums2[c]);
          if (logu < lp2) {
            int t = sh.sums2[c]; sh.sums2[c] = sh.sums2[c + 1]; sh.sums2[c + 1] = t;
            t = sh.perm2[c]; sh.perm2[c] = sh.perm2[c + 1]; sh.perm2[c + 1] = t;
          }
        }
      }
    }
  }
  __syncthreads();
  r1 = sh.lat[0][sh.perm1[warp]][lane];
  if (do2) r2 = sh.lat[1][sh.perm2[warp]][lane];
  __syncthreads();
}

// One CTA per replicate, warp c = temperature c.
template <bool TAPE>
#ifndef ISING_MAXT
#define ISING_MAXT 1024      // threads per CTA the kernel is compiled for (32 temperatures)
#endif
#ifndef ISING_MINB
#define ISING_MINB 1
#endif
__global__ void __launch_bounds__(ISING_MAXT, ISING_MINB) ising_driver_kernel(const IsingDevice M, const RunParams P) {
  __shared__ IsingShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = M.N;
  unsigned long long thr[5];
  double prob[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) { thr[q] = M.thr[warp][q]; prob[q] = M.probas[warp][q]; }
  const double denom = (double)(P.m - P.k + 1);

  for (;;) {
    if (tid == 0) {
      const unsigned long long got = atomicAdd(P.work_counter, 1ull);
      sh.rep = (got < (unsigned long long)P.nrep) ? (long long)got : -1;
    }
    __syncthreads();
    const long long rep = sh.rep;
    __syncthreads();
    if (rep < 0) break;
    IsingU<TAPE> U(P, rep);
    // ---- rinit (R/ising.R:4-8,30-36): lattice c of chain 1 from uniforms [1024 c, 1024 c + 1024), then chain 2;
    // filled column-major: draw idx -> row idx % 32, column idx / 32; spin +1 iff U >= 0.5
    unsigned r1 = 0, r2 = 0;
    for (int j = 0; j < 32; ++j) {
      const unsigned long long i1 = 1024ull * warp + 32ull * j + lane;
      const unsigned long long i2 = 1024ull * (N + warp) + 32ull * j + lane;
      r1 |= (U.u(i1) >= 0.5 ? 1u : 0u) << j;
      r2 |= (U.u(i2) >= 0.5 ? 1u : 0u) << j;
    }
    unsigned long long iu = 2048ull * N;
    int s1 = ising_sum_rows(r1, lane), s2 = ising_sum_rows(r2, lane);
    long long mc = 0, co = 0;           // exact integer estimator sums of this temperature
    long long time = 0, tau = 0;
    bool met = false;
    if (P.mode == MODE_ESTIMATOR && P.k == 0) mc = s1;
    if (P.mode == MODE_CHAINS && lane == 0) {
      P.samples1[(rep * P.stride) * N + warp] = (double)s1;
      P.samples2[(rep * P.stride) * N + warp] = (double)s2;
    }
    // ---- time loop (generic drivers); every warp executes the same control flow
    for (;;) {
      bool go;
      if (time < P.lag) go = true;
      else {
        const bool capped = (P.max_it > 0 && time >= P.max_it) || U.bad;
        if (P.mode == MODE_MEETING) go = !met && !capped;
        else go = (!met || time < (tau > P.m ? tau : P.m)) && !capped;
      }
      // U.bad is per-thread: make the decision CTA-uniform
      if (TAPE) go = !__syncthreads_or(go ? 0 : 1);
      if (!go) break;
      time += 1;
      const bool coupled_step = (time > P.lag) && !met;
      const bool do2 = coupled_step;
      const double u_iter = U.u(iu);                               // R/ising.R:46,92
      iu += 1;
      if (u_iter < M.pswap) {
        if (lane == 0) { sh.sums1[warp] = s1; sh.sums2[warp] = s2; }
        ising_swap_move<TAPE>(M, sh, r1, r2, lane, warp, iu, do2, U);
        iu += (unsigned long long)(N - 1);
      } else {
        const unsigned long long base = iu + 1024ull * warp;
        if (do2) ising_sweep<TAPE, true>(r1, r2, lane, base, thr, prob, U);
        else ising_sweep<TAPE, false>(r1, r2, lane, base, thr, prob, U);
        iu += 1024ull * N;
      }
      s1 = ising_sum_rows(r1, lane);                               // R/ising.R:75,135-136
      if (do2) {
        s2 = ising_sum_rows(r2, lane);
        const int eq = __all_sync(0xffffffffu, r1 == r2) ? 1 : 0;
        if (__syncthreads_and(eq)) { met = true; tau = time; }     // R/ising.R:215
      }
      const int y = (time > P.lag && !coupled_step) ? s1 : s2;     // after meeting chain 2 copies chain 1
      if (P.mode == MODE_ESTIMATOR) {
        const bool in_window = (time <= P.lag) ? (time >= P.k) : (P.k <= time && time <= P.m);
        if (in_window) mc += s1;
        if ((time >= P.k + P.lag) && (coupled_step || time == P.lag)) co += corr_weight(time, P.k, P.m, P.lag) * (long long)(s1 - y);
      } else if (P.mode == MODE_CHAINS && lane == 0) {
        P.samples1[(rep * P.stride + time) * N + warp] = (double)s1;
        if (time > P.lag) P.samples2[(rep * P.stride + (time - P.lag)) * N + warp] = (double)y;
      }
    }
    // ---- results
    if (lane == 0) {
      if (warp == 0) {
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
        P.uest[rep * N + warp] = ue / denom;
        if (P.mcmc) P.mcmc[rep * N + warp] = dmc / denom;
        if (P.corr) P.corr[rep * N + warp] = dco / denom;
      }
    }
  }
}

inline int ising_launch(const IsingDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st,
                        int64_t* launches, std::string* err) {
  if (P.nbreaks >= 2) { *err = "ising: histogram bins are not available for this model"; return UBM_ERR_UNSUPPORTED; }
  if (P.mode == MODE_ESTIMATOR && P.h_kind != UBM_H_IDENTITY) { *err = "ising: h must be the identity on the natural statistics"; return UBM_ERR_UNSUPPORTED; }
  const int threads = 32 * M.N;
  auto kern = tape ? ising_driver_kernel<true> : ising_driver_kernel<false>;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
  if (e != cudaSuccess || per_sm < 1) { *err = "ising: occupancy query failed"; return UBM_ERR_CUDA; }
  int grid = (int)std::min<int64_t>((int64_t)nsm * per_sm, P.nrep);
  kern<<<grid, threads, 0, st>>>(M, P);
  *launches += 1;
  return UBM_OK;
}

}  // namespace ubm
