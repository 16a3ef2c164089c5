// ubm_mvn_single.cuh — C2, second pass: the single-chain steps of a replicate after its chains have met.
//
// A coupled replicate spends its first tau steps (about 1000 at d = 100) in the coupled kernel and the remaining
// max(m, tau) - tau steps (about 9000 at m = 10000) as ONE chain: R/unbiasedestimator.R:83-91 (after meeting the second
// chain is a copy), i.e. the single RW-MH kernel of R/get_mh_kernels.R:29-40 on the target of src/mvnorm.cpp:71-86, with the
// MCMC part of H_{k:m} and of estimator_bin_ accumulated on the fly.  ubm_mvn.cuh hands such a replicate over at its
// meeting time (the record sits in the replicate's own output rows, RunParams::ho_*); this kernel continues it to its horizon and writes the results.
//
// Why a second kernel: a step of the coupled kernel is a chain of latency-bound phases (per-slot section, two GEMM phases)
// whose length hardly depends on how many slots take part, and shared memory limits it to 16 slots (five rows per slot
// next to the three factors).  A single chain needs three rows per slot and two factors, so 32 slots fit: the same phases,
// twice the replicates per step.  Arithmetic, draw addressing and summation orders are those of the single-chain steps of
// ubm_mvn.cuh (the same device functions), so the result of a replicate does not depend on where its steps ran.
#pragma once
#include "ubm_mvn.cuh"

namespace ubm {

#define MVS_R 32         // slots per CTA
enum { MS_IDLE = 0, MS_RESUME = 1, MS_RUN = 2 };

struct MvsSlots {
  long long rep[MVS_R], time[MVS_R], tau[MVS_R];
  unsigned long long iu[MVS_R], in[MVS_R], pu[MVS_R];
  double pdf[MVS_R];
  int phase[MVS_R];
  int exhausted;
  int draw_next;      // next slot whose normals of the coming step are to be generated (a queue shared by all 16 warps)
};

__host__ __device__ inline size_t mvs_smem_bytes(int packLen, int DPS) {
  return (size_t)(2 * packLen + DPS + 3 * MVS_R * DPS + MVS_R * 32 + DPS + 8 * MVS_R) * sizeof(double) + sizeof(MvsSlots) + 16;
}

template <bool ZERO_MEAN, int NH>
__global__ void __launch_bounds__(MVN_T) mvn_single_kernel(const MvnDevice M, const RunParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, T = MVN_T, lane = tid & 31, warp = tid >> 5;
  const int d = M.d, DPS = M.DPS, nCG = M.nCG;
  double* p = reinterpret_cast<double*>(smem_raw);
  double* packB = p; p += M.packLen;
  double* packC = p; p += M.packLen;
  double* MEAN = p; p += DPS;
  double* X1 = p; p += MVS_R * DPS;
  double* XI = p; p += 2 * MVS_R * DPS;
  double* TG = p; p += MVS_R * 32;
  p += DPS;                          // (kept: the layout of mvs_smem_bytes)
  double* LU = p; p += 8 * MVS_R;
  MvsSlots& S = *reinterpret_cast<MvsSlots*>(p);
  for (int i = tid; i < M.packLen; i += T) { packB[i] = M.packB[i]; packC[i] = M.packC[i]; }
  for (int i = tid; i < DPS; i += T) MEAN[i] = M.mean_pi[i];
  for (int i = tid; i < 3 * MVS_R * DPS + MVS_R * 32 + DPS + 8 * MVS_R; i += T) X1[i] = 0.0;
  if (tid < MVS_R) {
    S.rep[tid] = -1; S.phase[tid] = MS_IDLE; S.time[tid] = 0; S.tau[tid] = 0; S.iu[tid] = 0; S.in[tid] = 0; S.pu[tid] = 0; S.pdf[tid] = 0.0;
  }
  if (tid == 0) S.exhausted = 0;
  __syncthreads();

  const int dimh = NH * d;
  const double denom = (double)(P.m - P.k + 1);
  const unsigned FULL = 0xffffffffu;
  const int nbins = (P.nbreaks >= 2) ? P.nbreaks - 1 : 0;
  double bin_b0 = 0.0, bin_iw = 0.0;
  if (nbins) { bin_b0 = P.breaks[0]; bin_iw = (double)nbins / (P.breaks[nbins] - P.breaks[0]); }
  const unsigned long long ntodo = *P.ho_count;
  int cur = 0, stepno = 0;
  // MCMC numerators of this warp's two slots (slots warp and warp + 16): coordinate j = lane + 32 i, test-function block hh
  double am[2][NH][4];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
#pragma unroll
      for (int i = 0; i < 4; ++i) am[u][hh][i] = 0.0;
  const int kq = lane & 3;

#ifdef UBM_MVN_TIMING
  long long ta[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tn = 0, tp = clock64();
#define SMARK(i) do { long long now_ = clock64(); ta[i] += now_ - tp; tp = now_; } while (0)
#else
#define SMARK(i) do { } while (0)
#endif
  for (;;) {
#ifdef UBM_MVN_TIMING
    tp = clock64(); tn++;
#endif
    double* XIc = XI + cur * MVS_R * DPS;          // this step's rows (normals -> proposals)
    double* XIn = XI + (cur ^ 1) * MVS_R * DPS;    // filled by the producers for the next step ...
    const double* XIp = XIn;                       // ... and, until barrier 1, still holding the proposals of the step just computed
    // ================================================================ per-slot section (warp w owns slots w and w + 16)
    int my_active = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int s = warp + 16 * u;
      int ph = S.phase[s];
      const long long rep = S.rep[s];
      if (ph == MS_RUN) {
        // ---- log-density of the proposal (src/mvnorm.cpp:81-84) and the MH accept step (R/get_mh_kernels.R:33-38)
        const double ss1 = warp_butterfly(TG[s * 32 + lane]);
        const double pp1 = -0.5 * ss1 + M.cst;
        const double logu = LU[s * 8 + (int)(S.iu[s] & 7ull)];
        const double pdf1 = S.pdf[s];
        const int a1 = (logu < pp1 - pdf1) ? 1 : 0;
        const long long time = S.time[s] + 1;
        const long long tau = S.tau[s];
        __syncwarp();
        if (lane == 0) {
          S.iu[s] += 1;
          if (a1) S.pdf[s] = pp1;
          S.time[s] = time;
          S.in[s] += d;                                             // normals consumed by the step just finished
        }
        const int add_mcmc = (time <= P.lag) ? (time >= P.k) : (P.k <= time && time <= P.m);   // R/unbiasedestimator.R:90
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = lane + 32 * i;
          if (j < d) {
            const int o = s * DPS + j;
            if (a1) X1[o] = XIp[o];
            const double x1 = X1[o];
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
              const bool sq = (P.h_kind == UBM_H_SQUARE) || hh == 1;
              const double h1 = sq ? x1 * x1 : x1;
              if (add_mcmc) am[u][hh][i] = am[u][hh][i] + h1;
            }
            if (nbins && j == P.component && add_mcmc) {            // estimator_bin_ numerators (src/estimator_bin.cpp:15-22)
              const int b = bin_of(x1, P.breaks, P.nbreaks, bin_b0, bin_iw);
              if (b >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(P.bin_num + rep * nbins + b), 1ull);
            }
          }
        }
        __syncwarp();
        const bool capped = (P.max_it > 0 && time >= P.max_it);
        const long long horizon = (tau > P.m) ? tau : P.m;
        const bool finished = (time >= horizon) || (time >= P.lag && capped);
        if (finished) {
          if (lane == 0) {
            if (P.tau) P.tau[rep] = (double)tau;
            if (P.cost) P.cost[rep] = (double)(P.lag + 2 * (tau - P.lag) + ((time - tau) > 0 ? (time - tau) : 0));
            if (P.iter) P.iter[rep] = time;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = lane + 32 * i;
            if (j < d) {
#pragma unroll
              for (int hh = 0; hh < NH; ++hh) {
                const int e = hh * d + j;
                const double mc = am[u][hh][i], co = P.corr[rep * dimh + e];       // the correction sum is final since the meeting
                const double ue = mc + co;
                P.uest[rep * dimh + e] = ue / denom;
                P.mcmc[rep * dimh + e] = mc / denom;
                P.corr[rep * dimh + e] = co / denom;
              }
            }
          }
          ph = MS_IDLE;
        }
      } else if (ph == MS_RESUME) {
        ph = MS_RUN;        // the producers generated the d normals of its first step during the step that just ended
      }
      if (ph == MS_IDLE) {
        long long r = -1;
        if (lane == 0 && !S.exhausted) {
          const unsigned long long got = atomicAdd(P.work_counter, 1ull);
          if (got < ntodo) r = (long long)P.ho_list[got];
          else S.exhausted = 1;
        }
        r = __shfl_sync(FULL, r, 0);
        if (r >= 0) {
          ph = MS_RESUME;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = lane + 32 * i;
            if (j < d) {
              X1[s * DPS + j] = P.uest[r * dimh + j];
#pragma unroll
              for (int hh = 0; hh < NH; ++hh) am[u][hh][i] = P.mcmc[r * dimh + hh * d + j];
            }
          }
          if (lane == 0) {
            S.rep[s] = r; S.pdf[s] = P.cost[r];
            S.time[s] = P.iter[r]; S.tau[s] = (long long)P.tau[r];
            S.iu[s] = (unsigned long long)P.ho_int[r * 2 + 0]; S.in[s] = (unsigned long long)P.ho_int[r * 2 + 1];
            S.pu[s] = S.iu[s];
          }
        } else if (lane == 0) S.rep[s] = -1;
      }
      __syncwarp();
      if (lane == 0) S.phase[s] = ph;
      if (ph != MS_IDLE) my_active = 1;
    }
    if (tid == 0) S.draw_next = 0;
    SMARK(0);
    const int any_active = __syncthreads_or(my_active);      // CTA barrier 1: phases of this step are published
    if (!any_active) break;
    SMARK(1);
    // the d normals of the NEXT step of every live slot: 32 independent jobs handed out through a shared counter — to the
    // producer warps from the start of the step, to the consumer warps once their GEMM phases are done (with 32 slots the
    // draws, not the GEMMs, are the longer half of a step)
    // A job is 128 consecutive draws of the concatenation of the 32 rows (row length 4 nCG, 100 at d = 100): 25 jobs instead
    // of 32 row-sized ones with a quarter of the lanes idle; a lane's four draws never straddle two rows.
    auto draw_jobs = [&]() {
      const int D4 = 4 * nCG, njobs = (MVS_R * D4 + 127) >> 7;
      for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(&S.draw_next, 1);
        job = __shfl_sync(FULL, job, 0);
        if (job >= njobs) break;
        const int g = 128 * job + 4 * lane;
        const int s = g / D4, o = g - s * D4;
        const int ph = (s < MVS_R) ? S.phase[s] : MS_IDLE;
        const bool on = ph != MS_IDLE;
        const int sc = on ? s : 0;
        const unsigned long long base = (ph == MS_RESUME) ? S.in[sc] : S.in[sc] + d;
        draw_n4_lanes(P, S.rep[sc], base + (unsigned)o, on ? d - o : 0, XIn + sc * DPS + o, lane);
      }
    };

    if (warp < MVN_CW) {
      // ================================================================ CONSUMERS (warps 0-7)
      const int cw = warp;
      const unsigned mR = __ballot_sync(FULL, S.phase[lane] == MS_RUN);          // lane = slot
      const unsigned tm = ((mR & 0xffu) ? 1u : 0u) | ((mR & 0xff00u) ? 2u : 0u) | ((mR & 0xff0000u) ? 4u : 0u) | ((mR >> 24) ? 8u : 0u);
      if (tm) {
        const int rq = lane >> 2;                                                 // this lane's row inside a tile of 8 slots
        const double* const rc[4] = {XIc + rq * DPS, XIc + (8 + rq) * DPS, XIc + (16 + rq) * DPS, XIc + (24 + rq) * DPS};
        double acc[4][2], accB[4][2];
        int qmap[4];
        const int jt0 = M.jobs[cw][0], jt1 = M.jobs[cw][1];
        // ------------------------------------------------------------ proposals = x + in^T C_prop (in place: K loops, barrier, writes)
        if (jt0 != 255) mvn_gemm<0, false>(packB, d, jt0, lane, tm, rc, rc, nullptr, acc, qmap);
        if (jt1 != 255) mvn_gemm<0, false>(packB, d, jt1, lane, tm, rc, rc, nullptr, accB, qmap);
        consumer_bar();
        const int ntl = __popc(tm);
#pragma unroll
        for (int jb = 0; jb < MVN_JMAX; ++jb) {
          const int jt = jb ? jt1 : jt0;
          if (jt != 255) {
            const int n0 = 8 * jt + 2 * kq;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < ntl) {
                const int s = 8 * qmap[i] + rq;
                if ((mR >> s) & 1u) {
                  const double* base = X1 + s * DPS + n0;
                  const double v0 = jb ? accB[i][0] : acc[i][0], v1 = jb ? accB[i][1] : acc[i][1];
                  double2 o;
                  o.x = (n0 < d) ? (base[0] + v0) : 0.0;          // src/mvnorm.cpp:40-44
                  o.y = (n0 + 1 < d) ? (base[1] + v1) : 0.0;
                  *reinterpret_cast<double2*>(XIc + s * DPS + n0) = o;
                }
              }
            }
          }
        }
        consumer_bar();
        SMARK(2);
        // ------------------------------------------------------------ w = Cinv_pi^T (prop - mean), |w|^2 per 4-column group
        for (int jb = 0; jb < MVN_JMAX; ++jb) {
          const int jt = M.jobs[cw][jb];
          if (jt == 255) continue;
          mvn_gemm<ZERO_MEAN ? 0 : 1, false>(packC, d, jt, lane, tm, rc, rc, MEAN, acc, qmap);
          const int g = 2 * jt + (kq >> 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < ntl) {
              const int s = 8 * qmap[i] + rq;
              const double t = mvn_group_sq(acc[i][0], acc[i][1], lane);
              if (((mR >> s) & 1u) && (kq & 1)) TG[s * 32 + g] = t;
            }
          }
        }
      }
      SMARK(3);
      draw_jobs();
      SMARK(4);
    } else {
      // ================================================================ PRODUCERS (warps 8-15): draws of the NEXT step
      const int pw = warp - MVN_CW;
      {   // log-uniform rings of this warp's four slots (8 lanes per slot), kept ahead of the consumers
        const int s = pw + (lane >> 3) * MVN_PW, l8 = lane & 7;
        const bool on = S.phase[s] != MS_IDLE;
        unsigned long long pu = 0, target = 0;
        bool past_end = false;
        // a step consumes one log-uniform; the ring is topped up to six ahead every fourth step (or when a slot has fewer
        // than two left), so that the dependent Philox + log chain is paid once per four steps with four lanes busy
        if (on) {
          pu = S.pu[s];
          const unsigned long long iu = S.iu[s];
          if ((stepno & 3) == 0 || pu < iu + 2) target = iu + 6;
          const unsigned long long idx = pu + l8;
          if (idx < target) LU[s * 8 + (int)(idx & 7ull)] = ubm_log(draw_u_at<false>(P, S.rep[s], idx, &past_end));
        }
        __syncwarp();
        if (on && l8 == 0 && target > pu) S.pu[s] = target;
      }
      SMARK(3);
      draw_jobs();
      SMARK(4);
    }
    __syncthreads();      // CTA barrier 2: step computed, draws of the next step in place
    SMARK(5);
    cur ^= 1; ++stepno;
  }
#ifdef UBM_MVN_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 256))
    printf("mvn single timing (cycles/step, block 0, warp %d): steps %lld | slots %.0f barrier1 %.0f | %s %.0f, %s %.0f | draw queue %.0f | barrier2 %.0f\n",
           warp, tn, (double)ta[0] / tn, (double)ta[1] / tn, warp ? "-" : "phase B", (double)ta[2] / tn, warp ? "log-uniform ring" : "phase C",
           (double)ta[3] / tn, (double)ta[4] / tn, (double)ta[5] / tn);
#endif
}

// the two-pass launch of an estimator run: coupled pass (hand-off at the meeting), then the single-chain continuation
template <class Buf>
inline int mvn_launch_two_pass(const MvnDevice& M, RunParams P, int nsm, cudaStream_t st, Buf& scratch, int64_t* launches,
                               std::string* err) {
  const int nh = (P.h_kind == UBM_H_BOTH) ? 2 : 1;
  const size_t n = (size_t)P.nrep;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_int = take(n * 2 * 8), o_list = take(n * 4), o_count = take(8);
  if (scratch.bytes < off) { if (scratch.alloc(off) != cudaSuccess) { *err = "mvn: hand-off buffers"; return UBM_ERR_NOMEM; } }
  unsigned char* base = scratch.template as<unsigned char>();
  P.ho_int = reinterpret_cast<long long*>(base + o_int);
  P.ho_list = reinterpret_cast<int*>(base + o_list); P.ho_count = reinterpret_cast<unsigned long long*>(base + o_count);
  if (cudaMemsetAsync(P.ho_count, 0, 8, st) != cudaSuccess) { *err = "mvn: hand-off counter"; return UBM_ERR_CUDA; }
  int stl = mvn_launch(M, P, false, nsm, st, scratch, launches, err);      // pass 1 (scratch is not used by it)
  if (stl != UBM_OK) return stl;
  if (cudaMemsetAsync(P.work_counter, 0, 8, st) != cudaSuccess) { *err = "mvn: work counter"; return UBM_ERR_CUDA; }
  const size_t smem = mvs_smem_bytes(M.packLen, M.DPS);
  void (*kern)(const MvnDevice, const RunParams);
  const int variant = (M.zero_mean ? 2 : 0) | (nh == 2 ? 1 : 0);
  switch (variant) {
    case 0: kern = mvn_single_kernel<false, 1>; break;
    case 1: kern = mvn_single_kernel<false, 2>; break;
    case 2: kern = mvn_single_kernel<true, 1>; break;
    default: kern = mvn_single_kernel<true, 2>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = std::string("mvn: smem attribute (single-chain pass): ") + cudaGetErrorString(e); return UBM_ERR_CUDA; }
  const int grid = (int)std::min<int64_t>(nsm, (P.nrep + MVS_R - 1) / MVS_R);
  kern<<<grid, MVN_T, smem, st>>>(M, P);
  *launches += 1;
  return UBM_OK;
}
inline bool mvn_two_pass_fits(const MvnDevice& M, const RunParams& P) {   // every per-replicate output row is there to carry the hand-off
  return mvs_smem_bytes(M.packLen, M.DPS) <= 227 * 1024 && P.uest && P.mcmc && P.corr && P.tau && P.cost && P.iter;
}

}  // namespace ubm
