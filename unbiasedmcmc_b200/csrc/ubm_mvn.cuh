// ubm_mvn.cuh — C2: multivariate Normal target, RW-MH with reflection-max coupled proposals.
//
// Replaces, fused into one persistent kernel per launch:
//   drivers    R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//   kernels    R/get_mh_kernels.R:29-77 (d > 1 branch)
//   numerics   src/mvnorm.cpp:30-46 (fast_rmvnorm_cholesky_), :71-86 (fast_dmvnorm_cholesky_inverse_),
//              :185-236 (rmvnorm_reflection_max_coupling_)
//
// Design (B200): the three upper-triangular factors (proposal Cholesky, its inverse, the target's Cholesky
// inverse) are block-packed and stay in SHARED MEMORY for the life of the kernel; a CTA runs R replicate
// "slots" in lock step so that the five vector x triangular-matrix products of a coupled step become three
// small register-tiled GEMM phases ([rows x d] . [d x d]) on the FP64 pipe:
//   phase A  z = -(x2 - x1)^T Cinv_prop            (coupled slots)
//   phase B  prop = x + xi^T C_prop  (and eta)     (all active rows)
//   phase C  w = Cinv_pi^T (prop - mean_pi)        (all proposals) -> log-density
// Every output element is accumulated by ONE thread, sequentially in k with fma, so results are bit-identical
// to the oracle's plain loops.  Triangular work is balanced by pairing column group g with nCG-1-g.
// Slots are refilled from a global counter as soon as a replicate finishes (work stealing).
// Reductions over a d-vector use the canonical order "fma chain inside each 4-column group, xor-butterfly
// over the <= 32 groups" (oracle: sum_g4_*).
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "ubm_common.cuh"

namespace ubm {

struct MvnDevice {
  int d, DP, nCG, nPairs, zero_mean;
  int packLen;                 // doubles per packed matrix
  double cst;                  // src/mvnorm.cpp:74-75
  const double *mean_pi, *init_mean, *packA, *packB, *packC, *packI;   // device
  size_t offs[6];              // host-side: offsets inside the blob
};

// offset of column-group g inside a packed matrix: rows k = 0 .. min(4g+4, d)-1, 4 doubles per row
__host__ __device__ inline int mvn_group_rows(int g, int d) { int K = 4 * g + 4; return K < d ? K : d; }

inline void mvn_pack_matrix(const double* Ucm, int d, int nCG, std::vector<double>* blob) {
  for (int g = 0; g < nCG; ++g) {
    int K = mvn_group_rows(g, d);
    for (int k = 0; k < K; ++k)
      for (int c = 0; c < 4; ++c) {
        int j = 4 * g + c;
        blob->push_back((j < d && k <= j) ? Ucm[(size_t)j * d + k] : 0.0);
      }
  }
}

inline void mvn_pack_host(const ubm_model_mvn* s, MvnDevice* M, std::vector<double>* blob) {
  const int d = s->d;
  M->d = d; M->nCG = (d + 3) / 4; M->DP = 4 * M->nCG; M->nPairs = (M->nCG + 1) / 2;
  M->zero_mean = 1;
  for (int j = 0; j < d; ++j) if (s->mean_pi[j] != 0.0) M->zero_mean = 0;
  double halflogdet = 0.0;
  for (int j = 0; j < d; ++j) halflogdet = halflogdet + ubm_log(s->chol_inv_pi[(size_t)j * d + j]);
  M->cst = -(-halflogdet) - (d * 0.9189385);     // truncated constant kept verbatim (src/mvnorm.cpp:75)
  blob->clear();
  M->offs[0] = blob->size(); for (int j = 0; j < M->DP; ++j) blob->push_back(j < d ? s->mean_pi[j] : 0.0);
  M->offs[1] = blob->size(); for (int j = 0; j < M->DP; ++j) blob->push_back(j < d ? s->init_mean[j] : 0.0);
  M->offs[2] = blob->size(); mvn_pack_matrix(s->chol_inv_prop, d, M->nCG, blob);
  M->packLen = (int)(blob->size() - M->offs[2]);
  M->offs[3] = blob->size(); mvn_pack_matrix(s->chol_prop, d, M->nCG, blob);
  M->offs[4] = blob->size(); mvn_pack_matrix(s->chol_inv_pi, d, M->nCG, blob);
  M->offs[5] = blob->size(); mvn_pack_matrix(s->init_chol, d, M->nCG, blob);
}
inline void mvn_bind_device(MvnDevice* M, const double* dev) {
  M->mean_pi = dev + M->offs[0]; M->init_mean = dev + M->offs[1];
  M->packA = dev + M->offs[2]; M->packB = dev + M->offs[3]; M->packC = dev + M->offs[4]; M->packI = dev + M->offs[5];
}

enum { SL_IDLE = 0, SL_INIT = 1, SL_SINGLE = 2, SL_COUPLED = 3 };
#define MVN_MAXR 16

struct MvnSlots {   // per-slot scalars, in shared memory
  long long rep[MVN_MAXR], fin_rep[MVN_MAXR], time[MVN_MAXR], tau[MVN_MAXR];
  unsigned long long iu[MVN_MAXR], in[MVN_MAXR];
  double pdf1[MVN_MAXR], pdf2[MVN_MAXR], normz[MVN_MAXR], edotxi[MVN_MAXR], ppdf[2 * MVN_MAXR];
  double fin_tau[MVN_MAXR], fin_cost[MVN_MAXR];
  long long fin_iter[MVN_MAXR];
  int phase[MVN_MAXR], met[MVN_MAXR], ident[MVN_MAXR], acc1[MVN_MAXR], acc2[MVN_MAXR], bad[MVN_MAXR];
  int rowsA[MVN_MAXR], rowsB[2 * MVN_MAXR], rowsI[2 * MVN_MAXR], rowsC[2 * MVN_MAXR];
  int nA, nB, nI, nC, nActive, exhausted;
};

// One GEMM unit: RT input rows against the column group g of a packed matrix; acc[r][c] += in_r[k] M[k][4g+c].
template <int RT, bool SUBMEAN>
__device__ __forceinline__ void mvn_unit(const double* __restrict__ pack, int g, int d, const double* const* in,
                                         const double* __restrict__ mean, double (&acc)[RT][4]) {
  const int K = mvn_group_rows(g, d);
  // group offset: sum_{g' < g} 4 * min(4g'+4, d).  Groups with 4g'+4 <= d are full: 16 (1 + ... + g) = 8 g (g+1).
  // Only the last group can be clipped, and it has no successor, so the closed form is exact for every g.
  const double* mp = pack + 8 * g * (g + 1);
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const double2 m01 = *reinterpret_cast<const double2*>(mp + 4 * k);
    const double2 m23 = *reinterpret_cast<const double2*>(mp + 4 * k + 2);
    double mu = SUBMEAN ? mean[k] : 0.0;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      double v = in[r][k];
      if (SUBMEAN) v = v - mu;
      acc[r][0] = ubm_fma(v, m01.x, acc[r][0]);
      acc[r][1] = ubm_fma(v, m01.y, acc[r][1]);
      acc[r][2] = ubm_fma(v, m23.x, acc[r][2]);
      acc[r][3] = ubm_fma(v, m23.y, acc[r][3]);
    }
  }
}

__device__ __forceinline__ double warp_butterfly(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

struct MvnSmem {
  double *packA, *packB, *packC;
  double *X1, *X2, *XI, *ETA, *P1, *P2;   // [R][DP]
  double* TG;                             // [2R][32]
  double* ZERO;                           // [DP] zeros (padding rows of the last row group)
  MvnSlots* S;
};

__host__ __device__ inline size_t mvn_smem_bytes(int packLen, int DP, int R) {
  return (size_t)(3 * packLen + 6 * R * DP + 2 * R * 32 + DP) * sizeof(double) + sizeof(MvnSlots) + 16;
}

// row code: slot * 2 + which (which = 0: chain-1 row [XI -> P1], 1: chain-2 row [ETA -> P2])
template <bool TAPE, bool ZERO_MEAN>
__global__ void __launch_bounds__(128) mvn_driver_kernel(const MvnDevice M, const RunParams P, const int R,
                                                         double* __restrict__ acc_scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const int d = M.d, DP = M.DP, nCG = M.nCG, nPairs = M.nPairs;
  MvnSmem sm;
  {
    double* p = reinterpret_cast<double*>(smem_raw);
    sm.packA = p; p += M.packLen; sm.packB = p; p += M.packLen; sm.packC = p; p += M.packLen;
    sm.X1 = p; p += R * DP; sm.X2 = p; p += R * DP; sm.XI = p; p += R * DP; sm.ETA = p; p += R * DP;
    sm.P1 = p; p += R * DP; sm.P2 = p; p += R * DP;
    sm.TG = p; p += 2 * R * 32;
    sm.ZERO = p; p += DP;
    sm.S = reinterpret_cast<MvnSlots*>(p);
  }
  MvnSlots& S = *sm.S;
  for (int i = tid; i < M.packLen; i += T) { sm.packA[i] = M.packA[i]; sm.packB[i] = M.packB[i]; sm.packC[i] = M.packC[i]; }
  for (int i = tid; i < 6 * R * DP; i += T) sm.X1[i] = 0.0;
  for (int i = tid; i < 2 * R * 32 + DP; i += T) sm.TG[i] = 0.0;
  if (tid < MVN_MAXR) {
    S.rep[tid] = -1; S.fin_rep[tid] = -1; S.phase[tid] = SL_IDLE; S.time[tid] = 0; S.tau[tid] = 0; S.met[tid] = 0;
    S.iu[tid] = 0; S.in[tid] = 0; S.bad[tid] = 0;
  }
  if (tid == 0) { S.exhausted = 0; S.nActive = 0; }
  __syncthreads();

  const int dimh = (P.h_kind == UBM_H_BOTH) ? 2 * d : d;
  const double denom = (double)(P.m - P.k + 1);
  double* my_acc = acc_scratch ? acc_scratch + (size_t)blockIdx.x * R * 2 * dimh : nullptr;   // [R][2][dimh]
  const double* mean_pi = M.mean_pi;

  for (;;) {
    // ---------------------------------------------------------------- S1: slot bookkeeping (owner threads)
    if (tid < R) {
      const int s = tid;
      S.fin_rep[s] = -1;
      int ph = S.phase[s];
      if (ph != SL_IDLE) {
        long long time = S.time[s];
        bool met = S.met[s] != 0;
        bool finished;
        bool capped = (P.max_it > 0 && time >= P.max_it) || S.bad[s];
        if (P.mode == MODE_MEETING) finished = met || (time >= P.lag && capped);
        else {
          long long horizon = (S.tau[s] > P.m) ? S.tau[s] : P.m;
          finished = (met && time >= horizon) || (time >= P.lag && capped);
        }
        if (finished) {
          long long tau = S.tau[s];
          S.fin_rep[s] = S.rep[s];
          S.fin_tau[s] = met ? (double)tau : dinf();
          S.fin_iter[s] = time;
          if (!met) S.fin_cost[s] = dinf();
          else if (P.mode == MODE_MEETING) S.fin_cost[s] = (double)(P.lag + 2 * (tau - P.lag));
          else S.fin_cost[s] = (double)(P.lag + 2 * (tau - P.lag) + ((time - tau) > 0 ? (time - tau) : 0));
          if (S.bad[s]) atomicExch(P.status, UBM_ERR_TAPE);
          ph = SL_IDLE;
        } else if (time < P.lag) ph = SL_SINGLE;
        else ph = met ? SL_SINGLE : SL_COUPLED;
      }
      if (ph == SL_IDLE && !S.exhausted) {
        unsigned long long r = atomicAdd(P.work_counter, 1ull);
        if (r < (unsigned long long)P.nrep) {
          S.rep[s] = (long long)r; S.time[s] = 0; S.tau[s] = 0; S.met[s] = 0; S.iu[s] = 0; S.in[s] = 0; S.bad[s] = 0;
          ph = SL_INIT;
        } else {
          S.rep[s] = -1;
          S.exhausted = 1;   // benign race: every writer writes 1
        }
      }
      S.phase[s] = ph;
    }
    __syncthreads();
    if (tid == 0) {
      int nA = 0, nAct = 0;
      for (int s = 0; s < R; ++s) {
        if (S.phase[s] != SL_IDLE) ++nAct;
        if (S.phase[s] == SL_COUPLED) S.rowsA[nA++] = s;
      }
      S.nA = nA; S.nActive = nAct;
    }
    __syncthreads();
    // ---------------------------------------------------------------- write-out of finished replicates
    {
      for (int s = 0; s < R; ++s) {
        long long fr = S.fin_rep[s];
        if (fr < 0) continue;
        if (tid == 0) {
          if (P.tau) P.tau[fr] = S.fin_tau[s];
          if (P.cost) P.cost[fr] = S.fin_cost[s];
          if (P.iter) P.iter[fr] = S.fin_iter[s];
        }
        if (P.mode == MODE_ESTIMATOR) {
          double* a = my_acc + (size_t)s * 2 * dimh;
          for (int e = tid; e < dimh; e += T) {
            double mc = a[e], co = a[dimh + e];
            double ue = mc + co;
            P.uest[fr * dimh + e] = ue / denom;
            if (P.mcmc) P.mcmc[fr * dimh + e] = mc / denom;
            if (P.corr) P.corr[fr * dimh + e] = co / denom;
          }
        }
      }
    }
    if (S.nActive == 0) break;
    // ---------------------------------------------------------------- S2: draws xi (and init eta), diff
    for (int w = tid; w < R * nCG; w += T) {
      const int s = w / nCG, g = w - s * nCG;
      const int ph = S.phase[s];
      if (ph == SL_IDLE) continue;
      const long long rep = S.rep[s];
      const unsigned long long in0 = S.in[s];
      bool bad = false;
      double* xi = sm.XI + s * DP + 4 * g;
      double* eta = sm.ETA + s * DP + 4 * g;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = 4 * g + c;
        xi[c] = (j < d) ? draw_n_at<TAPE>(P, rep, in0 + j, &bad) : 0.0;
      }
      if (ph == SL_INIT) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = 4 * g + c;
          eta[c] = (j < d) ? draw_n_at<TAPE>(P, rep, in0 + d + j, &bad) : 0.0;
        }
      } else if (ph == SL_COUPLED) {
#pragma unroll
        for (int c = 0; c < 4; ++c) eta[c] = sm.X2[s * DP + 4 * g + c] - sm.X1[s * DP + 4 * g + c];   // mu2 - mu1 (:191)
      }
      if (bad) S.bad[s] = 1;
    }
    __syncthreads();
    if (tid < R && S.phase[tid] != SL_IDLE) S.in[tid] += (S.phase[tid] == SL_INIT) ? 2 * d : d;
    // ---------------------------------------------------------------- phase A: z = -(diff^T Cinv_prop) -> P2, |z|^2
    {
      const int nA = S.nA;
      const int nRG = (nA + 1) / 2;
      for (int u = tid; u < nRG * nPairs; u += T) {
        const int rg = u / nPairs, pr = u - rg * nPairs;
        const double* in[2];
        int slot[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          int idx = 2 * rg + r;
          slot[r] = (idx < nA) ? S.rowsA[idx] : -1;
          in[r] = (slot[r] >= 0) ? sm.ETA + slot[r] * DP : sm.ZERO;
        }
        const int gs[2] = {pr, nCG - 1 - pr};
        const int ng = (gs[0] == gs[1]) ? 1 : 2;
        for (int q = 0; q < ng; ++q) {
          const int g = gs[q];
          double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
          mvn_unit<2, false>(sm.packA, g, d, in, nullptr, acc);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (slot[r] < 0) continue;
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              double z = -acc[r][c];                              // :198
              sm.P2[slot[r] * DP + 4 * g + c] = z;
              t = ubm_fma(z, z, t);
            }
            sm.TG[(2 * slot[r]) * 32 + g] = t;
          }
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- S4a: normz per coupled slot
    for (int i = warp; i < S.nA; i += nwarp) {
      const int s = S.rowsA[i];
      double v = warp_butterfly(sm.TG[(2 * s) * 32 + lane]);
      if (lane == 0) S.normz[s] = sqrt(v);                        // :199
    }
    __syncthreads();
    // ---------------------------------------------------------------- S4b: e = z / normz -> P2, e . xi partials
    for (int w = tid; w < S.nA * nCG; w += T) {
      const int i = w / nCG, g = w - i * nCG;
      const int s = S.rowsA[i];
      const double nz = S.normz[s];
      if (nz < 1e-15) continue;
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = 4 * g + c;
        double e = sm.P2[s * DP + j] / nz;                        // :212
        sm.P2[s * DP + j] = e;
        t = ubm_fma(e, sm.XI[s * DP + j], t);                     // :216
      }
      sm.TG[(2 * s) * 32 + g] = t;
    }
    __syncthreads();
    // ---------------------------------------------------------------- S4c: e.xi, utilde, identical?
    for (int i = warp; i < S.nA; i += nwarp) {
      const int s = S.rowsA[i];
      const double nz = S.normz[s];
      double edx = warp_butterfly(sm.TG[(2 * s) * 32 + lane]);
      if (lane == 0) {
        if (nz < 1e-15) { S.ident[s] = 1; S.edotxi[s] = 0.0; }    // :200-210, no uniform drawn
        else {
          bool bad = false;
          double ut = draw_u_at<TAPE>(P, S.rep[s], S.iu[s], &bad);   // :213-215
          S.iu[s] += 1;
          if (bad) S.bad[s] = 1;
          double a = edx + nz;
          S.ident[s] = (ubm_log(ut) < (-0.5 * a * a + 0.5 * edx * edx)) ? 1 : 0;   // :217
          S.edotxi[s] = edx;
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- S4d: eta (non-identical only), row lists
    for (int w = tid; w < S.nA * nCG; w += T) {
      const int i = w / nCG, g = w - i * nCG;
      const int s = S.rowsA[i];
      if (S.ident[s]) continue;
      const double two_edx = 2. * S.edotxi[s];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = 4 * g + c;
        sm.ETA[s * DP + j] = sm.XI[s * DP + j] - two_edx * sm.P2[s * DP + j];   // :221
      }
    }
    if (tid == 0) {
      int nB = 0, nI = 0, nC = 0;
      for (int s = 0; s < R; ++s) {
        const int ph = S.phase[s];
        if (ph == SL_IDLE) continue;
        if (ph == SL_INIT) { S.rowsI[nI++] = 2 * s; S.rowsI[nI++] = 2 * s + 1; S.rowsC[nC++] = 2 * s; S.rowsC[nC++] = 2 * s + 1; }
        else {
          S.rowsB[nB++] = 2 * s; S.rowsC[nC++] = 2 * s;
          if (ph == SL_COUPLED && !S.ident[s]) { S.rowsB[nB++] = 2 * s + 1; S.rowsC[nC++] = 2 * s + 1; }
        }
      }
      S.nB = nB; S.nI = nI; S.nC = nC;
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase B: proposals = x + in^T C_prop ; init
    for (int pass = 0; pass < 2; ++pass) {
      const int nrows = pass ? S.nI : S.nB;
      const int* rows = pass ? S.rowsI : S.rowsB;
      const double* pack = pass ? M.packI : sm.packB;
      const int nRG = (nrows + 3) / 4;
      for (int u = tid; u < nRG * nPairs; u += T) {
        const int rg = u / nPairs, pr = u - rg * nPairs;
        const double* in[4];
        int code[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          int idx = 4 * rg + r;
          code[r] = (idx < nrows) ? rows[idx] : -1;
          in[r] = (code[r] < 0) ? sm.ZERO : ((code[r] & 1) ? sm.ETA : sm.XI) + (code[r] >> 1) * DP;
        }
        const int gs[2] = {pr, nCG - 1 - pr};
        const int ng = (gs[0] == gs[1]) ? 1 : 2;
        for (int q = 0; q < ng; ++q) {
          const int g = gs[q];
          double acc[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
          mvn_unit<4, false>(pack, g, d, in, nullptr, acc);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (code[r] < 0) continue;
            const int s = code[r] >> 1, which = code[r] & 1;
            const double* base = pass ? (M.init_mean + 4 * g) : ((which ? sm.X2 : sm.X1) + s * DP + 4 * g);
            double* out = (which ? sm.P2 : sm.P1) + s * DP + 4 * g;
#pragma unroll
            for (int c = 0; c < 4; ++c) out[c] = (4 * g + c < d) ? (base[c] + acc[r][c]) : 0.0;   // :40-44, :224-225
          }
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase C: w = Cinv_pi^T (prop - mean), |w|^2
    {
      const int nrows = S.nC;
      const int nRG = (nrows + 3) / 4;
      for (int u = tid; u < nRG * nPairs; u += T) {
        const int rg = u / nPairs, pr = u - rg * nPairs;
        const double* in[4];
        int code[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          int idx = 4 * rg + r;
          code[r] = (idx < nrows) ? S.rowsC[idx] : -1;
          in[r] = (code[r] < 0) ? sm.ZERO : ((code[r] & 1) ? sm.P2 : sm.P1) + (code[r] >> 1) * DP;
        }
        const int gs[2] = {pr, nCG - 1 - pr};
        const int ng = (gs[0] == gs[1]) ? 1 : 2;
        for (int q = 0; q < ng; ++q) {
          const int g = gs[q];
          double acc[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
          mvn_unit<4, !ZERO_MEAN>(sm.packC, g, d, in, mean_pi, acc);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (code[r] < 0) continue;
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) t = ubm_fma(acc[r][c], acc[r][c], t);
            sm.TG[code[r] * 32 + g] = t;
          }
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- S8: log-densities, accept / reject
    for (int i = warp; i < S.nC; i += nwarp) {
      const int code = S.rowsC[i];
      double ss = warp_butterfly(sm.TG[code * 32 + lane]);
      if (lane == 0) S.ppdf[code] = -0.5 * ss + M.cst;             // src/mvnorm.cpp:81-84
    }
    __syncthreads();
    if (tid < R) {
      const int s = tid;
      const int ph = S.phase[s];
      if (ph == SL_INIT) {
        S.acc1[s] = 1; S.acc2[s] = 1; S.pdf1[s] = S.ppdf[2 * s]; S.pdf2[s] = S.ppdf[2 * s + 1];
      } else if (ph != SL_IDLE) {
        bool bad = false;
        double logu = ubm_log(draw_u_at<TAPE>(P, S.rep[s], S.iu[s], &bad));   // R/get_mh_kernels.R:34,57
        S.iu[s] += 1;
        if (bad) S.bad[s] = 1;
        if (ph == SL_SINGLE) {
          double pp = S.ppdf[2 * s];
          int a1 = (logu < pp - S.pdf1[s]) ? 1 : 0;
          S.acc1[s] = a1; S.acc2[s] = 0;
          if (a1) S.pdf1[s] = pp;
          S.time[s] += 1;
        } else {
          const int ident = S.ident[s];
          double pp1 = S.ppdf[2 * s];
          double pp2 = ident ? pp1 : S.ppdf[2 * s + 1];
          int a1 = 0, a2 = 0;
          if (isfinite(pp1)) a1 = (logu < pp1 - S.pdf1[s]) ? 1 : 0;
          if (isfinite(pp2)) a2 = (logu < pp2 - S.pdf2[s]) ? 1 : 0;
          S.acc1[s] = a1; S.acc2[s] = a2;
          if (a1) S.pdf1[s] = pp1;
          if (a2) S.pdf2[s] = pp2;
          S.time[s] += 1;
          if (ident && a1 && a2) { S.met[s] = 1; S.tau[s] = S.time[s]; }
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- S9: state update, estimator, trajectories
    for (int w = tid; w < R * DP; w += T) {
      const int s = w / DP, j = w - s * DP;
      const int ph = S.phase[s];
      if (ph == SL_IDLE || j >= d) continue;
      const bool coupled_step = (ph == SL_COUPLED);
      if (S.acc1[s]) sm.X1[w] = sm.P1[w];
      if (S.acc2[s]) sm.X2[w] = (coupled_step && S.ident[s]) ? sm.P1[w] : sm.P2[w];
      const long long time = S.time[s];
      const long long rep = S.rep[s];
      const double x1 = sm.X1[w];
      const double x2 = sm.X2[w];
      if (P.mode == MODE_ESTIMATOR) {
        double* a = my_acc + (size_t)s * 2 * dimh;
        const int nh = (P.h_kind == UBM_H_BOTH) ? 2 : 1;
        for (int hh = 0; hh < nh; ++hh) {
          const bool sq = (P.h_kind == UBM_H_SQUARE) || hh == 1;
          const int e = hh * d + j;
          const double h1 = sq ? x1 * x1 : x1;
          if (ph == SL_INIT) {
            a[e] = (P.k == 0) ? h1 : 0.0;                          // R/unbiasedestimator.R:55-59
            a[dimh + e] = 0.0;
          } else {
            const bool in_window = (time <= P.lag) ? (time >= P.k) : (P.k <= time && time <= P.m);   // :66, :80, :90
            if (in_window) a[e] = a[e] + h1;
            // correction: after the last lag step (:70-72) and after every coupled step (:93-95)
            const bool corr_now = (time >= P.k + P.lag) && (coupled_step || (time == P.lag));
            if (corr_now) {
              const double h2 = sq ? x2 * x2 : x2;
              const double wgt = (double)corr_weight(time, P.k, P.m, P.lag);
              a[dimh + e] = a[dimh + e] + wgt * (h1 - h2);
            }
          }
        }
      } else if (P.mode == MODE_CHAINS) {
        double* S1 = P.samples1 + (rep * P.stride) * d;
        double* S2 = P.samples2 + (rep * P.stride) * d;
        if (ph == SL_INIT) { S1[j] = x1; S2[j] = x2; }
        else {
          S1[time * d + j] = x1;
          if (time > P.lag) S2[(time - P.lag) * d + j] = (coupled_step ? x2 : x1);   // R/coupled_chains.R:60-62,77-78
        }
      }
    }
    __syncthreads();
  }
}

struct DevBufFwd;   // (capi.cu's DevBuf)

template <class Buf>
inline int mvn_launch(const MvnDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st, Buf& scratch,
                      int64_t* launches, std::string* err) {
  if (P.nbreaks >= 2) { *err = "mvn: histogram bins are not available for this model yet"; return UBM_ERR_UNSUPPORTED; }
  const size_t budget = 227 * 1024;
  int R = MVN_MAXR;
  while (R > 1 && mvn_smem_bytes(M.packLen, M.DP, R) > budget) --R;
  size_t smem = mvn_smem_bytes(M.packLen, M.DP, R);
  if (smem > budget) { *err = "mvn: d too large for shared memory"; return UBM_ERR_UNSUPPORTED; }
  const int T = 128;
  int64_t need = (P.nrep + R - 1) / R;
  int grid = (int)std::min<int64_t>(nsm, need);
  const int dimh = (P.h_kind == UBM_H_BOTH) ? 2 * M.d : M.d;
  double* acc = nullptr;
  if (P.mode == MODE_ESTIMATOR) {
    size_t bytes = (size_t)grid * R * 2 * dimh * sizeof(double);
    if (scratch.bytes < bytes) { if (scratch.alloc(bytes) != cudaSuccess) { *err = "mvn: scratch alloc failed"; return UBM_ERR_NOMEM; } }
    acc = scratch.template as<double>();
  }
  void (*kern)(const MvnDevice, const RunParams, const int, double*);
  if (tape) kern = M.zero_mean ? mvn_driver_kernel<true, true> : mvn_driver_kernel<true, false>;
  else kern = M.zero_mean ? mvn_driver_kernel<false, true> : mvn_driver_kernel<false, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = std::string("mvn: smem attribute: ") + cudaGetErrorString(e); return UBM_ERR_CUDA; }
  kern<<<grid, T, smem, st>>>(M, P, R, acc);
  *launches += 1;
  return UBM_OK;
}

}  // namespace ubm
