// ubm_blasso.cuh — Bayesian lasso: Gibbs sampler on (beta, tau2, sigma2) and its coupling.
//
// Replaces, fused into one persistent kernel per launch:
//   drivers   R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//   kernels   R/bayesianlasso.R:30-32 (rinit), :34-50 (gibbs_kernel), :52-98 (coupled_gibbs_kernel)
//   numerics  src/blassoutil.cpp:17-41 (blassoconditional), :45-88 (blassoconditional_coupled) with fast_rmvnorm_ /
//             rmvnorm_max_coupling_ (src/mvnorm.cpp:9-26, :91-129), rinversegamma[_coupled] (R/invgamma_couplings.R:17-36),
//             rinvgaussian[_coupled] (src/inversegaussian.cpp:6-63)
//
// Design: one CTA (512 threads) per replicate, the p x p work in shared memory with the block linear algebra of
// ubm_pg.cuh (CTA Cholesky, column-parallel triangular inverse, Linv' Linv by all threads): per conditional
// A = XtX + diag(1 / tau2) -> L -> Linv -> A^-1 -> mean = A^-1 XtY, LLT of sigma2 A^-1; the regression draw and its maximal
// coupling by warp 0; the residual sum of squares by all threads (512 interleaved chains, butterfly per warp, warps added
// in order); the scalar Gibbs steps (inverse Gamma for sigma2, inverse Gaussian per component of 1 / tau2, and their
// maximal couplings) sequentially by thread 0 on the replicate's typed streams — the reference's draw order.
// Arithmetic: oracle/oracle_blasso.hpp, bit for bit.  A widening of the registry, not a tuned kernel.
#pragma once
#include <cmath>
#include <string>
#include <vector>
#include "ubm_common.cuh"
#include "ubm_pg.cuh"
#include "ubm_prims.cuh"

namespace ubm {

#define BL_MAXD 128     // 2 p + 1 <= 121

struct BlassoDevice {
  int n, p;
  double lambda2, alpha1, lgam;
  const double *X, *Y, *XtX, *XtY;   // device
  size_t offs[4];
};

inline void blasso_pack_host(const ubm_model_blasso* s, BlassoDevice* M, std::vector<double>* blob) {
  const int n = s->n, p = s->p;
  M->n = n; M->p = p;
  M->lambda2 = s->lambda * s->lambda;                 // R/bayesianlasso.R:28
  M->alpha1 = (n - 1) / 2.0 + p / 2.0;                // :27
  M->lgam = std::lgamma(M->alpha1);
  std::vector<double> XtX((size_t)p * p, 0.0), XtY(p, 0.0);   // :25-26, fma chains over the observations
  for (int a = 0; a < p; ++a) {
    for (int b = 0; b <= a; ++b) {
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc = ubm_fma(s->X[(size_t)a * n + i], s->X[(size_t)b * n + i], acc);
      XtX[(size_t)a * p + b] = acc; XtX[(size_t)b * p + a] = acc;
    }
    double t = 0.0;
    for (int i = 0; i < n; ++i) t = ubm_fma(s->X[(size_t)a * n + i], s->Y[i], t);
    XtY[a] = t;
  }
  blob->clear();
  auto put = [&](const double* src, size_t len) { size_t o = blob->size(); blob->insert(blob->end(), src, src + len); return o; };
  M->offs[0] = put(s->X, (size_t)n * p);
  M->offs[1] = put(s->Y, n);
  M->offs[2] = put(XtX.data(), XtX.size());
  M->offs[3] = put(XtY.data(), XtY.size());
}
inline void blasso_bind_device(BlassoDevice* M, const double* dev) {
  M->X = dev + M->offs[0]; M->Y = dev + M->offs[1]; M->XtX = dev + M->offs[2]; M->XtY = dev + M->offs[3];
}

// the replicate's sequential typed streams, as thread 0 sees them (counters live in shared memory between phases)
template <bool TAPE>
struct BlDraws {
  const RunParams& P;
  long long rep;
  unsigned long long in, iu;
  bool bad;
  __device__ BlDraws(const RunParams& p, long long r, unsigned long long in_, unsigned long long iu_) : P(p), rep(r), in(in_), iu(iu_), bad(false) {}
  __device__ double n() { return draw_n_at<TAPE>(P, rep, in++, &bad); }
  __device__ double u() { return draw_u_at<TAPE>(P, rep, iu++, &bad); }
};

struct BlShared {
  long long rep;
  unsigned long long in, iu;
  double rssw[2][16];
  int bad;
};

// A^-1 XtY and the lower factor of sigma2 A^-1 for one or two chains (src/blassoutil.cpp:21-30): A_c in MA_c, work MB_c
__device__ inline void bl_conditional(const BlassoDevice& M, const double* sXtX, const double* sXtY, const double* st1, const double* st2,
                                      bool two, double* MA1, double* MB1, double* MA2, double* MB2, double* mean1, double* mean2,
                                      double* tmp, int tid) {
  const int p = M.p, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < p * p; e += PG_T) {
    const int j = e / p, i = e - j * p;
    const double v = sXtX[e];
    MA1[e] = (i == j) ? v + 1.0 / st1[p + j] : v;
    if (two) MA2[e] = (i == j) ? v + 1.0 / st2[p + j] : v;
  }
  __syncthreads();
  pg_cta_chol(MA1, two ? MA2 : nullptr, tmp, p, tid);
  pg_tri_inverse_cols(MA1, MB1, two ? MA2 : nullptr, MB2, p, warp, lane, 0, PG_T / 32);
  __syncthreads();
  pg_ltl(MB1, MA1, p, tid, PG_T);                       // A^-1 = Linv' Linv
  if (two) pg_ltl(MB2, MA2, p, tid, PG_T);
  __syncthreads();
  {
    const int ch = tid >> 6, j = tid & 63;
    if (ch < (two ? 2 : 1) && j < p) {
      const double* Ai = ch ? MA2 : MA1;
      double acc = 0.0;
      for (int k = 0; k < p; ++k) acc = ubm_fma(Ai[k * p + j], sXtY[k], acc);
      (ch ? mean2 : mean1)[j] = acc;
    }
  }
  __syncthreads();
  const double g1 = st1[2 * p], g2 = two ? st2[2 * p] : 0.0;
  for (int e = tid; e < p * p; e += PG_T) { MA1[e] = MA1[e] * g1; if (two) MA2[e] = MA2[e] * g2; }
  __syncthreads();
  pg_cta_chol(MA1, two ? MA2 : nullptr, tmp, p, tid);   // LLT inside fast_rmvnorm_ / rmvnorm_max_coupling_
}

// sum((Y - X beta)^2) for one or two coefficient vectors -> sh.rssw[c][warp]; thread 0 adds the 16 warp sums in order
__device__ inline void bl_rss(const BlassoDevice& M, const double* b1, const double* b2, bool two, BlShared& sh, int tid) {
  const int n = M.n, p = M.p, warp = tid >> 5, lane = tid & 31;
  double a1 = 0.0, a2 = 0.0;
  for (int i = tid; i < n; i += PG_T) {
    double x1 = 0.0, x2 = 0.0;
    for (int j = 0; j < p; ++j) {
      const double xv = M.X[(size_t)j * n + i];
      x1 = ubm_fma(xv, b1[j], x1);
      if (two) x2 = ubm_fma(xv, b2[j], x2);
    }
    const double r1 = M.Y[i] - x1;
    a1 = ubm_fma(r1, r1, a1);
    if (two) { const double r2 = M.Y[i] - x2; a2 = ubm_fma(r2, r2, a2); }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    a1 = a1 + __shfl_xor_sync(0xffffffffu, a1, off);
    a2 = a2 + __shfl_xor_sync(0xffffffffu, a2, off);
  }
  if (lane == 0) { sh.rssw[0][warp] = a1; sh.rssw[1][warp] = a2; }
}

template <bool TAPE>
__global__ void __launch_bounds__(PG_T) blasso_driver_kernel(const BlassoDevice M, const RunParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = M.p, dim = 2 * p + 1;
  double* sp = reinterpret_cast<double*>(smem_raw);
  double* sXtX = sp; sp += p * p;
  double* MA1 = sp; sp += p * p;
  double* MB1 = sp; sp += p * p;
  double* MA2 = sp; sp += p * p;
  double* MB2 = sp; sp += p * p;
  double* sXtY = sp; sp += PG_MAXP;
  double* mean1 = sp; sp += PG_MAXP;
  double* mean2 = sp; sp += PG_MAXP;
  double* x1 = sp; sp += PG_MAXP;
  double* x2 = sp; sp += PG_MAXP;
  double* tmp = sp; sp += PG_MAXP;
  double* zz = sp; sp += PG_MAXP;          // tmp, zz contiguous: the Cholesky's diagonal scratch [2][PG_MAXP]
  double* st1 = sp; sp += BL_MAXD;         // chain states c(beta, tau2, sigma2)
  double* st2 = sp; sp += BL_MAXD;
  double* acc = sp; sp += 4 * BL_MAXD;     // mcmc [2 dim], corr [2 dim]
  BlShared& sh = *reinterpret_cast<BlShared*>(sp);
  const int dimh = (P.h_kind == UBM_H_BOTH) ? 2 * dim : dim;
  const double denom = (double)(P.m - P.k + 1);
  const int nbins = (P.mode == MODE_ESTIMATOR && P.nbreaks >= 2) ? P.nbreaks - 1 : 0;
  for (int e = tid; e < p * p; e += PG_T) sXtX[e] = M.XtX[e];
  for (int e = tid; e < p; e += PG_T) sXtY[e] = M.XtY[e];
  __syncthreads();

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const unsigned long long got = atomicAdd(P.work_counter, 1ull);
      sh.rep = (got < (unsigned long long)P.nrep) ? (long long)got : -1;
    }
    __syncthreads();
    const long long rep = sh.rep;
    __syncthreads();
    if (rep < 0) break;
    // ---- rinit (R/bayesianlasso.R:30-32): c(rep(0, p), rep(1, p), 1), no draws
    for (int e = tid; e < dim; e += PG_T) { const double v = (e < p) ? 0.0 : 1.0; st1[e] = v; st2[e] = v; }
    for (int e = tid; e < 4 * BL_MAXD; e += PG_T) acc[e] = 0.0;
    if (tid == 0) { sh.in = 0; sh.iu = 0; sh.bad = 0; }
    __syncthreads();
    long long* bn = (nbins && tid == P.component) ? P.bin_num + rep * nbins : nullptr;
    double bin_b0 = 0.0, bin_iw = 0.0;
    if (nbins) { bin_b0 = P.breaks[0]; bin_iw = (double)nbins / (P.breaks[nbins] - P.breaks[0]); }
    if (P.mode == MODE_ESTIMATOR && P.k == 0 && tid < dim) {
      const double x = st1[tid];
      if (P.h_kind == UBM_H_IDENTITY) acc[tid] = x;
      else if (P.h_kind == UBM_H_SQUARE) acc[tid] = x * x;
      else { acc[tid] = x; acc[dim + tid] = x * x; }
      if (bn) { const int b = bin_of(x, P.breaks, P.nbreaks, bin_b0, bin_iw); if (b >= 0) bn[b] += 1; }
    }
    if (P.mode == MODE_CHAINS && tid < dim) {
      P.samples1[(rep * P.stride) * dim + tid] = st1[tid];
      P.samples2[(rep * P.stride) * dim + tid] = st2[tid];
    }
    long long time = 0, tau = 0;
    bool met = false;
    // ---- time loop (generic drivers; CTA-uniform control flow)
    for (;;) {
      __syncthreads();
      bool go;
      const bool bad = sh.bad != 0;
      if (time < P.lag) go = !bad;
      else {
        const bool capped = (P.max_it > 0 && time >= P.max_it) || bad;
        if (P.mode == MODE_MEETING) go = !met && !capped;
        else go = (!met || time < (tau > P.m ? tau : P.m)) && !capped;
      }
      if (!go) break;
      time += 1;
      const bool coupled_step = (time > P.lag) && !met;
      // ---- beta | tau2, sigma2
      bool two = false;
      if (coupled_step) {                                             // R/bayesianlasso.R:60: are the conditionals the same?
        int eq = 1;
        if (tid <= p) eq = (tid < p) ? (st1[p + tid] == st2[p + tid]) : (st1[2 * p] == st2[2 * p]);
        two = !__syncthreads_and(eq);
      }
      bl_conditional(M, sXtX, sXtY, st1, st2, two, MA1, MB1, MA2, MB2, mean1, mean2, tmp, tid);
      if (warp == 0) {
        unsigned long long in0 = sh.in, iu0 = sh.iu;
        bool badr = false;
        for (int j = lane; j < p; j += 32) zz[j] = draw_n_at<TAPE>(P, rep, in0 + j, &badr);     // src/mvnorm.cpp:18-24 / :96
        in0 += p;
        __syncwarp();
        for (int j = lane; j < p; j += 32) {
          double a = 0.0;
          for (int i = 0; i <= j; ++i) a = ubm_fma(zz[i], MA1[i * p + j], a);
          x1[j] = a + mean1[j];
        }
        __syncwarp();
        badr = __any_sync(0xffffffffu, badr);                         // warp-uniform from here on (only lanes < p read the tape)
        if (two) {                                                    // rmvnorm_max_coupling_ (src/mvnorm.cpp:103-123)
          const double logu = ubm_log(draw_u_at<TAPE>(P, rep, iu0, &badr));
          iu0 += 1;
          const double d11 = pg_warp_dmvnorm(x1, mean1, MA1, p, lane, tmp);
          const double d12 = pg_warp_dmvnorm(x1, mean2, MA2, p, lane, tmp);
          if (d11 + logu < d12) {
            for (int j = lane; j < p; j += 32) x2[j] = x1[j];
          } else {
            bool accept = false;
            while (!accept && !badr) {
              for (int j = lane; j < p; j += 32) zz[j] = draw_n_at<TAPE>(P, rep, in0 + j, &badr);
              in0 += p;
              __syncwarp();
              for (int j = lane; j < p; j += 32) {
                double a = 0.0;
                for (int i = 0; i <= j; ++i) a = ubm_fma(zz[i], MA2[i * p + j], a);
                x2[j] = a + mean2[j];
              }
              __syncwarp();
              const double lu = ubm_log(draw_u_at<TAPE>(P, rep, iu0, &badr));
              iu0 += 1;
              const double d22 = pg_warp_dmvnorm(x2, mean2, MA2, p, lane, tmp);
              const double d21 = pg_warp_dmvnorm(x2, mean1, MA1, p, lane, tmp);
              if (d22 + lu > d21) accept = true;
              badr = __any_sync(0xffffffffu, badr);
            }
          }
        }
        __syncwarp();
        for (int j = lane; j < p; j += 32) {
          st1[j] = x1[j];
          if (coupled_step) st2[j] = two ? x2[j] : x1[j];
        }
        if (lane == 0) { sh.in = in0; sh.iu = iu0; }
        if (__any_sync(0xffffffffu, badr) && lane == 0) sh.bad = 1;
      }
      __syncthreads();
      // ---- sigma2 | beta, tau2 and tau2 | beta, sigma2
      bl_rss(M, st1, st2, two, sh, tid);
      __syncthreads();
      if (tid == 0) {
        double n1 = 0.0, n2 = 0.0;
        for (int w = 0; w < 16; ++w) { n1 = n1 + sh.rssw[0][w]; n2 = n2 + sh.rssw[1][w]; }
        double q1 = 0.0, q2 = 0.0;                                    // betaDbeta, src/blassoutil.cpp:34-37
        for (int j = 0; j < p; ++j) {
          q1 = q1 + st1[j] * st1[j] / st1[p + j];
          if (two) q2 = q2 + st2[j] * st2[j] / st2[p + j];
        }
        const double rate1 = 0.5 * (n1 + q1), rate2 = two ? 0.5 * (n2 + q2) : rate1;
        BlDraws<TAPE> D(P, rep, sh.in, sh.iu);
        if (!coupled_step) {                                          // gibbs_kernel, R/bayesianlasso.R:43-48
          const double g = 1.0 / prim_rgamma(M.alpha1, rate1, D);
          st1[2 * p] = g;
          const double sl = sqrt(M.lambda2 * g);
          for (int j = 0; j < p; ++j) st1[p + j] = 1.0 / prim_rinvgaussian(sl / fabs(st1[j]), M.lambda2, D);
        } else {                                                      // coupled_gibbs_kernel, :77-96
          double g1, g2;
          prim_positive_max_coupling(1, M.alpha1, M.alpha1, rate1, rate2, M.alpha1 * ubm_log(rate1) - M.lgam,
                                     M.alpha1 * ubm_log(rate2) - M.lgam, D, &g1, &g2);
          st1[2 * p] = g1; st2[2 * p] = g2;
          const double sl1 = sqrt(M.lambda2 * g1), sl2 = sqrt(M.lambda2 * g2);
          for (int j = 0; j < p; ++j) {
            if (g1 == g2 && st1[j] == st2[j]) {
              const double it = prim_rinvgaussian(sl1 / fabs(st1[j]), M.lambda2, D);
              st1[p + j] = 1.0 / it; st2[p + j] = st1[p + j];
            } else {
              double u1, u2;
              prim_positive_max_coupling(2, sl1 / fabs(st1[j]), sl2 / fabs(st2[j]), M.lambda2, M.lambda2, 0.0, 0.0, D, &u1, &u2);
              st1[p + j] = 1.0 / u1; st2[p + j] = 1.0 / u2;
            }
          }
        }
        sh.in = D.in; sh.iu = D.iu;
        if (D.bad) sh.bad = 1;
      }
      __syncthreads();
      if (coupled_step) {                                             // identical <- all(chain_state1 == chain_state2), :99
        const int eq = (tid < dim) ? (st1[tid] == st2[tid]) : 1;
        if (__syncthreads_and(eq)) { met = true; tau = time; }
      }
      // ---- estimator / trajectories
      if (tid < dim) {
        const double xa = st1[tid];
        const double xb = (time > P.lag && !coupled_step) ? xa : st2[tid];   // after meeting chain 2 copies chain 1
        if (P.mode == MODE_ESTIMATOR) {
          const bool in_window = (time <= P.lag) ? (time >= P.k) : (P.k <= time && time <= P.m);
          const bool corr_now = (time >= P.k + P.lag) && (coupled_step || time == P.lag);
          const long long wi = corr_now ? corr_weight(time, P.k, P.m, P.lag) : 0;
          const double wgt = (double)wi;
          const int nh = (P.h_kind == UBM_H_BOTH) ? 2 : 1;
          for (int hh = 0; hh < nh; ++hh) {
            const bool sq = (P.h_kind == UBM_H_SQUARE) || hh == 1;
            const int e = hh * dim + tid;
            const double h1 = sq ? xa * xa : xa;
            if (in_window) acc[e] = acc[e] + h1;
            if (corr_now) { const double h2 = sq ? xb * xb : xb; acc[2 * BL_MAXD + e] = acc[2 * BL_MAXD + e] + wgt * (h1 - h2); }
          }
          if (bn) {
            const int b1 = bin_of(xa, P.breaks, P.nbreaks, bin_b0, bin_iw);
            if (in_window && b1 >= 0) bn[b1] += 1;
            if (corr_now) {
              const int b2 = bin_of(xb, P.breaks, P.nbreaks, bin_b0, bin_iw);
              if (b1 >= 0) bn[b1] += wi;
              if (b2 >= 0) bn[b2] -= wi;
            }
          }
        } else if (P.mode == MODE_CHAINS) {
          P.samples1[(rep * P.stride + time) * dim + tid] = xa;
          if (time > P.lag) P.samples2[(rep * P.stride + (time - P.lag)) * dim + tid] = xb;
        }
      }
    }
    __syncthreads();
    // ---- results
    if (sh.bad) { met = false; if (tid == 0) atomicExch(P.status, UBM_ERR_TAPE); }
    if (tid == 0) {
      double cost = dinf();
      if (met) {
        if (P.mode == MODE_MEETING) cost = (double)(P.lag + 2 * (tau - P.lag));
        else cost = (double)(P.lag + 2 * (tau - P.lag) + ((time - tau) > 0 ? (time - tau) : 0));
      }
      if (P.tau) P.tau[rep] = met ? (double)tau : dinf();
      if (P.cost) P.cost[rep] = cost;
      if (P.iter) P.iter[rep] = time;
    }
    if (P.mode == MODE_ESTIMATOR && tid < dimh) {
      const double mc = acc[tid], co = acc[2 * BL_MAXD + tid];
      const double ue = mc + co;
      P.uest[rep * dimh + tid] = ue / denom;
      if (P.mcmc) P.mcmc[rep * dimh + tid] = mc / denom;
      if (P.corr) P.corr[rep * dimh + tid] = co / denom;
    }
  }
}

inline size_t blasso_smem_bytes(int p) {
  return (size_t)(5 * p * p + 7 * PG_MAXP + 2 * BL_MAXD + 4 * BL_MAXD) * sizeof(double) + sizeof(BlShared) + 16;
}

inline int blasso_launch(const BlassoDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st, int64_t* launches,
                         std::string* err) {
  const size_t smem = blasso_smem_bytes(M.p);
  if (smem > 227 * 1024) { *err = "bayesian lasso: p too large for shared memory"; return UBM_ERR_UNSUPPORTED; }
  auto kern = tape ? blasso_driver_kernel<true> : blasso_driver_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = std::string("bayesian lasso: smem attribute: ") + cudaGetErrorString(e); return UBM_ERR_CUDA; }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PG_T, smem);
  if (e != cudaSuccess || per_sm < 1) { *err = "bayesian lasso: occupancy query failed"; return UBM_ERR_CUDA; }
  const int grid = (int)std::min<int64_t>((int64_t)nsm * per_sm, P.nrep);
  kern<<<grid, PG_T, smem, st>>>(M, P);
  *launches += 1;
  return UBM_OK;
}

}  // namespace ubm
