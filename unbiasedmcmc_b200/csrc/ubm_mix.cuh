// ubm_mix.cuh — C1: univariate Normal-mixture target, RW-MH, one THREAD per replicate.
//
// Replaces, fused into one persistent kernel:
//   drivers      R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//   kernels      R/get_mh_kernels.R:29-77 (d = 1), inst/reproducebimodal/bimodal.run.R:14-57
//   couplings    R/rnorm_reflectionmax.R:18-39, R/get_max_coupling.R:24-39 + R/rnorm_max_coupling.R:17-23
//   estimators   R/H_bar.R (on the fly), src/estimator_bin.cpp:10-42 (on the fly, all bins)
//
// Layout: a warp takes 32 consecutive replicate indices from a global counter (work stealing), runs them to
// completion warp-synchronously, writes per-replicate results, takes the next 32.  Chain state, draw
// counters and H accumulators live in registers; the per-thread histogram (int32 numerators, [bin][thread])
// lives in shared memory, so nothing but O(1) numbers per replicate reaches HBM.
#pragma once
#include "ubm_common.cuh"

namespace ubm {

struct MixParams {
  int ncomp, coupling;
  double logw[8], mean[8], sd[8], logsd[8];
  double prop_sd, log_prop_sd, init_mean, init_sd;
};

#define UBM_LN_SQRT_2PI 0.918938533204672741780329736406

struct MixState { double x, pdf; };

struct MixModel {
  const MixParams& P;
  __device__ explicit MixModel(const MixParams& p) : P(p) {}

  // R dnorm(x, mu, sigma, log = TRUE)
  __device__ static double dnorm_log(double x, double mu, double sigma, double log_sigma) {
    double z = (x - mu) / sigma;
    return -((UBM_LN_SQRT_2PI + 0.5 * z * z) + log_sigma);
  }
  // README.Rmd:69-72
  __device__ __noinline__ double target(double x) const {
    double ev[8];
    double mx = -dinf();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < P.ncomp) {
        ev[i] = P.logw[i] + dnorm_log(x, P.mean[i], P.sd[i], P.logsd[i]);
        if (ev[i] > mx) mx = ev[i];
      }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < P.ncomp) {
        double a = ev[i] - mx;
        s = s + ((a == 0.0) ? 1.0 : ool_exp(a));   // ool_exp(0) == 1 exactly: skipping it changes no bit
      }
    }
    return mx + ool_log(s);
  }
  template <class D>
  __device__ void rinit(MixState& s, D& d) const {          // README.Rmd:80-84
    s.x = P.init_mean + P.init_sd * d.n();
    s.pdf = target(s.x);
  }
  template <class D>
  __device__ void single(MixState& s, D& d) const {         // R/get_mh_kernels.R:29-40
    double prop = d.n() * P.prop_sd + s.x;
    double ppdf = target(prop);
    double logu = ool_log(d.u());
    if (logu < ppdf - s.pdf) { s.x = prop; s.pdf = ppdf; }
  }
  template <class D>
  __device__ bool coupled(MixState& s1, MixState& s2, D& d) const {   // R/get_mh_kernels.R:42-77
    double p1, p2;
    bool ident;
    if (P.coupling == UBM_COUPLING_REFLECTION_MAX) {        // R/rnorm_reflectionmax.R:18-39
      double xdot = d.n();
      double z = (s1.x - s2.x) / P.prop_sd;
      double normz = sqrt(z * z);
      double e = z / normz;
      double utilde = d.u();
      double a = xdot + z;
      ident = ool_log(utilde) <
              (-((UBM_LN_SQRT_2PI + 0.5 * a * a) + 0.0)) - (-((UBM_LN_SQRT_2PI + 0.5 * xdot * xdot) + 0.0));
      double ydot = ident ? (xdot + z) : (xdot - 2 * (e * xdot) * e);
      p1 = s1.x + P.prop_sd * xdot;
      p2 = s2.x + P.prop_sd * ydot;
    } else {                                                // R/get_max_coupling.R:24-39
      double sg = P.prop_sd, lsg = P.log_prop_sd;
      p1 = s1.x + sg * d.n();
      if (dnorm_log(p1, s1.x, sg, lsg) + ool_log(d.u()) < dnorm_log(p1, s2.x, sg, lsg)) {
        ident = true; p2 = p1;
      } else {
        ident = false;
        bool reject = true;
        p2 = 0.0;
        while (reject && !d.bad) {
          p2 = s2.x + sg * d.n();
          reject = dnorm_log(p2, s2.x, sg, lsg) + ool_log(d.u()) < dnorm_log(p2, s1.x, sg, lsg);
        }
      }
    }
    double pdf1 = target(p1), pdf2;
    if (ident) { p2 = p1; pdf2 = pdf1; } else pdf2 = target(p2);
    double logu = ool_log(d.u());
    bool acc1 = false, acc2 = false;
    if (isfinite(pdf1)) acc1 = logu < pdf1 - s1.pdf;
    if (isfinite(pdf2)) acc2 = logu < pdf2 - s2.pdf;
    if (acc1) { s1.x = p1; s1.pdf = pdf1; }
    if (acc2) { s2.x = p2; s2.pdf = pdf2; }
    return ident && acc1 && acc2;
  }
};

__device__ inline void h_eval1(int h_kind, double x, double* h) {
  if (h_kind == UBM_H_IDENTITY) h[0] = x;
  else if (h_kind == UBM_H_SQUARE) h[0] = x * x;
  else { h[0] = x; h[1] = x * x; }
}

// One thread per replicate.  Shared memory: int hist[nbins][blockDim.x] (bins on) + breaks.
template <bool TAPE>
#ifndef MIX_MINB
#define MIX_MINB 5      // CTAs per SM asked of the register allocator (96 registers): 3 -> 5.5e6, 4 -> 6.1e6, 5 -> 6.6e6 est/s on C1
#endif
__global__ void __launch_bounds__(128, MIX_MINB) mix_driver_kernel(const __grid_constant__ MixParams MP, const RunParams P) {
  extern __shared__ unsigned char smem_raw[];
  const int nbins = (P.mode == MODE_ESTIMATOR && P.nbreaks >= 2) ? P.nbreaks - 1 : 0;
  double* s_breaks = reinterpret_cast<double*>(smem_raw);
  int* s_hist = reinterpret_cast<int*>(smem_raw + sizeof(double) * ((P.nbreaks + 1) & ~1));
  const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
  for (int i = tid; i < P.nbreaks; i += nthr) s_breaks[i] = P.breaks[i];
  __syncthreads();
  double b0 = 0, inv_w = 0;
  if (nbins) { b0 = s_breaks[0]; inv_w = (double)nbins / (s_breaks[nbins] - s_breaks[0]); }
  const MixModel model(MP);
  const int dimh = (P.h_kind == UBM_H_BOTH) ? 2 : 1;
  const double denom = (double)(P.m - P.k + 1);

  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.work_counter, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned long long)P.nrep) break;
    const int64_t r = (int64_t)base + lane;
    const bool active = r < P.nrep;
    bool done = !active;

    ScalarDraws<TAPE, true> D;
    D.init(P, active ? r : 0);
    MixState s1, s2;
    double mc[2] = {0, 0}, co[2] = {0, 0}, h1[2], h2[2];
    int64_t time = 0, tau = 0;
    bool met = false;
    for (int b = 0; b < nbins; ++b) s_hist[b * nthr + tid] = 0;
    auto bin_add = [&](double x, int w) {
      if (nbins) {
        int b = bin_of(x, s_breaks, P.nbreaks, b0, inv_w);
        if (b >= 0) s_hist[b * nthr + tid] += w;
      }
    };
    double* S1 = nullptr; double* S2 = nullptr;
    if (P.mode == MODE_CHAINS && active) { S1 = P.samples1 + r * P.stride; S2 = P.samples2 + r * P.stride; }

    if (active) {
      model.rinit(s1, D);
      model.rinit(s2, D);
      if (P.mode == MODE_ESTIMATOR) {
        if (P.k == 0) { h_eval1(P.h_kind, s1.x, h1); mc[0] = h1[0]; mc[1] = (dimh > 1) ? h1[1] : 0.0; bin_add(s1.x, 1); }
      } else if (P.mode == MODE_CHAINS) { S1[0] = s1.x; S2[0] = s2.x; }
      // lag single steps
      for (int64_t t = 0; t < P.lag; ++t) {
        time += 1;
        model.single(s1, D);
        if (P.mode == MODE_ESTIMATOR) {
          if (time >= P.k) {
            h_eval1(P.h_kind, s1.x, h1);
            mc[0] = mc[0] + h1[0]; if (dimh > 1) mc[1] = mc[1] + h1[1];
            bin_add(s1.x, 1);
          }
        } else if (P.mode == MODE_CHAINS) S1[time] = s1.x;
      }
      if (P.mode == MODE_ESTIMATOR && time >= P.k + P.lag) {
        int64_t wi = corr_weight(time, P.k, P.m, P.lag);
        double w = (double)wi;
        h_eval1(P.h_kind, s1.x, h1); h_eval1(P.h_kind, s2.x, h2);
        co[0] = co[0] + w * (h1[0] - h2[0]); if (dimh > 1) co[1] = co[1] + w * (h1[1] - h2[1]);
        bin_add(s1.x, (int)wi); bin_add(s2.x, -(int)wi);
      }
    }
    // main loop, warp-synchronous: every lane advances one time step per trip until all lanes are done
    for (;;) {
      if (!done) {
        bool go;
        if (P.mode == MODE_MEETING) go = !met && (P.max_it <= 0 || time < P.max_it);
        else go = (!met || time < (tau > P.m ? tau : P.m)) && (P.max_it <= 0 || time < P.max_it);
        if (D.bad) go = false;
        if (!go) done = true;
      }
      if (__all_sync(0xffffffffu, done)) break;
      if (!done) {
        time += 1;
        if (met) {
          model.single(s1, D);
          s2 = s1;
          if (P.mode == MODE_ESTIMATOR && P.k <= time && time <= P.m) {
            h_eval1(P.h_kind, s1.x, h1);
            mc[0] = mc[0] + h1[0]; if (dimh > 1) mc[1] = mc[1] + h1[1];
            bin_add(s1.x, 1);
          }
        } else {
          bool ident = model.coupled(s1, s2, D);
          if (ident) { met = true; tau = time; }
          if (P.mode == MODE_ESTIMATOR) {
            if (P.k <= time && time <= P.m) {
              h_eval1(P.h_kind, s1.x, h1);
              mc[0] = mc[0] + h1[0]; if (dimh > 1) mc[1] = mc[1] + h1[1];
              bin_add(s1.x, 1);
            }
            if (time >= P.k + P.lag) {
              int64_t wi = corr_weight(time, P.k, P.m, P.lag);
              double w = (double)wi;
              h_eval1(P.h_kind, s1.x, h1); h_eval1(P.h_kind, s2.x, h2);
              co[0] = co[0] + w * (h1[0] - h2[0]); if (dimh > 1) co[1] = co[1] + w * (h1[1] - h2[1]);
              bin_add(s1.x, (int)wi); bin_add(s2.x, -(int)wi);
            }
          }
        }
        if (P.mode == MODE_CHAINS) { S1[time] = s1.x; S2[time - P.lag] = s2.x; }
      }
    }
    // results
    if (active) {
      if (D.bad) atomicExch(P.status, UBM_ERR_TAPE);
      double taud = met ? (double)tau : dinf();
      double costd = dinf();
      if (met) {
        if (P.mode == MODE_MEETING) costd = (double)(P.lag + 2 * (tau - P.lag));
        else costd = (double)(P.lag + 2 * (tau - P.lag) + ((time - tau) > 0 ? (time - tau) : 0));
      }
      if (P.tau) P.tau[r] = taud;
      if (P.cost) P.cost[r] = costd;
      if (P.iter) P.iter[r] = time;
      if (P.mode == MODE_ESTIMATOR) {
        for (int i = 0; i < dimh; ++i) {
          double ue = mc[i] + co[i];
          if (P.uest) P.uest[r * dimh + i] = ue / denom;
          if (P.mcmc) P.mcmc[r * dimh + i] = mc[i] / denom;
          if (P.corr) P.corr[r * dimh + i] = co[i] / denom;
        }
        if (P.bins_out)
          for (int b = 0; b < nbins; ++b) P.bins_out[r * nbins + b] = (double)s_hist[b * nthr + tid] / denom;
      }
    }
    // integer bin numerators of the uncensored replicates: warp-reduce, then one atomic per bin per warp.
    // Integer sums are exact, so the aggregate does not depend on the hand-out order.
    if (nbins) {
      const bool counted = active && met;
      for (int b = 0; b < nbins; ++b) {
        long long c = counted ? (long long)s_hist[b * nthr + tid] : 0ll;
        long long c2 = c * c;
        if (__any_sync(0xffffffffu, c != 0)) {
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, off);
            c2 += __shfl_xor_sync(0xffffffffu, c2, off);
          }
          if (lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(P.bin_acc) + b, (unsigned long long)c);
            atomicAdd(reinterpret_cast<unsigned long long*>(P.bin_acc) + nbins + b, (unsigned long long)c2);
          }
        }
      }
    }
  }
}

}  // namespace ubm
