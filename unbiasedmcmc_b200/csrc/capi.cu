// capi.cu — the extern "C" boundary of libubm.so (include/ubm.h) and the host-side plumbing behind it:
// handle (device, stream, error text), plan (uploaded model + device result buffers), launches.
// Replaces the R `.Call` boundary of the reference (src/RcppExports.cpp:388-423) for the coupled-chain
// hot path.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ubm_common.cuh"
#include "ubm_mix.cuh"
#include "ubm_mvn.cuh"
#include "ubm_mvn_single.cuh"
#include "ubm_ising.cuh"
#include "ubm_pg.cuh"
#include "ubm_varsel.cuh"
#include "ubm_summary.cuh"
#include "ubm_post.cuh"
#include "ubm_peaks.cuh"
#include "ubm_prims.cuh"
#include "ubm_blasso.cuh"

using namespace ubm;

struct ubm_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  cudaDeviceProp prop{};
  long long* pt_counts = nullptr;     // device [2 * ISING_MAXN]: swap-move counters of the last Ising run (R/ising.R:48-49,89-91)
  int pt_n = 0;
};

static int fail(ubm_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, UBM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) {
    if (p) { cudaFree(p); p = nullptr; }
    bytes = n;
    if (n == 0) return cudaSuccess;
    return cudaMalloc(&p, n);
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct ubm_plan {
  ubm_handle* h = nullptr;
  int model_kind = 0, mode = MODE_ESTIMATOR, h_kind = 0;
  int dim = 0, dimh = 0, nbins = 0, nbreaks = 0, component = 0;
  int64_t k = 0, m = 1, lag = 1, max_it = 0, max_nrep = 0, stride = 0;
  // models
  MixParams mix{};
  MvnDevice mvn{};
  IsingDevice ising{};
  PgDevice pg{};
  VarselDevice varsel{};
  BlassoDevice blasso{};
  DevBuf model_blob;
  // results
  DevBuf uest, mcmc, corr, tau, cost, iter, bins_out, bin_acc, bin_num, breaks, summary, counter, status, scratch;
  DevBuf tape_u, tape_n, tape_e, samples1, samples2;
  int64_t len_u = 0, len_n = 0, len_e = 0;
  bool want_bins_out = false, want_parts = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  int64_t last_nrep = 0;
  ~ubm_plan() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
  }
};

// ---------------------------------------------------------------------------------------------------
// model registration (host side of the device registry)
// ---------------------------------------------------------------------------------------------------
static int model_dims(int32_t kind, const void* model, int* dim) {
  if (!model) return UBM_ERR_INVALID;
  switch (kind) {
    case UBM_MODEL_NORMAL_MIXTURE: *dim = 1; return UBM_OK;
    case UBM_MODEL_MVN: *dim = ((const ubm_model_mvn*)model)->d; return UBM_OK;
    case UBM_MODEL_ISING_PT: *dim = ((const ubm_model_ising_pt*)model)->nchains; return UBM_OK;
    case UBM_MODEL_PG_LOGISTIC: *dim = ((const ubm_model_pg_logistic*)model)->p; return UBM_OK;
    case UBM_MODEL_VARSEL: *dim = ((const ubm_model_varsel*)model)->p; return UBM_OK;
    case UBM_MODEL_BLASSO: *dim = 2 * ((const ubm_model_blasso*)model)->p + 1; return UBM_OK;
    default: return UBM_ERR_UNSUPPORTED;
  }
}

static int register_model(ubm_plan* p, int32_t kind, const void* model) {
  ubm_handle* h = p->h;
  p->model_kind = kind;
  if (kind == UBM_MODEL_NORMAL_MIXTURE) {
    const ubm_model_normal_mixture* s = (const ubm_model_normal_mixture*)model;
    if (s->ncomp < 1 || s->ncomp > 8) return fail(h, UBM_ERR_INVALID, "normal mixture: 1..8 components");
    if (s->coupling != UBM_COUPLING_REFLECTION_MAX && s->coupling != UBM_COUPLING_MAX)
      return fail(h, UBM_ERR_UNSUPPORTED, "normal mixture: unknown coupling");
    if (!(s->proposal_sd > 0)) return fail(h, UBM_ERR_INVALID, "normal mixture: proposal_sd must be > 0");
    MixParams& M = p->mix;
    M.ncomp = s->ncomp; M.coupling = s->coupling;
    for (int i = 0; i < s->ncomp; ++i) {
      M.logw[i] = ubm_log(s->weights[i]); M.mean[i] = s->means[i]; M.sd[i] = s->sds[i]; M.logsd[i] = ubm_log(s->sds[i]);
    }
    M.prop_sd = s->proposal_sd; M.log_prop_sd = ubm_log(s->proposal_sd);
    M.init_mean = s->init_mean; M.init_sd = s->init_sd;
    p->dim = 1;
    return UBM_OK;
  }
  if (kind == UBM_MODEL_MVN) {
    const ubm_model_mvn* s = (const ubm_model_mvn*)model;
    if (s->d < 2 || s->d > 128) return fail(h, UBM_ERR_INVALID, "mvn: need 2 <= d <= 128 (R stops for d = 1, R/mvnorm_couplings.R:96)");
    p->dim = s->d;
    std::vector<double> blob;
    mvn_pack_host(s, &p->mvn, &blob);
    CU(p->model_blob.alloc(blob.size() * sizeof(double)));
    CU(cudaMemcpyAsync(p->model_blob.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    mvn_bind_device(&p->mvn, p->model_blob.as<double>());
    return UBM_OK;
  }
  if (kind == UBM_MODEL_ISING_PT) {
    const ubm_model_ising_pt* s = (const ubm_model_ising_pt*)model;
    if (s->size != 32) return fail(h, UBM_ERR_UNSUPPORTED, "ising: the device kernel is specialised for 32 x 32 lattices");
    if (s->nchains < 1 || s->nchains > ISING_MAXN) return fail(h, UBM_ERR_INVALID, "ising: 1..32 temperatures");
    if (s->scan != UBM_ISING_SCAN_SEQUENTIAL) return fail(h, UBM_ERR_UNSUPPORTED, "ising: unknown scan order");
    if (!(s->proba_swapmove >= 0.0 && s->proba_swapmove <= 1.0)) return fail(h, UBM_ERR_INVALID, "ising: proba_swapmove in [0,1]");
    ising_pack_host(s, &p->ising);
    p->dim = s->nchains;
    return UBM_OK;
  }
  if (kind == UBM_MODEL_PG_LOGISTIC) {
    const ubm_model_pg_logistic* s = (const ubm_model_pg_logistic*)model;
    if (s->p < 1 || s->p > PG_PLIM || s->n < 1 || s->n > 4096) return fail(h, UBM_ERR_INVALID, "pg logistic: 1 <= p <= 60, 1 <= n <= 4096");
    if (s->w_coupling != 0 && s->w_coupling != 1) return fail(h, UBM_ERR_INVALID, "pg logistic: w_coupling must be 0 ('max') or 1 ('rej_samp')");
    if (s->beta_coupling != 0 && s->beta_coupling != 1) return fail(h, UBM_ERR_INVALID, "pg logistic: beta_coupling must be 0 (maximal coupling) or 1 (common random numbers)");
    if (!s->X || !s->Y || !s->b || !s->B) return fail(h, UBM_ERR_INVALID, "pg logistic: null data");
    p->dim = s->p;
    std::vector<double> blob;
    pg_pack_host(s, &p->pg, &blob);
    CU(p->model_blob.alloc(blob.size() * sizeof(double)));
    CU(cudaMemcpyAsync(p->model_blob.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    pg_bind_device(&p->pg, p->model_blob.as<double>());
    return UBM_OK;
  }
  if (kind == UBM_MODEL_BLASSO) {
    const ubm_model_blasso* s = (const ubm_model_blasso*)model;
    if (s->p < 1 || s->p > PG_PLIM || s->n < 3 || s->n > 4096) return fail(h, UBM_ERR_INVALID, "bayesian lasso: 1 <= p <= 60, 3 <= n <= 4096");
    if (!s->X || !s->Y) return fail(h, UBM_ERR_INVALID, "bayesian lasso: null data");
    if (!(s->lambda > 0.0)) return fail(h, UBM_ERR_INVALID, "bayesian lasso: lambda > 0");
    p->dim = 2 * s->p + 1;
    std::vector<double> blob;
    blasso_pack_host(s, &p->blasso, &blob);
    CU(p->model_blob.alloc(blob.size() * sizeof(double)));
    CU(cudaMemcpyAsync(p->model_blob.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    blasso_bind_device(&p->blasso, p->model_blob.as<double>());
    return UBM_OK;
  }
  if (kind == UBM_MODEL_VARSEL) {
    const ubm_model_varsel* s = (const ubm_model_varsel*)model;
    if (s->p < 2 || s->p > 1024 || s->n < 1 || s->s0 < 1 || s->s0 > 128) return fail(h, UBM_ERR_INVALID, "variable selection: 2 <= p <= 1024, 1 <= s0 <= 128");
    if (!s->X || !s->Y) return fail(h, UBM_ERR_INVALID, "variable selection: null data");
    if (!(s->proportion_singleflip >= 0.0 && s->proportion_singleflip <= 1.0)) return fail(h, UBM_ERR_INVALID, "variable selection: proportion_singleflip in [0,1]");
    VarselDevice& V = p->varsel;
    V.n = s->n; V.p = s->p; V.s0 = s->s0 < s->p ? s->s0 : s->p; V.nwords = (s->p + 31) / 32;
    V.g = s->g; V.kappa = s->kappa; V.prop_flip = s->proportion_singleflip;
    double y2 = 0.0;
    for (int i = 0; i < s->n; ++i) y2 = ubm_fma(s->Y[i], s->Y[i], y2);          // R/variableselection.R:6
    V.Y2 = y2; V.lg1 = ubm_log(s->g + 1.); V.gg = s->g / (s->g + 1.); V.logp = ubm_log((double)s->p);
    p->dim = s->p;
    DevBuf dX, dY;
    const size_t np_ = (size_t)s->n * s->p, pp_ = (size_t)s->p * s->p;
    CU(dX.alloc(np_ * 8)); CU(dY.alloc((size_t)s->n * 8));
    CU(p->model_blob.alloc((pp_ + s->p) * 8));
    CU(cudaMemcpyAsync(dX.p, s->X, np_ * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dY.p, s->Y, (size_t)s->n * 8, cudaMemcpyHostToDevice, h->stream));
    double* G = p->model_blob.as<double>();
    varsel_gram_kernel<<<(unsigned)((pp_ + 255) / 256), 256, 0, h->stream>>>(dX.as<double>(), dY.as<double>(), s->n, s->p, G, G + pp_);
    h->launches += 1;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    V.G = G; V.v = G + pp_;
    return UBM_OK;
  }
  return fail(h, UBM_ERR_UNSUPPORTED, "model kind not in the device registry (no CPU fallback)");
}

static int plan_build(ubm_handle* h, int mode, int32_t model_kind, const void* model, int32_t h_kind,
                      const ubm_bins* bins, int64_t k, int64_t m, int64_t lag, int64_t max_it, int64_t max_nrep,
                      int64_t stride, ubm_plan** out) {
  if (!h) return UBM_ERR_INVALID;
  if (!model || !out) return fail(h, UBM_ERR_INVALID, "null model / out");
  if (lag < 1) return fail(h, UBM_ERR_INVALID, "lag must be >= 1 (R/coupled_chains.R:50)");
  if (max_nrep < 1) return fail(h, UBM_ERR_INVALID, "nrep must be >= 1");
  if (mode == MODE_ESTIMATOR) {
    if (k < 0) return fail(h, UBM_ERR_INVALID, "k must be >= 0");
    if (k > m) return fail(h, UBM_ERR_INVALID, "error: k has to be less than m");       // R/unbiasedestimator.R:44-47
    if (lag > m) return fail(h, UBM_ERR_INVALID, "error: lag has to be less than m");   // :48-51
  }
  if (h_kind < UBM_H_IDENTITY || h_kind > UBM_H_BOTH) return fail(h, UBM_ERR_UNSUPPORTED, "unknown test function h");
  CU(cudaSetDevice(h->device));
  ubm_plan* p = new ubm_plan();
  p->h = h; p->mode = mode; p->h_kind = h_kind;
  p->k = k; p->m = m; p->lag = lag; p->max_it = max_it; p->max_nrep = max_nrep; p->stride = stride;
  int st = register_model(p, model_kind, model);
  if (st != UBM_OK) { delete p; return st; }
  p->dimh = (h_kind == UBM_H_BOTH) ? 2 * p->dim : p->dim;
  // combinations the device registry does not hold fail here, before anything is allocated or enqueued
  if (mode == MODE_ESTIMATOR && h_kind != UBM_H_IDENTITY && (model_kind == UBM_MODEL_ISING_PT || model_kind == UBM_MODEL_VARSEL)) {
    delete p;
    return fail(h, UBM_ERR_UNSUPPORTED, "this model's state is integer valued (natural statistics / inclusion indicators): h must be the identity");
  }
  if (mode == MODE_ESTIMATOR && bins && bins->nbreaks >= 2) {
    if (bins->nbreaks > 257) { delete p; return fail(h, UBM_ERR_INVALID, "at most 256 bins"); }
    if (bins->component < 0 || bins->component >= p->dim) { delete p; return fail(h, UBM_ERR_INVALID, "bins: component out of range"); }
    for (int i = 1; i < bins->nbreaks; ++i)
      if (!(bins->breaks[i] > bins->breaks[i - 1])) { delete p; return fail(h, UBM_ERR_INVALID, "bins: breaks must increase"); }
    p->nbreaks = bins->nbreaks; p->nbins = bins->nbreaks - 1; p->component = bins->component;
  }
  auto alloc_all = [&]() -> cudaError_t {
    cudaError_t e;
    size_t n = (size_t)max_nrep;
    if ((e = p->tau.alloc(n * 8)) != cudaSuccess) return e;
    if ((e = p->cost.alloc(n * 8)) != cudaSuccess) return e;
    if ((e = p->iter.alloc(n * 8)) != cudaSuccess) return e;
    if (mode == MODE_ESTIMATOR) {
      if ((e = p->uest.alloc(n * p->dimh * 8)) != cudaSuccess) return e;
      if ((e = p->mcmc.alloc(n * p->dimh * 8)) != cudaSuccess) return e;
      if ((e = p->corr.alloc(n * p->dimh * 8)) != cudaSuccess) return e;
    }
    if (p->nbins) {
      if ((e = p->bin_acc.alloc((size_t)2 * p->nbins * 8)) != cudaSuccess) return e;
      if ((e = p->breaks.alloc((size_t)p->nbreaks * 8)) != cudaSuccess) return e;
      if (model_kind != UBM_MODEL_NORMAL_MIXTURE && (e = p->bin_num.alloc(n * p->nbins * 8)) != cudaSuccess) return e;
      if ((e = cudaMemcpyAsync(p->breaks.p, bins->breaks, (size_t)p->nbreaks * 8, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return e;
    }
    if ((e = p->summary.alloc((size_t)ubm_summary_len(p->dimh, p->nbins) * 8)) != cudaSuccess) return e;
    if ((e = p->counter.alloc(8)) != cudaSuccess) return e;
    if ((e = p->status.alloc(4)) != cudaSuccess) return e;
    if (mode == MODE_CHAINS) {
      size_t sz = n * (size_t)stride * p->dim * 8;
      if ((e = p->samples1.alloc(sz)) != cudaSuccess) return e;
      if ((e = p->samples2.alloc(sz)) != cudaSuccess) return e;
    }
    if ((e = cudaEventCreate(&p->e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&p->e1)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&p->e2)) != cudaSuccess) return e;
    return cudaSuccess;
  };
  cudaError_t e = alloc_all();
  if (e != cudaSuccess) { delete p; return fail(h, e == cudaErrorMemoryAllocation ? UBM_ERR_NOMEM : UBM_ERR_CUDA, std::string("plan alloc: ") + cudaGetErrorString(e)); }
  *out = p;
  return UBM_OK;
}

static int plan_set_tapes(ubm_plan* p, const ubm_draws* d, int64_t nrep) {
  ubm_handle* h = p->h;
  auto up = [&](DevBuf& b, const double* src, int64_t len) -> cudaError_t {
    if (!src || len <= 0) return b.alloc(0);
    cudaError_t e = b.alloc((size_t)nrep * len * 8);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(b.p, src, (size_t)nrep * len * 8, cudaMemcpyHostToDevice, h->stream);
  };
  CU(up(p->tape_u, d->tape_u, d->len_u));
  CU(up(p->tape_n, d->tape_n, d->len_n));
  CU(up(p->tape_e, d->tape_e, d->len_e));
  p->len_u = d->tape_u ? d->len_u : 0; p->len_n = d->tape_n ? d->len_n : 0; p->len_e = d->tape_e ? d->len_e : 0;
  return UBM_OK;
}

static int plan_run(ubm_plan* p, int draws_mode, uint64_t seed, int64_t nrep, int64_t rep_offset) {
  ubm_handle* h = p->h;
  if (nrep < 1 || nrep > p->max_nrep) return fail(h, UBM_ERR_INVALID, "nrep outside the plan's capacity");
  CU(cudaSetDevice(h->device));
  RunParams P{};
  P.mode = p->mode; P.h_kind = p->h_kind;
  P.k = p->k; P.m = p->m; P.lag = p->lag; P.max_it = p->max_it;
  if (p->mode == MODE_CHAINS) {   // storage cap acts as max_iterations (include/ubm.h)
    int64_t cap = p->stride - 1;
    if (P.max_it <= 0 || P.max_it > cap) P.max_it = cap;
  }
  P.nrep = nrep; P.rep_offset = rep_offset;
  P.draws_mode = draws_mode; P.seed = seed;
  P.tape_u = p->tape_u.as<double>(); P.tape_n = p->tape_n.as<double>(); P.tape_e = p->tape_e.as<double>();
  P.len_u = p->len_u; P.len_n = p->len_n; P.len_e = p->len_e;
  P.component = p->component; P.nbreaks = p->nbreaks; P.breaks = p->breaks.as<double>();
  P.uest = p->uest.as<double>(); P.mcmc = p->mcmc.as<double>(); P.corr = p->corr.as<double>();
  P.tau = p->tau.as<double>(); P.cost = p->cost.as<double>(); P.iter = p->iter.as<int64_t>();
  P.bins_out = p->want_bins_out ? p->bins_out.as<double>() : nullptr;
  P.bin_acc = p->bin_acc.as<long long>();
  P.bin_num = p->bin_num.as<long long>();
  P.samples1 = p->samples1.as<double>(); P.samples2 = p->samples2.as<double>(); P.stride = p->stride;
  P.work_counter = p->counter.as<unsigned long long>(); P.status = p->status.as<int>();
  if (p->model_kind == UBM_MODEL_ISING_PT) {
    if (!h->pt_counts) CU(cudaMalloc(&h->pt_counts, 2 * ISING_MAXN * sizeof(long long)));
    CU(cudaMemsetAsync(h->pt_counts, 0, 2 * ISING_MAXN * sizeof(long long), h->stream));
    P.pt_counts = h->pt_counts; h->pt_n = p->ising.N;
  }
  cudaStream_t st = h->stream;
  CU(cudaMemsetAsync(p->counter.p, 0, 8, st));
  CU(cudaMemsetAsync(p->status.p, 0, 4, st));
  if (p->nbins) CU(cudaMemsetAsync(p->bin_acc.p, 0, p->bin_acc.bytes, st));
  if (p->nbins && p->bin_num.p) CU(cudaMemsetAsync(p->bin_num.p, 0, (size_t)nrep * p->nbins * 8, st));
  const bool tape = draws_mode == UBM_DRAWS_TAPE;
  const int nsm = h->prop.multiProcessorCount;
  CU(cudaEventRecord(p->e0, st));
  if (p->model_kind == UBM_MODEL_NORMAL_MIXTURE) {
    const int threads = 128;
    size_t smem = sizeof(double) * ((p->nbreaks + 1) & ~1) + (size_t)p->nbins * threads * sizeof(int);
    auto kern = tape ? mix_driver_kernel<true> : mix_driver_kernel<false>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail(h, UBM_ERR_INVALID, "normal mixture: too many bins for shared memory");
    int64_t need = (nrep + threads - 1) / threads;
    int grid = (int)std::min<int64_t>((int64_t)nsm * per_sm, need);
    kern<<<grid, threads, smem, st>>>(p->mix, P);
    h->launches += 1;
  } else if (p->model_kind == UBM_MODEL_MVN) {
    // estimator runs on Philox streams take two passes: coupled phase, then the single-chain continuation with twice the slots
    const bool two_pass = p->mode == MODE_ESTIMATOR && !tape && mvn_two_pass_fits(p->mvn, P) && !getenv("UBM_MVN_ONE_PASS");
    int stl = two_pass ? mvn_launch_two_pass(p->mvn, P, nsm, st, p->scratch, &h->launches, &h->err)
                       : mvn_launch(p->mvn, P, tape, nsm, st, p->scratch, &h->launches, &h->err);
    if (stl != UBM_OK) return stl;
  } else if (p->model_kind == UBM_MODEL_VARSEL) {
    int stl = varsel_launch(p->varsel, P, tape, nsm, st, p->scratch, &h->launches, &h->err);
    if (stl != UBM_OK) return stl;
  } else if (p->model_kind == UBM_MODEL_PG_LOGISTIC) {
    int stl = pg_launch(p->pg, P, tape, nsm, st, &h->launches, &h->err);
    if (stl != UBM_OK) return stl;
  } else if (p->model_kind == UBM_MODEL_BLASSO) {
    int stl = blasso_launch(p->blasso, P, tape, nsm, st, &h->launches, &h->err);
    if (stl != UBM_OK) return stl;
  } else if (p->model_kind == UBM_MODEL_ISING_PT) {
    int stl = ising_launch(p->ising, P, tape, nsm, st, &h->launches, &h->err);
    if (stl != UBM_OK) return stl;
  } else {
    return fail(h, UBM_ERR_UNSUPPORTED, "model kind not in the device registry");
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(p->e1, st));
  if (p->nbins && p->bin_num.p) {
    bins_finalize_kernel<<<p->nbins, 256, 0, st>>>(nrep, p->nbins, P.tau, P.bin_num, (double)(p->m - p->k + 1), P.bins_out, P.bin_acc);
    h->launches += 1;
    CU(cudaGetLastError());
  }
  SummaryArgs A{};
  A.nrep = nrep; A.dimh = (p->mode == MODE_ESTIMATOR) ? p->dimh : 0; A.nbins = p->nbins;
  A.tau = P.tau; A.cost = P.cost; A.iter = P.iter; A.uest = (p->mode == MODE_ESTIMATOR) ? P.uest : nullptr;
  A.bin_acc = P.bin_acc; A.denom = (double)(p->m - p->k + 1); A.summary = p->summary.as<double>();
  if (p->mode == MODE_ESTIMATOR) {
    int cols = UBM_SUMMARY_HEAD + 2 * p->dimh + 2 * p->nbins;
    summary_kernel<<<cols, 1024, 0, st>>>(A);
  } else {
    summary_kernel<<<UBM_SUMMARY_HEAD, 1024, 0, st>>>(A);
  }
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaEventRecord(p->e2, st));
  p->last_nrep = nrep;
  return UBM_OK;
}

static int plan_fetch(ubm_plan* p, const ubm_estimator_out* out) {
  ubm_handle* h = p->h;
  cudaStream_t st = h->stream;
  CU(cudaStreamSynchronize(st));
  int status = 0;
  CU(cudaMemcpy(&status, p->status.p, 4, cudaMemcpyDeviceToHost));
  size_t n = (size_t)p->last_nrep;
  if (out) {
    if (p->mode == MODE_ESTIMATOR) {
      if (out->uestimator) CU(cudaMemcpy(out->uestimator, p->uest.p, n * p->dimh * 8, cudaMemcpyDeviceToHost));
      if (out->mcmcestimator) CU(cudaMemcpy(out->mcmcestimator, p->mcmc.p, n * p->dimh * 8, cudaMemcpyDeviceToHost));
      if (out->correction) CU(cudaMemcpy(out->correction, p->corr.p, n * p->dimh * 8, cudaMemcpyDeviceToHost));
      if (out->bins && p->nbins && p->want_bins_out)
        CU(cudaMemcpy(out->bins, p->bins_out.p, n * p->nbins * 8, cudaMemcpyDeviceToHost));
    }
    if (out->meetingtime) CU(cudaMemcpy(out->meetingtime, p->tau.p, n * 8, cudaMemcpyDeviceToHost));
    if (out->iteration) CU(cudaMemcpy(out->iteration, p->iter.p, n * 8, cudaMemcpyDeviceToHost));
    if (out->cost) CU(cudaMemcpy(out->cost, p->cost.p, n * 8, cudaMemcpyDeviceToHost));
    if (out->summary) {
      size_t len = (p->mode == MODE_ESTIMATOR) ? (size_t)ubm_summary_len(p->dimh, p->nbins) : (size_t)UBM_SUMMARY_HEAD;
      CU(cudaMemcpy(out->summary, p->summary.p, len * 8, cudaMemcpyDeviceToHost));
    }
  }
  if (status == UBM_ERR_TAPE) return fail(h, UBM_ERR_TAPE, "replay tape exhausted before a replicate finished");
  if (status != 0) return fail(h, status, "device-side failure");
  return UBM_OK;
}

// ---------------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------------
extern "C" {

int ubm_create(int device, ubm_handle** out) {
  if (!out) return UBM_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || device < 0 || device >= n) return UBM_ERR_CUDA;   // no CPU fallback
  ubm_handle* h = new ubm_handle();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&h->prop, device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return UBM_ERR_CUDA;
  }
  *out = h;
  return UBM_OK;
}

int ubm_destroy(ubm_handle* h) {
  if (!h) return UBM_OK;
  cudaSetDevice(h->device);
  if (h->pt_counts) cudaFree(h->pt_counts);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return UBM_OK;
}

const char* ubm_last_error(const ubm_handle* h) { return h ? h->err.c_str() : "null handle"; }

int ubm_device_info(const ubm_handle* h, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor) {
  if (!h) return UBM_ERR_INVALID;
  if (name && name_len > 0) { strncpy(name, h->prop.name, name_len - 1); name[name_len - 1] = 0; }
  if (sm_count) *sm_count = h->prop.multiProcessorCount;
  if (cc_major) *cc_major = h->prop.major;
  if (cc_minor) *cc_minor = h->prop.minor;
  return UBM_OK;
}

int64_t ubm_launch_count(const ubm_handle* h) { return h ? h->launches : 0; }

int ubm_measure_peaks(ubm_handle* h, double* fp64_tflops, double* int32_tops) {
  if (!h) return UBM_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  DevBuf out;
  CU(out.alloc(64));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  const int grid = h->prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float best_d = 1e30f, best_i = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0, h->stream));
    peak_dfma_kernel<<<grid, threads, 0, h->stream>>>(out.as<double>(), iters, 0.999999, 1e-9);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_d) best_d = ms;
    CU(cudaEventRecord(e0, h->stream));
    peak_imad_kernel<<<grid, threads, 0, h->stream>>>(out.as<unsigned>(), iters, 1664525u, 1013904223u);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_i) best_i = ms;
    h->launches += 2;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double ops = (double)grid * threads * iters * 64.0;   // fma / imad per launch
  if (fp64_tflops) *fp64_tflops = 2.0 * ops / (best_d * 1e-3) * 1e-12;
  if (int32_tops) *int32_tops = 2.0 * ops / (best_i * 1e-3) * 1e-12;
  return UBM_OK;
}

int ubm_measure_fp64_tensor_peak(ubm_handle* h, double* fp64_tensor_tflops) {
  if (!h || !fp64_tensor_tflops) return UBM_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  DevBuf out;
  CU(out.alloc(64));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  const int grid = h->prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0, h->stream));
    peak_dmma_kernel<<<grid, threads, 0, h->stream>>>(out.as<double>(), iters, 0.999999, 1e-9);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
    h->launches += 1;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double fma = (double)grid * (threads / 32) * iters * 8.0 * 256.0;   // 256 fma per m8n8k4
  *fp64_tensor_tflops = 2.0 * fma / (best * 1e-3) * 1e-12;
  return UBM_OK;
}

int ubm_ising_swap_counts(ubm_handle* h, int64_t* nswap_attempts, int64_t* nswap_accepts1, int64_t* nswap_accepts2) {
  if (!h) return UBM_ERR_INVALID;
  if (!h->pt_counts || h->pt_n < 1) return fail(h, UBM_ERR_INVALID, "no Ising parallel-tempering run on this handle yet");
  long long host[2 * ISING_MAXN];
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(host, h->pt_counts, sizeof(host), cudaMemcpyDeviceToHost));
  if (nswap_attempts) *nswap_attempts = host[0];
  for (int c = 0; c + 1 < h->pt_n; ++c) {
    if (nswap_accepts1) nswap_accepts1[c] = host[1 + c];
    if (nswap_accepts2) nswap_accepts2[c] = host[h->pt_n + c];
  }
  return UBM_OK;
}

int64_t ubm_summary_len(int32_t dimh, int32_t nbins) { return UBM_SUMMARY_HEAD + 2 * (int64_t)dimh + 2 * (int64_t)nbins; }

int ubm_model_dim(int32_t model_kind, const void* model, int32_t* dim) {
  if (!dim) return UBM_ERR_INVALID;
  return model_dims(model_kind, model, dim);
}
int ubm_h_dim(int32_t model_kind, const void* model, int32_t h_kind, int32_t* dimh) {
  int d = 0;
  int st = model_dims(model_kind, model, &d);
  if (st != UBM_OK) return st;
  if (h_kind < UBM_H_IDENTITY || h_kind > UBM_H_BOTH || !dimh) return UBM_ERR_INVALID;
  *dimh = (h_kind == UBM_H_BOTH) ? 2 * d : d;
  return UBM_OK;
}

int ubm_plan_create(ubm_handle* h, int32_t model_kind, const void* model, int32_t h_kind, const ubm_bins* bins,
                    int64_t k, int64_t m, int64_t lag, int64_t max_iterations, int64_t max_nrep, ubm_plan** out) {
  return plan_build(h, MODE_ESTIMATOR, model_kind, model, h_kind, bins, k, m, lag, max_iterations, max_nrep, 0, out);
}
int ubm_plan_launch(ubm_plan* p, uint64_t seed, int64_t nrep, int64_t rep_offset) {
  if (!p) return UBM_ERR_INVALID;
  return plan_run(p, UBM_DRAWS_PHILOX, seed, nrep, rep_offset);
}
int ubm_plan_fetch(ubm_plan* p, const ubm_estimator_out* out) {
  if (!p) return UBM_ERR_INVALID;
  return plan_fetch(p, out);
}
int ubm_plan_summary_dev(ubm_plan* p, void** dev_ptr, int64_t* len) {
  if (!p) return UBM_ERR_INVALID;
  if (dev_ptr) *dev_ptr = p->summary.p;
  if (len) *len = ubm_summary_len(p->dimh, p->nbins);
  return UBM_OK;
}
int ubm_plan_stream(ubm_plan* p, void** cuda_stream) {
  if (!p || !cuda_stream) return UBM_ERR_INVALID;
  *cuda_stream = (void*)p->h->stream;
  return UBM_OK;
}
int ubm_plan_last_ms(ubm_plan* p, float* ms_main_kernel, float* ms_total) {
  if (!p) return UBM_ERR_INVALID;
  ubm_handle* h = p->h;
  CU(cudaEventSynchronize(p->e2));
  if (ms_main_kernel) CU(cudaEventElapsedTime(ms_main_kernel, p->e0, p->e1));
  if (ms_total) CU(cudaEventElapsedTime(ms_total, p->e0, p->e2));
  return UBM_OK;
}
int ubm_plan_destroy(ubm_plan* p) {
  if (!p) return UBM_OK;
  cudaSetDevice(p->h->device);
  cudaStreamSynchronize(p->h->stream);
  delete p;
  return UBM_OK;
}

int ubm_sample_meetingtime(ubm_handle* h, int32_t model_kind, const void* model, const ubm_draws* draws, int64_t lag,
                           int64_t max_iterations, int64_t nrep, int64_t rep_offset, double* meetingtime) {
  if (!h) return UBM_ERR_INVALID;
  if (!draws || !meetingtime) return fail(h, UBM_ERR_INVALID, "null draws / output");
  ubm_plan* p = nullptr;
  int st = plan_build(h, MODE_MEETING, model_kind, model, UBM_H_IDENTITY, nullptr, 0, 1, lag, max_iterations, nrep, 0, &p);
  if (st != UBM_OK) return st;
  if (draws->mode == UBM_DRAWS_TAPE) st = plan_set_tapes(p, draws, nrep);
  if (st == UBM_OK) st = plan_run(p, draws->mode, draws->seed, nrep, rep_offset);
  if (st == UBM_OK) {
    ubm_estimator_out o{};
    o.meetingtime = meetingtime;
    st = plan_fetch(p, &o);
  }
  ubm_plan_destroy(p);
  return st;
}

int ubm_sample_unbiasedestimator(ubm_handle* h, int32_t model_kind, const void* model, const ubm_draws* draws,
                                 int32_t h_kind, const ubm_bins* bins, int64_t k, int64_t m, int64_t lag,
                                 int64_t max_iterations, int64_t nrep, int64_t rep_offset, const ubm_estimator_out* out) {
  if (!h) return UBM_ERR_INVALID;
  if (!draws || !out) return fail(h, UBM_ERR_INVALID, "null draws / out");
  ubm_plan* p = nullptr;
  int st = plan_build(h, MODE_ESTIMATOR, model_kind, model, h_kind, bins, k, m, lag, max_iterations, nrep, 0, &p);
  if (st != UBM_OK) return st;
  if (out->bins && p->nbins) {
    p->want_bins_out = true;
    if (p->bins_out.alloc((size_t)nrep * p->nbins * 8) != cudaSuccess) { ubm_plan_destroy(p); return fail(h, UBM_ERR_NOMEM, "bins_out alloc"); }
  }
  if (draws->mode == UBM_DRAWS_TAPE) st = plan_set_tapes(p, draws, nrep);
  if (st == UBM_OK) st = plan_run(p, draws->mode, draws->seed, nrep, rep_offset);
  if (st == UBM_OK) st = plan_fetch(p, out);
  ubm_plan_destroy(p);
  return st;
}

int ubm_sample_coupled_chains(ubm_handle* h, int32_t model_kind, const void* model, const ubm_draws* draws, int64_t m,
                              int64_t lag, int64_t max_iterations, int64_t stride, int64_t nrep, int64_t rep_offset,
                              double* samples1, double* samples2, double* meetingtime, int64_t* iteration, double* cost) {
  if (!h) return UBM_ERR_INVALID;
  if (!draws || !samples1 || !samples2) return fail(h, UBM_ERR_INVALID, "null draws / sample buffers");
  if (stride < lag + 2) return fail(h, UBM_ERR_INVALID, "stride must be at least lag + 2 rows");
  ubm_plan* p = nullptr;
  int st = plan_build(h, MODE_CHAINS, model_kind, model, UBM_H_IDENTITY, nullptr, 0, m, lag, max_iterations, nrep, stride, &p);
  if (st != UBM_OK) return st;
  if (draws->mode == UBM_DRAWS_TAPE) st = plan_set_tapes(p, draws, nrep);
  if (st == UBM_OK) {
    cudaError_t e1 = cudaMemsetAsync(p->samples1.p, 0xFF, p->samples1.bytes, h->stream);   // NaN fill: rows never written stay NaN
    cudaError_t e2 = cudaMemsetAsync(p->samples2.p, 0xFF, p->samples2.bytes, h->stream);
    if (e1 != cudaSuccess || e2 != cudaSuccess) st = fail(h, UBM_ERR_CUDA, "sample_coupled_chains: memset of the trajectory buffers failed");
    else st = plan_run(p, draws->mode, draws->seed, nrep, rep_offset);
  }
  if (st == UBM_OK) {
    ubm_estimator_out o{};
    o.meetingtime = meetingtime; o.iteration = iteration; o.cost = cost;
    st = plan_fetch(p, &o);
  }
  if (st == UBM_OK || st == UBM_ERR_TAPE) {
    if (cudaMemcpy(samples1, p->samples1.p, p->samples1.bytes, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(samples2, p->samples2.p, p->samples2.bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
      st = fail(h, UBM_ERR_CUDA, "sample_coupled_chains: copy of the trajectories failed");
  }
  ubm_plan_destroy(p);
  return st;
}

// argument checks shared by the estimators that read stored chains.  R's H_bar stops on a chain that did not meet
// ((k+lag):(Inf-1) is an error, R/H_bar.R:45) and estimator_bin_ converts meetingtime to int (src/estimator_bin.cpp:11):
// a censored replicate (meetingtime = +Inf) is rejected here instead of indexing past the stored rows.
static int check_stored_chains(ubm_handle* h, const char* who, int64_t k, int64_t m, int64_t lag, int64_t stride, int64_t nrep,
                               const double* meetingtime, const int64_t* iteration) {
  if (nrep < 1) return fail(h, UBM_ERR_INVALID, std::string(who) + ": nrep must be >= 1");
  if (lag < 1) return fail(h, UBM_ERR_INVALID, std::string(who) + ": lag must be >= 1");
  if (k < 0 || k > m) return fail(h, UBM_ERR_INVALID, std::string(who) + ": need 0 <= k <= m");
  if (m >= stride) return fail(h, UBM_ERR_INVALID, std::string(who) + ": m is beyond the stored rows (stride)");
  for (int64_t r = 0; r < nrep; ++r) {
    if (!(meetingtime[r] < 9.0e18)) return fail(h, UBM_ERR_INVALID, std::string(who) + ": a replicate did not meet (meetingtime = Inf); raise max_iterations");
    if (meetingtime[r] > (double)stride) return fail(h, UBM_ERR_INVALID, std::string(who) + ": a meeting time is beyond the stored rows (stride)");
    if (iteration) {   // R/H_bar.R:21-28
      if (k > iteration[r]) return fail(h, UBM_ERR_INVALID, "error: k has to be less than the horizon of the coupled chains");
      if (m > iteration[r]) return fail(h, UBM_ERR_INVALID, "error: m has to be less than the horizon of the coupled chains");
    }
  }
  return UBM_OK;
}

int ubm_h_bar(ubm_handle* h, int32_t dim, int32_t h_kind, int64_t k, int64_t m, int64_t lag, int64_t stride,
              int64_t nrep, const double* samples1, const double* samples2, const double* meetingtime,
              const int64_t* iteration, double* hbar) {
  if (!h) return UBM_ERR_INVALID;
  if (!samples1 || !samples2 || !meetingtime || !iteration || !hbar) return fail(h, UBM_ERR_INVALID, "null buffer");
  if (h_kind < UBM_H_IDENTITY || h_kind > UBM_H_BOTH) return fail(h, UBM_ERR_UNSUPPORTED, "unknown test function h");
  if (int stc = check_stored_chains(h, "H_bar", k, m, lag, stride, nrep, meetingtime, iteration)) return stc;
  CU(cudaSetDevice(h->device));
  const int dimh = (h_kind == UBM_H_BOTH) ? 2 * dim : dim;
  DevBuf s1, s2, mt, out;
  size_t sz = (size_t)nrep * stride * dim * 8;
  CU(s1.alloc(sz)); CU(s2.alloc(sz)); CU(mt.alloc((size_t)nrep * 8)); CU(out.alloc((size_t)nrep * dimh * 8));
  CU(cudaMemcpyAsync(s1.p, samples1, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(s2.p, samples2, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(mt.p, meetingtime, (size_t)nrep * 8, cudaMemcpyHostToDevice, h->stream));
  int64_t total = nrep * dimh;
  hbar_kernel<<<(unsigned)((total + 127) / 128), 128, 0, h->stream>>>(s1.as<double>(), s2.as<double>(), mt.as<double>(), dim, h_kind, k, m, lag, stride, nrep, out.as<double>());
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(hbar, out.p, (size_t)nrep * dimh * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return UBM_OK;
}

int ubm_h_bar_grid(ubm_handle* h, int32_t dim, int32_t h_kind, int32_t npairs, const int64_t* ks, const int64_t* ms,
                   int64_t lag, int64_t stride, int64_t nrep, const double* samples1, const double* samples2,
                   const double* meetingtime, const int64_t* iteration, double* hbar) {
  if (!h) return UBM_ERR_INVALID;
  if (!samples1 || !samples2 || !meetingtime || !iteration || !hbar || !ks || !ms || npairs < 1) return fail(h, UBM_ERR_INVALID, "null buffer");
  if (h_kind < UBM_H_IDENTITY || h_kind > UBM_H_BOTH) return fail(h, UBM_ERR_UNSUPPORTED, "unknown test function h");
  for (int32_t q = 0; q < npairs; ++q)
    if (int stc = check_stored_chains(h, "H_bar", ks[q], ms[q], lag, stride, nrep, meetingtime, iteration)) return stc;
  CU(cudaSetDevice(h->device));
  const int dimh = (h_kind == UBM_H_BOTH) ? 2 * dim : dim;
  DevBuf s1, s2, mt, out;
  const size_t sz = (size_t)nrep * stride * dim * 8, ob = (size_t)nrep * dimh * 8;
  CU(s1.alloc(sz)); CU(s2.alloc(sz)); CU(mt.alloc((size_t)nrep * 8)); CU(out.alloc(ob * npairs));
  CU(cudaMemcpyAsync(s1.p, samples1, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(s2.p, samples2, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(mt.p, meetingtime, (size_t)nrep * 8, cudaMemcpyHostToDevice, h->stream));
  const int64_t total = nrep * dimh;
  for (int32_t q = 0; q < npairs; ++q) {     // chains stay resident; one launch per pair on the same stream
    hbar_kernel<<<(unsigned)((total + 127) / 128), 128, 0, h->stream>>>(s1.as<double>(), s2.as<double>(), mt.as<double>(), dim, h_kind, ks[q], ms[q], lag, stride, nrep, out.as<double>() + (size_t)q * nrep * dimh);
    h->launches += 1;
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(hbar, out.p, ob * npairs, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return UBM_OK;
}

int ubm_tv_upper_bounds(ubm_handle* h, const double* meetingtime, int64_t nrep, int64_t lag, const int64_t* t, int32_t nt,
                        double* bound) {
  if (!h) return UBM_ERR_INVALID;
  if (!meetingtime || !t || !bound || nrep < 1 || nt < 1 || lag < 1) return fail(h, UBM_ERR_INVALID, "tv_upper_bounds: invalid arguments");
  for (int64_t r = 0; r < nrep; ++r)
    if (!(meetingtime[r] < 9.0e18)) return fail(h, UBM_ERR_INVALID, "tv_upper_bounds: a meeting time is Inf (raise max_iterations)");
  CU(cudaSetDevice(h->device));
  DevBuf mt, dt, out;
  CU(mt.alloc((size_t)nrep * 8)); CU(dt.alloc((size_t)nt * 8)); CU(out.alloc((size_t)nt * 8));
  CU(cudaMemcpyAsync(mt.p, meetingtime, (size_t)nrep * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(dt.p, t, (size_t)nt * 8, cudaMemcpyHostToDevice, h->stream));
  tv_bound_kernel<<<(unsigned)nt, 256, 0, h->stream>>>(mt.as<double>(), nrep, lag, reinterpret_cast<const int64_t*>(dt.p), out.as<double>());
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(bound, out.p, (size_t)nt * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return UBM_OK;
}

int64_t ubm_measure_size(int64_t k, int64_t m, int64_t lag, double meetingtime) {
  if (!(meetingtime < 9.0e18)) return -1;                       // censored
  const int64_t T = (int64_t)meetingtime;
  return (m - k + 1) + 2 * ((T - (k + lag) > 0) ? T - (k + lag) : 0);
}

int ubm_c_chains_to_measure(ubm_handle* h, int32_t dim, int64_t k, int64_t m, int64_t lag, int64_t stride, int64_t nrep,
                            const double* samples1, const double* samples2, const double* meetingtime,
                            const int64_t* iteration, int64_t cap, double* atoms, double* weights, int32_t* mcmc_part,
                            int64_t* sizes) {
  if (!h) return UBM_ERR_INVALID;
  if (!samples1 || !samples2 || !meetingtime || !iteration || !atoms || !weights || !mcmc_part || !sizes) return fail(h, UBM_ERR_INVALID, "null buffer");
  if (k > m) return fail(h, UBM_ERR_INVALID, "k must be <= than m");                     // R/c_chains_to_measure_as_list.R:20-22
  if (dim < 1 || nrep < 1 || cap < 1 || lag < 1 || k < 0) return fail(h, UBM_ERR_INVALID, "invalid sizes");
  for (int64_t r = 0; r < nrep; ++r) {
    if (m > iteration[r]) return fail(h, UBM_ERR_INVALID, "m has to be less than the time horizon of the coupled chains");   // :23-25
    const int64_t sz = ubm_measure_size(k, m, lag, meetingtime[r]);
    if (sz < 0) return fail(h, UBM_ERR_INVALID, "c_chains_to_measure: a replicate did not meet (meetingtime = Inf)");
    if (sz > cap) return fail(h, UBM_ERR_INVALID, "c_chains_to_measure: cap is smaller than the number of atoms of a replicate");
    sizes[r] = sz;
  }
  CU(cudaSetDevice(h->device));
  DevBuf s1, s2, mt, da, dw, dp;
  const size_t sz = (size_t)nrep * stride * dim * 8, rows = (size_t)nrep * cap;
  CU(s1.alloc(sz)); CU(s2.alloc(sz)); CU(mt.alloc((size_t)nrep * 8));
  CU(da.alloc(rows * dim * 8)); CU(dw.alloc(rows * 8)); CU(dp.alloc(rows * 4));
  CU(cudaMemcpyAsync(s1.p, samples1, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(s2.p, samples2, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(mt.p, meetingtime, (size_t)nrep * 8, cudaMemcpyHostToDevice, h->stream));
  const int64_t total = (int64_t)rows * dim;
  measure_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(s1.as<double>(), s2.as<double>(), mt.as<double>(), dim, k, m, lag, stride, nrep, cap, da.as<double>(), dw.as<double>(), dp.as<int>());
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(atoms, da.p, rows * dim * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(weights, dw.p, rows * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(mcmc_part, dp.p, rows * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return UBM_OK;
}

int ubm_estimator_bins(ubm_handle* h, int32_t dim, const ubm_bins* bins, int64_t k, int64_t m, int64_t lag,
                       int64_t stride, int64_t nrep, const double* samples1, const double* samples2,
                       const double* meetingtime, double* binvals) {
  if (!h) return UBM_ERR_INVALID;
  if (!bins || bins->nbreaks < 2 || !samples1 || !samples2 || !meetingtime || !binvals) return fail(h, UBM_ERR_INVALID, "null buffer / bins");
  if (bins->component < 0 || bins->component >= dim) return fail(h, UBM_ERR_INVALID, "bins: component out of range");
  if (int stc = check_stored_chains(h, "estimator_bin", k, m, lag, stride, nrep, meetingtime, nullptr)) return stc;
  CU(cudaSetDevice(h->device));
  const int nbins = bins->nbreaks - 1;
  DevBuf s1, s2, mt, br, out;
  size_t sz = (size_t)nrep * stride * dim * 8;
  CU(s1.alloc(sz)); CU(s2.alloc(sz)); CU(mt.alloc((size_t)nrep * 8)); CU(br.alloc((size_t)bins->nbreaks * 8));
  CU(out.alloc((size_t)nrep * nbins * 8));
  CU(cudaMemcpyAsync(s1.p, samples1, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(s2.p, samples2, sz, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(mt.p, meetingtime, (size_t)nrep * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(br.p, bins->breaks, (size_t)bins->nbreaks * 8, cudaMemcpyHostToDevice, h->stream));
  int64_t total = nrep * nbins;
  estimator_bin_kernel<<<(unsigned)((total + 127) / 128), 128, 0, h->stream>>>(s1.as<double>(), s2.as<double>(), mt.as<double>(), dim, bins->component, br.as<double>(), nbins, k, m, lag, stride, nrep, out.as<double>());
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(binvals, out.p, (size_t)nrep * nbins * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return UBM_OK;
}

// ---------------------------------------------------------------------------------------------------
// stand-alone primitives, batched over samples
// ---------------------------------------------------------------------------------------------------
static int prim_params(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, RunParams* P, DevBuf* tu,
                       DevBuf* tn, DevBuf* status) {
  if (!draws) return fail(h, UBM_ERR_INVALID, "null draws");
  if (n < 1) return fail(h, UBM_ERR_INVALID, "n must be >= 1");
  CU(cudaSetDevice(h->device));
  *P = RunParams{};
  P->nrep = n; P->rep_offset = rep_offset; P->draws_mode = draws->mode; P->seed = draws->seed;
  if (draws->mode == UBM_DRAWS_TAPE) {
    if (draws->tape_u && draws->len_u > 0) {
      CU(tu->alloc((size_t)n * draws->len_u * 8));
      CU(cudaMemcpyAsync(tu->p, draws->tape_u, (size_t)n * draws->len_u * 8, cudaMemcpyHostToDevice, h->stream));
      P->tape_u = tu->as<double>(); P->len_u = draws->len_u;
    }
    if (draws->tape_n && draws->len_n > 0) {
      CU(tn->alloc((size_t)n * draws->len_n * 8));
      CU(cudaMemcpyAsync(tn->p, draws->tape_n, (size_t)n * draws->len_n * 8, cudaMemcpyHostToDevice, h->stream));
      P->tape_n = tn->as<double>(); P->len_n = draws->len_n;
    }
  }
  CU(status->alloc(4));
  CU(cudaMemsetAsync(status->p, 0, 4, h->stream));
  P->status = status->as<int>();
  return UBM_OK;
}
static int prim_finish(ubm_handle* h, DevBuf* status) {
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  int st = 0;
  CU(cudaMemcpy(&st, status->p, 4, cudaMemcpyDeviceToHost));
  if (st == UBM_ERR_TAPE) return fail(h, UBM_ERR_TAPE, "replay tape exhausted");
  return UBM_OK;
}

static int scalar_coupling(ubm_handle* h, const ubm_draws* draws, int mode, int64_t n, int64_t rep_offset, double mu1,
                           double mu2, double s1, double s2, double* xy, int32_t* identical) {
  if (!h) return UBM_ERR_INVALID;
  if (!xy || !identical) return fail(h, UBM_ERR_INVALID, "null output");
  if (!(s1 > 0) || !(s2 > 0)) return fail(h, UBM_ERR_INVALID, "standard deviations must be > 0");
  RunParams P; DevBuf tu, tn, status, dxy, did;
  int st = prim_params(h, draws, n, rep_offset, &P, &tu, &tn, &status);
  if (st != UBM_OK) return st;
  CU(dxy.alloc((size_t)n * 16)); CU(did.alloc((size_t)n * 4));
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (draws->mode == UBM_DRAWS_TAPE)
    prim_scalar_coupling_kernel<true><<<grid, 128, 0, h->stream>>>(P, mode, mu1, mu2, s1, s2, ubm_log(s1), ubm_log(s2), dxy.as<double>(), did.as<int>());
  else
    prim_scalar_coupling_kernel<false><<<grid, 128, 0, h->stream>>>(P, mode, mu1, mu2, s1, s2, ubm_log(s1), ubm_log(s2), dxy.as<double>(), did.as<int>());
  h->launches += 1;
  st = prim_finish(h, &status);
  CU(cudaMemcpy(xy, dxy.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(identical, did.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return st;
}

int ubm_rnorm_max_coupling(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double mu1, double mu2,
                           double sigma1, double sigma2, double* xy, int32_t* identical) {
  return scalar_coupling(h, draws, 1, n, rep_offset, mu1, mu2, sigma1, sigma2, xy, identical);
}
int ubm_rnorm_reflectionmax(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double mu1, double mu2,
                            double sigma, double* xy, int32_t* identical) {
  return scalar_coupling(h, draws, 0, n, rep_offset, mu1, mu2, sigma, sigma, xy, identical);
}

static int positive_coupling(ubm_handle* h, const ubm_draws* draws, int kind, int mode, int64_t n, int64_t rep_offset, double a1,
                             double a2, double b1, double b2, double* out, int32_t* identical) {
  if (!h) return UBM_ERR_INVALID;
  if (!out || (mode == 1 && !identical)) return fail(h, UBM_ERR_INVALID, "null output");
  if (!(a1 > 0) || !(b1 > 0) || (mode == 1 && (!(a2 > 0) || !(b2 > 0)))) return fail(h, UBM_ERR_INVALID, "parameters must be > 0");
  RunParams P; DevBuf tu, tn, status, dout, did;
  int st = prim_params(h, draws, n, rep_offset, &P, &tu, &tn, &status);
  if (st != UBM_OK) return st;
  const double c1 = (kind == 2) ? 0.0 : a1 * ubm_log(b1) - std::lgamma(a1);
  const double c2 = (kind == 2 || mode == 0) ? 0.0 : a2 * ubm_log(b2) - std::lgamma(a2);
  const size_t ob = (size_t)n * 8 * (mode == 1 ? 2 : 1);
  CU(dout.alloc(ob)); CU(did.alloc((size_t)n * 4));
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (draws->mode == UBM_DRAWS_TAPE)
    prim_positive_coupling_kernel<true><<<grid, 128, 0, h->stream>>>(P, kind, mode, a1, a2, b1, b2, c1, c2, dout.as<double>(), did.as<int>());
  else
    prim_positive_coupling_kernel<false><<<grid, 128, 0, h->stream>>>(P, kind, mode, a1, a2, b1, b2, c1, c2, dout.as<double>(), did.as<int>());
  h->launches += 1;
  st = prim_finish(h, &status);
  CU(cudaMemcpy(out, dout.p, ob, cudaMemcpyDeviceToHost));
  if (mode == 1) CU(cudaMemcpy(identical, did.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return st;
}
int ubm_rgamma_coupled(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double alpha1, double alpha2,
                       double beta1, double beta2, double* xy, int32_t* identical) {
  return positive_coupling(h, draws, 0, 1, n, rep_offset, alpha1, alpha2, beta1, beta2, xy, identical);
}
int ubm_rinversegamma_coupled(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double alpha1, double alpha2,
                              double beta1, double beta2, double* xy, int32_t* identical) {
  return positive_coupling(h, draws, 1, 1, n, rep_offset, alpha1, alpha2, beta1, beta2, xy, identical);
}
int ubm_rinvgaussian_coupled(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, double mu1, double mu2,
                             double lambda1, double lambda2, double* xy, int32_t* identical) {
  return positive_coupling(h, draws, 2, 1, n, rep_offset, mu1, mu2, lambda1, lambda2, xy, identical);
}
int ubm_rpositive(ubm_handle* h, const ubm_draws* draws, int32_t kind, int64_t n, int64_t rep_offset, double a, double b, double* out) {
  if (kind < 0 || kind > 2) return h ? fail(h, UBM_ERR_INVALID, "kind: 0 Gamma, 1 inverse Gamma, 2 inverse Gaussian") : UBM_ERR_INVALID;
  return positive_coupling(h, draws, kind, 0, n, rep_offset, a, a, b, b, out, nullptr);
}

static int mvn_sample(ubm_handle* h, const ubm_draws* draws, int mode, int64_t n, int64_t rep_offset, int32_t d,
                      const double* mu1, const double* mu2, const double* chol, const double* chol_inv, double* out,
                      int32_t* identical) {
  if (!h) return UBM_ERR_INVALID;
  if (!mu1 || !chol || !out || (mode == 1 && (!mu2 || !chol_inv || !identical))) return fail(h, UBM_ERR_INVALID, "null argument");
  if (mode == 1 && d < 2)   // R/mvnorm_couplings.R:96-98
    return fail(h, UBM_ERR_INVALID, "function 'rmvnorm_reflectionmax' is meant for multivariate Normals; for univariate Normals, use 'rnorm_reflectionmax'");
  if (d < 1 || d > PRIM_MAXD) return fail(h, UBM_ERR_INVALID, "1 <= d <= 128");
  RunParams P; DevBuf tu, tn, status, dm1, dm2, dc, dci, dout, did;
  int st = prim_params(h, draws, n, rep_offset, &P, &tu, &tn, &status);
  if (st != UBM_OK) return st;
  const size_t vb = (size_t)d * 8, mb = (size_t)d * d * 8, ob = (size_t)n * d * 8 * (mode == 1 ? 2 : 1);
  CU(dm1.alloc(vb)); CU(dc.alloc(mb)); CU(dout.alloc(ob)); CU(did.alloc((size_t)n * 4));
  CU(cudaMemcpyAsync(dm1.p, mu1, vb, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(dc.p, chol, mb, cudaMemcpyHostToDevice, h->stream));
  if (mode == 1) {
    CU(dm2.alloc(vb)); CU(dci.alloc(mb));
    CU(cudaMemcpyAsync(dm2.p, mu2, vb, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dci.p, chol_inv, mb, cudaMemcpyHostToDevice, h->stream));
  }
  const unsigned grid = (unsigned)((n + 63) / 64);
  if (draws->mode == UBM_DRAWS_TAPE)
    prim_mvn_sample_kernel<true><<<grid, 64, 0, h->stream>>>(P, mode, d, dm1.as<double>(), dm2.as<double>(), dc.as<double>(), dci.as<double>(), dout.as<double>(), did.as<int>());
  else
    prim_mvn_sample_kernel<false><<<grid, 64, 0, h->stream>>>(P, mode, d, dm1.as<double>(), dm2.as<double>(), dc.as<double>(), dci.as<double>(), dout.as<double>(), did.as<int>());
  h->launches += 1;
  st = prim_finish(h, &status);
  CU(cudaMemcpy(out, dout.p, ob, cudaMemcpyDeviceToHost));
  if (mode == 1) CU(cudaMemcpy(identical, did.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return st;
}

static int mvn_max(ubm_handle* h, const ubm_draws* draws, int variant, int64_t n, int64_t rep_offset, int32_t d,
                   const double* mu1, const double* mu2, const double* A1, const double* A2, const double* B1,
                   const double* B2, double* xy, int32_t* identical, int32_t* nattempts) {
  if (!h) return UBM_ERR_INVALID;
  if (!mu1 || !mu2 || !A1 || !A2 || !xy || !identical || (variant == 0 && (!B1 || !B2))) return fail(h, UBM_ERR_INVALID, "null argument");
  if (d < 1 || d > PRIM_MAXD) return fail(h, UBM_ERR_INVALID, "1 <= d <= 128");
  const size_t vb = (size_t)d * 8, mb = (size_t)d * d * 8;
  std::vector<double> L1, L2;
  double cst1, cst2;
  if (variant == 1) {       // LLT of the covariances (src/mvnorm.cpp:12, :52), once for the whole batch
    L1.assign(A1, A1 + (size_t)d * d); L2.assign(A2, A2 + (size_t)d * d);
    pg_chol_lower_host(L1, d); pg_chol_lower_host(L2, d);
    for (int j = 0; j < d; ++j)
      if (!(L1[(size_t)j * d + j] > 0.0) || !(L2[(size_t)j * d + j] > 0.0)) return fail(h, UBM_ERR_INVALID, "rmvnorm_max: covariance not positive definite");
    double h1 = 0.0, h2 = 0.0;
    for (int j = 0; j < d; ++j) { h1 = h1 + ubm_log(L1[(size_t)j * d + j]); h2 = h2 + ubm_log(L2[(size_t)j * d + j]); }
    cst1 = -(h1) - (d * 0.9189385); cst2 = -(h2) - (d * 0.9189385);            // :55-56
    A1 = L1.data(); A2 = L2.data();
  } else {
    double h1 = 0.0, h2 = 0.0;
    for (int j = 0; j < d; ++j) { h1 = h1 + ubm_log(B1[(size_t)j * d + j]); h2 = h2 + ubm_log(B2[(size_t)j * d + j]); }
    cst1 = -(-h1) - (d * 0.9189385); cst2 = -(-h2) - (d * 0.9189385);          // :74-75
  }
  RunParams P; DevBuf tu, tn, status, dm1, dm2, da1, da2, db1, db2, dout, did, dna;
  int st = prim_params(h, draws, n, rep_offset, &P, &tu, &tn, &status);
  if (st != UBM_OK) return st;
  CU(dm1.alloc(vb)); CU(dm2.alloc(vb)); CU(da1.alloc(mb)); CU(da2.alloc(mb)); CU(db1.alloc(mb)); CU(db2.alloc(mb));
  CU(dout.alloc((size_t)n * 2 * d * 8)); CU(did.alloc((size_t)n * 4)); CU(dna.alloc((size_t)n * 4));
  CU(cudaMemcpyAsync(dm1.p, mu1, vb, cudaMemcpyHostToDevice, h->stream)); CU(cudaMemcpyAsync(dm2.p, mu2, vb, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(da1.p, A1, mb, cudaMemcpyHostToDevice, h->stream)); CU(cudaMemcpyAsync(da2.p, A2, mb, cudaMemcpyHostToDevice, h->stream));
  if (variant == 0) { CU(cudaMemcpyAsync(db1.p, B1, mb, cudaMemcpyHostToDevice, h->stream)); CU(cudaMemcpyAsync(db2.p, B2, mb, cudaMemcpyHostToDevice, h->stream)); }
  const unsigned grid = (unsigned)((n + 63) / 64);
  if (draws->mode == UBM_DRAWS_TAPE)
    prim_mvn_max_kernel<true><<<grid, 64, 0, h->stream>>>(P, variant, d, dm1.as<double>(), dm2.as<double>(), da1.as<double>(), da2.as<double>(), db1.as<double>(), db2.as<double>(), cst1, cst2, dout.as<double>(), did.as<int>(), dna.as<int>());
  else
    prim_mvn_max_kernel<false><<<grid, 64, 0, h->stream>>>(P, variant, d, dm1.as<double>(), dm2.as<double>(), da1.as<double>(), da2.as<double>(), db1.as<double>(), db2.as<double>(), cst1, cst2, dout.as<double>(), did.as<int>(), dna.as<int>());
  h->launches += 1;
  st = prim_finish(h, &status);
  CU(cudaMemcpy(xy, dout.p, (size_t)n * 2 * d * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(identical, did.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (nattempts) CU(cudaMemcpy(nattempts, dna.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return st;
}
int ubm_rmvnorm_max(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d, const double* mu1,
                    const double* mu2, const double* Sigma1, const double* Sigma2, double* xy, int32_t* identical,
                    int32_t* nattempts) {
  return mvn_max(h, draws, 1, n, rep_offset, d, mu1, mu2, Sigma1, Sigma2, nullptr, nullptr, xy, identical, nattempts);
}
int ubm_rmvnorm_max_chol(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d,
                         const double* mu1, const double* mu2, const double* chol1, const double* chol2,
                         const double* chol_inv1, const double* chol_inv2, double* xy, int32_t* identical,
                         int32_t* nattempts) {
  return mvn_max(h, draws, 0, n, rep_offset, d, mu1, mu2, chol1, chol2, chol_inv1, chol_inv2, xy, identical, nattempts);
}

int ubm_fast_rmvnorm_chol(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d,
                          const double* mean, const double* chol, double* out) {
  return mvn_sample(h, draws, 0, n, rep_offset, d, mean, nullptr, chol, nullptr, out, nullptr);
}
extern "C" int ubm_fast_dmvnorm_chol_inverse(ubm_handle* h, int64_t n, int32_t d, const double* x, const double* mean,
                                             const double* chol_inv, double* out);
// covariance forms (src/mvnorm.cpp:9-26 fast_rmvnorm_, :49-67 fast_dmvnorm_): the LLT of the covariance is taken on the host
// (the canonical left-looking factorisation of this engine), then the factor forms run
static int host_factor(ubm_handle* h, int32_t d, const double* covariance, std::vector<double>* L) {
  if (!covariance) return fail(h, UBM_ERR_INVALID, "null covariance");
  if (d < 1 || d > PRIM_MAXD) return fail(h, UBM_ERR_INVALID, "1 <= d <= 128");
  L->assign(covariance, covariance + (size_t)d * d);
  pg_chol_lower_host(*L, d);
  for (int j = 0; j < d; ++j) if (!((*L)[(size_t)j * d + j] > 0.0)) return fail(h, UBM_ERR_INVALID, "covariance is not positive definite");
  return UBM_OK;
}
int ubm_fast_rmvnorm(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d, const double* mean,
                     const double* covariance, double* out) {
  if (!h) return UBM_ERR_INVALID;
  std::vector<double> L;
  int st = host_factor(h, d, covariance, &L);
  if (st != UBM_OK) return st;
  std::vector<double> U((size_t)d * d, 0.0);                    // upper factor, column-major: U(i, j) = L(j, i)
  for (int j = 0; j < d; ++j) for (int i = 0; i <= j; ++i) U[(size_t)j * d + i] = L[(size_t)i * d + j];
  return mvn_sample(h, draws, 0, n, rep_offset, d, mean, nullptr, U.data(), nullptr, out, nullptr);
}
int ubm_fast_dmvnorm(ubm_handle* h, int64_t n, int32_t d, const double* x, const double* mean, const double* covariance,
                     double* out) {
  if (!h) return UBM_ERR_INVALID;
  std::vector<double> L;
  int st = host_factor(h, d, covariance, &L);
  if (st != UBM_OK) return st;
  std::vector<double> Li((size_t)d * d, 0.0), Ui((size_t)d * d, 0.0);
  for (int j = 0; j < d; ++j) {                                 // inverse of the lower factor, reciprocal diagonal (as pg_pack_host)
    Li[(size_t)j * d + j] = 1.0 / L[(size_t)j * d + j];
    for (int i = j + 1; i < d; ++i) {
      double acc = 0.0;
      for (int k = j; k < i; ++k) acc = ubm_fma(L[(size_t)k * d + i], Li[(size_t)j * d + k], acc);
      Li[(size_t)j * d + i] = -acc * (1.0 / L[(size_t)i * d + i]);
    }
  }
  for (int j = 0; j < d; ++j) for (int i = 0; i <= j; ++i) Ui[(size_t)j * d + i] = Li[(size_t)i * d + j];   // inverse of the upper factor
  return ubm_fast_dmvnorm_chol_inverse(h, n, d, x, mean, Ui.data(), out);
}
int ubm_rmvnorm_reflectionmax(ubm_handle* h, const ubm_draws* draws, int64_t n, int64_t rep_offset, int32_t d,
                              const double* mu1, const double* mu2, const double* chol, const double* chol_inv,
                              double* xy, int32_t* identical) {
  return mvn_sample(h, draws, 1, n, rep_offset, d, mu1, mu2, chol, chol_inv, xy, identical);
}
int ubm_fast_dmvnorm_chol_inverse(ubm_handle* h, int64_t n, int32_t d, const double* x, const double* mean,
                                  const double* chol_inv, double* out) {
  if (!h) return UBM_ERR_INVALID;
  if (!x || !mean || !chol_inv || !out || n < 1) return fail(h, UBM_ERR_INVALID, "null argument");
  if (d < 1 || d > PRIM_MAXD) return fail(h, UBM_ERR_INVALID, "1 <= d <= 128");
  CU(cudaSetDevice(h->device));
  double halflogdet = 0.0;
  for (int j = 0; j < d; ++j) halflogdet = halflogdet + ubm_log(chol_inv[(size_t)j * d + j]);
  const double cst = -(-halflogdet) - (d * 0.9189385);      // src/mvnorm.cpp:74-75
  DevBuf dx, dm, dci, dout;
  CU(dx.alloc((size_t)n * d * 8)); CU(dm.alloc((size_t)d * 8)); CU(dci.alloc((size_t)d * d * 8)); CU(dout.alloc((size_t)n * 8));
  CU(cudaMemcpyAsync(dx.p, x, (size_t)n * d * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(dm.p, mean, (size_t)d * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(dci.p, chol_inv, (size_t)d * d * 8, cudaMemcpyHostToDevice, h->stream));
  prim_dmvnorm_kernel<<<(unsigned)((n + 63) / 64), 64, 0, h->stream>>>(n, d, dx.as<double>(), dm.as<double>(), dci.as<double>(), cst, dout.as<double>());
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return UBM_OK;
}

}  // extern "C"
