// ubm_pg.cuh — C3: Polya-Gamma Gibbs sampler for Bayesian logistic regression, coupled.
//
// Replaces, fused into one persistent kernel per launch:
//   drivers   R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//   kernels   inst/reproducegermancredit/germancredit.run.R:23-49 (single, coupled, rinit)
//   numerics  src/PolyaGamma.cpp:65-226 (Devroye PG(1,z) sampler), src/logisticregressioncoupling.cpp:9-115
//             (logcosh, xbeta_, w_max_couplingC), src/logisticregression.cpp:32-56 (m_sigma_function_),
//             R/logistic_regression.R:180-209 (sample_beta), src/mvnorm.cpp:9-26, :49-67, :91-129 (rmvnorm_max)
//
// Design (B200): one CTA (512 threads) per replicate.
//   1. PG sweep: one thread per observation computes x_i'beta for both chains (X column-major in L2, coalesced),
//      then runs the Devroye sampler and the per-observation max coupling with divergent rejection loops.  Draws are
//      addressed by CELL (sweep, observation) in the replicate's typed Philox streams, so the sweep is parallel and
//      independent of scheduling (oracle/oracle_pg.hpp states the same contract).
//   2. Gram build A_c = X' diag(w_c) X + B^-1 for both chains at once, register-tiled: a thread owns a 4 x 4 block of
//      entries for one of four interleaved observation classes (i mod 4), X is streamed through double-buffered 32-row
//      tiles of a padded row-major copy; canonical order: four chains over i mod 4, added to B^-1 in turn
//      (the reference: one chain over i, src/logisticregression.cpp:40-48).
//   3. p x p linear algebra: Cholesky (right-looking, same per-entry fma sequence as the oracle) and Sigma = Linv' Linv by
//      the whole CTA; solves and the triangular inverse at warp level (warp 0 chain 1, warp 1 chain 2).
//   4. max-coupled multivariate Normal draw of (beta1, beta2) by warp 0.
// Everything lives in shared memory; HBM sees X (392 KB, L2-resident) and the per-replicate results.
//
// Replay (UBM_DRAWS_TAPE): the reference's own order of consumption — ONE sequential stream per replicate, observation
// after observation through w_max_couplingC / w_rejsamplerC with their rejection loops (src/logisticregressioncoupling.cpp:
// 42-113), then the normals and uniforms of the beta update.  How many draws an observation takes is data dependent, so
// thread 0 walks the observations (struct PgSeq); this is a parity path like sample_coupled_chains, not a fast one.
#pragma once
#include <string>
#include <vector>
#include "ubm_common.cuh"
#include "../../include/ubm_pg.h"

namespace ubm {

#define PG_T 512
#define PG_TILE 32
#define PG_MAXP 64
#define PG_GS (PG_MAXP + 4)      // row stride of the staged Gram tile (doubles), 16-byte aligned rows
#define PG_PI_D 3.141592653589793238462643383279502884197
#define PG_TRUNC_D 0.64

struct PgDevice {
  int n, p, ne, w_coupling, beta_coupling;
  const double *X, *invB, *ktk, *b, *cholB;   // device: X n x p col-major; invB p x p; ktk, b [p]; cholB upper p x p
  const double *Xt;                            // device: X again, row-major [ceil(n/32)*32][PG_GS], zero padded: the Gram tiles
  size_t offs[6];
};

// host-side precomputation (logisticregression_precomputation, R/logistic_regression.R:34-50, and chol(B) for rinit)
inline void pg_chol_lower_host(std::vector<double>& A, int p) {
  for (int j = 0; j < p; ++j) {
    double s = A[(size_t)j * p + j];
    for (int k = 0; k < j; ++k) s = ubm_fma(-A[(size_t)k * p + j], A[(size_t)k * p + j], s);
    double ljj = sqrt(s);
    A[(size_t)j * p + j] = ljj;
    for (int i = j + 1; i < p; ++i) {
      double t = A[(size_t)j * p + i];
      for (int k = 0; k < j; ++k) t = ubm_fma(-A[(size_t)k * p + i], A[(size_t)k * p + j], t);
      A[(size_t)j * p + i] = t / ljj;
    }
  }
  for (int j = 0; j < p; ++j) for (int i = 0; i < j; ++i) A[(size_t)j * p + i] = 0.0;
}
inline void pg_pack_host(const ubm_model_pg_logistic* s, PgDevice* M, std::vector<double>* blob) {
  const int n = s->n, p = s->p;
  M->n = n; M->p = p; M->ne = p * (p + 1) / 2; M->w_coupling = s->w_coupling; M->beta_coupling = s->beta_coupling;
  std::vector<double> LB(s->B, s->B + (size_t)p * p);
  pg_chol_lower_host(LB, p);
  std::vector<double> cholB((size_t)p * p, 0.0), Li((size_t)p * p, 0.0), invB((size_t)p * p, 0.0), ktk(p, 0.0);
  for (int j = 0; j < p; ++j) for (int i = 0; i <= j; ++i) cholB[(size_t)j * p + i] = LB[(size_t)i * p + j];
  for (int j = 0; j < p; ++j) {
    Li[(size_t)j * p + j] = 1.0 / LB[(size_t)j * p + j];
    for (int i = j + 1; i < p; ++i) {
      double acc = 0.0;
      for (int k = j; k < i; ++k) acc = ubm_fma(LB[(size_t)k * p + i], Li[(size_t)j * p + k], acc);
      Li[(size_t)j * p + i] = -acc * (1.0 / LB[(size_t)i * p + i]);
    }
  }
  for (int a = 0; a < p; ++a)
    for (int c = a; c < p; ++c) {
      double acc = 0.0;
      for (int i = c; i < p; ++i) acc = ubm_fma(Li[(size_t)a * p + i], Li[(size_t)c * p + i], acc);
      invB[(size_t)c * p + a] = acc; invB[(size_t)a * p + c] = acc;
    }
  for (int j = 0; j < p; ++j) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = ubm_fma(s->X[(size_t)j * n + i], s->Y[i] - 0.5, acc);
    for (int k = 0; k < p; ++k) acc = ubm_fma(invB[(size_t)k * p + j], s->b[k], acc);
    ktk[j] = acc;
  }
  blob->clear();
  auto put = [&](const double* src, size_t len) { size_t o = blob->size(); blob->insert(blob->end(), src, src + len); return o; };
  M->offs[0] = put(s->X, (size_t)n * p);
  M->offs[1] = put(invB.data(), invB.size());
  M->offs[2] = put(ktk.data(), ktk.size());
  M->offs[3] = put(s->b, p);
  M->offs[4] = put(cholB.data(), cholB.size());
  const int npad = (n + PG_TILE - 1) / PG_TILE * PG_TILE;
  std::vector<double> Xt((size_t)npad * PG_GS, 0.0);
  for (int i = 0; i < n; ++i) for (int j = 0; j < p; ++j) Xt[(size_t)i * PG_GS + j] = s->X[(size_t)j * n + i];
  if (blob->size() & 1) blob->push_back(0.0);      // 16-byte alignment of the tiles
  M->offs[5] = put(Xt.data(), Xt.size());
}
inline void pg_bind_device(PgDevice* M, const double* dev) {
  M->X = dev + M->offs[0]; M->invB = dev + M->offs[1]; M->ktk = dev + M->offs[2]; M->b = dev + M->offs[3];
  M->cholB = dev + M->offs[4]; M->Xt = dev + M->offs[5];
}

// ---- out-of-line helpers -----------------------------------------------------------------------------------------
// The Devroye sampler calls log / exp / log-pnorm and the typed draws at some thirty places; inlined, the kernel grew
// to 27 000 instructions (430 KB) and ncu showed `no_inst` (instruction-fetch) stalls on 50-80 % of the samples of the
// sweep.  One copy of each keeps the hot loops inside the instruction cache; the arithmetic is unchanged.
__device__ __noinline__ double pg_log(double x) { return ubm_log(x); }
__device__ __noinline__ double pg_exp(double x) { return ubm_exp(x); }
__device__ __noinline__ double pg_pnorm_log(double x) { return ubm_pnorm_log(x); }
// draw k of a typed stream of (replicate, cell, attempt): one Philox block per call, no cached state
__device__ __noinline__ double pg_cell_u(uint64_t seed, uint64_t rep_eff, uint64_t base, unsigned k) {
  const ubm_u32x4 b = ubm_stream_block(seed, rep_eff, UBM_STREAM_U, base | (k >> 2));
  const unsigned s = k & 3u;
  return ubm_u01_32((s == 0) ? b.w[0] : (s == 1) ? b.w[1] : (s == 2) ? b.w[2] : b.w[3]);
}
__device__ __noinline__ double pg_cell_u53(uint64_t seed, uint64_t rep_eff, uint64_t base, unsigned k, int stream) {
  const ubm_u32x4 b = ubm_stream_block(seed, rep_eff, stream, base | (k >> 1));
  return (k & 1) ? ubm_u01_53(b.w[2], b.w[3]) : ubm_u01_53(b.w[0], b.w[1]);
}
__device__ __noinline__ double pg_qnorm(double p) { return ubm_qnorm(p); }

// ---- per-cell typed draws (random access into the replicate's Philox streams) ----------------------------------
struct PgCell {   // a cell = (sweep, observation); an attempt = 32 blocks of each typed stream (oracle/oracle_pg.hpp: CellDraws)
  uint64_t seed, rep, cell, base, rep_eff;
  unsigned ku, kn, ke;
  __device__ PgCell(uint64_t seed_, uint64_t grep_, uint64_t cell_) : seed(seed_), rep(grep_), cell(cell_) { set_attempt(0); }
  __device__ __forceinline__ void set_attempt(uint64_t a) {
    base = (1ull << 59) | ((cell & ((1ull << 26) - 1)) << 33) | ((a & ((1ull << 28) - 1)) << 5);
    rep_eff = rep + ((a >> 28) << 44);
    ku = kn = ke = 0;
  }
  __device__ __forceinline__ double u() { return pg_cell_u(seed, rep_eff, base, (ku++) & 127u); }
  __device__ __forceinline__ double n() { return pg_qnorm(pg_cell_u53(seed, rep_eff, base, (kn++) & 63u, UBM_STREAM_N)); }
  __device__ __forceinline__ double e() { return -pg_log(pg_cell_u53(seed, rep_eff, base, (ke++) & 63u, UBM_STREAM_E)); }
  __device__ __forceinline__ bool overflow() const { return false; }
};
// replay: the replicate's typed tapes, consumed sequentially (oracle/oracle_pg.hpp: SeqDraws over Draws)
struct PgSeq {
  const double *tu, *tn, *te;
  long long lu, ln, le;
  unsigned long long iu, in, ie;
  bool bad;
  __device__ __forceinline__ void set_attempt(uint64_t) {}
  __device__ double u() { const unsigned long long i = iu++; if ((long long)i >= lu) { bad = true; return 0.5; } return tu[i]; }
  __device__ double n() { const unsigned long long i = in++; if ((long long)i >= ln) { bad = true; return 0.0; } return tn[i]; }
  __device__ double e() { const unsigned long long i = ie++; if ((long long)i >= le) { bad = true; return 1.0; } return te[i]; }
  __device__ __forceinline__ bool overflow() const { return bad; }
};

// PolyaGamma::a — src/PolyaGamma.cpp:65-79.  log(pi / 2) and log((n + 1/2) pi), n < 8, do not depend on the draw: they are
// evaluated once per kernel (by the same ubm_log, so the values are the oracle's) instead of at every call.
#define PG_NLOGK 8
static __device__ double g_pg_logs[1 + PG_NLOGK];
__global__ void pg_log_table_kernel() {
  const int t = threadIdx.x;
  if (t == 0) g_pg_logs[0] = ubm_log(0.5 * PG_PI_D);
  else if (t <= PG_NLOGK) g_pg_logs[t] = ubm_log(((t - 1) + 0.5) * PG_PI_D);
}
__device__ __noinline__ double pg_a(int n, double x) {
  const double K = (n + 0.5) * PG_PI_D;
  double y = 0;
  if (x > PG_TRUNC_D) y = K * pg_exp(-0.5 * K * K * x);
  else if (x > 0) {
    const double logK = (n < PG_NLOGK) ? g_pg_logs[1 + n] : pg_log(K);
    const double expnt = -1.5 * (g_pg_logs[0] + pg_log(x)) + logK - 2.0 * (n + 0.5) * (n + 0.5) / x;
    y = pg_exp(expnt);
  }
  return y;
}
// PolyaGamma::mass_texpon — src/PolyaGamma.cpp:89-104
__device__ inline double pg_mass_texpon(double Z) {
  const double t = PG_TRUNC_D;
  const double fz = 0.125 * PG_PI_D * PG_PI_D + 0.5 * Z * Z;
  const double b = sqrt(1.0 / t) * (t * Z - 1);
  const double a = sqrt(1.0 / t) * (t * Z + 1) * -1.0;
  const double x0 = pg_log(fz) + fz * t;
  const double xb = x0 - Z + pg_pnorm_log(b);
  const double xa = x0 + Z + pg_pnorm_log(a);
  const double qdivp = 4 / PG_PI_D * (pg_exp(xb) + pg_exp(xa));
  return 1.0 / (1.0 + qdivp);
}
// PolyaGamma::rtigauss — src/PolyaGamma.cpp:106-139
template <class R>
__device__ inline double pg_rtigauss(double Z, R& r) {
  Z = fabs(Z);
  const double t = PG_TRUNC_D;
  double X = t + 1.0;
  if ((1.0 / 0.64) > Z) {
    double alpha = 0.0;
    while (r.u() > alpha) {
      double E1 = r.e(); double E2 = r.e();
      while (E1 * E1 > 2 * E2 / t) { E1 = r.e(); E2 = r.e(); if (r.overflow()) break; }
      X = 1 + E1 * t;
      X = t / (X * X);
      alpha = pg_exp(-0.5 * Z * Z * X);
      if (r.overflow()) break;
    }
  } else {
    const double mu = 1.0 / Z;
    while (X > t) {
      double Y = r.n(); Y *= Y;
      const double half_mu = 0.5 * mu;
      const double mu_Y = mu * Y;
      X = mu + half_mu * mu_Y - half_mu * sqrt(4 * mu_Y + mu_Y * mu_Y);
      if (r.u() > mu / (mu + X)) X = mu * mu / X;
      if (r.overflow()) break;
    }
  }
  return X;
}
// PolyaGamma::draw_like_devroye — src/PolyaGamma.cpp:175-226
template <class R>
__device__ inline double pg_draw(double Zin, R& r) {
  const double Z = fabs(Zin) * 0.5;
  const double fz = 0.125 * PG_PI_D * PG_PI_D + 0.5 * Z * Z;
  double X = 0.0, S = 1.0, Y = 0.0;
  const double mass = pg_mass_texpon(Z);
  while (true) {
    if (r.u() < mass) X = PG_TRUNC_D + r.e() / fz;
    else X = pg_rtigauss(Z, r);
    S = pg_a(0, X);
    Y = r.u() * S;
    int n = 0;
    bool go = true;
    while (go) {
      ++n;
      if (n % 2 == 1) {
        S = S - pg_a(n, X);
        if (Y <= S) return 0.25 * X;
      } else {
        S = S + pg_a(n, X);
        if (Y > S) go = false;
      }
    }
    if (r.overflow()) return 0.25 * X;
  }
}
// logcosh — src/logisticregressioncoupling.cpp:9-17
__device__ inline double pg_logcosh(double x) {
  if (x > 0.) return x + pg_log(1.0 + pg_exp(-2.0 * x)) - 0.6931472;
  return -x + pg_log(1.0 + pg_exp(2.0 * x)) - 0.6931472;
}

// ---- warp-level p x p linear algebra on column-major shared-memory matrices (stride p) --------------------------
// In place: the lower triangle of A becomes L (A = L L'), the strict upper part is zeroed.  Right-looking: after column k
// is scaled, every remaining entry (i, c), k < c <= i, takes its k-th update A[i][c] = fma(-L[i][k], L[c][k], A[i][c]) —
// exactly the fma sequence k = 0, 1, ... of the oracle's left-looking chol_lower for that entry, but the updates of one
// step are independent.  By the whole CTA, for one matrix (A2 == nullptr) or two at once (threads split in halves):
// thread (ti, tc) updates entries (k+1+ti+.., k+1+tc+..) of the trailing block, so a column step costs the pivot chain
// (sqrt, one division) plus two or three independent fma per thread and two CTA barriers (a single warp walking ~40
// entries per lane took 2.3 K cycles per column).  `dg` [2][PG_MAXP] keeps the new diagonal until the end, so that the
// pivot can be read and the column scaled between the same two barriers.
__device__ inline void pg_cta_chol(double* A1, double* A2, double* dg, int p, int tid) {
  const bool two = (A2 != nullptr);
  const int nthr = two ? PG_T / 2 : PG_T;
  const int half = two ? tid / (PG_T / 2) : 0;
  const int t = two ? tid % (PG_T / 2) : tid;
  double* A = half ? A2 : A1;
  double* d = dg + half * PG_MAXP;
  const int tc = t & 15, ti = t >> 4, nti = nthr >> 4;
  for (int k = 0; k < p; ++k) {
    const double lkk = sqrt(A[k * p + k]);
    if (t == 0) d[k] = lkk;
    for (int i = k + 1 + t; i < p; i += nthr) A[k * p + i] = A[k * p + i] / lkk;
    for (int i = t; i < k; i += nthr) A[k * p + i] = 0.0;
    __syncthreads();
    for (int i = k + 1 + ti; i < p; i += nti) {
      const double nl = -A[k * p + i];
      for (int c = k + 1 + tc; c <= i; c += 16) A[c * p + i] = ubm_fma(nl, A[k * p + c], A[c * p + i]);
    }
    __syncthreads();
  }
  for (int k = t; k < p; k += nthr) A[k * p + k] = d[k];
  __syncthreads();
}
// solve L y = rhs (forward), y overwrites `v` (shared); terms taken in increasing column order.  The reciprocal diagonal
// sits in registers (lane k mod 32) and the column of the next step is requested one step ahead, so a step is one
// shuffle pair, one multiplication and one fma deep (oracle: y_i = s_i * (1 / L_ii)).
__device__ __noinline__ void pg_warp_forward(const double* L, int p, int lane, double* v) {
  double s[2], dg[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = lane + 32 * q;
    s[q] = (i < p) ? v[i] : 0.0;
    dg[q] = (i < p) ? 1.0 / L[i * p + i] : 1.0;
  }
  double c0 = (lane < p) ? L[lane] : 0.0, c1 = (lane + 32 < p) ? L[lane + 32] : 0.0;          // column 0
  for (int k = 0; k < p; ++k) {
    const double sk = __shfl_sync(0xffffffffu, (k >= 32) ? s[1] : s[0], k & 31);
    const double dk = __shfl_sync(0xffffffffu, (k >= 32) ? dg[1] : dg[0], k & 31);
    double n0 = 0.0, n1 = 0.0;
    if (k + 1 < p) { if (lane < p) n0 = L[(k + 1) * p + lane]; if (lane + 32 < p) n1 = L[(k + 1) * p + lane + 32]; }
    const double yk = sk * dk;
    if (lane == k) s[0] = yk; else if (lane > k && lane < p) s[0] = ubm_fma(-c0, yk, s[0]);
    if (lane + 32 == k) s[1] = yk; else if (lane + 32 > k && lane + 32 < p) s[1] = ubm_fma(-c1, yk, s[1]);
    c0 = n0; c1 = n1;
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) { const int i = lane + 32 * q; if (i < p) v[i] = s[q]; }
  __syncwarp();
}
// solve L' x = y (backward), in place in `v`; terms taken from the last row upwards (as the oracle)
__device__ __noinline__ void pg_warp_backward(const double* L, int p, int lane, double* v) {
  double s[2], dg[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = lane + 32 * q;
    s[q] = (i < p) ? v[i] : 0.0;
    dg[q] = (i < p) ? 1.0 / L[i * p + i] : 1.0;
  }
  // row k of L, element i: L(k, i) = L[i * p + k]
  double c0 = (lane < p) ? L[lane * p + p - 1] : 0.0, c1 = (lane + 32 < p) ? L[(lane + 32) * p + p - 1] : 0.0;
  for (int k = p - 1; k >= 0; --k) {
    const double sk = __shfl_sync(0xffffffffu, (k >= 32) ? s[1] : s[0], k & 31);
    const double dk = __shfl_sync(0xffffffffu, (k >= 32) ? dg[1] : dg[0], k & 31);
    double n0 = 0.0, n1 = 0.0;
    if (k > 0) { if (lane < p) n0 = L[lane * p + k - 1]; if (lane + 32 < p) n1 = L[(lane + 32) * p + k - 1]; }
    const double xk = sk * dk;
    if (lane == k) s[0] = xk; else if (lane < k) s[0] = ubm_fma(-c0, xk, s[0]);
    if (lane + 32 == k) s[1] = xk; else if (lane + 32 < k) s[1] = ubm_fma(-c1, xk, s[1]);
    c0 = n0; c1 = n1;
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) { const int i = lane + 32 * q; if (i < p) v[i] = s[q]; }
  __syncwarp();
}
// Li = L^-1 (lower) of one or two factors, by the warps `w0 .. w0 + nw - 1`: a warp takes whole columns j (the columns are
// independent forward substitutions L x = e_j) and runs them right-looking with its lanes on the rows: once x_k is known
// every row i > k takes its term fma(L(i,k), x_k, s_i) — each row still sees k = j, j + 1, ... in order, the oracle's
// chain — and x_{k+1} = -s_{k+1} (1 / L_{k+1,k+1}) follows by one multiplication and a shuffle.  (One lane per column
// walking the left-looking loops was a single dependent chain of 1176 fma: 60 K cycles.)
__device__ inline void pg_tri_inverse_cols(const double* LA, double* LiA, const double* LB, double* LiB, int p, int warp, int lane,
                                           int w0, int nw) {
  const int ncol = LB ? 2 * p : p;
  for (int cj = warp - w0; cj < ncol; cj += nw) {
    const bool second = cj >= p;
    const double* L = second ? LB : LA;
    double* Li = second ? LiB : LiA;
    const int j = second ? cj - p : cj;
    double s[2] = {0.0, 0.0}, rd[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) { const int i = lane + 32 * q; rd[q] = (i < p) ? 1.0 / L[i * p + i] : 0.0; }
    for (int i = lane; i < j; i += 32) Li[j * p + i] = 0.0;
    double yk = __shfl_sync(0xffffffffu, (j >= 32) ? rd[1] : rd[0], j & 31);     // Li(j,j) = 1 / L(j,j)
    double c0 = (lane > j && lane < p) ? L[j * p + lane] : 0.0, c1 = (lane + 32 > j && lane + 32 < p) ? L[j * p + lane + 32] : 0.0;
    for (int k = j; k < p; ++k) {
      if (lane == (k & 31)) Li[j * p + k] = yk;
      if (k + 1 == p) break;
      double n0 = 0.0, n1 = 0.0;                                                  // column k + 1 of L, one step ahead
      if (lane > k + 1 && lane < p) n0 = L[(k + 1) * p + lane];
      if (lane + 32 > k + 1 && lane + 32 < p) n1 = L[(k + 1) * p + lane + 32];
      s[0] = ubm_fma(c0, yk, s[0]);                                                // rows <= k hold c = 0: s stays untouched where it matters
      s[1] = ubm_fma(c1, yk, s[1]);
      const double cand0 = -s[0] * rd[0], cand1 = -s[1] * rd[1];
      yk = __shfl_sync(0xffffffffu, (k + 1 >= 32) ? cand1 : cand0, (k + 1) & 31);
      c0 = n0; c1 = n1;
    }
  }
}
// S = Li' Li (both triangles), S[a,c] = sum_{i >= c} Li(i,a) Li(i,c), c >= a; entries strided over `nthr` threads
__device__ inline void pg_ltl(const double* Li, double* S, int p, int t, int nthr) {
  const int ne = p * (p + 1) / 2;
  for (int e = t; e < ne; e += nthr) {
    // e -> (a, c), a <= c, enumerated a-major
    int a = 0, rem = e;
    while (rem >= p - a) { rem -= p - a; ++a; }
    const int c = a + rem;
    double s = 0.0;
    for (int i = c; i < p; ++i) s = ubm_fma(Li[a * p + i], Li[c * p + i], s);
    S[c * p + a] = s; S[a * p + c] = s;
  }
}
// fast_dmvnorm_ (src/mvnorm.cpp:49-67) with the lower factor Ls of the covariance; tmp: shared scratch [p]
__device__ __noinline__ double pg_warp_dmvnorm(const double* x, const double* mean, const double* Ls, int p, int lane, double* tmp) {
  for (int i = lane; i < p; i += 32) tmp[i] = x[i] - mean[i];
  __syncwarp();
  pg_warp_forward(Ls, p, lane, tmp);
  double res = 0.0;
  // the logs of the diagonal are taken by all lanes at once and then added in the reference order by lane 0
  double lg[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) { const int j = lane + 32 * q; lg[q] = (j < p) ? pg_log(Ls[j * p + j]) : 0.0; }
  double halflogdet = 0.0;
  for (int j = 0; j < p; ++j) {
    const double t = __shfl_sync(0xffffffffu, (j >= 32) ? lg[1] : lg[0], j & 31);
    halflogdet = halflogdet + t;
  }
  if (lane == 0) {
    const double cst = -(halflogdet) - (p * 0.9189385);
    double sq = 0.0;
    for (int i = 0; i < p; ++i) sq = ubm_fma(tmp[i], tmp[i], sq);
    res = -0.5 * sq + cst;
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, res, 0);
}

struct PgShared {   // scalars
  long long rep;
  unsigned long long in, iu, ie;   // sequential region of the N / U (/ E: replay) streams
  int flag;
  int bad;                         // replay tape exhausted
  int eq;                          // replay: all w pairs equal
};

// Gram build for one or two chains: A_c[j2, j1] (lower) = invB[j1, j2] + sum_i (X(i,j1) w_c[i]) X(i,j2), j1 <= j2, on the
// FP64 tensor pipe: 8 x 8 output tiles (tile row <= tile column), mma.m8n8k4 over groups of four observations.  A DMMA
// accumulates its four products as an fma chain in observation order (tools/dmma_probe.cu), so an entry's sum is exactly
// the oracle's (m_sigma): four chains over the observation groups g = i / 4 taken by class g mod 4, each sequential in
// i, every term one fma of the rounded product X(i,j1) w[i] with X(i,j2); then ((((invB + P0) + P1) + P2) + P3).
// Warp w holds class w mod 4 of the tiles of subset w / 4 (row-major runs of the tile list) in registers: at most
// PG_MAXTL accumulators of two doubles per chain.  Operands come from the staged 32-observation tile of the padded
// row-major copy of X (row stride 68 doubles: the fragment loads of a half-warp hit 16 distinct bank pairs).
#define PG_PLIM 60               // ceil(p / 8) <= 8 tile rows: 36 tiles, 9 per subset (PG_MAXTL = 9; 7 up to p = 56, the German-credit size)
__device__ __forceinline__ void pg_dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#ifdef UBM_PG_TIMING
static __device__ long long g_pg_gt[4];
#endif
template <bool TWO, int PG_MAXTL>
__device__ inline void pg_gram(const PgDevice& M, const double* w1, const double* w2, double* A1, double* A2,
                               double* xs /* [2][PG_TILE][PG_GS] */, int tid) {
  const int n = M.n, p = M.p;
  const int lane = tid & 31, warp = tid >> 5;
  const int cls = warp & 3, sub = warp >> 2;
  const int nt = (p + 7) / 8, ntile = nt * (nt + 1) / 2, per = (ntile + 3) / 4;
  const int t0 = sub * per, cnt = (ntile - t0 < per) ? ((ntile - t0 > 0) ? ntile - t0 : 0) : per;
  // this warp's tiles: (ta, tb) of tile t0 + t in row-major order of the upper triangle of tiles
  int ta[PG_MAXTL], tb[PG_MAXTL];
  {
    int a = 0, rem = t0;
    while (a < nt && rem >= nt - a) { rem -= nt - a; ++a; }
    int b = a + rem;
#pragma unroll
    for (int t = 0; t < PG_MAXTL; ++t) {
      ta[t] = (a < nt) ? a : 0; tb[t] = (a < nt) ? b : 0;
      if (++b >= nt) { ++a; b = a; }
    }
  }
  double a1[PG_MAXTL][2], a2[PG_MAXTL][2];
#pragma unroll
  for (int t = 0; t < PG_MAXTL; ++t) { a1[t][0] = a1[t][1] = 0.0; a2[t][0] = a2[t][1] = 0.0; }
  const int kk = lane & 3, rr = lane >> 2;                // fragment coordinates: observation kk of the group, column rr of the tile
  // tiles of 32 observations of the row-major padded copy of X, double-buffered in shared memory
  constexpr int NV = PG_TILE * PG_GS / 2;                 // double2 per tile
  constexpr int PER = (NV + PG_T - 1) / PG_T;
  __syncthreads();
  {
    const double2* src = reinterpret_cast<const double2*>(M.Xt);
    double2* dst = reinterpret_cast<double2*>(xs);
    for (int q = tid; q < NV; q += PG_T) dst[q] = src[q];
  }
  __syncthreads();
  int buf = 0;
  for (int i0 = 0; i0 < n; i0 += PG_TILE, buf ^= 1) {
    const int rows = (n - i0 < PG_TILE) ? (n - i0) : PG_TILE;
    const double* xt = xs + buf * (PG_TILE * PG_GS);
    const bool more = i0 + PG_TILE < n;
#ifdef UBM_PG_TIMING
    long long tg0 = clock64();
#endif
    if (more) {     // the next tile goes straight from L2 to the other buffer (cp.async: no registers, no wait here)
      const double2* src = reinterpret_cast<const double2*>(M.Xt + (size_t)(i0 + PG_TILE) * PG_GS);
      double2* dst = reinterpret_cast<double2*>(xs + (buf ^ 1) * (PG_TILE * PG_GS));
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int q = tid + u * PG_T;
        if (q < NV) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"((unsigned)__cvta_generic_to_shared(dst + q)), "l"(src + q));
      }
      asm volatile("cp.async.commit_group;\n" ::);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = cls + 4 * h;                          // this warp's groups of the tile: local groups cls and cls + 4
      if (4 * g < rows) {
        const int ii = 4 * g + kk;                        // rows past n are zero in the padded copy
        const double* xr = xt + ii * PG_GS + rr;
        const double wa = (i0 + ii < n) ? w1[i0 + ii] : 0.0;
        const double wb = (TWO && i0 + ii < n) ? w2[i0 + ii] : 0.0;
        double xa = 0.0, ya = 0.0, yb = 0.0;
        int last = -1;
#pragma unroll
        for (int t = 0; t < PG_MAXTL; ++t) {
          if (t < cnt) {
            if (ta[t] != last) { last = ta[t]; xa = xr[8 * last]; ya = xa * wa; if (TWO) yb = xa * wb; }
            const double xb = xr[8 * tb[t]];
            pg_dmma(a1[t][0], a1[t][1], ya, xb);
            if (TWO) pg_dmma(a2[t][0], a2[t][1], yb, xb);
          }
        }
      }
    }
#ifdef UBM_PG_TIMING
    long long tg1 = clock64();
#endif
    if (more) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#ifdef UBM_PG_TIMING
    long long tg2 = clock64();
#endif
    __syncthreads();
#ifdef UBM_PG_TIMING
    if (tid == 0 && blockIdx.x == 0) { long long tg3 = clock64(); g_pg_gt[0] += tg1 - tg0; g_pg_gt[1] += tg2 - tg1; g_pg_gt[2] += tg3 - tg2; g_pg_gt[3] += 1; }
#endif
  }
  // ((((invB + P0) + P1) + P2) + P3): the four classes add their sums in turn.  Accumulator fragment: rows rr, columns 2 kk + {0, 1}
  for (int turn = 0; turn < 4; ++turn) {
    __syncthreads();
    if (cls == turn) {
#pragma unroll
      for (int t = 0; t < PG_MAXTL; ++t)
        if (t < cnt) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int j1 = 8 * ta[t] + rr, j2 = 8 * tb[t] + 2 * kk + c;
            if (j1 <= j2 && j2 < p) {
              const double base = (turn == 0) ? M.invB[(size_t)j2 * p + j1] : A1[j1 * p + j2];
              const double v1 = base + a1[t][c];
              A1[j1 * p + j2] = v1;
              if (turn == 3) A1[j2 * p + j1] = v1;
              if (TWO) {
                const double base2 = (turn == 0) ? M.invB[(size_t)j2 * p + j1] : A2[j1 * p + j2];
                const double v2 = base2 + a2[t][c];
                A2[j1 * p + j2] = v2;
                if (turn == 3) A2[j2 * p + j1] = v2;
              }
            }
          }
        }
    }
  }
  __syncthreads();
}

template <bool TAPE>
__global__ void __launch_bounds__(PG_T) pg_driver_kernel(const PgDevice M, const RunParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = M.n, p = M.p;
  double* sp = reinterpret_cast<double*>(smem_raw);
  double* w1 = sp; sp += n;
  double* w2 = sp; sp += n;
  double* MA1 = sp; sp += p * p;   // chain 1: A -> L -> Sigma -> Ls
  double* MB1 = sp; sp += p * p;   //          Linv
  double* MA2 = sp; sp += p * p;
  double* MB2 = sp; sp += p * p;
  double* xs = sp; sp += 2 * PG_TILE * PG_GS;
  double* beta1 = sp; sp += PG_MAXP;
  double* beta2 = sp; sp += PG_MAXP;
  double* m1 = sp; sp += PG_MAXP;
  double* m2 = sp; sp += PG_MAXP;
  double* x1 = sp; sp += PG_MAXP;
  double* x2 = sp; sp += PG_MAXP;
  double* tmp = sp; sp += PG_MAXP;
  double* zz = sp; sp += PG_MAXP;
  double* acc = sp; sp += 4 * PG_MAXP;      // mcmc [2p], corr [2p]
  int* plist = reinterpret_cast<int*>(sp); sp += (n + 1) / 2;   // observations whose first coupling test failed
  PgShared& sh = *reinterpret_cast<PgShared*>(sp);
  const int dimh = (P.h_kind == UBM_H_BOTH) ? 2 * p : p;
  const double denom = (double)(P.m - P.k + 1);

  for (;;) {
    if (tid == 0) {
      const unsigned long long got = atomicAdd(P.work_counter, 1ull);
      sh.rep = (got < (unsigned long long)P.nrep) ? (long long)got : -1;
      sh.in = 0; sh.iu = 0; sh.ie = 0; sh.bad = 0;
    }
    __syncthreads();
    const long long rep = sh.rep;
    if (rep < 0) break;
    const uint64_t grep = (uint64_t)(P.rep_offset + rep);
    // ---- rinit (germancredit.run.R:47-49): beta ~ N(b, B) via fast_rmvnorm (src/mvnorm.cpp:9-26), chain 1 then chain 2
    if (warp == 0) {
      for (int c = 0; c < 2; ++c) {
        bool badr = false;
        for (int j = lane; j < p; j += 32) zz[j] = draw_n_at<TAPE>(P, rep, (uint64_t)(c * p + j), &badr);
        if (TAPE && __any_sync(0xffffffffu, badr) && lane == 0) sh.bad = 1;
        __syncwarp();
        for (int j = lane; j < p; j += 32) {
          double a = 0.0;
          for (int i = 0; i <= j; ++i) a = ubm_fma(zz[i], M.cholB[(size_t)j * p + i], a);
          (c ? beta2 : beta1)[j] = a + M.b[j];
        }
        __syncwarp();
      }
      if (lane == 0) sh.in = 2ull * p;
    }
    for (int e = tid; e < 4 * PG_MAXP; e += PG_T) acc[e] = 0.0;
    __syncthreads();
    if (P.mode == MODE_ESTIMATOR && P.k == 0 && tid < p) {
      const double x = beta1[tid];
      if (P.h_kind == UBM_H_IDENTITY) acc[tid] = x;
      else if (P.h_kind == UBM_H_SQUARE) acc[tid] = x * x;
      else { acc[tid] = x; acc[p + tid] = x * x; }
      if (P.nbreaks >= 2 && tid == P.component) {       // estimator_bin_ numerators (src/estimator_bin.cpp:15-38)
        const int nb = P.nbreaks - 1;
        const int b = bin_of(x, P.breaks, P.nbreaks, P.breaks[0], (double)nb / (P.breaks[nb] - P.breaks[0]));
        if (b >= 0) P.bin_num[rep * nb + b] += 1;
      }
    }
    if (P.mode == MODE_CHAINS && tid < p) {
      P.samples1[(rep * P.stride) * p + tid] = beta1[tid];
      P.samples2[(rep * P.stride) * p + tid] = beta2[tid];
    }
    long long time = 0, tau = 0;
    bool met = false;
#ifdef UBM_PG_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64(), tsteps = 0, tcoupled = 0;
#define PGMARK(i) do { long long now_ = clock64(); tacc[i] += now_ - tprev; tprev = now_; } while (0)
#else
#define PGMARK(i) do { } while (0)
#endif
    // ---- time loop (generic drivers; CTA-uniform control flow)
    for (;;) {
      bool go;
      if (time < P.lag) go = true;
      else {
        const bool capped = (P.max_it > 0 && time >= P.max_it) || (TAPE && sh.bad);
        if (P.mode == MODE_MEETING) go = !met && !capped;
        else go = (!met || time < (tau > P.m ? tau : P.m)) && !capped;
      }
      if (TAPE && sh.bad) go = false;
      if (!go) break;
      time += 1;
      const bool coupled_step = (time > P.lag) && !met;
      const unsigned long long sweep = (unsigned long long)time;
      __syncthreads();
#ifdef UBM_PG_TIMING
      tprev = clock64(); tsteps++; if (coupled_step) tcoupled++;
#endif
      // ---- 1. PG sweep.  Phase 1: one thread per observation draws w1 (attempt 0 of the cell) and runs the first
      // test of the max coupling (src/logisticregressioncoupling.cpp:87-98); failures go to a work list.
      int all_equal = 1;
      if (tid == 0) sh.flag = 0;
      __syncthreads();
      if (TAPE) {
        // replay: z_i for both chains by all threads, then thread 0 walks the observations on the sequential stream
        for (int i = tid; i < n; i += PG_T) {
          double xb1 = 0.0, xb2 = 0.0;
          for (int j = 0; j < p; ++j) {
            const double xv = M.X[(size_t)j * n + i];
            xb1 = ubm_fma(xv, beta1[j], xb1);
            if (coupled_step) xb2 = ubm_fma(xv, beta2[j], xb2);
          }
          w1[i] = fabs(xb1); w2[i] = fabs(xb2);
        }
        __syncthreads();
        if (tid == 0) {
          PgSeq r;
          r.tu = P.tape_u + rep * P.len_u; r.tn = P.tape_n + rep * P.len_n; r.te = P.tape_e + rep * P.len_e;
          r.lu = P.tape_u ? P.len_u : 0; r.ln = P.tape_n ? P.len_n : 0; r.le = P.tape_e ? P.len_e : 0;
          r.iu = sh.iu; r.in = sh.in; r.ie = sh.ie; r.bad = false;
          int eq = 1;
          for (int i = 0; i < n && !r.bad; ++i) {
            const double z1 = w1[i], z2 = w2[i];
            if (!coupled_step) { w1[i] = pg_draw(z1, r); continue; }
            double wa, wb;
            if (M.w_coupling == 1) {                                      // w_rejsamplerC (:50-74)
              const double z_min = z1 < z2 ? z1 : z2, z_max = z1 < z2 ? z2 : z1;
              const double w_min = pg_draw(z_min, r);
              const double log_u = pg_log(r.u());
              const double log_ratio = -0.5 * w_min * (z_max * z_max - z_min * z_min);
              double w_max;
              if (log_u < log_ratio) w_max = w_min;
              else w_max = pg_draw(z_max, r);
              wa = (z1 < z2) ? w_min : w_max; wb = (z1 < z2) ? w_max : w_min;
            } else {                                                      // w_max_couplingC (:87-113)
              wa = pg_draw(z1, r);
              const double u1 = r.u();
              const double la1 = pg_logcosh(z2 / 2.) - 0.5 * z2 * z2 * wa - (pg_logcosh(z1 / 2.) - 0.5 * z1 * z1 * wa);
              if (pg_log(u1) <= la1) wb = wa;
              else {
                bool accept = false;
                wb = wa;
                while (!accept && !r.bad) {
                  wb = pg_draw(z2, r);
                  const double u2 = r.u();
                  const double la2 = pg_logcosh(z1 / 2.) - 0.5 * z1 * z1 * wb - (pg_logcosh(z2 / 2.) - 0.5 * z2 * z2 * wb);
                  if (pg_log(u2) > la2) accept = true;
                }
              }
            }
            w1[i] = wa; w2[i] = wb;
            if (wa != wb) eq = 0;
          }
          sh.iu = r.iu; sh.in = r.in; sh.ie = r.ie; sh.eq = eq;
          if (r.bad) sh.bad = 1;
        }
        __syncthreads();
        all_equal = sh.eq;
      }
      for (int i = tid; i < n && !TAPE; i += PG_T) {
        double xb1 = 0.0, xb2 = 0.0;
        for (int j = 0; j < p; ++j) {                                     // xbeta_ (logisticregressioncoupling.cpp:22-32)
          const double xv = M.X[(size_t)j * n + i];
          xb1 = ubm_fma(xv, beta1[j], xb1);
          if (coupled_step) xb2 = ubm_fma(xv, beta2[j], xb2);
        }
        PgCell cell(P.seed, grep, sweep * (unsigned long long)n + (unsigned long long)i);
        const double z1 = fabs(xb1);
        if (coupled_step && M.w_coupling == 1) {                          // w_rejsamplerC (logisticregressioncoupling.cpp:42-75)
          const double z2 = fabs(xb2);
          const double z_min = z1 < z2 ? z1 : z2, z_max = z1 < z2 ? z2 : z1;
          const double w_min = pg_draw(z_min, cell);
          const double log_u = pg_log(cell.u());
          const double log_ratio = -0.5 * w_min * (z_max * z_max - z_min * z_min);
          double w_max;
          if (log_u < log_ratio) w_max = w_min;
          else { cell.set_attempt(1); w_max = pg_draw(z_max, cell); }
          const double wa = (z1 < z2) ? w_min : w_max, wb = (z1 < z2) ? w_max : w_min;
          w1[i] = wa; w2[i] = wb;
          if (wa != wb) all_equal = 0;
          continue;
        }
        const double a1 = pg_draw(z1, cell);
        w1[i] = a1;
        if (coupled_step) {
          const double z2 = fabs(xb2);
          const double u1 = cell.u();
          const double la1 = pg_logcosh(z2 / 2.) - 0.5 * z2 * z2 * a1 - (pg_logcosh(z1 / 2.) - 0.5 * z1 * z1 * a1);
          if (pg_log(u1) <= la1) w2[i] = a1;
          else { all_equal = 0; plist[atomicAdd(&sh.flag, 1)] = i; }
        }
      }
      __syncthreads();
      // Phase 2: the rejection loop of the max coupling (:99-110), one warp per listed observation, 32 attempts at a
      // time (attempt a draws from its own region of the cell, so the first accepted attempt is the sequential answer)
      if (coupled_step && !TAPE) {
        const int nlist = sh.flag;
        for (int q = warp; q < nlist; q += PG_T / 32) {
          const int i = plist[q];
          double xb1 = 0.0, xb2 = 0.0;
          for (int j = 0; j < p; ++j) {
            const double xv = M.X[(size_t)j * n + i];
            xb1 = ubm_fma(xv, beta1[j], xb1);
            xb2 = ubm_fma(xv, beta2[j], xb2);
          }
          const double z1 = fabs(xb1), z2 = fabs(xb2);
          PgCell cell(P.seed, grep, sweep * (unsigned long long)n + (unsigned long long)i);
          for (unsigned long long base = 1;; base += 32) {
            cell.set_attempt(base + lane);
            const double a2 = pg_draw(z2, cell);
            const double u2 = cell.u();
            const double la2 = pg_logcosh(z1 / 2.) - 0.5 * z1 * z1 * a2 - (pg_logcosh(z2 / 2.) - 0.5 * z2 * z2 * a2);
            const unsigned okm = __ballot_sync(0xffffffffu, pg_log(u2) > la2);
            if (okm) {
              const double win = __shfl_sync(0xffffffffu, a2, __ffs(okm) - 1);
              if (lane == 0) w2[i] = win;
              break;
            }
          }
        }
      }
      all_equal = __syncthreads_and(all_equal);
      PGMARK(0);
      // ---- 2. Gram build(s)
      if (p <= 56) {
        if (coupled_step) pg_gram<true, 7>(M, w1, w2, MA1, MA2, xs, tid);
        else pg_gram<false, 7>(M, w1, w2, MA1, MA2, xs, tid);
      } else {
        if (coupled_step) pg_gram<true, 9>(M, w1, w2, MA1, MA2, xs, tid);
        else pg_gram<false, 9>(M, w1, w2, MA1, MA2, xs, tid);
      }
      PGMARK(1);
      // ---- 3. m, L, Linv per chain (warp 0: chain 1, warp 1: chain 2)
      pg_cta_chol(MA1, coupled_step ? MA2 : nullptr, tmp, p, tid);   // src/logisticregression.cpp:49-50 (tmp, zz: scratch)
      PGMARK(3);
      // :55 lower.inverse(): the columns of both factors over all warps
      pg_tri_inverse_cols(MA1, MB1, coupled_step ? MA2 : nullptr, MB2, p, warp, lane, 0, PG_T / 32);
      __syncthreads();
      PGMARK(5);
      // :51 m = A^-1 b through the inverse factor: y = Linv b, m = Linv' y (independent dot products; oracle: m_sigma)
      {
        const int ch = tid >> 6, i = tid & 63;                     // threads 0..63: chain 1, 64..127: chain 2
        const bool on = ch < (coupled_step ? 2 : 1) && i < p;
        const double* Li = ch ? MB2 : MB1;
        double* yv = tmp + ch * PG_MAXP;
        if (on) {
          double a = 0.0;
          for (int k = 0; k <= i; ++k) a = ubm_fma(Li[k * p + i], M.ktk[k], a);
          yv[i] = a;
        }
        __syncthreads();
        if (on) {
          double a = 0.0;
          for (int r = i; r < p; ++r) a = ubm_fma(Li[i * p + r], yv[r], a);
          (ch ? m2 : m1)[i] = a;
        }
      }
      __syncthreads();
      PGMARK(4);
      if (!coupled_step) {
        // single kernel: beta = fast_rmvnorm_chol(1, m, Linv) (germancredit.run.R:27-28): y_j = sum_{i >= j} z_i Linv(i, j)
        if (warp == 0) {
          const unsigned long long in0 = sh.in;
          bool badr = false;
          for (int j = lane; j < p; j += 32) zz[j] = draw_n_at<TAPE>(P, rep, in0 + j, &badr);
          if (TAPE && __any_sync(0xffffffffu, badr) && lane == 0) sh.bad = 1;
          __syncwarp();
          for (int j = lane; j < p; j += 32) {
            double a = 0.0;
            for (int i = j; i < p; ++i) a = ubm_fma(zz[i], MB1[j * p + i], a);
            beta1[j] = a + m1[j];
          }
          if (lane == 0) sh.in = in0 + p;
        }
      } else if (M.beta_coupling == 1) {
        // common random numbers (pgg_coupled_kernel, vignettes/polyagammagibbs.Rmd:203-206): one increment, beta_c = m_c + z Linv_c
        if (warp == 0) {
          const unsigned long long in0 = sh.in;
          bool badr = false;
          for (int j = lane; j < p; j += 32) zz[j] = draw_n_at<TAPE>(P, rep, in0 + j, &badr);
          if (TAPE && __any_sync(0xffffffffu, badr) && lane == 0) sh.bad = 1;
          __syncwarp();
          for (int j = lane; j < p; j += 32) {
            double a = 0.0, b2 = 0.0;
            for (int i = j; i < p; ++i) { a = ubm_fma(zz[i], MB1[j * p + i], a); b2 = ubm_fma(zz[i], MB2[j * p + i], b2); }
            const double v1 = a + m1[j];
            beta1[j] = v1;
            beta2[j] = all_equal ? v1 : (b2 + m2[j]);                      // :207-211: identical iff every PG pair is
          }
          if (lane == 0) sh.in = in0 + p;
        }
        if (all_equal) { met = true; tau = time; }
      } else {
        // sample_beta (R/logistic_regression.R:180-209): Sigma_c = Linv_c' Linv_c, then LLT of Sigma_c (src/mvnorm.cpp:12,52)
        pg_ltl(MB1, MA1, p, tid, PG_T);                                  // all warps: the entries are independent
        pg_ltl(MB2, MA2, p, tid, PG_T);
        __syncthreads();
        pg_cta_chol(MA1, MA2, tmp, p, tid);
        if (warp == 0) {                                                  // rmvnorm_max_coupling_ (src/mvnorm.cpp:91-129)
          unsigned long long in0 = sh.in, iu0 = sh.iu;
          bool badr = false;
          // x1 = z U1 + m1, U1 = Ls1': y_j = sum_{i <= j} z_i Ls1(j, i)
          for (int j = lane; j < p; j += 32) zz[j] = draw_n_at<TAPE>(P, rep, in0 + j, &badr);
          in0 += p;
          __syncwarp();
          for (int j = lane; j < p; j += 32) {
            double a = 0.0;
            for (int i = 0; i <= j; ++i) a = ubm_fma(zz[i], MA1[i * p + j], a);
            x1[j] = a + m1[j];
          }
          __syncwarp();
          const double logu = ubm_log(draw_u_at<TAPE>(P, rep, iu0, &badr));   // :103-106
          iu0 += 1;
          const double d11 = pg_warp_dmvnorm(x1, m1, MA1, p, lane, tmp);
          const double d12 = pg_warp_dmvnorm(x1, m2, MA2, p, lane, tmp);
          if (d11 + logu < d12) {                                         // :107
            for (int j = lane; j < p; j += 32) x2[j] = x1[j];
          } else {
            bool accept = false;
            while (!accept && !(TAPE && __any_sync(0xffffffffu, badr))) {   // :113-123
              for (int j = lane; j < p; j += 32) zz[j] = draw_n_at<TAPE>(P, rep, in0 + j, &badr);
              in0 += p;
              __syncwarp();
              for (int j = lane; j < p; j += 32) {
                double a = 0.0;
                for (int i = 0; i <= j; ++i) a = ubm_fma(zz[i], MA2[i * p + j], a);
                x2[j] = a + m2[j];
              }
              __syncwarp();
              const double lu = ubm_log(draw_u_at<TAPE>(P, rep, iu0, &badr));
              iu0 += 1;
              const double d22 = pg_warp_dmvnorm(x2, m2, MA2, p, lane, tmp);
              const double d21 = pg_warp_dmvnorm(x2, m1, MA1, p, lane, tmp);
              if (d22 + lu > d21) accept = true;
            }
          }
          __syncwarp();
          for (int j = lane; j < p; j += 32) {
            beta1[j] = x1[j];
            beta2[j] = all_equal ? x1[j] : x2[j];                         // germancredit.run.R:38-41
          }
          if (lane == 0) { sh.in = in0; sh.iu = iu0; }
          if (TAPE && __any_sync(0xffffffffu, badr) && lane == 0) sh.bad = 1;
        }
        if (all_equal) { met = true; tau = time; }
      }
      __syncthreads();
      PGMARK(2);
      // ---- estimator / trajectories
      if (tid < p) {
        const double xa = beta1[tid];
        const double xb = (time > P.lag && !coupled_step) ? xa : beta2[tid];   // after meeting chain 2 copies chain 1
        if (P.mode == MODE_ESTIMATOR) {
          const bool in_window = (time <= P.lag) ? (time >= P.k) : (P.k <= time && time <= P.m);
          const bool corr_now = (time >= P.k + P.lag) && (coupled_step || time == P.lag);
          const double wgt = corr_now ? (double)corr_weight(time, P.k, P.m, P.lag) : 0.0;
          const int nh = (P.h_kind == UBM_H_BOTH) ? 2 : 1;
          for (int hh = 0; hh < nh; ++hh) {
            const bool sq = (P.h_kind == UBM_H_SQUARE) || hh == 1;
            const int e = hh * p + tid;
            const double h1 = sq ? xa * xa : xa;
            if (in_window) acc[e] = acc[e] + h1;
            if (corr_now) { const double h2 = sq ? xb * xb : xb; acc[2 * PG_MAXP + e] = acc[2 * PG_MAXP + e] + wgt * (h1 - h2); }
          }
          if (P.nbreaks >= 2 && tid == P.component) {   // the component's histogram: this thread is the only writer of the row
            const int nb = P.nbreaks - 1;
            const double b0 = P.breaks[0], iw = (double)nb / (P.breaks[nb] - P.breaks[0]);
            long long* bn = P.bin_num + rep * nb;
            const int b1 = bin_of(xa, P.breaks, P.nbreaks, b0, iw);
            if (in_window && b1 >= 0) bn[b1] += 1;
            if (corr_now) {
              const long long w = corr_weight(time, P.k, P.m, P.lag);
              const int b2 = bin_of(xb, P.breaks, P.nbreaks, b0, iw);
              if (b1 >= 0) bn[b1] += w;
              if (b2 >= 0) bn[b2] -= w;
            }
          }
        } else if (P.mode == MODE_CHAINS) {
          P.samples1[(rep * P.stride + time) * p + tid] = xa;
          if (time > P.lag) P.samples2[(rep * P.stride + (time - P.lag)) * p + tid] = xb;
        }
      }
    }
    __syncthreads();
#ifdef UBM_PG_TIMING
    if (tid == 0 && blockIdx.x == 0)
      printf("gram tile iterations %lld: dmma section %.0f store %.0f barrier %.0f cycles each\n", g_pg_gt[3], (double)g_pg_gt[0] / g_pg_gt[3],
             (double)g_pg_gt[1] / g_pg_gt[3], (double)g_pg_gt[2] / g_pg_gt[3]);
    if (tid == 0 && blockIdx.x == 0)
      printf("pg timing (cycles/step, block 0): steps %lld coupled %lld pg-sweep %.0f gram %.0f | chol %.0f solves %.0f tri-inverse %.0f rest(sample, coupled linalg) %.0f\n", tsteps,
             tcoupled, (double)tacc[0] / tsteps, (double)tacc[1] / tsteps, (double)tacc[3] / tsteps, (double)tacc[4] / tsteps,
             (double)tacc[5] / tsteps, (double)tacc[2] / tsteps);
#endif
    // ---- results
    if (TAPE && sh.bad) { met = false; if (tid == 0) atomicExch(P.status, UBM_ERR_TAPE); }
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
      const double mc = acc[tid], co = acc[2 * PG_MAXP + tid];
      const double ue = mc + co;
      P.uest[rep * dimh + tid] = ue / denom;
      if (P.mcmc) P.mcmc[rep * dimh + tid] = mc / denom;
      if (P.corr) P.corr[rep * dimh + tid] = co / denom;
    }
    __syncthreads();
  }
}

inline size_t pg_smem_bytes(int n, int p) {
  return (size_t)(2 * n + 4 * p * p + 2 * PG_TILE * PG_GS + 8 * PG_MAXP + 4 * PG_MAXP + (n + 1) / 2) * sizeof(double) + sizeof(PgShared) + 16;
}

inline int pg_launch(const PgDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st, int64_t* launches,
                     std::string* err) {
  const size_t smem = pg_smem_bytes(M.n, M.p);
  if (smem > 227 * 1024) { *err = "pg logistic: n / p too large for shared memory"; return UBM_ERR_UNSUPPORTED; }
  auto kern = tape ? pg_driver_kernel<true> : pg_driver_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = std::string("pg logistic: smem attribute: ") + cudaGetErrorString(e); return UBM_ERR_CUDA; }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PG_T, smem);
  if (e != cudaSuccess || per_sm < 1) { *err = "pg logistic: occupancy query failed"; return UBM_ERR_CUDA; }
  const int grid = (int)std::min<int64_t>((int64_t)nsm * per_sm, P.nrep);
  pg_log_table_kernel<<<1, 32, 0, st>>>();
  kern<<<grid, PG_T, smem, st>>>(M, P);
  *launches += 2;
  return UBM_OK;
}

}  // namespace ubm
