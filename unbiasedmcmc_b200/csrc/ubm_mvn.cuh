// ubm_mvn.cuh — C2: multivariate Normal target, RW-MH with reflection-max coupled proposals.
//
// Replaces, fused into one persistent kernel per launch:
//   drivers    R/meetingtime.R:24-48, R/coupled_chains.R:39-87, R/unbiasedestimator.R:43-104
//   kernels    R/get_mh_kernels.R:29-77 (d > 1 branch)
//   numerics   src/mvnorm.cpp:30-46 (fast_rmvnorm_cholesky_), :71-86 (fast_dmvnorm_cholesky_inverse_),
//              :185-236 (rmvnorm_reflection_max_coupling_), src/estimator_bin.cpp:10-42 (on the fly)
//
// Design (B200): a CTA runs R = 16 replicate "slots" in lock step, so that the five vector x triangular-matrix products of
// a coupled step become three small GEMM phases ([rows x d] . [d x d]) — and those run on the FP64 TENSOR pipe:
//   phase A  z = -(x2 - x1)^T Cinv_prop            (coupled slots)
//   phase B  prop = x + xi^T C_prop  (and eta)     (all active rows)
//   phase C  w = Cinv_pi^T (prop - mean_pi)        (all proposals) -> log-density
// One DMMA (mma.sync m8n8k4 f64) multiplies 8 slot rows x 4 k-values into an 8-column output tile.  It accumulates its
// four products as a chain of fused multiply-adds in k order, one rounding each (measured: tools/dmma_probe.cu, 128000 of
// 128000 outputs equal the fma chain), so a column sum evaluated by a chain of DMMAs over k = 0, 4, 8, ... is bit for bit
// the sequential fma loop of the oracle (oracle_models.hpp: vec_times_upper) — replay parity survives the tensor cores.
// The three upper-triangular factors are packed as ready-made B fragments (tile (jt, ks): 32 doubles, lane l holds
// U[4 ks + l%4][8 jt + l/4]) and stay in SHARED MEMORY for the life of the kernel (140 KB at d = 100); slot rows are read
// as A fragments (row stride == 4 mod 16 doubles: conflict-free).  Row tiles are static: tile q = chain (q >> 1), slots
// 8 (q & 1) .. + 7; a tile is skipped when none of its rows takes part in the phase, rows that do not are masked at the
// write.  Output tiles jt (2 jt + 2 k-steps each) are handed to the eight consumer warps by a longest-first table built
// on the host and balanced per SM sub-partition (the tensor pipe is per sub-partition).
//
// Thread roles (512 threads, warp-specialised): warps 0-7 (two per SM sub-partition) are CONSUMERS: they run the GEMM
// phases and the coupling algebra, synchronising among themselves on a named barrier; warps 8-15 are PRODUCERS: while
// step t is computed they generate every draw of step t+1 (the d normals per slot through Philox + AS241, the rare tail
// branch compacted over the warp, and the log-uniforms, into a small ring), so no random-number work sits on the critical
// path.  All 16 warps share the short per-slot section at the step boundary (accept/reject, state and estimator update,
// slot state machine): warp s owns slot s and keeps its estimator sums in registers.  Slots are refilled from a global
// counter as soon as a replicate finishes (work stealing).  Reductions over a d-vector use the canonical order "fma chain
// inside each 4-column group, xor-butterfly over the <= 32 groups" (oracle: sum_g4_*).  Two CTA-wide barriers per step.
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "ubm_common.cuh"

namespace ubm {

#define MVN_MAXR 16
#define MVN_CW 8       // consumer warps: warps 0-7 (two per SM sub-partition)
#define MVN_PW 8       // producer warps: warps 8-15 produce the draws of the next step
#define MVN_GT (32 * MVN_CW)               // consumer (GEMM) threads
#define MVN_T (32 * (MVN_CW + MVN_PW))     // threads per CTA
#define MVN_JMAX 2     // output tiles per consumer warp (d <= 128: 16 tiles over 8 warps)

struct MvnDevice {
  int d, DPS, nNT, nCG, zero_mean;
  int packLen;                 // doubles per packed matrix
  double cst;                  // src/mvnorm.cpp:74-75
  const double *mean_pi, *init_mean, *packA, *packB, *packC, *packI;   // device
  size_t offs[6];              // host-side: offsets inside the blob
  unsigned char jobs[MVN_CW][MVN_JMAX];    // output tiles of consumer warp w (255: none), the longer one first
};

// packed B fragments: tile (jt, ks), ks = 0 .. 2 jt + 1, at index jt (jt + 1) + ks; lane l of the tile holds
// U[k = 4 ks + (l & 3)][n = 8 jt + (l >> 2)], zero below the diagonal and past d.
__host__ __device__ __forceinline__ int mvn_tile_base(int jt) { return jt * (jt + 1); }
__host__ __device__ __forceinline__ int mvn_pack_len(int nNT) { return 32 * nNT * (nNT + 1); }

inline void mvn_pack_matrix(const double* Ucm, int d, int nNT, std::vector<double>* blob) {
  const size_t base = blob->size();
  blob->resize(base + mvn_pack_len(nNT), 0.0);
  for (int jt = 0; jt < nNT; ++jt)
    for (int ks = 0; ks < 2 * jt + 2; ++ks)
      for (int l = 0; l < 32; ++l) {
        const int k = 4 * ks + (l & 3), n = 8 * jt + (l >> 2);
        (*blob)[base + (size_t)(mvn_tile_base(jt) + ks) * 32 + l] = (n < d && k <= n) ? Ucm[(size_t)n * d + k] : 0.0;
      }
}

inline void mvn_pack_host(const ubm_model_mvn* s, MvnDevice* M, std::vector<double>* blob) {
  const int d = s->d;
  M->d = d; M->nNT = (d + 7) / 8; M->nCG = (d + 3) / 4;
  // row stride: >= 8 nNT (an A fragment of the last tile reads up to column 8 nNT - 1) and == 4 (mod 16) doubles, so that
  // the 4 rows x 4 consecutive doubles read by a half-warp fall into 16 distinct bank pairs
  M->DPS = 8 * M->nNT;
  while (M->DPS % 16 != 4) M->DPS += 4;
  M->zero_mean = 1;
  for (int j = 0; j < d; ++j) if (s->mean_pi[j] != 0.0) M->zero_mean = 0;
  double halflogdet = 0.0;
  for (int j = 0; j < d; ++j) halflogdet = halflogdet + ubm_log(s->chol_inv_pi[(size_t)j * d + j]);
  M->cst = -(-halflogdet) - (d * 0.9189385);     // truncated constant kept verbatim (src/mvnorm.cpp:75)
  M->packLen = mvn_pack_len(M->nNT);
  blob->clear();
  M->offs[0] = blob->size(); for (int j = 0; j < M->DPS; ++j) blob->push_back(j < d ? s->mean_pi[j] : 0.0);
  M->offs[1] = blob->size(); for (int j = 0; j < M->DPS; ++j) blob->push_back(j < d ? s->init_mean[j] : 0.0);
  M->offs[2] = blob->size(); mvn_pack_matrix(s->chol_inv_prop, d, M->nNT, blob);
  M->offs[3] = blob->size(); mvn_pack_matrix(s->chol_prop, d, M->nNT, blob);
  M->offs[4] = blob->size(); mvn_pack_matrix(s->chol_inv_pi, d, M->nNT, blob);
  M->offs[5] = blob->size(); mvn_pack_matrix(s->init_chol, d, M->nNT, blob);
  // output tiles -> consumer warps, folded longest-first: warp w takes the (w+1)-th longest tile and the (16-w)-th, so that
  // every warp runs about the same number of k-steps (a warp's K loop is a dependent chain of loads and DMMAs; the longest
  // warp sets the length of a phase)
  for (int w = 0; w < MVN_CW; ++w) {
    const int j0 = M->nNT - 1 - w, j1 = M->nNT - 1 - (2 * MVN_CW - 1 - w);
    M->jobs[w][0] = (j0 >= 0) ? (unsigned char)j0 : 255;
    M->jobs[w][1] = (j1 >= 0) ? (unsigned char)j1 : 255;
  }
}
inline void mvn_bind_device(MvnDevice* M, const double* dev) {
  M->mean_pi = dev + M->offs[0]; M->init_mean = dev + M->offs[1];
  M->packA = dev + M->offs[2]; M->packB = dev + M->offs[3]; M->packC = dev + M->offs[4]; M->packI = dev + M->offs[5];
}

enum { SL_IDLE = 0, SL_WAIT = 1, SL_INIT = 2, SL_SINGLE = 3, SL_COUPLED = 4 };

struct MvnSlots {   // per-slot scalars, in shared memory
  long long rep[MVN_MAXR], time[MVN_MAXR], tau[MVN_MAXR];
  unsigned long long iu[MVN_MAXR], in[MVN_MAXR], pu[MVN_MAXR];
  double pdf1[MVN_MAXR], pdf2[MVN_MAXR];
  int phase[MVN_MAXR], met[MVN_MAXR], ident[MVN_MAXR], bad[MVN_MAXR];
  int exhausted;
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MVN_GT) : "memory"); }

// D(8 x 8) += A(8 x 4) B(4 x 8): a = A[l / 4][l % 4], b = B[l % 4][l / 4], d0 / d1 = D[l / 4][2 (l % 4) + {0, 1}].
// The four products of an output are accumulated as an fma chain in k order (tools/dmma_probe.cu).
__device__ __forceinline__ void mvn_dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_butterfly(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// position of the (r + 1)-th set bit of m (r < popc(m)): a five-level popcount descent (__fns is a software loop, 5 % of the
// stall samples of the single-chain kernel)
__device__ __forceinline__ int mvn_nth_set_bit(unsigned m, int r) {
  int pos = 0;
  int c = __popc(m & 0xffffu);
  if (r >= c) { r -= c; pos = 16; m >>= 16; }
  c = __popc(m & 0xffu);
  if (r >= c) { r -= c; pos += 8; m >>= 8; }
  c = __popc(m & 0xfu);
  if (r >= c) { r -= c; pos += 4; m >>= 4; }
  c = __popc(m & 0x3u);
  if (r >= c) { r -= c; pos += 2; m >>= 2; }
  if (r >= (int)(m & 1u)) pos += 1;
  return pos;
}

// Four standard normals per lane: draws i0 .. i0 + 3 of replicate `rep`'s N stream (both per lane: the lanes of a warp may
// serve different replicates), written to dst[0..3] (shared memory; entries past nvalid are zeroed, lanes with nvalid <= 0 write
// nothing), executed by a whole warp.  AS241's central branch is evaluated in place for all four draws, branch-free: eight
// independent Horner chains (numerator, denominator) keep the FP64 pipe busy instead of one dependent chain at a time; the
// ~15 % of draws that fall in the tails are COMPACTED over the warp (ballot + rank), so that the log/sqrt branch runs once per
// warp for up to 32 tail draws instead of once per tail draw of the unluckiest lane.
__device__ __forceinline__ void draw_n4_lanes(const RunParams& P, long long rep, unsigned long long i0, int nvalid,
                                              double* dst, int lane) {
  const uint64_t grep = (uint64_t)(P.rep_offset + rep);
  double pr[4] = {0.5, 0.5, 0.5, 0.5};
  if (nvalid > 0) {
    const unsigned long long b0 = i0 >> 1;
    const ubm_u32x4 A0 = ubm_stream_block(P.seed, grep, UBM_STREAM_N, b0);
    const ubm_u32x4 A1 = ubm_stream_block(P.seed, grep, UBM_STREAM_N, b0 + 1);
    if (i0 & 1) {
      const ubm_u32x4 A2 = ubm_stream_block(P.seed, grep, UBM_STREAM_N, b0 + 2);
      pr[0] = ubm_u01_53(A0.w[2], A0.w[3]); pr[1] = ubm_u01_53(A1.w[0], A1.w[1]);
      pr[2] = ubm_u01_53(A1.w[2], A1.w[3]); pr[3] = ubm_u01_53(A2.w[0], A2.w[1]);
    } else {
      pr[0] = ubm_u01_53(A0.w[0], A0.w[1]); pr[1] = ubm_u01_53(A0.w[2], A0.w[3]);
      pr[2] = ubm_u01_53(A1.w[0], A1.w[1]); pr[3] = ubm_u01_53(A1.w[2], A1.w[3]);
    }
  }
  double zc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) zc[c] = ubm_qnorm_central(pr[c] - 0.5);
  unsigned bal[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double q = pr[c] - 0.5;
    const bool valid = c < nvalid;
    const bool tail = valid && !(fabs(q) <= 0.425);
    if (nvalid > 0 && !tail) dst[c] = valid ? zc[c] : 0.0;
    bal[c] = __ballot_sync(0xffffffffu, tail);
  }
  const int n0 = __popc(bal[0]), n1 = n0 + __popc(bal[1]), n2 = n1 + __popc(bal[2]), n3 = n2 + __popc(bal[3]);
  const unsigned long long dsta = (unsigned long long)dst;
  for (int first = 0; first < n3; first += 32) {      // one pass unless more than 32 of the 4 x 32 draws are in the tails
    const int i = first + lane;
    const int c = (i < n0) ? 0 : (i < n1) ? 1 : (i < n2) ? 2 : 3;
    const int rank = i - ((c == 0) ? 0 : (c == 1) ? n0 : (c == 2) ? n1 : n2);
    const unsigned m = (c == 0) ? bal[0] : (c == 1) ? bal[1] : (c == 2) ? bal[2] : bal[3];
    const int src = (i < n3) ? mvn_nth_set_bit(m, rank) : 0;
    const double p0 = __shfl_sync(0xffffffffu, pr[0], src), p1 = __shfl_sync(0xffffffffu, pr[1], src);
    const double p2 = __shfl_sync(0xffffffffu, pr[2], src), p3 = __shfl_sync(0xffffffffu, pr[3], src);
    double* const dsrc = (double*)__shfl_sync(0xffffffffu, dsta, src);
    if (i < n3) {
      const double pc = (c == 0) ? p0 : (c == 1) ? p1 : (c == 2) ? p2 : p3;
      dsrc[c] = ubm_qnorm_tail(pc, pc - 0.5);
    }
  }
}
// the same for one row: draws i0 + 4 lane .. + 3 of one replicate go to row[4 lane .. 4 lane + 3]
template <bool TAPE>
__device__ __forceinline__ void draw_n4_warp(const RunParams& P, long long rep, unsigned long long i0, int nvalid,
                                             double* __restrict__ row, int lane, bool* bad) {
  if (TAPE) {
    if (nvalid > 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) row[4 * lane + c] = (c < nvalid) ? draw_n_at<true>(P, rep, i0 + c, bad) : 0.0;
    }
    return;
  }
  draw_n4_lanes(P, rep, i0, nvalid, row + 4 * lane, lane);
}

struct MvnSmem {
  double *packA, *packB, *packC, *MEAN;
  double *X1, *X2;                // [R][DPS]
  double *XI;                     // [2][R][DPS]  double-buffered chain-1 row: normals of this step -> proposal 1 (in place)
  double *V2;                     // [R][DPS]     chain-2 row: z -> e -> eta -> proposal 2 (in place)
  double *TG;                     // [2R][32]     per 4-column group partial sums of squares
  double *ZERO;                   // [DPS]
  double *LU;                     // [R][8] ring of log-uniforms produced ahead
  MvnSlots* S;
};

__host__ __device__ inline size_t mvn_smem_bytes(int packLen, int DPS, int R) {
  return (size_t)(3 * packLen + DPS + 5 * R * DPS + 2 * R * 32 + DPS + 8 * R) * sizeof(double) + sizeof(MvnSlots) + 16;
}

#ifdef UBM_MVN_TIMING
static __device__ long long g_mvn_tk[4];      // debug builds only: thread (0,0): cycles in K loops, k-steps, DMMAs, calls
#endif
// The K loop of one output tile jt for this consumer warp: acc[i] = rows of the i-th ACTIVE row tile times the packed
// factor, k-steps in increasing order.  MODE 0: a = r0[k]; 1: a = r0[k] - mean[k]; 2: a = r1[k] - r0[k].
// NT (compile time) row tiles, no predication: a DMMA that is predicated off still occupies its issue slot of the tensor
// pipe (measured: the first version ran every k-step at the cost of four DMMAs whatever the mask).
template <int MODE, bool GLOBAL, int NT>
__device__ __forceinline__ void mvn_kloop(const double* __restrict__ bp, int nk, int kq, const double* const (&p0)[4],
                                          const double* const (&p1)[4], const double* __restrict__ mean, double (&acc)[4][2]) {
  // fragments of k-step ks + 1 are requested before the DMMAs of k-step ks are issued (register double buffer)
  double b = GLOBAL ? __ldg(bp) : bp[0];
  double a[NT], a1[NT], mu = 0.0;
#pragma unroll
  for (int i = 0; i < NT; ++i) { a[i] = p0[i][kq]; if (MODE == 2) a1[i] = p1[i][kq]; }
  if (MODE == 1) mu = mean[kq];
  for (int ks = 0; ks < nk; ++ks) {
    const int kn = (ks + 1 < nk) ? ks + 1 : ks;        // the last trip re-reads its own fragments (never used)
    const double bn = GLOBAL ? __ldg(bp + 32 * kn) : bp[32 * kn];
    const int ko = 4 * kn + kq;
    double an[NT], a1n[NT], mun = 0.0;
#pragma unroll
    for (int i = 0; i < NT; ++i) { an[i] = p0[i][ko]; if (MODE == 2) a1n[i] = p1[i][ko]; }
    if (MODE == 1) mun = mean[ko];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      double x = a[i];
      if (MODE == 1) x = x - mu;
      if (MODE == 2) x = a1[i] - x;
      mvn_dmma(acc[i][0], acc[i][1], x, b);
    }
    b = bn; mu = mun;
#pragma unroll
    for (int i = 0; i < NT; ++i) { a[i] = an[i]; if (MODE == 2) a1[i] = a1n[i]; }
  }
}
// r0 / r1: per row tile q (chain q >> 1, slots 8 (q & 1) ..), this lane's row (row l / 4 of the tile; the ZERO row where the
// slot does not exist).  The active tiles of `tmask` are compacted: acc[i] belongs to tile qmap[i], i < popc(tmask).
template <int MODE, bool GLOBAL>
__device__ __forceinline__ void mvn_gemm(const double* __restrict__ pack, int d, int jt, int lane, unsigned tmask,
                                         const double* const (&r0)[4], const double* const (&r1)[4],
                                         const double* __restrict__ mean, double (&acc)[4][2], int (&qmap)[4]) {
  const int kq = lane & 3;
  const double* p0[4];
  const double* p1[4];
  unsigned rest = tmask;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[i][0] = 0.0; acc[i][1] = 0.0;
    const int q = (__ffs(rest) - 1) & 3;                     // i-th set bit (3 past the count: never used)
    rest &= rest - 1;
    qmap[i] = q;
    p0[i] = (q == 0) ? r0[0] : (q == 1) ? r0[1] : (q == 2) ? r0[2] : r0[3];
    p1[i] = (q == 0) ? r1[0] : (q == 1) ? r1[1] : (q == 2) ? r1[2] : r1[3];
  }
  const int nk = min(2 * jt + 2, (d + 3) >> 2);
  const double* bp = pack + (size_t)mvn_tile_base(jt) * 32 + lane;
#ifdef UBM_MVN_TIMING
  const long long tk0_ = clock64();
#endif
  switch (__popc(tmask)) {
    case 1: mvn_kloop<MODE, GLOBAL, 1>(bp, nk, kq, p0, p1, mean, acc); break;
    case 2: mvn_kloop<MODE, GLOBAL, 2>(bp, nk, kq, p0, p1, mean, acc); break;
    case 3: mvn_kloop<MODE, GLOBAL, 3>(bp, nk, kq, p0, p1, mean, acc); break;
    case 4: mvn_kloop<MODE, GLOBAL, 4>(bp, nk, kq, p0, p1, mean, acc); break;
    default: break;
  }
#ifdef UBM_MVN_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0) { g_mvn_tk[0] += clock64() - tk0_; g_mvn_tk[1] += nk; g_mvn_tk[2] += (long long)nk * __popc(tmask); g_mvn_tk[3] += 1; }
#endif
}

// per 4-column group sum of squares of a fragment pair, in the canonical order fma(v3,v3,fma(v2,v2,fma(v1,v1,fma(v0,v0,0)))):
// lanes with even l % 4 hold columns 0,1 of the group, the next lane columns 2,3.  Returns the group sum on odd lanes.
__device__ __forceinline__ double mvn_group_sq(double v0, double v1, int lane) {
  const double lo = ubm_fma(v1, v1, ubm_fma(v0, v0, 0.0));
  const double prev = __shfl_up_sync(0xffffffffu, lo, 1);
  return ubm_fma(v1, v1, ubm_fma(v0, v0, prev));
}

// NH: number of test-function blocks accumulated per coordinate (1: identity or square, 2: both)
template <bool TAPE, bool ZERO_MEAN, int NH>
__global__ void __launch_bounds__(MVN_T) mvn_driver_kernel(const MvnDevice M, const RunParams P, const int R) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, T = MVN_T, lane = tid & 31, warp = tid >> 5;
  const int d = M.d, DPS = M.DPS, nCG = M.nCG;
  MvnSmem sm;
  {
    double* p = reinterpret_cast<double*>(smem_raw);
    sm.packA = p; p += M.packLen; sm.packB = p; p += M.packLen; sm.packC = p; p += M.packLen;
    sm.MEAN = p; p += DPS;
    sm.X1 = p; p += R * DPS; sm.X2 = p; p += R * DPS; sm.XI = p; p += 2 * R * DPS; sm.V2 = p; p += R * DPS;
    sm.TG = p; p += 2 * R * 32;
    sm.ZERO = p; p += DPS;
    sm.LU = p; p += 8 * R;
    sm.S = reinterpret_cast<MvnSlots*>(p);
  }
  MvnSlots& S = *sm.S;
  for (int i = tid; i < M.packLen; i += T) { sm.packA[i] = M.packA[i]; sm.packB[i] = M.packB[i]; sm.packC[i] = M.packC[i]; }
  for (int i = tid; i < DPS; i += T) sm.MEAN[i] = M.mean_pi[i];
  for (int i = tid; i < 5 * R * DPS + 2 * R * 32 + DPS + 8 * R; i += T) sm.X1[i] = 0.0;
  if (tid < MVN_MAXR) {
    S.rep[tid] = -1; S.phase[tid] = SL_IDLE; S.time[tid] = 0; S.tau[tid] = 0; S.met[tid] = 0;
    S.iu[tid] = 0; S.in[tid] = 0; S.pu[tid] = 0; S.bad[tid] = 0; S.ident[tid] = 0;
  }
  if (tid == 0) S.exhausted = 0;
  __syncthreads();

  const int dimh = NH * d;
  const double denom = (double)(P.m - P.k + 1);
  const unsigned FULL = 0xffffffffu;
  const int nbins = (P.mode == MODE_ESTIMATOR && P.nbreaks >= 2) ? P.nbreaks - 1 : 0;
  double bin_b0 = 0.0, bin_iw = 0.0;
  if (nbins) { bin_b0 = P.breaks[0]; bin_iw = (double)nbins / (P.breaks[nbins] - P.breaks[0]); }
  int cur = 0;
  // estimator sums of this warp's slot (slot s = warp): coordinate j = lane + 32 i, test-function block hh
  double am[NH][4], ac[NH][4];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh)
#pragma unroll
    for (int i = 0; i < 4; ++i) { am[hh][i] = 0.0; ac[hh][i] = 0.0; }
  // this lane's rows of the four static row tiles (chain q >> 1, slots 8 (q & 1) + lane / 4), per input array
  const int myrow[2] = {lane >> 2, 8 + (lane >> 2)};
  const int kq = lane & 3;

#ifdef UBM_MVN_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tsteps = 0;
#define TMARK(i) do { long long now_ = clock64(); tacc[i] += now_ - tprev; tprev = now_; } while (0)
#else
#define TMARK(i) do { } while (0)
#endif
  for (;;) {
#ifdef UBM_MVN_TIMING
    long long tprev = clock64(); tsteps++;
#endif
    double* XIc = sm.XI + cur * R * DPS;          // this step's chain-1 rows (normals -> proposals)
    double* XIn = sm.XI + (cur ^ 1) * R * DPS;    // filled by the producers for the next step ...
    const double* XIp = XIn;                      // ... and, until barrier 1, still holding the proposals of the step just computed
    // ================================================================ per-slot section (warp s owns slot s)
    // finishes the step just computed (densities -> accept/reject -> state/estimator update) and runs the slot state
    // machine (finish / write-out / fetch the next replicate).  No random-number work: uniforms come from the ring.
    int my_active = 0;
    if (warp < R) {
      const int s = warp;
      int ph = S.phase[s];
      long long rep = S.rep[s];
      long long time = S.time[s];
      int met = S.met[s];
      if (ph >= SL_INIT) {
        // ---- log-densities of the proposals (src/mvnorm.cpp:81-84) and the MH accept step
        const double ss1 = warp_butterfly(sm.TG[(2 * s) * 32 + lane]);
        const double ss2 = warp_butterfly(sm.TG[(2 * s + 1) * 32 + lane]);
        const double pp1 = -0.5 * ss1 + M.cst;
        const int ident = S.ident[s];
        const double pp2 = (ph == SL_COUPLED && ident) ? pp1 : (-0.5 * ss2 + M.cst);
        int a1 = 0, a2 = 0, add_mcmc = 0, add_corr = 0;
        long long wgt_i = 0;
        if (ph == SL_INIT) {
          a1 = 1; a2 = 1;
          if (lane == 0) { S.pdf1[s] = pp1; S.pdf2[s] = pp2; }
        } else {
          const double logu = sm.LU[s * 8 + (int)(S.iu[s] & 7ull)];      // log(runif(1)), R/get_mh_kernels.R:34,57
          const double pdf1 = S.pdf1[s], pdf2 = S.pdf2[s];
          if (ph == SL_SINGLE) {
            a1 = (logu < pp1 - pdf1) ? 1 : 0;
          } else {
            if (isfinite(pp1)) a1 = (logu < pp1 - pdf1) ? 1 : 0;
            if (isfinite(pp2)) a2 = (logu < pp2 - pdf2) ? 1 : 0;
            if (ident && a1 && a2) met = 1;
          }
          time += 1;
          __syncwarp();
          if (lane == 0) {
            S.iu[s] += 1;
            if (a1) S.pdf1[s] = pp1;
            if (a2) S.pdf2[s] = pp2;
            S.time[s] = time;
            if (met && !S.met[s]) { S.met[s] = 1; S.tau[s] = time; }
          }
          // what the estimator adds after this step (R/unbiasedestimator.R:66, :70-72, :80, :90, :93-95)
          add_mcmc = (time <= P.lag) ? (time >= P.k) : (P.k <= time && time <= P.m);
          add_corr = ((time >= P.k + P.lag) && (ph == SL_COUPLED || time == P.lag)) ? 1 : 0;
          if (add_corr && P.mode == MODE_ESTIMATOR) wgt_i = corr_weight(time, P.k, P.m, P.lag);
        }
        const double wgt = (double)wgt_i;
        // ---- state update, estimator sums, trajectories
        const bool coupled_step = (ph == SL_COUPLED);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = lane + 32 * i;
          if (j < d) {
            const int o = s * DPS + j;
            if (a1) sm.X1[o] = XIp[o];
            if (a2) sm.X2[o] = (coupled_step && ident) ? XIp[o] : sm.V2[o];
            const double x1 = sm.X1[o];
            const double x2 = sm.X2[o];
            if (P.mode == MODE_ESTIMATOR) {
#pragma unroll
              for (int hh = 0; hh < NH; ++hh) {
                const bool sq = (P.h_kind == UBM_H_SQUARE) || hh == 1;
                const double h1 = sq ? x1 * x1 : x1;
                if (ph == SL_INIT) {
                  am[hh][i] = (P.k == 0) ? h1 : 0.0;                        // R/unbiasedestimator.R:55-59
                  ac[hh][i] = 0.0;
                } else {
                  if (add_mcmc) am[hh][i] = am[hh][i] + h1;
                  if (add_corr) {
                    const double h2 = sq ? x2 * x2 : x2;
                    ac[hh][i] = ac[hh][i] + wgt * (h1 - h2);
                  }
                }
              }
              if (nbins && j == P.component) {      // estimator_bin_ numerators on the fly (src/estimator_bin.cpp:15-38)
                long long* bn = P.bin_num + rep * nbins;
                if ((ph == SL_INIT) ? (P.k == 0) : (add_mcmc != 0)) {
                  const int b = bin_of(x1, P.breaks, P.nbreaks, bin_b0, bin_iw);
                  if (b >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(bn + b), 1ull);
                }
                if (ph != SL_INIT && add_corr) {
                  const int b1 = bin_of(x1, P.breaks, P.nbreaks, bin_b0, bin_iw);
                  const int b2 = bin_of(x2, P.breaks, P.nbreaks, bin_b0, bin_iw);
                  if (b1 >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(bn + b1), (unsigned long long)wgt_i);
                  if (b2 >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(bn + b2), (unsigned long long)(-wgt_i));
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
        }
        __syncwarp();
        // ---- slot state machine
        bool finished;
        // replay: a tape is exhausted when the step just finished CONSUMED a draw past its end (the producers only
        // prefetch, and reading ahead of the consumer is not an error)
        bool tape_bad = false;
        if (TAPE) {
          const long long n_used = (long long)S.in[s] + ((ph == SL_INIT) ? 2 * d : d);
          const long long len_n = P.tape_n ? P.len_n : 0, len_u = P.tape_u ? P.len_u : 0;
          tape_bad = n_used > len_n || (long long)S.iu[s] > len_u;
          if (tape_bad && lane == 0) S.bad[s] = 1;
        }
        const bool capped = (P.max_it > 0 && time >= P.max_it) || tape_bad;
        if (P.mode == MODE_MEETING) finished = met || (time >= P.lag && capped);
        else {
          const long long tt = met ? S.tau[s] : 0;
          const long long horizon = (tt > P.m) ? tt : P.m;
          finished = (met && time >= horizon) || (time >= P.lag && capped);
        }
        if (ph == SL_INIT) finished = false;
        if (lane == 0) S.in[s] += (ph == SL_INIT) ? 2 * d : d;          // normals consumed by the step just finished
        if (P.ho_int && P.mode == MODE_ESTIMATOR && met && !finished && ph != SL_INIT) {
          // two-pass run: the replicate has met and still has single-chain steps to its horizon — it continues in the
          // single-chain kernel (ubm_mvn_single.cuh) from this record (kept in its own output rows, see RunParams), and the
          // slot is free for the next coupled pair
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = lane + 32 * i;
            if (j < d) {
              P.uest[rep * dimh + j] = sm.X1[s * DPS + j];
#pragma unroll
              for (int hh = 0; hh < NH; ++hh) {
                P.mcmc[rep * dimh + hh * d + j] = am[hh][i];
                P.corr[rep * dimh + hh * d + j] = ac[hh][i];
              }
            }
          }
          if (lane == 0) {
            P.cost[rep] = S.pdf1[s];
            P.iter[rep] = time; P.tau[rep] = (double)S.tau[s];
            P.ho_int[rep * 2 + 0] = (long long)S.iu[s]; P.ho_int[rep * 2 + 1] = (long long)S.in[s];
            __threadfence();
            const unsigned long long q = atomicAdd(P.ho_count, 1ull);
            P.ho_list[q] = (int)rep;
          }
          ph = SL_IDLE;
        } else if (finished) {
          if (lane == 0) {
            const long long tt = S.tau[s];
            double cost = dinf();
            if (met) {
              if (P.mode == MODE_MEETING) cost = (double)(P.lag + 2 * (tt - P.lag));
              else cost = (double)(P.lag + 2 * (tt - P.lag) + ((time - tt) > 0 ? (time - tt) : 0));
            }
            if (P.tau) P.tau[rep] = met ? (double)tt : dinf();
            if (P.cost) P.cost[rep] = cost;
            if (P.iter) P.iter[rep] = time;
            if (tape_bad) atomicExch(P.status, UBM_ERR_TAPE);
          }
          if (P.mode == MODE_ESTIMATOR) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = lane + 32 * i;
              if (j < d) {
#pragma unroll
                for (int hh = 0; hh < NH; ++hh) {
                  const int e = hh * d + j;
                  const double mc = am[hh][i], co = ac[hh][i];
                  const double ue = mc + co;
                  P.uest[rep * dimh + e] = ue / denom;
                  if (P.mcmc) P.mcmc[rep * dimh + e] = mc / denom;
                  if (P.corr) P.corr[rep * dimh + e] = co / denom;
                }
              }
            }
          }
          ph = SL_IDLE;
        } else if (time < P.lag) ph = SL_SINGLE;
        else ph = met ? SL_SINGLE : SL_COUPLED;
      } else if (ph == SL_WAIT) {
        ph = SL_INIT;       // the producers generated the 2d rinit normals during the step that just ended
      }
      if (ph == SL_IDLE) {
        long long r = -1;
        if (lane == 0 && !S.exhausted) {
          const unsigned long long got = atomicAdd(P.work_counter, 1ull);
          if (got < (unsigned long long)P.nrep) r = (long long)got;
          else S.exhausted = 1;
        }
        r = __shfl_sync(FULL, r, 0);
        if (r >= 0) {
          ph = SL_WAIT;
          if (lane == 0) {
            S.rep[s] = r; S.time[s] = 0; S.tau[s] = 0; S.met[s] = 0; S.iu[s] = 0; S.in[s] = 0; S.pu[s] = 0; S.bad[s] = 0;
          }
        } else if (lane == 0) S.rep[s] = -1;
      }
      __syncwarp();
      if (lane == 0) { S.phase[s] = ph; S.ident[s] = 0; }
      if (ph != SL_IDLE) my_active = 1;
    }
    const int any_active = __syncthreads_or(my_active);      // CTA barrier 1: phases of this step are published
    if (!any_active) break;
    TMARK(0);

    if (warp < MVN_CW) {
      // ================================================================ CONSUMERS (warps 0-7)
      const int cw = warp;
      // slot masks of this step (every consumer warp computes them: no table, no barrier)
      unsigned mA, mS, mI;
      {
        const int ph = (lane < R) ? S.phase[lane] : SL_IDLE;
        mA = __ballot_sync(FULL, ph == SL_COUPLED); mS = __ballot_sync(FULL, ph == SL_SINGLE); mI = __ballot_sync(FULL, ph == SL_INIT);
      }
      const int sl0 = myrow[0], sl1 = myrow[1];                 // the two slots this lane's fragment rows belong to
      const bool ex0 = sl0 < R, ex1 = sl1 < R;
      double acc[4][2];
      int qmap[4];
      unsigned mD = 0;       // coupled slots whose proposals differ (chain-2 rows of phases B and C)
      // ---------------------------------------------------------------- phase A: z = -((x2 - x1)^T Cinv_prop) -> V2, |z|^2
      if (mA) {
        const unsigned tm = ((mA & 0xffu) ? 1u : 0u) | ((mA >> 8) ? 2u : 0u);
        const double* const r0[4] = {ex0 ? sm.X1 + sl0 * DPS : sm.ZERO, ex1 ? sm.X1 + sl1 * DPS : sm.ZERO, sm.ZERO, sm.ZERO};
        const double* const r1[4] = {ex0 ? sm.X2 + sl0 * DPS : sm.ZERO, ex1 ? sm.X2 + sl1 * DPS : sm.ZERO, sm.ZERO, sm.ZERO};
        for (int jb = 0; jb < MVN_JMAX; ++jb) {
          const int jt = M.jobs[cw][jb];
          if (jt == 255) continue;
          mvn_gemm<2, false>(sm.packA, d, jt, lane, tm, r0, r1, nullptr, acc, qmap);
          const int n0 = 8 * jt + 2 * kq, g = 2 * jt + (kq >> 1);
          const int ntl = __popc(tm);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (i < ntl) {
              const int q = qmap[i];
              const int s = q ? sl1 : sl0;
              const double z0 = -acc[i][0], z1 = -acc[i][1];        // :198
              const double t = mvn_group_sq(z0, z1, lane);
              if ((mA >> s) & 1u) {
                *reinterpret_cast<double2*>(sm.V2 + s * DPS + n0) = make_double2(z0, z1);
                if (kq & 1) sm.TG[(2 * s) * 32 + g] = t;
              }
            }
          }
        }
        consumer_bar();
        // -------------------------------------------------------------- coupling algebra, one warp per coupled slot
        for (int s = cw; s < R; s += MVN_CW) {
          if (!((mA >> s) & 1u)) continue;
          const double nz = sqrt(warp_butterfly(sm.TG[(2 * s) * 32 + lane]));   // :199
          double z[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
          double t = 0.0;
          if (lane < nCG) {
#pragma unroll
            for (int c = 0; c < 4; ++c) z[c] = XIc[s * DPS + 4 * lane + c];      // xi, drawn ahead (:193-195)
            if (!(nz < 1e-15)) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                e[c] = sm.V2[s * DPS + 4 * lane + c] / nz;                       // :212
                t = ubm_fma(e[c], z[c], t);                                      // :216
              }
            }
          }
          const double edx = warp_butterfly(t);
          int ident = 0;
          if (nz < 1e-15) ident = 1;                                             // :200-210, no uniform consumed
          else {
            const double lut = sm.LU[s * 8 + (int)(S.iu[s] & 7ull)];             // log(utilde), :213-215
            const double a = edx + nz;
            ident = (lut < (-0.5 * a * a + 0.5 * edx * edx)) ? 1 : 0;            // :217
          }
          __syncwarp();
          if (lane == 0) { S.ident[s] = ident; if (!(nz < 1e-15)) S.iu[s] += 1; }
          if (!ident && lane < nCG) {
            const double two_edx = 2. * edx;
#pragma unroll
            for (int c = 0; c < 4; ++c) sm.V2[s * DPS + 4 * lane + c] = z[c] - two_edx * e[c];   // :221
          }
        }
        consumer_bar();
        mD = __ballot_sync(FULL, lane < R && ((mA >> lane) & 1u) && !S.ident[lane < R ? lane : 0]);
      }
      TMARK(1);
      // ---------------------------------------------------------------- phase B: proposals = x + in^T C_prop (in place)
      // and, for freshly fetched slots, x0 = init_mean + in^T init_chol (matrix read from global memory / L2).
      // Outputs overwrite their inputs: both K loops of every warp first, then a barrier, then the writes.
      const double* const rc[4] = {ex0 ? XIc + sl0 * DPS : sm.ZERO, ex1 ? XIc + sl1 * DPS : sm.ZERO,
                                   ex0 ? sm.V2 + sl0 * DPS : sm.ZERO, ex1 ? sm.V2 + sl1 * DPS : sm.ZERO};
      for (int pass = 0; pass < 2; ++pass) {
        const unsigned m1 = pass ? mI : (mA | mS), m2 = pass ? mI : mD;      // chain-1 rows, chain-2 rows
        if (!(m1 | m2)) continue;                                            // uniform over the consumer group
        const unsigned tm = ((m1 & 0xffu) ? 1u : 0u) | ((m1 >> 8) ? 2u : 0u) | ((m2 & 0xffu) ? 4u : 0u) | ((m2 >> 8) ? 8u : 0u);
        double accB[4][2];
        const int jt0 = M.jobs[cw][0], jt1 = M.jobs[cw][1];
        // two call sites each so that the matrix loads are LDS (shared) resp. LDG (global), not generic loads
        if (jt0 != 255) {
          if (pass == 0) mvn_gemm<0, false>(sm.packB, d, jt0, lane, tm, rc, rc, nullptr, acc, qmap);
          else mvn_gemm<0, true>(M.packI, d, jt0, lane, tm, rc, rc, nullptr, acc, qmap);
        }
        if (jt1 != 255) {
          if (pass == 0) mvn_gemm<0, false>(sm.packB, d, jt1, lane, tm, rc, rc, nullptr, accB, qmap);
          else mvn_gemm<0, true>(M.packI, d, jt1, lane, tm, rc, rc, nullptr, accB, qmap);
        }
        consumer_bar();   // every read of the input rows is done: overwrite them with the outputs
        const int ntl = __popc(tm);
#pragma unroll
        for (int jb = 0; jb < MVN_JMAX; ++jb) {
          const int jt = jb ? jt1 : jt0;
          if (jt != 255) {
            const int n0 = 8 * jt + 2 * kq;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < ntl) {
                const int q = qmap[i];
                const int s = (q & 1) ? sl1 : sl0, which = q >> 1;
                if (((which ? m2 : m1) >> s) & 1u) {
                  const double* base = pass ? (M.init_mean + n0) : ((which ? sm.X2 : sm.X1) + s * DPS + n0);
                  const double v0 = jb ? accB[i][0] : acc[i][0], v1 = jb ? accB[i][1] : acc[i][1];
                  double2 o;
                  o.x = (n0 < d) ? (base[0] + v0) : 0.0;          // :40-44, :224-225
                  o.y = (n0 + 1 < d) ? (base[1] + v1) : 0.0;
                  *reinterpret_cast<double2*>((which ? sm.V2 : XIc) + s * DPS + n0) = o;
                }
              }
            }
          }
        }
        if (pass == 0 && mI) consumer_bar();   // the init pass must not start before every write of this pass has landed in ITS rows' neighbours' reads
      }
      consumer_bar();
      TMARK(2);
      // ---------------------------------------------------------------- phase C: w = Cinv_pi^T (prop - mean), |w|^2
      {
        const unsigned m1 = mA | mS | mI, m2 = mD | mI;
        const unsigned tm = ((m1 & 0xffu) ? 1u : 0u) | ((m1 >> 8) ? 2u : 0u) | ((m2 & 0xffu) ? 4u : 0u) | ((m2 >> 8) ? 8u : 0u);
        for (int jb = 0; jb < MVN_JMAX; ++jb) {
          const int jt = M.jobs[cw][jb];
          if (jt == 255) continue;
          TMARK(5);
          mvn_gemm<ZERO_MEAN ? 0 : 1, false>(sm.packC, d, jt, lane, tm, rc, rc, sm.MEAN, acc, qmap);
          TMARK(6);
          const int g = 2 * jt + (kq >> 1);
          const int ntl = __popc(tm);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < ntl) {
              const int q = qmap[i];
              const int s = (q & 1) ? sl1 : sl0, which = q >> 1;
              const double t = mvn_group_sq(acc[i][0], acc[i][1], lane);
              if ((((which ? m2 : m1) >> s) & 1u) && (kq & 1)) sm.TG[(2 * s + which) * 32 + g] = t;
            }
          }
        }
      }
      TMARK(3);
    } else {
      // ================================================================ PRODUCERS (warps 8-15): draws of the NEXT step
      const int pw = warp - MVN_CW;
      for (int s = pw; s < R; s += MVN_PW) {
        const int ph = S.phase[s];
        if (ph == SL_IDLE) continue;
        const long long rep = S.rep[s];
        const unsigned long long base = (ph == SL_WAIT) ? 0ull : ((ph == SL_INIT) ? (unsigned long long)(2 * d) : S.in[s] + d);
        bool past_end = false;    // prefetch past the end of a tape yields a dummy value; only consumption is an error
        const int nvalid = (lane < nCG) ? d - 4 * lane : 0;
        draw_n4_warp<TAPE>(P, rep, base + 4 * lane, nvalid, XIn + s * DPS, lane, &past_end);
        if (ph == SL_WAIT)        // rinit of chain 2 (draw order: chain 1's d normals, then chain 2's)
          draw_n4_warp<TAPE>(P, rep, base + d + 4 * lane, nvalid, sm.V2 + s * DPS, lane, &past_end);
      }
      // log-uniform rings of this warp's slots (8 lanes per slot): keep them at least 4 draws ahead of the consumers
      // (they use at most 2 per step)
      {
        const int s = pw + (lane >> 3) * MVN_PW, l8 = lane & 7;
        const bool on = s < R && S.phase[s] != SL_IDLE;
        unsigned long long pu = 0, target = 0;
        bool past_end = false;
        if (on) {
          pu = S.pu[s]; target = S.iu[s] + 6;
          const unsigned long long idx = pu + l8;
          if (idx < target) sm.LU[s * 8 + (int)(idx & 7ull)] = ubm_log(draw_u_at<TAPE>(P, S.rep[s], idx, &past_end));
        }
        __syncwarp();
        if (on && l8 == 0 && target > pu) S.pu[s] = target;
      }
    }
    __syncthreads();      // CTA barrier 2: step computed, draws of the next step in place
    TMARK(4);
    cur ^= 1;
  }
#ifdef UBM_MVN_TIMING
  if (tid == 0 && blockIdx.x == 0)
    printf("mvn timing (cycles/step, block 0, thread 0): steps %lld slot %.0f A+coupling %.0f B %.0f C %.0f wait-for-producers %.0f | K loops of warp 0: %.0f cycles/step, %.1f k-steps/step, %.1f DMMA/step, %.1f calls/step\n", tsteps,
           (double)tacc[0] / tsteps, (double)tacc[1] / tsteps, (double)tacc[2] / tsteps, (double)tacc[3] / tsteps,
           (double)tacc[4] / tsteps, (double)g_mvn_tk[0] / tsteps, (double)g_mvn_tk[1] / tsteps, (double)g_mvn_tk[2] / tsteps, (double)g_mvn_tk[3] / tsteps);
  if (tid == 0 && blockIdx.x == 0) printf("   phase C split: before gemm %.0f, gemm call %.0f, after (epilogue + loop) %.0f\n", (double)tacc[5] / tsteps, (double)tacc[6] / tsteps, (double)tacc[3] / tsteps);
#endif
}

template <class Buf>
inline int mvn_launch(const MvnDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st, Buf& scratch,
                      int64_t* launches, std::string* err) {
  (void)scratch;
  const size_t budget = 227 * 1024;
  const int nh = (P.mode == MODE_ESTIMATOR && P.h_kind == UBM_H_BOTH) ? 2 : 1;
  int R = MVN_MAXR;
  while (R > 1 && mvn_smem_bytes(M.packLen, M.DPS, R) > budget) --R;
  size_t smem = mvn_smem_bytes(M.packLen, M.DPS, R);
  if (smem > budget) { *err = "mvn: d too large for shared memory"; return UBM_ERR_UNSUPPORTED; }
  int64_t need = (P.nrep + R - 1) / R;
  int grid = (int)std::min<int64_t>(nsm, need);
  void (*kern)(const MvnDevice, const RunParams, const int);
  const int variant = (tape ? 4 : 0) | (M.zero_mean ? 2 : 0) | (nh == 2 ? 1 : 0);
  switch (variant) {
    case 0: kern = mvn_driver_kernel<false, false, 1>; break;
    case 1: kern = mvn_driver_kernel<false, false, 2>; break;
    case 2: kern = mvn_driver_kernel<false, true, 1>; break;
    case 3: kern = mvn_driver_kernel<false, true, 2>; break;
    case 4: kern = mvn_driver_kernel<true, false, 1>; break;
    case 5: kern = mvn_driver_kernel<true, false, 2>; break;
    case 6: kern = mvn_driver_kernel<true, true, 1>; break;
    default: kern = mvn_driver_kernel<true, true, 2>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = std::string("mvn: smem attribute: ") + cudaGetErrorString(e); return UBM_ERR_CUDA; }
  kern<<<grid, MVN_T, smem, st>>>(M, P, R);
  *launches += 1;
  return UBM_OK;
}

}  // namespace ubm
