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
// to the oracle's plain loops (structural zeros only ever add exact zeros).  Triangular work is balanced by
// pairing column group g with nCG-1-g.  Slots are refilled from a global counter as soon as a replicate
// finishes (work stealing).  Reductions over a d-vector use the canonical order "fma chain inside each
// 4-column group, xor-butterfly over the <= 32 groups" (oracle: sum_g4_*).
//
// Thread roles (512 threads, warp-specialised): warps 0-7 (two per SM sub-partition, so that one fills the other's DFMA
// issue gaps) are CONSUMERS: they run the GEMM phases and the coupling algebra, synchronising among themselves on a
// named barrier; warps 8-15 are PRODUCERS: while step t is computed they generate every draw of step t+1 (the d normals
// per slot through Philox + AS241, the rare tail branch compacted over the warp, and the log-uniforms, into a small
// ring), so no random-number work sits on the critical path.  All 16 warps share the short per-slot section at the
// step boundary (accept/reject, state and estimator update, slot state machine), one warp per slot.  Tiles are 2 rows x
// 4 columns per thread over balanced half-units (see mvn_phase_compute); phases with more rows than one batch run in
// batches.  Two CTA-wide barriers per step.
//
// Shared-memory layout notes (ncu and tools/lds_probe.cu, profiles/r01_mvn.md: the first version had 12.5e9 bank
// conflicts; an LDS.128 is served per quarter-warp, one wavefront when its 16-byte words fall into distinct banks):
//   - packed group g starts at 8 g (g+1) + 4 g (+2 for the long groups) doubles, so that the two streams of a
//     quarter-warp (a short group and its long partner) read different banks;
//   - vectors use a row stride of DPS doubles with DPS/2 == 2 (mod 8) words: four consecutive rows plus the same rows an
//     odd number of words further on (the partner stream) cover the eight bank groups exactly once;
//   - the canonical cut of the long columns is `half` == 2 (mod 4) rows from the end (mvn_half);
//   - GEMM units are numbered row-group fastest, so a warp touches few column groups (matrix bytes dominate).
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "ubm_common.cuh"

namespace ubm {

struct MvnDevice {
  int d, DP, DPS, nCG, nPairs, zero_mean;
  int packLen;                 // doubles per packed matrix
  double cst;                  // src/mvnorm.cpp:74-75
  const double *mean_pi, *init_mean, *packA, *packB, *packC, *packI;   // device
  size_t offs[6];              // host-side: offsets inside the blob
};

// column group g holds rows k = 0 .. 4g+3 (rows >= d and entries below the diagonal are zero), 4 doubles per row.
// Pads: group g starts 4 g (+2 for the long groups g >= nPairs) doubles late, so that the 16-byte words read by one
// quarter-warp (a short group gA = pr and its partner gB = nCG-1-pr, at the same loop step) fall into different banks.
__host__ __device__ __forceinline__ int mvn_group_off(int g, int nPairs) { return 8 * g * (g + 1) + 4 * g + (g >= nPairs ? 2 : 0); }
__host__ __device__ __forceinline__ int mvn_pack_len(int nCG, int nPairs) { return mvn_group_off(nCG - 1, nPairs) + 16 * nCG + 16; }
// The canonical cut of the long columns (oracle: vec_times_upper): K_g > half is summed as [0, K_g - half) + [K_g - half, K_g).
// half == 2 (mod 4) keeps the two streams of a column group, `half` rows apart, in different banks.
__host__ __device__ __forceinline__ int mvn_half(int nCG) { return 4 * ((nCG + 1) / 2) + 2; }

inline void mvn_pack_matrix(const double* Ucm, int d, int nCG, std::vector<double>* blob) {
  size_t base = blob->size();
  const int nPairs = (nCG + 1) / 2;
  blob->resize(base + mvn_pack_len(nCG, nPairs), 0.0);
  for (int g = 0; g < nCG; ++g) {
    double* p = blob->data() + base + mvn_group_off(g, nPairs);
    for (int k = 0; k < 4 * g + 4; ++k)
      for (int c = 0; c < 4; ++c) {
        int j = 4 * g + c;
        p[4 * k + c] = (j < d && k <= j) ? Ucm[(size_t)j * d + k] : 0.0;
      }
  }
}

inline void mvn_pack_host(const ubm_model_mvn* s, MvnDevice* M, std::vector<double>* blob) {
  const int d = s->d;
  M->d = d; M->nCG = (d + 3) / 4; M->DP = 4 * M->nCG; M->nPairs = (M->nCG + 1) / 2;
  // row stride (in 16-byte words) == 2 (mod 8): four consecutive rows plus the same rows an odd number of words further
  // on (the partner stream) cover the eight bank groups exactly once.  Reads of the GEMM prefetch may run 16 bytes past a
  // row: that is the next row or the next array, never used.
  M->DPS = M->DP + 2 * ((((2 - 2 * M->nCG) % 8) + 8) % 8);
  M->zero_mean = 1;
  for (int j = 0; j < d; ++j) if (s->mean_pi[j] != 0.0) M->zero_mean = 0;
  double halflogdet = 0.0;
  for (int j = 0; j < d; ++j) halflogdet = halflogdet + ubm_log(s->chol_inv_pi[(size_t)j * d + j]);
  M->cst = -(-halflogdet) - (d * 0.9189385);     // truncated constant kept verbatim (src/mvnorm.cpp:75)
  M->packLen = mvn_pack_len(M->nCG, M->nPairs);
  blob->clear();
  M->offs[0] = blob->size(); for (int j = 0; j < M->DPS; ++j) blob->push_back(j < d ? s->mean_pi[j] : 0.0);
  M->offs[1] = blob->size(); for (int j = 0; j < M->DPS; ++j) blob->push_back(j < d ? s->init_mean[j] : 0.0);
  M->offs[2] = blob->size(); mvn_pack_matrix(s->chol_inv_prop, d, M->nCG, blob);
  M->offs[3] = blob->size(); mvn_pack_matrix(s->chol_prop, d, M->nCG, blob);
  M->offs[4] = blob->size(); mvn_pack_matrix(s->chol_inv_pi, d, M->nCG, blob);
  M->offs[5] = blob->size(); mvn_pack_matrix(s->init_chol, d, M->nCG, blob);
}
inline void mvn_bind_device(MvnDevice* M, const double* dev) {
  M->mean_pi = dev + M->offs[0]; M->init_mean = dev + M->offs[1];
  M->packA = dev + M->offs[2]; M->packB = dev + M->offs[3]; M->packC = dev + M->offs[4]; M->packI = dev + M->offs[5];
}

enum { SL_IDLE = 0, SL_WAIT = 1, SL_INIT = 2, SL_SINGLE = 3, SL_COUPLED = 4 };
#define MVN_MAXR 16
#ifndef MVN_RT
#define MVN_RT 2       // rows per GEMM register tile (x 4 columns)
#endif
#ifndef MVN_CW
#define MVN_CW 8       // consumer warps: warps 0-7 (two per SM sub-partition, so that one fills the other's DFMA issue gaps)
#endif
#ifndef MVN_PW
#define MVN_PW 8       // producer warps: warps 8-15 produce the draws of the next step
#endif
#define MVN_GT (32 * MVN_CW)               // consumer (GEMM) threads
#define MVN_T (32 * (MVN_CW + MVN_PW))     // threads per CTA

struct MvnSlots {   // per-slot scalars, in shared memory
  long long rep[MVN_MAXR], time[MVN_MAXR], tau[MVN_MAXR];
  unsigned long long iu[MVN_MAXR], in[MVN_MAXR], pu[MVN_MAXR];
  double pdf1[MVN_MAXR], pdf2[MVN_MAXR];
  int phase[MVN_MAXR], met[MVN_MAXR], ident[MVN_MAXR], bad[MVN_MAXR];
  int rowsA[MVN_MAXR], rowsB[2 * MVN_MAXR], rowsI[2 * MVN_MAXR], rowsC[2 * MVN_MAXR];
  int nA, nB, nI, nC;
  int exhausted;
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MVN_GT) : "memory"); }

// Operands of one two-row step of a half-unit: 8 matrix entries (rows k, k+1 of a 4-column group) and the RT row pairs.
template <int RT, int MODE>
struct MvnOperands {
  double2 m0, m1, m2, m3, v[RT], w[RT], mu;
  __device__ __forceinline__ void load(const double* __restrict__ mp, const double* const* in, const double* const* in2,
                                       const double* __restrict__ aux, int k) {
    m0 = *reinterpret_cast<const double2*>(mp); m1 = *reinterpret_cast<const double2*>(mp + 2);
    m2 = *reinterpret_cast<const double2*>(mp + 4); m3 = *reinterpret_cast<const double2*>(mp + 6);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      v[r] = *reinterpret_cast<const double2*>(in[r] + k);
      if (MODE == 2) w[r] = *reinterpret_cast<const double2*>(in2[r] + k);
    }
    if (MODE == 1) mu = *reinterpret_cast<const double2*>(aux + k);
  }
  __device__ __forceinline__ void fma_into(double (&acc)[RT][4]) const {
    double2 x[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      x[r] = v[r];
      if (MODE == 1) { x[r].x = x[r].x - mu.x; x[r].y = x[r].y - mu.y; }
      if (MODE == 2) { x[r].x = w[r].x - x[r].x; x[r].y = w[r].y - x[r].y; }
    }
    // row k for all accumulators, then row k+1: dependent fma pairs are 4 RT instructions apart
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      acc[r][0] = ubm_fma(x[r].x, m0.x, acc[r][0]);
      acc[r][1] = ubm_fma(x[r].x, m0.y, acc[r][1]);
      acc[r][2] = ubm_fma(x[r].x, m1.x, acc[r][2]);
      acc[r][3] = ubm_fma(x[r].x, m1.y, acc[r][3]);
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      acc[r][0] = ubm_fma(x[r].y, m2.x, acc[r][0]);
      acc[r][1] = ubm_fma(x[r].y, m2.y, acc[r][1]);
      acc[r][2] = ubm_fma(x[r].y, m3.x, acc[r][2]);
      acc[r][3] = ubm_fma(x[r].y, m3.y, acc[r][3]);
    }
  }
};

// One GEMM half-unit: RT input rows (smem, 16-byte aligned) against two consecutive row segments of a packed matrix,
// executed as ONE loop so that the lanes of a warp — whose segments differ — run the same instruction stream:
//   segment 1: n1 (even) two-row steps of column group g1 from k = 0   -> first[r][c]
//   segment 2: n2 two-row steps of column group g2 from k = k2 (even)  -> acc[r][c]
// each sum sequential in k.  The loop is unrolled by two steps with two operand sets (ping-pong): the operands of the
// next step are requested while this one is computed, without register copies (reads run at most 32 bytes past a
// group / row: padding, the next row or the next array, never used).
// MODE 0: plain input; 1: input minus `aux` vector (the target mean); 2: input = in2_r[k] - in_r[k] (x2 - x1).
template <int RT, int MODE>
__device__ __forceinline__ void mvn_unit(const double* __restrict__ pack, int nPairs, int g1, int n1, int g2, int k2, int n2,
                                         const double* const* in, const double* const* in2,
                                         const double* __restrict__ aux, double (&first)[RT][4], double (&acc)[RT][4]) {
  const int n = n1 + n2;
  const double* const mp2 = pack + mvn_group_off(g2, nPairs) + 4 * k2;
  const double* mp = (n1 > 0) ? pack + mvn_group_off(g1, nPairs) : mp2;
  int k = (n1 > 0) ? 0 : k2;
  MvnOperands<RT, MODE> A, B;
  A.load(mp, in, in2, aux, k);
  int it = 0;
  for (; it + 1 < n; it += 2) {
    if (it == n1 && it > 0) {      // segment 1 is complete (a few lanes, once per unit)
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { first[r][c] = acc[r][c]; acc[r][c] = 0.0; }
    }
    B.load(mp + 8, in, in2, aux, k + 2);                 // step it + 1 (odd: never the first of segment 2)
    A.fma_into(acc);
    const bool jump = (it + 2 == n1);
    mp = jump ? mp2 : mp + 16;
    k = jump ? k2 : k + 4;
    A.load(mp, in, in2, aux, k);                         // step it + 2
    B.fma_into(acc);
  }
  if (it < n) {                                          // odd number of steps: the last one is in A
    if (it == n1 && it > 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { first[r][c] = acc[r][c]; acc[r][c] = 0.0; }
    }
    A.fma_into(acc);
  }
  if (n1 > 0 && n2 == 0) {
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) { first[r][c] = acc[r][c]; acc[r][c] = 0.0; }
  }
}

__device__ __forceinline__ double warp_butterfly(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Standard normals i0 + 4 lane .. + 3 of a replicate's N stream, written to row[4 lane .. 4 lane + 3] (shared memory;
// entries past nvalid are zeroed, lanes with nvalid <= 0 write nothing), executed by a whole warp.  AS241's central
// branch is evaluated in place; the ~15 % of draws that fall in the tails are COMPACTED over the warp (ballot + rank),
// so that the log/sqrt branch runs once per warp for up to 32 tail draws instead of once per tail draw of the
// unluckiest lane.
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
  unsigned bal[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double q = pr[c] - 0.5;
    const bool valid = c < nvalid;
    const bool tail = valid && !(fabs(q) <= 0.425);
    if (nvalid > 0 && !tail) row[4 * lane + c] = valid ? ubm_qnorm_central(q) : 0.0;
    bal[c] = __ballot_sync(0xffffffffu, tail);
  }
  const int n0 = __popc(bal[0]), n1 = n0 + __popc(bal[1]), n2 = n1 + __popc(bal[2]), n3 = n2 + __popc(bal[3]);
  for (int first = 0; first < n3; first += 32) {      // one pass unless more than 32 of the 4 x 32 draws are in the tails
    const int i = first + lane;
    const int c = (i < n0) ? 0 : (i < n1) ? 1 : (i < n2) ? 2 : 3;
    const int rank = i - ((c == 0) ? 0 : (c == 1) ? n0 : (c == 2) ? n1 : n2);
    const unsigned m = (c == 0) ? bal[0] : (c == 1) ? bal[1] : (c == 2) ? bal[2] : bal[3];
    const int src = (i < n3) ? (int)__fns(m, 0, rank + 1) : 0;
    const double p0 = __shfl_sync(0xffffffffu, pr[0], src), p1 = __shfl_sync(0xffffffffu, pr[1], src);
    const double p2 = __shfl_sync(0xffffffffu, pr[2], src), p3 = __shfl_sync(0xffffffffu, pr[3], src);
    if (i < n3) {
      const double pc = (c == 0) ? p0 : (c == 1) ? p1 : (c == 2) ? p2 : p3;
      row[4 * src + c] = ubm_qnorm_tail(pc, pc - 0.5);
    }
  }
}

struct MvnSmem {
  double *packA, *packB, *packC, *MEAN;
  double *X1, *X2;                // [R][DPS]
  double *XI;                     // [2][R][DPS]  double-buffered chain-1 row: normals of this step -> proposal 1 (in place)
  double *V2;                     // [R][DPS]     chain-2 row: z -> e -> eta -> proposal 2 (in place)
  double *ACC;                    // [R][2][dimh] mcmc sums, correction sums (estimator mode)
  double *TG;                     // [2R][32]
  double *ZERO;                   // [DPS]
  double *LU;                     // [R][8] ring of log-uniforms produced ahead
  MvnSlots* S;
};

__host__ __device__ inline size_t mvn_smem_bytes(int packLen, int DPS, int R, int acc_per_slot) {
  return (size_t)(3 * packLen + DPS + 5 * R * DPS + R * acc_per_slot + 2 * R * 32 + DPS + 8 * R) * sizeof(double) +
         sizeof(MvnSlots) + 16;
}

#ifdef UBM_MVN_TIMING
static __device__ long long g_mvn_tloop;      // debug builds only: cycles thread (0,0) spent inside the GEMM loops
#endif
// GEMM phase over a compact row list (row code = slot * 2 + which).  Work is cut into BALANCED HALF-UNITS: column
// group gA = pr (K_A = 4 pr + 4 rows) is paired with gB = nCG-1-pr (K_B rows, K_A + K_B = 2 half), and the long
// column group is summed in two segments [0, K_B - half) and [K_B - half, K_B) — the engine's canonical order
// (oracle: vec_times_upper).  Lane 2i (h = 0) computes gA completely plus the first segment of gB, lane 2i+1 (h = 1)
// the second segment of gB: both stream exactly `half` matrix rows.  The partner's first-segment sums arrive by
// shuffle, so h = 0 owns the outputs of gA and h = 1 those of gB.  Unit index u = ((pr * nRG) + rg) * 2 + h.
// Must be called by all 32 lanes of every consumer warp (shuffles); every thread owns at most one output group.
template <int RT, int MODE>
__device__ __forceinline__ void mvn_phase_compute(const double* __restrict__ pack, const MvnSmem& sm, const double* V1,
                                                  const int* rows, int nrows, int DPS, int nCG, int nPairs,
                                                  const double* aux, int u, int (&code)[RT], int& g_out, bool& has_out,
                                                  double (&out)[RT][4]) {
  const int nRG = (nrows + RT - 1) / RT;
  const bool active = u < nRG * nPairs * 2;
  const int h = u & 1, ur = u >> 1;
  const int pr = active ? ur / nRG : 0, rg = active ? ur - pr * nRG : 0;
  const int gA = pr, gB = nCG - 1 - pr;
  const int KA = 4 * gA + 4, KB = 4 * gB + 4, half = mvn_half(nCG);
  const bool middle = (gA == gB);
  const int ks = (KB > half) ? KB - half : 0;
  const double* in[RT];
  const double* in2[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int idx = r * nRG + rg;      // the rows read by one load instruction are consecutive: distinct banks
    code[r] = (active && idx < nrows) ? rows[idx] : -1;
    if (MODE == 2) {
      in[r] = (code[r] < 0) ? sm.ZERO : sm.X1 + (code[r] >> 1) * DPS;
      in2[r] = (code[r] < 0) ? sm.ZERO : sm.X2 + (code[r] >> 1) * DPS;
    } else {
      in[r] = (code[r] < 0) ? sm.ZERO : ((code[r] & 1) ? sm.V2 : V1) + (code[r] >> 1) * DPS;
      in2[r] = in[r];
    }
  }
  double accA[RT][4], accB[RT][4];
#pragma unroll
  for (int r = 0; r < RT; ++r) { accA[r][0] = accA[r][1] = accA[r][2] = accA[r][3] = 0.0; accB[r][0] = accB[r][1] = accB[r][2] = accB[r][3] = 0.0; }
#ifdef UBM_MVN_TIMING
  const long long tl0_ = clock64();
#endif
  {
    // h = 0: all of gA, then the first segment [0, ks) of gB; h = 1: the second segment [ks, KB) of gB
    const int n1 = (active && h == 0 && !middle) ? KA / 2 : 0;
    const int n2 = !active ? 0 : (h == 0 ? ks / 2 : (KB - ks) / 2);
    mvn_unit<RT, MODE>(pack, nPairs, gA, n1, gB, h ? ks : 0, n2, in, in2, aux, accA, accB);
  }
#ifdef UBM_MVN_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0) g_mvn_tloop += clock64() - tl0_;
#endif
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double lo = __shfl_xor_sync(0xffffffffu, accB[r][c], 1);
      if (h == 1 && ks > 0) accB[r][c] = lo + accB[r][c];
      out[r][c] = h ? accB[r][c] : accA[r][c];
    }
  g_out = h ? gB : gA;
  has_out = active && !(h == 0 && middle);
}

template <bool TAPE, bool ZERO_MEAN>
__global__ void __launch_bounds__(MVN_T) mvn_driver_kernel(const MvnDevice M, const RunParams P, const int R,
                                                           const int acc_per_slot) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, T = MVN_T, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const int d = M.d, DPS = M.DPS, nCG = M.nCG, nPairs = M.nPairs;
  MvnSmem sm;
  {
    double* p = reinterpret_cast<double*>(smem_raw);
    sm.packA = p; p += M.packLen; sm.packB = p; p += M.packLen; sm.packC = p; p += M.packLen;
    sm.MEAN = p; p += DPS;
    sm.X1 = p; p += R * DPS; sm.X2 = p; p += R * DPS; sm.XI = p; p += 2 * R * DPS; sm.V2 = p; p += R * DPS;
    sm.ACC = p; p += R * acc_per_slot;
    sm.TG = p; p += 2 * R * 32;
    sm.ZERO = p; p += DPS;
    sm.LU = p; p += 8 * R;
    sm.S = reinterpret_cast<MvnSlots*>(p);
  }
  MvnSlots& S = *sm.S;
  for (int i = tid; i < M.packLen; i += T) { sm.packA[i] = M.packA[i]; sm.packB[i] = M.packB[i]; sm.packC[i] = M.packC[i]; }
  for (int i = tid; i < DPS; i += T) sm.MEAN[i] = M.mean_pi[i];
  for (int i = tid; i < 5 * R * DPS + R * acc_per_slot + 2 * R * 32 + DPS + 8 * R; i += T) sm.X1[i] = 0.0;
  if (tid < MVN_MAXR) {
    S.rep[tid] = -1; S.phase[tid] = SL_IDLE; S.time[tid] = 0; S.tau[tid] = 0; S.met[tid] = 0;
    S.iu[tid] = 0; S.in[tid] = 0; S.pu[tid] = 0; S.bad[tid] = 0; S.ident[tid] = 0;
  }
  if (tid == 0) { S.exhausted = 0; S.nA = S.nB = S.nI = S.nC = 0; }
  __syncthreads();

  const int dimh = (P.h_kind == UBM_H_BOTH) ? 2 * d : d;
  const double denom = (double)(P.m - P.k + 1);
  const unsigned FULL = 0xffffffffu;
  int cur = 0;
  const int RB = min(2 * MVN_MAXR, MVN_RT * ((MVN_GT / 2) / nPairs));   // rows per GEMM batch: one half-unit per consumer thread

#ifdef UBM_MVN_TIMING
  long long tacc[5] = {0, 0, 0, 0, 0}, tsteps = 0;
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
    // ================================================================ per-slot section (all 8 warps, one warp per slot)
    // finishes the step just computed (densities -> accept/reject -> state/estimator update) and runs the slot state
    // machine (finish / write-out / fetch the next replicate).  No random-number work: uniforms come from the ring.
    int my_active = 0;
    // one slot per warp; slots beyond the warp count go to the producer warps (they are not on the critical path)
    for (int s = warp; s < R; s = (s < nwarp && warp >= MVN_CW) ? nwarp + warp - MVN_CW : R) {
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
        double wgt = 0.0;
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
          if (add_corr && P.mode == MODE_ESTIMATOR) wgt = (double)corr_weight(time, P.k, P.m, P.lag);
        }
        // ---- state update, estimator sums, trajectories
        const bool coupled_step = (ph == SL_COUPLED);
        for (int j = lane; j < d; j += 32) {
          const int o = s * DPS + j;
          if (a1) sm.X1[o] = XIp[o];
          if (a2) sm.X2[o] = (coupled_step && ident) ? XIp[o] : sm.V2[o];
          const double x1 = sm.X1[o];
          const double x2 = sm.X2[o];
          if (P.mode == MODE_ESTIMATOR) {
            double* a = sm.ACC + s * acc_per_slot;
            const int nh = (P.h_kind == UBM_H_BOTH) ? 2 : 1;
            for (int hh = 0; hh < nh; ++hh) {
              const bool sq = (P.h_kind == UBM_H_SQUARE) || hh == 1;
              const int e = hh * d + j;
              const double h1 = sq ? x1 * x1 : x1;
              if (ph == SL_INIT) {
                a[e] = (P.k == 0) ? h1 : 0.0;                        // R/unbiasedestimator.R:55-59
                a[dimh + e] = 0.0;
              } else {
                if (add_mcmc) a[e] = a[e] + h1;
                if (add_corr) {
                  const double h2 = sq ? x2 * x2 : x2;
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
        if (finished) {
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
            const double* a = sm.ACC + s * acc_per_slot;
            const int nh = (P.h_kind == UBM_H_BOTH) ? 2 : 1;
            for (int j = lane; j < d; j += 32)
              for (int hh = 0; hh < nh; ++hh) {
                const int e = hh * d + j;
                const double mc = a[e], co = a[dimh + e];
                const double ue = mc + co;
                P.uest[rep * dimh + e] = ue / denom;
                if (P.mcmc) P.mcmc[rep * dimh + e] = mc / denom;
                if (P.corr) P.corr[rep * dimh + e] = co / denom;
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
      if (warp == 0) {      // row tables from ballots
        const int ph = (lane < R) ? S.phase[lane] : SL_IDLE;
        const unsigned mA = __ballot_sync(FULL, ph == SL_COUPLED), mS = __ballot_sync(FULL, ph == SL_SINGLE);
        const unsigned mI = __ballot_sync(FULL, ph == SL_INIT);
        const unsigned below = (1u << lane) - 1u;
        if (ph == SL_COUPLED) S.rowsA[__popc(mA & below)] = 2 * lane;
        if (ph == SL_INIT) { const int q = 2 * __popc(mI & below); S.rowsI[q] = 2 * lane; S.rowsI[q + 1] = 2 * lane + 1; }
        if (lane == 0) { S.nA = __popc(mA); S.nI = 2 * __popc(mI); }
        if (mA == 0) {      // no coupled slot: the B / C lists are known already
          const unsigned mBS = mS;
          if (ph == SL_SINGLE) { const int q = __popc(mBS & below); S.rowsB[q] = 2 * lane; S.rowsC[q] = 2 * lane; }
          const int nb = __popc(mBS);
          if (ph == SL_INIT) { const int q = nb + 2 * __popc(mI & below); S.rowsC[q] = 2 * lane; S.rowsC[q + 1] = 2 * lane + 1; }
          if (lane == 0) { S.nB = nb; S.nC = nb + 2 * __popc(mI); }
        }
      }
      consumer_bar();
      const int nA = S.nA;
      // ---------------------------------------------------------------- phase A: z = -((x2 - x1)^T Cinv_prop) -> V2, |z|^2
      if (nA > 0) {
        for (int r0 = 0; r0 < nA; r0 += RB) {
          int code[MVN_RT], g;
          bool has;
          double acc[MVN_RT][4];
          mvn_phase_compute<MVN_RT, 2>(sm.packA, sm, XIc, S.rowsA + r0, min(RB, nA - r0), DPS, nCG, nPairs, nullptr, tid, code, g, has, acc);
          if (has) {
#pragma unroll
            for (int r = 0; r < MVN_RT; ++r) {
              if (code[r] < 0) continue;
              const int s = code[r] >> 1;
              double t = 0.0;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const double z = -acc[r][c];                           // :198
                sm.V2[s * DPS + 4 * g + c] = z;
                t = ubm_fma(z, z, t);
              }
              sm.TG[(2 * s) * 32 + g] = t;
            }
          }
        }
        consumer_bar();
        // -------------------------------------------------------------- coupling algebra, one warp per coupled slot
        for (int i = warp; i < nA; i += MVN_CW) {
          const int s = S.rowsA[i] >> 1;
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
        if (warp == 0) {      // B / C row lists: chain-1 rows of all stepping slots, chain-2 rows where proposals differ
          const int ph = (lane < R) ? S.phase[lane] : SL_IDLE;
          const int differ = (ph == SL_COUPLED) && !S.ident[lane < R ? lane : 0];
          const unsigned m1 = __ballot_sync(FULL, ph == SL_COUPLED || ph == SL_SINGLE);
          const unsigned m2 = __ballot_sync(FULL, differ);
          const unsigned mI = __ballot_sync(FULL, ph == SL_INIT);
          const unsigned below = (1u << lane) - 1u;
          const int n1 = __popc(m1), n2 = __popc(m2);
          if (ph == SL_COUPLED || ph == SL_SINGLE) { const int q = __popc(m1 & below); S.rowsB[q] = 2 * lane; S.rowsC[q] = 2 * lane; }
          if (differ) { const int q = n1 + __popc(m2 & below); S.rowsB[q] = 2 * lane + 1; S.rowsC[q] = 2 * lane + 1; }
          if (ph == SL_INIT) { const int q = n1 + n2 + 2 * __popc(mI & below); S.rowsC[q] = 2 * lane; S.rowsC[q + 1] = 2 * lane + 1; }
          if (lane == 0) { S.nB = n1 + n2; S.nC = n1 + n2 + 2 * __popc(mI); }
        }
        consumer_bar();
      }
      TMARK(1);
      // ---------------------------------------------------------------- phase B: proposals = x + in^T C_prop (in place)
      // and, for freshly fetched slots, x0 = init_mean + in^T init_chol (matrix read from global memory / L2)
      for (int pass = 0; pass < 2; ++pass) {
        const int ntot = pass ? S.nI : S.nB;
        const int* rows_all = pass ? S.rowsI : S.rowsB;
        for (int r0 = 0; r0 < ntot; r0 += RB) {     // uniform over the consumer group; one batch unless 2R rows are live
          const int nrows = min(RB, ntot - r0);
          int code[MVN_RT], g;
          bool has;
          double acc[MVN_RT][4];
          // two call sites so that the matrix loads are LDS (shared) resp. LDG (global), not generic loads
          if (pass == 0) mvn_phase_compute<MVN_RT, 0>(sm.packB, sm, XIc, rows_all + r0, nrows, DPS, nCG, nPairs, nullptr, tid, code, g, has, acc);
          else mvn_phase_compute<MVN_RT, 0>(M.packI, sm, XIc, rows_all + r0, nrows, DPS, nCG, nPairs, nullptr, tid, code, g, has, acc);
          consumer_bar();   // every read of this batch's input rows is done: overwrite them with the outputs
          if (has) {
#pragma unroll
            for (int r = 0; r < MVN_RT; ++r) {
              if (code[r] < 0) continue;
              const int s = code[r] >> 1, which = code[r] & 1;
              const double* base = pass ? (M.init_mean + 4 * g) : ((which ? sm.X2 : sm.X1) + s * DPS + 4 * g);
              double* outp = (which ? sm.V2 : XIc) + s * DPS + 4 * g;
#pragma unroll
              for (int c = 0; c < 4; ++c) outp[c] = (4 * g + c < d) ? (base[c] + acc[r][c]) : 0.0;   // :40-44, :224-225
            }
          }
        }
      }
      consumer_bar();
      TMARK(2);
      // ---------------------------------------------------------------- phase C: w = Cinv_pi^T (prop - mean), |w|^2
      {
        const int ntot = S.nC;
        for (int r0 = 0; r0 < ntot; r0 += RB) {
          const int nrows = min(RB, ntot - r0);
          int code[MVN_RT], g;
          bool has;
          double acc[MVN_RT][4];
          mvn_phase_compute<MVN_RT, ZERO_MEAN ? 0 : 1>(sm.packC, sm, XIc, S.rowsC + r0, nrows, DPS, nCG, nPairs, sm.MEAN, tid, code, g, has, acc);
          if (has) {
#pragma unroll
            for (int r = 0; r < MVN_RT; ++r) {
              if (code[r] < 0) continue;
              double t = 0.0;
#pragma unroll
              for (int c = 0; c < 4; ++c) t = ubm_fma(acc[r][c], acc[r][c], t);
              sm.TG[code[r] * 32 + g] = t;
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
    printf("mvn timing (cycles/step, block 0, thread 0): steps %lld slot %.0f A+coupling %.0f B %.0f C %.0f wait-for-producers %.0f | GEMM loops %.0f\n", tsteps,
           (double)tacc[0] / tsteps, (double)tacc[1] / tsteps, (double)tacc[2] / tsteps, (double)tacc[3] / tsteps,
           (double)tacc[4] / tsteps, (double)g_mvn_tloop / tsteps);
#endif
}

template <class Buf>
inline int mvn_launch(const MvnDevice& M, const RunParams& P, bool tape, int nsm, cudaStream_t st, Buf& scratch,
                      int64_t* launches, std::string* err) {
  (void)scratch;
  if (P.nbreaks >= 2) { *err = "mvn: histogram bins are not available for this model yet"; return UBM_ERR_UNSUPPORTED; }
  const size_t budget = 227 * 1024;
  const int dimh = (P.h_kind == UBM_H_BOTH) ? 2 * M.d : M.d;
  const int acc_per_slot = (P.mode == MODE_ESTIMATOR) ? 2 * dimh : 0;
  int R = MVN_MAXR;
  while (R > 1 && mvn_smem_bytes(M.packLen, M.DPS, R, acc_per_slot) > budget) --R;
  size_t smem = mvn_smem_bytes(M.packLen, M.DPS, R, acc_per_slot);
  if (smem > budget || M.nPairs > MVN_GT / 2) { *err = "mvn: d too large for shared memory"; return UBM_ERR_UNSUPPORTED; }
  int64_t need = (P.nrep + R - 1) / R;
  int grid = (int)std::min<int64_t>(nsm, need);
  void (*kern)(const MvnDevice, const RunParams, const int, const int);
  if (tape) kern = M.zero_mean ? mvn_driver_kernel<true, true> : mvn_driver_kernel<true, false>;
  else kern = M.zero_mean ? mvn_driver_kernel<false, true> : mvn_driver_kernel<false, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *err = std::string("mvn: smem attribute: ") + cudaGetErrorString(e); return UBM_ERR_CUDA; }
  kern<<<grid, MVN_T, smem, st>>>(M, P, R, acc_per_slot);
  *launches += 1;
  return UBM_OK;
}

}  // namespace ubm
