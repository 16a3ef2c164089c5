/*
 * ubm_numerics.h — the NUMERICS CONTRACT of the ubm-b200 engine.
 *
 * The reference package (pierrejacob/unbiasedmcmc) takes every primitive random draw and every
 * transcendental from the R runtime (unif_rand / norm_rand / exp_rand, log, exp, dnorm, pnorm; see
 * reference src/RRNG.cpp:50-95, src/mvnorm.cpp:16,36,104,194,214, src/ising.cpp:57,85).  R is not
 * part of the reference tree and a GPU cannot call it, so this header DEFINES the replacement for that
 * layer, once, as `static inline` functions that compile unchanged for the host (g++) and the device
 * (nvcc, sm_100a).  Every function is written with explicit fma() and a fixed operation order; the
 * build disables implicit contraction on both sides (nvcc -fmad=false, g++ -ffp-contract=off), and uses
 * only IEEE-754 correctly-rounded primitives (+ - * / sqrt fma), so host and device produce IDENTICAL
 * bits.  That is what makes meeting times, costs and bin counts comparable bit-for-bit between the CUDA
 * engine and the CPU oracle (oracle/ includes this header; the product never includes oracle/).
 *
 * Contents
 *   - Philox4x32-10 counter-based generator (Salmon et al., SC'11), typed per-replicate streams
 *   - U(0,1) from one 32-bit word (R's unif_rand also has 32-bit resolution), 53-bit U for N / E draws
 *   - ubm_log, ubm_exp, ubm_log1p (|err| <~ 1 ulp, validated against libm in tests/test_numerics.py)
 *   - ubm_qnorm (Wichura AS241 PPND16, the algorithm behind R's qnorm / inversion norm_rand)
 *   - ubm_pnorm_log_upper-style helpers live in ubm_pg.h (only the Polya-Gamma path needs them)
 *   - ubm_sum32: the canonical 32-way interleaved + butterfly reduction order
 */
#ifndef UBM_NUMERICS_H
#define UBM_NUMERICS_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define UBM_HD __host__ __device__ __forceinline__
#else
#define UBM_HD static inline
#endif

/* ------------------------------------------------------------------------------------------------
 * bit casts
 * ---------------------------------------------------------------------------------------------- */
UBM_HD uint64_t ubm_d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
UBM_HD double ubm_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
UBM_HD double ubm_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t w[4]; } ubm_u32x4;

#define UBM_PHILOX_M0 0xD2511F53u
#define UBM_PHILOX_M1 0xCD9E8D57u
#define UBM_PHILOX_W0 0x9E3779B9u
#define UBM_PHILOX_W1 0xBB67AE85u

UBM_HD void ubm_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#if defined(__CUDA_ARCH__)
  *hi = __umulhi(a, b);
  *lo = a * b;
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
#endif
}

UBM_HD ubm_u32x4 ubm_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                   uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    ubm_mulhilo32(UBM_PHILOX_M0, c0, &hi0, &lo0);
    ubm_mulhilo32(UBM_PHILOX_M1, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += UBM_PHILOX_W0;
    k1 += UBM_PHILOX_W1;
  }
  ubm_u32x4 out;
  out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
  return out;
}

/* Typed streams.  Draw #i of type T of replicate r under seed s lives in Philox block
 *   key = (s_lo, s_hi),  ctr = (blk_lo, (T << 28) | blk_hi, r_lo, r_hi)
 * with blk = i / 4 for U (one 32-bit word per draw) and blk = i / 2 for N and E (two words per draw).
 * Random access by construction: any lane can produce any draw. */
#define UBM_STREAM_U 0u
#define UBM_STREAM_N 1u
#define UBM_STREAM_E 2u

UBM_HD ubm_u32x4 ubm_stream_block(uint64_t seed, uint64_t rep, uint32_t type, uint64_t blk) {
  return ubm_philox4x32_10((uint32_t)blk, (type << 28) | ((uint32_t)(blk >> 32) & 0x0FFFFFFFu),
                           (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

/* (w + 1/2) 2^-32: exact in binary64, strictly inside (0,1), symmetric about 1/2. */
UBM_HD double ubm_u01_32(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386962890625e-10; }
/* 53-bit version from two words, also exact and strictly inside (0,1). */
UBM_HD double ubm_u01_53(uint32_t hi, uint32_t lo) {
  double a = (double)(hi >> 5);  /* 27 bits */
  double b = (double)(lo >> 6);  /* 26 bits */
  return (a * 67108864.0 + b + 0.5) * 1.1102230246251565404236316680908203125e-16; /* 2^-53 */
}

/* ------------------------------------------------------------------------------------------------
 * log, exp, log1p — fixed-order fma polynomials; every step is an IEEE primitive
 * ---------------------------------------------------------------------------------------------- */
#define UBM_LN2_HI 6.93147180369123816490e-01
#define UBM_LN2_LO 1.90821492927058770002e-10
#define UBM_INV_LN2 1.44269504088896338700e+00

/* Natural log.  x = 2^k m with m in [sqrt(1/2), sqrt(2)); f = m - 1; s = f / (2 + f);
 * log(m) = 2s + s R(s^2), R an odd-power tail polynomial (fdlibm e_log.c layout, fma evaluation).
 * Special values: log(0) = -inf, log(<0) = NaN, log(inf) = inf, log(NaN) = NaN. */
UBM_HD double ubm_log(double x) {
  uint64_t ix = ubm_d2u(x);
  int k = 0;
  if ((int64_t)ix < (int64_t)0x0010000000000000LL) {          /* x < 2^-1022 (incl. negatives) */
    if ((ix << 1) == 0) return -INFINITY;                      /* +-0 */
    if ((int64_t)ix < 0) return NAN;                           /* negative */
    x *= 18014398509481984.0;                                  /* 2^54: scale subnormal */
    ix = ubm_d2u(x);
    k = -54;
  }
  if (ix >= 0x7FF0000000000000ULL) return x + x;               /* inf or NaN */
  k += (int)(ix >> 52) - 1023;
  uint64_t mant = ix & 0x000FFFFFFFFFFFFFULL;
  /* normalise m into [sqrt(1/2), sqrt(2)) */
  uint64_t add = (mant + 0x95F6400000000ULL) & 0x0010000000000000ULL;   /* 1 iff m >= sqrt(2) */
  double m = ubm_u2d(mant | (add ^ 0x3FF0000000000000ULL));
  k += (int)(add >> 52);
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double dk = (double)k;
  double t1 = w * ubm_fma(w, ubm_fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01),
                          3.999999999940941908e-01);
  double t2 = z * ubm_fma(w, ubm_fma(w, ubm_fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                                     2.857142874366239149e-01), 6.666666666666735130e-01);
  double R = t2 + t1;
  double hfsq = 0.5 * f * f;
  /* log(1+f) = f - (hfsq - s (hfsq + R)) */
  double inner = ubm_fma(s, hfsq + R, ubm_fma(dk, UBM_LN2_LO, 0.0));
  return ubm_fma(dk, UBM_LN2_HI, 0.0) - ((hfsq - inner) - f);
}

/* exp.  k = round(x / ln2), r = x - k ln2 (two-step Cody-Waite), exp(r) by a degree-13 Taylor/Horner
 * polynomial in fma form (truncation < 4e-18 on |r| <= ln2/2), scaled by 2^k in two exact steps.
 * exp(0) == 1 exactly.  Overflow -> +inf, underflow -> 0 (gradual, through the second scaling). */
UBM_HD double ubm_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return INFINITY;
  if (x < -745.1332191019412) return 0.0;
  double kd = rint(x * UBM_INV_LN2);
  int k = (int)kd;
  double r = ubm_fma(-kd, UBM_LN2_HI, x);
  r = ubm_fma(-kd, UBM_LN2_LO, r);
  double p = 1.6059043836821614599e-10;              /* 1/13! */
  p = ubm_fma(p, r, 2.0876756987868098979e-09);      /* 1/12! */
  p = ubm_fma(p, r, 2.5052108385441718775e-08);      /* 1/11! */
  p = ubm_fma(p, r, 2.7557319223985890653e-07);      /* 1/10! */
  p = ubm_fma(p, r, 2.7557319223985892511e-06);      /* 1/9!  */
  p = ubm_fma(p, r, 2.4801587301587301566e-05);      /* 1/8!  */
  p = ubm_fma(p, r, 1.9841269841269841253e-04);      /* 1/7!  */
  p = ubm_fma(p, r, 1.3888888888888889419e-03);      /* 1/6!  */
  p = ubm_fma(p, r, 8.3333333333333332177e-03);      /* 1/5!  */
  p = ubm_fma(p, r, 4.1666666666666664354e-02);      /* 1/4!  */
  p = ubm_fma(p, r, 1.6666666666666665741e-01);      /* 1/3!  */
  p = ubm_fma(p, r, 0.5);
  p = ubm_fma(p, r, 1.0);
  p = ubm_fma(p, r, 1.0);
  /* scale by 2^k = 2^k1 * 2^k2 with both factors normal */
  int k1 = k / 2, k2 = k - k1;
  double s1 = ubm_u2d((uint64_t)(1023 + k1) << 52);
  double s2 = ubm_u2d((uint64_t)(1023 + k2) << 52);
  return (p * s1) * s2;
}

/* log1p for x > -1: the classic compensated form log(u) + (x - (u - 1)) / u with u = 1 + x
 * (exact when u == 1). */
UBM_HD double ubm_log1p(double x) {
  double u = 1.0 + x;
  if (u == 1.0) return x;
  double c = x - (u - 1.0);
  return ubm_log(u) + c / u;
}

/* ------------------------------------------------------------------------------------------------
 * qnorm: Wichura (1988) algorithm AS241, PPND16 — the routine behind R's qnorm5 and therefore behind
 * R's default ("Inversion") norm_rand.  One uniform in, one standard normal out.
 * ---------------------------------------------------------------------------------------------- */
/* central branch: |q| <= 0.425, q = p - 1/2 */
UBM_HD double ubm_qnorm_central(double q) {
  double r = 0.180625 - q * q;
  double num = 2509.0809287301226727;
  num = ubm_fma(num, r, 33430.575583588128105);
  num = ubm_fma(num, r, 67265.770927008700853);
  num = ubm_fma(num, r, 45921.953931549871457);
  num = ubm_fma(num, r, 13731.693765509461125);
  num = ubm_fma(num, r, 1971.5909503065514427);
  num = ubm_fma(num, r, 133.14166789178437745);
  num = ubm_fma(num, r, 3.387132872796366608);
  double den = 5226.495278852545925;
  den = ubm_fma(den, r, 28729.085735721942674);
  den = ubm_fma(den, r, 39307.89580009271061);
  den = ubm_fma(den, r, 21213.794301586595867);
  den = ubm_fma(den, r, 5394.1960214247511077);
  den = ubm_fma(den, r, 687.1870074920579083);
  den = ubm_fma(den, r, 42.313330701600911252);
  den = ubm_fma(den, r, 1.0);
  return q * num / den;
}
/* tail branches: |q| > 0.425 */
UBM_HD double ubm_qnorm_tail(double p, double q) {
  double pp = (q < 0.0) ? p : (1.0 - p);
  double r = sqrt(-ubm_log(pp));
  double val;
  if (r <= 5.0) {
    r -= 1.6;
    double num = 7.7454501427834140764e-4;
    num = ubm_fma(num, r, 0.0227238449892691845833);
    num = ubm_fma(num, r, 0.24178072517745061177);
    num = ubm_fma(num, r, 1.27045825245236838258);
    num = ubm_fma(num, r, 3.64784832476320460504);
    num = ubm_fma(num, r, 5.7694972214606914055);
    num = ubm_fma(num, r, 4.6303378461565452959);
    num = ubm_fma(num, r, 1.42343711074968357734);
    double den = 1.05075007164441684324e-9;
    den = ubm_fma(den, r, 5.475938084995344946e-4);
    den = ubm_fma(den, r, 0.0151986665636164571966);
    den = ubm_fma(den, r, 0.14810397642748007459);
    den = ubm_fma(den, r, 0.68976733498510000455);
    den = ubm_fma(den, r, 1.6763848301838038494);
    den = ubm_fma(den, r, 2.05319162663775882187);
    den = ubm_fma(den, r, 1.0);
    val = num / den;
  } else {
    r -= 5.0;
    double num = 2.01033439929228813265e-7;
    num = ubm_fma(num, r, 2.71155556874348757815e-5);
    num = ubm_fma(num, r, 0.0012426609473880784386);
    num = ubm_fma(num, r, 0.026532189526576123093);
    num = ubm_fma(num, r, 0.29656057182850489123);
    num = ubm_fma(num, r, 1.7848265399172913358);
    num = ubm_fma(num, r, 5.4637849111641143699);
    num = ubm_fma(num, r, 6.6579046435011037772);
    double den = 2.04426310338993978564e-15;
    den = ubm_fma(den, r, 1.4215117583164458887e-7);
    den = ubm_fma(den, r, 1.8463183175100546818e-5);
    den = ubm_fma(den, r, 7.868691311456132591e-4);
    den = ubm_fma(den, r, 0.0148753612908506148525);
    den = ubm_fma(den, r, 0.13692988092273580531);
    den = ubm_fma(den, r, 0.59983220655588793769);
    den = ubm_fma(den, r, 1.0);
    val = num / den;
  }
  return (q < 0.0) ? -val : val;
}
UBM_HD double ubm_qnorm(double p) {
  double q = p - 0.5;
  if (fabs(q) <= 0.425) return ubm_qnorm_central(q);
  return ubm_qnorm_tail(p, q);
}

/* ------------------------------------------------------------------------------------------------
 * typed draws from Philox blocks
 * ---------------------------------------------------------------------------------------------- */
UBM_HD double ubm_block_u(const ubm_u32x4 b, uint32_t slot4) { return ubm_u01_32(b.w[slot4]); }
UBM_HD double ubm_block_u53(const ubm_u32x4 b, uint32_t slot2) {
  return ubm_u01_53(b.w[2 * slot2], b.w[2 * slot2 + 1]);
}
/* stateless random-access versions (one Philox call per draw) */
UBM_HD double ubm_philox_u(uint64_t seed, uint64_t rep, uint64_t i) {
  ubm_u32x4 b = ubm_stream_block(seed, rep, UBM_STREAM_U, i >> 2);
  uint32_t s = (uint32_t)(i & 3u);
  uint32_t w = (s == 0) ? b.w[0] : (s == 1) ? b.w[1] : (s == 2) ? b.w[2] : b.w[3];
  return ubm_u01_32(w);
}
UBM_HD double ubm_philox_u53(uint64_t seed, uint64_t rep, uint32_t type, uint64_t i) {
  ubm_u32x4 b = ubm_stream_block(seed, rep, type, i >> 1);
  return (i & 1u) ? ubm_u01_53(b.w[2], b.w[3]) : ubm_u01_53(b.w[0], b.w[1]);
}
UBM_HD double ubm_philox_n(uint64_t seed, uint64_t rep, uint64_t i) {
  return ubm_qnorm(ubm_philox_u53(seed, rep, UBM_STREAM_N, i));
}
UBM_HD double ubm_philox_e(uint64_t seed, uint64_t rep, uint64_t i) {
  return -ubm_log(ubm_philox_u53(seed, rep, UBM_STREAM_E, i));
}

/* ------------------------------------------------------------------------------------------------
 * canonical reduction order: 32 interleaved partial sums, then an xor-butterfly (16, 8, 4, 2, 1).
 * On the device a warp holds partial q in lane q and runs the butterfly with __shfl_xor_sync; on the
 * host ubm_butterfly32 replays the same tree.  IEEE addition is commutative, so every lane ends with
 * the same bits.
 * ---------------------------------------------------------------------------------------------- */
UBM_HD double ubm_butterfly32(double* s /* 32 partials, overwritten */) {
  for (int off = 16; off >= 1; off >>= 1) {
    double t[32];
    for (int q = 0; q < 32; ++q) t[q] = s[q] + s[q ^ off];
    for (int q = 0; q < 32; ++q) s[q] = t[q];
  }
  return s[0];
}

#endif /* UBM_NUMERICS_H */
