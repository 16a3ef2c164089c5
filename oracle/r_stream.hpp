// r_stream.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Emulation of the R runtime's DEFAULT random number stream, so that the oracle can be run on the very stream an R
// session would have used and be compared with the outputs printed in the reference's README.md (the only literal
// outputs in the reference tree: README.md:146 mean cost 2072.788, README.md:171 "0.005471002 +/- 0.1458675", both
// produced after set.seed(1)).
//
// R is not part of /root/reference; what follows restates, from the published algorithms, what the R sources do:
//   * set.seed(seed)       src/main/RNG.c: RNG_Init — 50 rounds of the LCG seed = 69069 seed + 1 ("initial scrambling"),
//                          then 625 more rounds fill dummy[0..624]; FixupSeeds sets dummy[0] = mti = 624.
//   * unif_rand()          Mersenne-Twister MT19937 (Matsumoto & Nishimura 1998): genrand with the standard tempering,
//                          scaled by 2.3283064365386963e-10, then fixup() into the open interval.
//   * norm_rand()          kind "Inversion": u = unif_rand(); u = (int)(2^27 u) + unif_rand(); qnorm5(u / 2^27).
//                          qnorm5 = Wichura's AS241 PPND16 (plain Horner evaluation, no fused multiply-add).
//   * exp_rand()           Ahrens & Dieter (1972) algorithm SA with the table q[k] = sum_{i<=k} ln(2)^i / i!.
// Pinned by the widely published first draws of set.seed(1), set.seed(42), set.seed(123) (tests/test_r_stream.py) and,
// end to end, by reproducing the README literals through the oracle's drivers.
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

struct RStream {
  uint32_t mt[624];
  int mti = 625;

  explicit RStream(uint32_t seed = 1) { set_seed(seed); }

  void set_seed(uint32_t seed) {
    for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;      // initial scrambling
    uint32_t dummy[625];
    for (int j = 0; j < 625; ++j) { seed = 69069u * seed + 1u; dummy[j] = seed; }
    for (int j = 0; j < 624; ++j) mt[j] = dummy[j + 1];          // dummy[0] is mti
    mti = 624;                                                    // FixupSeeds(MERSENNE_TWISTER, initial = 1)
  }

  uint32_t genrand() {
    const int N = 624, M = 397;
    const uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    if (mti >= N) {
      int kk;
      for (kk = 0; kk < N - M; ++kk) {
        uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
      }
      for (; kk < N - 1; ++kk) {
        uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
      }
      uint32_t y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
      mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }

  double unif_rand() {
    const double i2_32m1 = 2.328306437080797e-10;                 // 1 / (2^32 - 1)
    double x = (double)genrand() * 2.3283064365386963e-10;        // [0, 1)
    if (x <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return x;
  }

  // AS241 PPND16 as coded in R's qnorm5 (lower tail, not log): plain Horner sums
  static double qnorm5(double p) {
    double q = p - 0.5, r, val;
    if (std::fabs(q) <= 0.425) {
      r = .180625 - q * q;
      val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                     45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                  133.14166789178437745) * r + 3.387132872796366608) /
            (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
              42.313330701600911252) * r + 1.);
      return val;
    }
    r = (q < 0) ? p : (0.5 - p + 0.5);                            // R_DT_CIv
    r = std::sqrt(-std::log(r));
    if (r <= 5.) {
      r += -1.6;
      val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
                 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
              4.6303378461565452959) * r + 1.42343711074968357734) /
            (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
                 .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
              2.05319162663775882187) * r + 1.);
    } else {
      r += -5.;
      val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
                 .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
              5.4637849111641143699) * r + 6.6579046435011037772) /
            (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                 7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
              .59983220655588793769) * r + 1.);
    }
    return (q < 0.0) ? -val : val;
  }

  double norm_rand() {
    const double BIG = 134217728.0;                               // 2^27
    double u = unif_rand();
    u = (double)(int)(BIG * u) + unif_rand();
    return qnorm5(u / BIG);
  }

  double exp_rand() {
    static const double q[] = {0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
                               0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
                               0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
                               0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000};
    double a = 0.;
    double u = unif_rand();
    while (u <= 0. || u >= 1.) u = unif_rand();
    for (;;) {
      u += u;
      if (u > 1.) break;
      a += q[0];
    }
    u -= 1.;
    if (u <= q[0]) return a + u;
    int i = 0;
    double ustar = unif_rand(), umin = ustar;
    do {
      ustar = unif_rand();
      if (umin > ustar) umin = ustar;
      i++;
    } while (u > q[i]);
    return a + umin * q[0];
  }
};

}  // namespace orc
