/*
 * ubm_pg.h — numerics contract, part 2: log Phi(x) for the Polya-Gamma path.
 *
 * The vendored BayesLogit sampler of the reference (src/PolyaGamma.cpp:89-104, mass_texpon) calls R's
 * pnorm(x, 0, 1, lower = TRUE, log.p = TRUE) through RNG::p_norm (src/RRNG.cpp:92-95).  R's pnorm is Cody's
 * (1969) rational Chebyshev algorithm (nmath/pnorm.c, pnorm_both); it is restated here for the lower tail on the
 * log scale, with the engine's ubm_exp / ubm_log / ubm_log1p, so that host and device agree bit for bit.
 * Validated against scipy.special.log_ndtr in tests/test_numerics.py.
 */
#ifndef UBM_PG_H
#define UBM_PG_H

#include "ubm_numerics.h"

#define UBM_1_SQRT_2PI 0.398942280401432677939946059934

/* log P(Z <= x) */
UBM_HD double ubm_pnorm_log(double x) {
  const double y = fabs(x);
  if (y <= 0.67448975) {
    double xnum = 0.0, xden = 0.0;
    if (y > 1.1102230246251565e-16) {
      const double xsq = x * x;
      xnum = 0.065682337918207449113 * xsq;
      xden = xsq;
      xnum = (xnum + 2.2352520354606839287) * xsq;   xden = (xden + 47.20258190468824187) * xsq;
      xnum = (xnum + 161.02823106855587881) * xsq;   xden = (xden + 976.09855173777669322) * xsq;
      xnum = (xnum + 1067.6894854603709582) * xsq;   xden = (xden + 10260.932208618978205) * xsq;
    }
    const double temp = x * (xnum + 18154.981253343561249) / (xden + 45507.789335026729956);
    return ubm_log(0.5 + temp);
  }
  double temp;
  if (y <= 5.656854249492380195206754896838) {       /* sqrt(32) */
    double xnum = 1.0765576773720192317e-8 * y, xden = y;
    xnum = (xnum + 0.39894151208813466764) * y;   xden = (xden + 22.266688044328115691) * y;
    xnum = (xnum + 8.8831497943883759412) * y;    xden = (xden + 235.38790178262499861) * y;
    xnum = (xnum + 93.506656132177855979) * y;    xden = (xden + 1519.377599407554805) * y;
    xnum = (xnum + 597.27027639480026226) * y;    xden = (xden + 6485.558298266760755) * y;
    xnum = (xnum + 2494.5375852903726711) * y;    xden = (xden + 18615.571640885098091) * y;
    xnum = (xnum + 6848.1904505362823326) * y;    xden = (xden + 34900.952721145977266) * y;
    xnum = (xnum + 11602.651437647350124) * y;    xden = (xden + 38912.003286093271411) * y;
    temp = (xnum + 9842.7148383839780218) / (xden + 19685.429676859990727);
  } else {
    const double xsq = 1.0 / (x * x);
    double xnum = 0.02307344176494017303 * xsq, xden = xsq;
    xnum = (xnum + 0.21589853405795699) * xsq;       xden = (xden + 1.28426009614491121) * xsq;
    xnum = (xnum + 0.1274011611602473639) * xsq;     xden = (xden + 0.468238212480865118) * xsq;
    xnum = (xnum + 0.022235277870649807) * xsq;      xden = (xden + 0.0659881378689285515) * xsq;
    xnum = (xnum + 0.001421619193227893466) * xsq;   xden = (xden + 0.00378239633202758244) * xsq;
    temp = xsq * (xnum + 2.9112874951168792e-5) / (xden + 7.29751555083966205e-5);
    temp = (UBM_1_SQRT_2PI - temp) / y;
  }
  /* do_del: split y^2 = xsq^2 + del with xsq = trunc(16 y) / 16 to keep exp() accurate */
  const double xs = trunc(y * 16.0) / 16.0;
  const double del = (y - xs) * (y + xs);
  if (x <= 0.0) return (-xs * xs * 0.5) + (-del * 0.5) + ubm_log(temp);          /* the small tail itself */
  return ubm_log1p(-ubm_exp(-xs * xs * 0.5) * ubm_exp(-del * 0.5) * temp);        /* 1 - upper tail */
}

#endif /* UBM_PG_H */
