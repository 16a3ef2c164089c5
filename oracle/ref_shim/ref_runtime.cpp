// ref_runtime.cpp — TEST INFRASTRUCTURE: the R-math stand-ins behind oracle/ref_shim/Rmath.h.
// Primitive draws come from typed replay tapes (or from the emulated R session stream of oracle/r_stream.hpp), so the
// unmodified reference sources and the oracle restatement can be fed the very same draws.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include "Rmath.h"
#include "ref_runtime.h"
#include "../r_stream.hpp"
#include "../../include/ubm_pg.h"

namespace {
thread_local const double *g_u = nullptr, *g_n = nullptr, *g_e = nullptr;
thread_local int64_t g_lu = 0, g_ln = 0, g_le = 0, g_iu = 0, g_in = 0, g_ie = 0;
thread_local int g_exhausted = 0;
thread_local orc::RStream* g_rs = nullptr;
[[noreturn]] void off_path(const char* what) {
  fprintf(stderr, "oracle/_ref: %s is not on the hot path and has no stand-in\n", what);
  abort();
}
}  // namespace

extern "C" {
void ref_tape_set(const double* u, int64_t lu, const double* n, int64_t ln, const double* e, int64_t le) {
  g_u = u; g_n = n; g_e = e; g_lu = u ? lu : 0; g_ln = n ? ln : 0; g_le = e ? le : 0;
  g_iu = g_in = g_ie = 0; g_exhausted = 0; g_rs = nullptr;
}
void ref_tape_used(int64_t* iu, int64_t* in, int64_t* ie, int32_t* exhausted) {
  if (iu) *iu = g_iu; if (in) *in = g_in; if (ie) *ie = g_ie; if (exhausted) *exhausted = g_exhausted;
}
void ref_use_rsession(void* rstream) { g_rs = reinterpret_cast<orc::RStream*>(rstream); g_exhausted = 0; }
}

double unif_rand() {
  if (g_rs) return g_rs->unif_rand();
  if (g_iu >= g_lu) { g_exhausted = 1; ++g_iu; return 0.5; }
  return g_u[g_iu++];
}
double norm_rand() {
  if (g_rs) return g_rs->norm_rand();
  if (g_in >= g_ln) { g_exhausted = 1; ++g_in; return 0.0; }
  return g_n[g_in++];
}
double exp_rand() {
  if (g_rs) return g_rs->exp_rand();
  if (g_ie >= g_le) { g_exhausted = 1; ++g_ie; return 1.0; }
  return g_e[g_ie++];
}
double rnorm(double mu, double sigma) { return mu + sigma * norm_rand(); }          // nmath/rnorm.c
double runif(double a, double b) {                                                   // nmath/runif.c
  double u;
  do { u = unif_rand(); } while (u <= 0 || u >= 1);
  return a + (b - a) * u;
}
double rexp(double scale) { return scale * exp_rand(); }                             // nmath/rexp.c
double rchisq(double) { off_path("rchisq"); }
double rgamma(double, double) { off_path("rgamma"); }
double rbeta(double, double) { off_path("rbeta"); }
double pgamma(double, double, double, int, int) { off_path("pgamma"); }
double lgammafn(double) { off_path("lgammafn"); }
double dbeta(double, double, double, int) { off_path("dbeta"); }
double pnorm(double x, double mu, double sigma, int lower_tail, int log_p) {
  double z = (x - mu) / sigma;
  if (!lower_tail) z = -z;
  const double lp = ubm_pnorm_log(z);
  return log_p ? lp : ubm_exp(lp);
}
