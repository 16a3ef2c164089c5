// Rmath.h — TEST INFRASTRUCTURE: stand-in for R's nmath, which is not part of the reference tree.
// The three primitive generators read the typed replay tapes installed with ref_tape_set() (oracle/ref_shim/ref_runtime.cpp);
// the distribution wrappers restate R's definitions on top of them (rnorm = mu + sigma * norm_rand(), rexp = scale *
// exp_rand(), runif = a + (b - a) * unif_rand()); pnorm is Cody's algorithm as restated in include/ubm_pg.h.
#ifndef UBM_REF_SHIM_RMATH_H
#define UBM_REF_SHIM_RMATH_H
double unif_rand();
double norm_rand();
double exp_rand();
double rnorm(double mu, double sigma);
double runif(double a, double b);
double rexp(double scale);
double rchisq(double df);
double rgamma(double shape, double scale);
double rbeta(double a, double b);
double pnorm(double x, double mu, double sigma, int lower_tail, int log_p);
double pgamma(double x, double shape, double scale, int lower_tail, int log_p);
double lgammafn(double x);
double dbeta(double x, double a, double b, int give_log);
#endif
