// R.h — TEST INFRASTRUCTURE: stand-in for the R API header (see Rcpp.h in this directory).
#ifndef UBM_REF_SHIM_R_H
#define UBM_REF_SHIM_R_H
#include <cstdio>
inline void GetRNGstate() {}
inline void PutRNGstate() {}
inline void R_CheckUserInterrupt() {}
#define Rprintf printf
#endif
