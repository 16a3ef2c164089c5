// ref_runtime.h — TEST INFRASTRUCTURE: tape control of the R-math stand-ins (oracle/ref_shim/ref_runtime.cpp).
#ifndef UBM_REF_RUNTIME_H
#define UBM_REF_RUNTIME_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* install typed tapes for the calling thread; cursors restart at 0 */
void ref_tape_set(const double* u, int64_t lu, const double* n, int64_t ln, const double* e, int64_t le);
/* how many draws of each type were consumed since ref_tape_set, and whether a tape ran out */
void ref_tape_used(int64_t* iu, int64_t* in, int64_t* ie, int32_t* exhausted);
/* draw from an emulated R session (orc::RStream*) instead of tapes */
void ref_use_rsession(void* rstream);
#ifdef __cplusplus
}
#endif
#endif
