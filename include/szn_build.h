/* szn_build.h — build identity of libszn.so (kept apart from szn.h so that asking for it does not change what is hashed).
 * The reference has no counterpart: it is Python (no build).  Used by bench.py (`libszn_build_id`) and tests/test_abi.py. */
#ifndef SZN_BUILD_H
#define SZN_BUILD_H
#ifdef __cplusplus
extern "C" {
#endif
/* first 16 hex digits of sha256 over the kernel sources (csrc/Makefile: HASHED, in that order) and the nvcc flags */
const char* szn_build_id(void);
#ifdef __cplusplus
}
#endif
#endif
