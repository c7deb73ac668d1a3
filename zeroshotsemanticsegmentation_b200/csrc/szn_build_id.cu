// Which sources this libszn.so was built from: the Makefile hashes the kernel sources (every .cu / .cuh / .h that holds
// device or launch code, and include/szn.h) together with the compiler flags and passes the first 16 hex digits in.
// nvcc object files are not reproducible byte for byte (temporary-file names end up in them), so a hash of the .so cannot
// tell whether two builds are the same code; this can.  bench.py prints it and matches profiles/r02_umma_traffic.json by it;
// tests/test_abi.py checks it against the sources in the tree (a stale libszn.so fails).
#include "../../include/szn_build.h"
#ifndef SZN_SOURCE_HASH
#define SZN_SOURCE_HASH "unknown"
#endif
extern "C" const char* szn_build_id(void) { return SZN_SOURCE_HASH; }
