// Internal helpers shared by the .cu files behind the C ABI declared in include/szn.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/szn.h"

// Default of the kernels added last (re-blocked conv1_1 weight gradient: SZN_CONV1_1_WGRAD_V2; the Python side has the
// same switch for the pool routing codes: engine.py SZN_POOL_CODE).  The switches are read per call, so one process can run
// both forms (tests, A/B runs).
#define SZN_NEW_KERNELS_DEFAULT 1

namespace szn {
inline int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}
int set_error(int code, const char* msg);  // records msg for szn_last_error(), returns code
void count_launch();                       // one more kernel of this library launched (szn_launch_count)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    return set_error(SZN_ERR_CUDA, buf);
  }
  count_launch();
  return 0;
}
}  // namespace szn
