// Internal helpers shared by the .cu files behind the C ABI declared in include/szn.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/szn.h"

// Default of the kernels that are new in this build (pool routing codes, re-blocked conv1_1 weight gradient, two conv1_1
// tensor-core CTAs per SM).  Each has its own environment switch, read per call, so that one process can run both forms.
#define SZN_NEW_KERNELS_DEFAULT 0

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
