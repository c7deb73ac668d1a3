// Internal helpers shared by the .cu files behind the C ABI declared in include/szn.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/szn.h"

namespace szn {
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
