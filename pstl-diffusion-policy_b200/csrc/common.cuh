// common.cuh — error plumbing shared by the translation units of libpstl_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pstl.h"

void pstl_set_error(const char* fmt, ...);
void pstl_count_launch(void);  // api.cu: kernels launched by this library (pstl_launch_count)

#define PSTL_CHECK_ARG(cond, msg)                   \
  do {                                              \
    if (!(cond)) {                                  \
      pstl_set_error("%s: %s", __func__, msg);      \
      return PSTL_ERR_ARG;                          \
    }                                               \
  } while (0)

#define PSTL_CUDA(call)                                                               \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      pstl_set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_));        \
      return PSTL_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

#define PSTL_LAUNCH_CHECK()                                                           \
  do {                                                                                \
    cudaError_t e_ = cudaGetLastError();                                              \
    if (e_ != cudaSuccess) {                                                          \
      pstl_set_error("%s: launch failed -> %s", __func__, cudaGetErrorString(e_));    \
      return PSTL_ERR_CUDA;                                                           \
    }                                                                                 \
    pstl_count_launch();                                                              \
  } while (0)

static inline int pstl_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
