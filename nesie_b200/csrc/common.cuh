// Shared helpers for libnesie_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nesie_b200.h"

namespace nesie {

// Last error text of the calling thread (read by nesie_last_error()).
void set_error(const char *fmt, ...);

inline int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return NESIE_OK;
}

#define NESIE_REQUIRE(cond, msg)                                  \
  do {                                                            \
    if (!(cond)) {                                                \
      nesie::set_error("%s: %s", __func__, msg);                  \
      return NESIE_ERR_INVALID_ARG;                               \
    }                                                             \
  } while (0)

#define NESIE_CUDA(call)                                                         \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      nesie::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
      return (int)e_;                                                            \
    }                                                                            \
  } while (0)

// The reference's squared distance, exactly as nvcc contracts it (default -fmad=true):
//   t = dy*dy; t = fma(dx,dx,t); d = fma(dz,dz,t)      (a minus b)
__device__ __forceinline__ float sqdist_ref(float ax, float ay, float az, float bx, float by,
                                            float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace nesie
