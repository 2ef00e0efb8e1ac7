// Shared helpers for libconvnp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CNP_API extern "C" __attribute__((visibility("default")))

// thread-local last-error string, see cnp_last_error()
void cnp_set_error(const char* fmt, ...);

#define CNP_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      cnp_set_error(__VA_ARGS__);              \
      return -1;                               \
    }                                          \
  } while (0)

// Return the launch status without synchronising (callee never synchronises).
#define CNP_LAUNCH_CHECK(name)                                             \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      cnp_set_error("%s: %s", name, cudaGetErrorString(e__));              \
      return (int)e__;                                                     \
    }                                                                      \
  } while (0)

// exp(-104) underflows to 0 in fp32 (round-to-nearest of 6.8e-46 against the 1.4e-45 denormal),
// so every set-conv term with 0.5*d^2/s^2 > 104 is exactly zero in the dense reference einsum.
// Truncating at that radius leaves only exact zeros out of the sums.
#define CNP_EXP_CUTOFF 104.0f

__device__ __forceinline__ float cnp_rbf(float a, float b, float scale2) {
  float d = a - b;
  return expf(-0.5f * (d * d) / scale2);
}

__device__ __forceinline__ float cnp_grid_pt(double start, double res, int i) {
  return (float)(start + (double)i * res);
}

static inline int cnp_cdiv(int a, int b) { return (a + b - 1) / b; }
