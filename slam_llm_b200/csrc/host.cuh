// Host-side helpers shared by every translation unit of libslam_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace slam {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

#define SLAM_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      ::slam::set_error(__VA_ARGS__);  \
      return -1;                       \
    }                                  \
  } while (0)

// after a kernel launch: surface launch-configuration errors without synchronising
#define SLAM_LAUNCH_CHECK(name)                                                  \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    ::slam::count_launch();                                                      \
    if (e__ != cudaSuccess) {                                                    \
      ::slam::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return static_cast<int>(e__);                                              \
    }                                                                            \
  } while (0)

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace slam
