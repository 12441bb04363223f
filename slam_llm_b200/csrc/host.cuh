// Host-side helpers shared by every translation unit of libslam_b200.so.
#pragma once
#include <cstdlib>
#include <utility>
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

// Every kernel of the library is launched with programmatic stream serialization (PDL): it may start while its predecessor
// in the stream is still draining, runs `pdl_trigger(); pdl_wait();` first (common.cuh) and so only overlaps its launch latency
// and block scheduling with the predecessor's tail - ~1300 launches per training step make the inter-kernel bubbles matter.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool enabled = []() {          // SLAM_PDL=0: plain stream order (A/B measurements)
    const char* e = getenv("SLAM_PDL");
    return e == nullptr || e[0] != '0';
  }();
  cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace slam
