// Error checking and launch-size helpers shared by the CUDA translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace tapes {

#define TAPES_CUDA_CHECK(expr)                                                                  \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " \
                               + __FILE__ + ":" + std::to_string(__LINE__));                    \
  } while (0)

inline unsigned grid_for(uint64_t items, unsigned block) {
  uint64_t g = (items + block - 1) / block;
  if (g == 0) g = 1;
  if (g > 0x7fffffffull) throw std::runtime_error("grid too large");
  return (unsigned)g;
}

}  // namespace tapes
