// Host-side helpers shared by all translation units of libatdn_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/atdn_b200.h"

namespace atdn {

// thread-local error string (atdn_last_error)
int set_error(int code, const char* fmt, ...);

#define ATDN_REQUIRE(cond, code, ...)                         \
  do {                                                        \
    if (!(cond)) return ::atdn::set_error((code), __VA_ARGS__); \
  } while (0)

#define ATDN_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return ::atdn::set_error((int)_e, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// returns 0 when the current device is sm_100 (cached per device), ATDN_ERR_ARCH otherwise
int require_sm100();

}  // namespace atdn
