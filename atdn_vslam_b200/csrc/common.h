// Host-side helpers shared by all translation units of libatdn_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

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

// One-time PER-DEVICE configuration of a launch site: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a property of
// the function on ONE device, so a process that drives several GPUs must repeat it on each of them.
//   static DeviceOnce once;  if (once.pending()) { ATDN_CUDA(cudaFuncSetAttribute(...)); once.done(); }
// Lock-free; two threads racing on the same device both configure (idempotent).
class DeviceOnce {
 public:
  bool pending() {
    dev_ = 0;
    cudaGetDevice(&dev_);
    return dev_ < 0 || dev_ >= 128 || !(mask_[dev_ >> 6].load(std::memory_order_acquire) >> (dev_ & 63) & 1u);
  }
  void done() {
    if (dev_ >= 0 && dev_ < 128) mask_[dev_ >> 6].fetch_or(uint64_t(1) << (dev_ & 63), std::memory_order_release);
  }

 private:
  static thread_local int dev_;
  std::atomic<uint64_t> mask_[2] = {};
};

// multiprocessor count of the current device (cached per device)
int num_sms();

// environment switches (A/B experiments), read ONCE per process
struct EnvSwitches {
  bool no_out_tma, b_resident, corr_no_pair;
  bool pdl;       // ATDN_PDL=1: programmatic dependent launch of the flow-net kernels (tc_ptx.cuh); off by default
  int corr_dbg;
};
const EnvSwitches& env_switches();

}  // namespace atdn
