// Error plumbing and device checks of the C ABI (include/atdn_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include "common.h"

namespace atdn {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

thread_local int DeviceOnce::dev_ = 0;

namespace {
struct DevInfo {
  std::atomic<int> state{0};   // 0 = unknown, 1 = sm_100, 2 = other
  int major = 0, minor = 0, sms = 0;
};
DevInfo g_dev[128];

DevInfo* dev_info(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 128) return nullptr;
  if (dev_out) *dev_out = dev;
  DevInfo* d = &g_dev[dev];
  if (d->state.load(std::memory_order_acquire) == 0) {
    cudaDeviceGetAttribute(&d->major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d->minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, dev);
    d->state.store((d->major == 10 && d->minor == 0) ? 1 : 2, std::memory_order_release);
  }
  return d;
}
}  // namespace

int require_sm100() {
  int dev = 0;
  DevInfo* d = dev_info(&dev);
  if (!d) return set_error(ATDN_ERR_ARCH, "cudaGetDevice failed or device index out of range");
  if (d->state.load(std::memory_order_acquire) == 1) return 0;
  return set_error(ATDN_ERR_ARCH, "device %d is sm_%d%d; libatdn_b200 contains sm_100a code only (no fallback)", dev, d->major, d->minor);
}

int num_sms() {
  DevInfo* d = dev_info(nullptr);
  return d ? d->sms : 0;
}

const EnvSwitches& env_switches() {
  static const EnvSwitches s = [] {
    auto on = [](const char* name) { const char* e = getenv(name); return e && e[0] == '1'; };
    EnvSwitches v;
    v.no_out_tma = on("ATDN_NO_OUT_TMA");
    v.b_resident = on("ATDN_B_RESIDENT");
    v.corr_no_pair = on("ATDN_CORR_NO_PAIR");
    v.pdl = on("ATDN_PDL");               // opt-in: measured neutral to slightly negative on the batched step
    const char* dbg = getenv("ATDN_CORR_DBG");
    v.corr_dbg = dbg ? atoi(dbg) : 0;
    return v;
  }();
  return s;
}

}  // namespace atdn

extern "C" const char* atdn_last_error(void) { return atdn::g_err; }
extern "C" int atdn_version(void) { return 100; }
extern "C" int atdn_check_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return atdn::set_error((int)e, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  return atdn::require_sm100();
}

// ------------------------------------------------------------------------------------------------
// Host-side pose chain (no device work): the sequential tail of the path after ONE device->host copy of all
// relative poses.  fp32 arithmetic in the operation order of atdn_vslam/utils/transforms.py:54-119 and
// slam_framework/neural_slam.py:204-215, 288-302 (this translation unit is compiled without FMA contraction).
// ------------------------------------------------------------------------------------------------
static void mat4_mul(const float* a, const float* b, float* o) {
  float t[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = a[i * 4] * b[j];   // sequential fused multiply-adds: the order of the CPU sgemm behind torch.matmul
      for (int k = 1; k < 4; ++k) acc = fmaf(a[i * 4 + k], b[k * 4 + j], acc);
      t[i * 4 + j] = acc;
    }
  for (int i = 0; i < 16; ++i) o[i] = t[i];
}

extern "C" int atdn_pose_chain(const float* rot, const float* cos_sin, const float* tr, int64_t num, float rot_threshold_rad,
                               float tr_threshold, float* poses, int32_t* is_key) {
  ATDN_REQUIRE(rot && tr && poses && is_key && num >= 0, ATDN_ERR_ARG, "atdn_pose_chain: null argument");
  static const float eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float cur[16], prop[16];
  for (int i = 0; i < 16; ++i) cur[i] = prop[i] = poses[i] = eye[i];
  is_key[0] = 1;                                  // frame 0 is always registered (neural_slam.py:218-225)
  for (int64_t t = 0; t < num; ++t) {
    const float* r = rot + 3 * t;
    float c1, c2, c3, s1, s2, s3;
    if (cos_sin) {   // caller-computed cos / sin (bit-identical to the caller's math library, e.g. torch's)
      const float* cs = cos_sin + 6 * t;
      c1 = cs[0]; c2 = cs[1]; c3 = cs[2]; s1 = cs[3]; s2 = cs[4]; s3 = cs[5];
    } else {
      c1 = cosf(r[0]); c2 = cosf(r[1]); c3 = cosf(r[2]);
      s1 = sinf(r[0]); s2 = sinf(r[1]); s3 = sinf(r[2]);
    }
    float m[16] = {c1 * c3 + s1 * s2 * s3, c3 * s1 * s2 - c1 * s3, c2 * s1, tr[3 * t],
                   c2 * s3, c2 * c3, -s2, tr[3 * t + 1],
                   c1 * s2 * s3 - c3 * s1, c1 * c3 * s2 + s1 * s3, c1 * c2, tr[3 * t + 2],
                   0, 0, 0, 1};
    mat4_mul(cur, m, cur);
    mat4_mul(prop, m, prop);
    const float alpha = atan2f(prop[2], prop[10]);
    const float beta = atan2f(-prop[6], sqrtf(1.0f - prop[6] * prop[6]));
    const float gamma = atan2f(prop[4], prop[5]);
    const float nr = sqrtf(alpha * alpha + beta * beta + gamma * gamma);
    const float nt = sqrtf(prop[3] * prop[3] + prop[7] * prop[7] + prop[11] * prop[11]);
    const int key = (nr > rot_threshold_rad) || (nt > tr_threshold);
    if (key)
      for (int i = 0; i < 16; ++i) prop[i] = eye[i];
    is_key[t + 1] = key;
    for (int i = 0; i < 16; ++i) poses[16 * (t + 1) + i] = cur[i];
  }
  return 0;
}
