// Error plumbing + device check of the C ABI (include/frx.h).
#include "common.cuh"

namespace frx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace frx

extern "C" {

int frx_abi_version(void) { return FRX_ABI_VERSION; }

const char* frx_last_error(void) { return frx::g_err; }

int frx_device_check(int ordinal) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    frx::set_error("no CUDA device visible (%s); libfrx_b200 has no CPU fallback",
                   e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return FRX_E_DEVICE;
  }
  if (ordinal < 0 || ordinal >= n) {
    frx::set_error("device ordinal %d out of range (0..%d)", ordinal, n - 1);
    return FRX_E_DEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, ordinal);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, ordinal);
  if (major != 10) {
    frx::set_error("device %d is sm_%d%d; libfrx_b200 is built for sm_100a only", ordinal, major, minor);
    return FRX_E_DEVICE;
  }
  return FRX_OK;
}

}  // extern "C"
