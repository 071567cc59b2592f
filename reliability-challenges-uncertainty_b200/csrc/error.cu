#include "common.cuh"

namespace rcu {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
}  // namespace rcu

extern "C" int rcu_abi_version(void) { return RCU_ABI_VERSION; }
extern "C" const char* rcu_last_error(void) { return rcu::get_error(); }
extern "C" int rcu_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    rcu::set_error("no CUDA device available (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    cudaGetLastError();
    return RCU_ECUDA;
  }
  RCU_CHECK_ARG(device >= 0 && device < n, "device %d out of range (%d present)", device, n);
  cudaDeviceProp prop;
  RCU_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    rcu::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return RCU_ENOTSUP;
  }
  return RCU_OK;
}
