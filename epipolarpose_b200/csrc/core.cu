// Error state, version, device check.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void epb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" __attribute__((visibility("default"))) int epb_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) const char* epb_last_error(void) { return g_err; }

extern "C" __attribute__((visibility("default"))) int epb_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    epb_set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return EPB_ENOGPU;
  }
  int dev = 0, major = 0;
  EPB_CUDA(cudaGetDevice(&dev));
  EPB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    epb_set_error("device %d has compute capability %d.x; libepb.so is built for sm_100a only", dev, major);
    return EPB_ENOGPU;
  }
  return EPB_OK;
}
