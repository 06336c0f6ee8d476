// Error state, version, device check.
#include "common.cuh"
#include <string.h>
#include <mutex>
#include <vector>

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

// ---------------------------------------------------------------------------------------
// Internal scratch (per-CTA partials, coefficient tables, operand planes).  One buffer per
// (purpose, device, stream), so calls on different devices or streams never share scratch;
// it only ever GROWS, by allocating a new buffer and retiring the old one WITHOUT freeing it
// (a kernel still in flight or a captured CUDA graph may hold the old pointer), and it
// refuses to allocate while the stream is being captured (cudaMalloc is illegal there): run
// one eager call of the same shape first.  New buffers are zero-filled (stream ordered).
struct WsEntry { int kind, dev; cudaStream_t st; void* p; size_t cap; };
static std::mutex g_ws_mu;
static std::vector<WsEntry> g_ws;

int epb_workspace(int kind, size_t bytes, cudaStream_t st, void** out) {
  int dev = 0;
  EPB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_ws_mu);
  WsEntry* e = nullptr;
  for (auto& w : g_ws)
    if (w.kind == kind && w.dev == dev && w.st == st) { e = &w; break; }
  if (e && e->cap >= bytes) {
    *out = e->p;
    return EPB_OK;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  EPB_CUDA(cudaStreamIsCapturing(st, &cs));
  if (cs != cudaStreamCaptureStatusNone) {
    epb_set_error("internal scratch (kind %d) would have to grow to %zu bytes during CUDA graph "
                  "capture: run one eager step of the same shape first", kind, bytes);
    return EPB_EINVAL;
  }
  size_t cap = bytes < 4096 ? 4096 : bytes;
  void* p = nullptr;
  EPB_CUDA(cudaMalloc(&p, cap));
  EPB_CUDA(cudaMemsetAsync(p, 0, cap, st));
  if (e) { e->p = p; e->cap = cap; }             // the old buffer is retired, not freed
  else g_ws.push_back({kind, dev, st, p, cap});
  *out = p;
  return EPB_OK;
}
