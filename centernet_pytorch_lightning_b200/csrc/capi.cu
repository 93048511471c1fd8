// Library-wide pieces of the C ABI: error text, launch counter, host-buffer convenience entry.
#include "cnb_common.cuh"
#include <mutex>
#include <string.h>

namespace cnb {

static thread_local char t_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

// Device staging arena for the *_host entry points (grown on demand; belongs to the device that was current when it
// was allocated and is re-allocated when the caller has switched devices).
struct Arena {
  void* dev = nullptr;
  size_t bytes = 0;
  int device = -1;
  std::mutex mu;
  int ensure(size_t need) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != device && dev) {
      cudaSetDevice(device);
      cudaFree(dev);
      cudaSetDevice(cur);
      dev = nullptr;
      bytes = 0;
    }
    device = cur;
    if (need <= bytes) return CNB_OK;
    if (dev) cudaFree(dev);
    dev = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&dev, need);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
      return CNB_ERR_CUDA;
    }
    bytes = need;
    return CNB_OK;
  }
};
static Arena g_arena;

}  // namespace cnb

using namespace cnb;

extern "C" int cnb_version(void) { return 100; }
extern "C" const char* cnb_last_error(void) { return t_err; }
extern "C" unsigned long long cnb_launch_count(void) { return g_launches.load(); }

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int cnb_ctdet_decode_host(const float* heat, const float* wh, const float* reg, float* out,
                                     int B, int C, int H, int W, int K) {
  CNB_CHECK_ARG(heat && wh && out, "ctdet_decode_host: null pointer");
  CNB_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1 && K >= 1, "ctdet_decode_host: bad shape");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device");
    return CNB_ERR_NO_DEVICE;
  }
  const size_t hw = (size_t)H * W;
  const size_t n_heat = (size_t)B * C * hw * 4, n_2 = (size_t)B * 2 * hw * 4, n_out = (size_t)B * K * 6 * 4;
  const size_t ws = cnb_ctdet_decode_workspace_bytes(B, C, H, W, K);
  CNB_CHECK_ARG(ws > 0, "ctdet_decode_host: unsupported shape");
  std::lock_guard<std::mutex> lock(g_arena.mu);
  const size_t total = up256(n_heat) + 2 * up256(n_2) + up256(n_out) + up256(ws);
  int rc = g_arena.ensure(total);
  if (rc) return rc;
  unsigned char* p = (unsigned char*)g_arena.dev;
  float* d_heat = (float*)p; p += up256(n_heat);
  float* d_wh = (float*)p;   p += up256(n_2);
  float* d_reg = (float*)p;  p += up256(n_2);
  float* d_out = (float*)p;  p += up256(n_out);
  void* d_ws = p;
  cudaStream_t st = 0;
  CNB_CUDA(cudaMemcpyAsync(d_heat, heat, n_heat, cudaMemcpyHostToDevice, st));
  CNB_CUDA(cudaMemcpyAsync(d_wh, wh, n_2, cudaMemcpyHostToDevice, st));
  if (reg) CNB_CUDA(cudaMemcpyAsync(d_reg, reg, n_2, cudaMemcpyHostToDevice, st));
  rc = cnb_ctdet_decode(d_heat, d_wh, reg ? d_reg : nullptr, d_out, B, C, H, W, K, d_ws, ws, st);
  if (rc) return rc;
  CNB_CUDA(cudaMemcpyAsync(out, d_out, n_out, cudaMemcpyDeviceToHost, st));
  CNB_CUDA(cudaStreamSynchronize(st));
  return CNB_OK;
}
