// api.cu -- context lifetime and error text of the C ABI (include/p1b200.h).
#include "common.cuh"

namespace p1 {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace p1

using namespace p1;

extern "C" int p1_version(void) { return 100; }  // 0.1.0

extern "C" const char *p1_last_error(void) { return g_error; }

extern "C" int p1_ctx_create(int device, p1_ctx **out) {
  P1_REQUIRE(out, "null out");
  *out = nullptr;
  int count = 0;
  P1_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) {
    set_error("p1_ctx_create: device %d not present (%d CUDA devices)", device, count);
    return P1_EINVAL;
  }
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  P1_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("p1_ctx_create: libp1b200 is built for sm_100a (B200) only; device %d is "
              "sm_%d%d", device, prop.major, prop.minor);
    return P1_EINVAL;
  }
  // scratch comes from the device's default stream-ordered pool; keep freed
  // blocks cached instead of returning them to the OS at every sync
  cudaMemPool_t pool;
  P1_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = UINT64_MAX;
  P1_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  p1_ctx *ctx = new p1_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->d_flags = nullptr;
  ctx->h_flags = nullptr;
  if (cudaMalloc((void **)&ctx->d_flags, 4 * sizeof(int)) != cudaSuccess ||
      cudaMallocHost((void **)&ctx->h_flags, 4 * sizeof(int)) != cudaSuccess) {
    set_error("p1_ctx_create: allocation failed: %s",
              cudaGetErrorString(cudaGetLastError()));
    p1_ctx_destroy(ctx);
    return P1_ECUDA;
  }
  *out = ctx;
  return P1_OK;
}

extern "C" int p1_ctx_destroy(p1_ctx *ctx) {
  if (!ctx) return P1_OK;
  DeviceGuard guard(ctx->device);
  if (ctx->d_flags) cudaFree(ctx->d_flags);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  delete ctx;
  return P1_OK;
}
