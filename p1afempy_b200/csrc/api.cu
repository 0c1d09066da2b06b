// api.cu -- context lifetime and error text of the C ABI (include/p1b200.h).
#include "common.cuh"
#include <vector>

namespace p1 {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace p1

struct p1_prof {
  struct Span { const char *name; cudaEvent_t a, b; };
  struct Total { const char *name; double ms; long launches; };
  bool on = false;
  std::vector<Span> pending;
  std::vector<cudaEvent_t> spare;
  std::vector<Total> totals;
  cudaEvent_t get() {
    if (!spare.empty()) { cudaEvent_t e = spare.back(); spare.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void drain() {
    for (auto &sp : pending) {
      float ms = 0.f;
      cudaEventSynchronize(sp.b);
      cudaEventElapsedTime(&ms, sp.a, sp.b);
      bool hit = false;
      for (auto &t : totals)
        if (strcmp(t.name, sp.name) == 0) { t.ms += ms; t.launches++; hit = true; break; }
      if (!hit) totals.push_back({sp.name, ms, 1});
      spare.push_back(sp.a);
      spare.push_back(sp.b);
    }
    pending.clear();
  }
};

namespace p1 {
void prof_begin(p1_ctx *ctx, const char *name, cudaStream_t s) {
  if (!ctx || !ctx->prof || !ctx->prof->on) return;
  p1_prof::Span sp{name, ctx->prof->get(), ctx->prof->get()};
  cudaEventRecord(sp.a, s);
  ctx->prof->pending.push_back(sp);
}
void prof_end(p1_ctx *ctx, cudaStream_t s) {
  if (!ctx || !ctx->prof || !ctx->prof->on || ctx->prof->pending.empty()) return;
  cudaEventRecord(ctx->prof->pending.back().b, s);
}
}  // namespace p1

using namespace p1;

extern "C" int p1_profile_enable(p1_ctx *ctx, int on) {
  P1_REQUIRE(ctx, "null ctx");
  DeviceGuard guard(ctx->device);
  if (!ctx->prof) ctx->prof = new p1_prof();
  ctx->prof->drain();
  ctx->prof->on = on != 0;
  return P1_OK;
}

extern "C" int p1_profile_reset(p1_ctx *ctx) {
  P1_REQUIRE(ctx, "null ctx");
  DeviceGuard guard(ctx->device);
  if (ctx->prof) { ctx->prof->drain(); ctx->prof->totals.clear(); }
  return P1_OK;
}

// text: one line per timed span, "name total_ms launches\n"
extern "C" int p1_profile_report(p1_ctx *ctx, char *buf, int64_t cap) {
  P1_REQUIRE(ctx && buf && cap > 0, "bad buffer");
  DeviceGuard guard(ctx->device);
  buf[0] = 0;
  if (!ctx->prof) return P1_OK;
  ctx->prof->drain();
  int64_t used = 0;
  for (auto &t : ctx->prof->totals) {
    int n = snprintf(buf + used, (size_t)(cap - used), "%s %.6f %ld\n", t.name, t.ms,
                     t.launches);
    if (n < 0 || used + n >= cap) break;
    used += n;
  }
  return P1_OK;
}

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
  ctx->prof = nullptr;
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
  if (ctx->prof) { ctx->prof->drain(); delete ctx->prof; }
  delete ctx;
  return P1_OK;
}
