// api.cu -- context lifetime and error text of the C ABI (include/p1b200.h).
#include "common.cuh"
#include <mutex>
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

struct p1_arena {
  struct Block { void *ptr; size_t bytes; bool used; cudaStream_t last; unsigned long long tick; };
  std::vector<Block> blocks;
  std::mutex lock;
  unsigned long long clock = 0;
  // unused blocks beyond this many bytes go back to the driver, least recently used
  // first (not largest first: that evicts the record block the next call of a loop asks
  // for again -- 34 GB at 256M triangles, re-allocated every step); torch's allocator
  // cannot reclaim what sits here (p1_ctx_set_cache_limit)
  size_t limit = (size_t)48 << 30;
};

namespace p1 {

int arena_alloc(p1_ctx *ctx, void **ptr, size_t bytes, cudaStream_t s) {
  *ptr = nullptr;
  bytes = (bytes + 511) & ~(size_t)511;
  p1_arena *ar = ctx->arena;
  {
    std::lock_guard<std::mutex> g(ar->lock);
    int best = -1;
    for (int i = 0; i < (int)ar->blocks.size(); ++i) {
      auto &b = ar->blocks[i];
      if (b.used || b.bytes < bytes || b.bytes > bytes + bytes / 8 + 4096) continue;
      if (best < 0 || b.bytes < ar->blocks[best].bytes) best = i;
    }
    if (best >= 0) {
      auto &b = ar->blocks[best];
      // work queued on another stream may still read the block
      if (b.last != s) cudaStreamSynchronize(b.last);
      b.used = true;
      b.last = s;
      b.tick = ++ar->clock;
      *ptr = b.ptr;
      return P1_OK;
    }
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    // give cached blocks back and try once more
    {
      std::lock_guard<std::mutex> g(ar->lock);
      cudaDeviceSynchronize();
      for (auto it = ar->blocks.begin(); it != ar->blocks.end();)
        if (!it->used) { cudaFree(it->ptr); it = ar->blocks.erase(it); } else ++it;
    }
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    set_error("out of device memory: %zu bytes: %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return P1_ECUDA;
  }
  std::lock_guard<std::mutex> g(ar->lock);
  ar->blocks.push_back({p, bytes, true, s, ++ar->clock});
  *ptr = p;
  return P1_OK;
}

void arena_free(p1_ctx *ctx, void *ptr, cudaStream_t s) {
  p1_arena *ar = ctx->arena;
  std::lock_guard<std::mutex> g(ar->lock);
  for (auto &b : ar->blocks)
    if (b.ptr == ptr) { b.used = false; b.last = s; b.tick = ++ar->clock; break; }
  size_t idle = 0;
  for (auto &b : ar->blocks)
    if (!b.used) idle += b.bytes;
  while (idle > ar->limit) {
    int big = -1;  // (the least recently used idle block)
    for (int i = 0; i < (int)ar->blocks.size(); ++i)
      if (!ar->blocks[i].used && (big < 0 || ar->blocks[i].tick < ar->blocks[big].tick)) big = i;
    if (big < 0) break;
    cudaStreamSynchronize(ar->blocks[big].last);  // queued work may still use it
    cudaFree(ar->blocks[big].ptr);
    idle -= ar->blocks[big].bytes;
    ar->blocks.erase(ar->blocks.begin() + big);
  }
}

bool prof_enabled(const p1_ctx *ctx) { return ctx && ctx->prof && ctx->prof->on; }

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
  DeviceGuard guard(ctx);
  if (!ctx->prof) ctx->prof = new p1_prof();
  ctx->prof->drain();
  ctx->prof->on = on != 0;
  return P1_OK;
}

extern "C" int p1_profile_reset(p1_ctx *ctx) {
  P1_REQUIRE(ctx, "null ctx");
  DeviceGuard guard(ctx);
  if (ctx->prof) { ctx->prof->drain(); ctx->prof->totals.clear(); }
  return P1_OK;
}

// text: one line per timed span, "name total_ms launches\n"
extern "C" int p1_profile_report(p1_ctx *ctx, char *buf, int64_t cap) {
  P1_REQUIRE(ctx && buf && cap > 0, "bad buffer");
  DeviceGuard guard(ctx);
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

// returns every cached, unused scratch block to the driver
extern "C" int p1_ctx_trim(p1_ctx *ctx) {
  P1_REQUIRE(ctx, "null ctx");
  DeviceGuard guard(ctx);
  cold_cache_clear(ctx);
  std::lock_guard<std::mutex> g(ctx->arena->lock);
  cudaDeviceSynchronize();
  auto &bl = ctx->arena->blocks;
  for (auto it = bl.begin(); it != bl.end();)
    if (!it->used) { cudaFree(it->ptr); it = bl.erase(it); } else ++it;
  return P1_OK;
}

extern "C" int p1_ctx_set_cache_limit(p1_ctx *ctx, int64_t bytes) {
  P1_REQUIRE(ctx && bytes >= 0, "bad argument");
  std::lock_guard<std::mutex> g(ctx->arena->lock);
  ctx->arena->limit = (size_t)bytes;
  return P1_OK;
}

extern "C" int p1_version(void) { return 200; }  // 0.2.0

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
  p1_ctx *ctx = new p1_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->d_flags = nullptr;
  ctx->h_flags = nullptr;
  ctx->prof = nullptr;
  ctx->cold = nullptr;
  ctx->rowptr_resident = 0;
  ctx->arena = new p1_arena();
  if (cudaMalloc((void **)&ctx->d_flags, 8 * sizeof(int)) != cudaSuccess ||
      cudaMallocHost((void **)&ctx->h_flags, 8 * sizeof(int)) != cudaSuccess) {
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
  DeviceGuard guard(ctx->device);  // (not the locking form: the lock dies with the context)
  if (ctx->d_flags) cudaFree(ctx->d_flags);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->prof) { ctx->prof->drain(); delete ctx->prof; }
  cold_cache_clear(ctx);
  if (ctx->arena) {
    cudaDeviceSynchronize();
    for (auto &b : ctx->arena->blocks) cudaFree(b.ptr);
    delete ctx->arena;
  }
  delete ctx;
  return P1_OK;
}
