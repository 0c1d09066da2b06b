// common.cuh -- context, error plumbing, pool allocation, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>

#include "../../include/p1b200.h"

namespace p1 {

void set_error(const char *fmt, ...);

#define P1_CUDA(expr)                                                         \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      p1::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,        \
                    cudaGetErrorString(e__));                                 \
      return P1_ECUDA;                                                        \
    }                                                                         \
  } while (0)

#define P1_TRY(expr)                                                          \
  do {                                                                        \
    int r__ = (expr);                                                         \
    if (r__ != P1_OK) return r__;                                             \
  } while (0)

#define P1_REQUIRE(cond, msg)                                                 \
  do {                                                                        \
    if (!(cond)) {                                                            \
      p1::set_error("%s: %s", __func__, msg);                                 \
      return P1_EINVAL;                                                       \
    }                                                                         \
  } while (0)

// device-side status word (one per context, zeroed by the calls that read it)
enum : int {
  FLAG_INDEX_RANGE = 1,    // element index outside [0, n_nodes)
  FLAG_NONMANIFOLD = 2,    // directed edge appears twice
  FLAG_BOUNDARY_MISS = 4,  // boundary edge not found among element edges
  FLAG_UNCOVERED = 8,      // boundary half-edge not listed in any boundary
  FLAG_OVERFLOW = 32,      // topology-only plan: a node has more than 8 incident elements
  FLAG_RETRY = 16,         // a row contradicted the closed-fan shortcut: re-plan exactly
  FLAG_SLOWPATH = 64,      // p1_assemble_cold: the mesh needs the plan route (nothing written)
};

}  // namespace p1

struct p1_prof;
struct p1_arena;
struct p1_cold_cache;
struct p1_ctx {
  int device;
  int sm_count;
  int *d_flags;        // 8 ints: [0] status bits, [1] max node index, [2..7] list lengths
  int *h_flags;        // pinned mirror
  p1_prof *prof;       // per-kernel CUDA-event timing, off unless enabled
  p1_arena *arena;     // cached scratch blocks
  p1_cold_cache *cold; // CUDA graphs of small cold assemblies (plan.cu), or null
  int rowptr_resident; // tiles of the fused row-pointer kernel resident at once (0: not asked yet)
  // The status words, their mirror and the profiler are one per context: every entry
  // point holds this lock from start to return (DeviceGuard), so calls on one context
  // from several host threads serialise instead of clearing each other's flags.
  // (p1_assemble_cold keeps its status in caller memory and only needs the lock for
  // the profiler.)
  std::recursive_mutex lock;
};

namespace p1 {

// Scratch memory comes from a context-owned caching arena: freed blocks are
// kept and handed out again to requests of (nearly) the same size.  The
// assembly of one mesh repeats the same request sizes every call, so after the
// first call no cudaMalloc happens on the hot path.  (cudaMallocAsync's pool
// was tried first: it splits large free blocks for small requests and then
// regrows, which showed up as 40 ms spikes every few calls.)
int arena_alloc(p1_ctx *ctx, void **ptr, size_t bytes, cudaStream_t s);
void arena_free(p1_ctx *ctx, void *ptr, cudaStream_t s);

template <typename T>
inline int pool_alloc(p1_ctx *ctx, T **ptr, size_t count, cudaStream_t s) {
  if (count == 0) count = 1;
  return arena_alloc(ctx, (void **)ptr, count * sizeof(T), s);
}
inline void pool_free(p1_ctx *ctx, void *ptr, cudaStream_t s) {
  if (ptr) arena_free(ctx, ptr, s);
}

struct DeviceGuard {
  int prev;
  std::recursive_mutex *lock = nullptr;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  // entry points: select the context's device AND serialise on the context
  explicit DeviceGuard(p1_ctx *ctx) : lock(&ctx->lock) {
    lock->lock();
    cudaGetDevice(&prev);
    if (prev != ctx->device) cudaSetDevice(ctx->device);
  }
  ~DeviceGuard() {
    cudaSetDevice(prev);
    if (lock) lock->unlock();
  }
};

inline unsigned grid_for(int64_t n, int block, int per_thread = 1) {
  int64_t per_block = (int64_t)block * per_thread;
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// Per-kernel timing with CUDA events on the launching stream (bench.py reads
// it through p1_profile_report).  No-ops unless p1_profile_enable(ctx, 1).
bool prof_enabled(const p1_ctx *ctx);
// releases the cached CUDA graphs of p1_assemble_cold and their scratch (plan.cu)
void cold_cache_clear(p1_ctx *ctx);
void prof_begin(p1_ctx *ctx, const char *name, cudaStream_t s);
void prof_end(p1_ctx *ctx, cudaStream_t s);
#define P1_TIMED(ctx, name, s, ...)                                           \
  do {                                                                        \
    p1::prof_begin((ctx), (name), (s));                                       \
    __VA_ARGS__;                                                              \
    p1::prof_end((ctx), (s));                                                 \
  } while (0)

// exclusive prefix sum of n int32 values; out may alias in; out[n] = total
// when write_total is true (out then has n+1 entries).
int exclusive_scan_i32(p1_ctx *ctx, const int *in, int *out, int64_t n,
                       bool write_total, cudaStream_t s);
// in-place exclusive scan of n block sums by one block; sums[n] = total (-1: beyond int32)
void scan_sums(int *sums, int n, cudaStream_t s);

}  // namespace p1
