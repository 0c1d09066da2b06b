// scan.cu -- exclusive prefix sum of int32 counters (reduce / scan-sums /
// downsweep).  Traffic: 2 reads + 1 write of the array; the arrays scanned on
// the hot path are per-node counters (4 B per node), a few per cent of the
// bytes the assembly moves, so the simple three-kernel form is kept.
#include "common.cuh"

namespace p1 {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_inclusive(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns the exclusive
// prefix, *total receives the block sum (valid in every thread).
template <int THREADS>
__device__ __forceinline__ int block_exclusive(int v, int *total) {
  __shared__ int warp_sums[THREADS / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = warp_inclusive(v, lane);
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < THREADS / 32 ? warp_sums[lane] : 0;
    int winc = warp_inclusive(w, lane);
    if (lane < THREADS / 32) warp_sums[lane] = winc - w;
    if (lane == THREADS / 32 - 1) block_total = winc;
  }
  __syncthreads();
  int res = inc - v + warp_sums[warp];
  *total = block_total;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const int *__restrict__ in, int64_t n,
                   int *__restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) sum += in[idx];
  }
  int total;
  block_exclusive<SCAN_THREADS>(sum, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
scan_sums_kernel(int *__restrict__ block_sums, int nb) {
  // the grand total is tracked in 64 bits: a sum beyond INT32_MAX (a row pointer that
  // int32 cannot hold) is reported as -1 instead of wrapping silently
  long long carry = 0;
  for (int start = 0; start < nb; start += 1024) {
    int idx = start + threadIdx.x;
    int v = idx < nb ? block_sums[idx] : 0;
    int total;
    int ex = block_exclusive<1024>(v, &total);
    if (idx < nb) block_sums[idx] = (int)carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) block_sums[nb] = carry > 0x7fffffffLL ? -1 : (int)carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_down_kernel(const int *__restrict__ in, int *__restrict__ out, int64_t n,
                 const int *__restrict__ block_sums, int write_total) {
  // thread t owns SCAN_ITEMS consecutive items so that its partial prefix is
  // a plain running sum
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE +
                       (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    sum += v[i];
  }
  int total;
  int ex = block_exclusive<SCAN_THREADS>(sum, &total);
  int run = block_sums[blockIdx.x] + ex;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
  }
  if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
    out[n] = block_sums[gridDim.x];
}

void scan_sums(int *sums, int n, cudaStream_t s) { scan_sums_kernel<<<1, 1024, 0, s>>>(sums, n); }

int exclusive_scan_i32(p1_ctx *ctx, const int *in, int *out, int64_t n,
                       bool write_total, cudaStream_t s) {
  if (n < 0) return P1_EINVAL;
  const unsigned nb = grid_for(n, SCAN_TILE);
  int *sums = nullptr;
  P1_TRY(pool_alloc(ctx, &sums, (size_t)nb + 1, s));
  P1_TIMED(ctx, "scan", s, {
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, n, sums);
    scan_sums_kernel<<<1, 1024, 0, s>>>(sums, (int)nb);
    scan_down_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, out, n, sums,
                                                 write_total ? 1 : 0);
  });
  P1_CUDA(cudaGetLastError());
  pool_free(ctx, sums, s);
  return P1_OK;
}

}  // namespace p1
