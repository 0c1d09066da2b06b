// plan.cu -- symbolic phase: node -> incident (element, vertex) lists in the
// fixed order (element, vertex), neighbour pairs, CSR row pointer.
//
// Replaces, for the assembly hot path, what scipy derives on every
// coo_matrix(...).tocsr() call: coo_tocsr (counting sort by row),
// csr_sort_indices (per-row sort) and csr_sum_duplicates
// (scipy/sparse/_coo.py:368-439, _compressed.py:1072-1130), but on the 3M
// (node, element) incidences instead of the 9M triplets: row i of any P1
// matrix has the columns {i} U {next, prev vertex of every element at i}.
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 256;

// ---------------------------------------------------------------- count
__global__ void __launch_bounds__(THREADS)
count_kernel(const int *__restrict__ elements, int64_t M, int N, int rb, int re,
             int *__restrict__ cnt, int *__restrict__ flags) {
  int mx = -1;
  bool bad = false;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    int e[3];
    e[0] = __ldg(elements + 3 * t);
    e[1] = __ldg(elements + 3 * t + 1);
    e[2] = __ldg(elements + 3 * t + 2);
    if ((unsigned)e[0] >= (unsigned)N || (unsigned)e[1] >= (unsigned)N ||
        (unsigned)e[2] >= (unsigned)N) {
      bad = true;
      continue;
    }
    mx = max(mx, max(e[0], max(e[1], e[2])));
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (e[k] >= rb && e[k] < re) atomicAdd(cnt + (e[k] - rb), 1);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx >= 0) atomicMax(flags + 1, mx);
  if (bad) atomicOr(flags, FLAG_INDEX_RANGE);
}

// -------------------------------------------------------------- scatter
// cnt holds the per-node counts on entry and zeros on exit.
__global__ void __launch_bounds__(THREADS)
scatter_kernel(const int *__restrict__ elements, int64_t M, int N, int rb, int re,
               int *__restrict__ cnt, const int *__restrict__ rec_off,
               uint32_t *__restrict__ rec) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    int e[3];
    e[0] = __ldg(elements + 3 * t);
    e[1] = __ldg(elements + 3 * t + 1);
    e[2] = __ldg(elements + 3 * t + 2);
    if ((unsigned)e[0] >= (unsigned)N || (unsigned)e[1] >= (unsigned)N ||
        (unsigned)e[2] >= (unsigned)N)
      continue;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (e[k] >= rb && e[k] < re) {
        const int a = e[k] - rb;
        const int slot = atomicSub(cnt + a, 1) - 1;
        rec[rec_off[a] + slot] = ((uint32_t)t << 2) | (uint32_t)k;
      }
  }
}

// --------------------------------------------------------------- rowlen
// One thread per owned row: put its records into the fixed order, look the
// neighbours up once, count distinct columns.
__global__ void __launch_bounds__(THREADS)
rowlen_kernel(const int *__restrict__ elements, int n_own, int rb,
              const int *__restrict__ rec_off, uint32_t *__restrict__ rec,
              int2 *__restrict__ nbr, int *__restrict__ rowlen,
              int *__restrict__ big_rows, int *__restrict__ n_big) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_own) return;
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  if (deg == 0) { rowlen[a] = 0; return; }
  if (deg > DEG_MAX) {
    big_rows[atomicAdd(n_big, 1)] = a;
    rowlen[a] = 0;  // filled in by rowlen_big_kernel
    return;
  }
  uint32_t r[DEG_MAX];
  for (int i = 0; i < deg; ++i) {  // insertion sort while loading
    const uint32_t v = rec[beg + i];
    int j = i;
    while (j > 0 && r[j - 1] > v) { r[j] = r[j - 1]; --j; }
    r[j] = v;
  }
  int2 nb[DEG_MAX];
  for (int i = 0; i < deg; ++i) {
    rec[beg + i] = r[i];
    const int64_t T = r[i] >> 2;
    const int k = r[i] & 3;
    const int nx = __ldg(elements + 3 * T + next_local(k));
    const int pv = __ldg(elements + 3 * T + prev_local(k));
    nb[i] = make_int2(nx, pv);
    nbr[beg + i] = nb[i];
  }
  int cols[COL_MAX];
  rowlen[a] = unique_columns(a + rb, nb, deg, cols);
}

// Rows with more than DEG_MAX records: one block per row, O(deg^2) without
// any capacity limit (pathological fans only; never on a shape-regular mesh).
__global__ void __launch_bounds__(THREADS)
sort_big_kernel(const int *__restrict__ elements, const int *__restrict__ big_rows,
                const int *__restrict__ rec_off, uint32_t *__restrict__ rec,
                uint32_t *__restrict__ tmp, int2 *__restrict__ nbr) {
  const int a = big_rows[blockIdx.x];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  // rank sort (records of one row are distinct)
  for (int i = threadIdx.x; i < deg; i += blockDim.x) {
    const uint32_t v = rec[beg + i];
    int rank = 0;
    for (int j = 0; j < deg; ++j) rank += rec[beg + j] < v;
    tmp[beg + rank] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < deg; i += blockDim.x) {
    const uint32_t v = tmp[beg + i];
    rec[beg + i] = v;
    const int64_t T = v >> 2;
    const int k = v & 3;
    nbr[beg + i] = make_int2(__ldg(elements + 3 * T + next_local(k)),
                             __ldg(elements + 3 * T + prev_local(k)));
  }
}

// candidate j of a big row: 0 = the row itself, 1+2i = next_i, 2+2i = prev_i
__device__ __forceinline__ int big_candidate(int self, const int2 *nb, int j) {
  if (j == 0) return self;
  const int2 p = nb[(j - 1) >> 1];
  return ((j - 1) & 1) ? p.y : p.x;
}

__global__ void __launch_bounds__(THREADS)
rowlen_big_kernel(const int *__restrict__ big_rows, int rb,
                  const int *__restrict__ rec_off, const int2 *__restrict__ nbr,
                  unsigned char *__restrict__ first, int *__restrict__ rowlen) {
  const int a = big_rows[blockIdx.x];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  const int2 *nb = nbr + beg;
  unsigned char *fl = first + 2 * (int64_t)beg + a;
  const int ncand = 2 * deg + 1;
  int mine = 0;
  for (int j = threadIdx.x; j < ncand; j += blockDim.x) {
    const int c = big_candidate(a + rb, nb, j);
    bool is_first = true;
    for (int i = 0; i < j && is_first; ++i)
      is_first = big_candidate(a + rb, nb, i) != c;
    fl[j] = is_first;
    mine += is_first;
  }
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  atomicAdd(&total, mine);
  __syncthreads();
  if (threadIdx.x == 0) rowlen[a] = total;
}

}  // namespace p1

using namespace p1;

extern "C" int p1_plan_create(p1_ctx *ctx, const int32_t *elements,
                              int64_t n_elements, int64_t n_nodes,
                              int64_t row_begin, int64_t row_end, void *stream,
                              p1_plan **out) {
  P1_REQUIRE(ctx && out, "null ctx/out");
  P1_REQUIRE(n_elements >= 0 && n_nodes >= 0, "negative size");
  P1_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n_nodes,
             "row range outside [0, n_nodes]");
  if (n_nodes > 0x7fffffffLL || n_elements >= (1LL << 30) ||
      3 * n_elements + n_nodes > 0x7fffffffLL) {
    set_error("p1_plan_create: mesh too large for int32 device indices "
              "(M=%lld N=%lld)", (long long)n_elements, (long long)n_nodes);
    return P1_ERANGE;
  }
  P1_REQUIRE(elements || n_elements == 0, "null elements");
  DeviceGuard guard(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int n_own = (int)(row_end - row_begin);
  const int rb = (int)row_begin, re = (int)row_end;
  const int N = (int)n_nodes;
  const int64_t M = n_elements;

  p1_plan *p = new p1_plan();
  memset(p, 0, sizeof(*p));
  p->ctx = ctx; p->stream = s; p->elements = elements; p->M = M; p->N = n_nodes;
  p->row_begin = row_begin; p->row_end = row_end;
  *out = nullptr;

  int rc = P1_OK;
  int *cnt = nullptr, *d_nbig = nullptr;
  uint32_t *tmp = nullptr;
  auto fail = [&](int code) { p1_plan_destroy(p); return code; };

  // flags: [0] status, [1] max index, [2] n_big
  if (cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), s) != cudaSuccess ||
      cudaMemsetAsync(ctx->d_flags + 1, 0xff, sizeof(int), s) != cudaSuccess) {
    set_error("p1_plan_create: memset failed");
    return fail(P1_ECUDA);
  }
  if ((rc = pool_alloc(&p->rec_off, (size_t)n_own + 1, s))) return fail(rc);
  if ((rc = pool_alloc(&p->indptr, (size_t)n_own + 1, s))) return fail(rc);
  if ((rc = pool_alloc(&cnt, (size_t)n_own + 1, s))) return fail(rc);
  cudaMemsetAsync(cnt, 0, ((size_t)n_own + 1) * sizeof(int), s);

  const unsigned eg = min(grid_for(M, THREADS), (unsigned)(ctx->sm_count * 32));
  P1_TIMED(ctx, "count", s,
           count_kernel<<<eg, THREADS, 0, s>>>(elements, M, N, rb, re, cnt, ctx->d_flags));
  if ((rc = exclusive_scan_i32(ctx, cnt, p->rec_off, n_own, true, s))) return fail(rc);
  // n_rec is needed on the host to size rec/nbr; for the whole matrix it is
  // 3M, for a row range it is read back
  int64_t n_rec = 3 * M;
  if (!(rb == 0 && re == N)) {
    int h = 0;
    cudaMemcpyAsync(&h, p->rec_off + n_own, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      set_error("p1_plan_create: sync failed: %s",
                cudaGetErrorString(cudaGetLastError()));
      return fail(P1_ECUDA);
    }
    n_rec = h;
  }
  p->n_rec = n_rec;
  if ((rc = pool_alloc(&p->rec, (size_t)n_rec, s))) return fail(rc);
  if ((rc = pool_alloc(&p->nbr, (size_t)n_rec, s))) return fail(rc);
  if ((rc = pool_alloc(&p->big_rows, (size_t)(n_rec / (DEG_MAX + 1)) + 1, s)))
    return fail(rc);
  P1_TIMED(ctx, "scatter", s,
           scatter_kernel<<<eg, THREADS, 0, s>>>(elements, M, N, rb, re, cnt, p->rec_off,
                                                 p->rec));
  d_nbig = ctx->d_flags + 2;
  // cnt is all zero again: reuse it as rowlen
  P1_TIMED(ctx, "rowlen", s,
           rowlen_kernel<<<grid_for(n_own, THREADS), THREADS, 0, s>>>(
               elements, n_own, rb, p->rec_off, p->rec, p->nbr, cnt, p->big_rows, d_nbig));
  cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 4 * sizeof(int),
                  cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) {
    set_error("p1_plan_create: %s", cudaGetErrorString(cudaGetLastError()));
    pool_free(cnt, s);
    return fail(P1_ECUDA);
  }
  if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
    set_error("p1_plan_create: element index out of range [0, %lld)",
              (long long)n_nodes);
    pool_free(cnt, s);
    return fail(P1_EINDEX);
  }
  p->max_index = ctx->h_flags[1];
  p->n_big = ctx->h_flags[2];
  if (p->n_big > 0) {
    if ((rc = pool_alloc(&tmp, (size_t)n_rec, s)) ||
        (rc = pool_alloc(&p->big_first, (size_t)(2 * n_rec + n_own + 1), s))) {
      pool_free(cnt, s);
      return fail(rc);
    }
    sort_big_kernel<<<p->n_big, THREADS, 0, s>>>(elements, p->big_rows, p->rec_off,
                                                 p->rec, tmp, p->nbr);
    rowlen_big_kernel<<<p->n_big, THREADS, 0, s>>>(p->big_rows, rb, p->rec_off,
                                                   p->nbr, p->big_first, cnt);
    pool_free(tmp, s);
  }
  if ((rc = exclusive_scan_i32(ctx, cnt, p->indptr, n_own, true, s))) {
    pool_free(cnt, s);
    return fail(rc);
  }
  int h_nnz = 0;
  cudaMemcpyAsync(&h_nnz, p->indptr + n_own, sizeof(int), cudaMemcpyDeviceToHost, s);
  pool_free(cnt, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("p1_plan_create: %s", cudaGetErrorString(e));
    return fail(P1_ECUDA);
  }
  p->nnz = h_nnz;
  *out = p;
  return P1_OK;
}

extern "C" int p1_plan_destroy(p1_plan *p) {
  if (!p) return P1_OK;
  DeviceGuard guard(p->ctx->device);
  void *blocks[] = {p->rec_off, p->rec, p->nbr, p->indptr, p->big_rows,
                    p->big_first, p->twin};
  for (void *b : blocks) pool_free(b, p->stream);
  delete p;
  return P1_OK;
}

extern "C" int p1_plan_info(const p1_plan *p, int64_t *nnz, int64_t *max_index,
                            int64_t *n_elements, int64_t *n_nodes) {
  P1_REQUIRE(p, "null plan");
  if (nnz) *nnz = p->nnz;
  if (max_index) *max_index = p->max_index;
  if (n_elements) *n_elements = p->M;
  if (n_nodes) *n_nodes = p->N;
  return P1_OK;
}
