// plan.cu -- symbolic phase: node -> incident (element, vertex) records with
// their neighbour pair, and the CSR row pointer.
//
// Replaces, for the assembly hot path, what scipy derives on every
// coo_matrix(...).tocsr() call: coo_tocsr (counting sort by row),
// csr_sort_indices (per-row sort) and csr_sum_duplicates
// (scipy/sparse/_coo.py:368-439, _compressed.py:1072-1130), but on the 3M
// (node, element) incidences instead of the 9M triplets: row i of any P1
// matrix has the columns {i} U {next, prev vertex of every element at i}.
//
// HBM traffic per triangle (int32): count 12 B read (+ node counters, L2
// atomics), scatter 12 B read + 36 B write, rowlen 24 B read + 2 B write.
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
               uint32_t *__restrict__ rec, int2 *__restrict__ nbr) {
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
        const int pos = rec_off[a] + atomicSub(cnt + a, 1) - 1;
        rec[pos] = ((uint32_t)t << 2) | (uint32_t)k;
        nbr[pos] = make_int2(e[k == 2 ? 0 : k + 1],
                             e[k == 0 ? 2 : k - 1] | (k << NODE_BITS));
      }
  }
}

// --------------------------------------------------------------- rowlen
// One thread per owned row: number of distinct values in
// {self} U {next_i} U {prev_i}.  No per-thread arrays: the records of a row
// are a few contiguous bytes that stay in L1 while they are compared.
__global__ void __launch_bounds__(THREADS)
rowlen_kernel(int n_own, int rb, const int *__restrict__ rec_off,
              const int2 *__restrict__ nbr, int *__restrict__ rowlen,
              int *__restrict__ big_rows, int *__restrict__ n_big) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_own) return;
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  if (deg == 0) { rowlen[a] = 0; return; }
  if (deg > DEG_BIG) {
    big_rows[atomicAdd(n_big, 1)] = a;
    rowlen[a] = 0;  // filled in by rowlen_big_kernel
    return;
  }
  const int self = a + rb;
  const int2 *nb = nbr + beg;
  int n = 1;
  for (int i = 0; i < deg; ++i) {
    const int x = nb[i].x, y = nbr_prev(nb[i]);
    bool fx = x != self, fy = y != self && y != x;
    for (int j = 0; j < deg; ++j) {
      const int2 q = nb[j];
      const int qx = q.x, qy = nbr_prev(q);
      // x is new unless an earlier next equals it
      fx = fx && !(j < i && qx == x);
      // y is new unless ANY next equals it or an earlier prev does
      fy = fy && qx != y && !(j < i && qy == y);
    }
    n += fx + fy;
  }
  rowlen[a] = n;
}

// Rows with more than DEG_BIG records: one block per row, O(deg^2) without
// any capacity limit (pathological fans only; never on a shape-regular mesh).
__global__ void __launch_bounds__(THREADS)
sort_big_kernel(const int *__restrict__ big_rows, const int *__restrict__ rec_off,
                uint32_t *__restrict__ rec, int2 *__restrict__ nbr,
                uint32_t *__restrict__ tmp_rec, int2 *__restrict__ tmp_nbr) {
  const int a = big_rows[blockIdx.x];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  // rank sort by (element, vertex); records of one row are distinct
  for (int i = threadIdx.x; i < deg; i += blockDim.x) {
    const uint32_t v = rec[beg + i];
    int rank = 0;
    for (int j = 0; j < deg; ++j) rank += rec[beg + j] < v;
    tmp_rec[beg + rank] = v;
    tmp_nbr[beg + rank] = nbr[beg + i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < deg; i += blockDim.x) {
    rec[beg + i] = tmp_rec[beg + i];
    nbr[beg + i] = tmp_nbr[beg + i];
  }
}

// candidate j of a big row: 0 = the row itself, 1+2i = next_i, 2+2i = prev_i
__device__ __forceinline__ int big_candidate(int self, const int2 *nb, int j) {
  if (j == 0) return self;
  const int2 p = nb[(j - 1) >> 1];
  return ((j - 1) & 1) ? nbr_prev(p) : p.x;
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

// ------------------------------------------------------ (T,k) row order
__global__ void __launch_bounds__(THREADS)
sort_rows_kernel(int n_own, const int *__restrict__ rec_off,
                 const uint32_t *__restrict__ rec, const int2 *__restrict__ nbr,
                 uint32_t *__restrict__ out_rec, int2 *__restrict__ out_nbr) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_own) return;
  const int beg = rec_off[a], end = rec_off[a + 1];
  for (int i = beg; i < end; ++i) {
    const uint32_t v = rec[i];
    int rank = 0;
    for (int j = beg; j < end; ++j) rank += rec[j] < v;
    out_rec[beg + rank] = v;
    out_nbr[beg + rank] = nbr[i];
  }
}

int ensure_sorted(p1_plan *p, cudaStream_t s) {
  if (p->sorted) return P1_OK;
  const int n_own = (int)(p->row_end - p->row_begin);
  uint32_t *r2 = nullptr;
  int2 *n2 = nullptr;
  P1_TRY(pool_alloc(&r2, (size_t)p->n_rec, s));
  P1_TRY(pool_alloc(&n2, (size_t)p->n_rec, s));
  if (n_own > 0)
    P1_TIMED(p->ctx, "sort_rows", s,
             sort_rows_kernel<<<grid_for(n_own, THREADS), THREADS, 0, s>>>(
                 n_own, p->rec_off, p->rec, p->nbr, r2, n2));
  P1_CUDA(cudaGetLastError());
  pool_free(p->rec, s);
  pool_free(p->nbr, s);
  p->rec = r2;
  p->nbr = n2;
  p->sorted = true;
  return P1_OK;
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
  if (n_nodes > (int64_t)NODE_MASK || n_elements >= (1LL << 30) ||
      3 * n_elements + n_nodes > 0x7fffffffLL) {
    set_error("p1_plan_create: mesh too large for int32 device indices "
              "(M=%lld N=%lld; limits: N < 2^29, 3M+N < 2^31)",
              (long long)n_elements, (long long)n_nodes);
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
  auto fail = [&](int code) { pool_free(cnt, s); p1_plan_destroy(p); return code; };

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
  if ((rc = pool_alloc(&p->big_rows, (size_t)(n_rec / (DEG_BIG + 1)) + 1, s)))
    return fail(rc);
  P1_TIMED(ctx, "scatter", s,
           scatter_kernel<<<eg, THREADS, 0, s>>>(elements, M, N, rb, re, cnt, p->rec_off,
                                                 p->rec, p->nbr));
  d_nbig = ctx->d_flags + 2;
  // cnt is all zero again: reuse it as rowlen
  if (n_own > 0)
    P1_TIMED(ctx, "rowlen", s,
             rowlen_kernel<<<grid_for(n_own, THREADS), THREADS, 0, s>>>(
                 n_own, rb, p->rec_off, p->nbr, cnt, p->big_rows, d_nbig));
  cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 4 * sizeof(int),
                  cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) {
    set_error("p1_plan_create: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(P1_ECUDA);
  }
  if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
    set_error("p1_plan_create: element index out of range [0, %lld)",
              (long long)n_nodes);
    return fail(P1_EINDEX);
  }
  p->max_index = ctx->h_flags[1];
  p->n_big = ctx->h_flags[2];
  if (p->n_big > 0) {
    uint32_t *tmp_rec = nullptr;
    int2 *tmp_nbr = nullptr;
    if ((rc = pool_alloc(&tmp_rec, (size_t)n_rec, s)) ||
        (rc = pool_alloc(&tmp_nbr, (size_t)n_rec, s)) ||
        (rc = pool_alloc(&p->big_first, (size_t)(2 * n_rec + n_own + 1), s))) {
      pool_free(tmp_rec, s);
      pool_free(tmp_nbr, s);
      return fail(rc);
    }
    sort_big_kernel<<<p->n_big, THREADS, 0, s>>>(p->big_rows, p->rec_off, p->rec, p->nbr,
                                                 tmp_rec, tmp_nbr);
    rowlen_big_kernel<<<p->n_big, THREADS, 0, s>>>(p->big_rows, rb, p->rec_off,
                                                   p->nbr, p->big_first, cnt);
    pool_free(tmp_rec, s);
    pool_free(tmp_nbr, s);
  }
  if ((rc = exclusive_scan_i32(ctx, cnt, p->indptr, n_own, true, s))) return fail(rc);
  int h_nnz = 0;
  cudaMemcpyAsync(&h_nnz, p->indptr + n_own, sizeof(int), cudaMemcpyDeviceToHost, s);
  pool_free(cnt, s);
  cnt = nullptr;
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
