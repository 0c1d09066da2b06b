// plan.cuh -- the p1_plan object and helpers shared by the row-owner kernels.
#pragma once
#include "common.cuh"

struct p1_plan {
  p1_ctx *ctx;
  cudaStream_t stream;      // the stream the plan was built on (frees are ordered on it)
  const int32_t *elements;  // borrowed, M x 3 (must outlive the plan)
  int64_t M, N;
  int64_t row_begin, row_end;  // owned rows
  int64_t n_rec;               // incident (element, vertex) records of owned rows
  int64_t nnz;
  int max_index;
  int *rec_off;        // n_own+1 : node -> first record
  uint32_t *rec;       // n_rec   : (T<<2)|k ascending per node
  int2 *nbr;           // n_rec   : (next, prev) node of the record
  int *indptr;         // n_own+1
  int *big_rows;       // local row ids with more than DEG_MAX records
  int n_big;
  unsigned char *big_first;  // scratch flags for big rows (2*n_rec+n_own) or null
  int *twin;           // lazily built: 3M, twin half-edge id T'*3+k' or -1
  int64_t n_open;      // half-edges without twin (valid once twin is built)
};

namespace p1 {

constexpr int DEG_MAX = 16;               // records per row on the fast path
constexpr int COL_MAX = 2 * DEG_MAX + 1;  // distinct columns a fast row can have

__device__ __forceinline__ int next_local(int k) { return k == 2 ? 0 : k + 1; }
__device__ __forceinline__ int prev_local(int k) { return k == 0 ? 2 : k - 1; }

// Sorted, duplicate-free column list of one row: {self} U {next_i} U {prev_i}.
// `cols` must hold 2*deg+1 ints.  Returns the number of distinct columns.
__device__ __forceinline__ int unique_columns(int self, const int2 *nb, int deg,
                                              int *cols) {
  int n = 0;
  cols[n++] = self;
  for (int i = 0; i < deg; ++i) {
    const int2 p = nb[i];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = h == 0 ? p.x : p.y;
      // sorted insert, skipping duplicates
      int j = n;
      while (j > 0 && cols[j - 1] > c) --j;
      if (j > 0 && cols[j - 1] == c) continue;
      for (int m = n; m > j; --m) cols[m] = cols[m - 1];
      cols[j] = c;
      ++n;
    }
  }
  return n;
}

__device__ __forceinline__ int find_column(const int *cols, int n, int c) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cols[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

int ensure_twins(p1_plan *plan, cudaStream_t s);

}  // namespace p1
