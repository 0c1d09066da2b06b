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
  uint32_t *rec;       // n_rec   : (T<<2)|k, arrival order until `sorted`
  int2 *nbr;           // n_rec   : x = next node, y = prev node | k<<29
  bool sorted;         // records of every row ascending in (T,k)
  int *indptr;         // n_own+1
  int *big_rows;       // local row ids with more than DEG_BIG records
  int n_big;
  unsigned char *big_first;  // scratch flags for big rows (2*n_rec+n_own) or null
  int *twin;           // lazily built: 3M, twin half-edge id T'*3+k' or -1
  int64_t n_open;      // half-edges without twin (valid once twin is built)
};

namespace p1 {

// rows with more records than this take the block-per-row path
constexpr int DEG_BIG = 40;
// node ids must leave the top 3 bits of an int free (k is packed into nbr.y)
constexpr int NODE_BITS = 29;
constexpr int NODE_MASK = (1 << NODE_BITS) - 1;

__device__ __forceinline__ int next_local(int k) { return k == 2 ? 0 : k + 1; }
__device__ __forceinline__ int prev_local(int k) { return k == 0 ? 2 : k - 1; }
__device__ __forceinline__ int nbr_prev(int2 p) { return p.y & NODE_MASK; }
__device__ __forceinline__ int nbr_k(int2 p) { return (int)((unsigned)p.y >> NODE_BITS); }

int ensure_twins(p1_plan *plan, cudaStream_t s);
// puts the records of every row into ascending (element, vertex) order: the
// order of the reference's sequential np.bincount (load vectors)
int ensure_sorted(p1_plan *plan, cudaStream_t s);

}  // namespace p1
