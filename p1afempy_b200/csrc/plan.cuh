// plan.cuh -- the p1_plan object and the 8-lane row analysis shared by the
// symbolic (row length) and numeric (fill) kernels.
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
  int4 *rec;           // n_rec   : x = next node, y = prev node | k<<29, z = element,
                       //           w = 0; arrival order within a row until `sorted`
  bool sorted;         // records of every row ascending in (element, vertex)
  bool exact;          // row lengths from the all-pairs kernel (no hash shortcut)
  unsigned long long *acc;  // n_own: low 32 bits = #records, high 32 bits = sum over the
                       // records of hash(next) - hash(prev): 0 for a closed fan
  int *indptr;         // n_own+1
  int *slow_rows;      // rows the 8-lane kernels do not take: more than 8 records
                       // or repeated neighbours (non-manifold / degenerate)
  int n_slow;
  int *big_rows;       // rows with more than DEG_BIG records
  int n_big;
  unsigned char *big_first;  // scratch flags for big rows (2*n_rec+n_own) or null
  int *twin;           // lazily built: 3M, twin half-edge id T'*3+k' or -1
  int64_t n_open;      // half-edges without twin (valid once twin is built)
};

namespace p1 {

// 32-bit mixer (lowbias32) for the closed-fan test
__host__ __device__ __forceinline__ unsigned mix32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

constexpr int GROUP = 8;      // lanes per row in the fast kernels = max records
constexpr int DEG_BIG = 40;   // rows with more records take the block-per-row path
// node ids must leave the top 3 bits of an int free (k is packed into rec.y)
constexpr int NODE_BITS = 29;
constexpr int NODE_MASK = (1 << NODE_BITS) - 1;

__device__ __forceinline__ int next_local(int k) { return k == 2 ? 0 : k + 1; }
__device__ __forceinline__ int prev_local(int k) { return k == 0 ? 2 : k - 1; }
__device__ __forceinline__ int rec_next(const int4 &r) { return r.x; }
__device__ __forceinline__ int rec_prev(const int4 &r) { return r.y & NODE_MASK; }
__device__ __forceinline__ int rec_k(const int4 &r) { return (int)((unsigned)r.y >> NODE_BITS); }
__device__ __forceinline__ int rec_elem(const int4 &r) { return r.z; }

// What one lane of an 8-lane group knows about "its" record and its row after
// the all-pairs comparison (shuffles only).  A row is `simple` when it has at
// most 8 records, all `next` distinct, all `prev` distinct and none equal to
// the row itself -- every row of an edge-manifold mesh of valence <= 8.
struct RowLane {
  bool valid;     // this lane holds a record
  bool simple;    // group-uniform
  int nx, pv, k, elem;
  int rk;         // rank of nx among the row's nexts = canonical record order
  int src;        // lane whose next == my prev (the element before me), or -1
  int mate;       // lane whose prev == my next (the element after me), or -1
  bool extra;     // my prev is a column no `next` provides (open fan end)
  int slot_nx, slot_pv, slot_self;  // positions in the sorted row
  int len;        // number of distinct columns
};

// `xch` is one int2 slot per thread of the block (shared memory); lanes of a
// group exchange (next, prev) through it.  (Shuffles would do, but 16 of them
// per lane saturate the ADU pipe -- measured 93 % -- while broadcast LDS is
// one wavefront per warp.)
__device__ __forceinline__ RowLane analyse_row(int self, int beg, int deg,
                                               const int4 *__restrict__ rec,
                                               unsigned gmask, int lane8, int shift,
                                               int2 *xch) {
  RowLane o;
  o.valid = lane8 < deg && deg <= GROUP;
  int4 r = make_int4(-1, -2, 0, 0);
  if (o.valid) r = __ldg(rec + beg + lane8);
  o.nx = rec_next(r);
  o.pv = o.valid ? rec_prev(r) : -2;
  o.k = rec_k(r);
  o.elem = rec_elem(r);
  int2 *grp = xch + (threadIdx.x & ~(GROUP - 1));
  grp[lane8] = make_int2(o.nx, o.pv);
  __syncwarp(gmask);
  bool ok = deg <= GROUP;
  int rk = 0, src = -1, mate = -1, below_pv = 0, below_self = 0;
#pragma unroll
  for (int j = 0; j < GROUP; ++j) {
    const int2 q = grp[j];
    if (j < deg) {
      if (q.x == self || q.y == self) ok = false;
      if (o.valid) {
        if (j != lane8 && (q.x == o.nx || q.y == o.pv)) ok = false;
        rk += q.x < o.nx;
        below_pv += q.x < o.pv;
        if (q.x == o.pv) src = j;
        if (q.y == o.nx) mate = j;
      }
      below_self += q.x < self;
    }
  }
  o.simple = (__ballot_sync(gmask, !ok) & gmask) == 0;
  o.rk = rk; o.src = src; o.mate = mate;
  o.extra = o.valid && src < 0;
  const unsigned emask = (__ballot_sync(gmask, o.extra) & gmask) >> shift;
  o.slot_nx = (self < o.nx) + rk;
  o.slot_pv = (self < o.pv) + below_pv;
  o.slot_self = below_self;
  if (emask) {  // open fan: the extra prev columns shift everything above them
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      const int qp = grp[j].y;
      if ((emask >> j) & 1) {
        o.slot_nx += qp < o.nx;
        o.slot_pv += qp < o.pv;
        o.slot_self += qp < self;
      }
    }
  }
  o.len = 1 + deg + __popc(emask);
  return o;
}

int ensure_twins(p1_plan *plan, cudaStream_t s);
// puts the records of every row into ascending (element, vertex) order: the
// order of the reference's sequential np.bincount (load vectors)
int ensure_sorted(p1_plan *plan, cudaStream_t s);

}  // namespace p1
