// assemble.cu -- numeric phase: the owner of a CSR row sums the entries its
// records carry (written by the scatter, plan.cu: the local matrix of every
// element is computed ONCE, by the thread that owns the element, with the
// reference's arithmetic, plan.cuh: *Values) in a fixed order and writes the
// finished row (sorted unique columns, explicit zeros kept).
//
// Fast path: 8 lanes per row, one 32-byte record per lane (one coalesced
// 256-bit load), all-pairs comparison through shared memory (plan.cuh:
// analyse_row).  HBM traffic per triangle: 96 B records + 42 B CSR output.
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 256;

// ------------------------------------------------------ fast path: 8 lanes/row
__global__ void __launch_bounds__(THREADS)
fill_rows_kernel(int n_own, int rb, int exact,
                 const unsigned long long *__restrict__ acc,
                 const int *__restrict__ rec_off, const Rec *__restrict__ rec,
                 const int *__restrict__ indptr, int *__restrict__ indices,
                 double *__restrict__ data, int *__restrict__ flags) {
  __shared__ int2 xch[THREADS];
  __shared__ double s_vp[THREADS];
  __shared__ double s_vd[THREADS];
  const int lane = threadIdx.x & 31, lane8 = lane & (GROUP - 1);
  const int shift = lane & ~(GROUP - 1);
  const unsigned gmask = 0xffu << shift;
  const int gbase = threadIdx.x & ~(GROUP - 1);
  const int a = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GROUP);
  if (a >= n_own) return;
  // the three loads that depend on the row id only are issued together: the
  // kernel is bound by the latency of its dependent loads, not by bandwidth
  const unsigned long long v = __ldg(acc + a);
  const int beg = __ldg(rec_off + a);
  const int64_t base = __ldg(indptr + a);
  const int deg = (int)(unsigned)v;
  // group-uniform exits: empty rows; rows the slow / big kernels take
  if (deg == 0 || deg > GROUP) return;
  if (!exact && (unsigned)(v >> 32) != 0u) return;
  const int self = a + rb;
  const RowLane r = analyse_row(self, beg, deg, rec, gmask, lane8, shift, xch);
  if (!r.simple || (!exact && r.len != 1 + deg)) {
    // exact plan: the row is in slow_rows.  Hash plan: the shortcut
    // "closed fan => 1 + #records columns" does not hold for this row.
    if (!exact && lane8 == 0) atomicOr(flags, FLAG_RETRY);
    return;
  }
  s_vp[threadIdx.x] = r.vp;
  if (r.valid) s_vd[gbase + r.rk] = r.vd;  // canonical order: ascending `next`
  __syncwarp(gmask);
  if (r.valid) {
    // column `next`: my (k,k+1) entry plus the (k,k+2) entry of the element
    // after me in the fan (two terms: order-free)
    const double off = r.mate < 0 ? r.vn : r.vn + s_vp[gbase + r.mate];
    if (indices) indices[base + r.slot_nx] = r.nx;
    data[base + r.slot_nx] = off;
    if (r.extra) {
      if (indices) indices[base + r.slot_pv] = r.pv;
      data[base + r.slot_pv] = r.vp;
    }
  }
  if (lane8 == 0) {
    double diag = 0.;
    for (int q = 0; q < deg; ++q) diag += s_vd[gbase + q];
    if (indices) indices[base + r.slot_self] = self;
    data[base + r.slot_self] = diag;
  }
}

// -------------------------------------------- slow path: one thread per row
// open fans (boundary nodes), more than 8 records, repeated neighbours
// (non-manifold / degenerate elements): sorted unique columns by insertion into
// a private shared-memory slice, values accumulated in the output row itself.
constexpr int SLOW_THREADS = 64;
constexpr int SLOW_COLS = 2 * DEG_BIG + 1;

__device__ __forceinline__ int lower_bound(const int *cols, int n, int c) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cols[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int insert_unique(int *cols, int n, int c) {
  int j = n;
  while (j > 0 && cols[j - 1] > c) --j;
  if (j > 0 && cols[j - 1] == c) return n;
  for (int m = n; m > j; --m) cols[m] = cols[m - 1];
  cols[j] = c;
  return n + 1;
}

__global__ void __launch_bounds__(SLOW_THREADS)
fill_slow_kernel(const int *__restrict__ slow_rows, int n_slow, int rb,
                 const int *__restrict__ rec_off, const Rec *__restrict__ rec,
                 const int *__restrict__ indptr, int *__restrict__ indices,
                 double *__restrict__ data) {
  __shared__ int cols_s[SLOW_THREADS * SLOW_COLS];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_slow) return;
  const int a = slow_rows[idx];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  const int self = a + rb;
  const Rec *nb = rec + beg;
  int *cs = cols_s + threadIdx.x * SLOW_COLS;
  int n = insert_unique(cs, 0, self);
  for (int i = 0; i < deg; ++i) {
    n = insert_unique(cs, n, rec_next(nb[i]));
    n = insert_unique(cs, n, rec_prev(nb[i]));
  }
  const int64_t base = indptr[a];
  double *vs = data + base;
  for (int j = 0; j < n; ++j) {
    if (indices) indices[base + j] = cs[j];
    vs[j] = 0.;
  }
  const int js = lower_bound(cs, n, self);
  // fixed summation order: ascending `next`, ties by `prev`
  for (int j = 0; j < n; ++j) {
    const int c = cs[j];
    int last_pv = -1;
    for (;;) {  // records with next == c in ascending prev order
      int pick = -1, pick_pv = 0x7fffffff;
      for (int i = 0; i < deg; ++i) {
        if (rec_next(nb[i]) != c) continue;
        const int pv = rec_prev(nb[i]);
        if (pv > last_pv && pv < pick_pv) { pick = i; pick_pv = pv; }
      }
      if (pick < 0) break;
      // every record with this (next, prev) pair (duplicates: identical elements)
      for (int i = 0; i < deg; ++i) {
        if (rec_next(nb[i]) != c || rec_prev(nb[i]) != pick_pv) continue;
        const Rec p = nb[i];
        vs[js] += p.vd;
        vs[j] += p.vn;
        vs[lower_bound(cs, n, pick_pv)] += p.vp;
      }
      last_pv = pick_pv;
    }
  }
}

// ------------------------------------------------ big path: one block per row
__device__ __forceinline__ int big_candidate2(int self, const Rec *nb, int j) {
  if (j == 0) return self;
  const Rec p = nb[(j - 1) >> 1];
  return ((j - 1) & 1) ? rec_prev(p) : rec_next(p);
}

__global__ void __launch_bounds__(THREADS)
fill_big_kernel(const int *__restrict__ big_rows, int rb,
                const int *__restrict__ rec_off, const Rec *__restrict__ rec,
                const unsigned char *__restrict__ first,
                const int *__restrict__ indptr, int *__restrict__ indices,
                double *__restrict__ data) {
  const int a = big_rows[blockIdx.x];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  const Rec *nb = rec + beg;
  const unsigned char *fl = first + 2 * (int64_t)beg + a;
  const int self = a + rb;
  const int ncand = 2 * deg + 1;
  const int64_t base = indptr[a];
  for (int j = threadIdx.x; j < ncand; j += blockDim.x) {
    if (!fl[j]) continue;
    const int c = big_candidate2(self, nb, j);
    int rank = 0;
    for (int i = 0; i < ncand; ++i)
      rank += fl[i] && big_candidate2(self, nb, i) < c;
    double sum = 0.;
    for (int i = 0; i < deg; ++i) {  // records are in (next, prev) order
      const Rec p = nb[i];
      const int px = rec_next(p), py = rec_prev(p);
      if (c == self) sum += p.vd;
      if (px == c) sum += p.vn;
      if (py == c) sum += p.vp;
    }
    if (indices) indices[base + rank] = c;
    data[base + rank] = sum;
  }
}

// ------------------------------------------- re-assembly: refresh the values
template <class V>
__global__ void __launch_bounds__(THREADS)
refresh_kernel(const int *__restrict__ elements, int64_t n_rec,
               Rec *__restrict__ rec, const int *__restrict__ rec_elem, V values) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = rec_elem[i];
    const int e[3] = {__ldg(elements + 3 * t), __ldg(elements + 3 * t + 1),
                      __ldg(elements + 3 * t + 2)};
    double v[3][3];
    values.rows(t, e, v);
    Rec r = rec[i];
    const int k = rec_k(r);
    r.vd = k == 0 ? v[0][0] : (k == 1 ? v[1][0] : v[2][0]);
    r.vn = k == 0 ? v[0][1] : (k == 1 ? v[1][1] : v[2][1]);
    r.vp = k == 0 ? v[0][2] : (k == 1 ? v[1][2] : v[2][2]);
    store_rec(rec + i, r);
  }
}

}  // namespace p1

using namespace p1;

extern "C" int p1_assemble_csr(const p1_plan *cplan, int op, const double *coords,
                               const double *local, int32_t *indptr,
                               int32_t *indices, double *data, void *stream) {
  P1_REQUIRE(cplan, "null plan");
  P1_REQUIRE(data || cplan->nnz == 0, "null data");
  p1_plan *plan = const_cast<p1_plan *>(cplan);
  p1_ctx *ctx = plan->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int n_own = (int)(plan->row_end - plan->row_begin);
  const int rb = (int)plan->row_begin;
  const int64_t M = plan->M;
  if (op == P1_OP_STORED) {
    if (plan->val_op == 0) {
      set_error("p1_assemble_csr: P1_OP_STORED, but the plan was not created with "
                "p1_plan_create_for and no operator has been assembled on it yet");
      return P1_EINVAL;
    }
  } else {
    // new coefficients on an existing plan: recompute the values in place
    P1_TRY(require_elements(plan, "assembling a new operator on an existing plan"));
    const unsigned rg = min(grid_for(plan->n_rec, THREADS), (unsigned)(ctx->sm_count * 32));
    const double2 *xy = (const double2 *)coords;
    switch (op) {
      case P1_OP_STIFFNESS:
      case P1_OP_MASS:
        P1_REQUIRE(coords || M == 0, "stiffness/mass need coords");
        break;
      case P1_OP_LOCAL:
        P1_REQUIRE(local || M == 0, "P1_OP_LOCAL needs the local matrices");
        break;
      default:
        set_error("p1_assemble_csr: unknown op %d", op);
        return P1_EINVAL;
    }
    if (plan->n_rec > 0) {
      prof_begin(ctx, "refresh", s);
      if (op == P1_OP_STIFFNESS)
        refresh_kernel<<<rg, THREADS, 0, s>>>(plan->elements, plan->n_rec, plan->rec,
                                              plan->rec_elem, StiffnessValues{xy});
      else if (op == P1_OP_MASS)
        refresh_kernel<<<rg, THREADS, 0, s>>>(plan->elements, plan->n_rec, plan->rec,
                                              plan->rec_elem, MassValues{xy});
      else
        refresh_kernel<<<rg, THREADS, 0, s>>>(plan->elements, plan->n_rec, plan->rec,
                                              plan->rec_elem, LocalValues{local});
      prof_end(ctx, s);
    }
    plan->val_op = op;
  }
  if (indptr)
    P1_CUDA(cudaMemcpyAsync(indptr, plan->indptr, ((size_t)n_own + 1) * sizeof(int),
                            cudaMemcpyDeviceToDevice, s));
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), s));
  if (n_own > 0)
    P1_TIMED(ctx, "fill_rows", s,
             (fill_rows_kernel<<<grid_for((int64_t)n_own * GROUP, THREADS), THREADS, 0, s>>>(
                 n_own, rb, plan->exact ? 1 : 0, plan->acc, plan->rec_off, plan->rec,
                 plan->indptr, indices, data, ctx->d_flags)));
  if (plan->n_slow > 0)
    P1_TIMED(ctx, "fill_slow", s,
             (fill_slow_kernel<<<grid_for(plan->n_slow, SLOW_THREADS), SLOW_THREADS, 0, s>>>(
                 plan->slow_rows, plan->n_slow, rb, plan->rec_off, plan->rec, plan->indptr,
                 indices, data)));
  if (plan->n_big > 0)
    fill_big_kernel<<<plan->n_big, THREADS, 0, s>>>(
        plan->big_rows, rb, plan->rec_off, plan->rec, plan->big_first, plan->indptr,
        indices, data);
  P1_CUDA(cudaGetLastError());
  if (!plan->exact) {
    // the closed-fan shortcut is verified while filling: read the verdict
    P1_CUDA(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int),
                            cudaMemcpyDeviceToHost, s));
    P1_CUDA(cudaStreamSynchronize(s));
    if (ctx->h_flags[0] & FLAG_RETRY) {
      set_error("p1_assemble_csr: the mesh contradicts the plan's closed-fan shortcut "
                "(non-manifold vertex or degenerate element); rebuild the plan with "
                "P1_PLAN_EXACT");
      return P1_ERETRY;
    }
  }
  return P1_OK;
}
