// assemble.cu -- numeric phase.
//
// Step 1 (one thread per triangle): the coefficients every entry of the local
// 3x3 matrix derives from, computed ONCE per element with the reference's
// arithmetic (compile with -fmad=false): /root/reference/p1afempy/
// solvers.py:31-45 (stiffness: a, b, c), :176-181 (mass: area/6, area/12).
// Local matrices of the generalised stiffness / weighted mass come as 9 values
// per element (local_ops.cu).
// Step 2: the owner of a CSR row gathers the entries (k,k), (k,k+1), (k,k+2)
// of its incident elements, sums duplicates in a fixed order and writes the
// finished row (sorted unique columns, explicit zeros kept).
//
// Fast path: 8 lanes per row, one record per lane, all-pairs comparison through
// shared memory (plan.cuh: analyse_row).  HBM traffic per triangle: 12 B indices
// + gathered coordinates + 24 B coefficients (step 1); 48 B records + 24 B
// gathered coefficients + 42 B CSR output (step 2).
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 256;

// ------------------------------------------------ per-element coefficients
__global__ void __launch_bounds__(THREADS)
elem_stiffness_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                      int64_t M, int rb, int re, double *__restrict__ abc) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e0 = __ldg(elements + 3 * t), e1 = __ldg(elements + 3 * t + 1),
              e2 = __ldg(elements + 3 * t + 2);
    // only elements that touch an owned row are ever gathered (multi-GPU)
    if (!((e0 >= rb && e0 < re) || (e1 >= rb && e1 < re) || (e2 >= rb && e2 < re))) continue;
    const double2 z0 = __ldg(coords + e0);
    const double2 z1 = __ldg(coords + e1);
    const double2 z2 = __ldg(coords + e2);
    const double d21x = z1.x - z0.x, d21y = z1.y - z0.y;
    const double d31x = z2.x - z0.x, d31y = z2.y - z0.y;
    const double area4 = 4. * (0.5 * (d21x * d31y - d21y * d31x));
    abc[3 * t] = (d21x * d31x + d21y * d31y) / area4;
    abc[3 * t + 1] = (d31x * d31x + d31y * d31y) / area4;
    abc[3 * t + 2] = (d21x * d21x + d21y * d21y) / area4;
  }
}

__global__ void __launch_bounds__(THREADS)
elem_mass_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                 int64_t M, int rb, int re, double2 *__restrict__ on_off) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e0 = __ldg(elements + 3 * t), e1 = __ldg(elements + 3 * t + 1),
              e2 = __ldg(elements + 3 * t + 2);
    if (!((e0 >= rb && e0 < re) || (e1 >= rb && e1 < re) || (e2 >= rb && e2 < re))) continue;
    const double2 z0 = __ldg(coords + e0);
    const double2 z1 = __ldg(coords + e1);
    const double2 z2 = __ldg(coords + e2);
    const double d21x = z1.x - z0.x, d21y = z1.y - z0.y;
    const double d31x = z2.x - z0.x, d31y = z2.y - z0.y;
    const double area = 0.5 * (d21x * d31y - d21y * d31x);
    on_off[t] = make_double2(area / 6., area / 12.);
  }
}

// row k of the local matrix of element T: entries (k,k), (k,k+1), (k,k+2)
struct StiffnessOp {
  const double *__restrict__ abc;  // M x 3
  __device__ __forceinline__ void row(int64_t T, int k, double &vd, double &vn,
                                      double &vp) const {
    const double a = __ldg(abc + 3 * T), b = __ldg(abc + 3 * T + 1),
                 c = __ldg(abc + 3 * T + 2);
    if (k == 0) { vd = -2. * a + b + c; vn = a - b; vp = a - c; }
    else if (k == 1) { vd = b; vn = -a; vp = a - b; }
    else { vd = c; vn = a - c; vp = -a; }
  }
};

struct MassOp {
  const double2 *__restrict__ on_off;  // M: (area/6, area/12)
  __device__ __forceinline__ void row(int64_t T, int, double &vd, double &vn,
                                      double &vp) const {
    const double2 m = __ldg(on_off + T);
    vd = m.x; vn = m.y; vp = m.y;
  }
};

struct LocalOp {
  const double *__restrict__ local;  // M x 9 row-major
  __device__ __forceinline__ void row(int64_t T, int k, double &vd, double &vn,
                                      double &vp) const {
    const double *m = local + 9 * T + 3 * k;
    vd = __ldg(m + k);
    vn = __ldg(m + next_local(k));
    vp = __ldg(m + prev_local(k));
  }
};

// ------------------------------------------------------ fast path: 8 lanes/row
template <class Op>
__global__ void __launch_bounds__(THREADS)
fill_rows_kernel(int n_own, int rb, int exact,
                 const unsigned long long *__restrict__ acc,
                 const int *__restrict__ rec_off, const int4 *__restrict__ rec,
                 const int *__restrict__ indptr, Op op, int *__restrict__ indices,
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
  double vd = 0., vn = 0., vp = 0.;
  if (r.valid) op.row((int64_t)r.elem, r.k, vd, vn, vp);
  s_vp[threadIdx.x] = vp;
  if (r.valid) s_vd[gbase + r.rk] = vd;  // canonical order: ascending `next`
  __syncwarp(gmask);
  if (r.valid) {
    // column `next`: my (k,k+1) entry plus the (k,k+2) entry of the element
    // after me in the fan (two terms: order-free)
    const double off = r.mate < 0 ? vn : vn + s_vp[gbase + r.mate];
    if (indices) indices[base + r.slot_nx] = r.nx;
    data[base + r.slot_nx] = off;
    if (r.extra) {
      if (indices) indices[base + r.slot_pv] = r.pv;
      data[base + r.slot_pv] = vp;
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

template <class Op>
__global__ void __launch_bounds__(SLOW_THREADS)
fill_slow_kernel(const int *__restrict__ slow_rows, int n_slow, int rb,
                 const int *__restrict__ rec_off, const int4 *__restrict__ rec,
                 const int *__restrict__ indptr, Op op, int *__restrict__ indices,
                 double *__restrict__ data) {
  __shared__ int cols_s[SLOW_THREADS * SLOW_COLS];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_slow) return;
  const int a = slow_rows[idx];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  const int self = a + rb;
  const int4 *nb = rec + beg;
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
  // fixed summation order: ascending `next`, ties in stored order
  for (int j = 0; j < n; ++j) {
    const int c = cs[j];
    for (int i = 0; i < deg; ++i) {
      const int4 p = nb[i];
      if (rec_next(p) != c) continue;
      double vd, vn, vp;
      op.row((int64_t)rec_elem(p), rec_k(p), vd, vn, vp);
      vs[js] += vd;
      vs[j] += vn;
      vs[lower_bound(cs, n, rec_prev(p))] += vp;
    }
  }
}

// ------------------------------------------------ big path: one block per row
__device__ __forceinline__ int big_candidate2(int self, const int4 *nb, int j) {
  if (j == 0) return self;
  const int4 p = nb[(j - 1) >> 1];
  return ((j - 1) & 1) ? rec_prev(p) : rec_next(p);
}

template <class Op>
__global__ void __launch_bounds__(THREADS)
fill_big_kernel(const int *__restrict__ big_rows, int rb,
                const int *__restrict__ rec_off, const int4 *__restrict__ rec,
                const unsigned char *__restrict__ first,
                const int *__restrict__ indptr, Op op, int *__restrict__ indices,
                double *__restrict__ data) {
  const int a = big_rows[blockIdx.x];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  const int4 *nb = rec + beg;
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
    for (int i = 0; i < deg; ++i) {  // records are in (element, vertex) order
      const int4 p = nb[i];
      const int px = rec_next(p), py = rec_prev(p);
      if (c != self && px != c && py != c) continue;
      double vd, vn, vp;
      op.row((int64_t)rec_elem(p), rec_k(p), vd, vn, vp);
      if (c == self) sum += vd;
      if (px == c) sum += vn;
      if (py == c) sum += vp;
    }
    if (indices) indices[base + rank] = c;
    data[base + rank] = sum;
  }
}

template <class Op>
int launch_fill(const p1_plan *p, Op op, int32_t *indices, double *data,
                cudaStream_t s) {
  const int n_own = (int)(p->row_end - p->row_begin);
  const int rb = (int)p->row_begin;
  if (n_own > 0)
    P1_TIMED(p->ctx, "fill_rows", s,
             (fill_rows_kernel<Op><<<grid_for((int64_t)n_own * GROUP, THREADS), THREADS, 0, s>>>(
                 n_own, rb, p->exact ? 1 : 0, p->acc, p->rec_off, p->rec, p->indptr, op,
                 indices, data, p->ctx->d_flags)));
  if (p->n_slow > 0)
    P1_TIMED(p->ctx, "fill_slow", s,
             (fill_slow_kernel<Op><<<grid_for(p->n_slow, SLOW_THREADS), SLOW_THREADS, 0, s>>>(
                 p->slow_rows, p->n_slow, rb, p->rec_off, p->rec, p->indptr, op, indices,
                 data)));
  if (p->n_big > 0)
    fill_big_kernel<Op><<<p->n_big, THREADS, 0, s>>>(
        p->big_rows, rb, p->rec_off, p->rec, p->big_first, p->indptr, op, indices, data);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

}  // namespace p1

using namespace p1;

extern "C" int p1_assemble_csr(const p1_plan *plan, int op, const double *coords,
                               const double *local, int32_t *indptr,
                               int32_t *indices, double *data, void *stream) {
  P1_REQUIRE(plan, "null plan");
  P1_REQUIRE(data || plan->nnz == 0, "null data");
  p1_ctx *ctx = plan->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_own = plan->row_end - plan->row_begin;
  const int64_t M = plan->M;
  if (indptr)
    P1_CUDA(cudaMemcpyAsync(indptr, plan->indptr, (n_own + 1) * sizeof(int),
                            cudaMemcpyDeviceToDevice, s));
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), s));
  const unsigned eg = min(grid_for(M, THREADS), (unsigned)(ctx->sm_count * 32));
  int rc = P1_OK;
  double *coef = nullptr;
  switch (op) {
    case P1_OP_STIFFNESS:
      P1_REQUIRE(coords || M == 0, "stiffness needs coords");
      P1_TRY(pool_alloc(ctx, &coef, (size_t)3 * M, s));
      if (M > 0)
        P1_TIMED(ctx, "elem_coef", s,
                 elem_stiffness_kernel<<<eg, THREADS, 0, s>>>(
                     (const double2 *)coords, plan->elements, M, (int)plan->row_begin,
                     (int)plan->row_end, coef));
      rc = launch_fill(plan, StiffnessOp{coef}, indices, data, s);
      break;
    case P1_OP_MASS:
      P1_REQUIRE(coords || M == 0, "mass needs coords");
      P1_TRY(pool_alloc(ctx, &coef, (size_t)2 * M, s));
      if (M > 0)
        P1_TIMED(ctx, "elem_coef", s,
                 elem_mass_kernel<<<eg, THREADS, 0, s>>>(
                     (const double2 *)coords, plan->elements, M, (int)plan->row_begin,
                     (int)plan->row_end, (double2 *)coef));
      rc = launch_fill(plan, MassOp{(const double2 *)coef}, indices, data, s);
      break;
    case P1_OP_LOCAL:
      P1_REQUIRE(local || M == 0, "P1_OP_LOCAL needs the local matrices");
      rc = launch_fill(plan, LocalOp{local}, indices, data, s);
      break;
    default:
      set_error("p1_assemble_csr: unknown op %d", op);
      return P1_EINVAL;
  }
  pool_free(ctx, coef, s);
  P1_TRY(rc);
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
