// assemble.cu -- numeric phase: the owner of a CSR row gathers the local-matrix
// rows of its incident elements, sums duplicates and writes the finished row
// (sorted unique columns, explicit zeros kept).
//
// Arithmetic mirrors the reference expression by expression (compile with
// -fmad=false): /root/reference/p1afempy/solvers.py:31-45 (stiffness),
// :176-181 (mass).  Local matrices of the generalised stiffness / weighted
// mass come precomputed per element (local_ops.cu).
//
// Fast path: 8 lanes per row, one record per lane, all-pairs comparison by
// warp shuffles (plan.cuh: analyse_row) -- no per-thread arrays, no shared
// memory, no serial sort.  HBM traffic per triangle: 48 B records + gathered
// coordinates (L1/L2 resident on locally ordered meshes) + 42 B CSR output.
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 256;

struct Tri {  // vertex coordinates in ELEMENT order (v0, v1, v2)
  double x0, y0, x1, y1, x2, y2;
};

// (self, next, prev) -> element order, given the local index k of `self`
__device__ __forceinline__ Tri element_order(int k, double2 s, double2 n, double2 p) {
  Tri t;
  if (k == 0) { t.x0 = s.x; t.y0 = s.y; t.x1 = n.x; t.y1 = n.y; t.x2 = p.x; t.y2 = p.y; }
  else if (k == 1) { t.x0 = p.x; t.y0 = p.y; t.x1 = s.x; t.y1 = s.y; t.x2 = n.x; t.y2 = n.y; }
  else { t.x0 = n.x; t.y0 = n.y; t.x1 = p.x; t.y1 = p.y; t.x2 = s.x; t.y2 = s.y; }
  return t;
}

struct StiffnessOp {
  static constexpr bool needs_coords = true;
  // row k of the local stiffness matrix: (k,k), (k,k+1), (k,k+2)
  __device__ __forceinline__ void row(int64_t, int k, const Tri &t, double &vd,
                                      double &vn, double &vp) const {
    const double d21x = t.x1 - t.x0, d21y = t.y1 - t.y0;
    const double d31x = t.x2 - t.x0, d31y = t.y2 - t.y0;
    const double area4 = 4. * (0.5 * (d21x * d31y - d21y * d31x));
    const double a = (d21x * d31x + d21y * d31y) / area4;
    const double b = (d31x * d31x + d31y * d31y) / area4;
    const double c = (d21x * d21x + d21y * d21y) / area4;
    if (k == 0) { vd = -2. * a + b + c; vn = a - b; vp = a - c; }
    else if (k == 1) { vd = b; vn = -a; vp = a - b; }
    else { vd = c; vn = a - c; vp = -a; }
  }
};

struct MassOp {
  static constexpr bool needs_coords = true;
  __device__ __forceinline__ void row(int64_t, int, const Tri &t, double &vd,
                                      double &vn, double &vp) const {
    const double d21x = t.x1 - t.x0, d21y = t.y1 - t.y0;
    const double d31x = t.x2 - t.x0, d31y = t.y2 - t.y0;
    const double area = 0.5 * (d21x * d31y - d21y * d31x);
    vd = area / 6.;
    vn = area / 12.;
    vp = vn;
  }
};

struct LocalOp {
  static constexpr bool needs_coords = false;
  const double *__restrict__ local;  // M x 9 row-major
  __device__ __forceinline__ void row(int64_t T, int k, const Tri &, double &vd,
                                      double &vn, double &vp) const {
    const double *m = local + 9 * T + 3 * k;
    vd = __ldg(m + k);
    vn = __ldg(m + next_local(k));
    vp = __ldg(m + prev_local(k));
  }
};

// ------------------------------------------------------ fast path: 8 lanes/row
template <class Op>
__global__ void __launch_bounds__(THREADS)
fill_rows_kernel(int n_own, int rb, const int *__restrict__ rec_off,
                 const int4 *__restrict__ rec, const int *__restrict__ indptr,
                 const double2 *__restrict__ coords, Op op,
                 int *__restrict__ indices, double *__restrict__ data) {
  __shared__ int2 xch[THREADS];
  __shared__ double2 s_zn[THREADS];
  __shared__ double s_vp[THREADS];
  __shared__ double s_vd[THREADS];
  const int lane = threadIdx.x & 31, lane8 = lane & (GROUP - 1);
  const int shift = lane & ~(GROUP - 1);
  const unsigned gmask = 0xffu << shift;
  const int gbase = threadIdx.x & ~(GROUP - 1);
  const int a = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GROUP);
  if (a >= n_own) return;
  const int beg = __ldg(rec_off + a), deg = __ldg(rec_off + a + 1) - beg;
  if (deg == 0 || deg > GROUP) return;  // group-uniform
  const int self = a + rb;
  const RowLane r = analyse_row(self, beg, deg, rec, gmask, lane8, shift, xch);
  if (!r.simple) return;  // group-uniform; fill_slow_kernel takes the row

  Tri t;
  if (Op::needs_coords) {
    const double2 zs = __ldg(coords + self);
    double2 zn = make_double2(0., 0.);
    if (r.valid) zn = __ldg(coords + r.nx);
    s_zn[threadIdx.x] = zn;
    __syncwarp(gmask);
    // the previous vertex is the `next` of the element before me in the fan
    double2 zp = s_zn[gbase + (r.src < 0 ? 0 : r.src)];
    if (r.extra) zp = __ldg(coords + r.pv);  // open fan end
    t = element_order(r.k, zs, zn, zp);
  }
  double vd = 0., vn = 0., vp = 0.;
  if (r.valid) op.row((int64_t)r.elem, r.k, t, vd, vn, vp);
  s_vp[threadIdx.x] = vp;
  if (r.valid) s_vd[gbase + r.rk] = vd;  // canonical order: ascending `next`
  __syncwarp(gmask);
  const int64_t base = __ldg(indptr + a);
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
// more than 8 records, or repeated neighbours (non-manifold / degenerate
// elements): sorted unique columns by insertion into a private shared-memory
// slice, values accumulated in the output row itself.
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
                 const int *__restrict__ indptr, const double2 *__restrict__ coords,
                 Op op, int *__restrict__ indices, double *__restrict__ data) {
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
  double2 zs = make_double2(0., 0.);
  if (Op::needs_coords) zs = coords[self];
  // fixed summation order: ascending `next`, ties in stored order
  for (int j = 0; j < n; ++j) {
    const int c = cs[j];
    for (int i = 0; i < deg; ++i) {
      const int4 p = nb[i];
      if (rec_next(p) != c) continue;
      const int pv = rec_prev(p), k = rec_k(p);
      Tri t;
      if (Op::needs_coords) t = element_order(k, zs, coords[c], coords[pv]);
      double vd, vn, vp;
      op.row((int64_t)rec_elem(p), k, t, vd, vn, vp);
      vs[js] += vd;
      vs[j] += vn;
      vs[lower_bound(cs, n, pv)] += vp;
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
                const int *__restrict__ indptr, const double2 *__restrict__ coords,
                Op op, int *__restrict__ indices, double *__restrict__ data) {
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
      const int k = rec_k(p);
      Tri t;
      if (Op::needs_coords)
        t = element_order(k, coords[self], coords[px], coords[py]);
      double vd, vn, vp;
      op.row((int64_t)rec_elem(p), k, t, vd, vn, vp);
      if (c == self) sum += vd;
      if (px == c) sum += vn;
      if (py == c) sum += vp;
    }
    if (indices) indices[base + rank] = c;
    data[base + rank] = sum;
  }
}

template <class Op>
int launch_fill(const p1_plan *p, Op op, const double *coords, int32_t *indices,
                double *data, cudaStream_t s) {
  const int n_own = (int)(p->row_end - p->row_begin);
  const int rb = (int)p->row_begin;
  const double2 *xy = (const double2 *)coords;
  if (n_own > 0)
    P1_TIMED(p->ctx, "fill_rows", s,
             (fill_rows_kernel<Op><<<grid_for((int64_t)n_own * GROUP, THREADS), THREADS, 0, s>>>(
                 n_own, rb, p->rec_off, p->rec, p->indptr, xy, op, indices, data)));
  if (p->n_slow > 0)
    P1_TIMED(p->ctx, "fill_slow", s,
             (fill_slow_kernel<Op><<<grid_for(p->n_slow, SLOW_THREADS), SLOW_THREADS, 0, s>>>(
                 p->slow_rows, p->n_slow, rb, p->rec_off, p->rec, p->indptr, xy, op,
                 indices, data)));
  if (p->n_big > 0)
    fill_big_kernel<Op><<<p->n_big, THREADS, 0, s>>>(
        p->big_rows, rb, p->rec_off, p->rec, p->big_first, p->indptr, xy, op, indices,
        data);
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
  DeviceGuard guard(plan->ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_own = plan->row_end - plan->row_begin;
  if (indptr)
    P1_CUDA(cudaMemcpyAsync(indptr, plan->indptr, (n_own + 1) * sizeof(int),
                            cudaMemcpyDeviceToDevice, s));
  switch (op) {
    case P1_OP_STIFFNESS:
      P1_REQUIRE(coords, "stiffness needs coords");
      return launch_fill(plan, StiffnessOp{}, coords, indices, data, s);
    case P1_OP_MASS:
      P1_REQUIRE(coords, "mass needs coords");
      return launch_fill(plan, MassOp{}, coords, indices, data, s);
    case P1_OP_LOCAL:
      P1_REQUIRE(local, "P1_OP_LOCAL needs the local matrices");
      return launch_fill(plan, LocalOp{local}, coords, indices, data, s);
    default:
      set_error("p1_assemble_csr: unknown op %d", op);
      return P1_EINVAL;
  }
}
