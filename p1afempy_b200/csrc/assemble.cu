// assemble.cu -- numeric phase: one owner per CSR row gathers the local-matrix
// rows of its incident elements (fixed order: element, vertex), sums duplicates
// and writes the finished row (sorted unique columns, explicit zeros kept).
//
// Arithmetic mirrors the reference expression by expression (compile with
// -fmad=false): /root/reference/p1afempy/solvers.py:31-45 (stiffness),
// :176-181 (mass).  Local matrices of the generalised stiffness / weighted
// mass come precomputed per element (local_ops.cu).
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 128;

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

// ------------------------------------------------------------ fill rows
// One block = ROWS consecutive rows = one contiguous span of the CSR arrays.
// Every thread builds its row (sorted unique columns by insertion, values
// accumulated in place) inside a shared-memory image of that span; the block
// then streams the image out with fully coalesced stores.  HBM traffic per
// triangle: 24 B neighbour pairs (+12 B records for P1_OP_LOCAL) + gathered
// coordinates (L1/L2 resident on locally ordered meshes) + 42 B CSR output.
constexpr int ROWS = THREADS;        // rows per block
constexpr int STAGE_CAP = ROWS * 20; // CSR entries staged per pass

__device__ __forceinline__ int lower_bound(const int *cols, int n, int c) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cols[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// sorted insert without duplicates; returns the new length
__device__ __forceinline__ int insert_unique(int *cols, int n, int c) {
  int j = n;
  while (j > 0 && cols[j - 1] > c) --j;
  if (j > 0 && cols[j - 1] == c) return n;
  for (int m = n; m > j; --m) cols[m] = cols[m - 1];
  cols[j] = c;
  return n + 1;
}

template <class Op>
__device__ __forceinline__ void build_row(int self, const int2 *__restrict__ nb,
                                          const uint32_t *__restrict__ rec, int deg,
                                          const double2 *__restrict__ coords, const Op &op,
                                          int *cs, double *vs, int len) {
  int n = insert_unique(cs, 0, self);
  for (int i = 0; i < deg; ++i) {
    const int2 p = nb[i];
    n = insert_unique(cs, n, p.x);
    n = insert_unique(cs, n, nbr_prev(p));
  }
  for (int j = 0; j < len; ++j) vs[j] = 0.;
  const int js = lower_bound(cs, len, self);
  double2 zs = make_double2(0., 0.);
  if (Op::needs_coords) zs = __ldg(coords + self);
  // canonical summation order: records by ascending `next` (the order of the
  // columns), independent of the order the scatter happened to produce
  for (int j = 0; j < len; ++j) {
    const int c = cs[j];
    for (int i = 0; i < deg; ++i) {
      const int2 p = nb[i];
      if (p.x != c) continue;
      const int pv = nbr_prev(p), k = nbr_k(p);
      Tri t;
      if (Op::needs_coords)
        t = element_order(k, zs, __ldg(coords + c), __ldg(coords + pv));
      double vd, vn, vp;
      op.row(Op::needs_coords ? 0 : (int64_t)(rec[i] >> 2), k, t, vd, vn, vp);
      vs[js] += vd;
      vs[j] += vn;
      vs[lower_bound(cs, len, pv)] += vp;
    }
  }
}

template <class Op>
__global__ void __launch_bounds__(THREADS)
fill_rows_kernel(int n_own, int rb, const int *__restrict__ rec_off,
                 const uint32_t *__restrict__ rec, const int2 *__restrict__ nbr,
                 const int *__restrict__ indptr, const double2 *__restrict__ coords,
                 Op op, int *__restrict__ indices, double *__restrict__ data) {
  __shared__ int cols_s[STAGE_CAP];
  __shared__ double vals_s[STAGE_CAP];
  __shared__ int chunk_rows;
  const int row0 = blockIdx.x * ROWS;
  const int nrows = min(ROWS, n_own - row0);
  const int tid = threadIdx.x;
  const int a = row0 + tid;
  int my_off = 0, my_len = 0, beg = 0, deg = 0;
  if (tid < nrows) {
    my_off = indptr[a];
    my_len = indptr[a + 1] - my_off;
    beg = rec_off[a];
    deg = rec_off[a + 1] - beg;
  }
  int first = 0;  // first row of the current pass (block-uniform)
  while (first < nrows) {
    const int base = indptr[row0 + first];
    int last = nrows;  // one pass in the common case
    if (indptr[row0 + nrows] - base > STAGE_CAP) {
      __syncthreads();
      if (tid == 0) {
        int r = first;
        while (r < nrows && indptr[row0 + r + 1] - base <= STAGE_CAP) ++r;
        chunk_rows = r == first ? first + 1 : r;  // an oversize row passes alone
      }
      __syncthreads();
      last = chunk_rows;
    }
    const int span = indptr[row0 + last] - base;
    const bool staged = span <= STAGE_CAP;  // false only for a lone big row
    if (staged && tid >= first && tid < last && deg > 0 && deg <= DEG_BIG)
      build_row(a + rb, nbr + beg, rec + beg, deg, coords, op,
                cols_s + (my_off - base), vals_s + (my_off - base), my_len);
    __syncthreads();
    if (staged) {
      for (int i = tid; i < span; i += THREADS) {
        if (indices) indices[(int64_t)base + i] = cols_s[i];
        data[(int64_t)base + i] = vals_s[i];
      }
    }
    __syncthreads();
    first = last;
  }
}

__device__ __forceinline__ int big_candidate2(int self, const int2 *nb, int j) {
  if (j == 0) return self;
  const int2 p = nb[(j - 1) >> 1];
  return ((j - 1) & 1) ? nbr_prev(p) : p.x;
}

// one block per row with more than DEG_BIG records (see plan.cu); runs AFTER
// fill_rows_kernel and overwrites whatever that kernel staged for these rows
template <class Op>
__global__ void __launch_bounds__(THREADS)
fill_big_kernel(const int *__restrict__ big_rows, int rb,
                const int *__restrict__ rec_off, const uint32_t *__restrict__ rec,
                const int2 *__restrict__ nbr, const unsigned char *__restrict__ first,
                const int *__restrict__ indptr, const double2 *__restrict__ coords,
                Op op, int *__restrict__ indices, double *__restrict__ data) {
  const int a = big_rows[blockIdx.x];
  const int beg = rec_off[a], deg = rec_off[a + 1] - beg;
  const int2 *nb = nbr + beg;
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
      const int2 p = nb[i];
      const int px = p.x, py = nbr_prev(p);
      if (c != self && px != c && py != c) continue;
      const int k = nbr_k(p);
      Tri t;
      if (Op::needs_coords)
        t = element_order(k, coords[self], coords[px], coords[py]);
      double vd, vn, vp;
      op.row((int64_t)(rec[beg + i] >> 2), k, t, vd, vn, vp);
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
  if (n_own > 0)
    P1_TIMED(p->ctx, "fill_rows", s,
             (fill_rows_kernel<Op><<<grid_for(n_own, ROWS), THREADS, 0, s>>>(
                 n_own, (int)p->row_begin, p->rec_off, p->rec, p->nbr, p->indptr,
                 (const double2 *)coords, op, indices, data)));
  if (p->n_big > 0)
    fill_big_kernel<Op><<<p->n_big, THREADS, 0, s>>>(
        p->big_rows, (int)p->row_begin, p->rec_off, p->rec, p->nbr, p->big_first,
        p->indptr, (const double2 *)coords, op, indices, data);
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
