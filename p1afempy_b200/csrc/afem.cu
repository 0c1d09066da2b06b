// afem.cu -- the callers either side of the assembly path:
//   * element -> neighbours (mesh.py:495-528) from the half-edge twins;
//   * newest vertex bisection (refinement.py:8-31, 508-718): edge marking with
//     closure, numbering of the new nodes in edge order, midpoints and embedded
//     values, refined boundaries, children of every element -- the reference's
//     numbering bit for bit (edge numbering itself: p1_geometric_data);
//   * the consumer side of solve_laplace (solvers.py:601-697): Neumann load,
//     y = A x on the device CSR (Dirichlet lift, energy), and a Jacobi-preconditioned
//     conjugate gradient on the free nodes, so that the matrix never leaves the GPU.
#include "plan.cuh"
#include <vector>

namespace p1 {

constexpr int AT = 256;
static inline unsigned agrid(const p1_ctx *ctx, int64_t n) {
  return min(grid_for(n, AT), (unsigned)(ctx->sm_count * 32));
}

// ------------------------------------------------------------ neighbours
__global__ void __launch_bounds__(AT)
neighbours_kernel(const int *__restrict__ twin, int64_t n, long long *__restrict__ out) {
  for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < n;
       h += (int64_t)gridDim.x * blockDim.x) {
    const int t = twin[h];
    out[h] = t >= 0 ? t / 3 : -1;
  }
}

// ------------------------------------------------------------------ NVB
// refinement.py:575-576: every edge of a marked element gets a new node
__global__ void __launch_bounds__(AT)
nvb_mark_kernel(const long long *__restrict__ marked, int64_t n_marked, int64_t M,
                const long long *__restrict__ e2e, int *__restrict__ flag, int *__restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_marked;
       i += (int64_t)gridDim.x * blockDim.x) {
    const long long t = marked[i];
    if (t < 0 || t >= M) { *bad = 1; continue; }
#pragma unroll
    for (int k = 0; k < 3; ++k) flag[e2e[3 * t + k]] = 1;
  }
}
// refinement.py:581-592: a marked edge forces the reference edge of its elements
__global__ void __launch_bounds__(AT)
nvb_closure_kernel(int64_t M, const long long *__restrict__ e2e, int *__restrict__ flag,
                   int *__restrict__ changed) {
  bool any = false;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    const long long r = e2e[3 * t];
    if (!flag[r] && (flag[e2e[3 * t + 1]] || flag[e2e[3 * t + 2]])) {
      flag[r] = 1;
      any = true;
    }
  }
  if (any) *changed = 1;
}
// children per element (refinement.py:625-637) and split flags per boundary edge
__global__ void __launch_bounds__(AT)
nvb_count_kernel(int64_t M, const long long *__restrict__ e2e, const int *__restrict__ flag,
                 int *__restrict__ n_child, int64_t K, const long long *__restrict__ b2e,
                 int *__restrict__ b_split) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < max(M, K);
       t += (int64_t)gridDim.x * blockDim.x) {
    if (t < M) {
      const int f0 = flag[e2e[3 * t]], f1 = flag[e2e[3 * t + 1]], f2 = flag[e2e[3 * t + 2]];
      n_child[t] = f0 ? 2 + f1 + f2 : 1;
    }
    if (t < K) b_split[t] = flag[b2e[t]] != 0;
  }
}
// new nodes: midpoints and embedded values (refinement.py:8-31); new_id: exclusive
// scan of the edge flags = the reference's numbering of the new nodes by edge index
__global__ void __launch_bounds__(AT)
nvb_nodes_kernel(int64_t E, const int *__restrict__ flag, const int *__restrict__ new_id,
                 const long long *__restrict__ e2n, const double2 *__restrict__ coords,
                 const double *__restrict__ embed, int64_t N, double2 *__restrict__ out_coords,
                 double *__restrict__ out_embed) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < E;
       g += (int64_t)gridDim.x * blockDim.x) {
    if (!flag[g]) continue;
    const long long a = e2n[2 * g], b = e2n[2 * g + 1];
    const double2 za = coords[a], zb = coords[b];
    const int64_t id = N + new_id[g];
    out_coords[id] = make_double2((za.x + zb.x) / 2., (za.y + zb.y) / 2.);
    if (embed) out_embed[id] = (embed[a] + embed[b]) / 2.;
  }
}
// children (refinement.py:639-716; sorted: the layout of sort_for_longest_egde=True)
__global__ void __launch_bounds__(AT)
nvb_children_kernel(int64_t M, const int *__restrict__ elements,
                    const long long *__restrict__ e2e, const int *__restrict__ flag,
                    const int *__restrict__ new_id, const int *__restrict__ first, int64_t N,
                    int sorted, int *__restrict__ out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int v0 = elements[3 * t], v1 = elements[3 * t + 1], v2 = elements[3 * t + 2];
    const long long g0 = e2e[3 * t], g1 = e2e[3 * t + 1], g2 = e2e[3 * t + 2];
    const bool s0 = flag[g0], s1 = flag[g1], s2 = flag[g2];
    const int m0 = (int)(N + new_id[g0]), m1 = (int)(N + new_id[g1]), m2 = (int)(N + new_id[g2]);
    int *o = out + 3 * (int64_t)first[t];
    auto put = [&](int j, int a, int b, int c) { o[3 * j] = a; o[3 * j + 1] = b; o[3 * j + 2] = c; };
    if (!s0) {
      put(0, v0, v1, v2);
    } else if (!s1 && !s2) {
      put(0, v2, v0, m0); put(1, v1, v2, m0);
    } else if (s1 && !s2) {
      put(0, v2, v0, m0); put(1, m0, v1, m1); put(2, v2, m0, m1);
    } else if (!s1 && s2) {
      put(0, m0, v2, m2); put(1, v0, m0, m2); put(2, v1, v2, m0);
    } else if (sorted) {
      put(0, v0, m0, m2); put(1, m0, v1, m1); put(2, m2, m1, v2); put(3, m1, m2, m0);
    } else {
      put(0, m0, v2, m2); put(1, v0, m0, m2); put(2, m0, v1, m1); put(3, v2, m0, m1);
    }
  }
}
// refined boundary j (refinement.py:600-613): the unsplit edges in order, then
// (b0, new) of the split ones, then (new, b1); before[r] = split edges before r (all
// boundaries concatenated), lo / hi: the boundary's rows, s_lo: before[lo]
__global__ void __launch_bounds__(AT)
nvb_boundary_kernel(int64_t lo, int64_t hi, int s_lo, int n_split,
                    const int *__restrict__ bnd, const long long *__restrict__ b2e,
                    const int *__restrict__ flag, const int *__restrict__ before,
                    const int *__restrict__ new_id, int64_t N, int *__restrict__ out) {
  const int64_t r = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= hi) return;
  const int b0 = bnd[2 * r], b1 = bnd[2 * r + 1];
  const int s = before[r] - s_lo;        // split edges of this boundary before r
  const int64_t keep = (hi - lo) - n_split;
  if (!flag[b2e[r]]) {
    const int64_t p = (r - lo) - s;
    out[2 * p] = b0; out[2 * p + 1] = b1;
  } else {
    const int mid = (int)(N + new_id[b2e[r]]);
    out[2 * (keep + s)] = b0; out[2 * (keep + s) + 1] = mid;
    out[2 * (keep + n_split + s)] = mid; out[2 * (keep + n_split + s) + 1] = b1;
  }
}
// refinement.py:556-567: rotate every element so that its first edge is the longest
// (argmax: the first maximum)
__global__ void __launch_bounds__(AT)
nvb_longest_kernel(int64_t M, const double2 *__restrict__ coords, int *__restrict__ elements) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int v[3] = {elements[3 * t], elements[3 * t + 1], elements[3 * t + 2]};
    double len[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double2 a = coords[v[k]], b = coords[v[k == 2 ? 0 : k + 1]];
      const double dx = b.x - a.x, dy = b.y - a.y;
      len[k] = dx * dx + dy * dy;
    }
    int best = 0;
    if (len[1] > len[best]) best = 1;
    if (len[2] > len[best]) best = 2;
    if (best == 1) { elements[3 * t] = v[1]; elements[3 * t + 1] = v[2]; elements[3 * t + 2] = v[0]; }
    if (best == 2) { elements[3 * t] = v[2]; elements[3 * t + 1] = v[0]; elements[3 * t + 2] = v[1]; }
  }
}

// --------------------------------------------------------- solve_laplace
// solvers.py:601-618: bincount over neumann.flatten('F') -- the first endpoints in
// edge order, then the second ones.  phase 0 / 1 are two launches, so a node that is
// the first endpoint of one edge and the second of another adds in that order.
__global__ void __launch_bounds__(AT)
neumann_kernel(int phase, const int *__restrict__ edges, int64_t K,
               const double2 *__restrict__ coords, const double *__restrict__ g_mid, int N,
               double *__restrict__ acc, int *__restrict__ bad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= K) return;
  const int a = edges[2 * j], b = edges[2 * j + 1];
  if ((unsigned)a >= (unsigned)N || (unsigned)b >= (unsigned)N) { *bad = 1; return; }
  const double2 za = coords[a], zb = coords[b];
  const double dx = zb.x - za.x, dy = zb.y - za.y;
  const double share = sqrt(dx * dx + dy * dy) * g_mid[j] / 2.;
  atomicAdd(acc + (phase == 0 ? a : b), share);
}
__global__ void __launch_bounds__(AT)
add_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
           double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

// y = A x (+ y0 scaled): 4 lanes per row (a P1 row has about 7 entries); the partial
// sums of the lanes are added in a fixed order
constexpr int LANES = 4;
__global__ void __launch_bounds__(AT)
spmv_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices,
            const double *__restrict__ data, const double *__restrict__ x,
            const unsigned char *__restrict__ fixed, double alpha, const double *__restrict__ y0,
            double *__restrict__ y) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int lane = threadIdx.x & (LANES - 1);
  double sum = 0.;
  if (row < n_rows && !(fixed && fixed[row])) {
    const int end = indptr[row + 1];
    for (int p = indptr[row] + lane; p < end; p += LANES) sum += data[p] * x[indices[p]];
  }
#pragma unroll
  for (int d = LANES / 2; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d, LANES);
  if (row < n_rows && lane == 0) y[row] = (y0 ? y0[row] : 0.) + alpha * sum;
}
__global__ void __launch_bounds__(AT)
diag_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices,
            const double *__restrict__ data, double *__restrict__ diag) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  double d = 0.;
  for (int p = indptr[row]; p < indptr[row + 1]; ++p)
    if (indices[p] == row) d = data[p];
  diag[row] = d;
}

// dot products in a fixed tree: DOT_BLOCKS partial sums, then one block
constexpr int DOT_BLOCKS = 1024;
__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)blockDim.x / 32; ++w) t += sh[w];
  return t;  // valid in thread 0
}
// scal: [0] rz, [1] pq, [2] rz_new, [3] rr
// mode 0: partial[b] = sum a*b over the block's slice
__global__ void __launch_bounds__(AT)
dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
           double *__restrict__ partial) {
  __shared__ double sh[AT / 32];
  double v = 0.;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = min(n, lo + per);
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) v += a[i] * b[i];
  const double t = block_sum(v, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void __launch_bounds__(AT)
dot_final_kernel(int nb, const double *__restrict__ partial, double *__restrict__ out) {
  __shared__ double sh[AT / 32];
  double v = 0.;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) v += partial[i];
  const double t = block_sum(v, sh);
  if (threadIdx.x == 0) *out = t;
}
// x += alpha p, r -= alpha q, z = r / diag; partial sums of r.z and r.r (alpha = rz / pq)
__global__ void __launch_bounds__(AT)
cg_update_kernel(int64_t n, const double *__restrict__ scal, const double *__restrict__ p,
                 const double *__restrict__ q, const double *__restrict__ diag,
                 const unsigned char *__restrict__ fixed, double *__restrict__ x,
                 double *__restrict__ r, double *__restrict__ z, double *__restrict__ partial) {
  __shared__ double sh[AT / 32];
  const double alpha = scal[1] != 0. ? scal[0] / scal[1] : 0.;
  double rz = 0., rr = 0.;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = min(n, lo + per);
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    if (fixed[i]) continue;
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    const double zi = diag[i] != 0. ? ri / diag[i] : ri;
    z[i] = zi;
    rz += ri * zi;
    rr += ri * ri;
  }
  const double t = block_sum(rz, sh);
  __syncthreads();
  const double u = block_sum(rr, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x] = t; partial[DOT_BLOCKS + blockIdx.x] = u; }
}
// rz_new, rr from the partial sums; beta = rz_new / rz; rz := rz_new
__global__ void __launch_bounds__(AT)
cg_scalars_kernel(int nb, const double *__restrict__ partial, double *__restrict__ scal) {
  __shared__ double sh[AT / 32];
  double a = 0., b = 0.;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) { a += partial[i]; b += partial[DOT_BLOCKS + i]; }
  const double t = block_sum(a, sh);
  __syncthreads();
  const double u = block_sum(b, sh);
  if (threadIdx.x == 0) {
    scal[4] = scal[0] != 0. ? t / scal[0] : 0.;  // beta
    scal[0] = t;
    scal[3] = u;
  }
}
__global__ void __launch_bounds__(AT)
cg_direction_kernel(int64_t n, const double *__restrict__ scal, const double *__restrict__ z,
                    const unsigned char *__restrict__ fixed, double *__restrict__ p) {
  const double beta = scal[4];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fixed[i] ? 0. : z[i] + beta * p[i];
}
// r = fixed ? 0 : b - A x (x carries the Dirichlet values); z = r / diag; p = z
__global__ void __launch_bounds__(AT)
cg_init_kernel(int64_t n, const double *__restrict__ b, const double *__restrict__ ax,
               const double *__restrict__ diag, const unsigned char *__restrict__ fixed,
               double *__restrict__ r, double *__restrict__ z, double *__restrict__ p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = fixed[i] ? 0. : b[i] - ax[i];
    const double zi = diag[i] != 0. ? ri / diag[i] : ri;
    r[i] = ri; z[i] = zi; p[i] = zi;
  }
}

}  // namespace p1

using namespace p1;

extern "C" int p1_element_neighbours(const p1_plan *cplan, int64_t *neighbours, void *stream) {
  P1_REQUIRE(cplan && (neighbours || cplan->M == 0), "null argument");
  p1_plan *p = const_cast<p1_plan *>(cplan);
  p1_ctx *ctx = p->ctx;
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  P1_TRY(ensure_twins(p, s));
  if (p->M > 0)
    neighbours_kernel<<<agrid(ctx, 3 * p->M), AT, 0, s>>>(p->twin, 3 * p->M,
                                                          (long long *)neighbours);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_nvb_sort_longest_edge(p1_ctx *ctx, const double *coords, int32_t *elements,
                                        int64_t n_elements, void *stream) {
  P1_REQUIRE(ctx && (n_elements == 0 || (coords && elements)), "null argument");
  DeviceGuard guard(ctx);
  if (n_elements > 0)
    nvb_longest_kernel<<<agrid(ctx, n_elements), AT, 0, (cudaStream_t)stream>>>(
        n_elements, (const double2 *)coords, elements);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

// Stage 1 of refineNVB: which edges get a new node (marked elements + closure), the
// new nodes' numbers, where every element's children start.  Scratch is the caller's:
//   edge_flag (E), edge_new (E + 1), child_first (M + 1), bnd_before (K + 1)  int32.
// counts_host: [0] new nodes, [1] elements of the refined mesh, [2 + j] split edges of
// boundary j.  SYNCHRONISES.
extern "C" int p1_nvb_plan(p1_ctx *ctx, int64_t n_elements, int64_t n_edges,
                           const int64_t *element2edges, const int64_t *marked,
                           int64_t n_marked, const int64_t *boundary2edges,
                           const int64_t *boundary_offsets_host, int n_boundaries,
                           int32_t *edge_flag, int32_t *edge_new, int32_t *child_first,
                           int32_t *bnd_before, int64_t *counts_host, void *stream) {
  P1_REQUIRE(ctx && counts_host && n_elements >= 0 && n_edges >= 0 && n_marked >= 0 &&
                 n_boundaries >= 0, "bad argument");
  P1_REQUIRE(n_elements == 0 || (element2edges && edge_flag && edge_new && child_first),
             "null array");
  P1_REQUIRE(n_marked == 0 || marked, "null marked");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t M = n_elements, E = n_edges;
  const int64_t K = n_boundaries ? boundary_offsets_host[n_boundaries] : 0;
  P1_REQUIRE(K == 0 || (boundary2edges && bnd_before), "null boundary arrays");
  int *d = ctx->d_flags, *h = ctx->h_flags;
  P1_CUDA(cudaMemsetAsync(d, 0, 8 * sizeof(int), s));
  if (E > 0) P1_CUDA(cudaMemsetAsync(edge_flag, 0, (size_t)E * sizeof(int), s));
  if (n_marked > 0)
    nvb_mark_kernel<<<agrid(ctx, n_marked), AT, 0, s>>>((const long long *)marked, n_marked, M,
                                                        (const long long *)element2edges,
                                                        edge_flag, d + 1);
  for (;;) {  // the closure usually settles in a few rounds
    if (M == 0) break;
    P1_CUDA(cudaMemsetAsync(d, 0, sizeof(int), s));
    nvb_closure_kernel<<<agrid(ctx, M), AT, 0, s>>>(M, (const long long *)element2edges, edge_flag,
                                                    d);
    P1_CUDA(cudaMemcpyAsync(h, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    P1_CUDA(cudaStreamSynchronize(s));
    if (h[1]) {
      set_error("p1_nvb_plan: marked element index out of range [0, %lld)", (long long)M);
      return P1_EINDEX;
    }
    if (!h[0]) break;
  }
  int *counts = nullptr;  // children per element, then the boundary flags
  P1_TRY(pool_alloc(ctx, &counts, (size_t)(M + K) + 2, s));
  if (M + K > 0)
    nvb_count_kernel<<<agrid(ctx, max(M, K)), AT, 0, s>>>(
        M, (const long long *)element2edges, edge_flag, counts, K,
        (const long long *)boundary2edges, counts + M + 1);
  int rc = P1_OK;
  if ((rc = exclusive_scan_i32(ctx, edge_flag, edge_new, E, true, s)) ||
      (rc = exclusive_scan_i32(ctx, counts, child_first, M, true, s)) ||
      (K > 0 && (rc = exclusive_scan_i32(ctx, counts + M + 1, bnd_before, K, true, s)))) {
    pool_free(ctx, counts, s);
    return rc;
  }
  pool_free(ctx, counts, s);
  int tot[2] = {0, 0};
  P1_CUDA(cudaMemcpyAsync(&tot[0], edge_new + E, sizeof(int), cudaMemcpyDeviceToHost, s));
  P1_CUDA(cudaMemcpyAsync(&tot[1], child_first + M, sizeof(int), cudaMemcpyDeviceToHost, s));
  std::vector<int> before((size_t)n_boundaries + 1, 0);
  for (int j = 0; j <= n_boundaries && K > 0; ++j)
    P1_CUDA(cudaMemcpyAsync(&before[j], bnd_before + boundary_offsets_host[j], sizeof(int),
                            cudaMemcpyDeviceToHost, s));
  P1_CUDA(cudaStreamSynchronize(s));
  if (tot[0] < 0 || tot[1] < 0) {
    set_error("p1_nvb_plan: the refined mesh exceeds int32 indices");
    return P1_ERANGE;
  }
  counts_host[0] = tot[0];
  counts_host[1] = tot[1];
  for (int j = 0; j < n_boundaries; ++j) counts_host[2 + j] = K > 0 ? before[j + 1] - before[j] : 0;
  return P1_OK;
}

// Stage 2: the refined mesh.  new_coords: (N + new nodes) x 2 with the old coordinates
// already in place (likewise new_embed, or NULL); new_elements: counts[1] x 3;
// new_boundaries_host[j]: device pointer to (K_j + split_j) x 2 int32.
extern "C" int p1_nvb_apply(p1_ctx *ctx, const double *coords, int64_t n_nodes,
                            const int32_t *elements, int64_t n_elements, int64_t n_edges,
                            const int64_t *element2edges, const int64_t *edge2nodes,
                            const int32_t *boundary_nodes, const int64_t *boundary2edges,
                            const int64_t *boundary_offsets_host, int n_boundaries,
                            const int32_t *edge_flag, const int32_t *edge_new,
                            const int32_t *child_first, const int32_t *bnd_before,
                            const int64_t *counts_host, const double *embed, int sorted,
                            double *new_coords, double *new_embed, int32_t *new_elements,
                            int32_t *const *new_boundaries_host, void *stream) {
  P1_REQUIRE(ctx && counts_host, "null argument");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t M = n_elements, E = n_edges, N = n_nodes;
  if (E > 0 && counts_host[0] > 0)
    nvb_nodes_kernel<<<agrid(ctx, E), AT, 0, s>>>(E, edge_flag, edge_new,
                                                  (const long long *)edge2nodes,
                                                  (const double2 *)coords, embed, N,
                                                  (double2 *)new_coords, new_embed);
  if (M > 0)
    nvb_children_kernel<<<agrid(ctx, M), AT, 0, s>>>(M, elements,
                                                     (const long long *)element2edges, edge_flag,
                                                     edge_new, child_first, N, sorted,
                                                     new_elements);
  std::vector<int> before((size_t)n_boundaries + 1, 0);
  for (int j = 0, acc = 0; j < n_boundaries; ++j) {
    before[j] = acc;
    acc += (int)counts_host[2 + j];
  }
  for (int j = 0; j < n_boundaries; ++j) {
    const int64_t lo = boundary_offsets_host[j], hi = boundary_offsets_host[j + 1];
    if (hi > lo)
      nvb_boundary_kernel<<<grid_for(hi - lo, AT), AT, 0, s>>>(
          lo, hi, before[j], (int)counts_host[2 + j], boundary_nodes,
          (const long long *)boundary2edges, edge_flag, bnd_before, edge_new, N,
          new_boundaries_host[j]);
  }
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

// b + Neumann contributions (solvers.py:601-618); g_mid[j] = g(midpoint of edge j)
extern "C" int p1_apply_neumann(p1_ctx *ctx, const double *coords, int64_t n_nodes,
                                const int32_t *edges, int64_t n_edges, const double *g_mid,
                                const double *b, double *out, void *stream) {
  P1_REQUIRE(ctx && n_nodes >= 0 && n_edges >= 0, "bad argument");
  P1_REQUIRE(n_nodes == 0 || (b && out), "null b/out");
  P1_REQUIRE(n_edges == 0 || (coords && edges && g_mid), "null edge data");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  double *acc = nullptr;
  P1_TRY(pool_alloc(ctx, &acc, (size_t)n_nodes + 1, s));
  cudaMemsetAsync(acc, 0, ((size_t)n_nodes + 1) * sizeof(double), s);
  cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s);
  for (int phase = 0; phase < 2 && n_edges > 0; ++phase)
    neumann_kernel<<<grid_for(n_edges, AT), AT, 0, s>>>(phase, edges, n_edges,
                                                        (const double2 *)coords, g_mid,
                                                        (int)n_nodes, acc, ctx->d_flags);
  if (n_nodes > 0) add_kernel<<<agrid(ctx, n_nodes), AT, 0, s>>>(n_nodes, b, acc, out);
  cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  pool_free(ctx, acc, s);
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    set_error("p1_apply_neumann: %s", cudaGetErrorString(e));
    return P1_ECUDA;
  }
  if (ctx->h_flags[0]) {
    set_error("p1_apply_neumann: boundary node index out of range");
    return P1_EINDEX;
  }
  return P1_OK;
}

// y = y0 + alpha * A x on a device CSR (y0 may be NULL; y may alias y0): the Dirichlet
// lift b - A x_D (solvers.py:679-681) and the energy x.A x (solvers.py:695)
extern "C" int p1_csr_spmv(p1_ctx *ctx, int64_t n_rows, const int32_t *indptr,
                           const int32_t *indices, const double *data, const double *x,
                           double alpha, const double *y0, double *y, void *stream) {
  P1_REQUIRE(ctx && n_rows >= 0 && (n_rows == 0 || (indptr && x && y)), "bad argument");
  DeviceGuard guard(ctx);
  if (n_rows > 0)
    spmv_kernel<<<grid_for(n_rows * LANES, AT), AT, 0, (cudaStream_t)stream>>>(
        n_rows, indptr, indices, data, x, nullptr, alpha, y0, y);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

// a . b in a fixed summation tree; SYNCHRONISES (the result is a host scalar)
extern "C" int p1_dot(p1_ctx *ctx, int64_t n, const double *a, const double *b,
                      double *result_host, void *stream) {
  P1_REQUIRE(ctx && result_host && n >= 0 && (n == 0 || (a && b)), "bad argument");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  *result_host = 0.;
  if (n == 0) return P1_OK;
  double *partial = nullptr;
  P1_TRY(pool_alloc(ctx, &partial, (size_t)DOT_BLOCKS + 1, s));
  dot_kernel<<<DOT_BLOCKS, AT, 0, s>>>(n, a, b, partial);
  dot_final_kernel<<<1, AT, 0, s>>>(DOT_BLOCKS, partial, partial + DOT_BLOCKS);
  cudaMemcpyAsync(result_host, partial + DOT_BLOCKS, sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  pool_free(ctx, partial, s);
  P1_CUDA(e);
  return P1_OK;
}

// Jacobi-preconditioned conjugate gradient for A x = b on the FREE nodes (fixed[i] != 0:
// x[i] is a Dirichlet value and stays; its row and column are lifted to the right-hand
// side, solvers.py:676-693) -- what keeps the matrix on the GPU where the reference
// hands it to a direct solver.  x: initial guess at the free nodes.  Stops when
// |r| <= tol * |r0| (checked every 8 iterations) or after max_iter iterations.
extern "C" int p1_pcg(p1_ctx *ctx, int64_t n, const int32_t *indptr, const int32_t *indices,
                      const double *data, const double *b, const unsigned char *fixed, double *x,
                      double tol, int max_iter, int *iterations_host,
                      double *relative_residual_host, void *stream) {
  P1_REQUIRE(ctx && n >= 0 && max_iter >= 0 && tol >= 0., "bad argument");
  P1_REQUIRE(n == 0 || (indptr && b && fixed && x), "null argument");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  if (iterations_host) *iterations_host = 0;
  if (relative_residual_host) *relative_residual_host = 0.;
  if (n == 0) return P1_OK;
  double *work = nullptr;  // r, z, p, q, diag (n each), partial (2 * DOT_BLOCKS), scal (8)
  P1_TRY(pool_alloc(ctx, &work, (size_t)5 * n + 2 * DOT_BLOCKS + 8, s));
  double *r = work, *z = r + n, *p = z + n, *q = p + n, *diag = q + n;
  double *partial = diag + n, *scal = partial + 2 * DOT_BLOCKS;
  const unsigned g = agrid(ctx, n), gs = grid_for(n * LANES, AT);
  cudaMemsetAsync(scal, 0, 8 * sizeof(double), s);
  diag_kernel<<<grid_for(n, AT), AT, 0, s>>>(n, indptr, indices, data, diag);
  spmv_kernel<<<gs, AT, 0, s>>>(n, indptr, indices, data, x, nullptr, 1., nullptr, q);
  cg_init_kernel<<<g, AT, 0, s>>>(n, b, q, diag, fixed, r, z, p);
  dot_kernel<<<DOT_BLOCKS, AT, 0, s>>>(n, r, z, partial);
  dot_final_kernel<<<1, AT, 0, s>>>(DOT_BLOCKS, partial, scal);          // rz
  dot_kernel<<<DOT_BLOCKS, AT, 0, s>>>(n, r, r, partial);
  dot_final_kernel<<<1, AT, 0, s>>>(DOT_BLOCKS, partial, scal + 3);      // rr
  double host[8];
  cudaMemcpyAsync(host, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  const double rr0 = host[3];
  double rr = rr0;
  int it = 0;
  while (e == cudaSuccess && it < max_iter && rr > tol * tol * rr0 && rr0 > 0.) {
    for (int k = 0; k < 8 && it < max_iter; ++k, ++it) {
      spmv_kernel<<<gs, AT, 0, s>>>(n, indptr, indices, data, p, fixed, 1., nullptr, q);
      dot_kernel<<<DOT_BLOCKS, AT, 0, s>>>(n, p, q, partial);
      dot_final_kernel<<<1, AT, 0, s>>>(DOT_BLOCKS, partial, scal + 1);  // pq
      cg_update_kernel<<<DOT_BLOCKS, AT, 0, s>>>(n, scal, p, q, diag, fixed, x, r, z, partial);
      cg_scalars_kernel<<<1, AT, 0, s>>>(DOT_BLOCKS, partial, scal);
      cg_direction_kernel<<<g, AT, 0, s>>>(n, scal, z, fixed, p);
    }
    cudaMemcpyAsync(host, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, s);
    e = cudaStreamSynchronize(s);
    rr = host[3];
  }
  pool_free(ctx, work, s);
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    set_error("p1_pcg: %s", cudaGetErrorString(e));
    return P1_ECUDA;
  }
  if (iterations_host) *iterations_host = it;
  if (relative_residual_host) *relative_residual_host = rr0 > 0. ? sqrt(rr / rr0) : 0.;
  return P1_OK;
}

// ------------------------------------------------- element blocks (multi-GPU eta_R)
// The estimator of an element needs its three neighbours, hence the complete fans of
// its three nodes.  A rank that owns the element block [m0, m1) therefore works on the
// sub-mesh of all elements that touch a node of the block (the block and a ring of
// neighbours): the indicators of the block's elements computed there are the global
// ones bit for bit.
namespace p1 {
__global__ void __launch_bounds__(AT)
block_nodes_kernel(const int *__restrict__ elements, int64_t m0, int64_t m1, int N,
                   unsigned char *__restrict__ node_flag, int *__restrict__ bad) {
  for (int64_t i = 3 * m0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * m1;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int v = elements[i];
    if ((unsigned)v >= (unsigned)N) { *bad = 1; continue; }
    node_flag[v] = 1;
  }
}
__global__ void __launch_bounds__(AT)
block_keep_kernel(const int *__restrict__ elements, int64_t M, int N,
                  const unsigned char *__restrict__ node_flag, int *__restrict__ keep,
                  int *__restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M; t += stride) {
    const int a = __ldg(elements + 3 * t), b = __ldg(elements + 3 * t + 1),
              c = __ldg(elements + 3 * t + 2);
    if ((unsigned)a >= (unsigned)N || (unsigned)b >= (unsigned)N || (unsigned)c >= (unsigned)N) {
      *bad = 1;
      keep[t] = 0;
      continue;
    }
    keep[t] = (node_flag[a] | node_flag[b] | node_flag[c]) != 0;
  }
}
__global__ void __launch_bounds__(AT)
block_gather_kernel(const int *__restrict__ elements, int64_t M, const int *__restrict__ pos,
                    int *__restrict__ sub_elements, int *__restrict__ sub_ids) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M; t += stride) {
    const int p = pos[t];
    if (pos[t + 1] == p) continue;
    sub_elements[3 * (int64_t)p] = elements[3 * t];
    sub_elements[3 * (int64_t)p + 1] = elements[3 * t + 1];
    sub_elements[3 * (int64_t)p + 2] = elements[3 * t + 2];
    if (sub_ids) sub_ids[p] = (int)t;
  }
}
}  // namespace p1

// Stage 1: node_flag[v] = 1 for the nodes of the block (N bytes, caller's), pos (M + 1
// ints, caller's) = exclusive scan of "element touches a flagged node".
// counts_host: [0] elements of the sub-mesh, [1] position of element m0 in it (the
// block's elements follow one another there).  SYNCHRONISES.
extern "C" int p1_block_submesh(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                                int64_t n_nodes, int64_t m0, int64_t m1,
                                unsigned char *node_flag, int32_t *pos, int64_t *counts_host,
                                void *stream) {
  P1_REQUIRE(ctx && counts_host && n_elements >= 0 && n_nodes >= 0 && m0 >= 0 && m0 <= m1 &&
                 m1 <= n_elements, "bad argument");
  P1_REQUIRE(n_elements == 0 || (elements && node_flag && pos), "null array");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t M = n_elements;
  counts_host[0] = counts_host[1] = 0;
  if (M == 0) return P1_OK;
  cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s);
  cudaMemsetAsync(node_flag, 0, (size_t)n_nodes, s);
  if (m1 > m0)
    block_nodes_kernel<<<agrid(ctx, 3 * (m1 - m0)), AT, 0, s>>>(elements, m0, m1, (int)n_nodes,
                                                                node_flag, ctx->d_flags);
  int *keep = nullptr;
  P1_TRY(pool_alloc(ctx, &keep, (size_t)M + 1, s));
  block_keep_kernel<<<agrid(ctx, M), AT, 0, s>>>(elements, M, (int)n_nodes, node_flag, keep,
                                                 ctx->d_flags);
  int rc = exclusive_scan_i32(ctx, keep, pos, M, true, s);
  pool_free(ctx, keep, s);
  if (rc) return rc;
  int h[2] = {0, 0};
  cudaMemcpyAsync(&h[0], pos + M, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(&h[1], pos + m0, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s);
  P1_CUDA(cudaStreamSynchronize(s));
  if (ctx->h_flags[0]) {
    set_error("p1_block_submesh: element index out of range [0, %lld)", (long long)n_nodes);
    return P1_EINDEX;
  }
  counts_host[0] = h[0];
  counts_host[1] = h[1];
  return P1_OK;
}

// Stage 2: the sub-mesh's elements (counts[0] x 3) and, optionally, their global indices.
extern "C" int p1_block_submesh_fill(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                                     const int32_t *pos, int32_t *sub_elements,
                                     int32_t *sub_ids, void *stream) {
  P1_REQUIRE(ctx && n_elements >= 0 && (n_elements == 0 || (elements && pos && sub_elements)),
             "bad argument");
  DeviceGuard guard(ctx);
  if (n_elements > 0)
    block_gather_kernel<<<agrid(ctx, n_elements), AT, 0, (cudaStream_t)stream>>>(
        elements, n_elements, pos, sub_elements, sub_ids);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

// ------------------------------------------------ multi-GPU: global row pointer
// Interleaved ownership (chunks of 2^c rows dealt round-robin): a rank's piece is
// placed in the global CSR by the offset of each of its chunks.  chunk_nnz: the stored
// entries of each of my chunks (one launch); after the all-gather of all ranks' values
// chunk_offsets scans them in global chunk order (chunk q * world + r sits at
// every[r * per_rank + q]) and keeps mine (one block).
namespace p1 {
__global__ void __launch_bounds__(AT)
chunk_nnz_kernel(const int *__restrict__ indptr, int n_own, int cshift, int per_rank,
                 long long *__restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= per_rank) return;
  const long long lo = min((long long)q << cshift, (long long)n_own);
  const long long hi = min(((long long)q + 1) << cshift, (long long)n_own);
  out[q] = (long long)indptr[hi] - indptr[lo];
}
__global__ void __launch_bounds__(1024)
chunk_offsets_kernel(const long long *__restrict__ every, int world, int per_rank, int rank,
                     long long *__restrict__ mine, long long *__restrict__ total) {
  __shared__ long long warp_sums[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int n = world * per_rank;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int g = base + threadIdx.x;           // global chunk index
    const int r = g % world, q = g / world;
    const long long v = g < n ? every[(long long)r * per_rank + q] : 0;
    long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    long long before = 0;
    for (int w = 0; w < warp; ++w) before += warp_sums[w];
    const long long carry = carry_s;
    if (g < n && r == rank) mine[q] = carry + before + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + before + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}
}  // namespace p1

extern "C" int p1_chunk_nnz(p1_ctx *ctx, const int32_t *indptr, int64_t n_own, int chunk_log2,
                            int64_t per_rank, int64_t *out, void *stream) {
  P1_REQUIRE(ctx && indptr && out && n_own >= 0 && per_rank >= 0 && chunk_log2 >= 0 &&
                 chunk_log2 < 31, "bad argument");
  DeviceGuard guard(ctx);
  if (per_rank > 0)
    chunk_nnz_kernel<<<grid_for(per_rank, AT), AT, 0, (cudaStream_t)stream>>>(
        indptr, (int)n_own, chunk_log2, (int)per_rank, (long long *)out);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_chunk_offsets(p1_ctx *ctx, const int64_t *every, int world, int64_t per_rank,
                                int rank, int64_t *mine, int64_t *total, void *stream) {
  P1_REQUIRE(ctx && every && mine && total && world >= 1 && rank >= 0 && rank < world &&
                 per_rank >= 0, "bad argument");
  DeviceGuard guard(ctx);
  chunk_offsets_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
      (const long long *)every, world, (int)per_rank, rank, (long long *)mine,
      (long long *)total);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}
