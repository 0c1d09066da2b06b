// local_ops.cu -- per-triangle kernels: geometry, triplet streams, points for
// user callbacks, local matrices of the generalised stiffness / weighted mass,
// local load vectors, and the per-node gathers that turn them into the
// reference's bincount results.
//
// One thread per triangle; element indices are read coalesced, coordinates are
// gathered (element order of NVB meshes is spatially local, so the gathers hit
// L1/L2).  Arithmetic mirrors the reference expression by expression
// (-fmad=false): /root/reference/p1afempy/mesh.py:10-51, solvers.py:50-142,
// 156-182, 278-377, 414-426, 540-598.
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 256;
// quadrature points fetched together by the streaming kernels: a thread keeps QB
// loads of the [n_qp x M] arrays in flight (local loads at 64M: 2.70 ms with 1, 2.37 with 4,
// 3.14 with 8)
#ifndef P1_QB
#define P1_QB 4
#endif
constexpr int QB = P1_QB;
#ifndef P1_INTEGRATE_BLOCKS
#define P1_INTEGRATE_BLOCKS 8192
#endif

struct Verts {
  int e0, e1, e2;
  double2 z0, z1, z2;
};

__device__ __forceinline__ void load_indices(const int *__restrict__ elements,
                                             int64_t t, int &e0, int &e1, int &e2) {
  e0 = __ldg(elements + 3 * t);
  e1 = __ldg(elements + 3 * t + 1);
  e2 = __ldg(elements + 3 * t + 2);
}

__device__ __forceinline__ Verts load_verts(const int *__restrict__ elements,
                                            const double2 *__restrict__ coords,
                                            int64_t t) {
  Verts v;
  load_indices(elements, t, v.e0, v.e1, v.e2);
  v.z0 = __ldg(coords + v.e0);
  v.z1 = __ldg(coords + v.e1);
  v.z2 = __ldg(coords + v.e2);
  return v;
}

#define P1_FOR_ELEMENTS(t, M)                                                 \
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (M);  \
       t += (int64_t)gridDim.x * blockDim.x)

// ------------------------------------------------------------- geometry
__global__ void __launch_bounds__(THREADS)
geometry_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                int64_t M, double *__restrict__ area, double2 *__restrict__ d21,
                double2 *__restrict__ d31) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double px = v.z1.x - v.z0.x, py = v.z1.y - v.z0.y;
    const double qx = v.z2.x - v.z0.x, qy = v.z2.y - v.z0.y;
    if (area) area[t] = 0.5 * (px * qy - py * qx);
    if (d21) d21[t] = make_double2(px, py);
    if (d31) d31[t] = make_double2(qx, qy);
  }
}

// ------------------------------------------------------ triplet streams
// entry k of element t: row = vertex k%3, col = vertex k/3 (solvers.py:34-37)
template <bool STIFFNESS>
__global__ void __launch_bounds__(THREADS)
triplets_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                int64_t M, int64_t *__restrict__ rows, int64_t *__restrict__ cols,
                double *__restrict__ vals) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double px = v.z1.x - v.z0.x, py = v.z1.y - v.z0.y;
    const double qx = v.z2.x - v.z0.x, qy = v.z2.y - v.z0.y;
    const double area = 0.5 * (px * qy - py * qx);
    double m[9];
    if (STIFFNESS) {
      const double area4 = 4. * area;
      const double a = (px * qx + py * qy) / area4;
      const double b = (qx * qx + qy * qy) / area4;
      const double c = (px * px + py * py) / area4;
      m[0] = -2. * a + b + c; m[1] = a - b; m[2] = a - c;
      m[3] = a - b; m[4] = b; m[5] = -a;
      m[6] = a - c; m[7] = -a; m[8] = c;
    } else {
      const double on = area / 6., off = area / 12.;
      m[0] = on; m[1] = off; m[2] = off; m[3] = off; m[4] = on; m[5] = off;
      m[6] = off; m[7] = off; m[8] = on;
    }
    const int e[3] = {v.e0, v.e1, v.e2};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      rows[9 * t + k] = e[k % 3];
      cols[9 * t + k] = e[k / 3];
      vals[9 * t + k] = m[k];
    }
  }
}

// ------------------------------------------- points for user callbacks
__global__ void __launch_bounds__(THREADS)
quad_points_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                   int64_t M, const double2 *__restrict__ points, int n_qp,
                   double2 *__restrict__ xq) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double ax = v.z1.x - v.z0.x, ay = v.z1.y - v.z0.y;
    const double bx = v.z2.x - v.z0.x, by = v.z2.y - v.z0.y;
    for (int q = 0; q < n_qp; ++q) {
      const double eta = points[q].x, xi = points[q].y;
      xq[(int64_t)q * M + t] = make_double2(v.z0.x + eta * ax + xi * bx,
                                            v.z0.y + eta * ay + xi * by);
    }
  }
}

__global__ void __launch_bounds__(THREADS)
nodal_at_quad_kernel(const double *__restrict__ u, const int *__restrict__ elements,
                     int64_t M, const double2 *__restrict__ points, int n_qp,
                     double *__restrict__ uq) {
  P1_FOR_ELEMENTS(t, M) {
    int e0, e1, e2;
    load_indices(elements, t, e0, e1, e2);
    const double u1 = __ldg(u + e0), u2 = __ldg(u + e1), u3 = __ldg(u + e2);
    for (int q = 0; q < n_qp; ++q) {
      const double eta = points[q].x, xi = points[q].y;
      const double h1 = 1 - eta - xi;
      uq[(int64_t)q * M + t] = u1 * h1 + u2 * eta + u3 * xi;
    }
  }
}

__global__ void __launch_bounds__(THREADS)
centroids_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                 int64_t M, double2 *__restrict__ out) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double px = v.z1.x - v.z0.x, py = v.z1.y - v.z0.y;
    const double qx = v.z2.x - v.z0.x, qy = v.z2.y - v.z0.y;
    out[t] = make_double2(v.z0.x + (px + qx) / 3., v.z0.y + (py + qy) / 3.);
  }
}

__global__ void __launch_bounds__(THREADS)
boundary_edges_kernel(const double2 *__restrict__ coords, const int *__restrict__ edges,
                      int64_t K, double *__restrict__ length,
                      double2 *__restrict__ midpoint) {
  P1_FOR_ELEMENTS(j, K) {
    const double2 a = coords[edges[2 * j]], b = coords[edges[2 * j + 1]];
    if (midpoint) midpoint[j] = make_double2((a.x + b.x) / 2., (a.y + b.y) / 2.);
    if (length) {
      const double dx = b.x - a.x, dy = b.y - a.y;
      length[j] = sqrt(dx * dx + dy * dy);
    }
  }
}

// ------------------------------------ local matrices, general stiffness
// solvers.py:306-366
__global__ void __launch_bounds__(THREADS)
local_general_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                     int64_t M, const double *__restrict__ a11q,
                     const double *__restrict__ a12q, const double *__restrict__ a21q,
                     const double *__restrict__ a22q, const double *__restrict__ weights,
                     int n_qp, double *__restrict__ local) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double dz1x = v.z1.x - v.z0.x, dz1y = v.z1.y - v.z0.y;
    const double dz2x = v.z2.x - v.z0.x, dz2y = v.z2.y - v.z0.y;
    const double alpha = dz2y, beta = -dz2x, gamma = -dz1y, delta = dz1x;
    const double areas = 0.5 * (dz1x * dz2y - dz1y * dz2x);
    double a = 0., b = 0., c = 0., d = 0.;
    for (int q0 = 0; q0 < n_qp; q0 += 4) {  // 16 loads in flight
      double v11[4], v12[4], v21[4], v22[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = (int64_t)(q0 + u) * M + t;
        const bool in = q0 + u < n_qp;
        v11[u] = in ? __ldg(a11q + i) : 0.;
        v12[u] = in ? __ldg(a12q + i) : 0.;
        v21[u] = in ? __ldg(a21q + i) : 0.;
        v22[u] = in ? __ldg(a22q + i) : 0.;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q0 + u < n_qp) {
          const double w = weights[q0 + u];
          a += w * v11[u];
          b += w * v12[u];
          c += w * v21[u];
          d += w * v22[u];
        }
    }
    const double pre = -1. / (2. * areas);
    const double b11 = pre * (a * alpha * alpha + b * beta * alpha +
                              c * alpha * beta + d * beta * beta);
    const double b12 = pre * (a * alpha * gamma + b * delta * alpha +
                              c * gamma * beta + d * beta * delta);
    const double b21 = pre * (a * alpha * gamma + b * beta * gamma +
                              c * alpha * delta + d * beta * delta);
    const double b22 = pre * (a * gamma * gamma + b * delta * gamma +
                              c * gamma * delta + d * delta * delta);
    double *m = local + 9 * t;
    m[0] = b11 + b12 + b21 + b22;
    m[1] = -(b11 + b21);
    m[2] = -(b12 + b22);
    m[3] = -(b11 + b12);
    m[4] = b11;
    m[5] = b12;
    m[6] = -(b21 + b22);
    m[7] = b21;
    m[8] = b22;
  }
}

// ---------------------------------------- local matrices, weighted mass
// solvers.py:86-129
__global__ void __launch_bounds__(THREADS)
local_wmass_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                   int64_t M, const double *__restrict__ phiq,
                   const double *__restrict__ weights, const double2 *__restrict__ points,
                   int n_qp, double *__restrict__ local) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double dz1x = v.z1.x - v.z0.x, dz1y = v.z1.y - v.z0.y;
    const double dz2x = v.z2.x - v.z0.x, dz2y = v.z2.y - v.z0.y;
    const double areas = 0.5 * (dz1x * dz2y - dz1y * dz2x);
    double m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = 0.;
    for (int q0 = 0; q0 < n_qp; q0 += QB) {
      double fv[QB];
#pragma unroll
      for (int u = 0; u < QB; ++u)
        fv[u] = q0 + u < n_qp ? __ldg(phiq + (int64_t)(q0 + u) * M + t) : 0.;
#pragma unroll
      for (int u = 0; u < QB; ++u)
        if (q0 + u < n_qp) {
          const double eta = points[q0 + u].x, xi = points[q0 + u].y;
          const double h[3] = {1 - eta - xi, eta, xi};
          const double base = 2 * areas * weights[q0 + u] * fv[u];
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) m[3 * k + l] += base * h[k] * h[l];
        }
    }
    double *out = local + 9 * t;
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = m[k];
  }
}

// --------------------------------------------------- local load vectors
// solvers.py:564-591: L[t][j] += 2.*w*areas*hat_j*f(x_q)
__global__ void __launch_bounds__(THREADS)
local_load_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                  int64_t M, const double *__restrict__ fq,
                  const double *__restrict__ weights, const double2 *__restrict__ points,
                  int n_qp, double *__restrict__ L) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double px = v.z1.x - v.z0.x, py = v.z1.y - v.z0.y;
    const double qx = v.z2.x - v.z0.x, qy = v.z2.y - v.z0.y;
    const double area = 0.5 * (px * qy - py * qx);
    double l0 = 0., l1 = 0., l2 = 0.;
    // QB points at a time in flight; the sums stay in the reference's order
    for (int q0 = 0; q0 < n_qp; q0 += QB) {
      double fv[QB];
#pragma unroll
      for (int u = 0; u < QB; ++u)
        fv[u] = q0 + u < n_qp ? __ldg(fq + (int64_t)(q0 + u) * M + t) : 0.;
#pragma unroll
      for (int u = 0; u < QB; ++u)
        if (q0 + u < n_qp) {
          const double eta = points[q0 + u].x, xi = points[q0 + u].y;
          const double s = 2. * weights[q0 + u] * area;
          l0 += s * (1. - eta - xi) * fv[u];
          l1 += s * eta * fv[u];
          l2 += s * xi * fv[u];
        }
    }
    L[3 * t] = l0; L[3 * t + 1] = l1; L[3 * t + 2] = l2;
  }
}

// solvers.py:415-424: share[t] = area4*fsT/12.
__global__ void __launch_bounds__(THREADS)
local_load_midpoint_kernel(const double2 *__restrict__ coords,
                           const int *__restrict__ elements, int64_t M,
                           const double *__restrict__ f_centroid,
                           double *__restrict__ share) {
  P1_FOR_ELEMENTS(t, M) {
    const Verts v = load_verts(elements, coords, t);
    const double px = v.z1.x - v.z0.x, py = v.z1.y - v.z0.y;
    const double qx = v.z2.x - v.z0.x, qy = v.z2.y - v.z0.y;
    const double area4 = 4. * (0.5 * (px * qy - py * qx));
    share[t] = area4 * f_centroid[t] / 12.;
  }
}

// ------------------------------------------------ integral of f(u_h)
// solvers.py:457-472: L[t] += w * f(u_q) * 2. * areas (left to right); the sum
// over the elements is a fixed-order tree (numpy's pairwise sum differs in the
// last bits; the integral is a scalar with tolerance 1e-12).
__global__ void __launch_bounds__(THREADS)
integrate_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                 int64_t M, const double *__restrict__ fq,
                 const double *__restrict__ weights, int n_qp,
                 double *__restrict__ partial) {
  __shared__ double red[THREADS];
  double acc = 0.;
  // every block owns a fixed contiguous slice, every thread a fixed stride in it:
  // the summation order depends on (M, grid) only
  const int64_t per_block = (M + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per_block;
  const int64_t hi = lo + per_block < M ? lo + per_block : M;
  for (int64_t t = lo + threadIdx.x; t < hi; t += THREADS) {
    const Verts v = load_verts(elements, coords, t);
    const double px = v.z1.x - v.z0.x, py = v.z1.y - v.z0.y;
    const double qx = v.z2.x - v.z0.x, qy = v.z2.y - v.z0.y;
    const double area = 0.5 * (px * qy - py * qx);
    double l = 0.;
    for (int q0 = 0; q0 < n_qp; q0 += QB) {
      double fv[QB];
#pragma unroll
      for (int u = 0; u < QB; ++u)
        fv[u] = q0 + u < n_qp ? __ldg(fq + (int64_t)(q0 + u) * M + t) : 0.;
#pragma unroll
      for (int u = 0; u < QB; ++u)
        if (q0 + u < n_qp) l += weights[q0 + u] * fv[u] * 2. * area;
    }
    acc += l;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(THREADS)
sum_partials_kernel(const double *__restrict__ partial, int n, double *__restrict__ out) {
  __shared__ double red[THREADS];
  double acc = 0.;
  for (int i = threadIdx.x; i < n; i += THREADS) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// ------------------------------------------------------ per-node gathers
// b[a] = sum of the node's local contributions in the ORDER of the reference's
// sequential np.bincount: ascending (element, vertex) for the cubature rule
// (bincount over elements.flatten(), solvers.py:593-596), ascending (vertex,
// element) for the legacy rule (elements.flatten('F'), solvers.py:422-425).
// The row's <= 8 element codes (T << 2 | k) are one 32-byte sector: they are
// sorted in registers, nothing is written back.
__device__ __forceinline__ unsigned gather_key(int code, bool vertex_major) {
  const unsigned c = (unsigned)code;
  return vertex_major ? ((c & 3u) << 29) | (c >> 2) : c;
}

template <bool VERTEX_MAJOR, int STRIDE>
__global__ void __launch_bounds__(THREADS)
gather_nodes_kernel(int n_own, Rows rows, const double *__restrict__ L,
                    double *__restrict__ b) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_own) return;
  const int deg = rows.deg(a);
  double sum = 0.;
  if (deg == 0) { b[a] = sum; return; }
  const int *code = rows.topo ? nullptr : rows.elem(a, deg);
  if (deg <= GROUP) {
    int cc[GROUP];
    if (rows.topo) {
      const int4 *slot = rows.topo + (int64_t)a * GROUP;
#pragma unroll
      for (int i = 0; i < GROUP; ++i) cc[i] = i < deg ? __ldg(slot + i).z : 0;
    } else {
      const int4 c0 = __ldg(reinterpret_cast<const int4 *>(code));
      const int4 c1 = __ldg(reinterpret_cast<const int4 *>(code) + 1);
      cc[0] = c0.x; cc[1] = c0.y; cc[2] = c0.z; cc[3] = c0.w;
      cc[4] = c1.x; cc[5] = c1.y; cc[6] = c1.z; cc[7] = c1.w;
    }
    unsigned key[GROUP];
#pragma unroll
    for (int i = 0; i < GROUP; ++i)
      key[i] = i < deg ? gather_key(cc[i], VERTEX_MAJOR) : 0xffffffffu;
    // rank by counting (keys of one row are distinct), then add in rank order
    int rank[GROUP];
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      rank[i] = 0;
#pragma unroll
      for (int j = 0; j < GROUP; ++j) rank[i] += key[j] < key[i];
    }
    double val[GROUP];
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      val[i] = 0.;
      if (i < deg) {
        const int t = cc[i] >> 2, k = cc[i] & 3;
        val[i] = __ldg(L + (STRIDE == 3 ? 3 * (int64_t)t + k : (int64_t)t));
      }
    }
    for (int q = 0; q < deg; ++q) {
#pragma unroll
      for (int i = 0; i < GROUP; ++i)
        if (rank[i] == q) sum += val[i];
    }
  } else {
    // more than 8 incidences: selection by key, no scratch
    unsigned last = 0;
    bool first = true;
    for (int q = 0; q < deg; ++q) {
      unsigned best = 0xffffffffu;
      int pick = 0;
      if (rows.topo) {
        topo_row_records(rows, a, n_own, [&](int, int4 r) {
          const unsigned key = gather_key(r.z, VERTEX_MAJOR);
          if ((first || key > last) && key < best) { best = key; pick = r.z; }
        });
      } else {
        for (int i = 0; i < deg; ++i) {
          const unsigned key = gather_key(code[i], VERTEX_MAJOR);
          if ((first || key > last) && key < best) { best = key; pick = code[i]; }
        }
      }
      const int t = pick >> 2, k = pick & 3;
      sum += __ldg(L + (STRIDE == 3 ? 3 * (int64_t)t + k : (int64_t)t));
      last = best;
      first = false;
    }
  }
  b[a] = sum;
}

// -------------------------------------------------------- index helpers
__global__ void __launch_bounds__(THREADS)
narrow_i64_kernel(const long long *__restrict__ src, int64_t n, int *__restrict__ dst) {
  P1_FOR_ELEMENTS(i, n) {
    const long long v = src[i];
    dst[i] = (v < 0 || v > 0x7fffffffLL) ? -1 : (int)v;
  }
}
__global__ void __launch_bounds__(THREADS)
narrow_u32_kernel(const unsigned *__restrict__ src, int64_t n, int *__restrict__ dst) {
  P1_FOR_ELEMENTS(i, n) {
    const unsigned v = src[i];
    dst[i] = v > 0x7fffffffu ? -1 : (int)v;
  }
}
__global__ void __launch_bounds__(THREADS)
widen_kernel(const int *__restrict__ src, int64_t n, long long *__restrict__ dst) {
  P1_FOR_ELEMENTS(i, n) dst[i] = src[i];
}

__global__ void __launch_bounds__(THREADS)
check_indices_kernel(const int *__restrict__ idx, int64_t n, int limit,
                     int *__restrict__ flags) {
  bool bad = false;
  P1_FOR_ELEMENTS(i, n) bad = bad || (unsigned)__ldg(idx + i) >= (unsigned)limit;
  if (bad) atomicOr(flags, FLAG_INDEX_RANGE);
}

static inline unsigned egrid(const p1_ctx *ctx, int64_t n) {
  unsigned g = grid_for(n, THREADS);
  unsigned cap = (unsigned)ctx->sm_count * 32u;
  return g < cap ? g : cap;
}

// copies a small host array to pool memory (valid in stream order)
template <typename T>
static int upload_small(p1_ctx *ctx, T **dst, const T *src_host, size_t count,
                        cudaStream_t s) {
  P1_TRY(pool_alloc(ctx, dst, count, s));
  P1_CUDA(cudaMemcpyAsync(*dst, src_host, count * sizeof(T), cudaMemcpyHostToDevice, s));
  return P1_OK;
}

}  // namespace p1

using namespace p1;

extern "C" int p1_narrow_indices(p1_ctx *ctx, const void *src, int src_kind,
                                 int64_t count, int32_t *dst, void *stream) {
  P1_REQUIRE(ctx && (count == 0 || (src && dst)), "null argument");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  if (count == 0) return P1_OK;
  if (src_kind == 0)
    narrow_i64_kernel<<<egrid(ctx, count), THREADS, 0, s>>>((const long long *)src, count, dst);
  else if (src_kind == 1)
    narrow_u32_kernel<<<egrid(ctx, count), THREADS, 0, s>>>((const unsigned *)src, count, dst);
  else if (src_kind == 2)
    P1_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(int), cudaMemcpyDeviceToDevice, s));
  else
    P1_REQUIRE(false, "src_kind must be 0 (int64), 1 (uint32) or 2 (int32)");
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_check_indices(p1_ctx *ctx, const int32_t *indices, int64_t count,
                                int64_t n_nodes, void *stream) {
  P1_REQUIRE(ctx && (count == 0 || indices), "null argument");
  P1_REQUIRE(n_nodes >= 0 && n_nodes <= 0x7fffffffLL, "bad n_nodes");
  DeviceGuard guard(ctx);
  if (count == 0) return P1_OK;
  cudaStream_t s = (cudaStream_t)stream;
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), s));
  check_indices_kernel<<<egrid(ctx, count), THREADS, 0, s>>>(indices, count, (int)n_nodes,
                                                            ctx->d_flags);
  P1_CUDA(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s));
  P1_CUDA(cudaStreamSynchronize(s));
  if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
    set_error("index out of range [0, %lld)", (long long)n_nodes);
    return P1_EINDEX;
  }
  return P1_OK;
}

extern "C" int p1_widen_indices(p1_ctx *ctx, const int32_t *src, int64_t count,
                                int64_t *dst, void *stream) {
  P1_REQUIRE(ctx && (count == 0 || (src && dst)), "null argument");
  DeviceGuard guard(ctx);
  if (count == 0) return P1_OK;
  widen_kernel<<<egrid(ctx, count), THREADS, 0, (cudaStream_t)stream>>>(
      src, count, (long long *)dst);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_geometry(p1_ctx *ctx, const double *coords, const int32_t *elements,
                           int64_t M, double *area, double *d21, double *d31,
                           void *stream) {
  P1_REQUIRE(ctx && (M == 0 || (coords && elements)), "null argument");
  DeviceGuard guard(ctx);
  if (M == 0) return P1_OK;
  geometry_kernel<<<egrid(ctx, M), THREADS, 0, (cudaStream_t)stream>>>(
      (const double2 *)coords, elements, M, area, (double2 *)d21, (double2 *)d31);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

static int triplets(p1_ctx *ctx, bool stiffness, const double *coords,
                    const int32_t *elements, int64_t M, int64_t *rows, int64_t *cols,
                    double *vals, void *stream) {
  P1_REQUIRE(ctx && (M == 0 || (coords && elements && rows && cols && vals)),
             "null argument");
  DeviceGuard guard(ctx);
  if (M == 0) return P1_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (stiffness)
    triplets_kernel<true><<<egrid(ctx, M), THREADS, 0, s>>>(
        (const double2 *)coords, elements, M, rows, cols, vals);
  else
    triplets_kernel<false><<<egrid(ctx, M), THREADS, 0, s>>>(
        (const double2 *)coords, elements, M, rows, cols, vals);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_mass_triplets(p1_ctx *ctx, const double *coords,
                                const int32_t *elements, int64_t M, int64_t *rows,
                                int64_t *cols, double *vals, void *stream) {
  return triplets(ctx, false, coords, elements, M, rows, cols, vals, stream);
}
extern "C" int p1_stiffness_triplets(p1_ctx *ctx, const double *coords,
                                     const int32_t *elements, int64_t M, int64_t *rows,
                                     int64_t *cols, double *vals, void *stream) {
  return triplets(ctx, true, coords, elements, M, rows, cols, vals, stream);
}

extern "C" int p1_quadrature_points(p1_ctx *ctx, const double *coords,
                                    const int32_t *elements, int64_t M,
                                    const double *points_host, int n_qp, double *xq,
                                    void *stream) {
  P1_REQUIRE(ctx && points_host && n_qp > 0, "bad rule");
  P1_REQUIRE(M == 0 || (coords && elements && xq), "null argument");
  DeviceGuard guard(ctx);
  if (M == 0) return P1_OK;
  cudaStream_t s = (cudaStream_t)stream;
  double *pts = nullptr;
  P1_TRY(upload_small(ctx, &pts, points_host, (size_t)2 * n_qp, s));
  quad_points_kernel<<<egrid(ctx, M), THREADS, 0, s>>>(
      (const double2 *)coords, elements, M, (const double2 *)pts, n_qp, (double2 *)xq);
  pool_free(ctx, pts, s);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_nodal_at_quadrature(p1_ctx *ctx, const double *u,
                                      const int32_t *elements, int64_t M,
                                      const double *points_host, int n_qp, double *uq,
                                      void *stream) {
  P1_REQUIRE(ctx && points_host && n_qp > 0, "bad rule");
  P1_REQUIRE(M == 0 || (u && elements && uq), "null argument");
  DeviceGuard guard(ctx);
  if (M == 0) return P1_OK;
  cudaStream_t s = (cudaStream_t)stream;
  double *pts = nullptr;
  P1_TRY(upload_small(ctx, &pts, points_host, (size_t)2 * n_qp, s));
  nodal_at_quad_kernel<<<egrid(ctx, M), THREADS, 0, s>>>(
      u, elements, M, (const double2 *)pts, n_qp, uq);
  pool_free(ctx, pts, s);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_centroids(p1_ctx *ctx, const double *coords, const int32_t *elements,
                            int64_t M, double *centroids, void *stream) {
  P1_REQUIRE(ctx && (M == 0 || (coords && elements && centroids)), "null argument");
  DeviceGuard guard(ctx);
  if (M == 0) return P1_OK;
  centroids_kernel<<<egrid(ctx, M), THREADS, 0, (cudaStream_t)stream>>>(
      (const double2 *)coords, elements, M, (double2 *)centroids);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_boundary_edges(p1_ctx *ctx, const double *coords, const int32_t *edges,
                                 int64_t K, double *length, double *midpoint,
                                 void *stream) {
  P1_REQUIRE(ctx && (K == 0 || (coords && edges)), "null argument");
  DeviceGuard guard(ctx);
  if (K == 0) return P1_OK;
  boundary_edges_kernel<<<egrid(ctx, K), THREADS, 0, (cudaStream_t)stream>>>(
      (const double2 *)coords, edges, K, length, (double2 *)midpoint);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_local_general_stiffness(p1_ctx *ctx, const double *coords,
                                          const int32_t *elements, int64_t M,
                                          const double *a11q, const double *a12q,
                                          const double *a21q, const double *a22q,
                                          const double *weights_host, int n_qp,
                                          double *local, void *stream) {
  P1_REQUIRE(ctx && weights_host && n_qp > 0, "bad rule");
  P1_REQUIRE(M == 0 || (coords && elements && a11q && a12q && a21q && a22q && local),
             "null argument");
  DeviceGuard guard(ctx);
  if (M == 0) return P1_OK;
  cudaStream_t s = (cudaStream_t)stream;
  double *w = nullptr;
  P1_TRY(upload_small(ctx, &w, weights_host, (size_t)n_qp, s));
  P1_TIMED(ctx, "local_general", s,
           (local_general_kernel<<<egrid(ctx, M), THREADS, 0, s>>>(
               (const double2 *)coords, elements, M, a11q, a12q, a21q, a22q, w, n_qp, local)));
  pool_free(ctx, w, s);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_local_weighted_mass(p1_ctx *ctx, const double *coords,
                                      const int32_t *elements, int64_t M,
                                      const double *phiq, const double *weights_host,
                                      const double *points_host, int n_qp,
                                      double *local, void *stream) {
  P1_REQUIRE(ctx && weights_host && points_host && n_qp > 0, "bad rule");
  P1_REQUIRE(M == 0 || (coords && elements && phiq && local), "null argument");
  DeviceGuard guard(ctx);
  if (M == 0) return P1_OK;
  cudaStream_t s = (cudaStream_t)stream;
  double *w = nullptr, *pts = nullptr;
  P1_TRY(upload_small(ctx, &w, weights_host, (size_t)n_qp, s));
  P1_TRY(upload_small(ctx, &pts, points_host, (size_t)2 * n_qp, s));
  P1_TIMED(ctx, "local_wmass", s,
           (local_wmass_kernel<<<egrid(ctx, M), THREADS, 0, s>>>(
               (const double2 *)coords, elements, M, phiq, w, (const double2 *)pts, n_qp, local)));
  pool_free(ctx, w, s);
  pool_free(ctx, pts, s);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_integrate_quadrature(p1_ctx *ctx, const double *coords,
                                       const int32_t *elements, int64_t M,
                                       const double *fq, const double *weights_host,
                                       int n_qp, double *result_host, void *stream) {
  P1_REQUIRE(ctx && weights_host && n_qp > 0 && result_host, "bad argument");
  P1_REQUIRE(M == 0 || (coords && elements && fq), "null argument");
  DeviceGuard guard(ctx);
  *result_host = 0.;
  if (M == 0) return P1_OK;
  cudaStream_t s = (cudaStream_t)stream;
  // fixed: the summation tree must not depend on the device (8192 slices: the last wave
  // of blocks is short on any SM count; 1024 left a third of the B200 idle for half the time)
  const int blocks = P1_INTEGRATE_BLOCKS;
  double *w = nullptr, *partial = nullptr;
  P1_TRY(upload_small(ctx, &w, weights_host, (size_t)n_qp, s));
  P1_TRY(pool_alloc(ctx, &partial, (size_t)blocks + 1, s));
  P1_TIMED(ctx, "integrate", s,
           (integrate_kernel<<<blocks, THREADS, 0, s>>>((const double2 *)coords, elements, M, fq,
                                                       w, n_qp, partial)));
  sum_partials_kernel<<<1, THREADS, 0, s>>>(partial, blocks, partial + blocks);
  P1_CUDA(cudaMemcpyAsync(result_host, partial + blocks, sizeof(double),
                          cudaMemcpyDeviceToHost, s));
  P1_CUDA(cudaStreamSynchronize(s));
  pool_free(ctx, w, s);
  pool_free(ctx, partial, s);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_load_vector(const p1_plan *plan, const double *coords,
                              const double *fq, const double *weights_host,
                              const double *points_host, int n_qp, double *b,
                              void *stream) {
  P1_REQUIRE(plan && weights_host && points_host && n_qp > 0, "bad plan/rule");
  const int64_t M = plan->M;
  const int n_own = (int)plan->n_own;
  P1_REQUIRE(n_own == 0 || b, "null b");
  P1_REQUIRE(M == 0 || (coords && fq), "null argument");
  p1_ctx *ctx = plan->ctx;
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  P1_TRY(require_elements(plan, "p1_load_vector"));
  double *w = nullptr, *pts = nullptr, *L = nullptr;
  P1_TRY(upload_small(ctx, &w, weights_host, (size_t)n_qp, s));
  P1_TRY(upload_small(ctx, &pts, points_host, (size_t)2 * n_qp, s));
  P1_TRY(pool_alloc(ctx, &L, (size_t)3 * M, s));
  if (M > 0)
    P1_TIMED(ctx, "load_local", s,
             local_load_kernel<<<egrid(plan->ctx, M), THREADS, 0, s>>>(
                 (const double2 *)coords, plan->elements, M, fq, w, (const double2 *)pts, n_qp,
                 L));
  if (n_own > 0)
    P1_TIMED(ctx, "load_gather", s,
             (gather_nodes_kernel<false, 3><<<grid_for(n_own, THREADS), THREADS, 0, s>>>(
                 n_own, rows_of(plan), L, b)));
  pool_free(ctx, L, s);
  pool_free(ctx, w, s);
  pool_free(ctx, pts, s);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}

extern "C" int p1_load_vector_midpoint(const p1_plan *plan, const double *coords,
                                       const double *f_centroid, double *b,
                                       void *stream) {
  P1_REQUIRE(plan, "null plan");
  const int64_t M = plan->M;
  const int n_own = (int)plan->n_own;
  P1_REQUIRE(n_own == 0 || b, "null b");
  P1_REQUIRE(M == 0 || (coords && f_centroid), "null argument");
  p1_ctx *ctx = plan->ctx;
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  P1_TRY(require_elements(plan, "p1_load_vector_midpoint"));
  double *share = nullptr;
  P1_TRY(pool_alloc(ctx, &share, (size_t)M, s));
  if (M > 0)
    local_load_midpoint_kernel<<<egrid(plan->ctx, M), THREADS, 0, s>>>(
        (const double2 *)coords, plan->elements, M, f_centroid, share);
  if (n_own > 0)
    gather_nodes_kernel<true, 1><<<grid_for(n_own, THREADS), THREADS, 0, s>>>(
        n_own, rows_of(plan), share, b);
  pool_free(ctx, share, s);
  P1_CUDA(cudaGetLastError());
  return P1_OK;
}
