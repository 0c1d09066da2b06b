// geometry.cu -- half-edge twins, the reference's edge numbering
// (/root/reference/p1afempy/mesh.py:87-161 provide_geometric_data) and the
// residual estimator (/root/reference/p1afempy/indicators.py:7-86).
//
// The reference matches the two directions of an edge through two scipy
// COO -> find() lexsorts.  Here the match comes from the plan's per-node fans:
// the twin of half-edge (a -> x) of element T is the half-edge (x -> a) of the
// element at node a whose PREVIOUS vertex is x.  Integer outputs are
// bit-identical to the reference; jumps are sums of at most two terms, hence
// order-free, hence bit-identical as well.
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 256;

// flags layout in ctx->d_flags: [0] status bits, [1] #half-edges without twin,
// [2] #forward entries, [3] #backward entries

// half-edge id 3*T + k from a record's element code (T << 2) | k
__device__ __forceinline__ int halfedge_of(int code) { return 3 * (code >> 2) + (code & 3); }

// BACKWARD_ONLY: store only the twins of half-edges a -> x with x < a -- all the
// edge numbering reads.  The 4-byte stores land in other rows' sectors (the
// elements around a node are far apart in the element order): every store that
// misses L2 costs a sector fill and a sector write-back, 11.8 GB of DRAM traffic at
// 64M for 0.8 GB of twins; half the stores, half of that.
template <bool BACKWARD_ONLY>
__global__ void __launch_bounds__(THREADS)
twin_kernel(int n_nodes, Rows rows, int *__restrict__ twin, int *__restrict__ flags) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_nodes) return;
  const int deg = rows.deg(a);
  if (deg == 0) return;
  if (rows.topo && deg > GROUP) return;  // twin_overflow_kernel
  int open_edges = 0;
  bool nonmanifold = false;
  const Rec *rec = rows.topo ? nullptr : rows.ptr(a, deg);
  const int *code = rows.topo ? nullptr : rows.elem(a, deg);
  if (deg <= GROUP) {
    // the row's records and element codes in registers, all pairs unrolled
    int nx[GROUP], pv[GROUP], cd[GROUP];
    if (rows.topo) {
      const int4 *slot = rows.topo + (int64_t)a * GROUP;
#pragma unroll
      for (int i = 0; i < GROUP; ++i) {
        nx[i] = -1 - i; pv[i] = -20 - i; cd[i] = 0;
        if (i < deg) {
          const int4 t = __ldg(slot + i);
          nx[i] = t.x; pv[i] = t.y & NODE_MASK; cd[i] = t.z;
        }
      }
    } else {
      const int4 c0 = __ldg(reinterpret_cast<const int4 *>(code));
      const int4 c1 = __ldg(reinterpret_cast<const int4 *>(code) + 1);
      const int cc[GROUP] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < GROUP; ++i) {
        nx[i] = -1 - i; pv[i] = -20 - i; cd[i] = cc[i];
        if (i < deg) {
          const Rec r = load_rec(rec + i);
          nx[i] = r.nx; pv[i] = rec_prev(r);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      int found = -1, hits = 0;
#pragma unroll
      for (int s = 0; s < GROUP; ++s) {
        if (pv[s] == nx[i]) { found = cd[s]; ++hits; }
        if (s > i && nx[s] == nx[i]) nonmanifold = true;  // same directed edge twice
      }
      if (i < deg) {
        const int he = halfedge_of(cd[i]);
        const bool keep = !BACKWARD_ONLY || nx[i] < a;
        if (hits == 0) {
          if (keep) twin[he] = -1;
          ++open_edges;
        } else {
          // the twin (target -> a) is edge prev(k_s) of the matching element
          if (keep) twin[he] = 3 * (found >> 2) + prev_local(found & 3);
          nonmanifold = nonmanifold || hits > 1;
        }
      }
    }
  } else {
    for (int i = 0; i < deg; ++i) {
      const int target = rec_next(rec[i]);
      int found = -1, hits = 0;
      for (int s = 0; s < deg; ++s)
        if (rec_prev(rec[s]) == target) { found = code[s]; ++hits; }
      const int he = halfedge_of(code[i]);
      const bool keep = !BACKWARD_ONLY || target < a;
      if (hits == 0) {
        if (keep) twin[he] = -1;
        ++open_edges;
      } else {
        if (keep) twin[he] = 3 * (found >> 2) + prev_local(found & 3);
        nonmanifold = nonmanifold || hits > 1;
      }
      for (int s = i + 1; s < deg; ++s)
        if (rec_next(rec[s]) == target) nonmanifold = true;
    }
  }
  if (open_edges) atomicAdd(flags + 1, open_edges);
  if (nonmanifold) atomicOr(flags, FLAG_NONMANIFOLD);
}

// Topology-only plans that know the slot of every half-edge (p1_plan::slot_of):
// the row thread keeps its results in its own 32 bytes (tw[8a + i], coalesced) and
// an element-major pass gathers twin[h] = tw[slot_of[h]].  A gather that misses L2
// costs one sector read; the scattered store it replaces costs a fill and a
// write-back (full twins at 64M: 3.0 ms with scattered stores).
__global__ void __launch_bounds__(THREADS)
twin_rows_kernel(int n_nodes, const unsigned long long *__restrict__ acc,
                 const int4 *__restrict__ topo, int *__restrict__ tw,
                 int *__restrict__ flags) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_nodes) return;
  const int deg = (int)(unsigned)acc[a];
  if (deg == 0 || deg > GROUP) return;  // (more than 8: twin_overflow_kernel)
  int open_edges = 0;
  bool nonmanifold = false;
  int nx[GROUP], pv[GROUP], cd[GROUP], out[GROUP];
  const int4 *slot = topo + (int64_t)a * GROUP;
#pragma unroll
  for (int i = 0; i < GROUP; ++i) {
    nx[i] = -1 - i; pv[i] = -20 - i; cd[i] = 0;
    if (i < deg) {
      const int4 t = __ldg(slot + i);
      nx[i] = t.x; pv[i] = t.y & NODE_MASK; cd[i] = t.z;
    }
  }
#pragma unroll
  for (int i = 0; i < GROUP; ++i) {
    int found = -1, hits = 0;
#pragma unroll
    for (int s = 0; s < GROUP; ++s) {
      if (pv[s] == nx[i]) { found = cd[s]; ++hits; }
      if (s > i && nx[s] == nx[i]) nonmanifold = true;  // same directed edge twice
    }
    out[i] = -1;
    if (i < deg) {
      if (hits == 0) ++open_edges;
      else out[i] = 3 * (found >> 2) + prev_local(found & 3);
      nonmanifold = nonmanifold || hits > 1;
    }
  }
  int4 *dst = reinterpret_cast<int4 *>(tw + (int64_t)a * GROUP);
  dst[0] = make_int4(out[0], out[1], out[2], out[3]);
  if (deg > 4) dst[1] = make_int4(out[4], out[5], out[6], out[7]);
  if (open_edges) atomicAdd(flags + 1, open_edges);
  if (nonmanifold) atomicOr(flags, FLAG_NONMANIFOLD);
}

__global__ void __launch_bounds__(THREADS)
twin_gather_kernel(int64_t n_halfedges, const int *__restrict__ slot_of,
                   const int *__restrict__ tw, int *__restrict__ twin) {
  for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < n_halfedges;
       h += (int64_t)gridDim.x * blockDim.x)
    twin[h] = __ldg(tw + __ldg(slot_of + h));
}

// Rows of a topology-only plan with more than 8 records (few: TSPILL_CAP): the
// thread of the FIRST spilled record of a row matches all of the row's records.
// tw != null: results go to tw[where] (gathered later through slot_of), else
// straight to twin[half-edge] (backward_only: only for half-edges a -> x, x < a).
__global__ void __launch_bounds__(THREADS)
twin_overflow_kernel(Rows rows, int n_rows, int *__restrict__ tw, int *__restrict__ twin,
                     int backward_only, int *__restrict__ flags) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= rows.n_tspill) return;
  const int a = rows.tspill[s].w;
  for (int j = 0; j < s; ++j)
    if (rows.tspill[j].w == a) return;  // not the first of its row
  int open_edges = 0;
  bool nonmanifold = false;
  topo_row_records(rows, a, n_rows, [&](int where, int4 r) {
    const int target = r.x;
    int found = -1, hits = 0, same = 0;
    topo_row_records(rows, a, n_rows, [&](int, int4 q) {
      if ((q.y & NODE_MASK) == target) { found = q.z; ++hits; }
      same += q.x == target;
    });
    nonmanifold = nonmanifold || hits > 1 || same > 1;
    const int result = hits == 0 ? -1 : 3 * (found >> 2) + prev_local(found & 3);
    open_edges += hits == 0;
    if (tw) tw[where] = result;
    else if (!backward_only || target < a) twin[halfedge_of(r.z)] = result;
  });
  if (open_edges) atomicAdd(flags + 1, open_edges);
  if (nonmanifold) atomicOr(flags, FLAG_NONMANIFOLD);
}

// position p of the reference's concatenated half-edge list:
//   p = k*M + m            for element m, local edge k   (I = e[m][k], J = e[m][k+1])
//   p = 3M + r             for boundary row r            (I = b[r][1], J = b[r][0])
// Edge ids = exclusive scan, in p order, of the flag I < J.  The list is cut into
// tiles of ET positions that never straddle two k regions: tile (k, b) covers the
// elements [b*ET, (b+1)*ET) for one k, so the block that reads those elements ONCE
// produces the flags (one byte each) and the counts of its three tiles; the
// boundary rows form a fourth region.  (The first version walked p and re-read the
// element array once per k, wrote 4-byte flags and scanned them in three passes:
// 5.6 GB of traffic at 64M for what needs 2.0 GB.)
constexpr int ET = 2048;  // positions per tile (ET / THREADS per thread)

__device__ __forceinline__ void tile_counts(int c0, int c1, int c2, int *out0, int *out1,
                                            int *out2) {
  __shared__ int cnt[3];
  if (threadIdx.x < 3) cnt[threadIdx.x] = 0;
  __syncthreads();
  c0 = __reduce_add_sync(0xffffffffu, c0);
  c1 = __reduce_add_sync(0xffffffffu, c1);
  c2 = __reduce_add_sync(0xffffffffu, c2);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(cnt, c0); atomicAdd(cnt + 1, c1); atomicAdd(cnt + 2, c2);
  }
  __syncthreads();
  if (threadIdx.x == 0) { *out0 = cnt[0]; if (out1) *out1 = cnt[1]; if (out2) *out2 = cnt[2]; }
}

__global__ void __launch_bounds__(THREADS)
forward_flags_kernel(const int *__restrict__ elements, int64_t M,
                     const int *__restrict__ bnd, int64_t K, int64_t tiles_per_k,
                     unsigned char *__restrict__ fwd, int *__restrict__ tile_sums,
                     int *__restrict__ flags) {
  const int64_t b = blockIdx.x;
  int c0 = 0, c1 = 0, c2 = 0, nb = 0;
  if (b < tiles_per_k) {
    // all loads of the thread's ET / THREADS elements first (one at a time this kernel
    // ran at 1.5 TB/s)
    constexpr int PER = ET / THREADS;
    int e[PER][3];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int64_t m = b * ET + j * THREADS + threadIdx.x;
      e[j][0] = e[j][1] = e[j][2] = 0;
      if (m < M) {
        e[j][0] = __ldg(elements + 3 * m);
        e[j][1] = __ldg(elements + 3 * m + 1);
        e[j][2] = __ldg(elements + 3 * m + 2);
      }
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int64_t m = b * ET + j * THREADS + threadIdx.x;
      if (m >= M) continue;
      const int f0 = e[j][0] < e[j][1], f1 = e[j][1] < e[j][2], f2 = e[j][2] < e[j][0];
      fwd[m] = (unsigned char)f0;
      fwd[M + m] = (unsigned char)f1;
      fwd[2 * M + m] = (unsigned char)f2;
      c0 += f0; c1 += f1; c2 += f2;
      nb += (e[j][1] < e[j][0]) + (e[j][2] < e[j][1]) + (e[j][0] < e[j][2]);
    }
    tile_counts(c0, c1, c2, tile_sums + b, tile_sums + tiles_per_k + b,
                tile_sums + 2 * tiles_per_k + b);
  } else {
    const int64_t t = b - tiles_per_k;
    for (int64_t r = t * ET + threadIdx.x; r < min((t + 1) * (int64_t)ET, K); r += THREADS) {
      const int I = bnd[2 * r + 1], J = bnd[2 * r];
      fwd[3 * M + r] = (unsigned char)(I < J);
      c0 += I < J;
      nb += J < I;
    }
    tile_counts(c0, 0, 0, tile_sums + 3 * tiles_per_k + t, nullptr, nullptr);
  }
  const int nf = __reduce_add_sync(0xffffffffu, c0 + c1 + c2);
  nb = __reduce_add_sync(0xffffffffu, nb);
  if ((threadIdx.x & 31) == 0) {
    if (nf) atomicAdd(flags + 2, nf);
    if (nb) atomicAdd(flags + 3, nb);
  }
}

// num[p] = tile_prefix[tile of p] + exclusive count of the flags before p in its tile
__global__ void __launch_bounds__(THREADS)
edge_ids_kernel(const unsigned char *__restrict__ fwd, int64_t M, int64_t K,
                int64_t tiles_per_k, const int *__restrict__ tile_prefix,
                int *__restrict__ num) {
  __shared__ int warp_sums[THREADS / 32];
  const int64_t b = blockIdx.x;
  const int region = b < 3 * tiles_per_k ? (int)(b / tiles_per_k) : 3;
  const int64_t t = b - (int64_t)region * tiles_per_k;
  const int64_t len = region < 3 ? M : K;
  const int64_t base = (int64_t)region * M;  // region 3 starts at 3M as well
  constexpr int PER = ET / THREADS;
  const int64_t i0 = t * ET + (int64_t)threadIdx.x * PER;  // PER consecutive positions
  int v[PER], sum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = i0 + i < len ? fwd[base + i0 + i] : 0;
    sum += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_sums[w];
  int run = tile_prefix[b] + before + inc - sum;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    if (i0 + i < len) num[base + i0 + i] = run;
    run += v[i];
  }
}

// num[] = exclusive scan of fwd[]: the edge id of every forward entry
__global__ void __launch_bounds__(THREADS)
element_edges_kernel(const int *__restrict__ elements, int64_t M,
                     const int *__restrict__ num, const int *__restrict__ twin,
                     long long *__restrict__ element2edges,
                     long long *__restrict__ edge2nodes, int64_t cap_edges) {
  for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < 3 * M;
       h += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = h / 3;
    const int k = (int)(h - 3 * m);
    const int I = __ldg(elements + 3 * m + k);
    const int J = __ldg(elements + 3 * m + next_local(k));
    const int64_t p = (int64_t)k * M + m;
    if (I < J) {
      const int id = num[p];
      element2edges[h] = id;
      // (more forward than backward half-edges -- boundary edges missing or listed
      // the wrong way round -- is reported by the caller; do not write past the
      // (3M + K + 1) / 2 rows a consistent mesh needs)
      if (id < cap_edges) {
        edge2nodes[2 * (int64_t)id] = I;
        edge2nodes[2 * (int64_t)id + 1] = J;
      }
    } else {
      const int t = twin[h];
      if (t >= 0) {
        const int64_t tm = t / 3;
        const int tk = t - 3 * (int)tm;
        element2edges[h] = num[(int64_t)tk * M + tm];
      } else {
        element2edges[h] = -1;  // set by boundary_edges_kernel, else uncovered
      }
    }
  }
}

// finds the element half-edge b0 -> b1 through the fan of b0
__device__ __forceinline__ int64_t locate_halfedge(int b0, int b1,
                                                   const Rows &rows) {
  const int end = rows.deg(b0);
  if (rows.topo) {
    const int4 *slot = rows.topo + (int64_t)b0 * GROUP;
    for (int i = 0; i < end && i < GROUP; ++i)
      if (slot[i].x == b1) return (int64_t)halfedge_of(slot[i].z);
    if (end > GROUP)
      for (int j = 0; j < rows.n_tspill; ++j) {
        const int4 r = rows.tspill[j];
        if (r.w == b0 && r.x == b1) return (int64_t)halfedge_of(r.z);
      }
    return -1;
  }
  const Rec *rec = rows.ptr(b0, end);
  const int *rec_elem = rows.elem(b0, end);
  for (int i = 0; i < end; ++i)
    if (rec_next(rec[i]) == b1)
      return (int64_t)halfedge_of(rec_elem[i]);
  return -1;
}

__global__ void __launch_bounds__(THREADS)
boundary_numbers_kernel(const int *__restrict__ bnd, int64_t K, int64_t M, int n_nodes,
                        Rows rows, const int *__restrict__ num,
                        long long *__restrict__ element2edges,
                        long long *__restrict__ edge2nodes,
                        long long *__restrict__ boundary2edges, int *__restrict__ flags,
                        int64_t cap_edges) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= K) return;
  const int b0 = bnd[2 * r], b1 = bnd[2 * r + 1];
  if ((unsigned)b0 >= (unsigned)n_nodes || (unsigned)b1 >= (unsigned)n_nodes) {
    atomicOr(flags, FLAG_INDEX_RANGE);
    return;
  }
  const int64_t h = locate_halfedge(b0, b1, rows);
  if (h < 0) {
    atomicOr(flags, FLAG_BOUNDARY_MISS);
    boundary2edges[r] = -1;
    return;
  }
  const int64_t m = h / 3;
  const int k = (int)(h - 3 * m);
  if (b0 < b1) {
    // the element half-edge is the forward one; the appended entry inherits
    boundary2edges[r] = num[(int64_t)k * M + m];
  } else {
    // the appended (reversed) entry is forward and owns the number
    const int id = num[3 * M + r];
    boundary2edges[r] = id;
    if (id < cap_edges) {
      edge2nodes[2 * (int64_t)id] = b1;
      edge2nodes[2 * (int64_t)id + 1] = b0;
    }
    element2edges[h] = id;
  }
}

// ------------------------------------------------------------ estimator
// indicators.py:47-62: dudn[3t+0] = dudn21, [3t+1] = dudn32, [3t+2] = dudn13
__global__ void __launch_bounds__(THREADS)
eta_local_kernel(const double2 *__restrict__ coords, const double *__restrict__ x,
                 const int *__restrict__ elements, int64_t M, double *__restrict__ dudn) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e0 = __ldg(elements + 3 * t), e1 = __ldg(elements + 3 * t + 1),
              e2 = __ldg(elements + 3 * t + 2);
    const double2 z0 = __ldg(coords + e0), z1 = __ldg(coords + e1), z2 = __ldg(coords + e2);
    const double px = z1.x - z0.x, py = z1.y - z0.y;
    const double qx = z2.x - z0.x, qy = z2.y - z0.y;
    const double area2 = 2. * (0.5 * (px * qy - py * qx));
    const double x0 = __ldg(x + e0);
    const double u21 = __ldg(x + e1) - x0, u31 = __ldg(x + e2) - x0;
    const double cx = (qx * u21 - px * u31) / area2;
    const double cy = (qy * u21 - py * u31) / area2;
    const double dn21 = px * cx + py * cy;
    const double dn13 = -(qx * cx + qy * cy);
    const double dn32 = -(dn13 + dn21);
    dudn[3 * t] = dn21;
    dudn[3 * t + 1] = dn32;
    dudn[3 * t + 2] = dn13;
  }
}

// mode 0: Neumann  (indicators.py:69-76)  jump -= |E| g(mid)
// mode 1: Dirichlet (indicators.py:79-80) jump  = 0
__global__ void __launch_bounds__(THREADS)
eta_boundary_kernel(int mode, const int *__restrict__ bnd, int64_t K, int n_nodes,
                    Rows rows, const int *__restrict__ twin,
                    const double *__restrict__ term, double *__restrict__ dudn,
                    int *__restrict__ flags) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= K) return;
  const int b0 = bnd[2 * r], b1 = bnd[2 * r + 1];
  if ((unsigned)b0 >= (unsigned)n_nodes || (unsigned)b1 >= (unsigned)n_nodes) {
    atomicOr(flags, FLAG_INDEX_RANGE);
    return;
  }
  const int64_t h = locate_halfedge(b0, b1, rows);
  if (h < 0) { atomicOr(flags, FLAG_BOUNDARY_MISS); return; }
  if (mode == 0) {
    dudn[h] = dudn[h] - term[r];
  } else {
    dudn[h] = 0.;
    if (twin[h] >= 0) dudn[twin[h]] = 0.;
  }
}

// indicators.py:82-85
__global__ void __launch_bounds__(THREADS)
eta_final_kernel(const double2 *__restrict__ coords, const int *__restrict__ elements,
                 int64_t M, const int *__restrict__ twin, const double *__restrict__ dudn,
                 const double *__restrict__ f_centroid, double *__restrict__ eta) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    double sq[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t h = 3 * t + k;
      const int tw = twin[h];
      double j = dudn[h];
      if (tw >= 0) j = j + __ldg(dudn + tw);
      sq[k] = j * j;
    }
    const int e0 = __ldg(elements + 3 * t), e1 = __ldg(elements + 3 * t + 1),
              e2 = __ldg(elements + 3 * t + 2);
    const double2 z0 = __ldg(coords + e0), z1 = __ldg(coords + e1), z2 = __ldg(coords + e2);
    const double px = z1.x - z0.x, py = z1.y - z0.y;
    const double qx = z2.x - z0.x, qy = z2.y - z0.y;
    const double area2 = 2. * (0.5 * (px * qy - py * qx));
    const double vol = 0.5 * area2 * f_centroid[t];
    eta[t] = (sq[0] + sq[1] + sq[2]) + vol * vol;
  }
}

static inline unsigned egrid(const p1_ctx *ctx, int64_t n) {
  unsigned g = grid_for(n, THREADS);
  unsigned cap = (unsigned)ctx->sm_count * 32u;
  return g < cap ? g : cap;
}

static int read_flags(p1_ctx *ctx, cudaStream_t s) {
  P1_CUDA(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int),
                          cudaMemcpyDeviceToHost, s));
  P1_CUDA(cudaStreamSynchronize(s));
  return P1_OK;
}

static int full_plan_required(const p1_plan *p, const char *who) {
  if (p->row_begin != 0 || p->row_end != p->N || p->own.world != 1) {
    set_error("%s needs a plan over all rows", who);
    return P1_EINVAL;
  }
  return P1_OK;
}

// backward_only: the caller reads twin[h] only for half-edges I -> J with J < I
// (edge numbering); a later caller that needs all of them rebuilds
int ensure_twins(p1_plan *p, cudaStream_t s, bool backward_only) {
  if (p->twin && (!p->twin_partial || backward_only)) return P1_OK;
  p1_ctx *ctx = p->ctx;
  P1_TRY(full_plan_required(p, "half-edge twins"));
  P1_TRY(require_elements(p, "half-edge twins (edge numbering, eta_R)"));
  if (!p->twin) P1_TRY(pool_alloc(ctx, &p->twin, (size_t)(3 * p->M), s));
  p->twin_partial = backward_only;
  P1_CUDA(cudaMemsetAsync(p->ctx->d_flags, 0, 8 * sizeof(int), s));
  if (p->N > 0 && p->topo && p->slot_of) {
    int *tw = nullptr;
    P1_TRY(pool_alloc(ctx, &tw, (size_t)p->N * GROUP + TSPILL_CAP, s));
    p->twin_partial = false;
    P1_TIMED(ctx, "twin", s, {
      twin_rows_kernel<<<grid_for(p->N, THREADS), THREADS, 0, s>>>((int)p->N, p->acc, p->topo, tw,
                                                                  p->ctx->d_flags);
      if (p->n_tspill > 0)
        twin_overflow_kernel<<<grid_for(p->n_tspill, THREADS), THREADS, 0, s>>>(
            rows_of(p), (int)p->N, tw, nullptr, 0, p->ctx->d_flags);
      twin_gather_kernel<<<min(grid_for(3 * p->M, THREADS), (unsigned)(ctx->sm_count * 32)),
                           THREADS, 0, s>>>(3 * p->M, p->slot_of, tw, p->twin);
    });
    pool_free(ctx, tw, s);
  } else if (p->N > 0) {
    if (backward_only)
      P1_TIMED(ctx, "twin", s, twin_kernel<true><<<grid_for(p->N, THREADS), THREADS, 0, s>>>(
          (int)p->N, rows_of(p), p->twin, p->ctx->d_flags));
    else
      P1_TIMED(ctx, "twin", s, twin_kernel<false><<<grid_for(p->N, THREADS), THREADS, 0, s>>>(
          (int)p->N, rows_of(p), p->twin, p->ctx->d_flags));
    if (p->topo && p->n_tspill > 0)
      twin_overflow_kernel<<<grid_for(p->n_tspill, THREADS), THREADS, 0, s>>>(
          rows_of(p), (int)p->n_own, nullptr, p->twin, backward_only ? 1 : 0, p->ctx->d_flags);
  }
  P1_CUDA(cudaGetLastError());
  P1_TRY(read_flags(p->ctx, s));
  if (p->ctx->h_flags[0] & FLAG_NONMANIFOLD) {
    pool_free(ctx, p->twin, s);
    p->twin = nullptr;
    set_error("mesh is not an edge-manifold: a directed edge occurs in two elements");
    return P1_ETOPOLOGY;
  }
  p->n_open = p->ctx->h_flags[1];
  return P1_OK;
}

}  // namespace p1

using namespace p1;

extern "C" int p1_geometric_data(const p1_plan *cplan, const int32_t *boundary_nodes,
                                 const int64_t *boundary_offsets_host, int n_boundaries,
                                 int64_t *element2edges, int64_t *edge2nodes,
                                 int64_t *boundary2edges, int64_t *n_edges_host,
                                 void *stream) {
  P1_REQUIRE(cplan && n_edges_host, "null plan/n_edges_host");
  P1_REQUIRE(n_boundaries >= 0 && (n_boundaries == 0 || boundary_offsets_host),
             "bad boundary list");
  p1_plan *p = const_cast<p1_plan *>(cplan);
  p1_ctx *ctx = p->ctx;
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t M = p->M;
  const int64_t K = n_boundaries ? boundary_offsets_host[n_boundaries] : 0;
  P1_REQUIRE(K == 0 || boundary_nodes, "null boundary_nodes");
  P1_REQUIRE(M == 0 || (element2edges && edge2nodes), "null output");
  if (3 * M + K > 0x7fffffffLL) { set_error("too many half-edges"); return P1_ERANGE; }
  // a one-shot (topology-only) plan: build only the twins the numbering reads
  P1_TRY(ensure_twins(p, s, p->topo != nullptr));
  const int64_t total = 3 * M + K;
  int *num = nullptr, *tile_sums = nullptr;
  unsigned char *fwd = nullptr;
  const int64_t tiles_per_k = (M + ET - 1) / ET, tiles_b = (K + ET - 1) / ET;
  const int64_t n_tiles = 3 * tiles_per_k + tiles_b;
  auto release = [&]() {
    pool_free(ctx, num, s);
    pool_free(ctx, tile_sums, s);
    pool_free(ctx, fwd, s);
  };
  int rc;
  if ((rc = pool_alloc(ctx, &num, (size_t)total + 1, s)) ||
      (rc = pool_alloc(ctx, &tile_sums, (size_t)n_tiles + 1, s)) ||
      (rc = pool_alloc(ctx, &fwd, (size_t)total + 1, s))) {
    release();
    return rc;
  }
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s));
  if (total > 0) {
    P1_TIMED(ctx, "edge_flags", s,
             forward_flags_kernel<<<(unsigned)(tiles_per_k + tiles_b), THREADS, 0, s>>>(
                 p->elements, M, boundary_nodes, K, tiles_per_k, fwd, tile_sums,
                 ctx->d_flags));
    if ((rc = exclusive_scan_i32(ctx, tile_sums, tile_sums, n_tiles, false, s))) {
      release();
      return rc;
    }
    P1_TIMED(ctx, "edge_ids", s,
             edge_ids_kernel<<<(unsigned)n_tiles, THREADS, 0, s>>>(fwd, M, K, tiles_per_k,
                                                                  tile_sums, num));
  }
  if (M > 0)
    P1_TIMED(ctx, "edge_numbers", s,
             element_edges_kernel<<<egrid(ctx, 3 * M), THREADS, 0, s>>>(
                 p->elements, M, num, p->twin, (long long *)element2edges,
                 (long long *)edge2nodes, (total + 1) / 2));
  if (K > 0)
    boundary_numbers_kernel<<<grid_for(K, THREADS), THREADS, 0, s>>>(
        boundary_nodes, K, M, (int)p->N, rows_of(p), num,
        (long long *)element2edges, (long long *)edge2nodes, (long long *)boundary2edges,
        ctx->d_flags, (total + 1) / 2);
  release();
  P1_CUDA(cudaGetLastError());
  P1_TRY(read_flags(ctx, s));
  const int st = ctx->h_flags[0];
  if (st & FLAG_INDEX_RANGE) {
    set_error("p1_geometric_data: boundary node index out of range");
    return P1_EINDEX;
  }
  if (ctx->h_flags[2] != ctx->h_flags[3] || (st & FLAG_BOUNDARY_MISS)) {
    set_error("p1_geometric_data: shape mismatch: %d forward vs %d backward half-edges%s "
              "(every boundary edge must be listed once, oriented like its element)",
              ctx->h_flags[2], ctx->h_flags[3],
              (st & FLAG_BOUNDARY_MISS) ? "; a boundary edge is not an element edge" : "");
    return P1_ETOPOLOGY;
  }
  *n_edges_host = ctx->h_flags[2];
  return P1_OK;
}

static int eta_r_impl(const p1_plan *cplan, const double *coords, const double *x,
                      const int32_t *dirichlet, int64_t n_dirichlet,
                      const int32_t *neumann, int64_t n_neumann,
                      const double *neumann_term, const double *f_centroid,
                      double *eta, void *stream, bool partial);

extern "C" int p1_eta_r(const p1_plan *cplan, const double *coords, const double *x,
                        const int32_t *dirichlet, int64_t n_dirichlet,
                        const int32_t *neumann, int64_t n_neumann,
                        const double *neumann_term, const double *f_centroid,
                        double *eta, void *stream) {
  return eta_r_impl(cplan, coords, x, dirichlet, n_dirichlet, neumann, n_neumann, neumann_term,
                    f_centroid, eta, stream, false);
}

// The estimator on a SUB-mesh (an element block and its ring, p1_block_submesh): open
// half-edges that no boundary lists are the cut through the mesh, not an error; the
// indicators of the elements next to the cut are meaningless and the caller drops them.
extern "C" int p1_eta_r_partial(const p1_plan *cplan, const double *coords, const double *x,
                                const int32_t *dirichlet, int64_t n_dirichlet,
                                const int32_t *neumann, int64_t n_neumann,
                                const double *neumann_term, const double *f_centroid,
                                double *eta, void *stream) {
  return eta_r_impl(cplan, coords, x, dirichlet, n_dirichlet, neumann, n_neumann, neumann_term,
                    f_centroid, eta, stream, true);
}

static int eta_r_impl(const p1_plan *cplan, const double *coords, const double *x,
                      const int32_t *dirichlet, int64_t n_dirichlet,
                      const int32_t *neumann, int64_t n_neumann,
                      const double *neumann_term, const double *f_centroid,
                      double *eta, void *stream, bool partial) {
  P1_REQUIRE(cplan, "null plan");
  p1_plan *p = const_cast<p1_plan *>(cplan);
  p1_ctx *ctx = p->ctx;
  const int64_t M = p->M;
  P1_REQUIRE(M == 0 || (coords && x && f_centroid && eta), "null argument");
  P1_REQUIRE(n_dirichlet >= 0 && n_neumann >= 0, "negative boundary size");
  P1_REQUIRE(n_dirichlet == 0 || dirichlet, "null dirichlet");
  P1_REQUIRE(n_neumann == 0 || (neumann && neumann_term), "null neumann");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  P1_TRY(ensure_twins(p, s));
  if (!partial && p->n_open != n_dirichlet + n_neumann) {
    set_error("p1_eta_r: shape mismatch: the mesh has %lld boundary edges but "
              "dirichlet+neumann list %lld", (long long)p->n_open,
              (long long)(n_dirichlet + n_neumann));
    return P1_ETOPOLOGY;
  }
  if (M == 0) return P1_OK;
  double *dudn = nullptr;
  P1_TRY(pool_alloc(ctx, &dudn, (size_t)(3 * M), s));
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s));
  P1_TIMED(ctx, "eta_local", s,
           eta_local_kernel<<<egrid(ctx, M), THREADS, 0, s>>>((const double2 *)coords, x,
                                                              p->elements, M, dudn));
  if (n_neumann > 0)
    eta_boundary_kernel<<<grid_for(n_neumann, THREADS), THREADS, 0, s>>>(
        0, neumann, n_neumann, (int)p->N, rows_of(p), p->twin,
        neumann_term, dudn, ctx->d_flags);
  if (n_dirichlet > 0)
    eta_boundary_kernel<<<grid_for(n_dirichlet, THREADS), THREADS, 0, s>>>(
        1, dirichlet, n_dirichlet, (int)p->N, rows_of(p), p->twin,
        nullptr, dudn, ctx->d_flags);
  P1_TIMED(ctx, "eta_final", s,
           eta_final_kernel<<<egrid(ctx, M), THREADS, 0, s>>>(
               (const double2 *)coords, p->elements, M, p->twin, dudn, f_centroid, eta));
  pool_free(ctx, dudn, s);
  P1_CUDA(cudaGetLastError());
  P1_TRY(read_flags(ctx, s));
  if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
    set_error("p1_eta_r: boundary node index out of range");
    return P1_EINDEX;
  }
  if (ctx->h_flags[0] & FLAG_BOUNDARY_MISS) {
    set_error("p1_eta_r: a dirichlet/neumann edge is not an element edge "
              "(boundary edges must be oriented like their element)");
    return P1_ETOPOLOGY;
  }
  return P1_OK;
}
