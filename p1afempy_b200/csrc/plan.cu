// plan.cu -- symbolic phase: node -> incident (element, vertex) records with
// their neighbour pair and local-matrix row, and the CSR row pointer.
//
// Replaces, for the assembly hot path, what scipy derives on every
// coo_matrix(...).tocsr() call: coo_tocsr (counting sort by row),
// csr_sort_indices (per-row sort) and csr_sum_duplicates
// (scipy/sparse/_coo.py:368-439, _compressed.py:1072-1130), but on the 3M
// (node, element) incidences instead of the 9M triplets: row i of any P1
// matrix has the columns {i} U {next, prev vertex of every element at i}.
//
// These kernels are bound by the SM's LSU rate for divergent accesses (about one
// 32 B sector per lane per 1.1-1.3 cycles: 201M incidences -> ~0.75 ms per
// access at 64M triangles), not by HBM bytes, so the layout minimises accesses
// per incidence: ONE 64-bit atomic (slot number + closed-fan hash; row a owns
// the fixed slots rec[8a..8a+7], so there is no counting pass, no offset scan)
// and ONE 256-bit store.
//
// Contents: the scatter (four slot kinds: records, topology, payload, slim),
// the spill lists for rows with more than 8 records, split / row lengths, the
// plan-free load vectors (p1_load_vector*_direct), row partitioning for
// multi-GPU, plan_build and the p1_plan_* entry points.
#include <cooperative_groups.h>

#include "plan.cuh"
#include <stdlib.h>
#include <type_traits>
#include <vector>

namespace p1 {

constexpr int THREADS = 256;
#ifndef P1_SCATTER_THREADS
#define P1_SCATTER_THREADS 256
#endif
#ifndef P1_SCATTER_WAVES
#define P1_SCATTER_WAVES 32
#endif
constexpr int SC_THREADS = P1_SCATTER_THREADS;  // scatter block size

// -------------------------------------------------------------- scatter
// The thread that owns element t computes the element's local matrix ONCE
// (value source V) and stores, for every owned vertex, one 32-byte record.
//   OVERFLOW == false: slot = old value of the row's counter; records that find
//                      slots >= 8 go to a spill list (cold block, SpillDesc).
//                      The atomic also accumulates hash(next) - hash(prev): the
//                      nexts and prevs of a closed fan are the same set, so an
//                      interior node ends at 0; a non-zero word marks the rows
//                      whose length is not 1 + #records (open fans = boundary
//                      nodes).  Verified row by row when the matrix is filled.
//   OVERFLOW == true : second pass, only when the spill list was too small: the
//                      rows with more than 8 records, into compact storage.
// Records that do not fit the 8 slots of their row (a node with more than 8
// elements: coarse-mesh vertices of an adaptive mesh, a few per cent of the nodes
// of a Delaunay mesh) go to a spill list; plan_build then gathers every such
// row's records into compact storage.  Only when the list overflows are the
// elements scanned a second time.  The descriptor lives in device memory and is
// read only by the rare thread that spills.
// TOPO 0: the 32-byte record (+ the optional element code array)
//      1: topology-only plan -- one 16-byte slot {next, prev|k, element code, 0}
//      2: vector payload -- one 16-byte slot {element code, 0, v[0] as two words}:
//         what a load vector needs of an incidence, no topology at all
//      3: slim plan -- one 8-byte slot {next, prev|k}; the fill recomputes the values
//      4: one 16-byte slot {next, prev|k, v[1] as two words} for an operator whose
//         local row is (2 v, v, v): the mass matrix (area/6 = 2 (area/12) exactly)
template <int TOPO>
__device__ __forceinline__ void store_slot(Rec *rec, int *rec_elem, int64_t pos, int nx,
                                           int pvk, const double (&v)[3], int code) {
  if constexpr (TOPO == 1) {
    reinterpret_cast<int4 *>(rec)[pos] = make_int4(nx, pvk, code, 0);
    // where the half-edge 3T+k sits: element-major consumers gather per-slot
    // results through it (geometry.cu: twins)
    if (rec_elem) rec_elem[3 * (code >> 2) + (code & 3)] = (int)pos;
  } else if constexpr (TOPO == 2) {
    reinterpret_cast<int4 *>(rec)[pos] =
        make_int4(code, 0, __double2loint(v[0]), __double2hiint(v[0]));
  } else if constexpr (TOPO == 3) {
    reinterpret_cast<int2 *>(rec)[pos] = make_int2(nx, pvk);
  } else if constexpr (TOPO == 4) {
    reinterpret_cast<int4 *>(rec)[pos] =
        make_int4(nx, pvk, __double2loint(v[1]), __double2hiint(v[1]));
  } else {
    Rec r;
    r.nx = nx; r.pvk = pvk;
    r.vd = v[0]; r.vn = v[1]; r.vp = v[2];
    store_rec(rec + pos, r);
    if (rec_elem) rec_elem[pos] = code;
  }
}

template <class V, bool OVERFLOW, int TOPO>
__device__ __forceinline__ void scatter_element(int64_t t, const int (&e)[3], int N,
                                                const Own &own,
                                                unsigned long long *__restrict__ acc,
                                                Rec *__restrict__ rec,
                                                int *__restrict__ rec_elem,
                                                int *__restrict__ cursor, const V &values,
                                                int &mx, bool &bad, int *flags) {
  if ((unsigned)e[0] >= (unsigned)N || (unsigned)e[1] >= (unsigned)N ||
      (unsigned)e[2] >= (unsigned)N) {
    bad = true;
    return;
  }
  mx = max(mx, max(e[0], max(e[1], e[2])));
  bool mine[3];
  int row[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mine[k] = own.owns(e[k], row[k]);
    if (OVERFLOW && mine[k]) mine[k] = (unsigned)acc[row[k]] > (unsigned)GROUP;
  }
  if (!(mine[0] || mine[1] || mine[2])) return;  // not my rows
  double v[3][3];
  double2 z0 = make_double2(0., 0.), z1 = z0, z2 = z0;
  if constexpr (V::needs_xy) {
    z0 = __ldg(values.coords + e[0]);
    z1 = __ldg(values.coords + e[1]);
    z2 = __ldg(values.coords + e[2]);
  }
  values.rows(t, e, z0, z1, z2, v);
  unsigned over = 0;  // vertices whose row had no free slot
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (mine[k]) {
      const int a = row[k];
      const int nx = e[k == 2 ? 0 : k + 1], pv = e[k == 0 ? 2 : k - 1];
      int64_t pos;
      if (OVERFLOW) {
        pos = atomicAdd(cursor + a, 1);
      } else {
        const unsigned d = mix32((unsigned)nx) - mix32((unsigned)pv);
        const unsigned slot =
            (unsigned)atomicAdd(acc + a, ((unsigned long long)d << 32) + 1ull);
        if (slot >= (unsigned)GROUP) {
          over |= 1u << k;
          continue;
        }
        pos = own.pos(a, slot);
      }
      store_slot<TOPO>(rec, rec_elem, pos, nx, pv | (k << NODE_BITS), v[k],
                       (int)((t << 2) | k));
    }
  if (over) {  // cold: a node with more than 8 elements
    if (TOPO == 1 && cursor) {  // `cursor` is the topology spill list
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if ((over >> k) & 1u) {
          const int idx = atomicAdd(flags + 5, 1);
          if (idx < TSPILL_CAP) {
            reinterpret_cast<int4 *>(cursor)[idx] =
                make_int4(e[k == 2 ? 0 : k + 1], e[k == 0 ? 2 : k - 1] | (k << NODE_BITS),
                          (int)((t << 2) | k), row[k]);
            if (rec_elem) rec_elem[3 * t + k] = N * GROUP + idx;
          }
        }
    } else if (TOPO == 2 && cursor) {  // payload spill list: {code, row, value}
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if ((over >> k) & 1u) {
          const int idx = atomicAdd(flags + 5, 1);
          if (idx < TSPILL_CAP)
            reinterpret_cast<int4 *>(cursor)[idx] =
                make_int4((int)((t << 2) | k), row[k], __double2loint(v[k][0]),
                          __double2hiint(v[k][0]));
        }
    } else if (TOPO != 0) {
      atomicOr(flags, FLAG_OVERFLOW);
    } else if (cursor) {  // (not the overflow pass: `cursor` is the spill descriptor)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if ((over >> k) & 1u)
          spill_record(reinterpret_cast<const SpillDesc *>(cursor), flags, row[k],
                       e[k == 2 ? 0 : k + 1], e[k == 0 ? 2 : k - 1] | (k << NODE_BITS), v[k][0],
                       v[k][1], v[k][2], (int)((t << 2) | k));
    }
  }
}

// Phase 1 of a triangle pair: one atomic per DISTINCT owned node; returns the slot of
// each of the 6 references, 5 bits each (30: full row and beyond).  (The kernel keeps
// no pipe busy: LSU wavefronts 55 %, issue 32 %, DRAM 60 % --
// profiles/r02_default_path_64M_ncu_full.csv.)
__device__ __forceinline__ unsigned pair_claim(const int (&id)[6], int N, const Own &own,
                                               unsigned long long *__restrict__ acc, int &mx,
                                               bool &bad) {
  bool valid[2], mine[6];
  int row[6];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    valid[j] = (unsigned)id[3 * j] < (unsigned)N && (unsigned)id[3 * j + 1] < (unsigned)N &&
               (unsigned)id[3 * j + 2] < (unsigned)N;
    bad = bad || !valid[j];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = 3 * j + k;
      row[i] = 0;
      mine[i] = valid[j] && own.owns(id[i], row[i]);
      if (valid[j]) mx = max(mx, id[i]);
    }
  }
  unsigned base[6], packed = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    base[i] = 0;
    bool leader = mine[i];
    unsigned rank = 0;
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (mine[i] && mine[j] && row[j] == row[i]) { leader = false; ++rank; base[i] = base[j]; }
    if (leader) {
      const int ej = i / 3, ek = i - 3 * ej;
      unsigned cnt = 1;
      unsigned h = mix32((unsigned)id[3 * ej + (ek == 2 ? 0 : ek + 1)]) -
                   mix32((unsigned)id[3 * ej + (ek == 0 ? 2 : ek - 1)]);
#pragma unroll
      for (int j = i + 1; j < 6; ++j)
        if (mine[j] && row[j] == row[i]) {
          const int fj = j / 3, fk = j - 3 * fj;
          ++cnt;
          h += mix32((unsigned)id[3 * fj + (fk == 2 ? 0 : fk + 1)]) -
               mix32((unsigned)id[3 * fj + (fk == 0 ? 2 : fk - 1)]);
        }
      base[i] = (unsigned)atomicAdd(acc + row[i], ((unsigned long long)h << 32) + cnt);
    }
    packed |= min(base[i] + rank, 30u) << (5 * i);
  }
  return packed;
}

// Phase 2: the values of the two triangles and one slot store per owned reference.
template <class V, int TOPO>
__device__ __forceinline__ void pair_deliver(int64_t t, const int (&id)[6], unsigned packed, int N,
                                             const Own &own, Rec *__restrict__ rec,
                                             int *__restrict__ rec_elem, const V &values,
                                             int *flags, int (&where)[6],
                                             const SpillDesc *spill = nullptr) {
  // where[i]: slot index of reference i (TOPO 1: the caller stores them, 48 bytes
  // per thread at once -- 4-byte stores per record cost 1.2 ms at 64M)
#pragma unroll
  for (int i = 0; i < 6; ++i) where[i] = -1;
  bool valid[2], touched[2], mine[6];
  int row[6];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    valid[j] = (unsigned)id[3 * j] < (unsigned)N && (unsigned)id[3 * j + 1] < (unsigned)N &&
               (unsigned)id[3 * j + 2] < (unsigned)N;
    touched[j] = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = 3 * j + k;
      row[i] = 0;
      mine[i] = valid[j] && own.owns(id[i], row[i]);
      touched[j] = touched[j] || mine[i];
    }
  }
  if (!(touched[0] || touched[1])) return;
  // (coordinates are NOT shared between the two triangles: lanes of one gather
  // instruction that hit the same sector already cost one LSU wavefront, and the
  // selects cost more than they save -- measured 2.20 ms with, 2.14 ms without)
  double2 z[6];
  if constexpr (V::needs_xy) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      z[i] = make_double2(0., 0.);
      if (touched[i / 3]) z[i] = __ldg(values.coords + id[i]);
    }
  }
  unsigned slot[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) slot[i] = mine[i] ? (packed >> (5 * i)) & 31u : 0xffu;
  bool over = false;  // some owned reference found its row full
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (!touched[j]) continue;
    const int e[3] = {id[3 * j], id[3 * j + 1], id[3 * j + 2]};
    double v[3][3];
    values.rows(t + j, e, z[3 * j], z[3 * j + 1], z[3 * j + 2], v);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = 3 * j + k;
      if (slot[i] >= (unsigned)GROUP) {
        if (mine[i]) over = true;
        continue;
      }
      const int64_t pos = own.pos(row[i], slot[i]);
      where[i] = (int)pos;
      store_slot<TOPO>(rec, TOPO == 1 ? nullptr : rec_elem, pos, e[k == 2 ? 0 : k + 1],
                       e[k == 0 ? 2 : k - 1] | (k << NODE_BITS), v[k],
                       (int)(((t + j) << 2) | k));
    }
  }
  if (over) {  // cold: a node with more than 8 elements; the values are computed again
    if (TOPO == 1 && spill) {  // `spill` is the topology spill list
#pragma unroll
      for (int i = 0; i < 6; ++i)
        if (mine[i] && slot[i] >= (unsigned)GROUP) {
          const int j = i / 3, k = i - 3 * j;
          const int idx = atomicAdd(flags + 5, 1);
          if (idx < TSPILL_CAP) {
            reinterpret_cast<int4 *>(const_cast<SpillDesc *>(spill))[idx] =
                make_int4(id[3 * j + (k == 2 ? 0 : k + 1)],
                          id[3 * j + (k == 0 ? 2 : k - 1)] | (k << NODE_BITS),
                          (int)(((t + j) << 2) | k), row[i]);
            where[i] = N * GROUP + idx;
          }
        }
    } else if (TOPO == 2 && spill) {  // payload spill list: {code, row, value}
      for (int j = 0; j < 2; ++j) {
        if (!touched[j]) continue;
        const int e[3] = {id[3 * j], id[3 * j + 1], id[3 * j + 2]};
        double v[3][3];
        double2 y0 = make_double2(0., 0.), y1 = y0, y2 = y0;
        if constexpr (V::needs_xy) {
          y0 = __ldg(values.coords + e[0]);
          y1 = __ldg(values.coords + e[1]);
          y2 = __ldg(values.coords + e[2]);
        }
        values.rows(t + j, e, y0, y1, y2, v);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = 3 * j + k;
          if (!(mine[i] && slot[i] >= (unsigned)GROUP)) continue;
          const int idx = atomicAdd(flags + 5, 1);
          if (idx < TSPILL_CAP)
            reinterpret_cast<int4 *>(const_cast<SpillDesc *>(spill))[idx] =
                make_int4((int)(((t + j) << 2) | k), row[i], __double2loint(v[k][0]),
                          __double2hiint(v[k][0]));
        }
      }
    } else if (TOPO != 0) {
      atomicOr(flags, FLAG_OVERFLOW);
    } else if (spill) {
      for (int j = 0; j < 2; ++j) {
        if (!touched[j]) continue;
        const int e[3] = {id[3 * j], id[3 * j + 1], id[3 * j + 2]};
        double v[3][3];
        double2 y0 = make_double2(0., 0.), y1 = y0, y2 = y0;
        if constexpr (V::needs_xy) {  // gathered again: keeping z[] alive costs 18 registers
          y0 = __ldg(values.coords + e[0]);
          y1 = __ldg(values.coords + e[1]);
          y2 = __ldg(values.coords + e[2]);
        }
        values.rows(t + j, e, y0, y1, y2, v);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = 3 * j + k;
          if (mine[i] && slot[i] >= (unsigned)GROUP)
            spill_record(spill, flags, row[i], e[k == 2 ? 0 : k + 1],
                         e[k == 0 ? 2 : k - 1] | (k << NODE_BITS), v[k][0], v[k][1], v[k][2],
                         (int)(((t + j) << 2) | k));
        }
      }
    }
  }
}

// Two consecutive triangles at once: NVB siblings share two of their three
// vertices, so the 6 vertex references are grouped in registers -- one atomic
// per DISTINCT node (typically 4 instead of 6).  The kernel is bound by LSU
// operations per lane; the 15 comparisons are free.
// (Requesting the six coordinates BEFORE the atomics, so that the two latencies overlap --
// same 80 registers -- was measured at 64M: stiffness scatter 2.19 -> 2.30 ms.  Like the
// two experiments noted in scatter_body, more requests in flight per warp make it slower:
// the kernel sits at the rate at which the memory system takes divergent requests, ~8 per
// triangle, not at their latency.)
template <class V, int TOPO>
__device__ __forceinline__ void scatter_pair(int64_t t, const int (&id)[6], int N,
                                             const Own &own,
                                             unsigned long long *__restrict__ acc,
                                             Rec *__restrict__ rec,
                                             int *__restrict__ rec_elem, const V &values,
                                             int &mx, bool &bad, int *flags, int (&where)[6],
                                             const SpillDesc *spill = nullptr) {
  const unsigned packed = pair_claim(id, N, own, acc, mx, bad);
  pair_deliver<V, TOPO>(t, id, packed, N, own, rec, rec_elem, values, flags, where, spill);
}

// `vec` != 0: elements is 16-byte aligned; a thread then reads 4 triangles =
// 12 indices as three 128-bit loads (the scalar form reads 12 B at stride 12 B
// and spends 0.45 ms just scanning 64M triangles).
// (forcing more resident blocks was measured: 40 or 32 registers spill and lose)
// PAIRS: group the references of two consecutive triangles (scatter_pair).  Pays
// when most references are owned (one GPU: 2.58 -> 2.20 ms at 64M); with 1/8 of the
// rows owned there is little to group and the extra registers cost more (0.76 ->
// 1.01 ms), so multi-GPU plans use the plain form.
// (register allocation of the PAIRS form, measured at 64M: 80 registers 2.20 ms --
// what ptxas picks without a min-blocks hint --, 99 registers 2.74 ms, 64 with
// spills 2.35 ms, 48 with spills 2.82 ms; 72 registers with blocks of 128 threads, 7 per
// SM: 2.32 ms; blocks of 128 threads at 80 registers: 2.34 ms)
template <class V, bool OVERFLOW, bool PAIRS, int TOPO>
__device__ __forceinline__ void
scatter_body(const int *__restrict__ elements, int64_t M, int N, const Own &own, int vec,
             unsigned long long *__restrict__ acc, Rec *__restrict__ rec,
             int *__restrict__ rec_elem, int *__restrict__ cursor, const V &values,
             int *__restrict__ flags) {
  int mx = -1;
  bool bad = false;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t quads = vec ? M / 4 : 0;
  const int4 *e4 = reinterpret_cast<const int4 *>(elements);
  // (prefetching the next quad's indices -- the kernel waits on memory: long-scoreboard
  // stalls 13 per issue at 24 warps per SM -- costs 36 registers, i.e. a resident block:
  // 2.13 -> 2.60 ms at 64M)
  for (int64_t g = tid; g < quads; g += stride) {
    const int4 p = __ldcs(e4 + 3 * g), q = __ldcs(e4 + 3 * g + 1), r = __ldcs(e4 + 3 * g + 2);  // streamed
    if constexpr (!OVERFLOW && PAIRS) {
      const int ab[6] = {p.x, p.y, p.z, p.w, q.x, q.y}, cd[6] = {q.z, q.w, r.x, r.y, r.z, r.w};
      int w0[6], w1[6];
      const SpillDesc *sp = reinterpret_cast<const SpillDesc *>(cursor);
      // (claiming both pairs before delivering the first -- eight atomics in flight per
      // thread instead of four -- was measured at 64M: stiffness scatter 2.13 -> 2.51 ms,
      // mass 2.19 -> 2.16: the atomics' latency is not what the kernel waits for)
      scatter_pair<V, TOPO>(4 * g, ab, N, own, acc, rec, rec_elem, values, mx, bad, flags, w0, sp);
      scatter_pair<V, TOPO>(4 * g + 2, cd, N, own, acc, rec, rec_elem, values, mx, bad, flags, w1,
                            sp);
      if (TOPO == 1 && rec_elem) {  // slot of the 12 half-edges: 3 x 16 bytes, aligned
        int4 *dst = reinterpret_cast<int4 *>(rec_elem + 12 * g);
        dst[0] = make_int4(w0[0], w0[1], w0[2], w0[3]);
        dst[1] = make_int4(w0[4], w0[5], w1[0], w1[1]);
        dst[2] = make_int4(w1[2], w1[3], w1[4], w1[5]);
      }
      continue;
    }
    const int ea[3] = {p.x, p.y, p.z}, eb[3] = {p.w, q.x, q.y}, ec[3] = {q.z, q.w, r.x},
              ed[3] = {r.y, r.z, r.w};
    scatter_element<V, OVERFLOW, TOPO>(4 * g, ea, N, own, acc, rec, rec_elem, cursor, values, mx, bad, flags);
    scatter_element<V, OVERFLOW, TOPO>(4 * g + 1, eb, N, own, acc, rec, rec_elem, cursor, values, mx, bad, flags);
    scatter_element<V, OVERFLOW, TOPO>(4 * g + 2, ec, N, own, acc, rec, rec_elem, cursor, values, mx, bad, flags);
    scatter_element<V, OVERFLOW, TOPO>(4 * g + 3, ed, N, own, acc, rec, rec_elem, cursor, values, mx, bad, flags);
  }
  for (int64_t t = 4 * quads + tid; t < M; t += stride) {
    const int e[3] = {__ldg(elements + 3 * t), __ldg(elements + 3 * t + 1),
                      __ldg(elements + 3 * t + 2)};
    scatter_element<V, OVERFLOW, TOPO>(t, e, N, own, acc, rec, rec_elem, cursor, values, mx, bad, flags);
  }
  if (!OVERFLOW) {
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx >= 0) atomicMax(flags + 1, mx);
    if (bad) atomicOr(flags, FLAG_INDEX_RANGE);
  }
}

template <class V, bool OVERFLOW, bool PAIRS, int TOPO>
__global__ void __launch_bounds__(SC_THREADS)
scatter_kernel(const int *__restrict__ elements, int64_t M, int N, Own own, int vec,
               unsigned long long *__restrict__ acc, Rec *__restrict__ rec,
               int *__restrict__ rec_elem, int *__restrict__ cursor, V values,
               int *__restrict__ flags) {
  scatter_body<V, OVERFLOW, PAIRS, TOPO>(elements, M, N, own, vec, acc, rec, rec_elem, cursor,
                                         values, flags);
}

// The same kernel with room for two blocks per SM stated: GeneralValues keeps 16 coefficient
// loads in flight and wants the registers for it -- without the hint ptxas settles on 103
// registers and the scatter takes 9.1 ms at 64M, with it on 120+ and 8.0 ms.  (A hint on
// the other value sources changes THEIR allocation, e.g. stiffness 80 -> 118 registers.)
template <class V, bool OVERFLOW, bool PAIRS, int TOPO>
__global__ void __launch_bounds__(SC_THREADS, 2)
scatter_kernel_wide(const int *__restrict__ elements, int64_t M, int N, Own own, int vec,
                    unsigned long long *__restrict__ acc, Rec *__restrict__ rec,
                    int *__restrict__ rec_elem, int *__restrict__ cursor, V values,
                    int *__restrict__ flags) {
  scatter_body<V, OVERFLOW, PAIRS, TOPO>(elements, M, N, own, vec, acc, rec, rec_elem, cursor,
                                         values, flags);
}

template <class V, bool OVERFLOW>
static void launch_one(bool pairs, unsigned grid, cudaStream_t s, const int *elements, int64_t M,
                       int N, Own own, int vec, unsigned long long *acc, Rec *rec,
                       int *rec_elem, int *cursor, V values, int *fl) {
  if constexpr (std::is_same<V, GeneralValues>::value && !OVERFLOW) {
    scatter_kernel_wide<V, OVERFLOW, false, 0><<<grid, SC_THREADS, 0, s>>>(
        elements, M, N, own, vec, acc, rec, rec_elem, cursor, values, fl);
    return;
  }
  if (pairs && !OVERFLOW)
    scatter_kernel<V, OVERFLOW, true, 0><<<grid, SC_THREADS, 0, s>>>(
        elements, M, N, own, vec, acc, rec, rec_elem, cursor, values, fl);
  else
    scatter_kernel<V, OVERFLOW, false, 0><<<grid, SC_THREADS, 0, s>>>(
        elements, M, N, own, vec, acc, rec, rec_elem, cursor, values, fl);
}

// slots instead of records (MODE 1: topology, 2: vector payload, 3: slim, 4: (2v, v, v) rows)
template <int MODE, class V>
static void launch_scatter_slots(p1_ctx *ctx, unsigned grid, cudaStream_t s, const int *elements,
                                 int64_t M, int N, Own own, unsigned long long *acc, void *slots_,
                                 V values, bool allow_pairs, int *slot_of = nullptr,
                                 int4 *tspill = nullptr, int *fl = nullptr) {
  const int vec = ((uintptr_t)elements & 15) == 0;
  bool pairs = allow_pairs && own.world == 1 && own.re - own.rb == N;
  Rec *slots = reinterpret_cast<Rec *>(slots_);
  if (!fl) fl = ctx->d_flags;
  if (pairs)
    scatter_kernel<V, false, true, MODE><<<grid, SC_THREADS, 0, s>>>(
        elements, M, N, own, vec, acc, slots, slot_of, reinterpret_cast<int *>(tspill), values,
        fl);
  else
    scatter_kernel<V, false, false, MODE><<<grid, SC_THREADS, 0, s>>>(
        elements, M, N, own, vec, acc, slots, slot_of, reinterpret_cast<int *>(tspill), values,
        fl);
}

template <bool OVERFLOW>
static void launch_scatter(p1_ctx *ctx, unsigned grid, cudaStream_t s, int op,
                           const int *elements, int64_t M, int N, Own own,
                           unsigned long long *acc, Rec *rec, int *rec_elem, int *cursor,
                           const double *coords, const double *local,
                           const p1_quad_args *qa, const double *d_weights,
                           const double *d_points, int *fl = nullptr) {
  const double2 *xy = (const double2 *)coords;
  if (!fl) fl = ctx->d_flags;
  // 4 triangles per thread; not for the source that streams four [n_qp x M] arrays,
  // which wants consecutive lanes on consecutive triangles (measured at 64M: 24.5 ms
  // with, 12.5 ms without; the unfused route through p1_local_general_stiffness is
  // 12.1 ms, so the drop-in layer uses that one)
  const int vec = ((uintptr_t)elements & 15) == 0 && op != P1_OP_GENERAL;
  // all rows owned (the weighted mass gains too since its values are prefetched: 5.8 -> 5.3 ms)
  const bool pairs = own.world == 1 && own.re - own.rb == N && op != P1_OP_GENERAL;
  if (op == P1_OP_GENERAL)
    launch_one<GeneralValues, OVERFLOW>(pairs, grid, s, elements, M, N, own, vec, acc, rec, rec_elem,
            cursor, GeneralValues{xy, qa->a11q, qa->a12q, qa->a21q, qa->a22q, d_weights, qa->n_qp, M}, fl);
  else if (op == P1_OP_WEIGHTED)
    launch_one<WeightedValues, OVERFLOW>(pairs, grid, s, elements, M, N, own, vec, acc, rec, rec_elem,
            cursor, WeightedValues{xy, qa->phiq, d_weights, (const double2 *)d_points, qa->n_qp, M}, fl);
  else if (op == P1_OP_STIFFNESS)
    launch_one<StiffnessValues, OVERFLOW>(pairs, grid, s, elements, M, N, own, vec, acc, rec, rec_elem,
            cursor, StiffnessValues{xy}, fl);
  else if (op == P1_OP_MASS)
    launch_one<MassValues, OVERFLOW>(pairs, grid, s, elements, M, N, own, vec, acc, rec, rec_elem,
            cursor, MassValues{xy}, fl);
  else if (op == P1_OP_LOCAL)
    launch_one<LocalValues, OVERFLOW>(pairs, grid, s, elements, M, N, own, vec, acc, rec, rec_elem,
            cursor, LocalValues{local}, fl);
  else
    launch_one<NoValues, OVERFLOW>(pairs, grid, s, elements, M, N, own, vec, acc, rec, rec_elem,
            cursor, NoValues{}, fl);
}

// ---------------------------------------------------------------- split
// counters: [0] n_big, [1] n_slow, [2] rows with more than 8 records
__global__ void __launch_bounds__(THREADS)
split_kernel(int n_own, const unsigned long long *__restrict__ acc,
             int *__restrict__ rowlen, int *__restrict__ slow_rows,
             int *__restrict__ big_rows, int *__restrict__ counters) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_own) return;
  const unsigned long long v = acc[a];
  const int deg = (int)(unsigned)v;
  if (deg == 0) { rowlen[a] = 0; return; }
  if ((unsigned)(v >> 32) == 0u && deg <= GROUP) { rowlen[a] = 1 + deg; return; }
  rowlen[a] = 0;  // written by rowlen_slow_kernel / rowlen_big_kernel
  if (deg > GROUP) atomicAdd(counters + 2, 1);
  if (deg > DEG_BIG) big_rows[atomicAdd(counters, 1)] = a;
  else slow_rows[atomicAdd(counters + 1, 1)] = a;
}

// exact plans: 8 lanes per row, all-pairs (plan.cuh: analyse_row)
__global__ void __launch_bounds__(THREADS)
rowlen_kernel(int n_own, Own own, Rows rows, int *__restrict__ rowlen,
              int *__restrict__ slow_rows, int *__restrict__ big_rows,
              int *__restrict__ counters) {
  __shared__ int2 xch[THREADS];
  const int lane = threadIdx.x & 31, lane8 = lane & (GROUP - 1);
  const int shift = lane & ~(GROUP - 1);
  const unsigned gmask = 0xffu << shift;
  const int a = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GROUP);
  if (a >= n_own) return;
  const int deg = rows.deg(a);
  if (deg == 0) {
    if (lane8 == 0) rowlen[a] = 0;
    return;
  }
  const RowLane r = analyse_row(own.global(a), rows.rec + (int64_t)a * GROUP, deg, gmask, lane8,
                                shift, xch);
  if (lane8 != 0) return;
  if (r.simple && r.len == 1 + deg) {  // closed, repeat-free fan: the fast fill takes it
    rowlen[a] = r.len;
    return;
  }
  rowlen[a] = 0;
  if (deg > GROUP) atomicAdd(counters + 2, 1);
  if (deg > DEG_BIG) big_rows[atomicAdd(counters, 1)] = a;
  else slow_rows[atomicAdd(counters + 1, 1)] = a;
}

__global__ void __launch_bounds__(THREADS)
overflow_count_kernel(int n_own, const unsigned long long *__restrict__ acc,
                      int *__restrict__ cnt) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_own) return;
  const int deg = (int)(unsigned)acc[a];
  cnt[a] = deg > GROUP ? deg : 0;
}

// rows with more than 8 records (they are all in the slow / big row lists): the
// row takes the next free range of the compact storage (the order of the ranges
// does not matter), its 8 slot records move to the head of it (cursor[a] then
// points behind them) ...
__global__ void __launch_bounds__(THREADS)
overflow_heads_kernel(const int *__restrict__ list, int n_list,
                      const unsigned long long *__restrict__ acc,
                      const Rec *__restrict__ rec, const int *__restrict__ rec_elem,
                      int *__restrict__ ovf_off, Rec *__restrict__ ovf_rec,
                      int *__restrict__ ovf_elem, int *__restrict__ cursor,
                      int *__restrict__ next_free) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_list) return;
  const int a = list[idx];
  const int deg = (int)(unsigned)acc[a];
  if (deg <= GROUP) return;
  const int off = atomicAdd(next_free, deg);
  ovf_off[a] = off;
  for (int i = 0; i < GROUP; ++i) {
    ovf_rec[off + i] = rec[(int64_t)a * GROUP + i];
    if (ovf_elem) ovf_elem[off + i] = rec_elem[(int64_t)a * GROUP + i];
  }
  cursor[a] = off + GROUP;
}
// ... and the spilled ones follow in arrival order
__global__ void __launch_bounds__(THREADS)
overflow_place_kernel(int n_spill, SpillDesc sp, int *__restrict__ cursor,
                      Rec *__restrict__ ovf_rec, int *__restrict__ ovf_elem) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_spill) return;
  const int pos = atomicAdd(cursor + sp.row[i], 1);
  ovf_rec[pos] = sp.rec[i];
  if (ovf_elem) ovf_elem[pos] = sp.elem[i];
}

// distinct values of {self} U {next_i} U {prev_i}, any number of records;
// the few contiguous records stay in L1 while they are compared
__global__ void __launch_bounds__(THREADS)
rowlen_slow_kernel(const int *__restrict__ slow_rows, const int *__restrict__ n_slow_dev,
                   Own own, Rows rows, int *__restrict__ rowlen) {
  // the list length is read on the device: the host has not synchronised yet
  const int n_slow = *n_slow_dev;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_slow;
       idx += gridDim.x * blockDim.x) {
  const int a = slow_rows[idx];
  const int deg = rows.deg(a);
  if (deg > GROUP && !rows.ovf_rec) continue;  // redone after the overflow pass
  const int self = own.global(a);
  int n = 1;
  for (int i = 0; i < deg; ++i) {
    const int2 ei = rows.edge(a, deg, i);
    const int x = ei.x, y = ei.y & NODE_MASK;
    bool fx = x != self, fy = y != self && y != x;
    for (int j = 0; j < deg; ++j) {
      const int2 ej = rows.edge(a, deg, j);
      const int qx = ej.x, qy = ej.y & NODE_MASK;
      fx = fx && !(j < i && qx == x);             // an earlier next equals it
      fy = fy && qx != y && !(j < i && qy == y);  // any next / an earlier prev
    }
    n += fx + fy;
  }
  rowlen[a] = n;
  }
}

// Rows with more than DEG_BIG records: one block per row, O(deg^2) without
// any capacity limit (pathological fans only; never on a shape-regular mesh).
// Rank sort by (next, prev, k): the canonical order of the big path.
__device__ __forceinline__ unsigned long long big_key(const Rec &r) {
  return ((unsigned long long)(unsigned)r.nx << 32) | (unsigned)r.pvk;
}
__global__ void __launch_bounds__(THREADS)
sort_big_kernel(const int *__restrict__ big_rows, Rows rows, Rec *__restrict__ tmp,
                int *__restrict__ tmp_elem) {
  const int a = big_rows[blockIdx.x];
  const int deg = rows.deg(a);
  Rec *row = rows.ptr(a, deg);
  int *el = rows.ovf_elem ? rows.elem(a, deg) : nullptr;
  const int64_t off = rows.ovf_off[a];
  for (int i = threadIdx.x; i < deg; i += blockDim.x) {
    const Rec v = row[i];
    const unsigned long long key = big_key(v);
    int rank = 0;
    for (int j = 0; j < deg; ++j) {
      const unsigned long long kj = big_key(row[j]);
      rank += kj < key || (kj == key && j < i);  // stable for identical keys
    }
    tmp[off + rank] = v;
    if (el) tmp_elem[off + rank] = el[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < deg; i += blockDim.x) {
    row[i] = tmp[off + i];
    if (el) el[i] = tmp_elem[off + i];
  }
}

// candidate j of a big row: 0 = the row itself, 1+2i = next_i, 2+2i = prev_i
__device__ __forceinline__ int big_candidate(int self, const Rec *nb, int j) {
  if (j == 0) return self;
  const Rec p = nb[(j - 1) >> 1];
  return ((j - 1) & 1) ? rec_prev(p) : rec_next(p);
}

__global__ void __launch_bounds__(THREADS)
rowlen_big_kernel(const int *__restrict__ big_rows, Own own, Rows rows,
                  unsigned char *__restrict__ first, int *__restrict__ rowlen) {
  const int a = big_rows[blockIdx.x];
  const int deg = rows.deg(a);
  const Rec *nb = rows.ptr(a, deg);
  unsigned char *fl = first + 2 * (int64_t)rows.ovf_off[a] + a;
  const int ncand = 2 * deg + 1;
  int mine = 0;
  for (int j = threadIdx.x; j < ncand; j += blockDim.x) {
    const int c = big_candidate(own.global(a), nb, j);
    bool is_first = true;
    for (int i = 0; i < j && is_first; ++i)
      is_first = big_candidate(own.global(a), nb, i) != c;
    fl[j] = is_first;
    mine += is_first;
  }
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  atomicAdd(&total, mine);
  __syncthreads();
  if (threadIdx.x == 0) rowlen[a] = total;
}


// ------------------------------------------------ load vectors without a plan
// b = bincount(elements, weights = local loads) in the reference's summation
// order needs, per node, only (element code, value) of every incidence: the
// scatter writes exactly that into the node's 8 slots and one thread per node
// sorts the codes in registers and adds.  No topology, no row pointer, and no
// random gather of the local loads (at 64M that gather read 13 GB for 1.6 GB of
// values: the three vertices of a triangle live in rows that are far apart).
struct LoadQuadValues {  // solvers.py:564-591, the arithmetic of local_load_kernel
  static constexpr bool needs_xy = true;
  const double2 *__restrict__ coords;
  const double *__restrict__ fq;       // [n_qp x M]
  const double *__restrict__ weights;  // device, n_qp
  const double2 *__restrict__ points;  // device, n_qp (eta, xi)
  int n_qp;
  int64_t M;
  __device__ __forceinline__ void rows(int64_t t, const int *, double2 z0, double2 z1,
                                       double2 z2, double (*v)[3]) const {
    const double px = z1.x - z0.x, py = z1.y - z0.y;
    const double qx = z2.x - z0.x, qy = z2.y - z0.y;
    const double area = 0.5 * (px * qy - py * qx);
    double l0 = 0., l1 = 0., l2 = 0.;
    for (int q = 0; q < n_qp; ++q) {
      const double2 pt = __ldg(points + q);
      const double f = __ldg(fq + (int64_t)q * M + t);
      const double s = 2. * __ldg(weights + q) * area;
      l0 += s * (1. - pt.x - pt.y) * f;
      l1 += s * pt.x * f;
      l2 += s * pt.y * f;
    }
    v[0][0] = l0; v[1][0] = l1; v[2][0] = l2;
  }
};
struct LoadMidpointValues {  // solvers.py:415-424: share = area4*fsT/12. for all three
  static constexpr bool needs_xy = true;
  const double2 *__restrict__ coords;
  const double *__restrict__ f_centroid;
  __device__ __forceinline__ void rows(int64_t t, const int *, double2 z0, double2 z1,
                                       double2 z2, double (*v)[3]) const {
    const double px = z1.x - z0.x, py = z1.y - z0.y;
    const double qx = z2.x - z0.x, qy = z2.y - z0.y;
    const double area4 = 4. * (0.5 * (px * qy - py * qx));
    const double share = area4 * __ldg(f_centroid + t) / 12.;
    v[0][0] = v[1][0] = v[2][0] = share;
  }
};

// The local loads of FOUR consecutive triangles in registers (the [n_qp x M]
// values of f are read as one 32-byte load per thread and point: 1 KB per warp
// instruction), handed to scatter_pair two triangles at a time.
struct PairLoads {
  static constexpr bool needs_xy = false;
  double l0[3], l1[3];  // the even and the odd triangle of the pair
  __device__ __forceinline__ void rows(int64_t t, const int *, double2, double2, double2,
                                       double (*v)[3]) const {
    const bool odd = (t & 1) != 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      v[k][0] = odd ? l1[k] : l0[k];
      v[k][1] = v[k][2] = 0.;
    }
  }
};

// points in flight per thread and resident blocks of the cubature kernel, measured at
// 64M / 16 points: (4, unhinted = 126 registers) 3.35 ms, (2, unhinted) 3.64, (8, unhinted)
// 5.01, (4, 3 blocks) 3.29, (2, 3 blocks) 3.20.  The midpoint kernel keeps the compiler's
// own choice (66 registers; a min-blocks hint of 1 made it 2.19 instead of 1.85 ms).
#ifndef P1_LOAD_MIN_BLOCKS
#define P1_LOAD_MIN_BLOCKS 3
#endif
#ifndef P1_LOAD_PF
#define P1_LOAD_PF 2
#endif
template <bool MIDPOINT>
__device__ __forceinline__ void load_scatter_body(const int *__restrict__ elements, int64_t M, int N, Own own,
                    const double2 *__restrict__ coords, const double *__restrict__ f,
                    const double *__restrict__ weights, const double2 *__restrict__ points,
                    int n_qp, int wide, unsigned long long *__restrict__ acc,
                    Rec *__restrict__ slots, int4 *__restrict__ spill_list,
                    int *__restrict__ flags) {
  int mx = -1;
  bool bad = false;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t quads = M / 4;
  const int4 *e4 = reinterpret_cast<const int4 *>(elements);
  for (int64_t g = tid; g < quads; g += stride) {
    const int4 p = __ldcs(e4 + 3 * g), q = __ldcs(e4 + 3 * g + 1), r = __ldcs(e4 + 3 * g + 2);  // streamed
    const int id[12] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w, r.x, r.y, r.z, r.w};
    double area[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      area[j] = 0.;
      if ((unsigned)id[3 * j] < (unsigned)N && (unsigned)id[3 * j + 1] < (unsigned)N &&
          (unsigned)id[3 * j + 2] < (unsigned)N) {
        const double2 z0 = __ldg(coords + id[3 * j]), z1 = __ldg(coords + id[3 * j + 1]),
                      z2 = __ldg(coords + id[3 * j + 2]);
        const double px = z1.x - z0.x, py = z1.y - z0.y;
        const double qx = z2.x - z0.x, qy = z2.y - z0.y;
        area[j] = 0.5 * (px * qy - py * qx);
      }
    }
    double l[4][3];
    if (MIDPOINT) {
      double fv[4];
      if (wide) {
        const double4 w4 = *reinterpret_cast<const double4 *>(f + 4 * g);
        fv[0] = w4.x; fv[1] = w4.y; fv[2] = w4.z; fv[3] = w4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) fv[j] = __ldg(f + 4 * g + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double area4 = 4. * area[j];
        l[j][0] = l[j][1] = l[j][2] = area4 * fv[j] / 12.;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) l[j][0] = l[j][1] = l[j][2] = 0.;
      // four points at a time: 128 bytes per thread in flight (the sums themselves
      // stay in the reference's order, point after point)
      for (int qp0 = 0; qp0 < n_qp; qp0 += P1_LOAD_PF) {
        double fv[P1_LOAD_PF][4];
#pragma unroll
        for (int u = 0; u < P1_LOAD_PF; ++u) {
          const double *fp = f + (int64_t)(qp0 + u) * M + 4 * g;
          fv[u][0] = fv[u][1] = fv[u][2] = fv[u][3] = 0.;
          if (qp0 + u < n_qp) {
            if (wide) {
              const double4 w4 = *reinterpret_cast<const double4 *>(fp);
              fv[u][0] = w4.x; fv[u][1] = w4.y; fv[u][2] = w4.z; fv[u][3] = w4.w;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) fv[u][j] = __ldg(fp + j);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < P1_LOAD_PF; ++u) {
          if (qp0 + u < n_qp) {
            const double2 pt = __ldg(points + qp0 + u);
            const double w = __ldg(weights + qp0 + u);
            const double h0 = 1. - pt.x - pt.y;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double s = 2. * w * area[j];
              l[j][0] += s * h0 * fv[u][j];
              l[j][1] += s * pt.x * fv[u][j];
              l[j][2] += s * pt.y * fv[u][j];
            }
          }
        }
      }
    }
    const int ab[6] = {id[0], id[1], id[2], id[3], id[4], id[5]},
              cd[6] = {id[6], id[7], id[8], id[9], id[10], id[11]};
    const PairLoads v01{{l[0][0], l[0][1], l[0][2]}, {l[1][0], l[1][1], l[1][2]}};
    const PairLoads v23{{l[2][0], l[2][1], l[2][2]}, {l[3][0], l[3][1], l[3][2]}};
    int unused[6];
    const SpillDesc *sp = reinterpret_cast<const SpillDesc *>(spill_list);
    scatter_pair<PairLoads, 2>(4 * g, ab, N, own, acc, slots, nullptr, v01, mx, bad, flags, unused,
                               sp);
    scatter_pair<PairLoads, 2>(4 * g + 2, cd, N, own, acc, slots, nullptr, v23, mx, bad, flags,
                               unused, sp);
  }
  for (int64_t t = 4 * quads + tid; t < M; t += stride) {  // the last M % 4 triangles
    const int e[3] = {__ldg(elements + 3 * t), __ldg(elements + 3 * t + 1),
                      __ldg(elements + 3 * t + 2)};
    if (MIDPOINT)
      scatter_element<LoadMidpointValues, false, 2>(t, e, N, own, acc, slots, nullptr,
                                                    reinterpret_cast<int *>(spill_list),
                                                    LoadMidpointValues{coords, f}, mx, bad, flags);
    else
      scatter_element<LoadQuadValues, false, 2>(t, e, N, own, acc, slots, nullptr,
                                                reinterpret_cast<int *>(spill_list),
                                                LoadQuadValues{coords, f, weights, points, n_qp, M},
                                                mx, bad, flags);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx >= 0) atomicMax(flags + 1, mx);
  if (bad) atomicOr(flags, FLAG_INDEX_RANGE);
}

__global__ void __launch_bounds__(SC_THREADS, P1_LOAD_MIN_BLOCKS)
load_scatter_cubature_kernel(const int *__restrict__ elements, int64_t M, int N, Own own,
                             const double2 *__restrict__ coords, const double *__restrict__ f,
                             const double *__restrict__ weights,
                             const double2 *__restrict__ points, int n_qp, int wide,
                             unsigned long long *__restrict__ acc, Rec *__restrict__ slots,
                             int4 *__restrict__ spill_list, int *__restrict__ flags) {
  load_scatter_body<false>(elements, M, N, own, coords, f, weights, points, n_qp, wide, acc, slots,
                           spill_list, flags);
}
__global__ void __launch_bounds__(SC_THREADS)
load_scatter_midpoint_kernel(const int *__restrict__ elements, int64_t M, int N, Own own,
                             const double2 *__restrict__ coords, const double *__restrict__ f,
                             int wide, unsigned long long *__restrict__ acc,
                             Rec *__restrict__ slots, int4 *__restrict__ spill_list,
                             int *__restrict__ flags) {
  load_scatter_body<true>(elements, M, N, own, coords, f, nullptr, nullptr, 0, wide, acc, slots,
                          spill_list, flags);
}

// ascending (element, vertex) for the cubature rule (bincount over
// elements.flatten(), solvers.py:593-596), ascending (vertex, element) for the
// legacy rule (elements.flatten('F'), solvers.py:422-425)
template <bool VERTEX_MAJOR>
__global__ void __launch_bounds__(THREADS)
sum_slots_kernel(int n, const unsigned long long *__restrict__ acc,
                 const int4 *__restrict__ slots, double *__restrict__ b) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int deg = (int)(unsigned)acc[a];
  double sum = 0.;
  if (deg >= 1 && deg <= GROUP) {  // (more than 8: flagged by the scatter, b is discarded)
    const int4 *row = slots + (int64_t)a * GROUP;
    unsigned key[GROUP];
    double val[GROUP];
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      key[i] = 0xffffffffu;
      val[i] = 0.;
      if (i < deg) {
        const int4 s = __ldg(row + i);
        const unsigned c = (unsigned)s.x;
        key[i] = VERTEX_MAJOR ? ((c & 3u) << 29) | (c >> 2) : c;
        val[i] = __hiloint2double(s.w, s.z);
      }
    }
    int rank[GROUP];
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      rank[i] = 0;
#pragma unroll
      for (int j = 0; j < GROUP; ++j) rank[i] += key[j] < key[i];
    }
    for (int q = 0; q < deg; ++q) {
#pragma unroll
      for (int i = 0; i < GROUP; ++i)
        if (rank[i] == q) sum += val[i];
    }
  }
  b[a] = sum;
}

// nodes with more than 8 incidences: the thread of the FIRST spilled entry of a node
// adds all of the node's values (8 slots + its spilled entries) in key order
__global__ void __launch_bounds__(THREADS)
sum_spilled_kernel(const unsigned long long *__restrict__ acc, const int4 *__restrict__ slots,
                   const int4 *__restrict__ spill_list, const int *__restrict__ n_spilled,
                   int vertex_major, double *__restrict__ b) {
  const int n = min(*n_spilled, TSPILL_CAP);
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int a = spill_list[s].y;
  for (int j = 0; j < s; ++j)
    if (spill_list[j].y == a) return;
  const int deg = (int)(unsigned)acc[a];
  auto key_of = [&](int code) {
    const unsigned c = (unsigned)code;
    return vertex_major ? ((c & 3u) << 29) | (c >> 2) : c;
  };
  double sum = 0.;
  unsigned last = 0;
  for (int q = 0; q < deg; ++q) {
    unsigned best = 0xffffffffu;
    double val = 0.;
    for (int i = 0; i < GROUP; ++i) {
      const int4 r = slots[(int64_t)a * GROUP + i];
      const unsigned key = key_of(r.x);
      if ((q == 0 || key > last) && key < best) { best = key; val = __hiloint2double(r.w, r.z); }
    }
    for (int j = 0; j < n; ++j) {
      const int4 r = spill_list[j];
      if (r.y != a) continue;
      const unsigned key = key_of(r.x);
      if ((q == 0 || key > last) && key < best) { best = key; val = __hiloint2double(r.w, r.z); }
    }
    if (best == 0xffffffffu) break;  // (entries beyond the list's capacity: call refused)
    sum += val;
    last = best;
  }
  b[a] = sum;
}

// f as 32-byte loads: the rows of the [n_qp x M] array must stay 32-byte aligned
static void launch_load_scatter(unsigned grid, cudaStream_t s, const int *elements, int64_t M,
                                int N, Own own, const LoadQuadValues &v, unsigned long long *acc,
                                int4 *slots, int4 *spill_list, int *flags) {
  const int wide = ((uintptr_t)v.fq & 31) == 0 && (M & 3) == 0;
  load_scatter_cubature_kernel<<<grid, SC_THREADS, 0, s>>>(
      elements, M, N, own, v.coords, v.fq, v.weights, v.points, v.n_qp, wide, acc,
      reinterpret_cast<Rec *>(slots), spill_list, flags);
}
static void launch_load_scatter(unsigned grid, cudaStream_t s, const int *elements, int64_t M,
                                int N, Own own, const LoadMidpointValues &v,
                                unsigned long long *acc, int4 *slots, int4 *spill_list,
                                int *flags) {
  const int wide = ((uintptr_t)v.f_centroid & 31) == 0;
  load_scatter_midpoint_kernel<<<grid, SC_THREADS, 0, s>>>(
      elements, M, N, own, v.coords, v.f_centroid, wide, acc, reinterpret_cast<Rec *>(slots),
      spill_list, flags);
}

template <class V>
static int load_vector_direct(p1_ctx *ctx, const int32_t *elements, int64_t M, int64_t n_nodes,
                              V values, bool vertex_major, bool allow_pairs, double *b,
                              cudaStream_t s, const char *who) {
  const int N = (int)n_nodes;
  const Own own{0, N, 0, 1, 0, 0};
  unsigned long long *acc = nullptr;
  int4 *slots = nullptr;
  int rc;
  if ((rc = pool_alloc(ctx, &acc, (size_t)N + 1, s))) return rc;
  int4 *spill_list = nullptr;  // {code, row, value} of the records beyond 8 per node
  if ((rc = pool_alloc(ctx, &slots, (size_t)N * GROUP, s)) ||
      (rc = pool_alloc(ctx, &spill_list, (size_t)TSPILL_CAP, s))) {
    pool_free(ctx, acc, s);
    pool_free(ctx, slots, s);
    return rc;
  }
  cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s);
  cudaMemsetAsync(acc, 0, ((size_t)N + 1) * sizeof(unsigned long long), s);
  const unsigned grid = min(grid_for(M, THREADS), (unsigned)(ctx->sm_count * 32));
  if (M > 0) {
    prof_begin(ctx, "load_scatter", s);
    if (((uintptr_t)elements & 15) == 0)
      launch_load_scatter(grid, s, elements, M, N, own, values, acc, slots, spill_list,
                          ctx->d_flags);
    else
      launch_scatter_slots<2>(ctx, grid, s, elements, M, N, own, acc, slots, values, allow_pairs,
                              nullptr, spill_list);
    prof_end(ctx, s);
  }
  if (N > 0) {
    prof_begin(ctx, "load_sum", s);
    if (vertex_major)
      sum_slots_kernel<true><<<grid_for(N, THREADS), THREADS, 0, s>>>(N, acc, slots, b);
    else
      sum_slots_kernel<false><<<grid_for(N, THREADS), THREADS, 0, s>>>(N, acc, slots, b);
    // nodes with more than 8 elements (the count of spilled records is read on the device)
    sum_spilled_kernel<<<grid_for(TSPILL_CAP, THREADS), THREADS, 0, s>>>(
        acc, slots, spill_list, ctx->d_flags + 5, vertex_major ? 1 : 0, b);
    prof_end(ctx, s);
  }
  cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  pool_free(ctx, acc, s);
  pool_free(ctx, slots, s);
  pool_free(ctx, spill_list, s);
  if (e != cudaSuccess) {
    set_error("%s: %s", who, cudaGetErrorString(e));
    return P1_ECUDA;
  }
  if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
    set_error("%s: element index out of range [0, %lld)", who, (long long)n_nodes);
    return P1_EINDEX;
  }
  if ((ctx->h_flags[0] & FLAG_OVERFLOW) || ctx->h_flags[5] > TSPILL_CAP) {
    set_error("%s: more than %d incidences beyond 8 per node; use a plan "
              "(P1_PLAN_KEEP_ELEMENTS) and p1_load_vector", who, TSPILL_CAP);
    return P1_ERETRY;
  }
  return P1_OK;
}

int require_elements(const p1_plan *p, const char *who) {
  if (p->rec_elem || p->topo) return P1_OK;
  set_error("%s needs a plan created with P1_PLAN_KEEP_ELEMENTS or P1_PLAN_TOPOLOGY", who);
  return P1_EINVAL;
}

// ------------------------------------------------------ balanced row ranges
__global__ void __launch_bounds__(THREADS)
sample_hist_kernel(const int *__restrict__ elements, int64_t M, int N, int stride,
                   int buckets, int *__restrict__ hist) {
  // pseudo-random 1/stride sample: a strided one is biased, because the position
  // of an element among its NVB siblings correlates with the age of its nodes
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M;
       t += (int64_t)gridDim.x * blockDim.x) {
    if (stride > 1 && (mix32((unsigned)t) % (unsigned)stride) != 0u) continue;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int e = __ldg(elements + 3 * t + k);
      if ((unsigned)e < (unsigned)N)
        atomicAdd(hist + (int)(((int64_t)e * buckets) / N), 1);
    }
  }
}

}  // namespace p1

using namespace p1;

// Row ranges with (nearly) equal numbers of stored entries, from a strided
// sample of the elements: row i holds about 1 + #incident elements entries.
// Deterministic in (elements, world): every rank computes the same answer
// without communication.  boundaries_host receives world+1 row indices.
extern "C" int p1_partition_rows(p1_ctx *ctx, const int32_t *elements,
                                 int64_t n_elements, int64_t n_nodes, int world,
                                 int64_t *boundaries_host, void *stream) {
  P1_REQUIRE(ctx && boundaries_host && world >= 1, "bad argument");
  P1_REQUIRE(n_elements >= 0 && n_nodes >= 0 && n_nodes <= 0x7fffffffLL, "bad size");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  const int buckets = 8192;
  const int stride = n_elements > (1 << 22) ? 16 : 1;
  std::vector<int> h(buckets, 0);
  if (n_elements > 0 && n_nodes > 0) {
    int *hist = nullptr;
    P1_TRY(pool_alloc(ctx, &hist, (size_t)buckets, s));
    P1_CUDA(cudaMemsetAsync(hist, 0, buckets * sizeof(int), s));
    sample_hist_kernel<<<min(grid_for(n_elements, THREADS), (unsigned)(ctx->sm_count * 16)),
                         THREADS, 0, s>>>(elements, n_elements, (int)n_nodes, stride,
                                          buckets, hist);
    P1_CUDA(cudaMemcpyAsync(h.data(), hist, buckets * sizeof(int), cudaMemcpyDeviceToHost, s));
    P1_CUDA(cudaStreamSynchronize(s));
    pool_free(ctx, hist, s);
  }
  // weight of bucket b = rows in it + sampled incidences * stride
  std::vector<double> cum(buckets + 1, 0.);
  for (int b = 0; b < buckets; ++b) {
    const int64_t lo = (int64_t)b * n_nodes / buckets, hi = (int64_t)(b + 1) * n_nodes / buckets;
    cum[b + 1] = cum[b] + (double)(hi - lo) + (double)h[b] * stride;
  }
  boundaries_host[0] = 0;
  boundaries_host[world] = n_nodes;
  int b = 0;
  for (int r = 1; r < world; ++r) {
    const double target = cum[buckets] * r / world;
    while (b < buckets - 1 && cum[b + 1] < target) ++b;
    const int64_t lo = (int64_t)b * n_nodes / buckets, hi = (int64_t)(b + 1) * n_nodes / buckets;
    const double w = cum[b + 1] - cum[b];
    const double frac = w > 0 ? (target - cum[b]) / w : 0.;
    int64_t row = lo + (int64_t)(frac * (double)(hi - lo));
    if (row < boundaries_host[r - 1]) row = boundaries_host[r - 1];
    if (row > n_nodes) row = n_nodes;
    boundaries_host[r] = row;
  }
  return P1_OK;
}

constexpr int PLAN_NO_SLIM = 1 << 16;   // internal: second attempt after an overflow
constexpr int PLAN_NO_TILED = 1 << 17;  // internal: the tiled plan declined this mesh

void p1::overflow_collect(p1_plan *p, int *cursor, int *next_free, cudaStream_t s) {
  p1_ctx *ctx = p->ctx;
  prof_begin(ctx, "overflow_collect", s);
  if (p->n_slow > 0)
    overflow_heads_kernel<<<grid_for(p->n_slow, THREADS), THREADS, 0, s>>>(
        p->slow_rows, p->n_slow, p->acc, p->rec, p->rec_elem, p->ovf_off, p->ovf_rec,
        p->ovf_elem, cursor, next_free);
  if (p->n_big > 0)
    overflow_heads_kernel<<<grid_for(p->n_big, THREADS), THREADS, 0, s>>>(
        p->big_rows, p->n_big, p->acc, p->rec, p->rec_elem, p->ovf_off, p->ovf_rec,
        p->ovf_elem, cursor, next_free);
  if (p->n_spill > 0)
    overflow_place_kernel<<<grid_for(p->n_spill, THREADS), THREADS, 0, s>>>(
        p->n_spill, p->spill, cursor, p->ovf_rec, p->ovf_elem);
  prof_end(ctx, s);
}

static int plan_build(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                      int64_t n_nodes, int64_t row_begin, int64_t row_end, int chunk_log2,
                      int world, int rank, int flags,
                      int op, const double *coords, const double *local,
                      const p1_quad_args *qa, void *stream, p1_plan **out) {
  P1_REQUIRE(ctx && out, "null ctx/out");
  P1_REQUIRE(n_elements >= 0 && n_nodes >= 0, "negative size");
  P1_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n_nodes,
             "row range outside [0, n_nodes]");
  P1_REQUIRE(world >= 1 && rank >= 0 && rank < world && chunk_log2 >= 0 && chunk_log2 < 31,
             "bad interleave");
  if (n_nodes > (int64_t)NODE_MASK || n_elements >= (1LL << 29) ||
      3 * n_elements + n_nodes > 0x7fffffffLL) {
    set_error("p1_plan_create: mesh too large for int32 device indices "
              "(M=%lld N=%lld; limits: N < 2^29, M < 2^29, 3M+N < 2^31)",
              (long long)n_elements, (long long)n_nodes);
    return P1_ERANGE;
  }
  P1_REQUIRE(elements || n_elements == 0, "null elements");
  P1_REQUIRE(op == 0 || op == P1_OP_STIFFNESS || op == P1_OP_MASS || op == P1_OP_LOCAL ||
                 op == P1_OP_GENERAL || op == P1_OP_WEIGHTED, "unknown op");
  const bool quad = op == P1_OP_GENERAL || op == P1_OP_WEIGHTED;
  P1_REQUIRE(!(op == P1_OP_STIFFNESS || op == P1_OP_MASS || quad) || coords ||
                 n_elements == 0, "this operator needs coords");
  P1_REQUIRE(!quad || (qa && qa->n_qp > 0 && qa->weights_host), "bad p1_quad_args");
  P1_REQUIRE(op != P1_OP_GENERAL || n_elements == 0 ||
                 (qa->a11q && qa->a12q && qa->a21q && qa->a22q), "null coefficient array");
  P1_REQUIRE(op != P1_OP_WEIGHTED || (qa->points_host && (n_elements == 0 || qa->phiq)),
             "null phiq/points");
  P1_REQUIRE(op != P1_OP_LOCAL || local || n_elements == 0,
             "P1_OP_LOCAL needs the local matrices");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  // One GPU, all rows, one operator, nothing to keep: the tiled path on request (tiled.cu).
  if ((flags & P1_PLAN_TILED) && op != 0 && world == 1 && row_begin == 0 &&
      row_end == n_nodes &&
      !(flags & (P1_PLAN_EXACT | P1_PLAN_KEEP_ELEMENTS | P1_PLAN_TOPOLOGY | P1_PLAN_FORCE_SLIM |
                 P1_PLAN_FORCE_RECORDS | P1_PLAN_TINY_SPILL | PLAN_NO_SLIM | PLAN_NO_TILED)) &&
      ((uintptr_t)elements & 15) == 0 && out) {
    const int rc = plan_build_tiled(ctx, elements, n_elements, n_nodes, op, coords, local, qa, s,
                                    out);
    if (rc != P1_ERETRY) return rc;  // (P1_ERETRY: a fan of more than 40 elements)
  }
  int wshift = -1;
  for (int b = 0; b < 31; ++b)
    if ((1 << b) == world) wshift = b;
  const Own own{(int)row_begin, (int)row_end, chunk_log2, world, rank, wshift};
  const int n_own = (int)own.count();
  const int N = (int)n_nodes;
  const int64_t M = n_elements;
  const bool keep = (flags & P1_PLAN_KEEP_ELEMENTS) != 0;

  p1_plan *p = new p1_plan();
  memset(p, 0, sizeof(*p));
  p->ctx = ctx; p->stream = s; p->elements = elements; p->M = M; p->N = n_nodes;
  p->row_begin = row_begin; p->row_end = row_end;
  p->own = own; p->n_own = n_own;
  p->exact = (flags & P1_PLAN_EXACT) != 0;
  *out = nullptr;

  int rc = P1_OK;
  int *rowlen = nullptr, *cnt = nullptr;
  double *d_weights = nullptr, *d_points = nullptr;
  SpillDesc spill{nullptr, nullptr, nullptr, 0};  // records of rows with more than 8
  SpillDesc *d_spill = nullptr;
  auto free_spill = [&]() {
    pool_free(ctx, spill.rec, s); pool_free(ctx, spill.row, s); pool_free(ctx, spill.elem, s);
    pool_free(ctx, d_spill, s);
    spill = SpillDesc{nullptr, nullptr, nullptr, 0};
    d_spill = nullptr;
  };
  auto fail = [&](int code) {
    free_spill();
    pool_free(ctx, rowlen, s);
    pool_free(ctx, cnt, s);
    pool_free(ctx, d_weights, s);
    pool_free(ctx, d_points, s);
    p1_plan_destroy(p);
    return code;
  };
  auto cuda_fail = [&](const char *what) {
    set_error("p1_plan_create: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    return fail(P1_ECUDA);
  };

  // flags: [0] status, [1] max index, [2] n_big, [3] n_slow, [4] rows with > 8 records
  if (cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s) != cudaSuccess ||
      cudaMemsetAsync(ctx->d_flags + 1, 0xff, sizeof(int), s) != cudaSuccess)
    return cuda_fail("memset");
  if ((rc = pool_alloc(ctx, &p->acc, (size_t)n_own + 1, s))) return fail(rc);
  if (flags & P1_PLAN_TOPOLOGY) {
    // edge numbering / eta_R / load vectors: one 16-byte slot per incidence, no values,
    // no row pointer, one synchronisation
    P1_REQUIRE(op == 0, "a topology-only plan carries no operator");
    if ((rc = pool_alloc(ctx, &p->topo, (size_t)n_own * GROUP, s))) return fail(rc);
    // slot of every half-edge (all rows owned, slot numbers fit an int)
    if ((rc = pool_alloc(ctx, &p->tspill, (size_t)TSPILL_CAP, s))) return fail(rc);
    if ((flags & P1_PLAN_SLOT_INDEX) && own.world == 1 && own.re - own.rb == N &&
        (int64_t)n_own * GROUP + TSPILL_CAP <= 0x7fffffffLL &&
        (rc = pool_alloc(ctx, &p->slot_of, (size_t)(3 * M) + 1, s)))
      return fail(rc);
    cudaMemsetAsync(p->acc, 0, ((size_t)n_own + 1) * sizeof(unsigned long long), s);
    const unsigned tg = min(grid_for(M, THREADS), (unsigned)(ctx->sm_count * 32));
    prof_begin(ctx, "scatter", s);
    launch_scatter_slots<1>(ctx, tg, s, elements, M, N, own, p->acc, p->topo, NoValues{}, true,
                            p->slot_of, p->tspill);
    prof_end(ctx, s);
    cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess)
      return cuda_fail("scatter");
    if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
      set_error("p1_plan_create: element index out of range [0, %lld)", (long long)n_nodes);
      return fail(P1_EINDEX);
    }
    if ((ctx->h_flags[0] & FLAG_OVERFLOW) || ctx->h_flags[5] > TSPILL_CAP) {
      set_error("p1_plan_create: P1_PLAN_TOPOLOGY holds 8 elements per node plus %d in "
                "total beyond that; use P1_PLAN_KEEP_ELEMENTS for this mesh", TSPILL_CAP);
      return fail(P1_ERETRY);
    }
    p->n_tspill = ctx->h_flags[5];
    p->max_index = ctx->h_flags[1];
    p->nnz = 0;
    *out = p;
    return P1_OK;
  }
  // Multi-GPU stiffness / mass: 8-byte slots, values recomputed by the fill from
  // gathered coordinates.  On one GPU that loses (every element's local row is then
  // computed three times: fill 1.9 -> 3.0 ms, step 4.55 -> 4.98 ms), but a rank that
  // owns 1/G of the rows meets most of its elements only once anyway, and its
  // scatter no longer gathers three coordinates for every element that touches one
  // of its rows (1 - (1 - 1/G)^3 of them).  (P1_PLAN_FORCE_SLIM: one-GPU measurements.)
  const bool slim = (own.world > 1 || n_own < N || (flags & P1_PLAN_FORCE_SLIM)) && !p->exact &&
                    !keep && !(flags & PLAN_NO_SLIM) &&
                    (op == P1_OP_STIFFNESS || op == P1_OP_MASS);
  if (slim) {
    if ((rc = pool_alloc(ctx, &p->slim, (size_t)n_own * GROUP, s))) return fail(rc);
    p->coords = (const double2 *)coords;
  } else if ((rc = pool_alloc(ctx, &p->rec, (size_t)n_own * GROUP, s))) {
    return fail(rc);
  }
  if (keep && (rc = pool_alloc(ctx, &p->rec_elem, (size_t)n_own * GROUP, s))) return fail(rc);
  if ((rc = pool_alloc(ctx, &p->indptr, (size_t)n_own + 1, s))) return fail(rc);
  if ((rc = pool_alloc(ctx, &rowlen, (size_t)n_own + 1, s))) return fail(rc);
  if ((rc = pool_alloc(ctx, &p->slow_rows, (size_t)n_own + 1, s))) return fail(rc);
  {
    const int64_t cap = (3 * M) / (DEG_BIG + 1);
    if ((rc = pool_alloc(ctx, &p->big_rows, (size_t)(cap < n_own ? cap : n_own) + 1, s)))
      return fail(rc);
  }
  cudaMemsetAsync(p->acc, 0, ((size_t)n_own + 1) * sizeof(unsigned long long), s);

  const unsigned eg = min(grid_for(M, THREADS), (unsigned)(ctx->sm_count * 32));
  if (quad) {  // the rule travels to the device once
    if ((rc = pool_alloc(ctx, &d_weights, (size_t)qa->n_qp, s))) return fail(rc);
    if ((rc = pool_alloc(ctx, &d_points, (size_t)2 * qa->n_qp, s))) return fail(rc);
    cudaMemcpyAsync(d_weights, qa->weights_host, qa->n_qp * sizeof(double),
                    cudaMemcpyHostToDevice, s);
    if (qa->points_host)
      cudaMemcpyAsync(d_points, qa->points_host, 2 * qa->n_qp * sizeof(double),
                      cudaMemcpyHostToDevice, s);
  }
  auto free_rule = [&]() {
    pool_free(ctx, d_weights, s);
    pool_free(ctx, d_points, s);
    d_weights = d_points = nullptr;
  };
  {  // spill list (see SpillDesc)
    int64_t cap = (3 * M) / 64;
    cap = cap < 4096 ? 4096 : (cap > (1 << 22) ? (1 << 22) : cap);
    if (flags & P1_PLAN_TINY_SPILL) cap = 1;  // tests: force the second pass over the elements
    spill.cap = (int)cap;
    if ((rc = pool_alloc(ctx, &spill.rec, (size_t)cap, s)) ||
        (rc = pool_alloc(ctx, &spill.row, (size_t)cap, s)) ||
        (keep && (rc = pool_alloc(ctx, &spill.elem, (size_t)cap, s))) ||
        (rc = pool_alloc(ctx, &d_spill, (size_t)1, s)))
      return fail(rc);
    // (pageable host source: the copy is staged before the call returns)
    cudaMemcpyAsync(d_spill, &spill, sizeof(spill), cudaMemcpyHostToDevice, s);
  }
  prof_begin(ctx, "scatter", s);
  if (slim)
    launch_scatter_slots<3>(ctx, eg, s, elements, M, N, own, p->acc, p->slim, NoValues{}, true);
  else
    launch_scatter<false>(ctx, eg, s, op, elements, M, N, own, p->acc, p->rec,
                          p->rec_elem, reinterpret_cast<int *>(d_spill), coords, local, qa,
                          d_weights, d_points);
  prof_end(ctx, s);
  p->val_op = op;
  const Rows rows0 = rows_of(p);
  if (n_own > 0) {
    if (p->exact)
      P1_TIMED(ctx, "rowlen", s,
               rowlen_kernel<<<grid_for((int64_t)n_own * GROUP, THREADS), THREADS, 0, s>>>(
                   n_own, own, rows0, rowlen, p->slow_rows, p->big_rows, ctx->d_flags + 2));
    else
      P1_TIMED(ctx, "split", s,
               split_kernel<<<grid_for(n_own, THREADS), THREADS, 0, s>>>(
                   n_own, p->acc, rowlen, p->slow_rows, p->big_rows, ctx->d_flags + 2));
  }
  // Common case (no row with more than 8 records): ONE synchronisation per plan.
  // The boundary rows are counted with the list length read on the device.
  const unsigned sg = (unsigned)(ctx->sm_count * 4);
  int h_nnz = 0;
  auto finish = [&]() -> int {   // slow rows -> scan -> read flags and nnz
    rowlen_slow_kernel<<<sg, THREADS, 0, s>>>(p->slow_rows, ctx->d_flags + 3, own, rows_of(p),
                                              rowlen);
    int r = exclusive_scan_i32(ctx, rowlen, p->indptr, n_own, true, s);
    if (r) return r;
    cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&h_nnz, p->indptr + n_own, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("p1_plan_create: %s", cudaGetErrorString(e));
      return P1_ECUDA;
    }
    return P1_OK;
  };
  if ((rc = finish())) return fail(rc);
  if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
    set_error("p1_plan_create: element index out of range [0, %lld)", (long long)n_nodes);
    return fail(P1_EINDEX);
  }
  if (h_nnz < 0) {
    set_error("p1_plan_create: more than 2^31 - 1 stored entries in these rows; split the rows "
              "(row_begin / row_end) so that every piece fits an int32 row pointer");
    return fail(P1_ERANGE);
  }
  p->max_index = ctx->h_flags[1];
  p->n_big = ctx->h_flags[2];
  p->n_slow = ctx->h_flags[3];
  if (ctx->h_flags[4] > 0 && slim) {
    // a node with more than 8 elements: this mesh takes the record path
    fail(P1_OK);
    return plan_build(ctx, elements, n_elements, n_nodes, row_begin, row_end, chunk_log2, world,
                      rank, flags | PLAN_NO_SLIM, op, coords, local, qa, stream, out);
  }
  if (ctx->h_flags[4] > 0) {
    // some rows have more than 8 records: give them compact storage and redo
    // their incidences (a second pass over the elements; never on an NVB mesh)
    if ((rc = pool_alloc(ctx, &cnt, (size_t)n_own + 1, s))) return fail(rc);
    if ((rc = pool_alloc(ctx, &p->ovf_off, (size_t)n_own + 1, s))) return fail(rc);
    const int n_spill = ctx->h_flags[5];
    if (n_spill <= spill.cap) {
      // the common case: a few rows, their records are already here.  Every such row
      // has exactly 8 records in its slots and the rest in the spill list, so the size
      // of the compact storage is known without counting; ovf_off / the cursors are
      // valid for those rows only.
      p->n_ovf = (int64_t)GROUP * ctx->h_flags[4] + n_spill;
      if ((rc = pool_alloc(ctx, &p->ovf_rec, (size_t)p->n_ovf, s))) return fail(rc);
      if (keep && (rc = pool_alloc(ctx, &p->ovf_elem, (size_t)p->n_ovf, s))) return fail(rc);
      p->spill = spill;
      p->n_spill = n_spill;
      overflow_collect(p, cnt, ctx->d_flags + 6, s);
      p->spill = SpillDesc{nullptr, nullptr, nullptr, 0};  // (freed with the builder's scratch)
    } else {
      // the spill list was too small: count, and scan the elements again for those rows
      overflow_count_kernel<<<grid_for(n_own, THREADS), THREADS, 0, s>>>(n_own, p->acc, cnt);
      if ((rc = exclusive_scan_i32(ctx, cnt, p->ovf_off, n_own, true, s))) return fail(rc);
      int h = 0;
      cudaMemcpyAsync(&h, p->ovf_off + n_own, sizeof(int), cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) return cuda_fail("overflow scan");
      p->n_ovf = h;
      if ((rc = pool_alloc(ctx, &p->ovf_rec, (size_t)p->n_ovf, s))) return fail(rc);
      if (keep && (rc = pool_alloc(ctx, &p->ovf_elem, (size_t)p->n_ovf, s))) return fail(rc);
      cudaMemcpyAsync(cnt, p->ovf_off, (size_t)n_own * sizeof(int), cudaMemcpyDeviceToDevice, s);
      prof_begin(ctx, "scatter_overflow", s);
      launch_scatter<true>(ctx, eg, s, op, elements, M, N, own, p->acc, p->ovf_rec,
                           p->ovf_elem, cnt, coords, local, qa, d_weights, d_points);
      prof_end(ctx, s);
    }
    pool_free(ctx, cnt, s);
    cnt = nullptr;
    const Rows rows = rows_of(p);
    if (p->n_big > 0) {
      Rec *tmp = nullptr;
      int *tmp_elem = nullptr;
      if ((rc = pool_alloc(ctx, &tmp, (size_t)p->n_ovf, s)) ||
          (rc = pool_alloc(ctx, &tmp_elem, (size_t)p->n_ovf, s)) ||
          (rc = pool_alloc(ctx, &p->big_first, (size_t)(2 * p->n_ovf + n_own + 1), s))) {
        pool_free(ctx, tmp, s);
        pool_free(ctx, tmp_elem, s);
        return fail(rc);
      }
      sort_big_kernel<<<p->n_big, THREADS, 0, s>>>(p->big_rows, rows, tmp, tmp_elem);
      rowlen_big_kernel<<<p->n_big, THREADS, 0, s>>>(p->big_rows, own, rows, p->big_first,
                                                     rowlen);
      pool_free(ctx, tmp, s);
      pool_free(ctx, tmp_elem, s);
    }
    if ((rc = finish())) return fail(rc);  // slow rows again (now with their records), rescan
  }
  pool_free(ctx, rowlen, s);
  rowlen = nullptr;
  free_rule();
  free_spill();
  p->nnz = h_nnz;
  *out = p;
  return P1_OK;
}

extern "C" int p1_plan_create(p1_ctx *ctx, const int32_t *elements,
                              int64_t n_elements, int64_t n_nodes,
                              int64_t row_begin, int64_t row_end, int flags,
                              void *stream, p1_plan **out) {
  return plan_build(ctx, elements, n_elements, n_nodes, row_begin, row_end, 0, 1, 0, flags,
                    0, nullptr, nullptr, nullptr, stream, out);
}

extern "C" int p1_plan_create_for(p1_ctx *ctx, const int32_t *elements,
                                  int64_t n_elements, int64_t n_nodes,
                                  int64_t row_begin, int64_t row_end, int flags, int op,
                                  const double *coords, const double *local,
                                  void *stream, p1_plan **out) {
  P1_REQUIRE(op != 0, "p1_plan_create_for needs an operator");
  P1_REQUIRE(op != P1_OP_GENERAL && op != P1_OP_WEIGHTED, "use p1_plan_create_quad");
  return plan_build(ctx, elements, n_elements, n_nodes, row_begin, row_end, 0, 1, 0, flags,
                    op, coords, local, nullptr, stream, out);
}

extern "C" int p1_plan_create_quad(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                                   int64_t n_nodes, int64_t row_begin, int64_t row_end,
                                   int flags, int op, const double *coords,
                                   const p1_quad_args *args, void *stream, p1_plan **out) {
  P1_REQUIRE(op == P1_OP_GENERAL || op == P1_OP_WEIGHTED,
             "p1_plan_create_quad: op must be P1_OP_GENERAL or P1_OP_WEIGHTED");
  return plan_build(ctx, elements, n_elements, n_nodes, row_begin, row_end, 0, 1, 0, flags,
                    op, coords, nullptr, args, stream, out);
}

extern "C" int p1_plan_create_interleaved(p1_ctx *ctx, const int32_t *elements,
                                          int64_t n_elements, int64_t n_nodes,
                                          int chunk_log2, int world, int rank, int flags,
                                          int op, const double *coords, const double *local,
                                          void *stream, p1_plan **out) {
  P1_REQUIRE(op != P1_OP_GENERAL && op != P1_OP_WEIGHTED,
             "interleaved plans take P1_OP_LOCAL for these operators");
  return plan_build(ctx, elements, n_elements, n_nodes, 0, n_nodes, chunk_log2, world, rank,
                    flags, op, coords, local, nullptr, stream, out);
}

extern "C" int p1_load_vector_direct(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                                     int64_t n_nodes, const double *coords, const double *fq,
                                     const double *weights_host, const double *points_host,
                                     int n_qp, double *b, void *stream) {
  P1_REQUIRE(ctx && b, "null ctx/b");
  P1_REQUIRE(n_elements >= 0 && n_nodes >= 0 && n_nodes <= (int64_t)NODE_MASK &&
                 n_elements < (1LL << 29), "bad size");
  P1_REQUIRE(n_elements == 0 || (elements && coords && fq), "null input");
  P1_REQUIRE(n_qp > 0 && weights_host && points_host, "bad rule");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  double *rule = nullptr;  // points (16-byte aligned pairs), then weights
  P1_TRY(pool_alloc(ctx, &rule, (size_t)3 * n_qp, s));
  cudaMemcpyAsync(rule, points_host, 2 * n_qp * sizeof(double), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(rule + 2 * n_qp, weights_host, n_qp * sizeof(double), cudaMemcpyHostToDevice,
                  s);
  const LoadQuadValues values{(const double2 *)coords, fq, rule + 2 * n_qp,
                              (const double2 *)rule, n_qp, n_elements};
  const int rc = load_vector_direct(ctx, elements, n_elements, n_nodes, values, false, false, b, s,
                                    "p1_load_vector_direct");
  pool_free(ctx, rule, s);
  return rc;
}

extern "C" int p1_load_vector_midpoint_direct(p1_ctx *ctx, const int32_t *elements,
                                              int64_t n_elements, int64_t n_nodes,
                                              const double *coords, const double *f_centroid,
                                              double *b, void *stream) {
  P1_REQUIRE(ctx && b, "null ctx/b");
  P1_REQUIRE(n_elements >= 0 && n_nodes >= 0 && n_nodes <= (int64_t)NODE_MASK &&
                 n_elements < (1LL << 29), "bad size");
  P1_REQUIRE(n_elements == 0 || (elements && coords && f_centroid), "null input");
  DeviceGuard guard(ctx);
  const LoadMidpointValues values{(const double2 *)coords, f_centroid};
  return load_vector_direct(ctx, elements, n_elements, n_nodes, values, true, true, b,
                            (cudaStream_t)stream, "p1_load_vector_midpoint_direct");
}

// ------------------------------------------------------- cold assembly, no host sync
// Everything the host used to wait for stays on the device: the status words are
// the caller's (flags, max index, list lengths, nnz), the outputs have a capacity
// instead of an exact size, the boundary-row list is walked with its length read on
// the device.  Meshes that need more than the common case (a node with more than 8
// elements, nnz above the capacity) are reported in the status word
// (P1_STATUS_SLOWPATH) and nothing is written; the caller then takes the plan route.
namespace p1 {
// Row lengths and row pointer without a row-length array or a list of irregular rows
// (replaces split + rowlen_slow + reduce / scan of sums / downsweep + finalize by reduce
// / scan of sums / final pass): closed fan: 1 + #records; open fan (a mesh-boundary
// node): the distinct neighbours are counted from its records on the spot; more than 8
// records: the plan route (FLAG_SLOWPATH).  Both passes recompute the lengths from the
// nodes' counters (8 bytes per row) instead of storing them.

// length of row a (see above) from its counter alone, or -1: an open fan, which
// open_row_length counts from the row's records.  big: more than 8 records
__device__ __forceinline__ int cold_row_length(unsigned long long v, bool &big) {
  const int deg = (int)(unsigned)v;
  if (deg > GROUP) { big = true; return 0; }
  if (deg == 0) return 0;
  if ((unsigned)(v >> 32) == 0u) return 1 + deg;  // closed fan (verified when the row is filled)
  return -1;
}

// open fan (1 <= deg <= 8): distinct values of {self} U {next_i} U {prev_i}.  All records
// are requested before the first is looked at: a thread meets these rows one after the
// other, and on a small mesh (one row in 128 on the boundary at 1M triangles) the slowest
// warp's chain of record misses was a third of the kernel (20 us before, 13 us after).
// (by value: a reference to the kernel's Rows would make every thread copy it to its stack)
__device__ __noinline__ int open_row_length(const Rec *rec, const int2 *slim, const int4 *half,
                                            int lay, int a, int self, int deg) {
  int ex[GROUP], ey[GROUP];
#pragma unroll
  for (int p = 0; p < GROUP; ++p) {
    int2 e = make_int2(self, self);
    if (p < deg) {
      const int64_t at = lay ? (int64_t)p * lay + a : (int64_t)a * GROUP + p;
      if (slim) {
        e = slim[at];
      } else if (half) {
        const int4 h = half[at];
        e = make_int2(h.x, h.y);
      } else {
        e = make_int2(rec[at].nx, rec[at].pvk);
      }
    }
    ex[p] = e.x;
    ey[p] = e.y & NODE_MASK;
  }
  int n = 1;
#pragma unroll
  for (int p = 0; p < GROUP; ++p) {
    const int x = ex[p], y = ey[p];
    bool fx = x != self, fy = y != self && y != x;
#pragma unroll
    for (int q = 0; q < GROUP; ++q) {
      fx = fx && !(q < p && ex[q] == x);
      fy = fy && ex[q] != y && !(q < p && ey[q] == y);
    }
    n += fx + fy;
  }
  return n;
}

// (A single-pass scan with decoupled look-back was measured at 33.5M rows: 0.34 ms with
// 1184 tiles of 1024 rows in flight and 0.34 ms with 296 persistent tiles of 4096 rows --
// a tile is read faster than its look-back walks over the tiles in flight -- against 0.29 ms
// for the four kernels replaced.  Reduce / scan of sums / final pass it is.)
constexpr int RP_THREADS = 256, RP_ITEMS = 8, RP_TILE = RP_THREADS * RP_ITEMS;

// What a cold assembly starts from, in one launch (three memset nodes cost a small mesh's
// graph 10 us): zeroed counters / tile sums / spill list, and the caller's status words
// (all zero but [1], the largest index met: -1).
__global__ void __launch_bounds__(256)
cold_init_kernel(ulonglong2 *__restrict__ acc2, size_t pairs, int *__restrict__ status) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride)
    acc2[i] = make_ulonglong2(0ull, 0ull);
  if (blockIdx.x == 0 && threadIdx.x < 8) status[threadIdx.x] = threadIdx.x == 1 ? -1 : 0;
}

// MODE 0: the tiles' sums; MODE 1: the row pointer, from the scanned sums.  MODE 2: both in
// ONE cooperative launch, for a grid that is resident at once (a mesh of up to ~4M
// triangles) -- the lengths stay in registers across the grid barrier and every tile adds
// up the sums of the tiles before it: at 1M triangles 30 us of 107 were the three launches'
// latencies.
constexpr int RP_REDUCE = 0, RP_FINAL = 1, RP_FUSED = 2;

template <int MODE>
__global__ void __launch_bounds__(RP_THREADS)
rowptr_kernel(int n_own, Own own, Rows rows, int *__restrict__ indptr,
              int *__restrict__ tile_sums /* n_tiles + 1; scanned when MODE 1 */,
              int *__restrict__ status, int64_t cap) {
  constexpr bool FINAL = MODE != RP_REDUCE;
  __shared__ int warp_sums[RP_THREADS / 32];
  __shared__ int base_s[RP_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int first = blockIdx.x * RP_TILE + threadIdx.x * RP_ITEMS;
  bool big = false;
  int len[RP_ITEMS], sum = 0;
  if (first + RP_ITEMS <= n_own) {  // (the arena's blocks are 512-byte aligned)
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(rows.acc + first);
#pragma unroll
    for (int i = 0; i < RP_ITEMS; i += 2) {
      const ulonglong2 v = src[i >> 1];
      len[i] = cold_row_length(v.x, big);
      len[i + 1] = cold_row_length(v.y, big);
    }
  } else {
#pragma unroll
    for (int i = 0; i < RP_ITEMS; ++i)
      len[i] = first + i < n_own ? cold_row_length(rows.acc[first + i], big) : 0;
  }
  unsigned open = 0;
#pragma unroll
  for (int i = 0; i < RP_ITEMS; ++i) open |= len[i] < 0 ? 1u << i : 0u;
  while (open) {  // (as many rounds as the warp's unluckiest thread has open rows)
    const int j = __ffs(open) - 1;
    open &= open - 1;
    const int n = open_row_length(rows.rec, rows.slim, rows.half, own.lay, first + j,
                                  own.global(first + j), rows.deg(first + j));
#pragma unroll
    for (int i = 0; i < RP_ITEMS; ++i) len[i] = i == j ? n : len[i];
  }
#pragma unroll
  for (int i = 0; i < RP_ITEMS; ++i) sum += len[i];
  int inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < RP_THREADS / 32; ++w) {
    before += w < warp ? warp_sums[w] : 0;
    total += warp_sums[w];
  }
  if (MODE != RP_FINAL) {
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
    if (big) atomicOr(status, FLAG_SLOWPATH);
    if (!FINAL) return;
  }
  int base;
  if (MODE == RP_FUSED) {
    __threadfence();
    cooperative_groups::this_grid().sync();
    int part = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += RP_THREADS) part += __ldcg(tile_sums + t);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) base_s[warp] = part;
    __syncthreads();
    base = 0;
#pragma unroll
    for (int w = 0; w < RP_THREADS / 32; ++w) base += base_s[w];
  } else {
    base = tile_sums[blockIdx.x];
  }
  int run = base + before + inc - sum;
  if (first + RP_ITEMS <= n_own && ((uintptr_t)indptr & 15) == 0) {
    int out[RP_ITEMS];
#pragma unroll
    for (int i = 0; i < RP_ITEMS; ++i) { out[i] = run; run += len[i]; }
    int4 *dst = reinterpret_cast<int4 *>(indptr + first);
#pragma unroll
    for (int i = 0; i < RP_ITEMS; i += 4)
      dst[i >> 2] = make_int4(out[i], out[i + 1], out[i + 2], out[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < RP_ITEMS; ++i) {
      if (first + i < n_own) indptr[first + i] = run;
      run += len[i];
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    // (-1: beyond int32, scan_sums_kernel; a resident grid's rows cannot get there)
    const int nnz = MODE == RP_FUSED ? base + total : tile_sums[gridDim.x];
    indptr[n_own] = nnz;
    status[6] = nnz;
    if (nnz < 0 || (int64_t)nnz > cap) atomicOr(status, FLAG_SLOWPATH);
  }
}
}  // namespace p1

// Small meshes are launch-bound (1M triangles: 8 launches for 65 us of kernels), and an
// adaptive or time-stepping loop assembles into the same buffers again and again: the
// launch sequence of a cold assembly of at most COLD_GRAPH_MAX triangles is captured once
// per set of arguments (on a private stream: the caller's may be the legacy default
// stream) and replayed as ONE CUDA graph launch afterwards.  A cached graph keeps its
// scratch blocks; the least recently used of COLD_GRAPH_SLOTS entries makes room.
constexpr int64_t COLD_GRAPH_MAX = 1 << 22;
constexpr int COLD_GRAPH_SLOTS = 8;
struct ColdKey {
  const void *elements, *coords, *local, *indptr, *indices, *data, *status;
  int64_t M, N, capacity;
  int chunk_log2, world, rank, op;
  cudaStream_t stream;
  bool operator==(const ColdKey &o) const { return memcmp(this, &o, sizeof(ColdKey)) == 0; }
};
struct ColdGraph {
  ColdKey key;
  cudaGraphExec_t exec;
  void *acc, *slots;
  unsigned long long tick;
};
struct p1_cold_cache {
  std::vector<ColdGraph> entries;
  // arguments met once: a graph is only built for a call that repeats, so a caller who
  // brings fresh buffers every time (an adaptive loop) never pays for a capture
  ColdKey seen[16];
  int n_seen = 0, next_seen = 0;
  cudaStream_t capture = nullptr;
  unsigned long long clock = 0;
};

void p1::cold_cache_clear(p1_ctx *ctx) {
  if (!ctx || !ctx->cold) return;
  cudaDeviceSynchronize();
  for (auto &g : ctx->cold->entries) {
    cudaGraphExecDestroy(g.exec);
    pool_free(ctx, g.acc, g.key.stream);
    pool_free(ctx, g.slots, g.key.stream);
  }
  if (ctx->cold->capture) cudaStreamDestroy(ctx->cold->capture);
  delete ctx->cold;
  ctx->cold = nullptr;
}

extern "C" int p1_assemble_cold(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                                int64_t n_nodes, int chunk_log2, int world, int rank, int op,
                                const double *coords, const double *local,
                                const p1_quad_args *qa, int64_t capacity, int32_t *indptr,
                                int32_t *indices, double *data, int32_t *status, void *stream) {
  P1_REQUIRE(ctx && indptr && status, "null ctx/indptr/status");
  P1_REQUIRE(n_elements >= 0 && n_nodes >= 0 && capacity >= 0, "negative size");
  P1_REQUIRE(data || capacity == 0, "null data");
  P1_REQUIRE(world >= 1 && rank >= 0 && rank < world && chunk_log2 >= 0 && chunk_log2 < 31,
             "bad interleave");
  if (n_nodes > (int64_t)NODE_MASK || n_elements >= (1LL << 29) ||
      3 * n_elements + n_nodes > 0x7fffffffLL) {
    set_error("p1_assemble_cold: mesh too large for int32 device indices");
    return P1_ERANGE;
  }
  P1_REQUIRE(elements || n_elements == 0, "null elements");
  const bool quad = op == P1_OP_GENERAL || op == P1_OP_WEIGHTED;
  P1_REQUIRE(op == P1_OP_STIFFNESS || op == P1_OP_MASS || op == P1_OP_LOCAL || quad, "unknown op");
  P1_REQUIRE(op == P1_OP_LOCAL || coords || n_elements == 0, "this operator needs coords");
  P1_REQUIRE(op != P1_OP_LOCAL || local || n_elements == 0, "P1_OP_LOCAL needs the local matrices");
  P1_REQUIRE(!quad || (qa && qa->n_qp > 0 && qa->weights_host), "bad p1_quad_args");
  P1_REQUIRE(op != P1_OP_GENERAL || n_elements == 0 ||
                 (qa->a11q && qa->a12q && qa->a21q && qa->a22q), "null coefficient array");
  P1_REQUIRE(op != P1_OP_WEIGHTED || (qa->points_host && (n_elements == 0 || qa->phiq)),
             "null phiq/points");
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  int wshift = -1;
  for (int b = 0; b < 31; ++b)
    if ((1 << b) == world) wshift = b;
  Own own{0, (int)n_nodes, chunk_log2, world, rank, wshift, 0};
  const int n_own = (int)own.count();
  const int N = (int)n_nodes;
  const int64_t M = n_elements;
  const bool slim = world > 1 && (op == P1_OP_STIFFNESS || op == P1_OP_MASS);
  if (!slim) own.lay = n_own;  // records slot-major (plan.cuh: Own::lay)
  // one GPU, mass matrix: the local row of an incidence is (2v, v, v), so a record is 16
  // bytes {next, prev|k, v}: half the record bytes through HBM in both directions
  const bool half = !slim && op == P1_OP_MASS;

  // one zeroed block: the nodes' counters, the scan's tile sums, the (empty) spill list
  const int n_tiles = (n_own + RP_TILE - 1) / RP_TILE;
  const size_t words = ((size_t)n_own + 1 + (size_t)(n_tiles + 4) / 2 + 8 + 1) & ~(size_t)1;
  unsigned long long *acc = nullptr;
  Rec *rec = nullptr;
  int2 *slots = nullptr;
  int4 *halves = nullptr;
  double *d_weights = nullptr, *d_points = nullptr;
  auto release = [&]() {
    for (void *b : {(void *)acc, (void *)rec, (void *)slots, (void *)halves, (void *)d_weights,
                    (void *)d_points})
      pool_free(ctx, b, s);
  };
  auto alloc_rows = [&]() {  // the rows' 8 slots, in the layout of this call
    return slim   ? pool_alloc(ctx, &slots, (size_t)n_own * GROUP, s)
           : half ? pool_alloc(ctx, &halves, (size_t)n_own * GROUP, s)
                  : pool_alloc(ctx, &rec, (size_t)n_own * GROUP, s);
  };
  // the whole launch sequence, on stream q
  auto enqueue = [&](cudaStream_t q) {
    int *tile_sums = reinterpret_cast<int *>(acc + n_own + 1);
    // (capacity 0: a record that finds its row full is dropped; the row is then reported)
    SpillDesc *d_spill =
        reinterpret_cast<SpillDesc *>(acc + n_own + 1 + (size_t)(n_tiles + 4) / 2);
#ifdef P1_AB_MEMSET
    cudaMemsetAsync(status, 0, 8 * sizeof(int), q);
    cudaMemsetAsync(status + 1, 0xff, sizeof(int), q);
    cudaMemsetAsync(acc, 0, words * sizeof(unsigned long long), q);
#else
    cold_init_kernel<<<min(grid_for((int64_t)(words / 2), 256), (unsigned)(ctx->sm_count * 8)), 256,
                       0, q>>>(reinterpret_cast<ulonglong2 *>(acc), words / 2, status);
#endif
    if (quad) {
      cudaMemcpyAsync(d_weights, qa->weights_host, qa->n_qp * sizeof(double),
                      cudaMemcpyHostToDevice, q);
      if (qa->points_host)
        cudaMemcpyAsync(d_points, qa->points_host, 2 * qa->n_qp * sizeof(double),
                        cudaMemcpyHostToDevice, q);
    }
    const unsigned eg = min(grid_for(M, THREADS), (unsigned)(ctx->sm_count * 32));
    prof_begin(ctx, "scatter", q);
    if (slim)
      launch_scatter_slots<3>(ctx, eg, q, elements, M, N, own, acc, slots, NoValues{}, true,
                              nullptr, nullptr, status);
    else if (half)
      launch_scatter_slots<4>(ctx, eg, q, elements, M, N, own, acc, halves,
                              MassValues{(const double2 *)coords}, true, nullptr, nullptr, status);
    else
      launch_scatter<false>(ctx, eg, q, op, elements, M, N, own, acc, rec, nullptr,
                            reinterpret_cast<int *>(d_spill), coords, local, qa, d_weights,
                            d_points, status);
    prof_end(ctx, q);
    const Rows rows{acc, rec, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, slots, halves};
    if (n_own > 0) {
      if (ctx->rowptr_resident == 0) {  // tiles that are resident at once, or -1
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rowptr_kernel<RP_FUSED>,
                                                          RP_THREADS, 0) != cudaSuccess)
          cudaGetLastError();
        ctx->rowptr_resident = per_sm > 0 ? per_sm * ctx->sm_count : -1;
      }
      const bool fused = n_tiles <= ctx->rowptr_resident;
      P1_TIMED(ctx, fused ? "rowptr_fused" : "rowptr", q, {
        if (fused) {
          int n_arg = n_own;
          Own own_arg = own;
          Rows rows_arg = rows;
          int64_t cap_arg = capacity;
          void *args[] = {&n_arg, &own_arg, &rows_arg, &indptr, &tile_sums, &status, &cap_arg};
          cudaLaunchCooperativeKernel((const void *)rowptr_kernel<RP_FUSED>, dim3(n_tiles),
                                      dim3(RP_THREADS), args, 0, q);
        } else {
          rowptr_kernel<RP_REDUCE><<<n_tiles, RP_THREADS, 0, q>>>(n_own, own, rows, indptr,
                                                                 tile_sums, status, capacity);
          scan_sums(tile_sums, n_tiles, q);
          rowptr_kernel<RP_FINAL><<<n_tiles, RP_THREADS, 0, q>>>(n_own, own, rows, indptr,
                                                                tile_sums, status, capacity);
        }
      });
      launch_fill_cold(ctx, own, n_own, rows, slim ? op : 0, (const double2 *)coords, indptr,
                       indices, data, status, q);
    } else {
      cudaMemsetAsync(indptr, 0, sizeof(int), q);
    }
  };
  int rc = P1_OK;

  // ---- small mesh, same arguments as before: one graph launch
  const bool graphable = !quad && M > 0 && M <= COLD_GRAPH_MAX && n_own > 0 && !prof_enabled(ctx);
  if (graphable) {
    if (!ctx->cold) ctx->cold = new p1_cold_cache();
    p1_cold_cache *cc = ctx->cold;
    ColdKey key;
    memset(&key, 0, sizeof(key));
    key.elements = elements; key.coords = coords; key.local = local; key.indptr = indptr;
    key.indices = indices; key.data = data; key.status = status;
    key.M = M; key.N = n_nodes; key.capacity = capacity;
    key.chunk_log2 = chunk_log2; key.world = world; key.rank = rank; key.op = op;
    key.stream = s;
    for (auto &g : cc->entries)
      if (g.key == key) {
        g.tick = ++cc->clock;
        P1_CUDA(cudaGraphLaunch(g.exec, s));
        return P1_OK;
      }
    bool repeated = false;
    for (int i = 0; i < cc->n_seen; ++i) repeated |= cc->seen[i] == key;
    if (!repeated) {
      cc->seen[cc->next_seen] = key;
      cc->next_seen = (cc->next_seen + 1) % 16;
      cc->n_seen = cc->n_seen < 16 ? cc->n_seen + 1 : 16;
    }
    // the second call with these arguments: capture (nothing executes while capturing)
    if (repeated && !cc->capture &&
        cudaStreamCreateWithFlags(&cc->capture, cudaStreamNonBlocking) != cudaSuccess) {
      cc->capture = nullptr;
      cudaGetLastError();
    }
    if (repeated && cc->capture &&
        (rc = pool_alloc(ctx, &acc, words, s)) == P1_OK &&
        (rc = alloc_rows()) == P1_OK) {
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      bool ok = cudaStreamBeginCapture(cc->capture, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        enqueue(cc->capture);
        ok = cudaStreamEndCapture(cc->capture, &graph) == cudaSuccess && graph != nullptr;
      }
      if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      if (ok) {
        if ((int)cc->entries.size() >= COLD_GRAPH_SLOTS) {  // the least recently used leaves
          int old = 0;
          for (int i = 1; i < (int)cc->entries.size(); ++i)
            if (cc->entries[i].tick < cc->entries[old].tick) old = i;
          ColdGraph &g = cc->entries[old];
          cudaStreamSynchronize(g.key.stream);
          cudaGraphExecDestroy(g.exec);
          pool_free(ctx, g.acc, g.key.stream);
          pool_free(ctx, g.slots, g.key.stream);
          cc->entries.erase(cc->entries.begin() + old);
        }
        cc->entries.push_back(ColdGraph{
            key, exec, acc, slim ? (void *)slots : (half ? (void *)halves : (void *)rec),
            ++cc->clock});
        P1_CUDA(cudaGraphLaunch(exec, s));
        return P1_OK;
      }
      cudaGetLastError();  // capture refused: the plain launch sequence below
    }
    release();
    acc = nullptr; rec = nullptr; slots = nullptr; halves = nullptr;
    if (rc) return rc;
  }

  if ((rc = pool_alloc(ctx, &acc, words, s)) || (rc = alloc_rows()) ||
      (quad && ((rc = pool_alloc(ctx, &d_weights, (size_t)qa->n_qp, s)) ||
                (rc = pool_alloc(ctx, &d_points, (size_t)2 * qa->n_qp, s))))) {
    release();
    return rc;
  }
  enqueue(s);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("p1_assemble_cold: %s", cudaGetErrorString(cudaGetLastError()));
    rc = P1_ECUDA;
  }
  release();
  return rc;
}

extern "C" int64_t p1_plan_rows(const p1_plan *p) { return p ? p->n_own : -1; }

extern "C" int p1_plan_destroy(p1_plan *p) {
  if (!p) return P1_OK;
  DeviceGuard guard(p->ctx);
  void *blocks[] = {p->acc, p->rec, p->rec_elem, p->topo, p->slot_of, p->tspill, p->slim,
                    p->ovf_off, p->ovf_rec, p->ovf_elem, p->indptr, p->slow_rows, p->big_rows,
                    p->big_first, p->twin, p->slotcode, p->fb_rows, p->tctr, p->ovf_cursor,
                    p->spill.rec, p->spill.row, p->spill.elem, p->d_spill, p->d_weights,
                    p->d_points};
  for (void *b : blocks) pool_free(p->ctx, b, p->stream);
  delete p;
  return P1_OK;
}

extern "C" int p1_plan_info(const p1_plan *p, int64_t *nnz, int64_t *max_index,
                            int64_t *n_elements, int64_t *n_nodes) {
  P1_REQUIRE(p, "null plan");
  if (nnz) *nnz = p->nnz;
  if (max_index) *max_index = p->max_index;
  if (n_elements) *n_elements = p->M;
  if (n_nodes) *n_nodes = p->N;
  return P1_OK;
}
