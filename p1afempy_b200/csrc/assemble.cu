// assemble.cu -- numeric phase: the owner of a CSR row sums the entries its
// records carry (written by the scatter, plan.cu: the local matrix of every
// element is computed ONCE, by the thread that owns the element, with the
// reference's arithmetic, plan.cuh: *Values) in a fixed order and writes the
// finished row (sorted unique columns, explicit zeros kept).
//
// Fast path: one thread per row, its <= 8 records in registers (fill_rows_kernel).
// HBM traffic per triangle: 96 B records + 42 B CSR output.
#include "plan.cuh"

namespace p1 {

constexpr int THREADS = 256;

// ---------------------------------------------------- fast path: thread per row
// Thread t of a block owns row row0+t.  It pulls its <= 8 records straight into
// registers -- a record is exactly one 32-byte sector, so a per-thread 256-bit
// load wastes nothing (staging the records through shared memory was measured
// slower: bank conflicts on the per-row reads) -- does the all-pairs comparison
// fully unrolled (rank of every column, the mate record sharing each
// off-diagonal entry, verification that the fan is closed and repeat-free),
// builds the row in a shared-memory image of the block's contiguous CSR span,
// and the block streams the image out with coalesced stores.
#ifndef P1_FILL_MIN_BLOCKS
#define P1_FILL_MIN_BLOCKS 20
#endif
#ifndef P1_FILL_ROWS
#define P1_FILL_ROWS 32
#endif
constexpr int FR = P1_FILL_ROWS;  // rows per block = threads per block
constexpr int CAP_OUT = FR * 10;  // CSR entries staged per block (9 per regular row + slack)

__device__ void slow_row_inline(int self, int deg, const Rec *nb, int *cs_out, double *vs_out);

// One row from its <= 8 records: cs / vs point at the row's place (the block's
// staged image or the output itself; cs may be null).  False when the fan is not
// closed and repeat-free.
// HALF: the 16-byte slots {next, prev|k, v} of p1_assemble_cold's mass matrix (local row
// (2v, v, v)) instead of records.
// stride: distance between the row's slots, in slots (1, or Own::lay)
template <bool HALF = false>
__device__ __forceinline__ bool row_from_records(int self, int deg, const Rec *row, int *cs,
                                                 double *vs, int64_t stride = 1) {
  int nx[GROUP], pv[GROUP], slot[GROUP];
  double vd[GROUP], vn[GROUP], vp[GROUP], mvp[GROUP];
  int slot_self = 0;
#pragma unroll
  for (int i = 0; i < GROUP; ++i) {
    nx[i] = 0x7fffffff - i;  // never below a valid column, never equal to one
    pv[i] = -2 - i;
    vd[i] = vn[i] = vp[i] = 0.;
    if (i < deg) {
      if constexpr (HALF) {
        const int4 h = __ldg(reinterpret_cast<const int4 *>(row) + i * stride);
        nx[i] = h.x; pv[i] = h.y & NODE_MASK;
        vn[i] = vp[i] = __hiloint2double(h.w, h.z);
        vd[i] = 2. * vn[i];
      } else {
        const Rec r = load_rec(row + i * stride);
        nx[i] = r.nx; pv[i] = rec_prev(r);
        vd[i] = r.vd; vn[i] = r.vn; vp[i] = r.vp;
      }
    }
  }
  // all pairs, in registers.  (Measured and dropped: the mate as an INDEX with the (k,k+2)
  // entries fetched from a per-thread column in shared memory, and "no repeats" tested on
  // the ranks instead of pair by pair -- 140 of 1030 instructions less, 16 shared-memory
  // operations more per row: 1.99 -> 2.09 ms at 64M.  The kernel is bound by its LSU
  // wavefronts (61 % of peak) and HBM (58 %), not by issue slots (46 %).)
  bool ok = true;
#pragma unroll
  for (int i = 0; i < GROUP; ++i) {
    int below = 0;
    bool mated = false;
    mvp[i] = 0.;
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      below += nx[j] < nx[i];
      if (pv[j] == nx[i]) { mvp[i] = vp[j]; mated = true; }
      if (j > i) ok = ok && nx[j] != nx[i] && pv[j] != pv[i];
    }
    const bool live = i < deg;
    // closed, repeat-free fan: every next has a mate, nothing equals the row
    ok = ok && (!live || (mated && nx[i] != self && pv[i] != self));
    slot[i] = below + (self < nx[i]);
    slot_self += live && nx[i] < self;
  }
  if (!ok) return false;
  // diagonal in a fixed order: park the (k,k) entries at their columns and
  // add them up in ascending column order
#pragma unroll
  for (int i = 0; i < GROUP; ++i)
    if (i < deg) vs[slot[i]] = vd[i];
  double diag = 0.;
  for (int q = 0; q <= deg; ++q)
    if (q != slot_self) diag += vs[q];
#pragma unroll
  for (int i = 0; i < GROUP; ++i)
    if (i < deg) {
      if (cs) cs[slot[i]] = nx[i];
      vs[slot[i]] = vn[i] + mvp[i];
    }
  if (cs) cs[slot_self] = self;
  vs[slot_self] = diag;
  return true;
}

#ifndef P1_FILL_MIN_BLOCKS_HALF
#define P1_FILL_MIN_BLOCKS_HALF P1_FILL_MIN_BLOCKS
#endif
template <bool HALF = false>
__global__ void __launch_bounds__(FR, HALF ? P1_FILL_MIN_BLOCKS_HALF : P1_FILL_MIN_BLOCKS)
fill_rows_kernel(int n_own, Own own, int exact,
                 const unsigned long long *__restrict__ acc, const Rec *__restrict__ rec,
                 const int *__restrict__ indptr, int *__restrict__ indices,
                 double *__restrict__ data, int *__restrict__ flags, int inline_slow = 0) {
  __shared__ int cols_s[CAP_OUT];
  __shared__ double vals_s[CAP_OUT];
  // (p1_assemble_cold: the output may be too small, or the mesh needs the plan route)
  if (*(volatile const int *)flags & (FLAG_SLOWPATH | FLAG_INDEX_RANGE)) return;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * FR;
  const int nrows = min(FR, n_own - row0);
  const int obeg = __ldg(indptr + row0), oend = __ldg(indptr + row0 + nrows);
  const int span = oend - obeg;
  const bool out_staged = span <= CAP_OUT;  // block-uniform
  int deg = 0, ob = 0;
  unsigned hash = 0;
  if (tid < nrows) {
    const unsigned long long v = __ldg(acc + row0 + tid);
    deg = (int)(unsigned)v;
    hash = (unsigned)(v >> 32);
    ob = __ldg(indptr + row0 + tid);
  }
  // the row's 8 slots (HALF: 16 bytes each), `stride` slots apart
  const int64_t stride = own.lay ? own.lay : 1;
  const int64_t first = own.pos(row0 + tid, 0);
  const Rec *row = HALF ? reinterpret_cast<const Rec *>(reinterpret_cast<const int4 *>(rec) + first)
                        : rec + first;
  const bool process = deg >= 1 && deg <= GROUP && (exact || hash == 0u);
  if (process) {
    // staged: the row is built in the block's image; otherwise (irregular rows
    // inside the block made the span oversize) directly in the output
    int *cs = out_staged ? cols_s + (ob - obeg) : (indices ? indices + ob : nullptr);
    double *vs = out_staged ? vals_s + (ob - obeg) : data + ob;
    // exact plan: a row that fails is in slow_rows.  Hash plan: the shortcut
    // "closed fan => 1 + #records columns" does not hold for this row.
    if (!row_from_records<HALF>(own.global(row0 + tid), deg, row, cs, vs, stride) && !exact)
      atomicOr(flags, FLAG_RETRY);
  } else if (inline_slow && deg >= 1 && deg <= GROUP) {  // an open fan (p1_assemble_cold)
    int *cs = out_staged ? cols_s + (ob - obeg) : (indices ? indices + ob : nullptr);
    double *vs = out_staged ? vals_s + (ob - obeg) : data + ob;
    if (HALF || stride != 1) {  // the row's records side by side, in local memory
      Rec made[GROUP];
      for (int i = 0; i < deg; ++i) {
        if constexpr (HALF) {
          const int4 h = reinterpret_cast<const int4 *>(row)[i * stride];
          made[i].nx = h.x;
          made[i].pvk = h.y;
          made[i].vn = made[i].vp = __hiloint2double(h.w, h.z);
          made[i].vd = 2. * made[i].vn;
        } else {
          made[i] = row[i * stride];
        }
      }
      slow_row_inline(own.global(row0 + tid), deg, made, cs, vs);
    } else {
      slow_row_inline(own.global(row0 + tid), deg, row, cs, vs);
    }
  }
  __syncthreads();
  if (out_staged) {
    // rows the slow / big kernels own get whatever the image holds here; those
    // kernels run afterwards on the same stream and overwrite them
    for (int i = tid; i < span; i += FR) {
      if (indices) indices[(int64_t)obeg + i] = cols_s[i];
      data[(int64_t)obeg + i] = vals_s[i];
    }
  }
}

// Tiled plans: the regular rows the tiles did not complete (their fan crosses a tile
// boundary), listed in no particular order; the row is built in its place in the
// output.  The list's length is read on the device.
__global__ void __launch_bounds__(128)
fill_list_kernel(const int *__restrict__ list, const int *__restrict__ n_list,
                 const unsigned long long *__restrict__ acc, const Rec *__restrict__ rec,
                 const int *__restrict__ indptr, int *__restrict__ indices,
                 double *__restrict__ data, int *__restrict__ flags) {
  const int n = *n_list;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int a = list[idx];
    const int deg = (int)(unsigned)__ldg(acc + a);
    const int ob = __ldg(indptr + a);
    if (!row_from_records(a, deg, rec + (int64_t)a * GROUP, indices ? indices + ob : nullptr,
                          data + ob))
      atomicOr(flags, FLAG_RETRY);
  }
}

// ------------------------------------- fast path of slim plans: gather and compute
// (multi-GPU stiffness / mass, plan.cu: plan_build).  The slots hold only
// (next, prev|k): 64 bytes per row instead of 256.  The thread gathers the
// coordinates of its node and of the `next` of every record (the `prev` of a
// record is the `next` of the record before it in a closed fan; both indices come
// out of the all-pairs comparison and address per-thread columns in shared
// memory), and recomputes entry (k,k), (k,k+1), (k,k+2) of each incident element
// with the arithmetic of the scatter (plan.cuh: slim_values).  1830 instructions
// per row against ~900 of fill_rows_kernel: issue-bound, not HBM-bound.
#ifndef P1_GATHER_MIN_BLOCKS
#define P1_GATHER_MIN_BLOCKS 20
#endif
template <class V>
__global__ void __launch_bounds__(FR, P1_GATHER_MIN_BLOCKS)
fill_gather_kernel(int n_own, Own own, const unsigned long long *__restrict__ acc,
                   const int2 *__restrict__ slim, V values,
                   const int *__restrict__ indptr, int *__restrict__ indices,
                   double *__restrict__ data, int *__restrict__ flags, int inline_slow = 0) {
  __shared__ int cols_s[CAP_OUT];
  __shared__ double vals_s[CAP_OUT];
  // per-thread columns ([record][thread]: conflict-free for any record index):
  // the coordinates of every `next`, and the (k,k+2) entry of every record
  __shared__ double2 zs[(GROUP + 1) * FR];  // column GROUP: the row's own node
  __shared__ double vps[GROUP * FR];
  __shared__ unsigned char src_s[GROUP * FR];  // record whose `next` is my `prev`
  if (*(volatile const int *)flags & (FLAG_SLOWPATH | FLAG_INDEX_RANGE)) return;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * FR;
  const int nrows = min(FR, n_own - row0);
  const int obeg = __ldg(indptr + row0), oend = __ldg(indptr + row0 + nrows);
  const int span = oend - obeg;
  const bool out_staged = span <= CAP_OUT;  // block-uniform
  int deg = 0, ob = 0;
  unsigned hash = 0;
  if (tid < nrows) {
    const unsigned long long v = __ldg(acc + row0 + tid);
    deg = (int)(unsigned)v;
    hash = (unsigned)(v >> 32);
    ob = __ldg(indptr + row0 + tid);
  }
  const bool process = deg >= 1 && deg <= GROUP && hash == 0u;
  const int self = own.global(row0 + tid);
  if (process) {
    const int4 *row4 = reinterpret_cast<const int4 *>(slim + (int64_t)(row0 + tid) * GROUP);
    int4 q[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      q[h] = make_int4(0, 0, 0, 0);
      if (2 * h < deg) q[h] = __ldg(row4 + h);
    }
    zs[GROUP * FR + tid] = __ldg(values.coords + self);
    int nx[GROUP], pv[GROUP], kk[GROUP];
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      const int sx = (i & 1) ? q[i >> 1].z : q[i >> 1].x;
      const int sy = (i & 1) ? q[i >> 1].w : q[i >> 1].y;
      nx[i] = 0x7fffffff - i;  // never below a valid column, never equal to one
      pv[i] = -2 - i;
      kk[i] = 0;
      if (i < deg) {
        nx[i] = sx; pv[i] = sy & NODE_MASK; kk[i] = (int)((unsigned)sy >> NODE_BITS);
        zs[i * FR + tid] = __ldg(values.coords + sx);
      }
    }
    int *cs = out_staged ? cols_s + (ob - obeg) : indices + ob;
    double *vs = out_staged ? vals_s + (ob - obeg) : data + ob;
    const bool put_cols = out_staged || indices != nullptr;
    bool ok = true;
    int slot_self = 0;
    unsigned ranks = 0, sourced = 0;
    int meta[GROUP];  // slot | mate << 4
    double vn[GROUP];
    // all pairs: the rank of every column and the record AFTER each record (its
    // `prev` is my `next`: its (k,k+2) entry shares my column).  The record BEFORE
    // (its `next` is my `prev`: my third vertex) is the inverse permutation.
#pragma unroll
    for (int i = 0; i < GROUP; ++i) {
      int below = 0, mate = -1;
#pragma unroll
      for (int j = 0; j < GROUP; ++j) {
        below += nx[j] < nx[i];
        if (pv[j] == nx[i]) mate = j;
      }
      const bool live = i < deg;
      // closed, repeat-free fan: every next has a mate, the mates are all different
      // (so every prev has a source), nothing equals the row; distinct ranks <=>
      // distinct nexts
      ok = ok && (!live || (mate >= 0 && nx[i] != self && pv[i] != self));
      if (live) {
        ranks |= 1u << below;
        sourced |= 1u << max(mate, 0);
        src_s[max(mate, 0) * FR + tid] = (unsigned char)i;
      }
      slot_self += live && nx[i] < self;
      meta[i] = (below + (self < nx[i])) | (max(mate, 0) << 4);
    }
    ok = ok && ranks == (1u << deg) - 1u && sourced == (1u << deg) - 1u;
    if (ok) {
#pragma unroll
      for (int i = 0; i < GROUP; ++i) {
        vn[i] = 0.;
        if (i < deg) {
          // vertex c of the element is column (c - k) mod 3 of {self, next, prev}
          const int k = kk[i], src = src_s[i * FR + tid];
          const int c0 = k == 0 ? GROUP : (k == 1 ? src : i);
          const int c1 = k == 0 ? i : (k == 1 ? GROUP : src);
          const int c2 = k == 0 ? src : (k == 1 ? i : GROUP);
          double v[3][3];
          values.rows(0, nullptr, zs[c0 * FR + tid], zs[c1 * FR + tid], zs[c2 * FR + tid], v);
          const double vd = k == 0 ? v[0][0] : (k == 1 ? v[1][0] : v[2][0]);
          vn[i] = k == 0 ? v[0][1] : (k == 1 ? v[1][1] : v[2][1]);
          vps[i * FR + tid] = k == 0 ? v[0][2] : (k == 1 ? v[1][2] : v[2][2]);
          const int slot = meta[i] & 15;
          vs[slot] = vd;  // parked: the diagonal is summed in ascending column order
          if (put_cols) cs[slot] = nx[i];
        }
      }
    }
    if (!ok) {
      // (whatever was written is overwritten by the rebuilt exact plan)
      atomicOr(flags, FLAG_RETRY);
    } else {
      double diag = 0.;
      for (int c = 0; c <= deg; ++c)
        if (c != slot_self) diag += vs[c];
#pragma unroll
      for (int i = 0; i < GROUP; ++i)
        if (i < deg) vs[meta[i] & 15] = vn[i] + vps[(meta[i] >> 4) * FR + tid];
      if (put_cols) cs[slot_self] = self;
      vs[slot_self] = diag;
    }
  } else if (inline_slow && deg >= 1 && deg <= GROUP) {  // an open fan (p1_assemble_cold)
    Rec made[GROUP];
    const double2 za = __ldg(values.coords + self);
    for (int i = 0; i < deg; ++i) {
      const int2 e = slim[(int64_t)(row0 + tid) * GROUP + i];
      made[i].nx = e.x;
      made[i].pvk = e.y;
      slim_values(values, (int)((unsigned)e.y >> NODE_BITS), za, __ldg(values.coords + e.x),
                  __ldg(values.coords + (e.y & NODE_MASK)), made[i].vd, made[i].vn, made[i].vp);
    }
    slow_row_inline(self, deg, made,
                    out_staged ? cols_s + (ob - obeg) : (indices ? indices + ob : nullptr),
                    out_staged ? vals_s + (ob - obeg) : data + ob);
  }
  __syncthreads();
  if (out_staged) {
    for (int i = tid; i < span; i += FR) {
      if (indices) indices[(int64_t)obeg + i] = cols_s[i];
      data[(int64_t)obeg + i] = vals_s[i];
    }
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

// sorted unique columns {self} U {next_i} U {prev_i} by insertion into cs (2 deg + 1 ints)
__device__ __forceinline__ int slow_row_columns(int self, int deg, const Rec *nb, int *cs) {
  int n = insert_unique(cs, 0, self);
  for (int i = 0; i < deg; ++i) {
    n = insert_unique(cs, n, rec_next(nb[i]));
    n = insert_unique(cs, n, rec_prev(nb[i]));
  }
  return n;
}

// the values of such a row, accumulated in vs[0 .. n) in a fixed order: ascending
// `next`, ties by `prev`
__device__ __forceinline__ void slow_row_values(int self, int deg, const Rec *nb, const int *cs,
                                                int n, double *vs) {
  for (int j = 0; j < n; ++j) vs[j] = 0.;
  const int js = lower_bound(cs, n, self);
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

// p1_assemble_cold fills the rare irregular row (an open fan: a mesh-boundary node)
// from inside the fast kernels, by the thread that owns the row: no row list, no
// extra launch.  Same routine, same bits as fill_slow_kernel.  (noinline: the hot
// path keeps its registers)
// The irregular row a mesh has many of -- ONE open, repeat-free fan around a node of the
// mesh boundary: deg + 2 columns (the fan's first `prev` has no `next` to merge with),
// found by the all-pairs comparison of row_from_records, in registers.  False: not such a
// row.  Same sums as slow_row_values: an entry is next-value + prev-value of (at most) two
// records, the diagonal adds the (k,k) entries in ascending column of `next`.
// nb: global or local memory.
__device__ __forceinline__ bool open_fan_row(int self, int deg, const Rec *nb, int *cs,
                                             double *vs) {
  // Only the indices live in registers, the values are fetched again where they are
  // used: a callee that needs 90 registers made fill_rows_kernel spill one of its own
  // in the prologue of EVERY thread (+0.03 ms at 64M triangles).
  int nx[GROUP], pv[GROUP], meta[GROUP];  // meta: slot | mate << 4 (mate 8: none)
#pragma unroll
  for (int i = 0; i < GROUP; ++i) {
    nx[i] = 0x7fffffff - i;
    pv[i] = -2 - i;
    if (i < deg) {
      nx[i] = nb[i].nx;
      pv[i] = nb[i].pvk & NODE_MASK;
    }
  }
  bool ok = true;
  int slot_self = 0, unmated = 0, n_first = 0, first_col = 0, first_rec = 0;
#pragma unroll
  for (int i = 0; i < GROUP; ++i) {
    int below = 0, mate = GROUP;
    bool sourced = false;
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      below += nx[j] < nx[i];
      mate = pv[j] == nx[i] ? j : mate;
      sourced = sourced || nx[j] == pv[i];
      if (j > i) ok = ok && nx[j] != nx[i] && pv[j] != pv[i];
    }
    const bool live = i < deg;
    ok = ok && (!live || (nx[i] != self && pv[i] != self));
    unmated += live && mate == GROUP;
    if (live && !sourced) { ++n_first; first_col = pv[i]; first_rec = i; }
    meta[i] = (below + (self < nx[i])) | (mate << 4);
    slot_self += live && nx[i] < self;
  }
  if (!ok || unmated != 1 || n_first != 1) return false;
  int slot_first = self < first_col;
#pragma unroll
  for (int i = 0; i < GROUP; ++i) {
    slot_first += nx[i] < first_col;  // (unused slots hold columns above any node)
    meta[i] += first_col < nx[i];
  }
  slot_self += first_col < self;
  vs[slot_first] = 0.;
#pragma unroll
  for (int i = 0; i < GROUP; ++i)
    if (i < deg) vs[meta[i] & 15] = nb[i].vd;
  double diag = 0.;
  for (int q = 0; q < deg + 2; ++q)
    if (q != slot_self) diag += vs[q];
#pragma unroll
  for (int i = 0; i < GROUP; ++i)
    if (i < deg) {
      const int mate = meta[i] >> 4;
      if (cs) cs[meta[i] & 15] = nx[i];
      vs[meta[i] & 15] = nb[i].vn + (mate < GROUP ? nb[mate].vp : 0.);
    }
  if (cs) { cs[slot_first] = first_col; cs[slot_self] = self; }
  vs[slot_first] = 0. + nb[first_rec].vp;
  vs[slot_self] = diag;
  return true;
}

__device__ __noinline__ void slow_row_inline(int self, int deg, const Rec *nb, int *cs_out,
                                             double *vs_out) {
#ifndef P1_AB_OLD_OPEN
  if (open_fan_row(self, deg, nb, cs_out, vs_out)) return;
#endif
  int cs[2 * GROUP + 1];
  const int n = slow_row_columns(self, deg, nb, cs);
  if (cs_out)
    for (int j = 0; j < n; ++j) cs_out[j] = cs[j];
  slow_row_values(self, deg, nb, cs, n, vs_out);
}

// (slim plans: V recomputes the <= 8 records of the row into local memory)
template <class V>
__global__ void __launch_bounds__(SLOW_THREADS)
fill_slow_kernel(const int *__restrict__ slow_rows, int n_slow, Own own, Rows rows, V values,
                 const int *__restrict__ indptr, int *__restrict__ indices,
                 double *__restrict__ data, int *__restrict__ verify = nullptr,
                 const int *__restrict__ n_slow_dev = nullptr,
                 const int *__restrict__ status = nullptr) {
  __shared__ int cols_s[SLOW_THREADS * SLOW_COLS];
  // n_slow_dev: the list's length is read on the device and the grid strides over it
  if (status && (*(volatile const int *)status & (FLAG_SLOWPATH | FLAG_INDEX_RANGE))) return;
  if (n_slow_dev) n_slow = *n_slow_dev;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_slow;
       idx += gridDim.x * blockDim.x) {
  const int a = slow_rows[idx];
  const int deg = rows.deg(a);
  const int self = own.global(a);
  const Rec *nb;
  Rec made[GROUP];
  if constexpr (V::needs_xy) {
    const double2 za = __ldg(values.coords + self);
    for (int i = 0; i < deg && i < GROUP; ++i) {
      const int2 e = rows.slim[(int64_t)a * GROUP + i];
      made[i].nx = e.x;
      made[i].pvk = e.y;
      slim_values(values, (int)((unsigned)e.y >> NODE_BITS), za, __ldg(values.coords + e.x),
                  __ldg(values.coords + (e.y & NODE_MASK)), made[i].vd, made[i].vn, made[i].vp);
    }
    nb = made;
  } else {
    nb = rows.ptr(a, deg);
  }
  int *cs = cols_s + threadIdx.x * SLOW_COLS;
  const int n = slow_row_columns(self, deg, nb, cs);
  const int64_t base = indptr[a];
  if (verify && n != indptr[a + 1] - indptr[a]) {
    // tiled plans take the row length from the record count (closed fan: 1 + deg,
    // else 2 + deg) without looking at the records: this row is not that simple
    atomicOr(verify, FLAG_RETRY);
    continue;
  }
  if (indices)
    for (int j = 0; j < n; ++j) indices[base + j] = cs[j];
  slow_row_values(self, deg, nb, cs, n, data + base);
  }
}

// ------------------------------------------------ big path: one block per row
__device__ __forceinline__ int big_candidate2(int self, const Rec *nb, int j) {
  if (j == 0) return self;
  const Rec p = nb[(j - 1) >> 1];
  return ((j - 1) & 1) ? rec_prev(p) : rec_next(p);
}

__global__ void __launch_bounds__(THREADS)
fill_big_kernel(const int *__restrict__ big_rows, Own own, Rows rows,
                const unsigned char *__restrict__ first,
                const int *__restrict__ indptr, int *__restrict__ indices,
                double *__restrict__ data) {
  const int a = big_rows[blockIdx.x];
  const int deg = rows.deg(a);
  const Rec *nb = rows.ptr(a, deg);
  const unsigned char *fl = first + 2 * (int64_t)rows.ovf_off[a] + a;
  const int self = own.global(a);
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
// `live` = number of valid records per group of 8 slots (null: all n valid)
template <class V>
__global__ void __launch_bounds__(THREADS)
refresh_kernel(const int *__restrict__ elements, int64_t n,
               const unsigned long long *__restrict__ live, Rec *__restrict__ rec,
               const int *__restrict__ rec_elem, V values) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (live) {
      const unsigned deg = (unsigned)live[i >> 3];
      if (deg > (unsigned)GROUP || (unsigned)(i & 7) >= deg) continue;
    }
    const int64_t t = rec_elem[i] >> 2;
    const int e[3] = {__ldg(elements + 3 * t), __ldg(elements + 3 * t + 1),
                      __ldg(elements + 3 * t + 2)};
    double v[3][3];
    double2 z0 = make_double2(0., 0.), z1 = z0, z2 = z0;
    if constexpr (V::needs_xy) {
      z0 = __ldg(values.coords + e[0]);
      z1 = __ldg(values.coords + e[1]);
      z2 = __ldg(values.coords + e[2]);
    }
    values.rows(t, e, z0, z1, z2, v);
    Rec r = rec[i];
    const int k = rec_k(r);
    r.vd = k == 0 ? v[0][0] : (k == 1 ? v[1][0] : v[2][0]);
    r.vn = k == 0 ? v[0][1] : (k == 1 ? v[1][1] : v[2][1]);
    r.vp = k == 0 ? v[0][2] : (k == 1 ? v[1][2] : v[2][2]);
    store_rec(rec + i, r);
  }
}

template <class V>
static void launch_refresh(p1_plan *plan, V values, cudaStream_t s) {
  p1_ctx *ctx = plan->ctx;
  const int64_t n_slots = plan->n_own * GROUP;
  if (n_slots > 0)
    refresh_kernel<<<min(grid_for(n_slots, THREADS), (unsigned)(ctx->sm_count * 32)), THREADS,
                     0, s>>>(plan->elements, n_slots, plan->acc, plan->rec, plan->rec_elem,
                             values);
  if (plan->n_ovf > 0)
    refresh_kernel<<<min(grid_for(plan->n_ovf, THREADS), (unsigned)(ctx->sm_count * 32)),
                     THREADS, 0, s>>>(plan->elements, plan->n_ovf, nullptr, plan->ovf_rec,
                                      plan->ovf_elem, values);
}

}  // namespace p1

using namespace p1;

void p1::launch_fill_fallback(p1_plan *plan, int *indices, double *data, cudaStream_t s) {
  p1_ctx *ctx = plan->ctx;
  P1_TIMED(ctx, "fill_list", s,
           (fill_list_kernel<<<ctx->sm_count * 8, 128, 0, s>>>(
               plan->fb_rows, plan->tctr, plan->acc, plan->rec, plan->indptr, indices, data,
               ctx->d_flags)));
  if (plan->n_slow > 0)
    P1_TIMED(ctx, "fill_slow", s,
             (fill_slow_kernel<<<grid_for(plan->n_slow, SLOW_THREADS), SLOW_THREADS, 0, s>>>(
                 plan->slow_rows, plan->n_slow, plan->own, rows_of(plan), NoValues{},
                 plan->indptr, indices, data, ctx->d_flags)));
}

void p1::launch_fill_cold(p1_ctx *ctx, const Own &own, int n_own, const Rows &rows, int slim_op,
                          const double2 *coords, const int *indptr, int *indices, double *data,
                          int *status, cudaStream_t s) {
  if (n_own <= 0) return;
  prof_begin(ctx, "fill_rows", s);
  if (slim_op == P1_OP_STIFFNESS)
    fill_gather_kernel<<<grid_for(n_own, FR), FR, 0, s>>>(n_own, own, rows.acc, rows.slim,
                                                         StiffnessValues{coords}, indptr, indices,
                                                         data, status, 1);
  else if (slim_op == P1_OP_MASS)
    fill_gather_kernel<<<grid_for(n_own, FR), FR, 0, s>>>(n_own, own, rows.acc, rows.slim,
                                                         MassValues{coords}, indptr, indices, data,
                                                         status, 1);
  else if (rows.half)
    fill_rows_kernel<true><<<grid_for(n_own, FR), FR, 0, s>>>(
        n_own, own, 0, rows.acc, reinterpret_cast<const Rec *>(rows.half), indptr, indices, data,
        status, 1);
  else
    fill_rows_kernel<false><<<grid_for(n_own, FR), FR, 0, s>>>(n_own, own, 0, rows.acc, rows.rec,
                                                              indptr, indices, data, status, 1);
  prof_end(ctx, s);
}

// Slim plans (multi-GPU stiffness / mass, no row with more than 8 records): the
// values are computed while filling, from the plan's coordinates or the ones
// passed here.
template <class V>
static void launch_slim(p1_plan *plan, V values, int *indices, double *data, cudaStream_t s) {
  p1_ctx *ctx = plan->ctx;
  const int n_own = (int)plan->n_own;
  if (n_own > 0)
    P1_TIMED(ctx, "fill_rows", s,
             (fill_gather_kernel<<<grid_for(n_own, FR), FR, 0, s>>>(
                 n_own, plan->own, plan->acc, plan->slim, values, plan->indptr, indices, data,
                 ctx->d_flags)));
  if (plan->n_slow > 0)
    P1_TIMED(ctx, "fill_slow", s,
             (fill_slow_kernel<<<grid_for(plan->n_slow, SLOW_THREADS), SLOW_THREADS, 0, s>>>(
                 plan->slow_rows, plan->n_slow, plan->own, rows_of(plan), values, plan->indptr,
                 indices, data)));
}

static int assemble_slim(p1_plan *plan, int op, const double *coords, int32_t *indptr,
                         int32_t *indices, double *data, void *stream) {
  p1_ctx *ctx = plan->ctx;
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  if (op == P1_OP_STORED) op = plan->val_op;
  else if (coords) plan->coords = (const double2 *)coords;
  if (op != P1_OP_STIFFNESS && op != P1_OP_MASS) {
    set_error("p1_assemble_csr: this plan assembles P1_OP_STIFFNESS or P1_OP_MASS only "
              "(create it with P1_PLAN_KEEP_ELEMENTS for other operators)");
    return P1_EINVAL;
  }
  P1_REQUIRE(plan->coords || plan->M == 0, "stiffness/mass need coords");
  plan->val_op = op;
  if (indptr)
    P1_CUDA(cudaMemcpyAsync(indptr, plan->indptr, ((size_t)plan->n_own + 1) * sizeof(int),
                            cudaMemcpyDeviceToDevice, s));
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), s));
  if (op == P1_OP_STIFFNESS) launch_slim(plan, StiffnessValues{plan->coords}, indices, data, s);
  else launch_slim(plan, MassValues{plan->coords}, indices, data, s);
  P1_CUDA(cudaGetLastError());
  P1_CUDA(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, s));
  P1_CUDA(cudaStreamSynchronize(s));
  if (ctx->h_flags[0] & FLAG_RETRY) {
    set_error("p1_assemble_csr: the mesh contradicts the plan's closed-fan shortcut "
              "(non-manifold vertex or degenerate element); rebuild the plan with "
              "P1_PLAN_EXACT");
    return P1_ERETRY;
  }
  return P1_OK;
}

extern "C" int p1_assemble_csr(const p1_plan *cplan, int op, const double *coords,
                               const double *local, int32_t *indptr,
                               int32_t *indices, double *data, void *stream) {
  P1_REQUIRE(cplan, "null plan");
  P1_REQUIRE(data || cplan->nnz == 0, "null data");
  p1_plan *plan = const_cast<p1_plan *>(cplan);
  if (plan->tiled) return assemble_tiled(plan, op, indptr, indices, data, stream);
  if (plan->slim) return assemble_slim(plan, op, coords, indptr, indices, data, stream);
  if (plan->topo) {
    set_error("p1_assemble_csr: a P1_PLAN_TOPOLOGY plan has no records to assemble");
    return P1_EINVAL;
  }
  p1_ctx *ctx = plan->ctx;
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  const int n_own = (int)plan->n_own;
  const Own own = plan->own;
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
    prof_begin(ctx, "refresh", s);
    if (op == P1_OP_STIFFNESS) launch_refresh(plan, StiffnessValues{xy}, s);
    else if (op == P1_OP_MASS) launch_refresh(plan, MassValues{xy}, s);
    else launch_refresh(plan, LocalValues{local}, s);
    prof_end(ctx, s);
    plan->val_op = op;
  }
  if (indptr)
    P1_CUDA(cudaMemcpyAsync(indptr, plan->indptr, ((size_t)n_own + 1) * sizeof(int),
                            cudaMemcpyDeviceToDevice, s));
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), s));
  if (n_own > 0)
    P1_TIMED(ctx, "fill_rows", s,
             (fill_rows_kernel<false><<<grid_for(n_own, FR), FR, 0, s>>>(
                 n_own, own, plan->exact ? 1 : 0, plan->acc, plan->rec, plan->indptr, indices,
                 data, ctx->d_flags)));
  if (plan->n_slow > 0)
    P1_TIMED(ctx, "fill_slow", s,
             (fill_slow_kernel<<<grid_for(plan->n_slow, SLOW_THREADS), SLOW_THREADS, 0, s>>>(
                 plan->slow_rows, plan->n_slow, own, rows_of(plan), NoValues{}, plan->indptr,
                 indices, data)));
  if (plan->n_big > 0)
    fill_big_kernel<<<plan->n_big, THREADS, 0, s>>>(
        plan->big_rows, own, rows_of(plan), plan->big_first, plan->indptr, indices, data);
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
