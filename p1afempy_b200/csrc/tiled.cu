// tiled.cu -- one-GPU cold assembly without the record round trip.
//
// The record pipeline (plan.cu scatter + assemble.cu fill) sends every (node,
// element) incidence through HBM as a 32-byte record: 192 of the ~290 bytes per
// triangle a step moves.  But meshes arrive in an element order that is local in
// space (refineNVB keeps the children of an element together,
// refinement.py:633-716): on the 64M-triangle NVB mesh half of the nodes have
// their whole fan inside 27 consecutive elements, 91 % inside a tile of 1024.
// Here a block takes a tile of TC consecutive elements and finishes every row
// whose fan lies inside the tile in shared memory; only the incidences of the
// other rows (tile-boundary nodes, mesh-boundary nodes, nodes with more than 8
// elements) travel through records and are finished by the fill kernels of
// assemble.cu.  Meshes without that locality lose nothing but the tile
// bookkeeping: all their incidences take the record path.
//
//   count   (one thread per 4 triangles) ONE 64-bit atomic per distinct node of a
//           triangle pair, as in the scatter: count + closed-fan hash, and the
//           slot of every incidence in its row, remembered in 2 bytes per triangle
//   split + scan -> row pointer (closed fan: 1 + #records, open fan: 2 + #records;
//           both verified when the rows are filled)
//   tile    hash-free of atomics: the incidence that holds slot 0 of a node inserts
//           the node into the tile's table (write, barrier, read back), the nodes
//           are numbered in ascending id through their runs of consecutive ids,
//           incidences park (element, k) in the 8 cells of their node, a node whose
//           cells are all filled is complete: its row is built by one thread from
//           shared memory in the arithmetic order of fill_rows_kernel (same bits)
//           and streamed to its final place in the CSR arrays
//   fill    the remaining rows, from records (assemble.cu)
#include "plan.cuh"

namespace p1 {

constexpr int TC = 1024;    // elements per tile
constexpr int TT = 256;     // threads per tile: 4 elements each
constexpr int HT = 4096;    // hash table slots
constexpr int NMAX = 768;   // nodes a tile can own (3 per thread in the compaction)
constexpr int HMAX = 256;   // runs of consecutive node ids a tile can sort
constexpr int FOREIGN = 1023;  // node not owned by the tile
constexpr int IMG = TT * 9;    // CSR entries staged per batch of TT rows
constexpr int CT = 256;        // count kernel block

__device__ __forceinline__ unsigned tile_hash(int id) {
  return ((unsigned)id * 0x9E3779B1u) >> 20;  // top 12 bits: HT = 4096
}
static_assert(HT == 4096, "tile_hash yields 12 bits");

// ------------------------------------------------------------------ count
// references of two consecutive triangles grouped in registers (plan.cu:
// scatter_pair): one atomic per distinct node; slot[i] = position of reference i
// in its row (saturated to 15)
template <int NTRI>
__device__ __forceinline__ void count_pair(const int (&id)[6], int N,
                                           unsigned long long *__restrict__ acc, int &mx,
                                           bool &bad, unsigned (&slot)[6]) {
  bool valid[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    valid[j] = j < NTRI && (unsigned)id[3 * j] < (unsigned)N &&
               (unsigned)id[3 * j + 1] < (unsigned)N && (unsigned)id[3 * j + 2] < (unsigned)N;
    bad = bad || (j < NTRI && !valid[j]);
    if (valid[j]) mx = max(mx, max(id[3 * j], max(id[3 * j + 1], id[3 * j + 2])));
  }
  unsigned base[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const bool live = valid[i / 3];
    base[i] = 0;
    bool leader = live;
    unsigned rank = 0;
#pragma unroll
    for (int j = 0; j < i; ++j)
      if (live && valid[j / 3] && id[j] == id[i]) { leader = false; ++rank; base[i] = base[j]; }
    if (leader) {
      const int ej = i / 3, ek = i - 3 * ej;
      unsigned cnt = 1;
      unsigned h = mix32((unsigned)id[3 * ej + (ek == 2 ? 0 : ek + 1)]) -
                   mix32((unsigned)id[3 * ej + (ek == 0 ? 2 : ek - 1)]);
#pragma unroll
      for (int j = i + 1; j < 6; ++j)
        if (valid[j / 3] && id[j] == id[i]) {
          const int fj = j / 3, fk = j - 3 * fj;
          ++cnt;
          h += mix32((unsigned)id[3 * fj + (fk == 2 ? 0 : fk + 1)]) -
               mix32((unsigned)id[3 * fj + (fk == 0 ? 2 : fk - 1)]);
        }
      base[i] = (unsigned)atomicAdd(acc + id[i], ((unsigned long long)h << 32) + cnt);
    }
    slot[i] = live ? min(base[i] + rank, 15u) : 15u;
  }
}

__global__ void __launch_bounds__(CT)
count_kernel(const int *__restrict__ elements, int64_t M, int N,
             unsigned long long *__restrict__ acc, unsigned short *__restrict__ slotcode,
             int *__restrict__ flags) {
  int mx = -1;
  bool bad = false;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t quads = M / 4;
  const int4 *e4 = reinterpret_cast<const int4 *>(elements);
  for (int64_t g = tid; g < quads; g += stride) {
    const int4 p = __ldg(e4 + 3 * g), q = __ldg(e4 + 3 * g + 1), r = __ldg(e4 + 3 * g + 2);
    const int ab[6] = {p.x, p.y, p.z, p.w, q.x, q.y}, cd[6] = {q.z, q.w, r.x, r.y, r.z, r.w};
    unsigned s0[6], s1[6];
    count_pair<2>(ab, N, acc, mx, bad, s0);
    count_pair<2>(cd, N, acc, mx, bad, s1);
    const unsigned long long code =
        (unsigned long long)(s0[0] | s0[1] << 4 | s0[2] << 8) |
        (unsigned long long)(s0[3] | s0[4] << 4 | s0[5] << 8) << 16 |
        (unsigned long long)(s1[0] | s1[1] << 4 | s1[2] << 8) << 32 |
        (unsigned long long)(s1[3] | s1[4] << 4 | s1[5] << 8) << 48;
    *reinterpret_cast<unsigned long long *>(slotcode + 4 * g) = code;
  }
  for (int64_t t = 4 * quads + tid; t < M; t += stride) {  // the last M % 4 triangles
    const int id[6] = {__ldg(elements + 3 * t), __ldg(elements + 3 * t + 1),
                       __ldg(elements + 3 * t + 2), -1, -1, -1};
    unsigned s[6];
    count_pair<1>(id, N, acc, mx, bad, s);
    slotcode[t] = (unsigned short)(s[0] | s[1] << 4 | s[2] << 8);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx >= 0) atomicMax(flags + 1, mx);
  if (bad) atomicOr(flags, FLAG_INDEX_RANGE);
}

// Row lengths without looking at a record: closed fan -> 1 + #records; anything
// else is taken for a simple open fan (a mesh-boundary node): 2 + #records.  Both
// are verified when the rows are filled (FLAG_RETRY -> exact plan).
// counters: [0] n_big, [1] n_slow, [2] rows with more than 8 records, [5] their
// records beyond 8 (flags + 2 is passed)
__global__ void __launch_bounds__(256)
split_tiled_kernel(int n, const unsigned long long *__restrict__ acc, int *__restrict__ rowlen,
                   int *__restrict__ slow_rows, int *__restrict__ big_rows,
                   int *__restrict__ counters) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const unsigned long long v = acc[a];
  const int deg = (int)(unsigned)v;
  if (deg == 0) { rowlen[a] = 0; return; }
  const bool closed = (unsigned)(v >> 32) == 0u;
  rowlen[a] = deg + (closed ? 1 : 2);
  if (closed && deg <= GROUP) return;
  if (deg > GROUP) {
    atomicAdd(counters + 2, 1);
    atomicAdd(counters + 5, deg - GROUP);
  }
  if (deg > DEG_BIG) big_rows[atomicAdd(counters, 1)] = a;
  else slow_rows[atomicAdd(counters + 1, 1)] = a;
}

// ------------------------------------------------------------------- tile
struct TileArgs {
  const int *elements;
  int64_t M;
  int N;
  const unsigned long long *acc;
  const unsigned short *slotcode;
  const int *indptr;
  Rec *rec;
  const SpillDesc *spill;
  int *indices;
  double *data;
  int *fb_rows, *fb_count;
  int *flags;
};

// block-wide exclusive scan of one value per thread (TT threads); two barriers
__device__ __forceinline__ int tile_scan(int v, int *scratch, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < TT / 32; ++w) {
    const int s = scratch[w];
    before += w < warp ? s : 0;
    all += s;
  }
  __syncthreads();
  total = all;
  return before + inc - v;
}

__device__ __forceinline__ bool table_has(const int *T, int id) {
  unsigned s = tile_hash(id);
  for (;;) {
    const int v = T[s];
    if (v == -1) return false;  // (first: id may be -1 itself, the neighbour of node 0)
    if (v == id) return true;
    s = (s + 1) & (HT - 1);
  }
}

// largest r with sh[r] <= id (sh ascending, sh[0] <= id)
__device__ __forceinline__ int run_of(const int *sh, int H, int id) {
  int lo = 0, hi = H;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (sh[mid] <= id) lo = mid; else hi = mid;
  }
  return lo;
}

// One complete row from shared memory, in the arithmetic order of
// fill_rows_kernel: (k,k) entries parked at their columns and added in ascending
// column order; an off-diagonal entry is (k,k+1) of the record whose next it is
// plus (k,k+2) of the record whose prev it is.  G = 4 or 8 records at most.
template <int G, class V>
__device__ __forceinline__ bool tile_row(int self, int deg, const unsigned short *cell8,
                                         const int *eids, const double *vals, double *vs,
                                         int *cs) {
  int nx[G], pv[G];
  double vd[G], vn[G], vp[G];
  unsigned cw[4];
  if (G == 8) {
    const uint4 c8 = *reinterpret_cast<const uint4 *>(cell8);
    cw[0] = c8.x; cw[1] = c8.y; cw[2] = c8.z; cw[3] = c8.w;
  } else {
    const uint2 c4 = *reinterpret_cast<const uint2 *>(cell8);
    cw[0] = c4.x; cw[1] = c4.y; cw[2] = cw[3] = 0;
  }
#pragma unroll
  for (int i = 0; i < G; ++i) {
    nx[i] = 0x7fffffff - i;  // above every valid column, all different
    pv[i] = 0x7fffffff - G - i;
    vd[i] = vn[i] = vp[i] = 0.;
    if (i < deg) {
      const unsigned c = (cw[i >> 1] >> (16 * (i & 1))) & 0xffffu;
      const int el = (int)(c >> 2), k = (int)(c & 3u);
      nx[i] = eids[(k == 2 ? 0 : k + 1) * TC + el];
      pv[i] = eids[(k == 0 ? 2 : k - 1) * TC + el];
      V::unpack(k, vals + el, TC, vd[i], vn[i], vp[i]);
    }
  }
  bool ok = true;
  int slot_self = 0, sn[G], sp[G];
  unsigned mn = 0, mp = 0;
#pragma unroll
  for (int i = 0; i < G; ++i) {
    int rn = 0, rp = 0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      rn += nx[j] < nx[i];
      rp += pv[j] < pv[i];
    }
    if (i < deg) {
      mn |= 1u << rn;
      mp |= 1u << rp;
      ok = ok && nx[i] != self && pv[i] != self;
      slot_self += nx[i] < self;
    }
    sn[i] = rn + (self < nx[i]);
    sp[i] = rp + (self < pv[i]);
  }
  // closed, repeat-free fan: the nexts are all different, the prevs are all
  // different, and (checked below) they are the same set
  ok = ok && mn == (1u << deg) - 1u && mp == (1u << deg) - 1u;
#pragma unroll
  for (int i = 0; i < G; ++i)
    if (i < deg) { vs[sn[i]] = vd[i]; cs[sn[i]] = nx[i]; }
  cs[slot_self] = self;
  double diag = 0.;
  for (int q = 0; q <= deg; ++q)
    if (q != slot_self) diag += vs[q];
#pragma unroll
  for (int i = 0; i < G; ++i)
    if (i < deg) vs[sn[i]] = vn[i];
#pragma unroll
  for (int i = 0; i < G; ++i)
    if (i < deg) {
      ok = ok && cs[sp[i]] == pv[i];
      vs[sp[i]] += vp[i];
    }
  vs[slot_self] = diag;
  return ok;
}

// shared-memory layout (bytes); the image of the row phase reuses the table
constexpr int SM_T = 0;                        // int[HT]
constexpr int SM_JOF = SM_T + HT * 4;          // u16[HT]
constexpr int SM_XY = SM_JOF + HT * 2;         // double2[NMAX]
constexpr int SM_R1 = SM_XY + NMAX * 16;
constexpr int SM_VIMG = 0;                     // double[IMG]
constexpr int SM_CIMG = SM_VIMG + IMG * 8;     // int[IMG]
constexpr int SM_DIMG = SM_CIMG + IMG * 4;     // int[IMG]
static_assert(SM_DIMG + IMG * 4 <= SM_R1, "the image must fit the table region");
constexpr int SM_IDS = SM_R1;                  // int[NMAX]
constexpr int SM_CELLS = SM_IDS + NMAX * 4;    // u16[NMAX * 8]
constexpr int SM_ROWPOS = SM_CELLS + NMAX * 16;  // int[NMAX]
constexpr int SM_DEG = SM_ROWPOS + NMAX * 4;   // u8[NMAX]
constexpr int SM_STAT = SM_DEG + NMAX;         // u8[NMAX]
constexpr int SM_EIDS = SM_STAT + NMAX;        // int[3 * TC]
constexpr int SM_HEADS = SM_EIDS + 3 * TC * 4;   // int[HMAX]
constexpr int SM_SHEADS = SM_HEADS + HMAX * 4;   // int[HMAX]
constexpr int SM_RUN = SM_SHEADS + HMAX * 4;     // int[HMAX]
constexpr int SM_FBL = SM_RUN + HMAX * 4;        // int[NMAX]
constexpr int SM_CROW = SM_FBL + NMAX * 4;       // u16[NMAX]
constexpr int SM_MISC = SM_CROW + NMAX * 2;      // int[32]
constexpr int SM_VALS = SM_MISC + 32 * 4;        // double[NV * TC]
static_assert(SM_VALS % 16 == 0 && SM_CELLS % 16 == 0 && SM_EIDS % 16 == 0, "alignment");
template <class V> constexpr int tile_smem() { return SM_VALS + V::NV * TC * 8; }

// status of an owned node
enum : unsigned char { ST_COMPLETE = 0, ST_REGULAR = 1, ST_IRREGULAR = 2, ST_FOREIGN = 3 };

template <class V>
__global__ void __launch_bounds__(TT, V::NV <= 3 ? 2 : 1)
tile_kernel(TileArgs A, V values) {
  extern __shared__ __align__(16) unsigned char sm[];
  int *T = reinterpret_cast<int *>(sm + SM_T);
  unsigned short *jof = reinterpret_cast<unsigned short *>(sm + SM_JOF);
  double2 *xy = reinterpret_cast<double2 *>(sm + SM_XY);
  double *vimg = reinterpret_cast<double *>(sm + SM_VIMG);
  int *cimg = reinterpret_cast<int *>(sm + SM_CIMG);
  int *dimg = reinterpret_cast<int *>(sm + SM_DIMG);
  int *ids = reinterpret_cast<int *>(sm + SM_IDS);
  unsigned short *cells = reinterpret_cast<unsigned short *>(sm + SM_CELLS);
  int *rowpos = reinterpret_cast<int *>(sm + SM_ROWPOS);
  unsigned char *degs = sm + SM_DEG;
  unsigned char *stat = sm + SM_STAT;
  int *eids = reinterpret_cast<int *>(sm + SM_EIDS);
  int *heads = reinterpret_cast<int *>(sm + SM_HEADS);
  int *sheads = reinterpret_cast<int *>(sm + SM_SHEADS);
  int *runs = reinterpret_cast<int *>(sm + SM_RUN);
  int *fbl = reinterpret_cast<int *>(sm + SM_FBL);
  unsigned short *crow = reinterpret_cast<unsigned short *>(sm + SM_CROW);
  int *misc = reinterpret_cast<int *>(sm + SM_MISC);  // [0] #heads [1] n [2] #fb [3] fb base
  int *scratch = misc + 8;                            // 8 warp sums
  double *vals = reinterpret_cast<double *>(sm + SM_VALS);

  const int tid = threadIdx.x;
  // ---- phase 0: clear the table and the cells, load my 4 elements
  for (int i = tid; i < HT / 4; i += TT) reinterpret_cast<int4 *>(T)[i] = make_int4(-1, -1, -1, -1);
  for (int i = tid; i < NMAX; i += TT)
    reinterpret_cast<int4 *>(cells)[i] = make_int4(-1, -1, -1, -1);
  if (tid < 8) misc[tid] = 0;
  // thread tid owns elements tile + 4 tid .. + 3 (three 128-bit loads), or, for the
  // sources that stream [n_qp x M] arrays, tile + q TT + tid (consecutive lanes on
  // consecutive triangles)
  constexpr bool LM = V::lane_major;
  const int64_t tile0 = (int64_t)blockIdx.x * TC;
  const int64_t t0 = tile0 + (LM ? tid : 4 * tid);
  constexpr int ES = LM ? TT : 1;  // distance between my elements
  int nval = 0;
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) nval += t0 + (int64_t)q4 * ES < A.M;
  int id[12];
  unsigned code[4];
  if (!LM && nval == 4) {
    const int4 *e4 = reinterpret_cast<const int4 *>(A.elements + 3 * t0);
    const int4 p = __ldg(e4), q = __ldg(e4 + 1), r = __ldg(e4 + 2);
    id[0] = p.x; id[1] = p.y; id[2] = p.z; id[3] = p.w; id[4] = q.x; id[5] = q.y;
    id[6] = q.z; id[7] = q.w; id[8] = r.x; id[9] = r.y; id[10] = r.z; id[11] = r.w;
    const unsigned long long c = __ldg(reinterpret_cast<const unsigned long long *>(A.slotcode + t0));
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) code[q4] = (unsigned)(c >> (16 * q4)) & 0xffffu;
  } else {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      code[q4] = 0xfffu;
#pragma unroll
      for (int k = 0; k < 3; ++k) id[3 * q4 + k] = -1;
      if (q4 < nval) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
          id[3 * q4 + k] = __ldg(A.elements + 3 * (t0 + (int64_t)q4 * ES) + k);
        code[q4] = A.slotcode[t0 + (int64_t)q4 * ES];
      }
    }
  }
  // element q of thread tid is local element q * TT + tid (conflict-free columns)
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
    for (int k = 0; k < 3; ++k) eids[k * TC + q4 * TT + tid] = id[3 * q4 + k];
  unsigned own = 0, live = 0;  // references holding slot 0 of their node / valid ones
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const int q4 = i / 3, k = i - 3 * q4;
    if (q4 < nval) {
      live |= 1u << i;
      if (((code[q4] >> (4 * k)) & 15u) == 0u) own |= 1u << i;
    }
  }
  __syncthreads();

  // ---- phase A: the holder of slot 0 inserts its node (one writer per node;
  // different nodes that meet in a slot: write, barrier, read back, loser moves on)
  int hs[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) hs[i] = (int)tile_hash(id[i]);
  {
    unsigned pend = own;
    for (;;) {
      unsigned wrote = 0;
#pragma unroll
      for (int i = 0; i < 12; ++i)
        if ((pend >> i) & 1u) {
          if (T[hs[i]] == -1) { T[hs[i]] = id[i]; wrote |= 1u << i; }
          else hs[i] = (hs[i] + 1) & (HT - 1);
        }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 12; ++i)
        if ((wrote >> i) & 1u) {
          if (T[hs[i]] == id[i]) pend &= ~(1u << i);
          else hs[i] = (hs[i] + 1) & (HT - 1);
        }
      if (!__syncthreads_or(pend != 0u)) break;
    }
  }
  // the other references look their node up (not found: owned by another tile)
#pragma unroll
  for (int i = 0; i < 12; ++i)
    if (((live & ~own) >> i) & 1u) {
      unsigned s = (unsigned)hs[i];
      for (;;) {
        const int v = T[s];
        if (v == id[i]) { hs[i] = (int)s; break; }
        if (v == -1) { hs[i] = -1; break; }
        s = (s + 1) & (HT - 1);
      }
    }

  // ---- phase B: number the owned nodes in ascending id through their runs of
  // consecutive ids (head: id - 1 is not owned)
#pragma unroll
  for (int i = 0; i < 12; ++i)
    if ((own >> i) & 1u)
      if (!table_has(T, id[i] - 1)) {
        const int h = atomicAdd(&misc[0], 1);
        if (h < HMAX) heads[h] = id[i];
      }
  __syncthreads();
  const int H = misc[0];
  int n;
  if (H <= HMAX) {
    if (tid < H) {
      const int key = heads[tid];
      int r = 0;
      for (int q = 0; q < H; ++q) r += heads[q] < key;
      sheads[r] = key;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 12; ++i)
      if ((own >> i) & 1u)
        if (!table_has(T, id[i] + 1)) {  // last of its run: the run's length
          const int r = run_of(sheads, H, id[i]);
          runs[r] = id[i] - sheads[r] + 1;
        }
    __syncthreads();
    const int len = tid < H ? runs[tid] : 0;
    const int before = tile_scan(len, scratch, n);
    if (tid < H) runs[tid] = before;
    __syncthreads();
    if (n <= NMAX) {
#pragma unroll
      for (int i = 0; i < 12; ++i)
        if ((own >> i) & 1u) {
          const int r = run_of(sheads, H, id[i]);
          const int j = runs[r] + id[i] - sheads[r];
          jof[hs[i]] = (unsigned short)j;
          ids[j] = id[i];
        }
    }
  } else {  // too many runs to sort: number the nodes as they come
#pragma unroll
    for (int i = 0; i < 12; ++i)
      if ((own >> i) & 1u) {
        const int j = atomicAdd(&misc[1], 1);
        if (j < NMAX) { jof[hs[i]] = (unsigned short)j; ids[j] = id[i]; }
      }
    __syncthreads();
    n = misc[1];
  }
  __syncthreads();
  const bool degenerate = n > NMAX;  // block-uniform: every incidence takes the record path
  if (degenerate) n = 0;

  // ---- phase C: per node count / hash / coordinates; incidences park (element, k)
  for (int j = tid; j < n; j += TT) {
    const int node = ids[j];
    const unsigned long long a = __ldg(A.acc + node);
    const unsigned deg = (unsigned)a;
    degs[j] = (unsigned char)min(deg, 255u);
    stat[j] = ((unsigned)(a >> 32) == 0u && deg >= 1u && deg <= (unsigned)GROUP) ? ST_REGULAR
                                                                               : ST_IRREGULAR;
    if constexpr (V::needs_xy) xy[j] = __ldg(values.coords + node);
  }
  int jv[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    jv[i] = FOREIGN;
    if (((live >> i) & 1u) && hs[i] >= 0 && !degenerate) {
      const int q4 = i / 3, k = i - 3 * q4;
      jv[i] = jof[hs[i]];
      const unsigned pos = (code[q4] >> (4 * k)) & 15u;
      if (pos < (unsigned)GROUP) cells[jv[i] * GROUP + pos] = (unsigned short)(((q4 * TT + tid) << 2) | k);
    }
  }
  __syncthreads();

  // ---- phase D: a regular node whose cells are all filled is complete
  for (int j = tid; j < n; j += TT)
    if (stat[j] == ST_REGULAR) {
      const int deg = degs[j];
      bool full = true;
      for (int q = 0; q < deg; ++q) full = full && cells[j * GROUP + q] != 0xffffu;
      if (full) {
        stat[j] = ST_COMPLETE;
        rowpos[j] = __ldg(A.indptr + ids[j]);
      }
    }
  __syncthreads();
  int nc;
  {  // complete rows in ascending order: thread t compacts nodes 3t .. 3t+2
    int c = 0;
#pragma unroll
    for (int u = 0; u < NMAX / TT; ++u) {
      const int j = tid * (NMAX / TT) + u;
      c += j < n && stat[j] == ST_COMPLETE;
    }
    int at = tile_scan(c, scratch, nc);
#pragma unroll
    for (int u = 0; u < NMAX / TT; ++u) {
      const int j = tid * (NMAX / TT) + u;
      if (j < n && stat[j] == ST_COMPLETE) crow[at++] = (unsigned short)j;
    }
  }

  // ---- phase E: values of my elements; incidences of incomplete nodes become records
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    if (q4 >= nval) continue;
    const int el = q4 * TT + tid;
    const int e[3] = {id[3 * q4], id[3 * q4 + 1], id[3 * q4 + 2]};
    double2 z[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      z[k] = make_double2(0., 0.);
      if constexpr (V::needs_xy) {
        const int j = jv[3 * q4 + k];
        z[k] = j != FOREIGN ? xy[j] : __ldg(values.coords + e[k]);
      }
    }
    values.pack(t0 + (int64_t)q4 * ES, e, z[0], z[1], z[2], vals + el, TC);
    unsigned char st[3];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int j = jv[3 * q4 + k];
      st[k] = j != FOREIGN ? stat[j] : ST_FOREIGN;
      any = any || st[k] != ST_COMPLETE;
    }
    if (any) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (st[k] == ST_COMPLETE) continue;
        double vd, vn, vp;
        V::unpack(k, vals + el, TC, vd, vn, vp);
        const unsigned pos = (code[q4] >> (4 * k)) & 15u;
        const int nx = e[k == 2 ? 0 : k + 1], pvk = e[k == 0 ? 2 : k - 1] | (k << NODE_BITS);
        if (pos < (unsigned)GROUP) {
          Rec r;
          r.nx = nx; r.pvk = pvk; r.vd = vd; r.vn = vn; r.vp = vp;
          store_rec(A.rec + (int64_t)e[k] * GROUP + pos, r);
        } else {
          spill_record(A.spill, A.flags, e[k], nx, pvk, vd, vn, vp,
                       (int)(((t0 + (int64_t)q4 * ES) << 2) | k));
        }
        if (pos == 0u) {  // the holder of slot 0 lists a regular row for the record fill
          if (st[k] == ST_REGULAR) {
            fbl[atomicAdd(&misc[2], 1)] = e[k];
          } else if (st[k] == ST_FOREIGN) {  // (degenerate tile: nothing is owned)
            const unsigned long long a = __ldg(A.acc + e[k]);
            if ((unsigned)(a >> 32) == 0u && (unsigned)a <= (unsigned)GROUP)
              A.fb_rows[atomicAdd(A.fb_count, 1)] = e[k];
          }
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0 && misc[2] > 0) misc[3] = atomicAdd(A.fb_count, misc[2]);
  __syncthreads();
  for (int i = tid; i < misc[2]; i += TT) A.fb_rows[misc[3] + i] = fbl[i];

  // ---- phase F: the complete rows, TT at a time, staged and streamed out
  bool ok = true;
  for (int b0 = 0; b0 < nc; b0 += TT) {
    __syncthreads();  // the image replaces the table (first pass) / the previous batch
    const int r = b0 + tid;
    const bool active = r < nc;
    const int j = active ? crow[r] : 0;
    const int deg = active ? degs[j] : 0;
    int total;
    const int base = tile_scan(active ? deg + 1 : 0, scratch, total);
    if (active) {
      const int self = ids[j];
      if (__any_sync(__activemask(), deg > 4))
        ok = tile_row<8, V>(self, deg, cells + j * GROUP, eids, vals, vimg + base, cimg + base) && ok;
      else
        ok = tile_row<4, V>(self, deg, cells + j * GROUP, eids, vals, vimg + base, cimg + base) && ok;
      const int delta = rowpos[j] - base;
      for (int q = 0; q <= deg; ++q) dimg[base + q] = delta;
    }
    __syncthreads();
    for (int x = tid; x < total; x += TT) {
      const int64_t o = (int64_t)x + dimg[x];
      if (A.indices) A.indices[o] = cimg[x];
      A.data[o] = vimg[x];
    }
  }
  if (!ok) atomicOr(A.flags, FLAG_RETRY);
}

template <class V>
static int launch_tile(const TileArgs &args, V values, cudaStream_t s) {
  const int smem = tile_smem<V>();
  P1_CUDA(cudaFuncSetAttribute(tile_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const unsigned grid = grid_for(args.M, TC);
  tile_kernel<V><<<grid, TT, smem, s>>>(args, values);
  return P1_OK;
}

}  // namespace p1

using namespace p1;

// ------------------------------------------------------------ plan (symbolic)
int p1::plan_build_tiled(p1_ctx *ctx, const int32_t *elements, int64_t M, int64_t n_nodes,
                         int op, const double *coords, const double *local,
                         const p1_quad_args *qa, cudaStream_t s, p1_plan **out) {
  const int N = (int)n_nodes;
  p1_plan *p = new p1_plan();
  memset(p, 0, sizeof(*p));
  p->ctx = ctx; p->stream = s; p->elements = elements; p->M = M; p->N = n_nodes;
  p->row_begin = 0; p->row_end = n_nodes;
  p->own = Own{0, N, 0, 1, 0, 0};
  p->n_own = N;
  p->tiled = true;
  p->val_op = op;
  p->v_coords = coords; p->v_local = local;
  if (qa) p->v_qa = *qa;
  *out = nullptr;
  int rc = P1_OK;
  int *rowlen = nullptr;
  auto fail = [&](int code) {
    pool_free(ctx, rowlen, s);
    p1_plan_destroy(p);
    return code;
  };
  if (cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s) != cudaSuccess ||
      cudaMemsetAsync(ctx->d_flags + 1, 0xff, sizeof(int), s) != cudaSuccess) {
    set_error("p1_plan_create: memset: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(P1_ECUDA);
  }
  if ((rc = pool_alloc(ctx, &p->acc, (size_t)N + 1, s)) ||
      (rc = pool_alloc(ctx, &p->slotcode, (size_t)M + 8, s)) ||
      (rc = pool_alloc(ctx, &p->indptr, (size_t)N + 1, s)) ||
      (rc = pool_alloc(ctx, &rowlen, (size_t)N + 1, s)) ||
      (rc = pool_alloc(ctx, &p->slow_rows, (size_t)N + 1, s)) ||
      (rc = pool_alloc(ctx, &p->big_rows, (size_t)((3 * M) / (DEG_BIG + 1)) + 1, s)) ||
      (rc = pool_alloc(ctx, &p->fb_rows, (size_t)N + 1, s)) ||
      (rc = pool_alloc(ctx, &p->tctr, (size_t)8, s)) ||
      (rc = pool_alloc(ctx, &p->rec, (size_t)N * GROUP, s)))
    return fail(rc);
  if (op == P1_OP_GENERAL || op == P1_OP_WEIGHTED) {  // the rule travels to the device once
    if ((rc = pool_alloc(ctx, &p->d_weights, (size_t)qa->n_qp, s)) ||
        (rc = pool_alloc(ctx, &p->d_points, (size_t)2 * qa->n_qp, s)))
      return fail(rc);
    cudaMemcpyAsync(p->d_weights, qa->weights_host, qa->n_qp * sizeof(double),
                    cudaMemcpyHostToDevice, s);
    if (qa->points_host)
      cudaMemcpyAsync(p->d_points, qa->points_host, 2 * qa->n_qp * sizeof(double),
                      cudaMemcpyHostToDevice, s);
  }
  cudaMemsetAsync(p->acc, 0, ((size_t)N + 1) * sizeof(unsigned long long), s);
  if (M > 0)
    P1_TIMED(ctx, "count", s,
             (count_kernel<<<min(grid_for(M, CT, 4), (unsigned)(ctx->sm_count * 32)), CT, 0, s>>>(
                 elements, M, N, p->acc, p->slotcode, ctx->d_flags)));
  if (N > 0)
    P1_TIMED(ctx, "split", s,
             (split_tiled_kernel<<<grid_for(N, 256), 256, 0, s>>>(
                 N, p->acc, rowlen, p->slow_rows, p->big_rows, ctx->d_flags + 2)));
  if ((rc = exclusive_scan_i32(ctx, rowlen, p->indptr, N, true, s))) return fail(rc);
  int h_nnz = 0;
  cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(&h_nnz, p->indptr + N, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("p1_plan_create: %s", cudaGetErrorString(e));
    return fail(P1_ECUDA);
  }
  pool_free(ctx, rowlen, s);
  rowlen = nullptr;
  if (ctx->h_flags[0] & FLAG_INDEX_RANGE) {
    set_error("p1_plan_create: element index out of range [0, %lld)", (long long)n_nodes);
    return fail(P1_EINDEX);
  }
  if (h_nnz < 0) {
    set_error("p1_plan_create: more than 2^31 - 1 stored entries");
    return fail(P1_ERANGE);
  }
  p->max_index = ctx->h_flags[1];
  p->n_big = ctx->h_flags[2];
  p->n_slow = ctx->h_flags[3];
  p->n_gt8 = ctx->h_flags[4];
  p->n_spill = ctx->h_flags[7];
  p->nnz = h_nnz;
  if (p->n_big > 0) return fail(P1_ERETRY);  // pathological fans: the record plan handles them
  // rows with more than 8 records: spill list now, compact storage after the tiles
  p->spill.cap = p->n_spill > 0 ? p->n_spill : 1;
  if ((rc = pool_alloc(ctx, &p->spill.rec, (size_t)p->spill.cap, s)) ||
      (rc = pool_alloc(ctx, &p->spill.row, (size_t)p->spill.cap, s)) ||
      (rc = pool_alloc(ctx, &p->d_spill, (size_t)1, s)))
    return fail(rc);
  cudaMemcpyAsync(p->d_spill, &p->spill, sizeof(SpillDesc), cudaMemcpyHostToDevice, s);
  if (p->n_gt8 > 0) {
    p->n_ovf = (int64_t)GROUP * p->n_gt8 + p->n_spill;
    if ((rc = pool_alloc(ctx, &p->ovf_off, (size_t)N + 1, s)) ||
        (rc = pool_alloc(ctx, &p->ovf_cursor, (size_t)N + 1, s)) ||
        (rc = pool_alloc(ctx, &p->ovf_rec, (size_t)p->n_ovf, s)))
      return fail(rc);
  }
  *out = p;
  return P1_OK;
}

// ------------------------------------------------------------ assemble (numeric)
int p1::assemble_tiled(p1_plan *plan, int op, int32_t *indptr, int32_t *indices, double *data,
                       void *stream) {
  p1_ctx *ctx = plan->ctx;
  DeviceGuard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  if (op != P1_OP_STORED && op != plan->val_op) {
    set_error("p1_assemble_csr: this plan was created for one operator (p1_plan_create_for); "
              "create it with P1_PLAN_KEEP_ELEMENTS to assemble others");
    return P1_EINVAL;
  }
  const int N = (int)plan->N;
  if (indptr)
    P1_CUDA(cudaMemcpyAsync(indptr, plan->indptr, ((size_t)N + 1) * sizeof(int),
                            cudaMemcpyDeviceToDevice, s));
  P1_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), s));
  P1_CUDA(cudaMemsetAsync(plan->tctr, 0, 8 * sizeof(int), s));
  TileArgs args{plan->elements, plan->M, N, plan->acc, plan->slotcode, plan->indptr, plan->rec,
                plan->d_spill, indices, data, plan->fb_rows, plan->tctr, ctx->d_flags};
  const double2 *xy = (const double2 *)plan->v_coords;
  const p1_quad_args &qa = plan->v_qa;
  if (plan->M > 0) {
    prof_begin(ctx, "tile", s);
    int rc = P1_OK;
    switch (plan->val_op) {
      case P1_OP_STIFFNESS: rc = launch_tile(args, StiffnessValues{xy}, s); break;
      case P1_OP_MASS: rc = launch_tile(args, MassValues{xy}, s); break;
      case P1_OP_LOCAL: rc = launch_tile(args, LocalValues{plan->v_local}, s); break;
      case P1_OP_GENERAL:
        rc = launch_tile(args, GeneralValues{xy, qa.a11q, qa.a12q, qa.a21q, qa.a22q,
                                             plan->d_weights, qa.n_qp, plan->M}, s);
        break;
      case P1_OP_WEIGHTED:
        rc = launch_tile(args, WeightedValues{xy, qa.phiq, plan->d_weights,
                                              (const double2 *)plan->d_points, qa.n_qp, plan->M}, s);
        break;
      default:
        set_error("p1_assemble_csr: tiled plan without an operator");
        rc = P1_EINVAL;
    }
    prof_end(ctx, s);
    if (rc) return rc;
  }
  if (plan->n_gt8 > 0) overflow_collect(plan, plan->ovf_cursor, plan->tctr + 1, s);
  launch_fill_fallback(plan, indices, data, s);
  P1_CUDA(cudaGetLastError());
  P1_CUDA(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  P1_CUDA(cudaStreamSynchronize(s));
  if (ctx->h_flags[0] & FLAG_RETRY) {
    set_error("p1_assemble_csr: the mesh contradicts the plan's closed-fan shortcut "
              "(non-manifold vertex or degenerate element); rebuild the plan with "
              "P1_PLAN_EXACT");
    return P1_ERETRY;
  }
  return P1_OK;
}
