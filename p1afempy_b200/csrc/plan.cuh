// plan.cuh -- the p1_plan object, the 32-byte incidence record and the 8-lane
// row analysis shared by the symbolic (row length) and numeric (fill) kernels.
#pragma once
#include "common.cuh"

namespace p1 {
// One record per (node, incident element): exactly one 32-byte sector, written
// with one 256-bit store by the scatter and read with one 256-bit load per lane.
//   nx  : next vertex of the element after this node
//   pvk : previous vertex | local index k of this node in the element << 29
//   vd, vn, vp : entries (k,k), (k,k+1), (k,k+2) of the element's local matrix
struct __align__(32) Rec {
  int nx, pvk;
  double vd, vn, vp;
};

// Spill list for the records of rows with more than 8 records (see spill_record).
struct SpillDesc {
  Rec *rec;
  int *row, *elem;  // owning row; element code (null unless the plan keeps elements)
  int cap;
};

// Which rows a plan owns, and the local <-> global row numbering.
//   contiguous : rows [rb, re)                                   (world == 1)
//   interleaved: of the chunks of 2^cshift rows inside [rb, re), those with
//                chunk index % world == rank; local rows are the owned chunks
//                in ascending order.  Every rank then holds the same mix of
//                old (high-valence) and new nodes: balanced in every respect.
struct Own {
  int rb, re, cshift, world, rank;
  int wshift;  // log2(world) when world is a power of two, else -1
  // p1_assemble_cold stores the rows' 8 slots SLOT-major, slot i of row a at [i * lay + a]
  // (lay = number of owned rows): the fill's loads of slot i for 32 consecutive rows are
  // then contiguous (8 L1 wavefronts per 256-bit load instruction instead of 32).
  // 0: row-major, [a * 8 + i] (plans).
  int lay;
  __host__ __device__ __forceinline__ int64_t pos(int a, unsigned slot) const {
    const int rs = lay ? 1 : 8, ss = lay ? lay : 1;
    return (int64_t)a * rs + (int64_t)slot * ss;
  }
  __host__ __device__ __forceinline__ bool owns(int e, int &a) const {
    const unsigned x = (unsigned)(e - rb);
    if (x >= (unsigned)(re - rb)) return false;
    if (world == 1) { a = (int)x; return true; }
    const unsigned q = x >> cshift;
    unsigned owner, mine;  // the scatter tests 3 vertices of every element: keep it cheap
    if (wshift >= 0) { owner = q & ((1u << wshift) - 1u); mine = q >> wshift; }
    else { owner = q % (unsigned)world; mine = q / (unsigned)world; }
    if ((int)owner != rank) return false;
    a = (int)((mine << cshift) | (x & ((1u << cshift) - 1u)));
    return true;
  }
  __host__ __device__ __forceinline__ int global(int a) const {
    if (world == 1) return a + rb;
    const unsigned q = (unsigned)a >> cshift;
    return rb + (int)(((q * (unsigned)world + (unsigned)rank) << cshift) |
                      ((unsigned)a & ((1u << cshift) - 1u)));
  }
  // number of owned rows
  __host__ int64_t count() const {
    const int64_t n = (int64_t)re - rb;
    if (world == 1) return n;
    const int64_t c = (int64_t)1 << cshift;
    const int64_t full = n / c, tail = n % c;
    int64_t mine = (full / world) * c + ((full % world) > rank ? c : 0);
    if (tail && (full % world) == rank) mine += tail;
    return mine;
  }
};

}  // namespace p1

struct p1_plan {
  p1_ctx *ctx;
  cudaStream_t stream;      // the stream the plan was built on (frees are ordered on it)
  const int32_t *elements;  // borrowed, M x 3 (must outlive the plan)
  int64_t M, N;
  int64_t row_begin, row_end;  // row range the ownership is defined on
  p1::Own own;                 // which of those rows this plan owns
  int64_t n_own;               // number of owned (= local) rows
  int64_t nnz;
  int max_index;
  // Records: row a owns the 8 slots rec[8a .. 8a+7] (256 B); its slot number comes
  // from the same atomic that counts it, so there is no counting pass and no
  // offset array.  Rows with more than 8 records (none on an NVB mesh) keep ALL
  // their records in the compact overflow area instead: ovf_rec[ovf_off[a] ..].
  unsigned long long *acc;  // n_own: low 32 bits = #records, high 32 bits = sum over the
                       // records of hash(next) - hash(prev): 0 for a closed fan
  p1::Rec *rec;        // 8 * n_own
  int *rec_elem;       // 8 * n_own: (element << 2) | k of the record
                       // (P1_PLAN_KEEP_ELEMENTS) or null
  int4 *topo;          // P1_PLAN_TOPOLOGY: 8 * n_own slots {next, prev | k<<29, (element<<2)|k, 0}
                       // INSTEAD of rec / rec_elem / indptr (edge numbering, eta_R, load
                       // vectors need no values and no row pointer)
  int2 *slim;          // multi-GPU stiffness / mass plans: 8 * n_own slots {next, prev | k<<29}
                       // INSTEAD of rec; the fill gathers the coordinates and recomputes
                       // the entries (plan_build explains when that pays)
  const double2 *coords;  // borrowed with a slim plan (must outlive it)
  int *slot_of;        // with topo: 3M, slot index 8a+i of half-edge 3T+k (or null)
  int4 *tspill;        // with topo: the records that found their row full, {next, prev|k,
  int n_tspill;        // element code, ROW}; at most TSPILL_CAP, else the plan is refused.
                       // Their slot index (slot_of) is 8*N + position in this list.
  int *ovf_off;        // n_own+1 (only when some row overflows) or null
  p1::Rec *ovf_rec;    // n_ovf
  int *ovf_elem;       // n_ovf or null
  int64_t n_ovf;
  int val_op;          // which operator's values the records hold (0 = none)
  bool exact;          // row lengths from the all-pairs kernel (no hash shortcut)
  int *indptr;         // n_own+1
  int *slow_rows;      // rows the fast fill does not take: open fans, more than 8
                       // records, repeated neighbours (non-manifold / degenerate)
  int n_slow;
  int *big_rows;       // rows with more than DEG_BIG records
  int n_big;
  unsigned char *big_first;  // scratch flags for big rows (2*n_ovf+n_own) or null
  int *twin;           // lazily built: 3M, twin half-edge id T'*3+k' or -1
  bool twin_partial;   // only the entries of half-edges I -> J with J < I are valid
  int64_t n_open;      // half-edges without twin (valid once twin is built)
  // Tiled plans (tiled.cu): the count pass stored the slot of every incidence, the
  // tile kernel of p1_assemble_csr completes tile-interior rows in shared memory and
  // only the incidences of the other rows travel through `rec`.
  bool tiled;
  unsigned short *slotcode;  // M: slot of vertex k in bits 4k..4k+3 (15 = beyond 8)
  int *fb_rows;        // rows the tiles left to the record path (filled per assemble)
  int *tctr;           // 8 device counters: [0] #fb_rows, [1] next free overflow range
  int *ovf_cursor;     // n_own (only when some row has more than 8 records)
  int n_gt8, n_spill;  // rows with more than 8 records / their records beyond 8
  p1::SpillDesc spill; // spill list (capacity n_spill) and its device copy
  p1::SpillDesc *d_spill;
  const double *v_coords, *v_local;  // borrowed value sources of the stored operator
  p1_quad_args v_qa;                 // (device arrays borrowed)
  double *d_weights, *d_points;      // the rule on the device (quadrature operators)
};

namespace p1 {

// 32-bit mixer (lowbias32) for the closed-fan test
__host__ __device__ __forceinline__ unsigned mix32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

constexpr int GROUP = 8;      // lanes per row in the fast kernels = max records
constexpr int TSPILL_CAP = 1024;  // topology-only plans: records beyond 8 per row, in total
constexpr int DEG_BIG = 40;   // rows with more records take the block-per-row path
// node ids must leave the top 3 bits of an int free (k is packed into pvk)
constexpr int NODE_BITS = 29;
constexpr int NODE_MASK = (1 << NODE_BITS) - 1;

__device__ __forceinline__ int next_local(int k) { return k == 2 ? 0 : k + 1; }
__device__ __forceinline__ int prev_local(int k) { return k == 0 ? 2 : k - 1; }
__device__ __forceinline__ int rec_next(const Rec &r) { return r.nx; }
__device__ __forceinline__ int rec_prev(const Rec &r) { return r.pvk & NODE_MASK; }
__device__ __forceinline__ int rec_k(const Rec &r) { return (int)((unsigned)r.pvk >> NODE_BITS); }

// Where the records of a row live (see p1_plan).
struct Rows {
  const unsigned long long *acc;
  Rec *rec;
  int *rec_elem;
  const int *ovf_off;
  Rec *ovf_rec;
  int *ovf_elem;
  const int4 *topo;  // topology-only plans: everything in one 16-byte slot
  const int4 *tspill;  // ... and the few records beyond 8 per row (.w = row)
  int n_tspill;
  const int2 *slim;  // slim plans: {next, prev | k << 29}, values recomputed by the fill
  const int4 *half;  // p1_assemble_cold, mass matrix: {next, prev | k << 29, v}, local row (2v, v, v)
  __device__ __forceinline__ int deg(int a) const { return (int)(unsigned)acc[a]; }
  // (next, prev | k << 29) of record i of row a, whatever the storage
  __device__ __forceinline__ int2 edge(int a, int deg, int i) const {
    if (slim) return slim[(int64_t)a * 8 + i];
    if (half) {
      const int4 h = half[(int64_t)a * 8 + i];
      return make_int2(h.x, h.y);
    }
    const Rec *r = ptr(a, deg) + i;
    return make_int2(r->nx, r->pvk);
  }
  __device__ __forceinline__ Rec *ptr(int a, int deg) const {
    return deg <= 8 ? rec + (int64_t)a * 8 : ovf_rec + ovf_off[a];
  }
  __device__ __forceinline__ int *elem(int a, int deg) const {
    return deg <= 8 ? rec_elem + (int64_t)a * 8 : ovf_elem + ovf_off[a];
  }
};
inline Rows rows_of(const p1_plan *p) {
  return Rows{p->acc, p->rec, p->rec_elem, p->ovf_off, p->ovf_rec, p->ovf_elem, p->topo, p->tspill,
              p->n_tspill, p->slim};
}

// sm_100a has 256-bit global loads/stores (LDG.E.256 / STG.E.256)
// (.cs: a record is written once and read once, by another kernel; evict-first keeps a
// little more of the counters and coordinates in L2 -- scatter 2.20 -> 2.19 ms at 64M)
#ifndef P1_REC_ST
#define P1_REC_ST "st.global.cs.v4.b64"
#endif
#ifndef P1_REC_LD
#define P1_REC_LD "ld.global.v4.b64"
#endif
__device__ __forceinline__ void store_rec(Rec *p, const Rec &r) {
  asm volatile(P1_REC_ST " [%0], {%1,%2,%3,%4};"
               :: "l"(p), "l"(((long long)(unsigned)r.pvk << 32) | (unsigned)r.nx),
                  "l"(__double_as_longlong(r.vd)), "l"(__double_as_longlong(r.vn)),
                  "l"(__double_as_longlong(r.vp))
               : "memory");
}
__device__ __forceinline__ Rec load_rec(const Rec *p) {
  long long a, b, c, d;
  asm volatile(P1_REC_LD " {%0,%1,%2,%3}, [%4];"
               : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  Rec r;
  r.nx = (int)(unsigned)(a & 0xffffffffll);
  r.pvk = (int)(unsigned)((unsigned long long)a >> 32);
  r.vd = __longlong_as_double(b);
  r.vn = __longlong_as_double(c);
  r.vp = __longlong_as_double(d);
  return r;
}

// Records that do not fit the 8 slots of their row (a node with more than 8
// elements) go to a spill list; the descriptor lives in device memory and is read
// only by the rare thread that spills (plan.cu: scatter, tiled.cu: tile kernel).
// (called from a cold block behind ONE branch per triangle / pair: inside the store
// loop it cost 0.14 ms at 64M without ever running)
__device__ __forceinline__ void spill_record(const SpillDesc *sp, int *flags, int a, int nx,
                                             int pvk, double v0, double v1, double v2, int code) {
  const SpillDesc d = *sp;
  const int idx = atomicAdd(flags + 5, 1);  // counts every spilled record, kept or not
  if (idx >= d.cap) return;
  Rec r;
  r.nx = nx; r.pvk = pvk;
  r.vd = v0; r.vn = v1; r.vp = v2;
  store_rec(d.rec + idx, r);
  d.row[idx] = a;
  if (d.elem) d.elem[idx] = code;
}


// What one lane of an 8-lane group knows about "its" record and its row after
// the all-pairs comparison.  A row is `simple` when it has at most 8 records,
// all `next` distinct, all `prev` distinct and none equal to the row itself --
// every row of an edge-manifold mesh of valence <= 8.
struct RowLane {
  bool valid;     // this lane holds a record
  bool simple;    // group-uniform
  int nx, pv, k;
  double vd, vn, vp;
  int rk;         // rank of nx among the row's nexts = canonical record order
  int src;        // lane whose next == my prev (the element before me), or -1
  int mate;       // lane whose prev == my next (the element after me), or -1
  bool extra;     // my prev is a column no `next` provides (open fan end)
  int slot_nx, slot_pv, slot_self;  // positions in the sorted row
  int len;        // number of distinct columns
};

// `xch` is one int2 slot per thread of the block (shared memory); lanes of a
// group exchange (next, prev) through it.  (Shuffles would do, but 16 of them
// per lane saturate the ADU pipe -- measured 93 % -- while broadcast LDS is
// one wavefront per warp.)
__device__ __forceinline__ RowLane analyse_row(int self, const Rec *row, int deg,
                                               unsigned gmask, int lane8, int shift,
                                               int2 *xch) {
  RowLane o;
  o.valid = lane8 < deg && deg <= GROUP;
  Rec r;
  r.nx = -1; r.pvk = 0; r.vd = r.vn = r.vp = 0.;
  if (o.valid) r = load_rec(row + lane8);
  o.nx = r.nx;
  o.pv = o.valid ? rec_prev(r) : -2;
  o.k = rec_k(r);
  o.vd = r.vd; o.vn = r.vn; o.vp = r.vp;
  int2 *grp = xch + (threadIdx.x & ~(GROUP - 1));
  grp[lane8] = make_int2(o.nx, o.pv);
  __syncwarp(gmask);
  bool ok = deg <= GROUP;
  int rk = 0, src = -1, mate = -1, below_pv = 0, below_self = 0;
#pragma unroll
  for (int j = 0; j < GROUP; ++j) {
    const int2 q = grp[j];
    if (j < deg) {
      if (q.x == self || q.y == self) ok = false;
      if (o.valid) {
        if (j != lane8 && (q.x == o.nx || q.y == o.pv)) ok = false;
        rk += q.x < o.nx;
        below_pv += q.x < o.pv;
        if (q.x == o.pv) src = j;
        if (q.y == o.nx) mate = j;
      }
      below_self += q.x < self;
    }
  }
  o.simple = (__ballot_sync(gmask, !ok) & gmask) == 0;
  o.rk = rk; o.src = src; o.mate = mate;
  o.extra = o.valid && src < 0;
  const unsigned emask = (__ballot_sync(gmask, o.extra) & gmask) >> shift;
  o.slot_nx = (self < o.nx) + rk;
  o.slot_pv = (self < o.pv) + below_pv;
  o.slot_self = below_self;
  if (emask) {  // open fan: the extra prev columns shift everything above them
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      const int qp = grp[j].y;
      if ((emask >> j) & 1) {
        o.slot_nx += qp < o.nx;
        o.slot_pv += qp < o.pv;
        o.slot_self += qp < self;
      }
    }
  }
  o.len = 1 + deg + __popc(emask);
  return o;
}

// All records of row a of a topology-only plan: f(where, slot) with where = 8a+i for
// the slots, 8*n_rows + j for spilled record j.  (Rows with more than 8 records are
// few -- TSPILL_CAP bounds the list -- so scanning it is fine.)
template <class F>
__device__ __forceinline__ void topo_row_records(const Rows &rows, int a, int n_rows, F f) {
  const int deg = rows.deg(a);
  const int4 *slot = rows.topo + (int64_t)a * GROUP;
  for (int i = 0; i < deg && i < GROUP; ++i) f(a * GROUP + i, slot[i]);
  if (deg > GROUP)
    for (int j = 0; j < rows.n_tspill; ++j) {
      const int4 r = rows.tspill[j];
      if (r.w == a) f(n_rows * GROUP + j, r);
    }
}

int ensure_twins(p1_plan *plan, cudaStream_t s, bool backward_only = false);
// tiled.cu: the one-GPU cold path (count pass, tile kernel); P1_ERETRY from the
// plan builder means "use the record plan for this mesh"
int plan_build_tiled(p1_ctx *ctx, const int32_t *elements, int64_t M, int64_t n_nodes, int op,
                     const double *coords, const double *local, const p1_quad_args *qa,
                     cudaStream_t s, p1_plan **out);
int assemble_tiled(p1_plan *plan, int op, int32_t *indptr, int32_t *indices, double *data,
                   void *stream);
// plan.cu: rows with more than 8 records move to compact storage (8 slot records +
// the spilled ones); cursor: n_own scratch ints, next_free: a zeroed device counter
void overflow_collect(p1_plan *plan, int *cursor, int *next_free, cudaStream_t s);
// assemble.cu: the rows a tiled plan left to the record path (plan->fb_rows, counted on
// the device in plan->tctr[0]) and the slow rows
void launch_fill_fallback(p1_plan *plan, int *indices, double *data, cudaStream_t s);
struct Rows;
// assemble.cu: the fill kernel of p1_assemble_cold (the status words are read on the
// device, open fans are filled by the row's own thread; slim_op != 0: 8-byte slots,
// values recomputed)
void launch_fill_cold(p1_ctx *ctx, const Own &own, int n_own, const Rows &rows, int slim_op,
                      const double2 *coords, const int *indptr, int *indices, double *data,
                      int *status, cudaStream_t s);
// error unless the plan was created with P1_PLAN_KEEP_ELEMENTS
int require_elements(const p1_plan *plan, const char *who);

// ---- value sources for the scatter / refresh kernels (assemble.cu) ----------
// rows(t, e, z0, z1, z2, v): v[k][0..2] = entries (k,k), (k,k+1), (k,k+2) of
// element t; z0..z2 are the vertex coordinates (supplied by the caller, who can
// share them between neighbouring triangles) when needs_xy
// Tiled assembly (tiled.cu) parks what an element contributes in shared memory:
// NV doubles per element, written by pack() (strided by `ld`), from which
// unpack() rebuilds the entries (k,k), (k,k+1), (k,k+2) of vertex k with the
// expressions of rows(), bit for bit.  Sources without a compact form park all
// nine entries.
#define P1_GENERIC_PACK                                                               \
  static constexpr int NV = 9;                                                        \
  static constexpr bool lane_major = true;                                            \
  __device__ __forceinline__ void pack(int64_t t, const int *e, double2 z0, double2 z1, \
                                       double2 z2, double *out, int ld) const {       \
    double v[3][3];                                                                   \
    rows(t, e, z0, z1, z2, v);                                                        \
    _Pragma("unroll") for (int k = 0; k < 3; ++k)                                     \
      _Pragma("unroll") for (int c = 0; c < 3; ++c) out[(3 * k + c) * ld] = v[k][c]; \
  }                                                                                   \
  __device__ static __forceinline__ void unpack(int k, const double *in, int ld,      \
                                                double &vd, double &vn, double &vp) { \
    vd = in[(3 * k) * ld]; vn = in[(3 * k + 1) * ld]; vp = in[(3 * k + 2) * ld];      \
  }

struct NoValues {
  static constexpr bool needs_xy = false;
  P1_GENERIC_PACK
  __device__ __forceinline__ void rows(int64_t, const int *, double2, double2, double2,
                                       double (*v)[3]) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k][0] = v[k][1] = v[k][2] = 0.;
  }
};
struct StiffnessValues {  // solvers.py:31-45
  static constexpr bool needs_xy = true;
  const double2 *__restrict__ coords;
  __device__ __forceinline__ void rows(int64_t, const int *, double2 z0, double2 z1,
                                       double2 z2, double (*v)[3]) const {
    const double d21x = z1.x - z0.x, d21y = z1.y - z0.y;
    const double d31x = z2.x - z0.x, d31y = z2.y - z0.y;
    const double area4 = 4. * (0.5 * (d21x * d31y - d21y * d31x));
    const double a = (d21x * d31x + d21y * d31y) / area4;
    const double b = (d31x * d31x + d31y * d31y) / area4;
    const double c = (d21x * d21x + d21y * d21y) / area4;
    v[0][0] = -2. * a + b + c; v[0][1] = a - b; v[0][2] = a - c;
    v[1][0] = b;               v[1][1] = -a;    v[1][2] = a - b;
    v[2][0] = c;               v[2][1] = a - c; v[2][2] = -a;
  }
  // compact form: a, b, c
  static constexpr int NV = 3;
  static constexpr bool lane_major = false;
  __device__ __forceinline__ void pack(int64_t, const int *, double2 z0, double2 z1,
                                       double2 z2, double *out, int ld) const {
    const double d21x = z1.x - z0.x, d21y = z1.y - z0.y;
    const double d31x = z2.x - z0.x, d31y = z2.y - z0.y;
    const double area4 = 4. * (0.5 * (d21x * d31y - d21y * d31x));
    out[0] = (d21x * d31x + d21y * d31y) / area4;
    out[ld] = (d31x * d31x + d31y * d31y) / area4;
    out[2 * ld] = (d21x * d21x + d21y * d21y) / area4;
  }
  __device__ static __forceinline__ void unpack(int k, const double *in, int ld, double &vd,
                                                double &vn, double &vp) {
    const double a = in[0], b = in[ld], c = in[2 * ld];
    if (k == 0) { vd = -2. * a + b + c; vn = a - b; vp = a - c; }
    else if (k == 1) { vd = b; vn = -a; vp = a - b; }
    else { vd = c; vn = a - c; vp = -a; }
  }
};
struct MassValues {  // solvers.py:176-181
  static constexpr bool needs_xy = true;
  const double2 *__restrict__ coords;
  __device__ __forceinline__ void rows(int64_t, const int *, double2 z0, double2 z1,
                                       double2 z2, double (*v)[3]) const {
    const double d21x = z1.x - z0.x, d21y = z1.y - z0.y;
    const double d31x = z2.x - z0.x, d31y = z2.y - z0.y;
    const double area = 0.5 * (d21x * d31y - d21y * d31x);
    const double on = area / 6., off = area / 12.;
#pragma unroll
    for (int k = 0; k < 3; ++k) { v[k][0] = on; v[k][1] = off; v[k][2] = off; }
  }
  // compact form: area/6, area/12
  static constexpr int NV = 2;
  static constexpr bool lane_major = false;
  __device__ __forceinline__ void pack(int64_t, const int *, double2 z0, double2 z1,
                                       double2 z2, double *out, int ld) const {
    const double d21x = z1.x - z0.x, d21y = z1.y - z0.y;
    const double d31x = z2.x - z0.x, d31y = z2.y - z0.y;
    const double area = 0.5 * (d21x * d31y - d21y * d31x);
    out[0] = area / 6.;
    out[ld] = area / 12.;
  }
  __device__ static __forceinline__ void unpack(int, const double *in, int ld, double &vd,
                                                double &vn, double &vp) {
    vd = in[0]; vn = vp = in[ld];
  }
};
struct LocalValues {  // 3x3 local matrix given per element, row-major
  static constexpr bool needs_xy = false;
  P1_GENERIC_PACK
  const double *__restrict__ local;
  __device__ __forceinline__ void rows(int64_t t, const int *, double2, double2, double2,
                                       double (*v)[3]) const {
    const double *m = local + 9 * t;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      v[k][0] = __ldg(m + 3 * k + k);
      v[k][1] = __ldg(m + 3 * k + (k == 2 ? 0 : k + 1));
      v[k][2] = __ldg(m + 3 * k + (k == 0 ? 2 : k - 1));
    }
  }
};

// A slim slot turned back into record values: vertex k of the element is the
// row's node, k+1 its next, k+2 its prev -- rotate the three coordinates back
// into element order so that the arithmetic is the scatter's, bit for bit.
template <class V>
__device__ __forceinline__ void slim_values(const V &values, int k, double2 za, double2 zn,
                                            double2 zp, double &vd, double &vn, double &vp) {
  const double2 z0 = k == 0 ? za : (k == 1 ? zp : zn);
  const double2 z1 = k == 0 ? zn : (k == 1 ? za : zp);
  const double2 z2 = k == 0 ? zp : (k == 1 ? zn : za);
  double v[3][3];
  values.rows(0, nullptr, z0, z1, z2, v);
  vd = k == 0 ? v[0][0] : (k == 1 ? v[1][0] : v[2][0]);
  vn = k == 0 ? v[0][1] : (k == 1 ? v[1][1] : v[2][1]);
  vp = k == 0 ? v[0][2] : (k == 1 ? v[1][2] : v[2][2]);
}

// solvers.py:306-366 evaluated inside the scatter: the quadrature sums of the
// four coefficients ([n_qp x M] arrays) and the local matrix never touch HBM
// as an [M x 9] array.
struct GeneralValues {
  static constexpr bool needs_xy = true;
  P1_GENERIC_PACK
  const double2 *__restrict__ coords;
  const double *__restrict__ a11q, *__restrict__ a12q, *__restrict__ a21q,
      *__restrict__ a22q;
  const double *__restrict__ weights;  // device, n_qp
  int n_qp;
  int64_t M;
  __device__ __forceinline__ void rows(int64_t t, const int *, double2 z0, double2 z1,
                                       double2 z2, double (*v)[3]) const {
    const double dz1x = z1.x - z0.x, dz1y = z1.y - z0.y;
    const double dz2x = z2.x - z0.x, dz2y = z2.y - z0.y;
    const double alpha = dz2y, beta = -dz2x, gamma = -dz1y, delta = dz1x;
    const double areas = 0.5 * (dz1x * dz2y - dz1y * dz2x);
    double a = 0., b = 0., c = 0., d = 0.;
    for (int q0 = 0; q0 < n_qp; q0 += 4) {  // 16 loads in flight, sums in order
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
          const double w = __ldg(weights + q0 + u);
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
    // local matrix m[3r+c] (solvers.py:356-366); v[k] = (m[k][k], m[k][k+1], m[k][k+2])
    v[0][0] = b11 + b12 + b21 + b22; v[0][1] = -(b11 + b21); v[0][2] = -(b12 + b22);
    v[1][0] = b11;                   v[1][1] = b12;          v[1][2] = -(b11 + b12);
    v[2][0] = b22;                   v[2][1] = -(b21 + b22); v[2][2] = b21;
  }
};

// solvers.py:86-129 evaluated inside the scatter
struct WeightedValues {
  static constexpr bool needs_xy = true;
  P1_GENERIC_PACK
  const double2 *__restrict__ coords;
  const double *__restrict__ phiq;     // [n_qp x M]
  const double *__restrict__ weights;  // device, n_qp
  const double2 *__restrict__ points;  // device, n_qp (eta, xi)
  int n_qp;
  int64_t M;
  __device__ __forceinline__ void rows(int64_t t, const int *, double2 z0, double2 z1,
                                       double2 z2, double (*v)[3]) const {
    const double dz1x = z1.x - z0.x, dz1y = z1.y - z0.y;
    const double dz2x = z2.x - z0.x, dz2y = z2.y - z0.y;
    const double areas = 0.5 * (dz1x * dz2y - dz1y * dz2x);
    double m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = 0.;
    for (int q0 = 0; q0 < n_qp; q0 += 4) {  // four loads in flight, sums in order
      double fv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        fv[u] = q0 + u < n_qp ? __ldg(phiq + (int64_t)(q0 + u) * M + t) : 0.;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q0 + u < n_qp) {
          const double2 pt = __ldg(points + q0 + u);
          const double h[3] = {1 - pt.x - pt.y, pt.x, pt.y};
          const double base = 2 * areas * __ldg(weights + q0 + u) * fv[u];
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) m[3 * k + l] += base * h[k] * h[l];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      v[k][0] = m[3 * k + k];
      v[k][1] = m[3 * k + (k == 2 ? 0 : k + 1)];
      v[k][2] = m[3 * k + (k == 0 ? 2 : k - 1)];
    }
  }
};

}  // namespace p1
