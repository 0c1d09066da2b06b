/* p1b200.h -- C ABI of libp1b200.so: the P1 assembly hot path on NVIDIA B200.
 *
 * The reference (leuraph/p1afempy v0.2.16) has no FFI layer: its boundary is a
 * set of module-level Python functions over numpy arrays.  This header is what
 * a binding for that path binds instead; every entry point names the reference
 * code it replaces (file:line under /root/reference).  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - all array arguments are DEVICE pointers unless the name ends in _host;
 *  - row-major, C-contiguous; float64 values; int32 node/element indices on
 *    the device (p1_narrow_i64 converts what numpy hands over);
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *    calls are stream-ordered and asynchronous unless stated otherwise;
 *  - every function returns 0 on success or a negative P1_E* code; the text
 *    for the calling thread's last failure is p1_last_error();
 *  - inputs are never modified; outputs are caller-allocated; scratch memory
 *    comes from the context's caching arena and goes back to it before the
 *    call returns (reuse is stream-ordered), except what a p1_plan keeps.
 */
#ifndef P1B200_H
#define P1B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P1_OK 0
#define P1_EINVAL (-1)      /* bad argument                                   */
#define P1_EINDEX (-2)      /* element/boundary index outside [0, n_nodes)    */
#define P1_ECUDA (-3)       /* CUDA runtime error, see p1_last_error()        */
#define P1_ETOPOLOGY (-4)   /* mesh is not an edge-manifold / boundary edges  */
                            /* not covered (the reference raises ValueError,  */
                            /* mesh.py:150)                                   */
#define P1_ERANGE (-5)      /* size exceeds what int32 device indices hold    */
#define P1_ECOMM (-6)       /* NCCL / multi-GPU error                         */
#define P1_ERETRY (-7)      /* the plan's closed-fan shortcut was contradicted */
                            /* by the mesh (non-manifold vertex, degenerate   */
                            /* element): rebuild the plan with P1_PLAN_EXACT  */
                            /* and assemble again                             */

typedef struct p1_ctx p1_ctx;
typedef struct p1_plan p1_plan;

int p1_version(void);
const char *p1_last_error(void);

/* One context per (process, device).  Owns the error flags and the pool. */
int p1_ctx_create(int device, p1_ctx **out);
int p1_ctx_destroy(p1_ctx *ctx);
/* Scratch memory is cached in the context between calls (same sizes recur on
 * the same mesh); p1_ctx_trim returns the unused blocks to the driver. */
int p1_ctx_trim(p1_ctx *ctx);
/* Idle scratch beyond `bytes` is returned as soon as a block is released (largest
 * first); default 48 GB.  0 = cache nothing. */
int p1_ctx_set_cache_limit(p1_ctx *ctx, int64_t bytes);

/* numpy hands elements over as int64 (refinement.py:641) or uint32
 * (io_helpers.py:91); the device works on int32.  Values outside
 * [0, 2^31) are clamped to -1 so that the range check of p1_plan_create
 * reports them.  src_kind: 0 = int64, 1 = uint32, 2 = int32 (plain copy). */
int p1_narrow_indices(p1_ctx *ctx, const void *src, int src_kind,
                      int64_t count, int32_t *dst, void *stream);
/* Per-kernel device timing (CUDA events on the launching stream) for
 * bench.py's roofline line.  Off by default; p1_profile_report writes one
 * "name total_ms launches" line per kernel into buf and SYNCHRONISES. */
int p1_profile_enable(p1_ctx *ctx, int on);
int p1_profile_reset(p1_ctx *ctx);
int p1_profile_report(p1_ctx *ctx, char *buf, int64_t cap);

/* P1_EINDEX unless every value lies in [0, n_nodes).  The per-triangle calls
 * below gather by element index without checking; run this (or p1_plan_create,
 * which checks too) on untrusted input first.  SYNCHRONISES. */
int p1_check_indices(p1_ctx *ctx, const int32_t *indices, int64_t count,
                     int64_t n_nodes, void *stream);
/* int32 -> int64 widening of an index array (scipy switches the CSR index
 * dtype to int64 when max(9M, N) >= 2^31, scipy/sparse/_coo.py:419). */
int p1_widen_indices(p1_ctx *ctx, const int32_t *src, int64_t count,
                     int64_t *dst, void *stream);

/* ------------------------------------------------------------------------
 * Plan = everything that depends on `elements` only: the node -> incident
 * (element, local vertex) lists in the fixed order (element, vertex), the
 * neighbour pairs, and the CSR row pointer.  Replaces what scipy's
 * coo_tocsr + csr_sort_indices + csr_sum_duplicates derive on every call
 * (scipy/sparse/_coo.py:368-439, _compressed.py:1072-1130).
 *
 * row_begin/row_end restrict the plan to the rows [row_begin,row_end) of the
 * matrix (multi-GPU row ownership); pass 0, n_nodes for the whole matrix.
 * SYNCHRONISES the stream once (the caller needs nnz to allocate outputs).
 * ---------------------------------------------------------------------- */
#define P1_PLAN_DEFAULT 0
#define P1_PLAN_EXACT 1     /* count the columns of every row by comparison   */
#define P1_PLAN_KEEP_ELEMENTS 2 /* remember the element of every record: needed */
                            /* to assemble a second operator on the same plan, */
                            /* for load vectors, edge numbering and eta_R      */
#define P1_PLAN_TOPOLOGY 4  /* p1_plan_create only: 16 bytes per incidence, no     */
                            /* values and no row pointer -- all that load vectors, */
                            /* edge numbering and eta_R need.  8 elements per node */
                            /* plus 1024 in total beyond that; P1_ERETRY otherwise */
                            /* (use P1_PLAN_KEEP_ELEMENTS for such a mesh)         */
#define P1_PLAN_SLOT_INDEX 8 /* with P1_PLAN_TOPOLOGY: also record the slot of every   */
                             /* half-edge, so that p1_eta_r gets ALL half-edge twins  */
                             /* by a gather instead of scattered stores (+0.2 ms on   */
                             /* the plan, -1.5 ms on the twins at 64M); not worth it  */
                             /* for p1_geometric_data alone, which reads half of them */
/* Diagnostic flags (tests and measurements choose a storage the library would not
 * pick by itself; results are the same): */
#define P1_PLAN_FORCE_RECORDS 16 /* one GPU: 32-byte records through HBM for every  */
                                 /* incidence instead of the tiled path             */
#define P1_PLAN_FORCE_SLIM 32    /* one GPU: the 8-byte slots of multi-GPU plans     */
#define P1_PLAN_TINY_SPILL 64    /* spill list of one record: rows with more than 8  */
                                 /* records take the second pass over the elements   */
#define P1_PLAN_TILED 128        /* one GPU, all rows, p1_plan_create_for/_quad: the  */
                                 /* tiled path (see below; measured: 5 GB instead of  */
                                 /* 18 GB of DRAM traffic per 64M-triangle step, but  */
                                 /* instruction-bound and slower: DESIGN.md)          */
int p1_plan_create(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                   int64_t n_nodes, int64_t row_begin, int64_t row_end,
                   int flags, void *stream, p1_plan **out);
/* The cold path of one matrix; assemble with P1_OP_STORED afterwards.  op/coords/
 * local as for p1_assemble_csr; they are borrowed until the plan is destroyed.
 * The scatter that groups the incidences also computes the local matrix of every
 * element (once) and stores its rows in the records.  With P1_PLAN_TILED (one GPU,
 * all rows, no P1_PLAN_KEEP_ELEMENTS): a count pass fixes the row pointer, then
 * blocks of 1024 consecutive elements finish in shared memory every row whose
 * elements all lie in the block and write it straight into the CSR arrays; only the
 * other rows' incidences travel through HBM as 32-byte records. */
int p1_plan_create_for(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                       int64_t n_nodes, int64_t row_begin, int64_t row_end,
                       int flags, int op, const double *coords, const double *local,
                       void *stream, p1_plan **out);
/* Generalised stiffness / weighted mass with the quadrature sums and the local
 * matrix evaluated inside the scatter (no [M x 9] array through HBM).  Device
 * arrays are [n_qp x M], qp-major; the rule is given on the host. */
typedef struct p1_quad_args {
  const double *a11q, *a12q, *a21q, *a22q; /* P1_OP_GENERAL                    */
  const double *phiq;                      /* P1_OP_WEIGHTED                   */
  const double *weights_host;              /* n_qp                             */
  const double *points_host;               /* n_qp x 2 (eta, xi); WEIGHTED only */
  int n_qp;
} p1_quad_args;
int p1_plan_create_quad(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                        int64_t n_nodes, int64_t row_begin, int64_t row_end,
                        int flags, int op, const double *coords,
                        const p1_quad_args *args, void *stream, p1_plan **out);
/* Multi-GPU with balanced ownership: of the chunks of 2^chunk_log2 consecutive
 * rows, rank r owns those with chunk index % world == r; its local rows are the
 * owned chunks in ascending order (p1_plan_rows of them).  Every rank then holds
 * the same mix of old, high-valence and new nodes.  Interleaving the ranks'
 * pieces chunk by chunk gives the reference CSR bit for bit. */
int p1_plan_create_interleaved(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                               int64_t n_nodes, int chunk_log2, int world, int rank,
                               int flags, int op, const double *coords,
                               const double *local, void *stream, p1_plan **out);
int64_t p1_plan_rows(const p1_plan *plan); /* rows the plan owns (length of indptr - 1) */
int p1_plan_destroy(p1_plan *plan);
/* Multi-GPU: world+1 row boundaries such that every range holds about the same
 * number of stored entries (estimated from a strided sample of the elements;
 * deterministic, so all ranks agree without communication).  SYNCHRONISES. */
int p1_partition_rows(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                      int64_t n_nodes, int world, int64_t *boundaries_host,
                      void *stream);
/* nnz: stored entries in the owned rows; max_index: largest node index that
 * occurs in `elements` (the reference infers the shape of the stiffness and
 * mass matrices as max_index+1, solvers.py:46-47,153). */
int p1_plan_info(const p1_plan *plan, int64_t *nnz, int64_t *max_index,
                 int64_t *n_elements, int64_t *n_nodes);

/* What to assemble.  The value of local entry (r, c) of element T goes to
 * (row, col) = (elements[T][r], elements[T][c]). */
#define P1_OP_STORED 0    /* the values the plan's records already hold
                             (p1_plan_create_for, or the previous assemble)  */
#define P1_OP_STIFFNESS 1 /* solvers.py:15-47   get_stiffness_matrix          */
#define P1_OP_MASS 2      /* solvers.py:145-182 get_mass_matrix               */
#define P1_OP_GENERAL 4   /* solvers.py:278-377: local matrices computed inside the
                             scatter from p1_quad_args (p1_plan_create_quad only) */
#define P1_OP_WEIGHTED 5  /* solvers.py:50-142: likewise                       */
#define P1_OP_LOCAL 3     /* any 3x3 local matrix given per element, row-major
                             local[T*9 + 3*r + c] (general stiffness, weighted
                             mass after p1_local_* below)                     */

/* Assemble to CSR over the owned rows.
 *   indptr  : (row_end-row_begin+1) int32, local offsets starting at 0
 *   indices : nnz int32, sorted and unique per row, explicit zeros kept
 *   data    : nnz float64; an off-diagonal entry is the sum of its (at most two)
 *             contributions, a diagonal entry the sum of the (k,k) entries of
 *             the incident elements taken in ascending order of their `next`
 *             vertex (= ascending column): a fixed order, whatever the path
 * indptr and/or indices may be NULL (re-assembly on an unchanged mesh writes
 * values only).  `coords` (n_nodes x 2) is used by P1_OP_STIFFNESS/MASS,
 * `local` (n_elements x 9) by P1_OP_LOCAL; any op other than P1_OP_STORED
 * recomputes the record values first and needs P1_PLAN_KEEP_ELEMENTS.
 * Returns P1_ERETRY when the mesh contradicts a default plan (see above). */
int p1_assemble_csr(const p1_plan *plan, int op, const double *coords,
                    const double *local, int32_t *indptr, int32_t *indices,
                    double *data, void *stream);

/* The whole cold assembly (plan + fill of p1_plan_create_for / _quad / _interleaved and
 * p1_assemble_csr(P1_OP_STORED)) WITHOUT a host synchronisation: nothing the host
 * would have to wait for is needed to launch the next kernel.  The outputs have a
 * CAPACITY instead of an exact size (indices / data: `capacity` entries; any mesh
 * whose boundary nodes have one open fan each satisfies nnz <= 2 * rows + incidences
 * of the owned rows), indptr has rows + 1 entries and indptr[rows] = nnz.
 * status: 8 device ints, valid once the stream has passed the call:
 *   [0] P1_STATUS_* bits, [1] largest node index in `elements`, [6] nnz.
 * When a status bit is set NOTHING was written to indices / data; the caller takes
 * the plan route (P1_STATUS_SLOWPATH: a node with more than 8 elements, or nnz above
 * the capacity; P1_STATUS_RETRY: as P1_ERETRY) or reports the error (INDEX).
 * world / rank / chunk_log2 as for p1_plan_create_interleaved (1, 0, 0: all rows).
 * qa: P1_OP_GENERAL / P1_OP_WEIGHTED only, else NULL.  Scratch memory returns to the
 * context's arena in stream order.
 * Meshes of at most 4M triangles (P1_OP_STIFFNESS / MASS / LOCAL, profiler off): the
 * second call with exactly the same arguments (pointers, sizes, stream) records the
 * launch sequence as a CUDA graph and later calls replay it with ONE launch; the
 * arrays' CONTENT may change freely between calls.  The context keeps the 8 most
 * recently used graphs with their scratch until p1_ctx_trim / p1_ctx_destroy. */
#define P1_STATUS_INDEX 1
#define P1_STATUS_RETRY 16
#define P1_STATUS_SLOWPATH 64
int p1_assemble_cold(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                     int64_t n_nodes, int chunk_log2, int world, int rank, int op,
                     const double *coords, const double *local, const p1_quad_args *qa,
                     int64_t capacity, int32_t *indptr, int32_t *indices, double *data,
                     int32_t *status, void *stream);

/* solvers.py:156-182 get_mass_matrix_elements: the 9M triplets in the
 * reference's order (element-major, entry k: row = vertex k%3, col = k/3). */
int p1_mass_triplets(p1_ctx *ctx, const double *coords,
                     const int32_t *elements, int64_t n_elements,
                     int64_t *rows, int64_t *cols, double *vals, void *stream);
/* same stream of triplets for the stiffness matrix (solvers.py:34-47). */
int p1_stiffness_triplets(p1_ctx *ctx, const double *coords,
                          const int32_t *elements, int64_t n_elements,
                          int64_t *rows, int64_t *cols, double *vals,
                          void *stream);

/* mesh.py:10-51 get_area / get_directional_vectors (any output may be NULL):
 * area (M), d21 (M x 2), d31 (M x 2). */
int p1_geometry(p1_ctx *ctx, const double *coords, const int32_t *elements,
                int64_t n_elements, double *area, double *d21, double *d31,
                void *stream);

/* Points at which user coefficient callbacks are evaluated.
 *  xq[q][T][0..1] = z0 + eta_q*dz1 + xi_q*dz2           (solvers.py:326,582)
 *  uq[q][T]       = u1*(1-eta-xi) + u2*eta + u3*xi      (solvers.py:113-117)
 *  centroids[T]   = c0 + (d21+d31)/3                    (solvers.py:421,
 *                                                        indicators.py:84)
 * points_host: n_qp x 2 (eta, xi) in HOST memory. */
int p1_quadrature_points(p1_ctx *ctx, const double *coords,
                         const int32_t *elements, int64_t n_elements,
                         const double *points_host, int n_qp, double *xq,
                         void *stream);
int p1_nodal_at_quadrature(p1_ctx *ctx, const double *u,
                           const int32_t *elements, int64_t n_elements,
                           const double *points_host, int n_qp, double *uq,
                           void *stream);
int p1_centroids(p1_ctx *ctx, const double *coords, const int32_t *elements,
                 int64_t n_elements, double *centroids, void *stream);

/* solvers.py:278-377 get_general_stiffness_matrix: local 3x3 matrices from
 * the coefficient values at the quadrature points, a_ij_q[q*M + T]. */
int p1_local_general_stiffness(p1_ctx *ctx, const double *coords,
                               const int32_t *elements, int64_t n_elements,
                               const double *a11q, const double *a12q,
                               const double *a21q, const double *a22q,
                               const double *weights_host, int n_qp,
                               double *local, void *stream);
/* solvers.py:50-142 get_weighted_mass_matrix: phiq[q*M + T] = phi(u_q). */
int p1_local_weighted_mass(p1_ctx *ctx, const double *coords,
                           const int32_t *elements, int64_t n_elements,
                           const double *phiq, const double *weights_host,
                           const double *points_host, int n_qp, double *local,
                           void *stream);

/* solvers.py:540-598 get_right_hand_side_using_quadrature_rule; fq[q*M + T].
 * b has n_nodes entries; summation order per node = the reference's
 * sequential bincount order (element, vertex), i.e. bit-identical. */
int p1_load_vector(const p1_plan *plan, const double *coords,
                   const double *fq, const double *weights_host,
                   const double *points_host, int n_qp, double *b,
                   void *stream);
/* solvers.py:429-472 integrate_composition_nonlinear_with_fem (SURVEY.md 8(f),
 * "next" rank 1): sum over the triangles of sum_q w_q * fq[q*M+T] * 2. * area_T,
 * fq = f(u_q) with u_q from p1_nodal_at_quadrature.  The load vector of the
 * composition (solvers.py:475-537) is p1_load_vector fed with the same fq.
 * Fixed-order summation tree; SYNCHRONISES (the result is a host scalar). */
int p1_integrate_quadrature(p1_ctx *ctx, const double *coords, const int32_t *elements,
                            int64_t n_elements, const double *fq,
                            const double *weights_host, int n_qp, double *result_host,
                            void *stream);

/* solvers.py:414-426 legacy centroid rule; f_centroid[T]; summation order
 * per node = (vertex, element), as bincount over elements.flatten('F'). */
int p1_load_vector_midpoint(const p1_plan *plan, const double *coords,
                            const double *f_centroid, double *b, void *stream);

/* Cold load vectors without a plan (nothing is cached between calls, like the
 * reference's stateless get_right_hand_side*, solvers.py:380-426, 540-598): the
 * quadrature sum / the midpoint share of every element is computed by the thread
 * that scans the element and delivered to the element's three nodes as (element
 * code, value); each node then adds its values in the reference's bincount order.
 * Bit-identical to p1_load_vector / p1_load_vector_midpoint.  fq: [n_qp x M]
 * values of f at the quadrature points (p1_quadrature_points).  P1_ERETRY when
 * the nodes with more than 8 elements hold more than 1024 incidences beyond 8:
 * use a plan (P1_PLAN_KEEP_ELEMENTS) and the functions above. */
int p1_load_vector_direct(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                          int64_t n_nodes, const double *coords, const double *fq,
                          const double *weights_host, const double *points_host,
                          int n_qp, double *b, void *stream);
int p1_load_vector_midpoint_direct(p1_ctx *ctx, const int32_t *elements,
                                   int64_t n_elements, int64_t n_nodes,
                                   const double *coords, const double *f_centroid,
                                   double *b, void *stream);

/* mesh.py:87-161 provide_geometric_data.  boundary_nodes: concatenated
 * (K x 2) int32 rows of all boundaries; boundary_offsets_host: n_boundaries+1
 * row offsets (host).  Outputs: element2edges (M x 3), edge2nodes (E x 2: the
 * caller provides (3M+K+1)/2 rows, E = (3M+K)/2 are written; on the error path
 * below nothing is written past that capacity), boundary2edges (K);
 * *n_edges_host receives E.  SYNCHRONISES the stream. */
int p1_geometric_data(const p1_plan *plan, const int32_t *boundary_nodes,
                      const int64_t *boundary_offsets_host, int n_boundaries,
                      int64_t *element2edges, int64_t *edge2nodes,
                      int64_t *boundary2edges, int64_t *n_edges_host,
                      void *stream);
/* number of undirected edges the call above will report: (3M+K)/2; returns
 * P1_ETOPOLOGY when forward and backward half-edge counts differ. */

/* indicators.py:7-86 compute_eta_r.  dirichlet / neumann: (K x 2) int32;
 * neumann_term[j] = |E_j| * g(midpoint_j) (indicators.py:70-76, evaluated by
 * the caller because g is a user callback); f_centroid[T] = f(centroid_T).
 * eta has n_elements entries. */
int p1_eta_r(const p1_plan *plan, const double *coords, const double *x,
             const int32_t *dirichlet, int64_t n_dirichlet,
             const int32_t *neumann, int64_t n_neumann,
             const double *neumann_term, const double *f_centroid,
             double *eta, void *stream);
/* Multi-GPU eta_R: the indicators of an element block [m0, m1) only need the sub-mesh
 * of the elements that touch a node of the block.  p1_block_submesh flags the block's
 * nodes (node_flag: n_nodes bytes) and numbers the elements of the sub-mesh (pos:
 * n_elements + 1 ints; counts_host: [0] their number, [1] where the block starts among
 * them -- its elements stay consecutive); p1_block_submesh_fill writes them (and their
 * global indices); p1_eta_r_partial is p1_eta_r on such a sub-mesh: boundary edges not
 * listed because they are the cut through the mesh are no error.  The block's
 * indicators computed this way are the global ones bit for bit. */
int p1_block_submesh(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                     int64_t n_nodes, int64_t m0, int64_t m1, unsigned char *node_flag,
                     int32_t *pos, int64_t *counts_host, void *stream);
int p1_block_submesh_fill(p1_ctx *ctx, const int32_t *elements, int64_t n_elements,
                          const int32_t *pos, int32_t *sub_elements, int32_t *sub_ids,
                          void *stream);
int p1_eta_r_partial(const p1_plan *plan, const double *coords, const double *x,
                     const int32_t *dirichlet, int64_t n_dirichlet,
                     const int32_t *neumann, int64_t n_neumann,
                     const double *neumann_term, const double *f_centroid,
                     double *eta, void *stream);
/* Multi-GPU matrices, interleaved ownership: the stored entries of each of the calling
 * rank's chunks (out: per_rank int64, chunks beyond the rank's last one count 0) and,
 * from the all-gathered values of every rank (every[r * per_rank + q]), the offset of
 * each of its chunks in the GLOBAL CSR plus the global nnz -- one launch each around
 * the one collective of a multi-GPU assembly. */
int p1_chunk_nnz(p1_ctx *ctx, const int32_t *indptr, int64_t n_own, int chunk_log2,
                 int64_t per_rank, int64_t *out, void *stream);
int p1_chunk_offsets(p1_ctx *ctx, const int64_t *every, int world, int64_t per_rank, int rank,
                     int64_t *mine, int64_t *total, void *stream);
/* |E_j| and midpoints of boundary edges (indicators.py:70-76): length (K),
 * midpoint (K x 2). */
int p1_boundary_edges(p1_ctx *ctx, const double *coords,
                      const int32_t *edges, int64_t n_edges, double *length,
                      double *midpoint, void *stream);

/* ------------------------------------------------------------------------
 * The callers either side of the assembly path.
 * ---------------------------------------------------------------------- */

/* mesh.py:495-528 get_element_to_neighbours: neighbours[3 T + i] = the element that
 * shares edge i (vertex i -> vertex i+1) of element T, -1 on the boundary.  Needs a
 * plan that remembers its elements (P1_PLAN_KEEP_ELEMENTS or P1_PLAN_TOPOLOGY |
 * P1_PLAN_SLOT_INDEX) over all rows.  SYNCHRONISES when the twins are built. */
int p1_element_neighbours(const p1_plan *plan, int64_t *neighbours, void *stream);

/* refinement.py:508-718 refineNVB in two stages around the caller's allocation of the
 * refined mesh (p1_geometric_data provides element2edges / edge2nodes / boundary2edges).
 * refinement.py:556-567: sort_for_longest_egde rotates the elements IN PLACE. */
int p1_nvb_sort_longest_edge(p1_ctx *ctx, const double *coords, int32_t *elements,
                             int64_t n_elements, void *stream);
/* Stage 1 (refinement.py:574-592 and the counts of 625-637): edges of the marked
 * elements, closure over the reference edges, numbering of the new nodes in edge
 * order.  Scratch (device, int32): edge_flag (E), edge_new (E+1), child_first (M+1),
 * bnd_before (K+1).  counts_host: [0] new nodes, [1] refined elements, [2+j] split
 * edges of boundary j.  SYNCHRONISES. */
int p1_nvb_plan(p1_ctx *ctx, int64_t n_elements, int64_t n_edges,
                const int64_t *element2edges, const int64_t *marked, int64_t n_marked,
                const int64_t *boundary2edges, const int64_t *boundary_offsets_host,
                int n_boundaries, int32_t *edge_flag, int32_t *edge_new,
                int32_t *child_first, int32_t *bnd_before, int64_t *counts_host,
                void *stream);
/* Stage 2 (refinement.py:8-31, 594-716): midpoints (and embedded values) of the new
 * nodes behind the old ones, the refined boundaries (unsplit edges, then first halves,
 * then second halves), the children of every element in the reference's order.
 * new_coords / new_embed already hold the old nodes' values in their first n_nodes
 * entries; new_boundaries_host: host array of device pointers. */
int p1_nvb_apply(p1_ctx *ctx, const double *coords, int64_t n_nodes,
                 const int32_t *elements, int64_t n_elements, int64_t n_edges,
                 const int64_t *element2edges, const int64_t *edge2nodes,
                 const int32_t *boundary_nodes, const int64_t *boundary2edges,
                 const int64_t *boundary_offsets_host, int n_boundaries,
                 const int32_t *edge_flag, const int32_t *edge_new,
                 const int32_t *child_first, const int32_t *bnd_before,
                 const int64_t *counts_host, const double *embed, int sorted,
                 double *new_coords, double *new_embed, int32_t *new_elements,
                 int32_t *const *new_boundaries_host, void *stream);

/* solvers.py:601-618 apply_neumann: out = b + bincount over the Neumann edges of
 * |E| g(midpoint) / 2 (first endpoints, then second endpoints).  SYNCHRONISES. */
int p1_apply_neumann(p1_ctx *ctx, const double *coords, int64_t n_nodes,
                     const int32_t *edges, int64_t n_edges, const double *g_mid,
                     const double *b, double *out, void *stream);
/* y = y0 + alpha A x on a device CSR (int32 indices): the Dirichlet lift b - A x_D
 * (solvers.py:679-681) and A x of the energy (solvers.py:695).  y0 may be NULL. */
int p1_csr_spmv(p1_ctx *ctx, int64_t n_rows, const int32_t *indptr, const int32_t *indices,
                const double *data, const double *x, double alpha, const double *y0,
                double *y, void *stream);
/* a . b, fixed summation tree.  SYNCHRONISES. */
int p1_dot(p1_ctx *ctx, int64_t n, const double *a, const double *b, double *result_host,
           void *stream);
/* solvers.py:688-693 without leaving the device: Jacobi-preconditioned conjugate
 * gradient on the free nodes (fixed[i] != 0: x[i] keeps its Dirichlet value; rows and
 * columns of fixed nodes are lifted to the right-hand side).  SYNCHRONISES. */
int p1_pcg(p1_ctx *ctx, int64_t n, const int32_t *indptr, const int32_t *indices,
           const double *data, const double *b, const unsigned char *fixed, double *x,
           double tol, int max_iter, int *iterations_host, double *relative_residual_host,
           void *stream);

#ifdef __cplusplus
}
#endif
#endif /* P1B200_H */
