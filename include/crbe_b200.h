/*
 * crbe_b200.h -- C ABI of libcrbe_b200.so, the sm_100a implementation of the
 * CRBE hot path (Crouzeix-Raviart FEM + Backward Euler, ``BESCRFEM``) of
 * clemsadand/AirPollution.
 *
 * The reference has no FFI: its boundary is the Python class API of crbe.py.
 * Each entry point below replaces the arithmetic of the cited reference lines;
 * the host mirror (airpollution_b200/crbe.py) binds them with ctypes and keeps
 * the reference's class/attribute names.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (CRBE_ERR_*); the
 *     message is available from crbe_last_error() (thread local).
 *   - pointers named *_d are DEVICE pointers owned by the caller (the library
 *     never frees them); pointers named *_h are HOST pointers.  Values are
 *     IEEE float64, indices int32, sizes int64 -- as the reference produces.
 *   - opaque handles own library-internal device memory and are released by
 *     their *_destroy / *_free function.
 *   - all work is enqueued on the context's stream; functions that return
 *     host-side results synchronise that stream before returning.
 *   - one host thread per context.
 */
#ifndef CRBE_B200_H
#define CRBE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRBE_ABI_VERSION 1

#define CRBE_OK 0
#define CRBE_ERR_ARG (-1)      /* bad argument                                  */
#define CRBE_ERR_CUDA (-2)     /* CUDA runtime error                            */
#define CRBE_ERR_MESH (-3)     /* non-manifold / degenerate mesh                */
#define CRBE_ERR_SOLVER (-4)   /* BiCGStab breakdown or iteration limit         */
#define CRBE_ERR_COMM (-5)     /* NCCL / peer-access error                      */

typedef struct crbe_ctx crbe_ctx;
typedef struct crbe_topology crbe_topology;
typedef struct crbe_solver crbe_solver;

/* ---- errors / lifetime ------------------------------------------------- */
int crbe_abi_version(void);
const char* crbe_last_error(void);
int crbe_ctx_create(int device, crbe_ctx** out);
int crbe_ctx_set_stream(crbe_ctx* ctx, void* cuda_stream);   /* cudaStream_t; NULL = the default stream */
int crbe_ctx_synchronize(crbe_ctx* ctx);
int crbe_ctx_destroy(crbe_ctx* ctx);
/* plain copies on the context stream (so non-torch hosts need nothing else) */
int crbe_memcpy_h2d(crbe_ctx* ctx, void* dst_d, const void* src_h, int64_t bytes, int sync);
int crbe_memcpy_d2h(crbe_ctx* ctx, void* dst_h, const void* src_d, int64_t bytes, int sync);

/* ---- a-1 / a-2: DOF numbering and mesh geometry (crbe.py:50-154) -------- */
/* Phase 1: first-seen edge numbering of the triangle list (crbe.py:109-131).
 * tri_d: nt x 3 vertex ids.  Sizes are returned so the caller can allocate. */
int crbe_topology_create(crbe_ctx* ctx, const int32_t* tri_d, int64_t nt, int64_t nv,
                         crbe_topology** out, int64_t* n_segments, int64_t* n_boundary_segments,
                         int64_t* n_boundary_triangles);
/* Phase 2: write the arrays.
 *   t2s_d       nt x 3   triangle_to_segments                    (crbe.py:129)
 *   segments_d  N  x 2   [min,max] vertex ids in id order        (crbe.py:128)
 *   edge_slots_d N x 2   the (<=2) slots 3*t+a holding each edge; second = -1 on the boundary
 *   bnd_seg_d   Nb       sorted ids of edges seen once           (crbe.py:78-80)
 *   bnd_tri_d   Nbt      triangles with a boundary edge          (crbe.py:83-95)
 *   bnd_tri_seg_d Nbt    their first boundary edge in local order (crbe.py:92)        */
int crbe_topology_fill(crbe_topology* topo, int32_t* t2s_d, int32_t* segments_d, int32_t* edge_slots_d,
                       int32_t* bnd_seg_d, int32_t* bnd_tri_d, int32_t* bnd_tri_seg_d);
int crbe_topology_free(crbe_topology* topo);
/* midpoints (crbe.py:71), lengths (crbe.py:134-141), areas (crbe.py:143-154), diameter (crbe.py:98-106).
 * points_d: nv x 2.  Any output may be NULL. */
int crbe_mesh_geometry(crbe_ctx* ctx, const double* points_d, int64_t nv, const int32_t* tri_d, int64_t nt,
                       const int32_t* segments_d, int64_t n_seg, double* midpoints_d, double* lengths_d,
                       double* areas_d, double* diameter_h);

/* ---- a-6 (structure): CSR pattern from edge connectivity (crbe.py:352-354) */
/* indptr_d: N+1 (out).  The pattern equals scipy's canonical csr_matrix of the
 * COO triplets (sorted columns, duplicates merged, explicit zeros kept). */
int crbe_csr_pattern_count(crbe_ctx* ctx, const int32_t* t2s_d, const int32_t* edge_slots_d, int64_t n_seg,
                           int32_t* indptr_d, int64_t* nnz_h);
/* indices_d: nnz (out); scatter_pos_d: nt x 9 (out) CSR slot of local entry (a,b) of each triangle. */
int crbe_csr_pattern_fill(crbe_ctx* ctx, const int32_t* t2s_d, const int32_t* edge_slots_d, int64_t n_seg,
                          int64_t nt, const int32_t* indptr_d, int32_t* indices_d, int32_t* scatter_pos_d);
/* Element colouring of the dual graph (no two triangles of one colour share an
 * edge), deterministic.  colour_d: nt (out); order_d: nt (out) triangles grouped
 * by colour, ascending inside a colour; colour_offsets_h: 9 entries (out). */
int crbe_colour_elements(crbe_ctx* ctx, const int32_t* t2s_d, const int32_t* edge_slots_d, int64_t nt,
                         int32_t* colour_d, int32_t* order_d, int64_t* colour_offsets_h, int32_t* n_colours_h);

/* ---- a-3..a-6 (values): element matrices and global assembly ------------ */
/* Local matrices of every triangle, for parity checks: k_loc_d, a_loc_d: nt x 9,
 * m_loc_d: nt x 9 (crbe.py:249-313).  v_elem_d: NULL or nt x 2 per-element velocity. */
int crbe_element_matrices(crbe_ctx* ctx, const double* points_d, const int32_t* tri_d, const double* areas_d,
                          int64_t nt, double D, double vx, double vy, const double* v_elem_d,
                          double* k_loc_d, double* m_loc_d, double* a_loc_d);
/* One thread per element, colour by colour, no atomics (crbe.py:336-354).
 * Any of m_val_d, k_val_d, a_val_d (nnz each, structural pattern) may be NULL. */
int crbe_assemble(crbe_ctx* ctx, const double* points_d, const int32_t* tri_d, const double* areas_d,
                  const int32_t* scatter_pos_d, const int32_t* order_d, const int64_t* colour_offsets_h,
                  int32_t n_colours, int64_t nnz, double D, double vx, double vy, const double* v_elem_d,
                  double* m_val_d, double* k_val_d, double* a_val_d);
/* s = m + coef*(k + a) entry by entry in the reference's order (crbe.py:358,360,386). */
int crbe_system_values(crbe_ctx* ctx, int64_t nnz, const double* m_val_d, const double* k_val_d,
                       const double* a_val_d, double coef, double* s_val_d);

/* Time-varying velocity (BASELINE.json config 5; the reference's v is constant, crbe.py:309, and is the special case
 * v_elem_d == NULL).  A_loc (crbe.py:284-313) is linear in v with coefficients that depend on the triangle only:
 * crbe_solver_advection_plan lays them out once for the loaded solver (per triangle the inverse Jacobian and area/6 of
 * crbe.py:291-310, per row one word describing where its <= 2 triangles add).  points/tri/areas/scatter_pos are read during
 * the call only; edge_slots_d and k_val_d must stay alive while the plan is used.  Requires crbe_solver_set_system. */
int crbe_solver_advection_plan(crbe_solver* s, const double* points_d, const int32_t* tri_d, const double* areas_d, int64_t nt,
                               const int32_t* edge_slots_d, const int32_t* scatter_pos_d, const double* k_val_d);
/* Rebuild A(v) and the rows of the loaded solver system in ONE pass over the rows: each row forms m + coef*(k + a) in the
 * reference's order (crbe.py:358), applies the Dirichlet rows and the diagonal scaling and writes the solver's layout; nothing
 * of size nnz is materialised unless a_val_out_d / s_val_out_d (structural pattern) are given.  Bit-identical to
 * crbe_assemble + crbe_system_values + crbe_solver_set_system.  write_system != 0: the system rows; write_rhs != 0: the
 * Crank-Nicolson operator M - coef*(K + A) (crbe.py:386), which belongs to the old time level of a step -- rebuild the system
 * before a step (write_rhs = 0) and the operator after it (write_system = 0, write_rhs = 1).  A row that comes out unusable
 * (zero diagonal) is reported by the next crbe_solver_step*. */
int crbe_solver_update_advection(crbe_solver* s, const double* v_elem_d, double vx, double vy, double coef, int32_t write_system,
                                 int32_t write_rhs, double* a_val_out_d, double* s_val_out_d);

/* ---- kernels exposed for unit tests and profiling ----------------------- */
int crbe_spmv_csr(crbe_ctx* ctx, int64_t n, const int32_t* indptr_d, const int32_t* indices_d,
                  const double* val_d, const double* x_d, double* y_d);
int crbe_dot(crbe_ctx* ctx, int64_t n, const double* x_d, const double* y_d, double* out_h);
/* (rel_l2, l2, max) of crbe.py:447-453 */
int crbe_errors(crbe_ctx* ctx, int64_t n, const double* u_exact_d, const double* u_num_d, double* out3_h);
/* The raw sums behind crbe_errors for a block of DOFs: out3_h = sum error^2, sum u_exact^2, max error.  A partitioned solve adds
 * the first two and takes the maximum of the third over the ranks before forming the triple of crbe.py:447-453. */
int crbe_error_sums(crbe_ctx* ctx, int64_t n, const double* u_exact_d, const double* u_num_d, double* out3_h);

/* Plume diagnostics of the reference's analysis scripts (scripts/problem3_comprehensive_analysis2.py:60-302:
 * mass, centre of mass, spread, peak) in one pass over the DOFs: with the CR quadrature they are weighted sums
 * with weights_d = diag(M).  out8_h: sum w u, sum w u x, sum w u y, sum w u x^2, sum w u y^2, peak value,
 * peak DOF index, 0. */
int crbe_moments(crbe_ctx* ctx, int64_t n, const double* u_d, const double* weights_d, const double* midpoints_d,
                 double* out8_h);

/* ---- a-9..a-11: the per-step linear solve ------------------------------- */
typedef struct crbe_solve_info {
    int32_t iterations;      /* BiCGStab iterations of this solve                     */
    int32_t restarts;        /* restarts from the true residual                       */
    int32_t status;          /* 0 converged; 1 iteration limit; 2 breakdown (3, 4: see CRBE_SOLVER_NO_PREDICT, CRBE_ERR_COMM) */
    int32_t launches;        /* kernels launched by this call                         */
    double relres;           /* recurrence residual  ||r|| / ||b||  (Jacobi-scaled)   */
    double true_relres;      /* ||b - A x|| / ||b|| recomputed after convergence      */
    double bnorm;            /* ||b|| (Jacobi-scaled)                                  */
    double initial_relres;   /* ||b - A x0|| / ||b|| of the initial guess (crbe_solver_step*)  */
    int32_t guess_order;     /* order of the extrapolated initial guess used by this step      */
    int32_t reserved;
} crbe_solve_info;

#define CRBE_SOLVER_FUSED 1u          /* reserved (the iteration is the merged-reduction form)   */
#define CRBE_SOLVER_VERIFY 2u         /* recompute the true residual after every convergence   */
#define CRBE_SOLVER_INDEX32 64u       /* keep 32-bit column indices in the bulk-copy kernels even when every column - row offset
                                         of the matrix fits 16 bits (default: the 16-bit form, 8 bytes per row and SpMV less) */
#define CRBE_SOLVER_VERIFY_AUTO 32u   /* ... only after solves of more than 12 iterations or a restart (the gap between
                                         recurrence and true residual grows with the length of the recurrence) */
#define CRBE_SOLVER_GRAPH 4u          /* replay a step (head kernels, first batch of iterations, state download or end-of-step
                                         record) as one CUDA graph.  Step by step: once the same step shape has been seen twice
                                         (single GPU and peer-memory transport).  Inside the chunks of crbe_solver_steps_ring: on one
                                         GPU, for shapes in steady use, all ring positions at once; on several GPUs never (a capture
                                         on any rank stalls all of them)                                                        */
#define CRBE_SOLVER_EXTRAPOLATE 16u   /* start each step from the polynomial extrapolation of the last q+1 solutions instead of
                                         u^n (q = 1: 2 u^n - u^(n-1)), as far as the history of the time loop reaches    */
#define CRBE_SOLVER_EXTRAP_ORDER(q) (((q) & 7u) << 8)   /* q = 1..4 with CRBE_SOLVER_EXTRAPOLATE; 0 means 1 */
#define CRBE_SOLVER_EXTRAP_ADAPT 128u /* treat the order as an upper bound and pick the order per step from the measured
                                         initial residuals: which order wins depends on how smooth the solution is in
                                         time against the rounding noise of the earlier solves (higher orders amplify it) */
#define CRBE_SOLVER_TMA 8u            /* SpMV-type kernels fed by bulk async copies (cp.async.bulk
                                         + mbarrier pipeline through shared memory)               */
#define CRBE_SOLVER_ILU0 4096u        /* precondition BiCGStab with a multicolour ILU(0) factorisation instead of the diagonal: rows are
                                         coloured and renumbered so that the triangular solves are one data-parallel pass per colour.
                                         For the regimes where the diagonally scaled iteration needs tens to hundreds of iterations
                                         per step (dt D / h^2 >> 1); single GPU; starts every step from u^n or the extrapolated guess */
#define CRBE_SOLVER_NO_PREDICT 2048u  /* always store r and p in the update kernel.  Default: the kernel predicts ||r||^2 =
                                         (s,s) - (t,s)^2/(t,t) from the sums of the preceding kernel and, when that lies clearly
                                         below the stopping threshold, skips the stores (and the read of v) nobody would use:
                                         24 bytes per row less in the last iteration of a solve; same x, same (r,r)            */

/* Build the solver for the pattern (indptr/indices, structural, N rows) with
 * Dirichlet rows bnd_seg_d[0..nb) (crbe.py:397-402). */
int crbe_solver_create(crbe_ctx* ctx, int64_t n, const int32_t* indptr_d, const int32_t* indices_d,
                       int64_t nnz, const int32_t* bnd_seg_d, int64_t nb, crbe_solver** out);
/* Load the system: s_val_d = M + c(K+A) on the structural pattern; m_val_d the
 * mass matrix (its diagonal forms the BE right-hand side, crbe.py:384).
 * rhs_val_d: NULL for Backward Euler, or the values of M - c(K+A) for
 * Crank-Nicolson (crbe.py:386).  Applies the Dirichlet rows, folds the Jacobi
 * preconditioner into the stored rows. */
int crbe_solver_set_system(crbe_solver* s, const double* s_val_d, const double* m_val_d,
                           const double* rhs_val_d);
/* Defaults: rtol 1e-13, 10000 iterations, flags = TMA | VERIFY_AUTO | GRAPH | EXTRAPOLATE | EXTRAP_ORDER(4) | EXTRAP_ADAPT.
 * None of the flags changes the stopping rule ||r|| <= rtol ||b||. */
int crbe_solver_set_options(crbe_solver* s, double rtol, int32_t max_iterations, uint32_t flags);
/* One time step (crbe.py:419-426 without the lift): forms b from u_d (and
 * dt*source_d if not NULL), solves the Dirichlet system, leaves the un-lifted
 * solution in u_d.  The solver keeps copies of the last solutions of the loop
 * (right-hand side, extrapolated initial guess); crbe_solver_set_system starts
 * a new loop. */
int crbe_solver_step(crbe_solver* s, double* u_d, const double* source_d, double dt, crbe_solve_info* info_h);
/* The same step with two alternating vectors (single GPU; both crbe_solver_vector_length long, zero padded):
 * u_cur_d holds u^n and is left intact, the new solution is built in u_next_d (which should still hold u^(n-1)
 * from the call before, for the extrapolated initial guess).  u^n can therefore be downloaded by the host
 * during the whole next step without a staging copy. */
int crbe_solver_step_pingpong(crbe_solver* s, double* u_cur_d, double* u_next_d, const double* source_d, double dt,
                              crbe_solve_info* info_h);
/* The general form: a ring of count = 2..5 such vectors (bufs_h: host array of device pointers).  bufs[cur] holds u^n,
 * the buffers before it (cyclically) the earlier solutions of this time loop; u^(n+1) is built in bufs[(cur+1) % count],
 * which held the oldest one.  Call with cur advancing by one each step; the solver notices a new loop by itself.
 * Nothing is copied, an extrapolation of order q needs count >= q+1, and every solution stays intact for count-1
 * further steps. */
int crbe_solver_step_ring(crbe_solver* s, double* const* bufs_h, int32_t count, int32_t cur, const double* source_d, double dt,
                          crbe_solve_info* info_h);
/* n_steps consecutive steps of the same ring (cur advancing by one per step) with ONE host synchronisation per chunk of steps
 * instead of one per step: the loop of crbe.py:419-429 for the stretches in which the host has nothing to do between steps
 * (constant source_d -- NULL for the stock problem --, no row of `solutions` to fetch).  Convergence is enforced per step on
 * the device; a step that needs more iterations than were enqueued for it stops the chunk there and is finished the ordinary
 * way.  infos_h: n_steps records.  *done_h: steps completed (= n_steps unless an error is returned).  Results are bit-identical
 * to n_steps calls of crbe_solver_step_ring. */
int crbe_solver_steps_ring(crbe_solver* s, double* const* bufs_h, int32_t count, int32_t cur, int32_t n_steps,
                           const double* source_d, double dt, crbe_solve_info* infos_h, int32_t* done_h);
/* Solve  A x = b  for the loaded system (Dirichlet rows applied); x_d holds the initial guess. */
int crbe_solver_solve(crbe_solver* s, const double* b_d, double* x_d, crbe_solve_info* info_h);
/* b of crbe.py:384-402 for inspection: b_d (out). */
int crbe_solver_rhs(crbe_solver* s, const double* u_d, const double* source_d, double dt, double* b_d);
/* Independent check of a finished Backward-Euler step (what crbe.py:426 must reproduce): the residual of u_next in the Dirichlet
 * system whose right-hand side is built from u_prev (crbe.py:384-402),  ||b - A u_next|| / ||b||  in the solver's diagonally
 * scaled norm.  Reads the loaded matrix and the two vectors only; the time loop is not disturbed.  bnorm_h may be NULL. */
int crbe_solver_step_residual(crbe_solver* s, const double* u_prev_d, const double* u_next_d, const double* source_d, double dt,
                              double* relres_h, double* bnorm_h);
/* out = u with out[bnd[k]] += bc[k]   (the lift of crbe.py:429) */
int crbe_solver_lift(crbe_solver* s, const double* u_d, const double* bc_values_d, double* out_d);
/* solutions[step, :] = u_prev + lift (crbe.py:429) written straight into the host's history array, asynchronously on
 * `stream` (a cudaStream_t of the caller, normally NOT the context's stream, so that the transfer overlaps the next
 * step): the N values of u_d go down with one copy-engine transfer, the Nb boundary values bc_values_h go up, and the
 * device stores the lifted boundary entries through the device mapping of row_h.  row_h: page-locked host memory
 * (cudaHostAlloc / cudaHostRegister), N doubles; bc_values_h: page-locked, must stay valid until the stream has
 * passed this call.  Calls on one stream are ordered; do not issue it on two streams at once. */
int crbe_solver_store_lifted_async(crbe_solver* s, const double* u_d, const double* bc_values_h, double* row_h, void* stream);
int crbe_solver_mass_diagonal(crbe_solver* s, const double** mdiag_d_out);   /* diag(M), the weights of crbe_moments */
/* 16 or 32: width of the column indices the SpMV kernels stream for the loaded system and options */
int crbe_solver_index_bits(crbe_solver* s, int32_t* bits_h);
int crbe_solver_destroy(crbe_solver* s);

/* ---- row-block partitioned solve over several GPUs (one process per GPU) -- */
/* The reference is single-process; this is the multi-GPU extension of the same
 * solve.  Rank r owns a contiguous block of DOF rows; NCCL
 * carries the halo DOFs before each SpMV and the allreduce of the dot products. */
typedef struct crbe_comm crbe_comm;
int crbe_comm_unique_id_bytes(void);
int crbe_comm_unique_id(void* id_out_h);                       /* rank 0 creates it, the host broadcasts it */
int crbe_comm_create(crbe_ctx* ctx, int rank, int world, const void* unique_id_h, crbe_comm** out);
int crbe_comm_destroy(crbe_comm* comm);
/* Local rows only.  Column indices are local: [0, n_own) owned, ld + h for halo entry h
 * (ld = n_own rounded up to 256, see crbe_solver_vector_length).  Neighbour q receives the
 * owned entries send_idx_d[send_off[q] .. send_off[q+1]) and delivers recv_counts_h[q]
 * consecutive halo entries.  Vectors passed to crbe_solver_step must have
 * crbe_solver_vector_length doubles. */
int crbe_solver_create_partitioned(crbe_ctx* ctx, crbe_comm* comm, int64_t n_own, int64_t n_halo,
                                   const int32_t* indptr_d, const int32_t* indices_d, int64_t nnz,
                                   const int32_t* bnd_seg_d, int64_t nb, int32_t n_neigh,
                                   const int32_t* neigh_ranks_h, const int64_t* send_counts_h,
                                   const int32_t* send_idx_d, const int64_t* recv_counts_h, crbe_solver** out);
int crbe_solver_vector_length(crbe_solver* s, int64_t* len_h, int64_t* halo_offset_h);
/* Peer-memory transport (default on one NVLink node; replaces the NCCL calls of the partitioned solver): a ring of five
 * solution vectors, the gathered work vectors p, s, r^ and a mailbox live in one CUDA-IPC window per rank.  The halo exchange
 * is fused into the kernels that produce a gathered vector (their last CTA stores the boundary entries straight into the
 * neighbours' halo segments over NVLink and raises an epoch flag; the consuming SpMV walks its strip from the middle and waits
 * only when it reaches a tile that references a halo column), and so is the allreduce of the dot products (the producer's last
 * CTA deposits the partial sums in every rank's mailbox, the kernels that need the totals add the deposits in rank order at
 * their head).  Step 1, every rank: export (64-byte IPC handle; meta_h[2] = {ld, veclen}).  The host gathers handles and metas
 * of all ranks.  Step 2: connect (handles_h = world x 64 bytes in rank order; halo_seg_off_h[q] = offset of this rank's segment
 * inside neighbour q's halo region).  Afterwards the solution vectors handed to crbe_solver_step* must be those of the window
 * ring: crbe_solver_ring (all five, for crbe_solver_step_ring / crbe_solver_steps_ring) or crbe_solver_x (the first, for the
 * in-place crbe_solver_step).  CRBE_P2P_TIMEOUT_MS (environment, default ~18000): how long a kernel spins on a peer's flag
 * before it declares the peer dead; the running step then returns CRBE_ERR_COMM and the solver refuses further steps. */
int crbe_solver_p2p_export(crbe_solver* s, void* ipc_handle_out_h, int64_t* meta_h);
int crbe_solver_p2p_connect(crbe_solver* s, int rank, const void* handles_h, const int64_t* ld_all_h,
                            const int64_t* veclen_all_h, const int64_t* halo_seg_off_h);
int crbe_solver_x(crbe_solver* s, void** x_d_out);
int crbe_solver_ring(crbe_solver* s, void** bufs_out5_h, int32_t* count_h);
int crbe_solver_p2p_error(crbe_solver* s, int* err_h);       /* non-zero: a peer never signalled (1 halo entries, 2 dot products) */

/* ---- measurement -------------------------------------------------------- */
/* Per-kernel device time of the solver kernels, CUDA events on the context stream.
 * enable != 0 resets and starts the accumulation, 0 stops it.  crbe_solver_profile_read
 * fills 8 entries: init, pv, st, xr, p, s, residual, extrapolate: total ms and launch counts
 * (launches enqueued past convergence, which return at once, are not counted). */
int crbe_solver_profile(crbe_solver* s, int enable);
int crbe_solver_profile_read(crbe_solver* s, double* ms_h, int64_t* count_h);
/* Counters since the solver was created, out4_h: [0] update kernels that ran in their short last-iteration form (see
 * CRBE_SOLVER_NO_PREDICT), [1] chunks of steps enqueued with one host synchronisation (crbe_solver_steps_ring), [2] steps that
 * converged inside such chunks, [3] chunks cut short by a step that needed more iterations than were enqueued. */
int crbe_solver_counters(crbe_solver* s, int64_t* out4_h);
/* kernels launched through this context so far */
int crbe_ctx_launch_count(crbe_ctx* ctx, int64_t* count_h);

/* ---- test hooks (used by tests/ only; no reference counterpart) ------------------------------------------ */
/* int32 exclusive scan of the set-up kernels (mesh.cu), exposed for its unit test */
int crbe_test_exclusive_scan(crbe_ctx* ctx, const int32_t* in_d, int32_t* out_d, int64_t n, int64_t* total_h);
/* the solver's row-scaled ELL arrays (device pointers, tile-major layout) so that a test can compare them with scipy */
int crbe_solver_debug_ell(crbe_solver* s, int64_t* ld_h, const int32_t** ell_col_d, const double** ell_val_d,
                          const double** mscale_d, const double** dscale_d);
/* in-place sum of buf_d[0..count) over the ranks of a communicator (NCCL path) */
int crbe_comm_test_allreduce(crbe_comm* comm, double* buf_d, int count);

#ifdef __cplusplus
}
#endif
#endif /* CRBE_B200_H */
