/* devicekmc-b200 — C-ABI of the B200-native field-and-rate hot path of DeviceKMC.
 *
 * Plain C linkage, plain pointers and sizes.  Every `d_*` pointer is DEVICE memory owned by
 * the caller (the reference keeps them in `GPUBuffers`, gpu_buffers.h:16-55); every other
 * pointer is host memory.  All functions return a dkmc_status (0 = OK) and never print;
 * dkmc_last_error() returns the text of the last failure on the calling thread.
 * Work is issued on the context's stream and is COMPLETE on return unless stated otherwise
 * (the reference's callers read results right after each call).
 *
 * Each entry point names the reference interface it replaces (file:line in
 * manasakani/DeviceKMC, src/).  The reference-named `extern "C"` shim that forwards the
 * gpu_solvers.h signatures to these functions is devicekmc_b200/shim/gpu_solvers_shim.cu;
 * INTEGRATION.md shows how a maintainer links it.
 */
#ifndef DKMC_H
#define DKMC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DKMC_VERSION 100

typedef enum {
    DKMC_OK = 0,
    DKMC_ERR_CUDA = 1,          /* a CUDA runtime call or kernel failed */
    DKMC_ERR_ARG = 2,           /* invalid argument */
    DKMC_ERR_NOT_CONVERGED = 3, /* CG hit max_iter (solution still written) */
    DKMC_ERR_RNG_EXHAUSTED = 4, /* event loop consumed all supplied uniforms; call again */
    DKMC_ERR_NO_DEVICE = 5
} dkmc_status;

/* ELEMENT / EVENTTYPE values are the reference's 4-byte unscoped enums, utils.h:37-60 */
enum { DKMC_DEFECT = 0, DKMC_OXYGEN_DEFECT = 1, DKMC_VACANCY = 2, DKMC_O_EL = 3, DKMC_Hf_EL = 4,
       DKMC_Ni_EL = 5, DKMC_Ti_EL = 6, DKMC_Pt_EL = 7, DKMC_N_EL = 8, DKMC_NULL_ELEMENT = 9 };
enum { DKMC_VACANCY_GENERATION = 0, DKMC_VACANCY_RECOMBINATION = 1, DKMC_VACANCY_DIFFUSION = 2,
       DKMC_ION_DIFFUSION = 3, DKMC_NULL_EVENT = 4 };

typedef struct dkmc_ctx dkmc_ctx; /* workspace: stream, scratch arena, cached CSR tiling */

int dkmc_version(void);
const char *dkmc_last_error(void);

/* get_gpu_info / set_gpu, gpu_solvers.h:113-114 (kmc_events.cu:15-32) */
int dkmc_get_gpu_info(char *name, int name_cap, int dev);
int dkmc_set_gpu(int dev);
int dkmc_device_count(int *count);

/* Workspace bound to the current device.  The reference mallocs/frees scratch inside every
 * call (potential_solver_gpu.cu:397-493,735-779; kmc_events.cu:163-164,362); here it is a
 * persistent arena that grows on demand. */
int dkmc_ctx_create(dkmc_ctx **ctx);
int dkmc_ctx_destroy(dkmc_ctx *ctx);
int dkmc_ctx_set_stream(dkmc_ctx *ctx, void *cuda_stream); /* cudaStream_t; default 0 */
int dkmc_ctx_synchronize(dkmc_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py "gpu_launches") */
int dkmc_ctx_launch_count(dkmc_ctx *ctx, long long *count);

/* copytoConstMemory, gpu_solvers.h:205 (kmc_events.cu:369-375): per-layer zero-field
 * activation energies used by the rate table.  n_layers <= 16. */
int dkmc_set_layer_energies(dkmc_ctx *ctx, int n_layers, const double *E_gen, const double *E_rec,
                            const double *E_Vdiff, const double *E_Odiff);

/* ---- a1: neighbour graph.  Device::constructSiteNeighborList, Device.cpp:98-136,175-199 and
 * the padded table Device.cpp:68-80 (rows ascending in j, padded with -1), via a cell list.
 * `lattice` is a HOST double[3].  count: writes the max degree (Device::max_num_neighbors);
 * fill: writes d_neigh_idx[N*nn].  fill must follow count on the same positions. */
int dkmc_neighbor_count(dkmc_ctx *ctx, int N, const double *d_x, const double *d_y, const double *d_z,
                        const double *lattice, int pbc, double nn_dist, int *max_nn);
int dkmc_neighbor_fill(dkmc_ctx *ctx, int N, const double *d_x, const double *d_y, const double *d_z,
                       const double *lattice, int pbc, double nn_dist, int nn, int *d_neigh_idx);

/* the same builder for HOST arrays (h_x/h_y/h_z in, h_neigh_idx[N * *max_nn] out), so
 * that a host class which keeps both in host memory — the reference's Device — can replace its
 * O(N^2) loop (Device.cpp:98-136) with one call.  h_neigh_idx == NULL: only *max_nn is written; the
 * caller then sizes the table and calls again.  Shim: devicekmc_b200/shim/device_setup_shim.cpp. */
int dkmc_neighbor_table_host(dkmc_ctx *ctx, int N, const double *h_x, const double *h_y, const double *h_z,
                             const double *lattice, int pbc, double nn_dist, int *max_nn, int *h_neigh_idx);

/* ---- a2: CSR structure of K.  initialize_sparsity, gpu_solvers.h:43
 * (iterative_solvers_gpu.cu:96-109 -> Assemble_K_sparsity :2158-2208).  The arrays are
 * allocated by the library (as the reference's `int **` out-parameters are) and have the
 * GPUBuffers meaning: interior block m x m with the diagonal, interior-relative ascending
 * columns; contact blocks with contact-relative columns.  m = N - NL - NR. */
typedef struct {
    int m, nnz, left_nnz, right_nnz;
    int *d_row_ptr, *d_col;             /* GPUBuffers::Device_row_ptr_d / Device_col_indices_d */
    int *d_left_row_ptr, *d_left_col;   /* contact_left_row_ptr / contact_left_col_indices */
    int *d_right_row_ptr, *d_right_col; /* contact_right_row_ptr / contact_right_col_indices */
} dkmc_sparsity;
int dkmc_initialize_sparsity(dkmc_ctx *ctx, int N, int nn, const int *d_neigh_idx, int NL, int NR,
                             dkmc_sparsity *out);
int dkmc_free_sparsity(dkmc_ctx *ctx, dkmc_sparsity *sp);

/* ---- a3: charge state machine.  update_charge_gpu, gpu_solvers.h:127-130
 * (CPU semantics potential_solver.cpp:172-217).  Asynchronous like the reference's. */
int dkmc_update_charge(dkmc_ctx *ctx, const int *d_site_element, int *d_site_charge,
                       const int *d_neigh_idx, int N, int nn, const int *d_metals, int num_metals);

/* ---- a4 + a5: background potential.  background_potential_gpu_sparse, gpu_solvers.h:139-141
 * (potential_solver_gpu.cu:696-781; CPU semantics potential_solver.cpp:289-410).
 * Assembles K on the neighbour graph, solves the interior system with preconditioned CG (Jacobi + one
 * coarse unknown per uncharged-vacancy cluster) warm-started from d_site_potential_boundary, refines with
 * a double-double residual and writes the Dirichlet contacts (-Vd/2, +Vd/2).  Each CG run is ONE persistent
 * kernel (grid barriers between the vector and the SpMV phase, one reduction per iteration, scalars in
 * registers; csrc/pcg_persistent.cuh) — semantically the loop of iterative_solvers_gpu.cu:424-455. */
typedef struct {
    double rel_tol;      /* CG stop: ||r||_D^-1 <= rel_tol * ||b||_D^-1   (default 1e-12) */
    int max_iter;        /* per CG run (default 20000) */
    int refine_rounds;   /* max restarts on the double-double residual (default 4) */
    int check_every;     /* fallback path only (DKMC_LEGACY_CG=1 / no peer access): iterations between host
                            convergence polls (default 32); the persistent-kernel PCG decides on the device */
    int cluster_precond; /* 1 (default): add one coarse unknown per uncharged-vacancy cluster to the
                            Jacobi preconditioner; 0: plain Jacobi as the reference */
    double refine_tol;   /* relative tolerance of each restart's correction solve (default 1e-6) */
    double est_tol;      /* restart while max|M^-1 r_true| / max|x| exceeds this (default 1e-13) */
} dkmc_solver_opts;
typedef struct {
    int iterations;        /* total CG iterations incl. refinement runs */
    int refinements;       /* refinement rounds executed */
    double rel_residual;   /* final double-double ||b - A x||_2 / ||b||_2 */
    double assemble_ms, solve_ms; /* device time of the two phases */
    double est_error;      /* max|M^-1 (b - A x)| / max|x| of the returned solution */
} dkmc_solve_info;
void dkmc_default_solver_opts(dkmc_solver_opts *o);
int dkmc_background_potential_sparse(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int nn,
                                     const int *d_neigh_idx, int NL, int NR, double Vd,
                                     double high_G, double low_G, const int *d_site_element,
                                     const int *d_site_charge, const int *d_metals, int num_metals,
                                     double *d_site_potential_boundary,
                                     const dkmc_solver_opts *opts, dkmc_solve_info *info);

/* The context caches a few things keyed by the ADDRESS of the caller's arrays: the SpMV tiling (by d_row_ptr, m,
 * nnz), the cell grid of the pairwise sum (by d_x, d_sigma, N: positions and sigma are static in the reference),
 * the previous charges of the opt-in incremental pairwise update (by d_site_charge / d_site_potential_charge).
 * A caller that changes such an array IN PLACE (another structure at the same address, positions moved, phi_c
 * overwritten by a restart) calls dkmc_ctx_invalidate: the next calls rebuild them.  Matrix VALUES are never
 * cached: dkmc_spmv / dkmc_solve_cg read the val array they are given. */
int dkmc_ctx_invalidate(dkmc_ctx *ctx);

/* Internal row order of the solver (optional; default: the caller's order).  The public arrays keep the
 * caller's site order — the reference's puts all lattice atoms before all interstitials
 * (reorder_boundary.py:113-124), so the CSR bandwidth is ~0.7 N and an index range is not a spatial slab.
 * d_order[p] = the interior row (0 .. m-1, caller's numbering) that the solver puts at position p, e.g. the rows
 * sorted x-major by grid cell (device array of sp->m ints, copied).  a4/a5/8f-3 then solve P K P^T: the structure
 * is permuted once, the values (assembled in the caller's order: the diagonal keeps the reference's summation
 * order), right-hand side and starting vector are gathered every step and the solution is scattered back.
 * d_order = NULL removes the order.  dkmc_solver_csr: the CSR structure the solver really works on (for the
 * slab plan of dkmc_dist_background_potential, whose row ranges then count internal rows). */
int dkmc_solver_set_order(dkmc_ctx *ctx, const dkmc_sparsity *sp, const int *d_order);
int dkmc_solver_csr(dkmc_ctx *ctx, const dkmc_sparsity *sp, const int **d_row_ptr, const int **d_col,
                    const int **d_order);

/* ---- conduction-band edge.  update_CB_edge_gpu_sparse, gpu_solvers.h:121-123
 * (potential_solver_gpu.cu:595-694; CPU semantics Device::setLaplacePotential,
 * potential_solver.cpp:4-139): the same Kirchhoff system with the rule "high_G iff either site is a
 * metal", contacts at +q Vd / 2 (first NL sites) and -q Vd / 2 (last NR sites); called once per bias
 * point.  Same assembly kernel and refined PCG (plain Jacobi: no vacancy clusters under this rule),
 * warm-started from d_site_CB_edge. */
int dkmc_update_CB_edge_sparse(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double Vd,
                               double q, double high_G, double low_G, const int *d_site_element,
                               const int *d_metals, int num_metals, double *d_site_CB_edge,
                               const dkmc_solver_opts *opts, dkmc_solve_info *info);

/* building blocks of a4/a5, exported for parity tests and the roofline measurements */
int dkmc_assemble_K(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double Vd,
                    double high_G, double low_G, const int *d_site_element, const int *d_site_charge,
                    const int *d_metals, int num_metals, double *d_val, double *d_rhs);
/* y = A x, CSR FP64/int32 (replaces cusparseSpMV, iterative_solvers_gpu.cu:411,428) */
int dkmc_spmv(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col,
              const double *d_val, const double *d_x, double *d_y);
int dkmc_solve_cg(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col,
                  const double *d_val, const double *d_rhs, double *d_x,
                  const dkmc_solver_opts *opts, dkmc_solve_info *info);

/* ---- a6: pairwise Coulomb sum.  poisson_gridless_gpu, gpu_solvers.h:144-147
 * (potential_solver_gpu.cu:908-978; CPU semantics potential_solver.cpp:412-432, utils.h:102).
 * d_lattice/d_sigma/d_k are device scalars/arrays exactly as the reference passes them. */
int dkmc_poisson_gridless(dkmc_ctx *ctx, int pbc, int N, const double *d_lattice,
                          const double *d_sigma, const double *d_k, const double *d_x,
                          const double *d_y, const double *d_z, const int *d_site_charge,
                          double *d_site_potential_charge);
/* same sum restricted to the target range [row_begin,row_end) (slab-partitioned ranks) */
int dkmc_poisson_gridless_rows(dkmc_ctx *ctx, int pbc, int N, const double *d_lattice,
                               const double *d_sigma, const double *d_k, const double *d_x,
                               const double *d_y, const double *d_z, const int *d_site_charge,
                               int row_begin, int row_end, double *d_site_potential_charge);

/* Device::updatePotential (potential_solver.cpp:249-260) calls the CG and then the pairwise sum; they
 * are independent (both read only site_charge) and bound by different units (HBM vs the FP64 pipe).
 * _begin compacts the charged sites on the context's stream, then launches the sum on the context's
 * SIDE stream with a bounded residency per SM and returns at once; work issued afterwards on the
 * context's stream (dkmc_background_potential_sparse) runs concurrently with it.  _join makes the
 * context's stream wait for the sum, blocks until it is complete and reports its device time.
 * A dkmc_poisson_gridless(_rows) call with the same arguments as a pending _begin joins it instead
 * of recomputing — that is how the reference-named shim overlaps the two without touching
 * Device::updatePotential.  The inputs must not change between _begin and _join. */
int dkmc_poisson_gridless_begin(dkmc_ctx *ctx, int pbc, int N, const double *d_lattice,
                                const double *d_sigma, const double *d_k, const double *d_x,
                                const double *d_y, const double *d_z, const int *d_site_charge,
                                int row_begin, int row_end, double *d_site_potential_charge);
int dkmc_poisson_gridless_join(dkmc_ctx *ctx, double *pairwise_ms);
/* The erfc factor of a6 is exactly 0 in double precision beyond r = 26.45 sigma sqrt 2 (131 A at the
 * shipped sigma): by default (on = 1) the charged sites are binned into a cell grid every step and a
 * warp only visits the cells that come within that distance of its targets — the sum is unchanged,
 * at 1 M sites 56 % of the pairs are skipped.  Periodic devices (pbc = 1) always take the all-pairs
 * kernel.  dkmc_pairwise_pairs_evaluated: pairs the last cell-list sum evaluated (-1: none yet). */
int dkmc_ctx_set_pairwise_cells(dkmc_ctx *ctx, int on);
/* OPT-IN approximation, default 0 = off: truncate the sum at cutoff_sigmas * sigma.
 * erfc(r / (sigma sqrt 2)) < 1.6e-23 beyond 10 sigma, so phi_c changes by < 1e-21 of its largest
 * entries (far below the 1e-10 parity bound; sites whose every charge is farther away get exactly 0
 * instead of ~1e-25 V).  This changes the WORK of a6 (70x fewer pairs at 1 M sites); bench.py reports
 * it as a separate variant, never as the headline. */
int dkmc_ctx_set_pairwise_cutoff(dkmc_ctx *ctx, double cutoff_sigmas);
int dkmc_pairwise_pairs_evaluated(dkmc_ctx *ctx, long long *pairs);
/* Far field of the cell-list kernel (default on = 1).  A run of sources that lies at least 4 sigma sqrt 2
 * (19.8 A at the shipped sigma) from every target of a warp is evaluated as
 * erfc(t) / r = exp(-t^2) F(1/t^2) / (c sqrt(pi) r^2), t = c r — the same function (F by a polynomial fitted
 * to 1e-14 relative, tools/fit_erfcx.py) without the square root and the second division: ~40 instead of
 * ~62 FP64 instructions per pair.  on = 0: every pair takes the general formula (test hook).
 * dkmc_pairwise_pairs_far: pairs the last cell-list sum evaluated by the far-field formula (-1: none yet). */
int dkmc_ctx_set_pairwise_far_field(dkmc_ctx *ctx, int on);
int dkmc_pairwise_pairs_far(dkmc_ctx *ctx, long long *pairs);
/* OPT-IN incremental update, default 0 = off.  Between two KMC steps only the sites
 * touched by executed events change charge, so phi_c(new) = phi_c(old) + sum over the CHANGED sites of
 * (q_new - q_old) * kernel: O(N * n_changed) instead of O(N * n_charged).  With refresh_every = R > 0
 * a call on the same (d_site_charge, d_site_potential_charge, N, rows, pbc) as the previous one sums
 * only the charge differences into d_site_potential_charge — which must still hold the previous
 * result — and every R-th call is a full sum, bounding the rounding drift at ~R ulp of the largest
 * entries (norm-wise; an entry whose sources all left keeps a residue of that size instead of its
 * exact tiny value).  Like the cutoff this changes the WORK of a6: a bench variant, not the headline.
 * dkmc_pairwise_incremental_counts: how many full / difference sums have run. */
int dkmc_ctx_set_pairwise_incremental(dkmc_ctx *ctx, int refresh_every);
int dkmc_pairwise_incremental_counts(dkmc_ctx *ctx, long long *full_sums, long long *delta_sums);
/* share of each SM the overlapped pairwise kernel may occupy: CTAs per SM x threads per CTA */
int dkmc_ctx_set_pairwise_share(dkmc_ctx *ctx, int blocks_per_sm, int threads_per_block);

/* ---- a7: rate table.  build_event_list, kmc_events.cu:34-126 with the CPU semantics of
 * KMCProcess::update_events_and_rates, KMCProcess.cpp:67-164 (vacancy-diffusion barrier from
 * layer[i]).  d_event_type int32[N*nn], d_event_prob double[N*nn]. */
int dkmc_build_event_list(dkmc_ctx *ctx, int N, int nn, const int *d_neigh_idx, const int *d_site_layer,
                          const double *d_lattice, int pbc, const double *d_T_bg, const double *d_freq,
                          const double *d_sigma, const double *d_k, const double *d_x, const double *d_y,
                          const double *d_z, const double *d_potential_boundary,
                          const double *d_potential_charge, const int *d_site_element,
                          const int *d_site_charge, int *d_event_type, double *d_event_prob);

/* inclusive prefix sum (utils.h:91-99 / thrust::inclusive_scan kmc_events.cu:214), parallel
 * block-level scan; and upper_bound selection (KMCProcess.cpp:311 / kmc_events.cu:222):
 * first idx with cum[idx] > u * cum[n-1].  Exported for parity tests and the scan roofline. */
int dkmc_inclusive_scan(dkmc_ctx *ctx, long long n, const double *d_in, double *d_out);
int dkmc_select_event(dkmc_ctx *ctx, long long n, const double *d_cum, double u, long long *idx,
                      double *Psum);

/* ---- a7 + a8: one KMC step.  execute_kmc_step_gpu, gpu_solvers.h:196-201
 * (kmc_events.cu:146-365; CPU semantics KMCProcess.cpp:282-365).  Builds the rate table, then
 * runs the residence-time loop entirely on the device: select (prefix sums + search), execute,
 * zero the conflicting events, draw the residence time — until event_time >= 1/freq.
 * `uniforms` is the HOST array of the next n_uniforms numbers of the reference's
 * RandomNumberGenerator stream (random_num.h:4-23), two per executed event; *n_used tells the
 * caller how far to advance its generator.  Returns DKMC_ERR_RNG_EXHAUSTED if the loop needs
 * more numbers: call dkmc_kmc_step_continue with the following numbers of the stream.
 * events_out (host, may be NULL): (table idx, i, j, type) per executed event, up to max_events. */
typedef struct {
    int n_events;          /* events executed in this step */
    int n_used;            /* uniforms consumed */
    int n_exact_fallbacks; /* selections decided by the exact sequential replay */
    double event_time;     /* the last residence-time draw (what the reference returns): -ln(u2) / Psum with the
                              hierarchical Psum (the reference's is the sequential sum: equal to ~1e-15 relative) */
    double rate_ms, loop_ms;
} dkmc_step_info;
int dkmc_execute_kmc_step(dkmc_ctx *ctx, int N, int nn, const int *d_neigh_idx, const int *d_site_layer,
                          const double *d_lattice, int pbc, const double *d_T_bg, const double *d_freq,
                          const double *d_sigma, const double *d_k, const double *d_x, const double *d_y,
                          const double *d_z, const double *d_potential_boundary,
                          const double *d_potential_charge, int *d_site_element, int *d_site_charge,
                          const double *uniforms, int n_uniforms, int *events_out, int max_events,
                          dkmc_step_info *info);
int dkmc_kmc_step_continue(dkmc_ctx *ctx, const double *uniforms, int n_uniforms, int *events_out,
                           int max_events, dkmc_step_info *info);
/* test hook: force every selection through the exact sequential replay (1) or never (0, default) */
int dkmc_ctx_set_exact_select(dkmc_ctx *ctx, int mode);
/* device pointers of the last step's event tables (valid until the next step) */
int dkmc_last_event_tables(dkmc_ctx *ctx, const int **d_event_type, const double **d_event_prob);

/* ---- multi-GPU (no reference counterpart: the reference is single-GPU).
 * One process per GPU.  The interior rows of K are split into contiguous ranges of whole SpMV
 * tiles (an x-slab for x-major ordered sites); the PCG of a4/a5 then exchanges the halo of p
 * with ncclSend/ncclRecv and all-reduces [p.Ap] and [r.D^-1 r, cluster sums] per iteration.
 * Bootstrap: rank 0 calls dkmc_dist_unique_id, the host broadcasts the 128 bytes (e.g. through
 * torch.distributed), every rank calls dkmc_dist_init. */
#define DKMC_MAX_RANKS 64
#define DKMC_MAX_HALO_SEGMENTS 64
typedef struct {
    int world;
    int row_begin[DKMC_MAX_RANKS], row_end[DKMC_MAX_RANKS]; /* interior rows of every rank */
    int n_send, n_recv;                                      /* this rank's halo segments [begin,end) */
    int send_peer[DKMC_MAX_HALO_SEGMENTS], send_begin[DKMC_MAX_HALO_SEGMENTS], send_end[DKMC_MAX_HALO_SEGMENTS];
    int recv_peer[DKMC_MAX_HALO_SEGMENTS], recv_begin[DKMC_MAX_HALO_SEGMENTS], recv_end[DKMC_MAX_HALO_SEGMENTS];
} dkmc_dist_plan;
int dkmc_spmv_tile_nnz(void); /* rows are assigned to SpMV tiles by row start, this many nnz per tile */
int dkmc_dist_unique_id(char *id128);
int dkmc_dist_init(dkmc_ctx *ctx, int rank, int world, const char *id128);
int dkmc_dist_finalize(dkmc_ctx *ctx);
/* Optional: peer-memory windows for the per-iteration exchange of the distributed PCG (halo of the
 * search direction and the two all-reduces) — NVLink peer stores and flags inside small kernels
 * instead of three NCCL calls per iteration.  Every rank calls _alloc (m = interior rows; returns a
 * 64-byte CUDA IPC handle), the host all-gathers the handles in rank order, every rank calls _open
 * with the world * 64 bytes.  Without it dkmc_dist_background_potential uses NCCL throughout. */
int dkmc_dist_p2p_alloc(dkmc_ctx *ctx, int m, char *ipc_handle64);
int dkmc_dist_p2p_open(dkmc_ctx *ctx, const char *ipc_handles64);
/* all-gather of a caller array split by rows over the ranks (row_begin/row_end: HOST int[world]; rank r
 * holds rows [row_begin[r], row_end[r]) of d_buf on entry, every rank holds all of them on return):
 * through the peer windows when they are open and hold n doubles, else NCCL broadcasts.
 * Asynchronous on the context's stream. */
int dkmc_dist_allgather_rows(dkmc_ctx *ctx, double *d_buf, int n, const int *row_begin, const int *row_end);
/* background_potential_gpu_sparse (gpu_solvers.h:139-141) over `world` GPUs; on return every rank
 * holds the full d_site_potential_boundary. */
int dkmc_dist_background_potential(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double Vd,
                                   double high_G, double low_G, const int *d_site_element,
                                   const int *d_site_charge, const int *d_metals, int num_metals,
                                   double *d_site_potential_boundary, const dkmc_dist_plan *plan,
                                   const dkmc_solver_opts *opts, dkmc_solve_info *info);

/* on = 1: run the CG as one kernel per operation with host convergence polls (the round-1 path, kept as the
 * A/B baseline and as the fallback without peer access); default 0 (env DKMC_LEGACY_CG=1 sets it at create). */
int dkmc_ctx_set_legacy_cg(dkmc_ctx *ctx, int on);
/* The persistent-kernel CG comes in two recurrences.  on = 1 (the default from four GPUs on, for the restarts' correction solves): pipelined CG — both inner products of an
 * iteration are taken before its matrix product, so their (cross-GPU) reduction travels while the product runs
 * and an iteration has ONE synchronisation point.  on = 0: Chronopoulos-Gear CG, two synchronisation points per
 * iteration (the default on one GPU, where there is no NVLink round trip to hide; also the automatic fall-back
 * when one CTA's rows touch more than 24 uncharged-vacancy clusters or the cluster coarse space is off).
 * Both solve a5's system to the same refined accuracy (env DKMC_PCG_PIPELINED sets it at create). */
int dkmc_ctx_set_pcg_pipelined(dkmc_ctx *ctx, int on);

/* Measurement aid: with DKMC_PCG_PROF=1 in the environment the persistent PCG accumulates, in CTA 0, the
 * nanoseconds spent per phase: out[0] set-up of a solve, [1] vector phase, [2] barrier + halo wait, [3] SpMV
 * tiles, [4] cluster rows, [5] barrier + reduction wait; out[6] = iterations, out[7] = solves since the last
 * call (which resets the counters).  All zero when profiling is off. */
int dkmc_pcg_profile(dkmc_ctx *ctx, double *out8);

/* Measurement aid (no reference counterpart): achievable FP64 FMA throughput of this GPU in
 * TFLOP/s (8 independent DFMA chains per thread on every SM) — the roofline denominator of the
 * pairwise sum, which MEASURED_PEAKS.json does not carry. */
int dkmc_probe_fp64_tflops(dkmc_ctx *ctx, double *tflops);

/* ---- snapshots that do not stall the step.  Device::writeSnapshot (Device.cpp:236-252)
 * prints per site: element, position, potential_boundary + potential_charge, power — after a blocking
 * sync_GPUToHost of all seven site arrays (kmc_main.cpp:196-201, gpu_buffers.cpp:39-55).  begin: one
 * fused pass on the context's stream stages element, charge, the summed potential and (if given) the
 * power; a separate copy stream then moves the staged arrays into the caller's HOST buffers (page-
 * locked memory makes the copies truly asynchronous) while the caller goes on — the event loop may
 * mutate site_element / site_charge right away.  wait: blocks until the host buffers are complete;
 * ready: non-blocking query.  One snapshot in flight per context.  d_site_power / h_power may both be
 * NULL.  Multi-GPU: the site state is replicated, so rank 0 alone snapshots. */
int dkmc_snapshot_begin(dkmc_ctx *ctx, int N, const int *d_site_element, const int *d_site_charge,
                        const double *d_site_potential_boundary, const double *d_site_potential_charge,
                        const double *d_site_power, int *h_element, int *h_charge, double *h_potential,
                        double *h_power);
int dkmc_snapshot_ready(dkmc_ctx *ctx, int *ready);
int dkmc_snapshot_wait(dkmc_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* DKMC_H */
