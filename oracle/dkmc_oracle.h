/* TEST INFRASTRUCTURE ONLY — CPU restatement of the DeviceKMC field-and-rate hot path.
 *
 * This is the parity oracle for the CUDA path in devicekmc_b200/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path never links, imports or calls anything in oracle/.
 *
 * Pinning: every function below is checked against the UNMODIFIED reference CPU build
 * (oracle/_ref/libdkmc_ref.so, see oracle/Makefile + ref_harness.cpp) in
 * tests/test_oracle_vs_reference.py (runs where /root/reference exists) and against the
 * golden fixtures generated from that build (tests/golden/, script tests/golden/make_golden.py).
 * The reference ships no known-answer tests for this path.
 *
 * All `file:line` citations are relative to the reference tree (manasakani/DeviceKMC).
 */
#ifndef DKMC_ORACLE_H
#define DKMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ELEMENT / EVENTTYPE enum values, utils.h:37-60 */
enum { ORC_DEFECT = 0, ORC_OXYGEN_DEFECT = 1, ORC_VACANCY = 2, ORC_O_EL = 3, ORC_Hf_EL = 4,
       ORC_Ni_EL = 5, ORC_Ti_EL = 6, ORC_Pt_EL = 7, ORC_N_EL = 8, ORC_NULL_ELEMENT = 9 };
enum { ORC_VACANCY_GENERATION = 0, ORC_VACANCY_RECOMBINATION = 1, ORC_VACANCY_DIFFUSION = 2,
       ORC_ION_DIFFUSION = 3, ORC_NULL_EVENT = 4 };

/* site_dist, utils.cpp:100-137 (pbc only in y,z) */
double orc_site_dist(double x1, double y1, double z1, double x2, double y2, double z2,
                     const double *lattice, int pbc);

/* Device::constructSiteNeighborList / is_neighbor, Device.cpp:98-136,175-199.
 * method 0 = O(N^2) brute force exactly as the reference, 1 = cell list (same predicate).
 * Returns max degree (= Device::max_num_neighbors); deg[i] = #neighbours of i. */
int orc_neighbor_degrees(int N, const double *x, const double *y, const double *z,
                         const double *lattice, int pbc, double nn_dist, int *deg, int method);
/* padded neighbour table, Device.cpp:68-80: row i = ascending j, then -1 */
void orc_neighbor_fill(int N, const double *x, const double *y, const double *z,
                       const double *lattice, int pbc, double nn_dist, int nn, int *neigh_idx,
                       int method);

/* site -> layer id, KMCProcess.cpp:34-50 (last matching layer wins; -1 if outside) */
void orc_site_layers(int N, const double *x, int n_layers, const double *start_x,
                     const double *end_x, int *site_layer);

/* Device::updateCharge CPU branch, potential_solver.cpp:172-217 */
void orc_update_charge(int N, int nn, const int *neigh_idx, const int *element,
                       const int *metals, int num_metals, int *charge);

/* CSR structure of K = [left | interior | right] restricted to interior rows
 * (initialize_sparsity → Assemble_K_sparsity, iterative_solvers_gpu.cu:96-109,2158-2208;
 * interior block includes the diagonal, columns ascending, interior-relative indices;
 * contact blocks use contact-relative indices).  Pass 1 fills the row pointers. */
void orc_csr_row_ptr(int N, int nn, const int *neigh_idx, int NL, int NR, int *row_ptr,
                     int *left_row_ptr, int *right_row_ptr);
void orc_csr_fill(int N, int nn, const int *neigh_idx, int NL, int NR, const int *row_ptr,
                  int *col, const int *left_row_ptr, int *left_col, const int *right_row_ptr,
                  int *right_col);

/* K values + rhs following Device::background_potential, potential_solver.cpp:289-372
 * (same conductance rule, diagonal = sequential sum over ALL neighbours in ascending j,
 * rhs = -(K_left VL + K_right VR) accumulated in ascending j).  System: A x = rhs, x = phi_int. */
void orc_assemble_K(int N, int nn, const int *neigh_idx, int NL, int NR, const int *element,
                    const int *charge, const int *metals, int num_metals, double high_G,
                    double low_G, double Vd, const int *row_ptr, const int *col, double *val,
                    double *rhs);

/* Oracle's own solver for A x = rhs: symmetric-Jacobi-scaled CG with the recurrences of
 * solve_sparse_CG_Jacobi (iterative_solvers_gpu.cu:309-480) run to `tol` on ||r||/||b||,
 * followed by `refine` rounds of iterative refinement whose residual is accumulated in
 * __float128.  (The reference CPU path uses dense dgesv, potential_solver.cpp:379, which is
 * infeasible beyond N ~ 3e4 and is itself ~2e-9 from the exact solution; see DESIGN.md.)
 * x: in = initial guess, out = solution.  info[0]=iterations, info[1]=final scaled ||r||/||b||,
 * info[2]=quad-precision ||b-Ax||_inf after the last refinement. */
int orc_solve(int m, const int *row_ptr, const int *col, const double *val, const double *rhs,
              double *x, double tol, int max_iter, int refine, double *info);

/* convenience: assemble + solve + scatter with Dirichlet contacts, potential_solver.cpp:389-403 */
void orc_laplace_cb_edge(int N, int nn, const int *neigh_idx, int NL, int NR, const int *element,
                         const int *metals, int num_metals, double high_G, double low_G, double Vd,
                         double q, double *site_CB_edge, double tol, int max_iter, int refine,
                         double *info);
void orc_background_potential(int N, int nn, const int *neigh_idx, int NL, int NR,
                              const int *element, const int *charge, const int *metals,
                              int num_metals, double high_G, double low_G, double Vd,
                              double *site_potential_boundary, double tol, int max_iter,
                              int refine, double *info);

/* Device::poisson_gridless + v_solve, potential_solver.cpp:412-432, utils.h:102 */
void orc_poisson_gridless(int N, const double *x, const double *y, const double *z,
                          const double *lattice, int pbc, const int *charge, double sigma,
                          double k, double *site_potential_charge);
/* same, for a subset of target rows (bounded CPU-baseline samples) */
void orc_poisson_gridless_rows(int N, const double *x, const double *y, const double *z,
                               const double *lattice, int pbc, const int *charge, double sigma,
                               double k, int row_begin, int row_end, double *out_rows);

/* KMCProcess::update_events_and_rates, KMCProcess.cpp:67-164.  E tables indexed by layer. */
void orc_rate_table(int N, int nn, const int *neigh_idx, const int *site_layer,
                    const double *lattice, int pbc, double T_bg, double freq, double sigma,
                    double k, const double *x, const double *y, const double *z,
                    const double *pot_boundary, const double *pot_charge, const int *element,
                    const int *charge, const double *E_gen, const double *E_rec,
                    const double *E_Vdiff, const double *E_Odiff, int *event_type,
                    double *event_prob);

/* RandomNumberGenerator, random_num.h:4-23: std::mt19937 + uniform_real_distribution<double>(0,1)
 * as implemented by libstdc++ (generate_canonical<double,53>: two 32-bit draws per double). */
typedef struct { uint32_t mt[624]; int idx; } orc_rng;
void orc_rng_seed(orc_rng *r, uint32_t seed);
double orc_rng_uniform(orc_rng *r);

/* event loop of KMCProcess::executeKMCStep CPU branch, KMCProcess.cpp:297-358:
 * sequential inclusive_prefix_sum (utils.h:91-99), std::upper_bound, execute_event
 * (KMCProcess.cpp:187-256), conflict zeroing (330-352), event_time = -log(u)/Psum.
 * events[4*e + {0,1,2,3}] = table idx, i, j, type.  Returns #events (may exceed max_events;
 * only the first max_events are recorded).  event_type/event_prob are modified in place. */
void orc_set_event_limit(int n);
int orc_kmc_events(int N, int nn, const int *neigh_idx, int *event_type, double *event_prob,
                   int *element, int *charge, double freq, orc_rng *rng, double *event_time,
                   int *events, int max_events);

/* select only (one draw): first idx with cum[idx] > u*Psum on the sequential prefix sum.
 * Returns idx (== n if none), writes Psum. */
long orc_select_event(long n, const double *event_prob, double u, double *Psum);

#ifdef __cplusplus
}
#endif
#endif
