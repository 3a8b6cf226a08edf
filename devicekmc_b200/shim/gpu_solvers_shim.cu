// devicekmc-b200 — reference-named entry points for the field-and-rate hot path.
//
// This file is what a DeviceKMC maintainer compiles INSTEAD of the reference's
// potential_solver_gpu.cu / iterative_solvers_gpu.cu / kmc_events.cu.  It includes the
// reference's own headers (src/gpu_solvers.h, src/gpu_buffers.h), implements the `extern "C"`
// functions declared there for this path with their exact signatures, and forwards to the C-ABI
// of include/dkmc.h.  The reference's kmc_main.cpp, Device.cpp, KMCProcess.cpp,
// potential_solver.cpp and gpu_buffers.cpp then link against it unchanged.
//
//   nvcc -I<DeviceKMC>/src -I<devicekmc-b200>/include -c gpu_solvers_shim.cu
// (like the reference's own .cu files it is compiled WITHOUT -DUSE_CUDA: only the .cpp files get it)
//
// Conventions kept from the reference: every pointer argument is device memory
// owned by GPUBuffers; results are complete on return (update_charge_gpu stays asynchronous);
// errors are reported on stderr and execution continues (utils.h:145-153); the cuBLAS/cuSOLVER
// handles are accepted and ignored.
#include "gpu_solvers.h"

#include <cmath>
#include <algorithm>
#include <cstdio>
#include <vector>

#include "dkmc.h"

namespace {

dkmc_ctx *g_ctx = nullptr;
int g_n_layers = 0;
std::vector<double> g_E[4];

dkmc_ctx *ctx() {
    if (!g_ctx) {
        if (dkmc_ctx_create(&g_ctx) != DKMC_OK) {
            fprintf(stderr, "devicekmc-b200: %s\n", dkmc_last_error());
            g_ctx = nullptr;
        } else if (g_n_layers > 0) {
            dkmc_set_layer_energies(g_ctx, g_n_layers, g_E[0].data(), g_E[1].data(), g_E[2].data(), g_E[3].data());
        }
    }
    return g_ctx;
}

void report(int status, const char *where) {
    if (status != DKMC_OK) fprintf(stderr, "devicekmc-b200: %s: status %d: %s\n", where, status, dkmc_last_error());
}

dkmc_sparsity sparsity_of(const GPUBuffers &g, int m) {
    dkmc_sparsity sp;
    sp.m = m;
    sp.nnz = g.Device_nnz; sp.left_nnz = g.contact_left_nnz; sp.right_nnz = g.contact_right_nnz;
    sp.d_row_ptr = g.Device_row_ptr_d; sp.d_col = g.Device_col_indices_d;
    sp.d_left_row_ptr = g.contact_left_row_ptr; sp.d_left_col = g.contact_left_col_indices;
    sp.d_right_row_ptr = g.contact_right_row_ptr; sp.d_right_col = g.contact_right_col_indices;
    return sp;
}

}  // namespace

extern "C" {

void get_gpu_info(char *gpu_string, int dev) {  // gpu_solvers.h:113
    report(dkmc_get_gpu_info(gpu_string, 1000, dev), "get_gpu_info");
}

void set_gpu(int dev) { report(dkmc_set_gpu(dev), "set_gpu"); }  // gpu_solvers.h:114

// gpu_solvers.h:205 — called from the GPUBuffers constructor, before any context exists
void copytoConstMemory(std::vector<double> E_gen, std::vector<double> E_rec, std::vector<double> E_Vdiff,
                       std::vector<double> E_Odiff) {
    g_E[0] = E_gen; g_E[1] = E_rec; g_E[2] = E_Vdiff; g_E[3] = E_Odiff;
    g_n_layers = (int)E_gen.size();
    if (g_ctx) report(dkmc_set_layer_energies(g_ctx, g_n_layers, g_E[0].data(), g_E[1].data(), g_E[2].data(), g_E[3].data()),
                      "copytoConstMemory");
}

// The reference's site order puts all lattice atoms before all interstitials (reorder_boundary.py:113-124): the rows
// of K are then not spatially sorted.  Give the solver an internal row order — interior rows x-major by 3.6 A grid
// cell, ties in the caller's order — unless the input already has it (dkmc_solver_set_order; every public array keeps
// the caller's order).  Same rule as the Python mirror (devicekmc_b200/structures.py: cell_order).
static void register_solver_order(GPUBuffers &gpubuf, const dkmc_sparsity &sp, int n_contact) {
    const int N = gpubuf.N_, m = sp.m;
    if (m <= 0 || m != N - 2 * n_contact) return;
    std::vector<double> x(N), y(N), z(N);
    if (cudaMemcpy(x.data(), gpubuf.site_x, sizeof(double) * N, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(y.data(), gpubuf.site_y, sizeof(double) * N, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(z.data(), gpubuf.site_z, sizeof(double) * N, cudaMemcpyDeviceToHost) != cudaSuccess) return;
    double x0 = x[0];
    for (int i = 1; i < N; ++i) x0 = x[i] < x0 ? x[i] : x0;
    const double edge = 3.6;
    std::vector<long long> key(m);
    for (int r = 0; r < m; ++r) {
        const int i = n_contact + r;
        const long long cx = (long long)std::floor((x[i] - x0) / edge), cy = (long long)std::floor(y[i] / edge),
                        cz = (long long)std::floor(z[i] / edge);
        key[r] = ((cx + (1ll << 19)) << 42) | ((cy + (1ll << 20)) << 21) | (cz + (1ll << 20));
    }
    std::vector<int> order(m);
    for (int r = 0; r < m; ++r) order[r] = r;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
    bool identity = true;
    for (int r = 0; r < m && identity; ++r) identity = order[r] == r;
    if (identity) return;
    int *d_order = nullptr;
    if (cudaMalloc(&d_order, sizeof(int) * (size_t)m) != cudaSuccess) return;
    cudaMemcpy(d_order, order.data(), sizeof(int) * (size_t)m, cudaMemcpyHostToDevice);
    report(dkmc_solver_set_order(ctx(), &sp, d_order), "initialize_sparsity (solver row order)");
    cudaFree(d_order);
}

// gpu_solvers.h:43 — the CSR index buffers become members of gpubuf, as in the reference
void initialize_sparsity(GPUBuffers &gpubuf, int pbc, const double nn_dist, int num_atoms_contact) {
    (void)pbc; (void)nn_dist;  // the structure follows from gpubuf.neigh_idx (same cutoff, Device.cpp:98-136)
    dkmc_sparsity sp;
    int st = dkmc_initialize_sparsity(ctx(), gpubuf.N_, gpubuf.nn_, gpubuf.neigh_idx, num_atoms_contact,
                                      num_atoms_contact, &sp);
    report(st, "initialize_sparsity");
    if (st != DKMC_OK) return;
    gpubuf.Device_row_ptr_d = sp.d_row_ptr; gpubuf.Device_col_indices_d = sp.d_col; gpubuf.Device_nnz = sp.nnz;
    gpubuf.contact_left_row_ptr = sp.d_left_row_ptr; gpubuf.contact_left_col_indices = sp.d_left_col;
    gpubuf.contact_left_nnz = sp.left_nnz;
    gpubuf.contact_right_row_ptr = sp.d_right_row_ptr; gpubuf.contact_right_col_indices = sp.d_right_col;
    gpubuf.contact_right_nnz = sp.right_nnz;
    register_solver_order(gpubuf, sp, num_atoms_contact);
}

// gpu_solvers.h:127-130
void update_charge_gpu(ELEMENT *gpu_site_element, int *gpu_site_charge, int *gpu_neigh_idx, int N, int nn,
                       const ELEMENT *metals, const int num_metals) {
    report(dkmc_update_charge(ctx(), reinterpret_cast<const int *>(gpu_site_element), gpu_site_charge, gpu_neigh_idx, N,
                              nn, reinterpret_cast<const int *>(metals), num_metals),
           "update_charge_gpu");
}

// gpu_solvers.h:139-141
void background_potential_gpu_sparse(cublasHandle_t, cusolverDnHandle_t, GPUBuffers &gpubuf, const int N,
                                     const int N_left_tot, const int N_right_tot, const double d_Vd, const int pbc,
                                     const double d_high_G, const double d_low_G, const double nn_dist,
                                     const int num_metals, int kmc_step_count) {
    (void)nn_dist; (void)kmc_step_count;
    dkmc_sparsity sp = sparsity_of(gpubuf, N - N_left_tot - N_right_tot);
    dkmc_solve_info info = {};
    // Device::updatePotential (potential_solver.cpp:249-260) calls poisson_gridless_gpu on the same
    // gpubuf right after this function: start that sum now on the side stream so that it overlaps
    // the CG; the poisson_gridless_gpu call below then only joins it.
    report(dkmc_poisson_gridless_begin(ctx(), pbc, gpubuf.N_, gpubuf.lattice, gpubuf.sigma, gpubuf.k, gpubuf.site_x,
                                       gpubuf.site_y, gpubuf.site_z, gpubuf.site_charge, 0, gpubuf.N_,
                                       gpubuf.site_potential_charge),
           "poisson_gridless_gpu (early start)");
    int st = dkmc_background_potential_sparse(ctx(), &sp, N, gpubuf.nn_, gpubuf.neigh_idx, N_left_tot, N_right_tot, d_Vd,
                                              d_high_G, d_low_G, reinterpret_cast<const int *>(gpubuf.site_element),
                                              gpubuf.site_charge, reinterpret_cast<const int *>(gpubuf.metal_types),
                                              num_metals, gpubuf.site_potential_boundary, nullptr, &info);
    report(st, "background_potential_gpu_sparse");
    std::cout << "# CG steps: " << info.iterations << "\n";  // iterative_solvers_gpu.cu:457
}

// gpu_solvers.h:121-123.  Called once per bias point from Device::setLaplacePotential
// (potential_solver.cpp:15) when solve_current = 1.  The reference solves in volts and scales the
// whole vector by eV_to_J = 1.60217663e-19 afterwards (potential_solver_gpu.cu:672); the system is
// linear, so the contacts carry the factor here.
void update_CB_edge_gpu_sparse(cublasHandle_t, cusolverDnHandle_t, GPUBuffers &gpubuf, const int N,
                               const int N_left_tot, const int N_right_tot, const double d_Vd, const int pbc,
                               const double d_high_G, const double d_low_G, const double nn_dist,
                               const int num_metals) {
    (void)pbc; (void)nn_dist;
    dkmc_sparsity sp = sparsity_of(gpubuf, N - N_left_tot - N_right_tot);
    dkmc_solve_info info = {};
    int st = dkmc_update_CB_edge_sparse(ctx(), &sp, N, N_left_tot, N_right_tot, d_Vd, 1.60217663e-19, d_high_G, d_low_G,
                                        reinterpret_cast<const int *>(gpubuf.site_element),
                                        reinterpret_cast<const int *>(gpubuf.metal_types), num_metals,
                                        gpubuf.site_CB_edge, nullptr, &info);
    report(st, "update_CB_edge_gpu_sparse");
    std::cout << "# CG steps: " << info.iterations << "\n";  // iterative_solvers_gpu.cu:457
}

// gpu_solvers.h:144-147
void poisson_gridless_gpu(const int num_atoms_contact, const int pbc, const int N, const double *lattice,
                          const double *sigma, const double *k, const double *posx, const double *posy,
                          const double *posz, const int *site_charge, double *site_potential_charge) {
    (void)num_atoms_contact;
    // joins the sum started in background_potential_gpu_sparse when the arguments match, else computes
    report(dkmc_poisson_gridless(ctx(), pbc, N, lattice, sigma, k, posx, posy, posz, site_charge, site_potential_charge),
           "poisson_gridless_gpu");
}

// gpu_solvers.h:196-201.  The reference draws two numbers per executed event from `rng`
// (kmc_events.cu:221,348).  The device loop consumes pre-drawn numbers of a COPY of the generator
// and the caller's generator is then advanced by exactly the count used.
double execute_kmc_step_gpu(const int N, const int nn, const int *neigh_idx, const int *site_layer,
                            const double *lattice, const int pbc, const double *T_bg, const double *freq,
                            const double *sigma, const double *k, const double *posx, const double *posy,
                            const double *posz, const double *site_potential_boundary,
                            const double *site_potential_charge, const double *site_temperature,
                            ELEMENT *site_element, int *site_charge, RandomNumberGenerator &rng,
                            const int *neigh_idx_host) {
    (void)site_temperature; (void)neigh_idx_host;
    const int batch = 4096;
    std::vector<double> u(batch);
    dkmc_step_info info = {};
    auto draw_ahead = [&]() {
        RandomNumberGenerator ahead = rng;  // copy: peeking must not consume
        for (int i = 0; i < batch; ++i) u[i] = ahead.getRandomNumber();
    };
    draw_ahead();
    int st = dkmc_execute_kmc_step(ctx(), N, nn, neigh_idx, site_layer, lattice, pbc, T_bg, freq, sigma, k, posx, posy,
                                   posz, site_potential_boundary, site_potential_charge,
                                   reinterpret_cast<int *>(site_element), site_charge, u.data(), batch, nullptr, 0, &info);
    for (int i = 0; i < info.n_used; ++i) (void)rng.getRandomNumber();
    while (st == DKMC_ERR_RNG_EXHAUSTED) {
        draw_ahead();
        st = dkmc_kmc_step_continue(ctx(), u.data(), batch, nullptr, 0, &info);
        for (int i = 0; i < info.n_used; ++i) (void)rng.getRandomNumber();
    }
    report(st, "execute_kmc_step_gpu");
    if (st != DKMC_OK) return HUGE_VAL;  // a failed step must not leave the caller's `while (kmc_time < t)` spinning
    std::cout << "Number of KMC steps: " << info.n_events << "\n";  // kmc_events.cu:360
    return info.event_time;
}

}  // extern "C"
