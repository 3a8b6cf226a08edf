// devicekmc-b200 — a C++ host over the C-ABI, without any of the reference's headers.
//
// The loop of kmc_main.cpp:175-279 (solve_potential = perturb_structure = 1) for a host program that is
// NOT DeviceKMC: read an xyz file, keep the site arrays on the device, and per KMC step call
//     charge -> [pairwise sum on the side stream || K assembly + PCG] -> rate table + event loop.
// The random stream is the reference's (random_num.h:4-23): std::mt19937 +
// std::uniform_real_distribution<double>(0, 1), two numbers per executed event; the device loop is handed
// numbers drawn ahead from a COPY of the generator and the generator is advanced by the count it used.
//
//   g++ -std=c++17 -O2 -I include -I /usr/local/cuda/include examples/kmc_loop.cpp -o kmc_loop
//       -L devicekmc_b200/lib -ldkmc_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/devicekmc_b200/lib
//   ./a.out device.xyz 108.97557 25.575 25.575 144 6.0 5      (xyz, lattice, contact sites, Vd, steps)
//
// DeviceKMC itself does not need this file: its own host links the reference-named shim instead
// (devicekmc_b200/shim/).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "dkmc.h"

#define CK(call)                                                                                     \
    do {                                                                                             \
        int st_ = (call);                                                                            \
        if (st_ != DKMC_OK) {                                                                        \
            fprintf(stderr, "%s: status %d: %s\n", #call, st_, dkmc_last_error());                   \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                              \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

template <typename T>
static T *to_device(const std::vector<T> &h) {
    T *d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

int main(int argc, char **argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: %s device.xyz Lx Ly Lz n_contact Vd steps\n", argv[0]);
        return 2;
    }
    const double lattice[3] = {atof(argv[2]), atof(argv[3]), atof(argv[4])};
    const int n_contact = atoi(argv[5]);
    const double Vd = atof(argv[6]);
    const int steps = atoi(argv[7]);
    // parameters of the shipped test device (structures/single_devices/test_2.5nm/parameters.txt)
    const int pbc = 0;
    const double nn_dist = 3.5, sigma = 3.5e-10, k = 8.987552e9 / 23.0, T_bg = 300.0, freq = 1e14;
    const double high_G = 1.0, low_G = 1e-8;                                   // input_parser.cpp:392-393
    const std::vector<int> metals = {DKMC_Ti_EL, DKMC_N_EL};
    // structure_input.h:12-50: E_gen, E_rec, E_Vdiff, E_Odiff and the x-range of the five layers
    const double E_gen[5] = {0.0, 3.93, 3.93, 1.66, 1.73}, E_rec[5] = {0, 0, 0, 0, 0};
    const double E_Vdiff[5] = {0.0, 1.09, 1.09, 1.09, 0.0}, E_Odiff[5] = {0.76, 0.76, 0.76, 0.76, 2.8};
    const double layer_x0[5] = {-22.0, 0.0, 3.0, 48.1431, 52.6431}, layer_x1[5] = {0.0, 3.0, 48.1431, 52.6431, 90.0};

    // ---- xyz (utils.cpp:72-98)
    const std::map<std::string, int> element_of = {{"d", DKMC_DEFECT}, {"Od", DKMC_OXYGEN_DEFECT}, {"V", DKMC_VACANCY},
                                                   {"O", DKMC_O_EL}, {"Hf", DKMC_Hf_EL}, {"Ni", DKMC_Ni_EL},
                                                   {"Ti", DKMC_Ti_EL}, {"Pt", DKMC_Pt_EL}, {"N", DKMC_N_EL}};
    std::ifstream in(argv[1]);
    int N = 0;
    std::string line, name;
    in >> N;
    std::getline(in, line);
    std::getline(in, line);
    if (!in || N <= 2 * n_contact) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    std::vector<int> element(N), layer(N), charge(N, 0);
    std::vector<double> x(N), y(N), z(N), zeros(N, 0.0);
    for (int i = 0; i < N; ++i) {
        in >> name >> x[i] >> y[i] >> z[i];
        std::getline(in, line);
        auto it = element_of.find(name);
        if (!in || it == element_of.end()) { fprintf(stderr, "bad xyz row %d\n", i); return 2; }
        element[i] = it->second;
        layer[i] = -1;
        for (int l = 0; l < 5; ++l)                                            // KMCProcess.cpp:34-50: the last match wins
            if (layer_x0[l] <= x[i] && x[i] <= layer_x1[l]) layer[i] = l;
        if (layer[i] < 0) { fprintf(stderr, "Site #%d is not inside the device!\n", i); return 2; }
    }

    // ---- set-up: neighbour table (cell list on the GPU), device mirror, CSR structure of K
    dkmc_ctx *ctx = nullptr;
    CK(dkmc_ctx_create(&ctx));
    int nn = 0;
    CK(dkmc_neighbor_table_host(ctx, N, x.data(), y.data(), z.data(), lattice, pbc, nn_dist, &nn, nullptr));
    std::vector<int> neigh((size_t)N * nn);
    CK(dkmc_neighbor_table_host(ctx, N, x.data(), y.data(), z.data(), lattice, pbc, nn_dist, &nn, neigh.data()));
    printf("N = %d, max neighbours = %d\n", N, nn);
    int *d_element = to_device(element), *d_charge = to_device(charge), *d_layer = to_device(layer);
    int *d_neigh = to_device(neigh), *d_metals = to_device(metals);
    double *d_x = to_device(x), *d_y = to_device(y), *d_z = to_device(z);
    double *d_pb = to_device(zeros), *d_pc = to_device(zeros);
    double *d_lattice = to_device(std::vector<double>(lattice, lattice + 3));
    double *d_sigma = to_device(std::vector<double>{sigma}), *d_k = to_device(std::vector<double>{k});
    double *d_T = to_device(std::vector<double>{T_bg}), *d_freq = to_device(std::vector<double>{freq});
    if (!d_element || !d_charge || !d_layer || !d_neigh || !d_metals || !d_x || !d_y || !d_z || !d_pb || !d_pc ||
        !d_lattice || !d_sigma || !d_k || !d_T || !d_freq) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    CK(dkmc_set_layer_energies(ctx, 5, E_gen, E_rec, E_Vdiff, E_Odiff));
    dkmc_sparsity sp;
    CK(dkmc_initialize_sparsity(ctx, N, nn, d_neigh, n_contact, n_contact, &sp));
    {
        // the file is in the reference's site order (lattice atoms, then interstitials): give the solver an internal
        // row order — interior rows x-major by 3.6 A grid cell — unless the rows already have it; every array of this
        // program keeps the file's order (dkmc_solver_set_order)
        const double x0 = *std::min_element(x.begin(), x.end()), edge = 3.6;
        std::vector<long long> key(sp.m);
        std::vector<int> order(sp.m);
        for (int r = 0; r < sp.m; ++r) {
            const int i = n_contact + r;
            const long long cx = (long long)std::floor((x[i] - x0) / edge), cy = (long long)std::floor(y[i] / edge),
                            cz = (long long)std::floor(z[i] / edge);
            key[r] = ((cx + (1ll << 19)) << 42) | ((cy + (1ll << 20)) << 21) | (cz + (1ll << 20));
            order[r] = r;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
        bool sorted_already = true;
        for (int r = 0; r < sp.m && sorted_already; ++r) sorted_already = order[r] == r;
        if (!sorted_already) {
            int *d_order = to_device(order);
            if (!d_order) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
            CK(dkmc_solver_set_order(ctx, &sp, d_order));
            cudaFree(d_order);
            printf("solver row order: x-major grid cells (input not spatially sorted)\n");
        }
    }

    // ---- the KMC loop
    std::mt19937 rng(1);                                                       // rnd_seed_kmc, structure_input.h:8
    const int batch = 4096;
    std::vector<double> u(batch);
    auto draw_ahead = [&]() {
        std::mt19937 ahead = rng;                                              // peeking must not consume
        for (int i = 0; i < batch; ++i) u[i] = std::uniform_real_distribution<double>(0.0, 1.0)(ahead);
    };
    auto advance = [&](int n) {
        for (int i = 0; i < n; ++i) (void)std::uniform_real_distribution<double>(0.0, 1.0)(rng);
    };
    double kmc_time = 0.0;
    for (int step = 0; step < steps; ++step) {
        CK(dkmc_update_charge(ctx, d_element, d_charge, d_neigh, N, nn, d_metals, (int)metals.size()));
        CK(dkmc_poisson_gridless_begin(ctx, pbc, N, d_lattice, d_sigma, d_k, d_x, d_y, d_z, d_charge, 0, N, d_pc));
        dkmc_solve_info solve = {};
        int st = dkmc_background_potential_sparse(ctx, &sp, N, nn, d_neigh, n_contact, n_contact, Vd, high_G, low_G,
                                                  d_element, d_charge, d_metals, (int)metals.size(), d_pb, nullptr, &solve);
        double pairwise_ms = 0.0;
        CK(dkmc_poisson_gridless_join(ctx, &pairwise_ms));
        if (st != DKMC_OK && st != DKMC_ERR_NOT_CONVERGED) CK(st);
        dkmc_step_info info = {};
        draw_ahead();
        st = dkmc_execute_kmc_step(ctx, N, nn, d_neigh, d_layer, d_lattice, pbc, d_T, d_freq, d_sigma, d_k, d_x, d_y, d_z,
                                   d_pb, d_pc, d_element, d_charge, u.data(), batch, nullptr, 0, &info);
        advance(info.n_used);
        while (st == DKMC_ERR_RNG_EXHAUSTED) {
            draw_ahead();
            st = dkmc_kmc_step_continue(ctx, u.data(), batch, nullptr, 0, &info);
            advance(info.n_used);
        }
        CK(st);
        kmc_time += info.event_time;
        printf("step %d: %d CG iterations (%.2f ms), pairwise %.2f ms, %d events, KMC time is: %g\n", step,
               solve.iterations, solve.solve_ms, pairwise_ms, info.n_events, kmc_time);
    }
    CU(cudaMemcpy(element.data(), d_element, N * sizeof(int), cudaMemcpyDeviceToHost));
    CK(dkmc_free_sparsity(ctx, &sp));
    CK(dkmc_ctx_destroy(ctx));
    return 0;
}
