// devicekmc-b200 — host-side set-up replacement.
//
// Device::constructSiteNeighborList (Device.cpp:98-136) tests all N^2 site pairs on the host:
// seconds at 10^4 sites, hours at 10^6.  This file defines the SAME member function on top of the
// library's cell-list builder (dkmc_neighbor_table_host, include/dkmc.h): identical result — the
// same `dist < nn_dist && i != j` predicate evaluated with the same individually rounded FP64
// operations, rows ascending in j — in O(N).
//
// A maintainer either replaces the body in Device.cpp by the one below, or keeps Device.cpp
// untouched and lets this definition win at link time:
//     objcopy --weaken-symbol=_ZN6Device25constructSiteNeighborListEv Device.o
//     g++ -I<DeviceKMC>/src -I<devicekmc-b200>/include -c device_setup_shim.cpp
// Everything after the call in Device::Device (the zero-neighbour check, the padded neigh_idx
// table, Device.cpp:60-80) runs unchanged on the lists filled here.
#include "Device.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dkmc.h"

void Device::constructSiteNeighborList() {
    dkmc_ctx *ctx = nullptr;
    if (dkmc_ctx_create(&ctx) != DKMC_OK) {
        fprintf(stderr, "devicekmc-b200: constructSiteNeighborList: %s\n", dkmc_last_error());
        abort();  // no CPU fallback: the reference loop is not kept as a second path
    }
    const double lat[3] = {lattice[0], lattice[1], lattice[2]};
    int nn = 0;
    int st = dkmc_neighbor_table_host(ctx, N, site_x.data(), site_y.data(), site_z.data(), lat, pbc ? 1 : 0, nn_dist,
                                      &nn, nullptr);
    std::vector<int> table;
    if (st == DKMC_OK && nn > 0) {
        table.resize((size_t)N * nn);
        st = dkmc_neighbor_table_host(ctx, N, site_x.data(), site_y.data(), site_z.data(), lat, pbc ? 1 : 0, nn_dist, &nn,
                                      table.data());
    }
    if (st != DKMC_OK) {
        fprintf(stderr, "devicekmc-b200: constructSiteNeighborList: status %d: %s\n", st, dkmc_last_error());
        abort();
    }
    dkmc_ctx_destroy(ctx);

    #pragma omp parallel for
    for (int i = 0; i < N; i++) {
        const int *row = table.data() + (size_t)i * nn;
        int deg = 0;
        while (deg < nn && row[deg] >= 0) ++deg;
        site_neighbors.l[i].assign(row, row + deg);
    }
    if (nn > this->max_num_neighbors) this->max_num_neighbors = nn;
    site_neighbors.is_constructed = 1;
    std::cout << "Maximum number of neighbors in device is: " << this->max_num_neighbors << "\n";  // Device.cpp:135
}
