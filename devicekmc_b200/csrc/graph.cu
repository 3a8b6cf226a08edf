// devicekmc-b200 — neighbour graph (cell list), CSR structure of K, and the site-charge
// state machine.  Integer outputs, bit-exact against the reference:
//   a1  Device::constructSiteNeighborList / is_neighbor     Device.cpp:98-136,175-199
//       padded neigh_idx table                               Device.cpp:68-80
//   a2  initialize_sparsity -> Assemble_K_sparsity           iterative_solvers_gpu.cu:96-109,2158-2208
//   a3  Device::updateCharge (CPU branch)                    potential_solver.cpp:172-217
// The reference builds a1/a2 with O(N^2) all-pairs loops; here a uniform cell grid of edge
// >= nn_dist limits every site to its 27 surrounding cells.
#include "common.cuh"
#include "scan.cuh"

namespace dkmc {

// ---------------------------------------------------------------- bounding box
__global__ void __launch_bounds__(1024) bounds_kernel(int N, const double *__restrict__ x,
                                                      const double *__restrict__ y,
                                                      const double *__restrict__ z, double *out) {
    __shared__ double sh[6][32];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double v[3] = {x[i], y[i], z[i]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = fmin(lo[d], v[d]);
            hi[d] = fmax(hi[d], v[d]);
        }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    if (lane == 0)
        for (int d = 0; d < 3; ++d) { sh[d][w] = lo[d]; sh[3 + d][w] = hi[d]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        int d = threadIdx.x;
        double r = sh[d][0];
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) r = d < 3 ? fmin(r, sh[d][k]) : fmax(r, sh[d][k]);
        out[d] = r;
    }
}

struct GridDesc {
    int ncx, ncy, ncz, pbc;
    double minx, miny, minz, wx, wy, wz, ly, lz;
};

__device__ __forceinline__ int cell_of(const GridDesc &g, double x, double y, double z) {
    int cx = (int)floor((x - g.minx) / g.wx);
    int cy, cz;
    if (g.pbc) {
        double fy = y / g.ly; fy -= floor(fy); if (fy >= 1.0) fy = 0.0;
        double fz = z / g.lz; fz -= floor(fz); if (fz >= 1.0) fz = 0.0;
        cy = (int)(fy * g.ncy);
        cz = (int)(fz * g.ncz);
    } else {
        cy = (int)floor((y - g.miny) / g.wy);
        cz = (int)floor((z - g.minz) / g.wz);
    }
    cx = min(max(cx, 0), g.ncx - 1);
    cy = min(max(cy, 0), g.ncy - 1);
    cz = min(max(cz, 0), g.ncz - 1);
    return (cx * g.ncy + cy) * g.ncz + cz;
}

__global__ void cell_count_kernel(int N, GridDesc g, const double *__restrict__ x,
                                  const double *__restrict__ y, const double *__restrict__ z,
                                  int *__restrict__ cell_of_site, int *cell_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = cell_of(g, x[i], y[i], z[i]);
    cell_of_site[i] = c;
    atomicAdd(cell_count + c, 1);
}

__global__ void cell_fill_kernel(int N, const int *__restrict__ cell_of_site,
                                 const int *__restrict__ cell_start, int *cell_fill,
                                 int *__restrict__ cell_sites) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = cell_of_site[i];
    int slot = atomicAdd(cell_fill + c, 1);
    cell_sites[cell_start[c] + slot] = i;  // order inside a cell is irrelevant: rows are sorted later
}

// Visit every j != i in the 27 cells around site i with dist(i,j) < nn_dist.
template <typename F>
__device__ __forceinline__ void for_each_neighbor(int i, const GridDesc &g, double nn_dist,
                                                  const double *__restrict__ x,
                                                  const double *__restrict__ y,
                                                  const double *__restrict__ z,
                                                  const int *__restrict__ cell_of_site,
                                                  const int *__restrict__ cell_start,
                                                  const int *__restrict__ cell_sites, F f) {
    const double xi = x[i], yi = y[i], zi = z[i];
    const int c = cell_of_site[i];
    const int cz = c % g.ncz, cy = (c / g.ncz) % g.ncy, cx = c / (g.ncz * g.ncy);
    int ys[3], zs[3], ny = 0, nz = 0;
    for (int d = -1; d <= 1; ++d) {
        int yy = cy + d, zz = cz + d;
        if (g.pbc) {
            yy = ((yy % g.ncy) + g.ncy) % g.ncy;
            zz = ((zz % g.ncz) + g.ncz) % g.ncz;
        }
        if (yy >= 0 && yy < g.ncy) {
            bool dup = false;
            for (int k = 0; k < ny; ++k) dup |= (ys[k] == yy);
            if (!dup) ys[ny++] = yy;
        }
        if (zz >= 0 && zz < g.ncz) {
            bool dup = false;
            for (int k = 0; k < nz; ++k) dup |= (zs[k] == zz);
            if (!dup) zs[nz++] = zz;
        }
    }
    for (int dx = -1; dx <= 1; ++dx) {
        int xx = cx + dx;
        if (xx < 0 || xx >= g.ncx) continue;
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < nz; ++b) {
                int cc = (xx * g.ncy + ys[a]) * g.ncz + zs[b];
                int s0 = cell_start[cc], s1 = cell_start[cc + 1];
                for (int s = s0; s < s1; ++s) {
                    int j = cell_sites[s];
                    if (j == i) continue;
                    double d = site_dist_exact(xi, yi, zi, x[j], y[j], z[j], g.ly, g.lz, g.pbc);
                    if (d < nn_dist) f(j);
                }
            }
    }
}

__global__ void degree_kernel(int N, GridDesc g, double nn_dist, const double *__restrict__ x,
                              const double *__restrict__ y, const double *__restrict__ z,
                              const int *__restrict__ cell_of_site, const int *__restrict__ cell_start,
                              const int *__restrict__ cell_sites, int *__restrict__ deg, int *max_deg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (i < N) {
        for_each_neighbor(i, g, nn_dist, x, y, z, cell_of_site, cell_start, cell_sites, [&](int) { ++c; });
        deg[i] = c;
    }
    for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_deg, c);
}

__global__ void neighbor_fill_kernel(int N, int nn, GridDesc g, double nn_dist,
                                     const double *__restrict__ x, const double *__restrict__ y,
                                     const double *__restrict__ z, const int *__restrict__ cell_of_site,
                                     const int *__restrict__ cell_start, const int *__restrict__ cell_sites,
                                     int *__restrict__ neigh_idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int *row = neigh_idx + (size_t)i * nn;
    int c = 0;
    for_each_neighbor(i, g, nn_dist, x, y, z, cell_of_site, cell_start, cell_sites, [&](int j) {
        if (c < nn) {
            // insertion keeps the row ascending in j (Device.cpp:105-112 visits j = 0..N-1)
            int p = c;
            while (p > 0 && row[p - 1] > j) { row[p] = row[p - 1]; --p; }
            row[p] = j;
        }
        ++c;
    });
    for (int s = min(c, nn); s < nn; ++s) row[s] = -1;
}

// ---------------------------------------------------------------- CSR structure (a2)
__global__ void sparsity_count_kernel(int N, int nn, int NL, int NR, const int *__restrict__ neigh_idx,
                                      int *__restrict__ cnt_i, int *__restrict__ cnt_l,
                                      int *__restrict__ cnt_r) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int m = N - NL - NR;
    if (r >= m) return;
    const int *row = neigh_idx + (size_t)(r + NL) * nn;
    int ci = 1, cl = 0, cr = 0;  // the diagonal is part of the interior block
    for (int s = 0; s < nn; ++s) {
        int j = row[s];
        if (j < 0) break;
        if (j < NL) ++cl;
        else if (j >= N - NR) ++cr;
        else ++ci;
    }
    cnt_i[r] = ci; cnt_l[r] = cl; cnt_r[r] = cr;
}

__global__ void sparsity_fill_kernel(int N, int nn, int NL, int NR, const int *__restrict__ neigh_idx,
                                     const int *__restrict__ row_ptr, int *__restrict__ col,
                                     const int *__restrict__ lrp, int *__restrict__ lcol,
                                     const int *__restrict__ rrp, int *__restrict__ rcol) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int m = N - NL - NR;
    if (r >= m) return;
    int i = r + NL;
    const int *row = neigh_idx + (size_t)i * nn;
    int pi = row_ptr[r], pl = lrp[r], pr = rrp[r];
    bool diag_done = false;
    for (int s = 0; s < nn; ++s) {
        int j = row[s];
        if (j < 0) break;
        if (j < NL) lcol[pl++] = j;
        else if (j >= N - NR) rcol[pr++] = j - (N - NR);
        else {
            if (!diag_done && j > i) { col[pi++] = r; diag_done = true; }
            col[pi++] = j - NL;
        }
    }
    if (!diag_done) col[pi++] = r;
}

// ---------------------------------------------------------------- charge (a3)
__global__ void update_charge_kernel(int N, int nn, const int *__restrict__ element,
                                     const int *__restrict__ neigh_idx, const int *__restrict__ metals,
                                     int num_metals, int *__restrict__ charge) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int e = element[i];
    if (e != DKMC_VACANCY && e != DKMC_OXYGEN_DEFECT) return;
    const int *row = neigh_idx + (size_t)i * nn;
    int q = (e == DKMC_VACANCY) ? 2 : -2;
    int vnn = 0;
    for (int s = 0; s < nn; ++s) {
        int j = row[s];
        if (j < 0) break;
        int ej = element[j];
        bool metal = false;
        for (int k = 0; k < num_metals; ++k) metal |= (metals[k] == ej);
        if (e == DKMC_VACANCY) {
            if (ej == DKMC_VACANCY) ++vnn;
            if (metal || vnn >= 2) { q = 0; break; }
        } else if (metal) { q = 0; break; }
    }
    charge[i] = q;
}

static int make_grid(dkmc_ctx *ctx, int N, const double *d_x, const double *d_y, const double *d_z,
                     const double *lattice, int pbc, double nn_dist, GridDesc *g) {
    double *d_b = nullptr;
    int rc = ensure<double>(ctx, S_NB_BOUNDS, 6, &d_b);
    if (rc) return rc;
    DKMC_LAUNCH(ctx, bounds_kernel, 1, 1024, 0, N, d_x, d_y, d_z, d_b);
    double b[6];
    DKMC_CUDA(cudaMemcpyAsync(b, d_b, sizeof(b), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    double w = nn_dist * 1.001 + 1e-9;
    g->pbc = pbc ? 1 : 0;
    g->minx = b[0]; g->miny = b[1]; g->minz = b[2];
    g->wx = w;
    g->ncx = (int)floor((b[3] - b[0]) / w) + 1;
    g->ly = lattice[1]; g->lz = lattice[2];
    if (pbc) {
        g->ncy = (int)floor(lattice[1] / w); if (g->ncy < 1) g->ncy = 1;
        g->ncz = (int)floor(lattice[2] / w); if (g->ncz < 1) g->ncz = 1;
        g->wy = lattice[1] / g->ncy; g->wz = lattice[2] / g->ncz;
    } else {
        g->ncy = (int)floor((b[4] - b[1]) / w) + 1;
        g->ncz = (int)floor((b[5] - b[2]) / w) + 1;
        g->wy = w; g->wz = w;
    }
    long long ncell = (long long)g->ncx * g->ncy * g->ncz;
    DKMC_REQUIRE(ncell < (1ll << 30), "cell grid too large (positions span too far for nn_dist)");
    return DKMC_OK;
}

static int build_cells(dkmc_ctx *ctx, int N, const GridDesc &g, const double *d_x, const double *d_y,
                       const double *d_z, int **cell_of_site, int **cell_start, int **cell_sites) {
    int ncell = g.ncx * g.ncy * g.ncz;
    int *d_fill = nullptr, *d_tmp = nullptr;
    int rc;
    if ((rc = ensure<int>(ctx, S_NB_CELL_OF, N, cell_of_site))) return rc;
    if ((rc = ensure<int>(ctx, S_NB_CELL_START, (size_t)ncell + 1, cell_start))) return rc;
    if ((rc = ensure<int>(ctx, S_NB_CELL_FILL, (size_t)ncell + 1, &d_fill))) return rc;
    if ((rc = ensure<int>(ctx, S_NB_CELL_SITES, N, cell_sites))) return rc;
    if ((rc = ensure<int>(ctx, S_SCAN_BLOCK, (size_t)ceil_div(ncell, kScanTile) + 1, &d_tmp))) return rc;
    DKMC_CUDA(cudaMemsetAsync(d_fill, 0, ((size_t)ncell + 1) * sizeof(int), ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(*cell_start, 0, sizeof(int), ctx->stream));
    DKMC_LAUNCH(ctx, cell_count_kernel, ceil_div(N, 256), 256, 0, N, g, d_x, d_y, d_z, *cell_of_site, d_fill);
    // cell_start[c+1] = inclusive scan of counts
    if ((rc = inclusive_scan<int>(ctx, d_fill, ncell, *cell_start + 1, d_tmp))) return rc;
    DKMC_CUDA(cudaMemsetAsync(d_fill, 0, ((size_t)ncell + 1) * sizeof(int), ctx->stream));
    DKMC_LAUNCH(ctx, cell_fill_kernel, ceil_div(N, 256), 256, 0, N, *cell_of_site, *cell_start, d_fill, *cell_sites);
    return DKMC_OK;
}

}  // namespace dkmc

using namespace dkmc;

extern "C" {

int dkmc_neighbor_count(dkmc_ctx *ctx, int N, const double *d_x, const double *d_y, const double *d_z,
                        const double *lattice, int pbc, double nn_dist, int *max_nn) {
    DKMC_REQUIRE(ctx && d_x && d_y && d_z && lattice && max_nn, "null pointer");
    DKMC_REQUIRE(N > 0 && nn_dist > 0, "N and nn_dist must be positive");
    GridDesc g;
    int rc;
    if ((rc = make_grid(ctx, N, d_x, d_y, d_z, lattice, pbc, nn_dist, &g))) return rc;
    int *cell_of_site, *cell_start, *cell_sites, *d_deg;
    if ((rc = build_cells(ctx, N, g, d_x, d_y, d_z, &cell_of_site, &cell_start, &cell_sites))) return rc;
    if ((rc = ensure<int>(ctx, S_NB_DEG, (size_t)N + 1, &d_deg))) return rc;
    int *d_max = d_deg + N;
    DKMC_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream));
    DKMC_LAUNCH(ctx, degree_kernel, ceil_div(N, 128), 128, 0, N, g, nn_dist, d_x, d_y, d_z, cell_of_site,
                cell_start, cell_sites, d_deg, d_max);
    DKMC_CUDA(cudaMemcpyAsync(max_nn, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->grid.valid = true;
    ctx->grid.N = N;
    return DKMC_OK;
}

int dkmc_neighbor_fill(dkmc_ctx *ctx, int N, const double *d_x, const double *d_y, const double *d_z,
                       const double *lattice, int pbc, double nn_dist, int nn, int *d_neigh_idx) {
    DKMC_REQUIRE(ctx && d_x && d_y && d_z && lattice && d_neigh_idx, "null pointer");
    DKMC_REQUIRE(N > 0 && nn > 0, "N and nn must be positive");
    GridDesc g;
    int rc;
    if ((rc = make_grid(ctx, N, d_x, d_y, d_z, lattice, pbc, nn_dist, &g))) return rc;
    int *cell_of_site, *cell_start, *cell_sites;
    if ((rc = build_cells(ctx, N, g, d_x, d_y, d_z, &cell_of_site, &cell_start, &cell_sites))) return rc;
    DKMC_LAUNCH(ctx, neighbor_fill_kernel, ceil_div(N, 128), 128, 0, N, nn, g, nn_dist, d_x, d_y, d_z,
                cell_of_site, cell_start, cell_sites, d_neigh_idx);
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return DKMC_OK;
}

// the same builder for HOST arrays, for a host class (the reference's Device) that keeps
// positions and the neighbour table in host memory.  One upload of the positions, count + fill on
// the device, one download of the table.
int dkmc_neighbor_table_host(dkmc_ctx *ctx, int N, const double *h_x, const double *h_y, const double *h_z,
                             const double *lattice, int pbc, double nn_dist, int *max_nn, int *h_neigh_idx) {
    DKMC_REQUIRE(ctx && h_x && h_y && h_z && lattice && max_nn, "null pointer");
    DKMC_REQUIRE(N > 0 && nn_dist > 0, "N and nn_dist must be positive");
    double *d_pos;
    int rc;
    if ((rc = ensure<double>(ctx, S_NB_HOSTPOS, (size_t)3 * N, &d_pos))) return rc;
    double *d_x = d_pos, *d_y = d_pos + N, *d_z = d_pos + 2 * (size_t)N;
    DKMC_CUDA(cudaMemcpyAsync(d_x, h_x, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    DKMC_CUDA(cudaMemcpyAsync(d_y, h_y, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    DKMC_CUDA(cudaMemcpyAsync(d_z, h_z, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = dkmc_neighbor_count(ctx, N, d_x, d_y, d_z, lattice, pbc, nn_dist, max_nn))) return rc;
    if (!h_neigh_idx || *max_nn == 0) return DKMC_OK;
    int *d_tab;
    const size_t cells = (size_t)N * (size_t)*max_nn;
    if ((rc = ensure<int>(ctx, S_NB_HOSTTAB, cells, &d_tab))) return rc;
    if ((rc = dkmc_neighbor_fill(ctx, N, d_x, d_y, d_z, lattice, pbc, nn_dist, *max_nn, d_tab))) return rc;
    DKMC_CUDA(cudaMemcpyAsync(h_neigh_idx, d_tab, sizeof(int) * cells, cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return DKMC_OK;
}

int dkmc_initialize_sparsity(dkmc_ctx *ctx, int N, int nn, const int *d_neigh_idx, int NL, int NR,
                             dkmc_sparsity *out) {
    DKMC_REQUIRE(ctx && d_neigh_idx && out, "null pointer");
    int m = N - NL - NR;
    DKMC_REQUIRE(N > 0 && nn > 0 && NL >= 0 && NR >= 0 && m > 0, "need N - NL - NR > 0");
    memset(out, 0, sizeof(*out));
    out->m = m;
    int *cnt = nullptr, *tmp = nullptr;
    int rc;
    if ((rc = ensure<int>(ctx, S_SP_CNT, (size_t)3 * m, &cnt))) return rc;
    if ((rc = ensure<int>(ctx, S_SCAN_BLOCK, (size_t)ceil_div(m, kScanTile) + 1, &tmp))) return rc;
    DKMC_CUDA(cudaMalloc(&out->d_row_ptr, ((size_t)m + 1) * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&out->d_left_row_ptr, ((size_t)m + 1) * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&out->d_right_row_ptr, ((size_t)m + 1) * sizeof(int)));
    DKMC_CUDA(cudaMemsetAsync(out->d_row_ptr, 0, sizeof(int), ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(out->d_left_row_ptr, 0, sizeof(int), ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(out->d_right_row_ptr, 0, sizeof(int), ctx->stream));
    DKMC_LAUNCH(ctx, sparsity_count_kernel, ceil_div(m, 128), 128, 0, N, nn, NL, NR, d_neigh_idx, cnt,
                cnt + m, cnt + 2 * (size_t)m);
    if ((rc = inclusive_scan<int>(ctx, cnt, m, out->d_row_ptr + 1, tmp))) return rc;
    if ((rc = inclusive_scan<int>(ctx, cnt + m, m, out->d_left_row_ptr + 1, tmp))) return rc;
    if ((rc = inclusive_scan<int>(ctx, cnt + 2 * (size_t)m, m, out->d_right_row_ptr + 1, tmp))) return rc;
    int tot[3];
    DKMC_CUDA(cudaMemcpyAsync(&tot[0], out->d_row_ptr + m, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaMemcpyAsync(&tot[1], out->d_left_row_ptr + m, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaMemcpyAsync(&tot[2], out->d_right_row_ptr + m, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    out->nnz = tot[0]; out->left_nnz = tot[1]; out->right_nnz = tot[2];
    DKMC_CUDA(cudaMalloc(&out->d_col, ((size_t)tot[0] + 1) * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&out->d_left_col, ((size_t)tot[1] + 1) * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&out->d_right_col, ((size_t)tot[2] + 1) * sizeof(int)));
    DKMC_LAUNCH(ctx, sparsity_fill_kernel, ceil_div(m, 128), 128, 0, N, nn, NL, NR, d_neigh_idx,
                out->d_row_ptr, out->d_col, out->d_left_row_ptr, out->d_left_col, out->d_right_row_ptr,
                out->d_right_col);
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return DKMC_OK;
}

int dkmc_free_sparsity(dkmc_ctx *ctx, dkmc_sparsity *sp) {
    DKMC_REQUIRE(ctx && sp, "null pointer");
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    int **ptrs[6] = {&sp->d_row_ptr, &sp->d_col, &sp->d_left_row_ptr, &sp->d_left_col,
                     &sp->d_right_row_ptr, &sp->d_right_col};
    for (auto p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    if (ctx->tiling.row_ptr && ctx->tiling.d_tile_row) {  // tiling may reference the freed row_ptr
        cudaFree(ctx->tiling.d_tile_row);
        ctx->tiling = SpmvTiling();
    }
    return DKMC_OK;
}

int dkmc_update_charge(dkmc_ctx *ctx, const int *d_site_element, int *d_site_charge,
                       const int *d_neigh_idx, int N, int nn, const int *d_metals, int num_metals) {
    DKMC_REQUIRE(ctx && d_site_element && d_site_charge && d_neigh_idx, "null pointer");
    DKMC_REQUIRE(num_metals == 0 || d_metals != nullptr, "metals");
    DKMC_LAUNCH(ctx, update_charge_kernel, ceil_div(N, 256), 256, 0, N, nn, d_site_element, d_neigh_idx,
                d_metals, num_metals, d_site_charge);
    return DKMC_OK;  // asynchronous, as the reference's launch (potential_solver_gpu.cu:62)
}

}  // extern "C"
