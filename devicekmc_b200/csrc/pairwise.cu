// devicekmc-b200 — pairwise (gridless) Coulomb potential of the charged sites.
//   a6  Device::poisson_gridless + v_solve (CPU semantics)    potential_solver.cpp:412-432, utils.h:102
//       poisson_gridless_gpu / calculate_pairwise_interaction  potential_solver_gpu.cu:908-978
// phi_c[i] = sum_{j != i, q_j != 0} q_j * erfc(r_ij / (sigma*sqrt(2))) * k * q_e / r_ij,  r in metres.
// The reference launches N*ceil(N/512) blocks, one thread per (i,j) pair incl. uncharged j, and
// combines with a shared-memory tree + atomicAdd(double).  Here: the charged sites are
// compacted (ascending j = the CPU summation order), tiles of them are staged in shared memory
// and every thread owns one target site and accumulates in registers — no atomics, FP64-pipe
// bound, O(N * N_charged).
#define DKMC_CARVEOUT_MAXSHARED 1
#include "common.cuh"
#include "scan.cuh"
#include "erfc_coeffs.cuh"

#include <cstdlib>

namespace dkmc {

constexpr int kPwThreads = 128;      // CTA size when the sum has the GPU to itself
constexpr int kPwMaxThreads = 256;   // the CTA size is a launch parameter (targets per tile = 2 x CTA size)
constexpr int kPwTile = 256;       // charged sources per shared-memory tile
constexpr int kCompactBlock = 1024;
constexpr int kPwFullBlocksPerSm = 6;   // ~80 registers x 128 threads: six CTAs fill an SM
constexpr int kPwCellFullBlocksPerSm = 6;   // cell-list kernel: same footprint as the all-pairs kernel

struct __align__(32) ChargedSite { double x, y, z, q; };
constexpr double kErfcZero = 26.45;   // erfc_fast(t) is exactly 0 from here on

__global__ void __launch_bounds__(kCompactBlock) charged_count_kernel(int N, const int *__restrict__ charge,
                                                                    int *__restrict__ block_count) {
    __shared__ int sh[32];
    int i = blockIdx.x * kCompactBlock + threadIdx.x;
    int f = (i < N && charge[i] != 0) ? 1 : 0;
    unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_count[blockIdx.x] = v;
    }
}

// block_incl = inclusive scan of block_count.  Writes the charged sites in ascending site order.
__global__ void __launch_bounds__(kCompactBlock) charged_scatter_kernel(
    int N, int nblocks, const int *__restrict__ charge, const double *__restrict__ x,
    const double *__restrict__ y, const double *__restrict__ z, const int *__restrict__ block_count,
    const int *__restrict__ block_incl, ChargedSite *__restrict__ src, int *__restrict__ src_idx,
    int *__restrict__ total) {
    __shared__ int sh[32];
    int i = blockIdx.x * kCompactBlock + threadIdx.x;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int q = i < N ? charge[i] : 0;
    int f = q != 0 ? 1 : 0;
    unsigned b = __ballot_sync(0xffffffffu, f);
    int in_warp = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) sh[w] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = sh[threadIdx.x];
        int inc = warp_inclusive_scan_int(v, threadIdx.x);
        sh[threadIdx.x] = inc - v;
    }
    __syncthreads();
    int base = block_incl[blockIdx.x] - block_count[blockIdx.x];
    if (f) {
        int pos = base + sh[w] + in_warp;
        ChargedSite s;
        s.x = x[i]; s.y = y[i]; s.z = z[i]; s.q = (double)q;
        src[pos] = s;
        src_idx[pos] = i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *total = block_incl[nblocks - 1];
}

// 8f-2: dq = q - prev, prev = q.  With prev == nullptr-like first use the caller runs a full sum instead.
__global__ void pw_delta_kernel(int N, const int *__restrict__ charge, int *__restrict__ prev, int *__restrict__ dq) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int q = charge[i];
    dq[i] = q - prev[i];
    prev[i] = q;
}

// ---- branch-free FP64 building blocks (MUFU seed + one third-order Newton step)
__device__ __forceinline__ double rsqrt_fast(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a * y, y, 1.0);                 // 1 - a y^2
    double c = fma(0.375, e, 0.5) * e;              // e/2 + 3 e^2/8
    return fma(y, c, y);
}
__device__ __forceinline__ double rcp_fast(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    double c = fma(e, e, e);
    return fma(y, c, y);
}
// exp(-s) for 0 <= s <= 700 (no denormal results in that range)
__device__ __forceinline__ double exp_neg_fast(double s) {
    const double kMagic = 6755399441055744.0;      // 1.5 * 2^52: rounds to nearest integer
    double tmp = fma(-s, 1.4426950408889634, kMagic);
    int n = __double2loint(tmp);
    double nd = tmp - kMagic;
    double r = fma(nd, -6.93147180369123816490e-01, -s);
    r = fma(nd, -1.90821492927058770002e-10, r);
    double p = kExpC[kExpDeg];
#pragma unroll
    for (int i = kExpDeg - 1; i >= 0; --i) p = fma(p, r, kExpC[i]);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
// erfc(t) / 1  for t >= 0, relative error ~3e-15 (tools/fit_erfcx.py): exp(-t^2) * erfcx(t)
__device__ __forceinline__ double erfc_fast(double t) {
    double w = rcp_fast(t + kErfcxK);
    double v = fma(kErfcxB, w, kErfcxA);
    double g = kErfcxC[kErfcxDeg];
#pragma unroll
    for (int i = kErfcxDeg - 1; i >= 0; --i) g = fma(g, v, kErfcxC[i]);
    double s = fmin(t * t, 700.0);
    double e = exp_neg_fast(s);
    return t < kErfcZero ? g * e : 0.0;  // erfc < 1e-305 beyond: contributes nothing at 1e-10
}

// Far field, t = c r >= kErfcFarT0: erfc(t) / r = exp(-t^2) F(1/t^2) / (c sqrt(pi) r^2) — no square root and
// no division beyond 1/r^2 (tools/fit_erfcx.py: F by a degree-10 polynomial, 1e-14 relative).  Returns
// exp(-t^2) F(1/t^2) / r^2; the caller multiplies the sum by 1 / (c sqrt(pi)).  ~40 FP64 instructions per
// pair instead of ~62 for the general formula.  Exactly 0 from t = kErfcZero on, like erfc_fast.
__device__ __forceinline__ double erfc_far_over_r(double r2, double c2, double inv_c2) {
    const double y = rcp_fast(r2);          // 1 / r^2
    const double t2 = r2 * c2;              // t^2
    const double sv = y * inv_c2;           // 1 / t^2
    double f = kErfcFarC[kErfcFarDeg];
#pragma unroll
    for (int i = kErfcFarDeg - 1; i >= 0; --i) f = fma(f, sv, kErfcFarC[i]);
    const double e = exp_neg_fast(fmin(t2, 700.0));
    return t2 < kErfcZero * kErfcZero ? (f * e) * y : 0.0;
}

// phi_c[i] = k q_e sum_j q_j erfc(r_ij / (sigma sqrt 2)) / r_ij.  Two targets per thread, sources
// staged in shared memory, two sources per iteration: four independent FP64 chains per thread keep
// the FP64 pipe busy with few resident warps (the kernel shares the SMs with the CG), no atomics.
// Persistent CTAs: each fetches tiles of 2 x blockDim targets from an atomic counter, so the launch
// can be sized to a chosen residency per SM (the whole SM when the sum runs alone, a share of it
// when it runs beside the CG on the side stream).  Every target's sum runs over the compacted
// sources in ascending site order (even and odd sources in two accumulators), so the result does
// not depend on the tile schedule.
template <bool PBC>
__global__ void __launch_bounds__(kPwMaxThreads) pairwise_kernel(
    int row_begin, int row_end, const double *__restrict__ x, const double *__restrict__ y,
    const double *__restrict__ z, const int *__restrict__ n_src_ptr, const ChargedSite *__restrict__ src,
    const int *__restrict__ src_idx, const double *__restrict__ lattice, const double *__restrict__ sigma_ptr,
    const double *__restrict__ k_ptr, int *tile_counter, int *sm_count, int sm_quota, int accumulate, double *out) {
    __shared__ ChargedSite tile[kPwTile];
    __shared__ int tile_idx[kPwTile];
    __shared__ int s_tile;
    // Even spread when the kernel shares the SMs with the CG: the launch holds more CTAs than wanted,
    // every CTA registers on its SM, and those beyond the SM's quota leave at once (the tiles are
    // fetched dynamically, so nobody's work is lost).  Without this the block scheduler may pack the
    // persistent CTAs onto some of the SMs and leave the others to the CG alone.
    if (sm_count != nullptr) {
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_tile = atomicAdd(sm_count + (smid & 255u), 1);
            atomicAdd(sm_count + 256, 1);   // CTAs that have started (the gate on the main stream waits for all)
        }
        __syncthreads();
        if (s_tile >= sm_quota) return;
    }
    const int nsrc = *n_src_ptr;
    const double sigma = *sigma_ptr, kc = *k_ptr;
    const double ly = lattice[1], lz = lattice[2];
    const double inv_ly = 1.0 / ly, inv_lz = 1.0 / lz;
    const double cscale = 1e-10 / (sigma * sqrt(2.0));  // t = r[Angstrom] * cscale
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const int nthr = (int)blockDim.x, targets = 2 * nthr;
    const int n_tiles = (row_end - row_begin + targets - 1) / targets;

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int tile_id = s_tile;
        if (tile_id >= n_tiles) break;
        const int ia = row_begin + tile_id * targets + threadIdx.x, ib = ia + nthr;
        const bool va = ia < row_end, vb = ib < row_end;
        const double xa = va ? x[ia] : 0.0, ya = va ? y[ia] : 0.0, za = va ? z[ia] : 0.0;
        const double xb = vb ? x[ib] : 0.0, yb = vb ? y[ib] : 0.0, zb = vb ? z[ib] : 0.0;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;

        auto pair_term = [&](double xi, double yi, double zi, int i, const ChargedSite &s, int sidx) -> double {
            double dx = xi - s.x, dy = yi - s.y, dz = zi - s.z;
            if (PBC) {
                dy = fma(-rint(dy * inv_ly), ly, dy);
                dz = fma(-rint(dz * inv_lz), lz, dz);
            }
            double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            double rinv = rsqrt_fast(r2);
            double r = r2 * rinv;
            double term = s.q * erfc_fast(r * cscale) * rinv;
            term = (r2 == 0.0) ? s.q * inf : term;   // coincident sites: the reference divides by zero
            return (sidx == i) ? 0.0 : term;         // i != j (potential_solver.cpp:422)
        };

        for (int t0 = 0; t0 < nsrc; t0 += kPwTile) {
            const int nt = min(kPwTile, nsrc - t0);
            __syncthreads();
            for (int t = threadIdx.x; t < nt; t += nthr) {
                tile[t] = src[t0 + t];
                tile_idx[t] = src_idx[t0 + t];
            }
            __syncthreads();
            int t = 0;
            for (; t + 1 < nt; t += 2) {
                const ChargedSite s0 = tile[t], s1 = tile[t + 1];
                const int j0 = tile_idx[t], j1 = tile_idx[t + 1];
                a0 += pair_term(xa, ya, za, ia, s0, j0);
                b0 += pair_term(xb, yb, zb, ib, s0, j0);
                a1 += pair_term(xa, ya, za, ia, s1, j1);
                b1 += pair_term(xb, yb, zb, ib, s1, j1);
            }
            if (t < nt) {
                const ChargedSite s0 = tile[t];
                const int j0 = tile_idx[t];
                a0 += pair_term(xa, ya, za, ia, s0, j0);
                b0 += pair_term(xb, yb, zb, ib, s0, j0);
            }
        }
        // rinv is in 1/Angstrom: 1e10 converts to 1/m
        // accumulate: the sources are charge DIFFERENCES and out holds the previous potential (8f-2)
        if (va) { const double v = (a0 + a1) * (kc * kElementaryCharge * 1e10); out[ia] = accumulate ? out[ia] + v : v; }
        if (vb) { const double v = (b0 + b1) * (kc * kElementaryCharge * 1e10); out[ib] = accumulate ? out[ib] + v : v; }
    }
}

// ---------------------------------------------------------------- cell list of the charged sites
// erfc_fast returns exactly 0 for t = r / (sigma sqrt 2) >= kErfcZero (erfc < 1e-305 there), so a
// source farther than Rc = kErfcZero * sigma * sqrt 2 (131 A at sigma = 3.5 A) adds exactly nothing:
// skipping it is not an approximation.  The charged sites are binned every step into a uniform grid
// of edge Rc / 8; a warp visits only the cells whose box comes within Rc of the box of its 32
// targets (a warp-uniform decision: no divergence).  At 1 M sites 56 % of the pairs lie beyond Rc.
constexpr int kPwMaxCells = 8192;
constexpr int kPwCellsPerCutoff = 8;   // cell edge = cutoff / 8 (16 A): the visited volume is ~1.26x the cutoff sphere

struct PwGrid {
    double ox, oy, oz, h, inv_h, rc2;   // origin (A), cell edge, 1/edge, cutoff^2 (A^2)
    int ncx, ncy, ncz, reach;           // cells per axis; cells to look at on either side
};

// min / max of the coordinates: per-block partials, then one block
__global__ void __launch_bounds__(256) pw_bbox_partial_kernel(int N, const double *__restrict__ x, const double *__restrict__ y,
                                                              const double *__restrict__ z, double *__restrict__ part) {
    __shared__ double sh[6][8];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const double c[3] = {x[i], y[i], z[i]};
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], c[a]); hi[a] = fmax(hi[a], c[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        for (int a = 0; a < 3; ++a) { sh[a][w] = lo[a]; sh[3 + a][w] = hi[a]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sh[threadIdx.x][0];
        for (int k = 1; k < 8; ++k) v = threadIdx.x < 3 ? fmin(v, sh[threadIdx.x][k]) : fmax(v, sh[threadIdx.x][k]);
        part[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void pw_grid_setup_kernel(int nblocks, const double *__restrict__ part, const double *__restrict__ sigma_ptr,
                                     double cutoff_sigmas, PwGrid *g) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int b = 0; b < nblocks; ++b)
        for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], part[b * 6 + a]); hi[a] = fmax(hi[a], part[b * 6 + 3 + a]); }
    // exact: the distance beyond which erfc_fast is 0.  cutoff_sigmas > 0: the caller's (shorter) cutoff
    double rc = kErfcZero * (*sigma_ptr) * sqrt(2.0) * 1e10 * (1.0 + 1e-12);  // Angstrom, rounded up
    if (cutoff_sigmas > 0.0) rc = fmin(rc, cutoff_sigmas * (*sigma_ptr) * 1e10);
    static_assert(kPwCellsPerCutoff >= 1, "cells per cutoff");
    double h = rc / kPwCellsPerCutoff;
    int n[3];
    while (true) {
        for (int a = 0; a < 3; ++a) n[a] = (int)floor((hi[a] - lo[a]) / h) + 1;
        if ((long long)n[0] * n[1] * n[2] <= kPwMaxCells) break;
        h *= 1.25;
    }
    g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
    g->h = h; g->inv_h = 1.0 / h; g->rc2 = rc * rc;
    g->ncx = n[0]; g->ncy = n[1]; g->ncz = n[2];
    g->reach = (int)ceil(rc / h);
}

__device__ __forceinline__ int pw_cell_coord(double c, double o, double inv_h, int n) {
    int k = (int)floor((c - o) * inv_h);
    return k < 0 ? 0 : (k >= n ? n - 1 : k);
}

// cell of every charged site and the population of the cells
__global__ void pw_cell_count_kernel(const int *__restrict__ n_src_ptr, const ChargedSite *__restrict__ src,
                                     const PwGrid *__restrict__ gp, int *__restrict__ cell_of, int *__restrict__ cell_count) {
    const int n = *n_src_ptr;
    const PwGrid g = *gp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const ChargedSite s = src[e];
        const int c = (pw_cell_coord(s.x, g.ox, g.inv_h, g.ncx) * g.ncy + pw_cell_coord(s.y, g.oy, g.inv_h, g.ncy)) * g.ncz +
                      pw_cell_coord(s.z, g.oz, g.inv_h, g.ncz);
        cell_of[e] = c;
        atomicAdd(cell_count + c, 1);
    }
}

// exclusive scan of the (few thousand) cell populations by one block
__global__ void __launch_bounds__(1024) pw_cell_scan_kernel(const PwGrid *__restrict__ gp, const int *__restrict__ cell_count,
                                                            int *__restrict__ cell_start) {
    __shared__ int sh[32];
    __shared__ int carry_s;
    const int ncells = gp->ncx * gp->ncy * gp->ncz;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < ncells; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < ncells ? cell_count[i] : 0;
        int inc = warp_inclusive_scan_int(v, lane);
        if (lane == 31) sh[w] = inc;
        __syncthreads();
        if (w == 0) sh[lane] = warp_inclusive_scan_int(sh[lane], lane);
        __syncthreads();
        const int excl = carry_s + (w > 0 ? sh[w - 1] : 0) + inc - v;
        if (i < ncells) cell_start[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cell_start[ncells] = carry_s;
}

// charged sites grouped by cell, ascending site index inside a cell (deterministic): every entry takes the next
// free slot of its cell (atomic cursor, arbitrary order), then one thread per cell sorts the cell's few entries
// by site index.  (Round 1 ranked every entry against all earlier ones: O(n_charged^2).)
__global__ void __launch_bounds__(256) pw_cell_fill_kernel(const int *__restrict__ n_src_ptr, const ChargedSite *__restrict__ src,
                                                           const int *__restrict__ src_idx, const int *__restrict__ cell_of,
                                                           const int *__restrict__ cell_start, int *__restrict__ cursor,
                                                           ChargedSite *__restrict__ out, int *__restrict__ out_idx) {
    const int n = *n_src_ptr;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int c = cell_of[e];
        const int pos = cell_start[c] + atomicAdd(cursor + c, 1);
        out[pos] = src[e];
        out_idx[pos] = src_idx[e];
    }
}
__global__ void __launch_bounds__(128) pw_cell_sort_kernel(const PwGrid *__restrict__ gp, const int *__restrict__ cell_start,
                                                           ChargedSite *__restrict__ out, int *__restrict__ out_idx) {
    const int ncells = gp->ncx * gp->ncy * gp->ncz;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int a = cell_start[c], b = cell_start[c + 1];
    for (int k = a + 1; k < b; ++k) {   // insertion sort by site index
        const int key = out_idx[k];
        const ChargedSite v = out[k];
        int j = k;
        while (j > a && out_idx[j - 1] > key) { out_idx[j] = out_idx[j - 1]; out[j] = out[j - 1]; --j; }
        out_idx[j] = key; out[j] = v;
    }
}

// Two targets per lane (a warp owns 64 consecutive targets); the warp walks the cells near the box
// of its targets in (x, y, z) cell order and, inside a cell, the sources in ascending site index, two
// per iteration into two accumulators per target (four independent FP64 chains per lane, every source
// loaded once for two targets).  Same persistent tile scheduling and per-SM quota as pairwise_kernel.
template <int MINB>
__global__ void __launch_bounds__(kPwMaxThreads, MINB) pairwise_cells_kernel(
    int row_begin, int row_end, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
    const PwGrid *__restrict__ gp, const int *__restrict__ cell_start, const ChargedSite *__restrict__ src,
    const int *__restrict__ src_idx, const double *__restrict__ sigma_ptr, const double *__restrict__ k_ptr,
    int *tile_counter, int *sm_count, int sm_quota, unsigned sm_quota_linger_ns, unsigned long long *pair_counter,
    int accumulate, int far_on, double *out) {
    __shared__ int s_tile;
    if (sm_count != nullptr) {
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_tile = atomicAdd(sm_count + (smid & 255u), 1);
            atomicAdd(sm_count + 256, 1);   // CTAs that have started (the gate on the main stream waits for all)
        }
        __syncthreads();
        if (s_tile >= sm_quota) {
            // do not leave at once: an SM that frees a slot instantly would swallow the whole surplus
            // of the launch while busy SMs never receive their share
            if (sm_quota_linger_ns > 0) __nanosleep(sm_quota_linger_ns);
            return;
        }
    }
    const PwGrid g = *gp;
    const double sigma = *sigma_ptr, kc = *k_ptr;
    const double cscale = 1e-10 / (sigma * sqrt(2.0));
    const double c2 = cscale * cscale, inv_c2 = 1.0 / c2;
    const double far_scale = 1.0 / (cscale * 1.7724538509055160273);   // 1 / (c sqrt(pi))
    // far-field distance: t >= kErfcFarT0 with a margin for the rounding of the box arithmetic
    const double r_far = far_on ? kErfcFarT0 / cscale * (1.0 + 1e-9) : 1e300, r_far2 = far_on ? r_far * r_far : 1e300;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const int lane = threadIdx.x & 31;
    const int n_tiles = (row_end - row_begin + 63) / 64;   // a tile = the 64 targets of one warp
    unsigned long long my_pairs = 0, my_far = 0;

    while (true) {
        // every warp fetches its own tiles: no CTA-wide barrier, a warp with a short cell walk does
        // not wait for its neighbours
        int tile_id = 0;
        if (lane == 0) tile_id = atomicAdd(tile_counter, 1);
        tile_id = __shfl_sync(0xffffffffu, tile_id, 0);
        if (tile_id >= n_tiles) break;
        const int ia = row_begin + tile_id * 64 + lane, ib = ia + 32;
        const bool va = ia < row_end, vb = ib < row_end;
        const int ca = va ? ia : row_end - 1, cb = vb ? ib : row_end - 1;   // idle lanes shadow the last target
        const double xa = x[ca], ya = y[ca], za = z[ca], xb = x[cb], yb = y[cb], zb = z[cb];
        // box of the warp's 64 targets -> warp-uniform cell range
        double lo[3] = {fmin(xa, xb), fmin(ya, yb), fmin(za, zb)}, hi[3] = {fmax(xa, xb), fmax(ya, yb), fmax(za, zb)};
#pragma unroll
        for (int a = 0; a < 3; ++a)
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            }
        const int cx0 = max(0, pw_cell_coord(lo[0], g.ox, g.inv_h, g.ncx) - g.reach);
        const int cx1 = min(g.ncx - 1, pw_cell_coord(hi[0], g.ox, g.inv_h, g.ncx) + g.reach);
        const int cy0 = max(0, pw_cell_coord(lo[1], g.oy, g.inv_h, g.ncy) - g.reach);
        const int cy1 = min(g.ncy - 1, pw_cell_coord(hi[1], g.oy, g.inv_h, g.ncy) + g.reach);
        const int cz0 = max(0, pw_cell_coord(lo[2], g.oz, g.inv_h, g.ncz) - g.reach);
        const int cz1 = min(g.ncz - 1, pw_cell_coord(hi[2], g.oz, g.inv_h, g.ncz) + g.reach);
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;      // near field: general formula
        double fa0 = 0.0, fa1 = 0.0, fb0 = 0.0, fb1 = 0.0;  // far field: sums of q exp(-t^2) F / r^2

        auto pair_term = [&](double xi, double yi, double zi, int i, const ChargedSite &s, int sidx) -> double {
            double dx = xi - s.x, dy = yi - s.y, dz = zi - s.z;
            double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            double rinv = rsqrt_fast(r2);
            double r = r2 * rinv;
            double term = s.q * erfc_fast(r * cscale) * rinv;
            term = (r2 == 0.0) ? s.q * inf : term;   // coincident sites: the reference divides by zero
            return (sidx == i) ? 0.0 : term;         // i != j (potential_solver.cpp:422)
        };
        // a far run lies at least r_far from every target of the warp: t >= kErfcFarT0, r > 0, j != i
        auto far_term = [&](double xi, double yi, double zi, const ChargedSite &s) -> double {
            double dx = xi - s.x, dy = yi - s.y, dz = zi - s.z;
            double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            return s.q * erfc_far_over_r(r2, c2, inv_c2);
        };
        // the sources [s0, s1) of the sorted list, two per iteration, the next two requested ahead
        auto near_run = [&](int s0, int s1) {
            int e = s0;
            if (e + 1 < s1) {
                ChargedSite p0 = src[e], p1 = src[e + 1];
                int j0 = __ldg(src_idx + e), j1 = __ldg(src_idx + e + 1);
                for (; e + 3 < s1; e += 2) {
                    const ChargedSite n0 = src[e + 2], n1 = src[e + 3];
                    const int k0 = __ldg(src_idx + e + 2), k1 = __ldg(src_idx + e + 3);
                    a0 += pair_term(xa, ya, za, ia, p0, j0);
                    b0 += pair_term(xb, yb, zb, ib, p0, j0);
                    a1 += pair_term(xa, ya, za, ia, p1, j1);
                    b1 += pair_term(xb, yb, zb, ib, p1, j1);
                    p0 = n0; p1 = n1; j0 = k0; j1 = k1;
                }
                a0 += pair_term(xa, ya, za, ia, p0, j0);
                b0 += pair_term(xb, yb, zb, ib, p0, j0);
                a1 += pair_term(xa, ya, za, ia, p1, j1);
                b1 += pair_term(xb, yb, zb, ib, p1, j1);
                e += 2;
            }
            if (e < s1) {
                const ChargedSite p0 = src[e];
                const int j0 = __ldg(src_idx + e);
                a0 += pair_term(xa, ya, za, ia, p0, j0);
                b0 += pair_term(xb, yb, zb, ib, p0, j0);
            }
        };
        auto far_run = [&](int s0, int s1) {
            int e = s0;
            if (e + 1 < s1) {
                ChargedSite p0 = src[e], p1 = src[e + 1];
                for (; e + 3 < s1; e += 2) {
                    const ChargedSite n0 = src[e + 2], n1 = src[e + 3];
                    fa0 += far_term(xa, ya, za, p0);
                    fb0 += far_term(xb, yb, zb, p0);
                    fa1 += far_term(xa, ya, za, p1);
                    fb1 += far_term(xb, yb, zb, p1);
                    p0 = n0; p1 = n1;
                }
                fa0 += far_term(xa, ya, za, p0);
                fb0 += far_term(xb, yb, zb, p0);
                fa1 += far_term(xa, ya, za, p1);
                fb1 += far_term(xb, yb, zb, p1);
                e += 2;
            }
            if (e < s1) {
                const ChargedSite p0 = src[e];
                fa0 += far_term(xa, ya, za, p0);
                fb0 += far_term(xb, yb, zb, p0);
            }
        };

        for (int cx = cx0; cx <= cx1; ++cx) {
            const double bx0 = g.ox + cx * g.h, bx1 = bx0 + g.h;
            const double gx = fmax(0.0, fmax(bx0 - hi[0], lo[0] - bx1));
            for (int cy = cy0; cy <= cy1; ++cy) {
                const double by0 = g.oy + cy * g.h, by1 = by0 + g.h;
                const double gy = fmax(0.0, fmax(by0 - hi[1], lo[1] - by1));
                const double gxy = gx * gx + gy * gy;
                if (gxy > g.rc2) continue;
                // the cells of a z-column are contiguous in the sorted source list: find the z-range that
                // comes within the cutoff and walk its sources as ONE run
                const double dzmax = sqrt(g.rc2 - gxy);
                int za0 = pw_cell_coord(lo[2] - dzmax, g.oz, g.inv_h, g.ncz), za1 = pw_cell_coord(hi[2] + dzmax, g.oz, g.inv_h, g.ncz);
                za0 = max(za0, cz0); za1 = min(za1, cz1);
                if (za0 > za1) continue;
                const int c0 = (cx * g.ncy + cy) * g.ncz;
                const int s0 = __ldg(cell_start + c0 + za0), s1 = __ldg(cell_start + c0 + za1 + 1);
                const unsigned long long np = (unsigned long long)(s1 - s0) * ((va ? 1 : 0) + (vb ? 1 : 0));
                my_pairs += np;
                // the cells of the column that may hold a source closer than r_far to one of the targets take the
                // general formula; the run below and the run above them the far-field one (warp-uniform)
                int zn0 = za1 + 1, zn1 = za1;    // empty near range: the whole column is far
                if (gxy < r_far2) {
                    const double dzn = sqrt(r_far2 - gxy);
                    zn0 = max(za0, pw_cell_coord(lo[2] - dzn, g.oz, g.inv_h, g.ncz));
                    zn1 = min(za1, pw_cell_coord(hi[2] + dzn, g.oz, g.inv_h, g.ncz));
                    if (zn0 > zn1) { zn0 = za1 + 1; zn1 = za1; }
                }
                if (zn0 > za1) {
                    far_run(s0, s1);
                    my_far += np;
                } else {
                    const int n0 = __ldg(cell_start + c0 + zn0), n1 = __ldg(cell_start + c0 + zn1 + 1);
                    far_run(s0, n0);
                    near_run(n0, n1);
                    far_run(n1, s1);
                    my_far += (unsigned long long)((n0 - s0) + (s1 - n1)) * ((va ? 1 : 0) + (vb ? 1 : 0));
                }
            }
        }
        // rinv is in 1/Angstrom: 1e10 converts to 1/m; the far-field sums carry the factor 1 / (c sqrt(pi))
        // accumulate: the sources are charge DIFFERENCES and out holds the previous potential (8f-2)
        if (va) { const double v = ((a0 + a1) + (fa0 + fa1) * far_scale) * (kc * kElementaryCharge * 1e10); out[ia] = accumulate ? out[ia] + v : v; }
        if (vb) { const double v = ((b0 + b1) + (fb0 + fb1) * far_scale) * (kc * kElementaryCharge * 1e10); out[ib] = accumulate ? out[ib] + v : v; }
    }
    if (pair_counter != nullptr) {   // pairs evaluated (all | by the far-field formula), for the roofline
        for (int o = 16; o > 0; o >>= 1) {
            my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
            my_far += __shfl_xor_sync(0xffffffffu, my_far, o);
        }
        if (lane == 0) { atomicAdd(pair_counter, my_pairs); atomicAdd(pair_counter + 1, my_far); }
    }
}

// Main-stream gate after the side-stream launch: returns once every CTA of the pairwise launch has
// started (or after ~200 us), so that the persistent CTAs are spread over an idle GPU before the CG's
// kernels arrive — otherwise SMs that happen to be busy never receive their share.
__global__ void pw_gate_kernel(const int *started, int expected) {
    const long long t0 = clock64();
    while (*(volatile const int *)started < expected && clock64() - t0 < 400000) { }
}

// compaction of the charged sites (on the context's main stream) and the tile counter
static unsigned pw_linger_ns() {
    static const unsigned v = [] { const char *e = getenv("DKMC_PW_LINGER_NS"); return e ? (unsigned)atoi(e) : 0u; }();
    return v;
}

// Beside the CG the cell-list kernel runs in its 64-register build (launch bound 4 CTAs of 256 threads; ~400 bytes
// of spills): at a share of 3 x 128 threads per SM it then holds 24.5 k instead of 30.7 k registers, which leaves the
// persistent PCG 4 CTAs per SM instead of 3.  Measured at 1 M sites: the sum 26.5 -> 27.8 ms, the CG beside it
// 35.3 -> 31.5 ms, the step 39.0 -> 35.0 ms (25.6 -> 28.6 KMC steps/s).  Alone the 80-register build is used.
// DKMC_PW_LEAN=0: the 80-register build everywhere.
static bool pw_lean() {
    static const bool on = [] { const char *e = getenv("DKMC_PW_LEAN"); return e ? atoi(e) != 0 : true; }();
    return on;
}

static bool pw_gate_enabled() {
    static const bool on = [] { const char *e = getenv("DKMC_PW_GATE"); return e ? atoi(e) != 0 : false; }();
    return on;
}

struct PwCells {
    PwGrid *grid = nullptr;
    int *cell_start = nullptr;
    ChargedSite *src = nullptr;
    int *src_idx = nullptr;
    unsigned long long *pair_counter = nullptr;
};

// bins the compacted charged sites into the cell grid (all on the context's main stream)
static int pairwise_bin_cells(dkmc_ctx *ctx, int N, const double *d_x, const double *d_y, const double *d_z,
                              const double *d_sigma, const ChargedSite *src, const int *src_idx, const int *total,
                              PwCells *out) {
    constexpr int kBoxBlocks = 256;
    double *box;          // [6 * kBoxBlocks] partial min/max | PwGrid
    int *cells;           // cell_count[kPwMaxCells + 1] | cell_start[kPwMaxCells + 1] | cell_of[N]
    int rc;
    const size_t box_doubles = 6 * kBoxBlocks + (sizeof(PwGrid) + 7) / 8 + 4;
    if ((rc = ensure<double>(ctx, S_PW_BOX, box_doubles, &box))) return rc;
    if ((rc = ensure<int>(ctx, S_PW_CELLS, (size_t)2 * (kPwMaxCells + 1) + (size_t)N, &cells))) return rc;
    if ((rc = ensure<ChargedSite>(ctx, S_PW_SRC2, (size_t)N, &out->src))) return rc;
    if ((rc = ensure<int>(ctx, S_PW_IDX2, (size_t)N, &out->src_idx))) return rc;
    PwGrid *grid = reinterpret_cast<PwGrid *>(box + 6 * kBoxBlocks);
    unsigned long long *pair_counter = reinterpret_cast<unsigned long long *>(box + 6 * kBoxBlocks + (sizeof(PwGrid) + 7) / 8);
    int *cell_count = cells, *cell_start = cells + (kPwMaxCells + 1), *cell_of = cells + 2 * (kPwMaxCells + 1);
    auto &gc = ctx->pw_grid;
    if (gc.d_x != d_x || gc.d_sigma != d_sigma || gc.N != N || gc.box != box || gc.cutoff_sigmas != ctx->pw_cutoff_sigmas) {
        // positions are static: the box of the sites and the grid are computed once
        DKMC_LAUNCH(ctx, pw_bbox_partial_kernel, kBoxBlocks, 256, 0, N, d_x, d_y, d_z, box);
        DKMC_LAUNCH(ctx, pw_grid_setup_kernel, 1, 32, 0, kBoxBlocks, box, d_sigma, ctx->pw_cutoff_sigmas, grid);
        gc.d_x = d_x; gc.d_sigma = d_sigma; gc.N = N; gc.box = box; gc.cutoff_sigmas = ctx->pw_cutoff_sigmas;
    }
    DKMC_CUDA(cudaMemsetAsync(cell_count, 0, (kPwMaxCells + 1) * sizeof(int), ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(pair_counter, 0, 2 * sizeof(unsigned long long), ctx->stream));
    int grid_n = ceil_div(N, 256);
    if (grid_n > 1024) grid_n = 1024;
    DKMC_LAUNCH(ctx, pw_cell_count_kernel, grid_n, 256, 0, total, src, grid, cell_of, cell_count);
    DKMC_LAUNCH(ctx, pw_cell_scan_kernel, 1, 1024, 0, grid, cell_count, cell_start);
    DKMC_CUDA(cudaMemsetAsync(cell_count, 0, (kPwMaxCells + 1) * sizeof(int), ctx->stream));   // now the cells' fill cursors
    DKMC_LAUNCH(ctx, pw_cell_fill_kernel, grid_n, 256, 0, total, src, src_idx, cell_of, cell_start, cell_count, out->src, out->src_idx);
    DKMC_LAUNCH(ctx, pw_cell_sort_kernel, ceil_div(kPwMaxCells, 128), 128, 0, grid, cell_start, out->src, out->src_idx);
    out->grid = grid; out->cell_start = cell_start; out->pair_counter = pair_counter;
    return DKMC_OK;
}

static int pairwise_prepare(dkmc_ctx *ctx, int N, const double *d_x, const double *d_y, const double *d_z,
                            const int *d_site_charge, ChargedSite **src_out, int **src_idx_out, int **total_out,
                            int **tile_counter_out) {
    const int nb = ceil_div(N, kCompactBlock);
    int *counts, *tmp, *src_idx, *tile_counter;
    ChargedSite *src;
    int rc;
    // counts | inclusive | total
    if ((rc = ensure<int>(ctx, S_PW_COUNT, (size_t)2 * nb + 4, &counts))) return rc;
    if ((rc = ensure<int>(ctx, S_SCAN_BLOCK, (size_t)ceil_div(nb, kScanTile) + 1, &tmp))) return rc;
    if ((rc = ensure<ChargedSite>(ctx, S_PW_SRC, (size_t)N, &src))) return rc;
    if ((rc = ensure<int>(ctx, S_PW_FLAGS, (size_t)N, &src_idx))) return rc;
    if ((rc = ensure<int>(ctx, S_PW_TILECTR, 8 + 256, &tile_counter))) return rc;
    int *incl = counts + nb, *total = counts + 2 * nb;
    DKMC_CUDA(cudaMemsetAsync(tile_counter, 0, (2 + 256) * sizeof(int), ctx->stream));  // tile counter | CTAs per SM | CTAs started
    DKMC_LAUNCH(ctx, charged_count_kernel, nb, kCompactBlock, 0, N, d_site_charge, counts);
    if ((rc = inclusive_scan<int>(ctx, counts, nb, incl, tmp))) return rc;
    DKMC_LAUNCH(ctx, charged_scatter_kernel, nb, kCompactBlock, 0, N, nb, d_site_charge, d_x, d_y, d_z, counts, incl,
                src, src_idx, total);
    *src_out = src; *src_idx_out = src_idx; *total_out = total; *tile_counter_out = tile_counter;
    return DKMC_OK;
}

static int pairwise_launch(dkmc_ctx *ctx, cudaStream_t stream, int blocks_per_sm, int threads, bool shared_sms, int pbc, int row_begin, int row_end,
                           const double *d_lattice, const double *d_sigma, const double *d_k, const double *d_x,
                           const double *d_y, const double *d_z, const ChargedSite *src, const int *src_idx,
                           const int *total, int *tile_counter, double *d_out, const PwCells *cells, int accumulate = 0) {
    if (shared_sms) {
        // what the persistent PCG sizes its grid by: the registers this kernel holds per thread
        static int regs[3] = {0, 0, 0};
        const int which = cells ? 0 : (pbc ? 1 : 2);
        if (!regs[which]) {
            cudaFuncAttributes fa;
            cudaError_t e = cells ? (pw_lean() ? cudaFuncGetAttributes(&fa, pairwise_cells_kernel<4>) : cudaFuncGetAttributes(&fa, pairwise_cells_kernel<3>))
                                  : (pbc ? cudaFuncGetAttributes(&fa, pairwise_kernel<true>) : cudaFuncGetAttributes(&fa, pairwise_kernel<false>));
            regs[which] = e == cudaSuccess ? fa.numRegs : 96;
        }
        ctx->pw_regs_per_thread = regs[which];
    }
    if (cells) {  // one target per thread
        int grid = ctx->num_sms * blocks_per_sm;
        int *sm_count = nullptr;
        const int tiles = ceil_div(ceil_div(row_end - row_begin, 64), threads / 32);
        if (shared_sms) {
            const int fit = (kPwCellFullBlocksPerSm * kPwThreads) / threads;
            grid = ctx->num_sms * (fit > blocks_per_sm ? fit : blocks_per_sm);
            sm_count = tile_counter + 1;
        } else if (grid > tiles) {
            grid = tiles;
        }
        // The preferred carve-out is only a hint: launched with (almost) no shared memory of its own, this
        // kernel was given another L1/shared split than the CG's kernels in about one launch out of three,
        // and the two streams then ran one after the other (67 ms instead of 45 ms).  Asking for 16 KB of
        // (unused) dynamic shared memory, like the all-pairs kernel's staging tile, makes the split stick.
        static const int pad_smem = [] { const char *e = getenv("DKMC_PW_PAD_SMEM"); return e ? atoi(e) : 16384; }();
        if (pw_lean() && shared_sms) {
            DKMC_LAUNCH_ON(ctx, stream, pairwise_cells_kernel<4>, grid, threads, pad_smem, row_begin, row_end, d_x, d_y, d_z, cells->grid,
                           cells->cell_start, cells->src, cells->src_idx, d_sigma, d_k, tile_counter, sm_count, blocks_per_sm,
                           pw_linger_ns(), cells->pair_counter, accumulate, ctx->pw_far_field, d_out);
        } else {
            DKMC_LAUNCH_ON(ctx, stream, pairwise_cells_kernel<3>, grid, threads, pad_smem, row_begin, row_end, d_x, d_y, d_z, cells->grid,
                           cells->cell_start, cells->src, cells->src_idx, d_sigma, d_k, tile_counter, sm_count, blocks_per_sm,
                           pw_linger_ns(), cells->pair_counter, accumulate, ctx->pw_far_field, d_out);
        }
        if (shared_sms && pw_gate_enabled()) DKMC_LAUNCH(ctx, pw_gate_kernel, 1, 1, 0, sm_count + 256, grid);
        return DKMC_OK;
    }
    const int tiles = ceil_div(row_end - row_begin, 2 * threads);
    int grid = ctx->num_sms * blocks_per_sm;
    int *sm_count = nullptr;
    if (shared_sms) {  // launch what fits on an idle GPU; each SM keeps `blocks_per_sm` of them
        const int fit = (kPwFullBlocksPerSm * kPwThreads) / threads;
        grid = ctx->num_sms * (fit > blocks_per_sm ? fit : blocks_per_sm);
        sm_count = tile_counter + 1;
    } else if (grid > tiles) {
        grid = tiles;
    }
    if (pbc) {
        DKMC_LAUNCH_ON(ctx, stream, pairwise_kernel<true>, grid, threads, 0, row_begin, row_end, d_x, d_y, d_z, total,
                       src, src_idx, d_lattice, d_sigma, d_k, tile_counter, sm_count, blocks_per_sm, accumulate, d_out);
    } else {
        DKMC_LAUNCH_ON(ctx, stream, pairwise_kernel<false>, grid, threads, 0, row_begin, row_end, d_x, d_y, d_z, total,
                       src, src_idx, d_lattice, d_sigma, d_k, tile_counter, sm_count, blocks_per_sm, accumulate, d_out);
    }
    if (shared_sms && pw_gate_enabled()) DKMC_LAUNCH(ctx, pw_gate_kernel, 1, 1, 0, sm_count + 256, grid);
    return DKMC_OK;
}

// 8f-2 (opt-in).  Decides between a full sum and an update by the charge differences since the
// previous call on the same arrays.  Returns the charge array the sum has to run over (the charges
// themselves, or the differences) and whether the kernel accumulates into d_out.
static int pairwise_incremental(dkmc_ctx *ctx, int pbc, int N, const int *d_site_charge, int row_begin, int row_end,
                                const double *d_out, const int **charge_for_sum, int *accumulate) {
    *charge_for_sum = d_site_charge;
    *accumulate = 0;
    auto &inc = ctx->pw_inc;
    if (inc.refresh_every <= 0) { inc.valid = false; return DKMC_OK; }
    int *prev, *dq;
    int rc;
    if ((rc = ensure<int>(ctx, S_PW_PREVQ, (size_t)N, &prev))) return rc;
    if ((rc = ensure<int>(ctx, S_PW_DQ, (size_t)N, &dq))) return rc;
    const bool same = inc.valid && inc.d_charge == d_site_charge && inc.d_out == d_out && inc.N == N &&
                      inc.row_begin == row_begin && inc.row_end == row_end && inc.pbc == pbc;
    if (same && inc.since_full + 1 < inc.refresh_every) {
        DKMC_LAUNCH(ctx, pw_delta_kernel, ceil_div(N, 256), 256, 0, N, d_site_charge, prev, dq);
        *charge_for_sum = dq;
        *accumulate = 1;
        ++inc.since_full;
        ++inc.delta_sums;
        return DKMC_OK;
    }
    DKMC_CUDA(cudaMemcpyAsync(prev, d_site_charge, sizeof(int) * (size_t)N, cudaMemcpyDeviceToDevice, ctx->stream));
    inc.valid = true; inc.d_charge = d_site_charge; inc.d_out = d_out; inc.N = N;
    inc.row_begin = row_begin; inc.row_end = row_end; inc.pbc = pbc;
    inc.since_full = 0;
    ++inc.full_sums;
    return DKMC_OK;
}

}  // namespace dkmc

using namespace dkmc;

extern "C" {

int dkmc_poisson_gridless_begin(dkmc_ctx *ctx, int pbc, int N, const double *d_lattice, const double *d_sigma,
                                const double *d_k, const double *d_x, const double *d_y, const double *d_z,
                                const int *d_site_charge, int row_begin, int row_end,
                                double *d_site_potential_charge) {
    DKMC_REQUIRE(ctx && d_lattice && d_sigma && d_k && d_x && d_y && d_z && d_site_charge && d_site_potential_charge,
                 "null pointer");
    DKMC_REQUIRE(N > 0 && row_begin >= 0 && row_end <= N && row_begin <= row_end, "row range");
    DKMC_REQUIRE(!ctx->pw_pending.active, "a pairwise sum is already in flight: call dkmc_poisson_gridless_join");
    ChargedSite *src;
    int *src_idx, *total, *tile_counter;
    int rc, accumulate;
    const int *q_sum;
    if ((rc = pairwise_incremental(ctx, pbc, N, d_site_charge, row_begin, row_end, d_site_potential_charge, &q_sum, &accumulate))) return rc;
    if ((rc = pairwise_prepare(ctx, N, d_x, d_y, d_z, q_sum, &src, &src_idx, &total, &tile_counter))) return rc;
    PwCells cells;
    const bool use_cells = !pbc && ctx->pw_use_cells;
    if (use_cells && (rc = pairwise_bin_cells(ctx, N, d_x, d_y, d_z, d_sigma, src, src_idx, total, &cells))) return rc;
    // fork: the side stream starts after the compaction and everything issued before it
    DKMC_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    DKMC_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    DKMC_CUDA(cudaEventRecord(ctx->ev_pw0, ctx->side_stream));
    if (row_end > row_begin)
        if ((rc = pairwise_launch(ctx, ctx->side_stream, ctx->pw_side_blocks_per_sm, ctx->pw_side_threads, true, pbc, row_begin, row_end, d_lattice,
                                  d_sigma, d_k, d_x, d_y, d_z, src, src_idx, total, tile_counter,
                                  d_site_potential_charge, use_cells ? &cells : nullptr, accumulate))) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_pw1, ctx->side_stream));
    auto &pp = ctx->pw_pending;
    pp.active = true; pp.pbc = pbc; pp.N = N; pp.row_begin = row_begin; pp.row_end = row_end;
    pp.d_charge = d_site_charge; pp.d_out = d_site_potential_charge;
    return DKMC_OK;
}

int dkmc_poisson_gridless_join(dkmc_ctx *ctx, double *pairwise_ms) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    if (pairwise_ms) *pairwise_ms = 0.0;
    if (!ctx->pw_pending.active) return DKMC_OK;
    ctx->pw_pending.active = false;
    DKMC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_pw1, 0));  // later work on the main stream sees the result
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_pw1));
    if (pairwise_ms) {
        float ms = 0.f;
        DKMC_CUDA(cudaEventElapsedTime(&ms, ctx->ev_pw0, ctx->ev_pw1));
        *pairwise_ms = ms;
    }
    return DKMC_OK;
}

int dkmc_poisson_gridless_rows(dkmc_ctx *ctx, int pbc, int N, const double *d_lattice, const double *d_sigma,
                               const double *d_k, const double *d_x, const double *d_y, const double *d_z,
                               const int *d_site_charge, int row_begin, int row_end,
                               double *d_site_potential_charge) {
    DKMC_REQUIRE(ctx && d_lattice && d_sigma && d_k && d_x && d_y && d_z && d_site_charge && d_site_potential_charge,
                 "null pointer");
    DKMC_REQUIRE(N > 0 && row_begin >= 0 && row_end <= N && row_begin <= row_end, "row range");
    auto &pp = ctx->pw_pending;
    if (pp.active) {
        // the same sum was started ahead of time (dkmc_poisson_gridless_begin): just join it
        const bool same = pp.pbc == pbc && pp.N == N && pp.row_begin == row_begin && pp.row_end == row_end &&
                          pp.d_charge == d_site_charge && pp.d_out == d_site_potential_charge;
        int rc = dkmc_poisson_gridless_join(ctx, nullptr);
        if (rc || same) return rc;
    }
    if (row_begin == row_end) return DKMC_OK;
    ChargedSite *src;
    int *src_idx, *total, *tile_counter;
    int rc, accumulate;
    const int *q_sum;
    if ((rc = pairwise_incremental(ctx, pbc, N, d_site_charge, row_begin, row_end, d_site_potential_charge, &q_sum, &accumulate))) return rc;
    if ((rc = pairwise_prepare(ctx, N, d_x, d_y, d_z, q_sum, &src, &src_idx, &total, &tile_counter))) return rc;
    PwCells cells;
    const bool use_cells = !pbc && ctx->pw_use_cells;
    if (use_cells && (rc = pairwise_bin_cells(ctx, N, d_x, d_y, d_z, d_sigma, src, src_idx, total, &cells))) return rc;
    if ((rc = pairwise_launch(ctx, ctx->stream, use_cells ? kPwCellFullBlocksPerSm : kPwFullBlocksPerSm, kPwThreads, false, pbc,
                              row_begin, row_end, d_lattice, d_sigma, d_k, d_x, d_y, d_z, src, src_idx, total, tile_counter,
                              d_site_potential_charge, use_cells ? &cells : nullptr, accumulate))) return rc;
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return DKMC_OK;
}

int dkmc_poisson_gridless(dkmc_ctx *ctx, int pbc, int N, const double *d_lattice, const double *d_sigma,
                          const double *d_k, const double *d_x, const double *d_y, const double *d_z,
                          const int *d_site_charge, double *d_site_potential_charge) {
    return dkmc_poisson_gridless_rows(ctx, pbc, N, d_lattice, d_sigma, d_k, d_x, d_y, d_z, d_site_charge, 0, N,
                                      d_site_potential_charge);
}

int dkmc_ctx_set_pairwise_cells(dkmc_ctx *ctx, int on) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    ctx->pw_use_cells = on ? 1 : 0;
    return DKMC_OK;
}

int dkmc_ctx_set_pairwise_cutoff(dkmc_ctx *ctx, double cutoff_sigmas) {
    DKMC_REQUIRE(ctx != nullptr && cutoff_sigmas >= 0.0, "ctx / cutoff_sigmas >= 0");
    ctx->pw_cutoff_sigmas = cutoff_sigmas;
    return DKMC_OK;
}

int dkmc_ctx_set_pairwise_incremental(dkmc_ctx *ctx, int refresh_every) {
    DKMC_REQUIRE(ctx != nullptr && refresh_every >= 0, "ctx / refresh_every >= 0");
    ctx->pw_inc.refresh_every = refresh_every;
    ctx->pw_inc.valid = false;
    ctx->pw_inc.since_full = 0;
    return DKMC_OK;
}

int dkmc_pairwise_incremental_counts(dkmc_ctx *ctx, long long *full_sums, long long *delta_sums) {
    DKMC_REQUIRE(ctx != nullptr && full_sums != nullptr && delta_sums != nullptr, "null pointer");
    *full_sums = ctx->pw_inc.full_sums;
    *delta_sums = ctx->pw_inc.delta_sums;
    return DKMC_OK;
}

int dkmc_pairwise_pairs_evaluated(dkmc_ctx *ctx, long long *pairs) {
    DKMC_REQUIRE(ctx != nullptr && pairs != nullptr, "ctx/pairs");
    *pairs = -1;
    if (!ctx->slot_ptr[S_PW_BOX]) return DKMC_OK;
    const double *box = static_cast<const double *>(ctx->slot_ptr[S_PW_BOX]);
    const unsigned long long *pc = reinterpret_cast<const unsigned long long *>(box + 6 * 256 + (sizeof(PwGrid) + 7) / 8);
    unsigned long long h = 0;
    DKMC_CUDA(cudaMemcpyAsync(&h, pc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    *pairs = (long long)h;
    return DKMC_OK;
}

int dkmc_pairwise_pairs_far(dkmc_ctx *ctx, long long *pairs) {
    DKMC_REQUIRE(ctx != nullptr && pairs != nullptr, "ctx/pairs");
    *pairs = -1;
    if (!ctx->slot_ptr[S_PW_BOX]) return DKMC_OK;
    const double *box = static_cast<const double *>(ctx->slot_ptr[S_PW_BOX]);
    const unsigned long long *pc = reinterpret_cast<const unsigned long long *>(box + 6 * 256 + (sizeof(PwGrid) + 7) / 8);
    unsigned long long h = 0;
    DKMC_CUDA(cudaMemcpyAsync(&h, pc + 1, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    *pairs = (long long)h;
    return DKMC_OK;
}

int dkmc_ctx_set_pairwise_far_field(dkmc_ctx *ctx, int on) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    ctx->pw_far_field = on ? 1 : 0;
    return DKMC_OK;
}

int dkmc_ctx_set_pairwise_share(dkmc_ctx *ctx, int blocks_per_sm, int threads_per_block) {
    DKMC_REQUIRE(ctx != nullptr && blocks_per_sm >= 1 && blocks_per_sm <= 16, "blocks_per_sm in 1..16");
    DKMC_REQUIRE(threads_per_block >= 32 && threads_per_block <= kPwMaxThreads && threads_per_block % 32 == 0,
                 "threads_per_block: a multiple of 32 up to 256");
    ctx->pw_side_blocks_per_sm = blocks_per_sm;
    ctx->pw_side_threads = threads_per_block;
    return DKMC_OK;
}

}  // extern "C"
