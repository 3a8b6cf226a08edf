// devicekmc-b200 — background (boundary) potential: K assembly + Jacobi-preconditioned CG.
//   a4  Device::background_potential (CPU semantics)          potential_solver.cpp:289-410
//       background_potential_gpu_sparse / Assemble_A          potential_solver_gpu.cu:696-781,397-493
//   a5  solve_sparse_CG_Jacobi                                iterative_solvers_gpu.cu:309-480
// Design (B200-first, not the reference's launch sequence):
//   * one fused assembly kernel writes off-diagonals, the diagonal (sequential ascending-j sum,
//     the reference's rounding), 1/diag and the rhs — replaces 8 launches + 6 cudaMallocs;
//   * SpMV streams val/col of a 2048-nnz tile fully coalesced into shared memory, then reduces
//     each row in CSR order; algorithmic traffic 12*nnz + 20*m bytes, HBM-bound;
//   * CG scalars (rz, pAp, alpha, beta, convergence flag) never leave the device: reductions
//     finish in the last-arriving block in a fixed order (bitwise reproducible), kernels turn
//     into no-ops once converged, the host polls one flag every `check_every` iterations;
//   * iterative refinement with a double-double residual recovers the digits that cond(K)~1e7
//     (uncharged-vacancy clusters coupled by high_G inside a low_G oxide) takes from plain CG.
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#define DKMC_CARVEOUT_MAXSHARED 1
#include "common.cuh"
#include "scan.cuh"

namespace dkmc {

constexpr int kSpmvThreads = 256;
constexpr int kSpmvCap = 2048;              // shared products per block = 256 threads x 8
constexpr int kSpmvTile = kSpmvCap - 64;    // nnz per tile by row start: with rows <= 63 long a
                                            // tile (+1 alignment slot) always fits one pass
constexpr int kVecThreads = 256;
constexpr int kMaxPartials = 1 << 16;

// experiment switches (env DKMC_FLAGS): bit0 matrix loads cached (not streaming), bit1 streaming
// vector traffic in the CG update, bit2 L2 persistence window on the matrix values
static int g_flags = [] { const char *e = getenv("DKMC_FLAGS"); return e ? atoi(e) : 0; }();

struct CgScalars {
    double rz, rz_new, pAp, alpha, beta, bb, stop, resnorm2, bnorm2;
    int done, iters, max_iter, pad;     // pad != 0: a cross-CTA / cross-GPU wait timed out (error)
    unsigned int cnt_a, cnt_b, cnt_c, cnt_d;
    unsigned long long rseq_end, hseq_end;   // persistent PCG: sequence numbers reached (reductions, halo exchanges)
};

// ---------------------------------------------------------------- site class + assembly
__global__ void site_class_kernel(int N, const int *__restrict__ element, const int *__restrict__ charge,
                                  const int *__restrict__ metals, int num_metals,
                                  unsigned char *__restrict__ cls) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int e = element[i];
    bool metal = false;
    for (int k = 0; k < num_metals; ++k) metal |= (metals[k] == e);
    cls[i] = metal ? 1 : ((e == DKMC_VACANCY && charge != nullptr && charge[i] == 0) ? 2 : 0);
}

// conductance rules.  0, background potential (potential_solver.cpp:325-346): high_G iff (metal & metal)
// or (uncharged V & uncharged V).  1, CB-edge Laplace solve (potential_solver.cpp:58-70): high_G iff
// either site is a metal.
__device__ __forceinline__ bool is_high(unsigned char ci, unsigned char cj, int rule) {
    return rule == 0 ? (ci != 0 && ci == cj) : (ci == 1 || cj == 1);
}
__device__ __forceinline__ double conductance(unsigned char ci, unsigned char cj, double high_G, double low_G, int rule = 0) {
    return is_high(ci, cj, rule) ? high_G : low_G;
}

__global__ void __launch_bounds__(128) assemble_kernel(
    int m, int N, int NL, int NR, double VL, double VR, int rule, double high_G, double low_G,
    const unsigned char *__restrict__ cls, const int *__restrict__ row_ptr, const int *__restrict__ col,
    const int *__restrict__ lrp, const int *__restrict__ lcol, const int *__restrict__ rrp,
    const int *__restrict__ rcol, double *__restrict__ val, double *__restrict__ rhs,
    double *__restrict__ dinv) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const unsigned char ci = cls[r + NL];
    double diag = 0.0, ksub = 0.0;
    // ascending j: left contact, interior, right contact (potential_solver.cpp:350-372)
    for (int p = lrp[r]; p < lrp[r + 1]; ++p) {
        double G = conductance(ci, cls[lcol[p]], high_G, low_G, rule);
        diag = __dadd_rn(diag, G);
        ksub = __dadd_rn(ksub, __dmul_rn(-G, VL));
    }
    int dpos = -1;
    for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
        int c = col[p];
        if (c == r) { dpos = p; continue; }
        const unsigned char cj = cls[c + NL];
        double G = conductance(ci, cj, high_G, low_G, rule);
        val[p] = -G;
        diag = __dadd_rn(diag, G);
    }
    for (int p = rrp[r]; p < rrp[r + 1]; ++p) {
        double G = conductance(ci, cls[rcol[p] + (N - NR)], high_G, low_G, rule);
        diag = __dadd_rn(diag, G);
        ksub = __dadd_rn(ksub, __dmul_rn(-G, VR));
    }
    if (dpos >= 0) val[dpos] = diag;
    rhs[r] = -ksub;  // D*phi = -Ksub (potential_solver.cpp:379,396)
    if (dinv) dinv[r] = 1.0 / diag;
}

// ---------------------------------------------------------------- SpMV tiling
__device__ __forceinline__ int first_row_at(const int *__restrict__ row_ptr, int m, long long target) {
    int lo = 0, hi = m;  // first r in [0,m) with row_ptr[r] >= target
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row_ptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// tile t = rows whose first non-zero index lies in [t*T, (t+1)*T); info = (r0, r1, k0, k1)
__global__ void tile_rows_kernel(int m, int num_tiles, const int *__restrict__ row_ptr, int4 *__restrict__ tile_info) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    int r0 = first_row_at(row_ptr, m, (long long)t * kSpmvTile);
    int r1 = (t + 1 == num_tiles) ? m : first_row_at(row_ptr, m, (long long)(t + 1) * kSpmvTile);
    tile_info[t] = make_int4(r0, r1, row_ptr[r0], row_ptr[r1]);
}

// y = A x over one nnz tile per block.  MODE 0: y only.  MODE 1: also dot(w, y) -> *dot_out.
// MODE 2: y = b - A x (residual) and rr = sum (y^2 * dinv) -> *dot_out.
// Phase 1 streams the tile's val/col with all 8 elements of a thread in flight at once (one DRAM round
// trip per tile), gathers x through the read-only path and parks the products in shared memory;
// phase 2 adds each row's products in CSR order.
template <int MODE>
__global__ void __launch_bounds__(kSpmvThreads, MODE == 0 ? 8 : 6) spmv_tile_kernel(
    int num_tiles, int nnz, const int *__restrict__ row_ptr, const int *__restrict__ col, const double *__restrict__ val,
    const double *__restrict__ x, double *__restrict__ y, const int4 *__restrict__ tile_info,
    const double *__restrict__ w, const double *__restrict__ dinv, double *partials, unsigned int *counter,
    double *dot_out, const int *done_flag, int flags) {
    __shared__ __align__(16) double prod[kSpmvCap];
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    double local = 0.0;
    // grid-stride over the tiles: at most 6 CTAs per SM worth of partial sums for the fused dot
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int4 ti = tile_info[t];
        const int r0 = ti.x, r1 = ti.y;
        if (r0 < r1) {
            const int k0 = ti.z, k1 = ti.w;
            const int ka = k0 & ~1;
            // row bounds of this thread's first row, requested before the big loads
            int my_r = r0 + threadIdx.x, ra = 0, rb = 0;
            if (my_r < r1) { ra = row_ptr[my_r]; rb = row_ptr[my_r + 1]; }
            if (k1 - ka <= kSpmvCap) {
                // two batches of four 8-byte/4-byte loads per thread (32 registers -> 8 blocks per SM)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double v[4];
                    int c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        int k = k0 + (h * 4 + u) * kSpmvThreads + threadIdx.x;
                        bool ok = k < k1;
                        if (flags & 1) {  // keep the matrix in L2 across CG iterations
                            v[u] = ok ? __ldg(val + k) : 0.0;
                            c[u] = ok ? __ldg(col + k) : 0;
                        } else {
                            v[u] = ok ? __ldcs(val + k) : 0.0;
                            c[u] = ok ? __ldcs(col + k) : 0;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        int k = k0 + (h * 4 + u) * kSpmvThreads + threadIdx.x;
                        if (k < k1) prod[k - ka] = v[u] * __ldg(x + c[u]);
                    }
                }
                __syncthreads();
                // phase 2: each row adds its products in CSR order
                for (int r = my_r; r < r1; r += kSpmvThreads) {
                    if (r != my_r) { ra = row_ptr[r]; rb = row_ptr[r + 1]; }
                    double s = 0.0;
#pragma unroll 4
                    for (int k = ra - ka; k < rb - ka; ++k) s += prod[k];
                    if (MODE == 2) { s = w[r] - s; local += s * s * dinv[r]; }
                    y[r] = s;
                    if (MODE == 1) local += w[r] * s;
                }
            } else {  // rows too long for the staging buffer: direct path
                for (int r = my_r; r < r1; r += kSpmvThreads) {
                    double s = 0.0;
                    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) s += val[k] * __ldg(x + col[k]);
                    if (MODE == 2) { s = w[r] - s; local += s * s * dinv[r]; }
                    y[r] = s;
                    if (MODE == 1) local += w[r] * s;
                }
            }
        }
        __syncthreads();  // prod is reused by the next tile
    }
    if (MODE != 0) {
        double tot = block_sum(local, red);
        grid_sum_finish(tot, partials, counter, dot_out, red);
    }
}

// res = b - A x with double-double accumulation (TwoProd via FMA, TwoSum), rounded to double.
// Also accumulates ||res||^2_{D^-1} and ||b||^2_{D^-1}.
__global__ void __launch_bounds__(kSpmvThreads) residual_dd_kernel(
    int m, const int *__restrict__ row_ptr, const int *__restrict__ col, const double *__restrict__ val,
    const double *__restrict__ x, const double *__restrict__ b, const double *__restrict__ dinv,
    double *__restrict__ res, const int4 *__restrict__ tile_info, double *partials, unsigned int *counter,
    double *res_out) {
    __shared__ double hi[kSpmvCap];
    __shared__ double lo[kSpmvCap];
    __shared__ double red[32];
    const int4 ti = tile_info[blockIdx.x];
    const int r0 = ti.x, r1 = ti.y;
    double local = 0.0;
    if (r0 < r1) {
        const int k0 = ti.z, k1 = ti.w;
        const int cnt = k1 - k0;
        const bool staged = cnt <= kSpmvCap;
        if (staged) {
            for (int k = threadIdx.x; k < cnt; k += kSpmvThreads) {
                double a = val[k0 + k], xv = __ldg(x + col[k0 + k]);
                double p = a * xv;
                hi[k] = p;
                lo[k] = fma(a, xv, -p);
            }
        }
        __syncthreads();
        for (int r = r0 + threadIdx.x; r < r1; r += kSpmvThreads) {
            double sh = b[r], sl = 0.0;
            int a = row_ptr[r], e = row_ptr[r + 1];
            for (int k = a; k < e; ++k) {
                double ph, pl;
                if (staged) { ph = hi[k - k0]; pl = lo[k - k0]; }
                else { double av = val[k], xv = __ldg(x + col[k]); ph = av * xv; pl = fma(av, xv, -ph); }
                // TwoSum(sh, -ph)
                double t = sh - ph;
                double bb = t - sh;
                double err = (sh - (t - bb)) + (-ph - bb);
                sh = t;
                sl += err - pl;
            }
            double rr = sh + sl;
            res[r] = rr;
            local += rr * rr * dinv[r];
        }
    }
    double tot = block_sum(local, red);
    grid_sum_finish(tot, partials, counter, res_out, red);
}

// ---------------------------------------------------------------- cluster (coarse) preconditioner
// Uncharged vacancies that neighbour each other are tied by high_G = 1 inside an oxide whose
// other conductances are low_G = 1e-8: every such cluster adds an eigenvalue ~1e-7 to the
// Jacobi-scaled matrix, invisible to the D^-1 residual norm and costing CG thousands of
// iterations.  M^-1 = D^-1 + W E^-1 W^T adds one coarse unknown per cluster (W = cluster
// indicator vectors, E = W^T A W = conductance leaving the cluster) and removes them.
struct Precond {
    const double *dinv;
    const int *pos;        // row -> position in the sorted member list, -1 if not clustered
    const int *seg_start;  // per sorted position: first position of its cluster
    const int *seg_len;    // per sorted position: cluster size
    const int *mem_row;    // sorted member rows (by cluster label, then row)
    const double *w;       // 1 / (1_c^T A 1_c), stored at the cluster's first position
};

// z_i = (M^-1 r)_i
__device__ __forceinline__ double precond_apply(const Precond &P, int i, double ri, const double *rvec) {
    double z = ri * P.dinv[i];
    const int s = P.pos ? P.pos[i] : -1;
    if (s >= 0) {
        const int st = P.seg_start[s], len = P.seg_len[s];
        double sum = 0.0;
        for (int k = 0; k < len; ++k) sum += rvec[P.mem_row[st + k]];
        z += P.w[st] * sum;
    }
    return z;
}

// same with r = r_old - alpha*Ap formed on the fly for the other members of the cluster
__device__ __forceinline__ double precond_apply_updated(const Precond &P, int i, double ri_new, const double *r_old,
                                                        const double *Ap, double alpha) {
    double z = ri_new * P.dinv[i];
    const int s = P.pos ? P.pos[i] : -1;
    if (s >= 0) {
        const int st = P.seg_start[s], len = P.seg_len[s];
        double sum = 0.0;
        for (int k = 0; k < len; ++k) {
            int j = P.mem_row[st + k];
            sum += r_old[j] - alpha * Ap[j];
        }
        z += P.w[st] * sum;
    }
    return z;
}

// flag[r] = 1 iff interior row r is an uncharged vacancy with an uncharged-vacancy interior neighbour
__global__ void cluster_mark_kernel(int m, int NL, const unsigned char *__restrict__ cls,
                                    const int *__restrict__ row_ptr, const int *__restrict__ col,
                                    int *__restrict__ flag, int *__restrict__ pos) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    int f = 0;
    if (cls[r + NL] == 2) {
        for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
            int c = col[p];
            if (c != r && cls[c + NL] == 2) { f = 1; break; }
        }
    }
    flag[r] = f;
    pos[r] = -1;
}

constexpr int kCompactThreads = 1024;
constexpr int kClusterThreads = 512;   // one CTA; small enough to fit beside the overlapped pairwise kernel

__global__ void __launch_bounds__(kCompactThreads) flag_count_kernel(int n, const int *__restrict__ flag,
                                                                    int *__restrict__ block_count) {
    __shared__ int sh[32];
    int i = blockIdx.x * kCompactThreads + threadIdx.x;
    int f = (i < n && flag[i]) ? 1 : 0;
    unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_count[blockIdx.x] = v;
    }
}

// ascending list of flagged indices; pos[idx] = position in the list
__global__ void __launch_bounds__(kCompactThreads) flag_scatter_kernel(int n, int nblocks, const int *__restrict__ flag,
                                                                      const int *__restrict__ block_count,
                                                                      const int *__restrict__ block_incl,
                                                                      int *__restrict__ list, int *__restrict__ pos,
                                                                      int *__restrict__ total) {
    __shared__ int sh[32];
    int i = blockIdx.x * kCompactThreads + threadIdx.x;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int f = (i < n && flag[i]) ? 1 : 0;
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
    if (f) {
        int p = block_incl[blockIdx.x] - block_count[blockIdx.x] + sh[w] + in_warp;
        list[p] = i;
        pos[i] = p;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *total = block_incl[nblocks - 1];
}

// One CTA: connected components of the clustered rows (min-label propagation with pointer
// jumping), bitonic sort of (label, position) keys, cluster segments, row -> sorted position.
__global__ void __launch_bounds__(kClusterThreads) cluster_build_kernel(int NL, const unsigned char *__restrict__ cls,
                                                             const int *__restrict__ row_ptr,
                                                             const int *__restrict__ col, const int *n_cl_ptr,
                                                             const int *list, int *pos, int *lab,
                                                             unsigned long long *keys, int *mem_row,
                                                             int *seg_start, int *seg_len) {
    const int n = *n_cl_ptr;
    if (n <= 0) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int t = tid; t < n; t += nt) lab[t] = t;
    __syncthreads();
    volatile int *vlab = lab;
    while (true) {
        int changed = 0;
        for (int t = tid; t < n; t += nt) {
            int l = vlab[t];
            const int row = list[t];
            for (int p = row_ptr[row]; p < row_ptr[row + 1]; ++p) {
                int c = col[p];
                if (c != row && cls[c + NL] == 2) {
                    int s = pos[c];
                    if (s >= 0) l = min(l, vlab[s]);
                }
            }
            l = min(l, vlab[l]);
            if (l < vlab[t]) { vlab[t] = l; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    int P = 1;
    while (P < n) P <<= 1;
    for (int t = tid; t < P; t += nt)
        keys[t] = t < n ? (((unsigned long long)(unsigned)lab[t] << 32) | (unsigned)t) : ~0ull;
    __syncthreads();
    volatile unsigned long long *vk = keys;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P; t += nt) {
                int q = t ^ j;
                if (q > t) {
                    unsigned long long a = vk[t], b = vk[q];
                    bool asc = (t & k) == 0;
                    if ((a > b) == asc) { vk[t] = b; vk[q] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int s = tid; s < n; s += nt) {
        unsigned long long key = vk[s];
        int t = (int)(key & 0xffffffffu);
        mem_row[s] = list[t];
    }
    __syncthreads();
    for (int s = tid; s < n; s += nt) {
        unsigned lbl = (unsigned)(vk[s] >> 32);
        bool start = (s == 0) || ((unsigned)(vk[s - 1] >> 32) != lbl);
        if (start) {
            int len = 1;
            while (s + len < n && (unsigned)(vk[s + len] >> 32) == lbl) ++len;
            for (int k = 0; k < len; ++k) { seg_start[s + k] = s; seg_len[s + k] = len; }
        }
        pos[mem_row[s]] = s;  // row -> sorted position (the list positions are no longer needed)
    }
}

// w[start] = 1 / (1_c^T A 1_c), one thread per cluster
__global__ void cluster_weight_kernel(const int *n_cl_ptr, const int *__restrict__ pos,
                                      const int *__restrict__ seg_start, const int *__restrict__ seg_len,
                                      const int *__restrict__ mem_row, const int *__restrict__ row_ptr,
                                      const int *__restrict__ col, const double *__restrict__ val,
                                      double *__restrict__ w) {
    const int n = *n_cl_ptr;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        if (seg_start[s] != s) continue;
        const int len = seg_len[s];
        double E = 0.0;
        for (int k = 0; k < len; ++k) {
            int row = mem_row[s + k];
            for (int p = row_ptr[row]; p < row_ptr[row + 1]; ++p) {
                int c = col[p];
                int sc = pos[c];
                if (sc >= 0 && seg_start[sc] == s) E += val[p];
            }
        }
        w[s] = E > 0.0 ? 1.0 / E : 0.0;
    }
}

// ---------------------------------------------------------------- CG vector kernels
// p = z = M^-1 r; rz = r.z; bb = b.M^-1 b; sets the stop threshold and clears the flags
__global__ void __launch_bounds__(kVecThreads) cg_init_kernel(int m, const double *r, const double *b, Precond P,
                                                             double *__restrict__ p, double tol, int max_iter,
                                                             double *partials, CgScalars *sc) {
    __shared__ double red[32];
    __shared__ double red2[32];
    double lrz = 0.0, lbb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        double ri = r[i], bi = b[i];
        double z = precond_apply(P, i, ri, r);
        p[i] = z;
        lrz += ri * z;
        lbb += bi * precond_apply(P, i, bi, b);
    }
    double t1 = block_sum(lrz, red);
    double t2 = block_sum(lbb, red2);
    // two deterministic grid sums sharing one arrival counter
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = t1;
        partials[gridDim.x + blockIdx.x] = t2;
        __threadfence();
        is_last = (atomicAdd(&sc->cnt_a, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double a1 = 0.0, a2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        a1 += __ldcg(partials + i);
        a2 += __ldcg(partials + gridDim.x + i);
    }
    a1 = block_sum(a1, red);
    a2 = block_sum(a2, red2);
    if (threadIdx.x == 0) {
        sc->rz = a1;
        sc->bb = a2;
        double ref = a2 > 0.0 ? a2 : a1;
        sc->stop = tol * tol * ref;
        sc->done = (a1 <= sc->stop) ? 1 : 0;
        sc->iters = 0;
        sc->max_iter = max_iter;
        sc->cnt_a = 0u;
    }
}

__device__ __forceinline__ double2 ld2s(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void st2s(double *p, double2 v) { __stcs(reinterpret_cast<double2 *>(p), v); }
__device__ __forceinline__ double2 ld2(const double *p, bool vec) {
    if (vec) return *reinterpret_cast<const double2 *>(p);
    return make_double2(p[0], p[1]);
}
__device__ __forceinline__ void st2(double *p, double2 v, bool vec) {
    if (vec) *reinterpret_cast<double2 *>(p) = v;
    else { p[0] = v.x; p[1] = v.y; }
}
__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// alpha = rz/pAp; x += alpha p; r_new = r_old - alpha Ap; rz_new = r_new . M^-1 r_new;
// last block: beta, convergence.  r is ping-ponged so that cluster members can be re-formed.
// Two elements per thread per pass, 16-byte loads.
__global__ void __launch_bounds__(kVecThreads) cg_update_kernel(int m, double *x, const double *r_old,
                                                               double *r_new, const double *p,
                                                               const double *Ap, Precond P, double *partials,
                                                               CgScalars *sc, int flags) {
    __shared__ double red[32];
    if (sc->done) return;
    const double alpha = sc->rz / sc->pAp;
    const bool vx = aligned16(x);
    const bool stream = (flags & 2) != 0;
    double local = 0.0;
    const int n2 = (m + 1) >> 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += gridDim.x * blockDim.x) {
        const int i = 2 * j;
        if (i + 1 < m) {
            double2 p2 = ld2(p + i, true), r2 = stream ? ld2s(r_old + i) : ld2(r_old + i, true),
                    a2 = stream ? ld2s(Ap + i) : ld2(Ap + i, true), x2 = (stream && vx) ? ld2s(x + i) : ld2(x + i, vx);
            x2.x += alpha * p2.x; x2.y += alpha * p2.y;
            r2.x -= alpha * a2.x; r2.y -= alpha * a2.y;
            if (stream && vx) st2s(x + i, x2); else st2(x + i, x2, vx);
            st2(r_new + i, r2, true);
            local += r2.x * precond_apply_updated(P, i, r2.x, r_old, Ap, alpha);
            local += r2.y * precond_apply_updated(P, i + 1, r2.y, r_old, Ap, alpha);
        } else {
            x[i] += alpha * p[i];
            double ri = r_old[i] - alpha * Ap[i];
            r_new[i] = ri;
            local += ri * precond_apply_updated(P, i, ri, r_old, Ap, alpha);
        }
    }
    double tot = block_sum(local, red);
    if (grid_sum_finish(tot, partials, &sc->cnt_b, &sc->rz_new, red)) {
        double rzn = sc->rz_new;
        sc->alpha = alpha;
        sc->beta = rzn / sc->rz;
        sc->rz = rzn;
        sc->iters += 1;
        if (rzn <= sc->stop || sc->iters >= sc->max_iter || !(rzn == rzn)) sc->done = 1;
    }
}

// p = M^-1 r + beta p   (skipped once converged)
__global__ void __launch_bounds__(kVecThreads) cg_direction_kernel(int m, const double *r, Precond P, double *p,
                                                                  const CgScalars *sc) {
    if (sc->done) return;
    const double beta = sc->beta;
    const int n2 = (m + 1) >> 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += gridDim.x * blockDim.x) {
        const int i = 2 * j;
        if (i + 1 < m) {
            double2 r2 = ld2(r + i, true), p2 = ld2(p + i, true);
            p2.x = precond_apply(P, i, r2.x, r) + beta * p2.x;
            p2.y = precond_apply(P, i + 1, r2.y, r) + beta * p2.y;
            st2(p + i, p2, true);
        } else {
            p[i] = precond_apply(P, i, r[i], r) + beta * p[i];
        }
    }
}

// after a residual replacement: rz := r_true . M^-1 r_true decides convergence
__global__ void replace_scalars_kernel(CgScalars *sc) {
    if (threadIdx.x != 0) return;
    if (sc->iters >= sc->max_iter) { sc->done = 1; return; }
    sc->rz = sc->rz_new;
    sc->done = (sc->rz_new <= sc->stop) ? 1 : 0;
}

// out = v . M^-1 v   (true-residual norm in the preconditioner's metric)
__global__ void __launch_bounds__(kVecThreads) precond_norm_kernel(int m, const double *v, Precond P, double *partials,
                                                                  unsigned int *counter, double *out) {
    __shared__ double red[32];
    double local = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
        local += v[i] * precond_apply(P, i, v[i], v);
    double tot = block_sum(local, red);
    grid_sum_finish(tot, partials, counter, out, red);
}

__global__ void axpy_kernel(int m, double a, const double *__restrict__ xin, double *__restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) y[i] += a * xin[i];
}

__global__ void fill_kernel(int n, double v, double *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = v;
}

__global__ void diag_inverse_kernel(int m, const int *__restrict__ row_ptr, const int *__restrict__ col,
                                    const double *__restrict__ val, double *__restrict__ dinv) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double d = 1.0;
    for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p)
        if (col[p] == r) d = val[p];
    dinv[r] = 1.0 / d;
}

// ---------------------------------------------------------------- host side
static int get_tiling(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int4 **tile_row, int *num_tiles) {
    SpmvTiling &t = ctx->tiling;
    if (t.row_ptr != d_row_ptr || t.m != m || t.nnz != nnz) {
        if (t.d_tile_row) {
            DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
            DKMC_CUDA(cudaFree(t.d_tile_row));
            t.d_tile_row = nullptr;
        }
        t.num_tiles = ceil_div(nnz > 0 ? nnz : 1, kSpmvTile);
        DKMC_CUDA(cudaMalloc(&t.d_tile_row, ((size_t)t.num_tiles + 1) * sizeof(int4)));
        DKMC_LAUNCH(ctx, tile_rows_kernel, ceil_div(t.num_tiles + 1, 256), 256, 0, m, t.num_tiles, d_row_ptr,
                    reinterpret_cast<int4 *>(t.d_tile_row));
        t.row_ptr = d_row_ptr; t.m = m; t.nnz = nnz;
    }
    *tile_row = reinterpret_cast<const int4 *>(t.d_tile_row);
    *num_tiles = t.num_tiles;
    return DKMC_OK;
}

static int vec_grid(const dkmc_ctx *ctx, int m) {
    // CTAs per SM of the grid-stride vector kernels: 5 = what is resident at once (48 registers x 256 threads),
    // so every CTA makes 2-3 passes and the fixed cost per CTA (scalar loads, block reduction, ticket) and the
    // final sum over the partials shrink 2.6x.  Measured at 1 M sites (tools/vec_grid_experiment.py):
    // 106 us per CG iteration instead of 111 (one pass per thread, 16 per SM), 152 instead of 157 overlapped.
    static const int cps = [] { const char *e = getenv("DKMC_VEC_CPS"); int v = e ? atoi(e) : 5; return v < 1 ? 1 : (v > 16 ? 16 : v); }();
    int g = ceil_div((m + 1) / 2, kVecThreads);
    int cap = ctx->num_sms * cps;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

struct CgWork {
    double *r[2], *p, *Ap, *dinv, *res, *e, *partials;
    CgScalars *sc;
    const int4 *tile_row;
    int num_tiles;
    Precond P;
    int n_cl;   // clustered rows (length of the sorted member list), host copy
};

static int cg_workspace(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, CgWork *w) {
    int rc;
    if ((rc = ensure<double>(ctx, S_CG_R, m, &w->r[0]))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_Z, m, &w->r[1]))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_P, m, &w->p))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_AP, m, &w->Ap))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_DINV, m, &w->dinv))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_RES, m, &w->res))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_E, m, &w->e))) return rc;
    if ((rc = get_tiling(ctx, m, nnz, d_row_ptr, &w->tile_row, &w->num_tiles))) return rc;
    size_t np = (size_t)(w->num_tiles > ctx->num_sms * 16 ? w->num_tiles : ctx->num_sms * 16) * 2 + 64;
    if ((rc = ensure<double>(ctx, S_PARTIALS, np, &w->partials))) return rc;
    bool fresh = ctx->slot_ptr[S_SCALARS] == nullptr;
    if ((rc = ensure<CgScalars>(ctx, S_SCALARS, 4, &w->sc))) return rc;
    if (fresh) DKMC_CUDA(cudaMemsetAsync(w->sc, 0, 4 * sizeof(CgScalars), ctx->stream));
    w->P = Precond{w->dinv, nullptr, nullptr, nullptr, nullptr, nullptr};
    w->n_cl = 0;
    return DKMC_OK;
}

// Detects the uncharged-vacancy clusters of the current state and fills w->P.
static int build_clusters(dkmc_ctx *ctx, int m, int NL, const unsigned char *cls, const int *d_row_ptr,
                          const int *d_col, const double *d_val, CgWork *w) {
    int *ints;
    unsigned long long *keys;
    double *wts;
    int rc;
    const int nb = ceil_div(m, kCompactThreads);
    size_t P2 = 1;
    while (P2 < (size_t)m) P2 <<= 1;
    // flag | pos | list | lab | mem_row | seg_start | seg_len | block_count | block_incl | total
    if ((rc = ensure<int>(ctx, S_CL_INT, (size_t)7 * m + 2 * (size_t)nb + 8, &ints))) return rc;
    if ((rc = ensure<unsigned long long>(ctx, S_CL_KEYS, P2, &keys))) return rc;
    if ((rc = ensure<double>(ctx, S_CL_W, m, &wts))) return rc;
    int *tmp;
    if ((rc = ensure<int>(ctx, S_SCAN_BLOCK, (size_t)ceil_div(nb, kScanTile) + 1, &tmp))) return rc;
    int *flag = ints, *pos = ints + m, *list = ints + 2 * (size_t)m, *lab = ints + 3 * (size_t)m,
        *mem_row = ints + 4 * (size_t)m, *seg_start = ints + 5 * (size_t)m, *seg_len = ints + 6 * (size_t)m,
        *bcount = ints + 7 * (size_t)m, *bincl = bcount + nb, *total = bincl + nb;
    DKMC_LAUNCH(ctx, cluster_mark_kernel, ceil_div(m, 256), 256, 0, m, NL, cls, d_row_ptr, d_col, flag, pos);
    DKMC_LAUNCH(ctx, flag_count_kernel, nb, kCompactThreads, 0, m, flag, bcount);
    if ((rc = inclusive_scan<int>(ctx, bcount, nb, bincl, tmp))) return rc;
    DKMC_LAUNCH(ctx, flag_scatter_kernel, nb, kCompactThreads, 0, m, nb, flag, bcount, bincl, list, pos, total);
    DKMC_LAUNCH(ctx, cluster_build_kernel, 1, kClusterThreads, 0, NL, cls, d_row_ptr, d_col, total, list, pos, lab, keys, mem_row,
                seg_start, seg_len);
    DKMC_LAUNCH(ctx, cluster_weight_kernel, 64, 128, 0, total, pos, seg_start, seg_len, mem_row, d_row_ptr, d_col, d_val, wts);
    w->P = Precond{w->dinv, pos, seg_start, seg_len, mem_row, wts};
    // the persistent PCG sizes its reduction payload by the number of clustered rows
    DKMC_CUDA(cudaMemcpyAsync(&w->n_cl, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return DKMC_OK;
}

// SpMV over `ntiles` tiles starting at `tile_info` (register-staged CSR tiles).  Measured at 1 M sites: 67 us
// alone (75 % of the measured HBM peak).  Round 1 also carried a TMA-fed, a packed (4 B per non-zero) and a
// window-staged (2.6 B per non-zero) variant, all bit-identical and all slower inside the CG — the kernel is
// bound by the 26 M gathers of x, not by what it streams; they were removed in round 2.
template <int MODE>
static int launch_spmv(dkmc_ctx *ctx, int ntiles, int m, int nnz, const int *d_row_ptr, const int *d_col,
                       const double *d_val, const double *d_x, double *d_y, const int4 *tile_info, const double *w,
                       const double *dinv, double *partials, unsigned int *counter, double *dot_out,
                       const int *done_flag) {
    (void)m;
    if (ntiles <= 0) return DKMC_OK;
    int grid = ctx->num_sms * 6;
    if (MODE == 0 || grid > ntiles) grid = ntiles;  // no dot to finish: one tile per CTA
    DKMC_LAUNCH(ctx, spmv_tile_kernel<MODE>, grid, kSpmvThreads, 0, ntiles, nnz, d_row_ptr, d_col, d_val, d_x, d_y,
                tile_info, w, dinv, partials, counter, dot_out, done_flag, g_flags);
    return DKMC_OK;
}

// the persistent-kernel PCG (pcg_persistent.cuh) on one GPU; defined below, after the peer-window plumbing
static bool use_persistent_pcg(const dkmc_ctx *ctx);
static int run_pcg_persistent_single(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col,
                                     const double *d_val, const double *d_b, double *d_x, const CgWork &w, double tol,
                                     int max_iter, int *iters_out, int *converged, double *bb_out);

// Preconditioned CG on A x = b starting from x (in/out).  w.P must be set.
static int run_pcg(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col, const double *d_val,
                   const double *d_b, double *d_x, const CgWork &w, double tol, int max_iter, int check_every,
                   int *iters_out, int *converged, double *bb_out) {
    int persistent_its = 0;
    if (use_persistent_pcg(ctx)) {
        int rcp = run_pcg_persistent_single(ctx, m, nnz, d_row_ptr, d_col, d_val, d_b, d_x, w, tol, max_iter, iters_out, converged,
                                            bb_out);
        if (rcp != DKMC_OK || *converged || *iters_out < max_iter) return rcp;
        // max_iter reached without convergence (seen once: floating metal islands): do not give up
        // before the textbook two-reduction recurrence of the per-operation path below has had its turn, from the
        // current x
        persistent_its = *iters_out;
    }
    // per-operation path (DKMC_LEGACY_CG=1, or the fall-back above): three launches per iteration, the host polls a
    // flag every `check_every`
    const int vg = vec_grid(ctx, m);
    // r = b - A x, then z/p/rz/bb
    int rc0;
    if ((rc0 = launch_spmv<2>(ctx, w.num_tiles, m, nnz, d_row_ptr, d_col, d_val, d_x, w.r[0], w.tile_row, d_b, w.dinv,
                              w.partials, &w.sc->cnt_c, &w.sc->resnorm2, nullptr))) return rc0;
    DKMC_LAUNCH(ctx, cg_init_kernel, vg, kVecThreads, 0, m, w.r[0], d_b, w.P, w.p, tol, max_iter, w.partials, w.sc);
    CgScalars h;
    memset(&h, 0, sizeof(h));
    int launched = 0, cur = 0;
    const bool replace = (g_flags & 8) != 0;
    if (check_every < 1) check_every = 1;
    while (true) {
        for (int k = 0; k < check_every; ++k) {
            if ((rc0 = launch_spmv<1>(ctx, w.num_tiles, m, nnz, d_row_ptr, d_col, d_val, w.p, w.Ap, w.tile_row, w.p, nullptr,
                                      w.partials, &w.sc->cnt_c, &w.sc->pAp, &w.sc->done))) return rc0;
            DKMC_LAUNCH(ctx, cg_update_kernel, vg, kVecThreads, 0, m, d_x, w.r[cur], w.r[cur ^ 1], w.p, w.Ap, w.P,
                        w.partials, w.sc, g_flags);
            cur ^= 1;
            DKMC_LAUNCH(ctx, cg_direction_kernel, vg, kVecThreads, 0, m, w.r[cur], w.P, w.p, w.sc);
        }
        launched += check_every;
        if (replace) {
            // residual replacement: swap the drifting recurrence residual for the true one
            // (double-double), keep the search direction; convergence is then judged on the
            // TRUE residual, so no restart is needed to reach rounding-level accuracy
            DKMC_LAUNCH(ctx, residual_dd_kernel, w.num_tiles, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, d_x, d_b, w.dinv,
                        w.r[cur], w.tile_row, w.partials, &w.sc->cnt_d, &w.sc->resnorm2);
            DKMC_LAUNCH(ctx, precond_norm_kernel, vg, kVecThreads, 0, m, w.r[cur], w.P, w.partials, &w.sc->cnt_d,
                        &w.sc->rz_new);
            DKMC_LAUNCH(ctx, replace_scalars_kernel, 1, 32, 0, w.sc);
        }
        DKMC_CUDA(cudaMemcpyAsync(&h, w.sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, ctx->stream));
        DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h.done || launched >= max_iter) break;
    }
    *iters_out = h.iters + persistent_its;
    *converged = (h.rz <= h.stop) ? 1 : 0;
    if (bb_out) *bb_out = h.bb;
    return DKMC_OK;
}

// max_i |(M^-1 v)_i| and max_i |x_i| (bit patterns of non-negative doubles order like integers,
// so atomicMax is exact and order-independent)
__global__ void __launch_bounds__(kVecThreads) inf_norms_kernel(int m, const double *v, Precond P,
                                                               const double *__restrict__ x,
                                                               unsigned long long *out) {
    double mz = 0.0, mx = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        mz = fmax(mz, fabs(precond_apply(P, i, v[i], v)));
        mx = fmax(mx, fabs(x[i]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        mz = fmax(mz, __shfl_xor_sync(0xffffffffu, mz, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, (unsigned long long)__double_as_longlong(mz));
        atomicMax(out + 1, (unsigned long long)__double_as_longlong(mx));
    }
}

// True residual res = b - A x in double-double.  rel = ||res||_M^-1 / ||b||_M^-1 (energy scale);
// est = max|M^-1 res| / max|x|: a per-entry error estimate.  K spans conductances 1 and 1e-8, so
// the energy norm alone says nothing about the oxide entries (they weigh 1e-8 in it); est is what
// decides whether another refinement round is needed.
static int true_residual(dkmc_ctx *ctx, int m, const int *d_row_ptr, const int *d_col, const double *d_val,
                         const double *d_rhs, const double *d_x, const CgWork &w, double bb, double *rel,
                         double *est) {
    DKMC_LAUNCH(ctx, residual_dd_kernel, w.num_tiles, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, d_x, d_rhs, w.dinv,
                w.res, w.tile_row, w.partials, &w.sc->cnt_d, &w.sc->resnorm2);
    DKMC_LAUNCH(ctx, precond_norm_kernel, vec_grid(ctx, m), kVecThreads, 0, m, w.res, w.P, w.partials, &w.sc->cnt_d,
                &w.sc->resnorm2);
    unsigned long long *mx = reinterpret_cast<unsigned long long *>(w.sc + 1);
    DKMC_CUDA(cudaMemsetAsync(mx, 0, 2 * sizeof(unsigned long long), ctx->stream));
    DKMC_LAUNCH(ctx, inf_norms_kernel, vec_grid(ctx, m), kVecThreads, 0, m, w.res, w.P, d_x, mx);
    double h = 0.0, hm[2] = {0.0, 0.0};
    DKMC_CUDA(cudaMemcpyAsync(&h, &w.sc->resnorm2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaMemcpyAsync(hm, mx, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    *rel = bb > 0 ? sqrt(h / bb) : sqrt(h);
    *est = hm[1] > 0 ? hm[0] / hm[1] : hm[0];
    return DKMC_OK;
}

// Tolerance of a restart's correction solve: a restart only has to bring the per-entry error estimate from
// `est` below est_tol, so it is solved to est_tol / (kRestartMargin * est) relative to its own right-hand side,
// never tighter than refine_tol and never looser than 1e-2 (DKMC_RESTART_MARGIN=0: always refine_tol, the round-1
// behaviour).  If one restart falls short the next one finishes the job (refine_rounds).
static double restart_tolerance(const dkmc_solver_opts &o, double est) {
    static const double margin = [] { const char *e = getenv("DKMC_RESTART_MARGIN"); return e ? atof(e) : 30.0; }();
    if (!(margin > 0.0) || !(est > 0.0)) return o.refine_tol;
    double t = o.est_tol / (margin * est);
    if (t < o.refine_tol) t = o.refine_tol;
    if (t > 1e-2) t = 1e-2;
    return t;
}

static int solve_refined(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col,
                         const double *d_val, const double *d_rhs, double *d_x, CgWork &w,
                         const dkmc_solver_opts &o, dkmc_solve_info *info) {
    int iters = 0, conv = 0, total = 0, rc;
    double bb0 = 0.0, rel = 0.0, est = 0.0;
    if ((rc = run_pcg(ctx, m, nnz, d_row_ptr, d_col, d_val, d_rhs, d_x, w, o.rel_tol, o.max_iter, o.check_every, &iters, &conv, &bb0))) return rc;
    total += iters;
    bool all_conv = conv != 0;
    const int vg = vec_grid(ctx, m);
    // The recurrence residual drifts from the true one and CG's energy norm is blind to the
    // low_G part of the device: restart on the double-double residual (each restart is solved
    // relative to ITS OWN right-hand side, i.e. on the scale of what is still wrong) until the
    // per-entry error estimate is at rounding level or the rounds are used up.
    int rounds = 0;
    static const bool trace = getenv("DKMC_SOLVE_TRACE") != nullptr;
    if ((rc = true_residual(ctx, m, d_row_ptr, d_col, d_val, d_rhs, d_x, w, bb0, &rel, &est))) return rc;
    if (trace) fprintf(stderr, "dkmc solve: first %d its -> rel %.2e est %.2e\n", iters, rel, est);
    while (rounds < o.refine_rounds && est > o.est_tol) {
        DKMC_LAUNCH(ctx, fill_kernel, vg, kVecThreads, 0, m, 0.0, w.e);
        const double rtol = restart_tolerance(o, est);
        if ((rc = run_pcg(ctx, m, nnz, d_row_ptr, d_col, d_val, w.res, w.e, w, rtol, o.max_iter, o.check_every, &iters, &conv, nullptr))) return rc;
        if (trace) fprintf(stderr, "dkmc solve: restart tol %.1e: %d its\n", rtol, iters);
        total += iters;
        DKMC_LAUNCH(ctx, axpy_kernel, vg, kVecThreads, 0, m, 1.0, w.e, d_x);
        ++rounds;
        if ((rc = true_residual(ctx, m, d_row_ptr, d_col, d_val, d_rhs, d_x, w, bb0, &rel, &est))) return rc;
    }
    if (info) {
        info->iterations = total;
        info->refinements = rounds;
        info->rel_residual = rel;
        info->est_error = est;
    }
    return (all_conv || est <= o.est_tol) ? DKMC_OK : DKMC_ERR_NOT_CONVERGED;
}

// ================================================================ distributed PCG (x-slab rows per rank)
// One process per GPU.  Rank r owns the interior rows [row_begin, row_end) (whole SpMV tiles) of
// the full-length vectors; every iteration moves the halo of p (contiguous index ranges, to the
// x-neighbours for x-major ordered sites) with ncclSend/ncclRecv and two small all-reduces:
// [p.Ap] and [r.D^-1 r, W^T r] — the cluster coarse space is applied from the all-reduced cluster
// sums, so clusters may straddle slab faces.  K assembly and the cluster detection are replicated
// (0.4 ms + 0.4 ms at 1 M sites); the solve itself is partitioned.
// Peer-memory exchange for the per-iteration traffic of the distributed PCG (one process per GPU,
// windows shared through CUDA IPC, NVLink peer stores).  Every rank owns a window
//   [ p vector (m doubles) | reduction slots [2][world][kP2pRedCap] | flags [4][world] ]
// that all ranks map.  An all-reduce is ONE small kernel per rank: push my partial sums into every
// rank's slot, fence, raise my flag everywhere, wait for everybody's flag here, add the slots in rank
// order (deterministic, identical on all ranks).  The halo of p is pushed into the neighbours' p
// vectors the same way.  Slots and flags are double-buffered by the parity of a running operation
// count; a rank can only start operation n+2 after every rank has raised its flag for n+1, i.e.
// after everybody finished reading the slots of n.  Latency: one NVLink store + flag round (~3 us)
// instead of ~20-40 us per NCCL call, of which the PCG needs three per iteration.
constexpr int kP2pRedCap = 1 << 15;               // doubles per rank and reduction
constexpr long long kP2pTimeoutCycles = 40000000000ll;  // ~20 s: a lost peer becomes an error, not a hang (a peer that
                                                         // is merely late — host work between two steps — must not)

struct P2pPeers {
    unsigned char *base[DKMC_MAX_RANKS];
    int world, rank;
    size_t red_off, red2_off, flag_off;   // reduction slots of the per-op kernels / of the persistent PCG, flags
    size_t vec2_off;                      // second gather vector (pipelined PCG)
};

// window = [ vector (m doubles) | second vector | slots [2][world][kP2pRedCap] | the same again | flags [8][DKMC_MAX_RANKS] ]
static size_t window_layout(int m, int world, P2pPeers *P) {
    const size_t pbytes = (((size_t)m + 64) * sizeof(double) + 255) & ~(size_t)255;
    const size_t rbytes = (size_t)2 * world * kP2pRedCap * sizeof(double);
    const size_t fbytes = (size_t)8 * DKMC_MAX_RANKS * sizeof(unsigned long long);
    P->vec2_off = pbytes;
    P->red_off = 2 * pbytes;
    P->red2_off = 2 * pbytes + rbytes;
    P->flag_off = 2 * pbytes + 2 * rbytes;
    return 2 * pbytes + 2 * rbytes + fbytes;
}

struct DistState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // peer-memory windows (dkmc_dist_p2p_alloc / _open)
    unsigned char *win = nullptr;
    size_t win_bytes = 0;
    int m_cap = 0;
    bool p2p = false;
    unsigned long long seq = 0;    // reductions (slots and flags 0/1 by parity)
    unsigned long long hseq = 0;   // halo exchanges (flags 2/3 by parity)
    unsigned long long pseq_r = 0, pseq_h = 0;   // the persistent PCG's own sequences (second slot bank, flags 4-7)
    P2pPeers peers;
};

// one GPU: the persistent PCG runs the same protocol against a "window" in ordinary device memory
struct SelfWindow {
    unsigned char *base = nullptr;
    int m_cap = 0;
    unsigned long long pseq_r = 0, pseq_h = 0;
    P2pPeers peers;
};
static DistState *dist_of(dkmc_ctx *ctx) { return static_cast<DistState *>(ctx->dist); }

// flags are written with release and polled with acquire semantics at system scope: ordering comes
// from the flag accesses themselves, not from (much more expensive) system-wide fences
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long *p2p_flag(const P2pPeers &P, int at_rank, int buf, int of_rank) {
    return reinterpret_cast<unsigned long long *>(P.base[at_rank] + P.flag_off) + (size_t)buf * P.world + of_rank;
}

// every rank waits until all ranks have raised flag `buf` to `seq` here; lane r watches rank r
__device__ __forceinline__ void p2p_wait_all(const P2pPeers &P, int buf, unsigned long long seq, int *err) {
    for (int r = threadIdx.x; r < P.world; r += blockDim.x) {
        const unsigned long long *f = p2p_flag(P, P.rank, buf, r);
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > kP2pTimeoutCycles) { *err = 1; break; }
        }
    }
}

struct P2pHalo {
    int n_send, n_recv;
    int send_peer[DKMC_MAX_HALO_SEGMENTS], send_begin[DKMC_MAX_HALO_SEGMENTS], send_end[DKMC_MAX_HALO_SEGMENTS];
    int recv_peer[DKMC_MAX_HALO_SEGMENTS];
};

#include "pcg_persistent.cuh"
#include "pcg_pipelined.cuh"

static bool use_persistent_pcg(const dkmc_ctx *ctx) { return !ctx->legacy_cg; }
constexpr int kPcgProfLen = 8 + 2 * 2048;   // phase sums of CTA 0 | per CTA: SpMV-phase time, wait at barrier R

// CTAs per SM of the persistent PCG.  All CTAs must be resident together (they meet at grid barriers), so the
// grid is sized to what fits: beside the overlapped pairwise kernel that is what its bounded residency leaves.
// Three instantiations trade registers per thread (loads in flight) against CTAs per SM: *variant = 0, 1, 2 for
// launch bounds of 6, 5, 4 CTAs per SM (40, 48, 64 registers).
typedef void (*PcgKernel)(const PcgArgs);
// family 0: Chronopoulos-Gear (two synchronisations per iteration), family 1: pipelined (one).  Variants 0, 1, 2 =
// launch bounds of 6, 5, 4 CTAs per SM (40, 48, 64 registers); variant 3 = 8 CTAs (32 registers), an experiment that
// gained nothing beside the pairwise kernel (5 instead of 4 resident CTAs, more spills) and is never chosen by default.
static const PcgKernel kPcgKernels[2][4] = {
    {pcg_persistent_kernel<6, false>, pcg_persistent_kernel<5, false>, pcg_persistent_kernel<4, false>, pcg_persistent_kernel<8, false>},
    {pcg_pipelined_kernel<6, false>, pcg_pipelined_kernel<5, false>, pcg_pipelined_kernel<4, false>, pcg_pipelined_kernel<8, false>}};
static const PcgKernel kPcgKernelsProf[2][4] = {
    {pcg_persistent_kernel<6, true>, pcg_persistent_kernel<5, true>, pcg_persistent_kernel<4, true>, pcg_persistent_kernel<8, true>},
    {pcg_pipelined_kernel<6, true>, pcg_pipelined_kernel<5, true>, pcg_pipelined_kernel<4, true>, pcg_pipelined_kernel<8, true>}};

static int pcg_ctas_per_sm(dkmc_ctx *ctx, int family, int *variant) {
    static int occ_all[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, regs_all[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    int *occ = occ_all[family], *regs = regs_all[family];
    if (!occ[0]) {
        for (int v = 0; v < 4; ++v) {
            cudaFuncAttributes fa;
            regs[v] = cudaFuncGetAttributes(&fa, kPcgKernels[family][v]) == cudaSuccess ? fa.numRegs : 40 + 8 * v;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[v], kPcgKernels[family][v], kSpmvThreads, 0) != cudaSuccess || occ[v] < 1) occ[v] = 1;
        }
    }
    // DKMC_PCG_CPS="alone_cps,alone_variant,overlap_cps,overlap_variant" overrides (0 / -1: default)
    static int cfg[4] = {0, -1, 0, -1};
    static bool cfg_read = false;
    if (!cfg_read) { const char *e = getenv("DKMC_PCG_CPS"); if (e) sscanf(e, "%d,%d,%d,%d", &cfg[0], &cfg[1], &cfg[2], &cfg[3]); cfg_read = true; }
    // beside the pairwise kernel only while that kernel is still running: a restart launched after the sum has
    // finished (short sums: small devices, several GPUs) gets the whole GPU
    const bool overlapped = ctx->pw_pending.active && cudaEventQuery(ctx->ev_pw1) == cudaErrorNotReady;
    int v = 0;
    const int cv = cfg[overlapped ? 3 : 1];
    if (cv >= 0 && cv < 4) v = cv;
    int cps = occ[v];
    if (overlapped) {
        // registers and threads the pairwise CTAs hold on every SM (allocation granularity: 8 registers per thread)
        const int pw_threads = ctx->pw_side_blocks_per_sm * ctx->pw_side_threads;
        const int pw_regs = pw_threads * ((ctx->pw_regs_per_thread + 7) & ~7);
        const int mine = kSpmvThreads * ((regs[v] + 7) & ~7);
        int fit = (65536 - pw_regs) / (mine > 0 ? mine : 1);
        const int fit_threads = (2048 - pw_threads) / kSpmvThreads;
        if (fit > fit_threads) fit = fit_threads;
        if (fit < cps) cps = fit;
    }
    const int want = cfg[overlapped ? 2 : 0];
    if (want > 0 && want < cps) cps = want;
    *variant = v;
    return cps < 1 ? 1 : cps;
}

struct PcgGeometry {
    int ra, rb, t0, t1, n_cl;
    int row_end_all[DKMC_MAX_RANKS];
    P2pPeers peers;
    P2pHalo halo;
    unsigned long long *pseq_r, *pseq_h;
};

static int run_pcg_persistent_family(dkmc_ctx *ctx, int family, int m, const int *d_row_ptr, const int *d_col, const double *d_val,
                                     const double *d_b, double *d_x, const CgWork &w, const PcgGeometry &geo, double tol,
                                     int max_iter, int *iters_out, int *converged, double *bb_out, int *fallback);

// The persistent-kernel PCG: the pipelined kernel (one synchronisation per iteration) unless it is switched off
// (dkmc_ctx_set_pcg_pipelined / DKMC_PCG_PIPELINED=0) or some CTA's rows touch more clusters than its shared-memory
// table holds — then, on every rank alike, the Chronopoulos-Gear kernel.
static int run_pcg_persistent(dkmc_ctx *ctx, int m, const int *d_row_ptr, const int *d_col, const double *d_val,
                              const double *d_b, double *d_x, const CgWork &w, const PcgGeometry &geo, double tol,
                              int max_iter, int *iters_out, int *converged, double *bb_out) {
    // auto (-1): on several GPUs only — on one GPU there is no NVLink round trip to hide and the extra vector costs
    // more than the second barrier (measured at 1 M sites: 150 vs 145 us per iteration beside the pairwise sum).
    // Never without the cluster coarse space: the pipelined recurrences stagnate on the undeflated matrix.
    // Only for the restarts' correction solves (tolerance 1e-6 .. 1e-2 relative to their own right-hand side, far
    // above the floor of the pipelined recurrences): the first solve of a step runs to 1e-12, where the pipelined
    // recurrence residual stagnates in about one step out of six (measured, 1 M sites) — and the restarts are two
    // thirds of a step's iterations.
    // auto: from four GPUs on (at two the hidden round trip is worth less than the restarts' extra iterations:
    // 42.3 against 44.3 KMC steps/s at 1 M sites)
    const bool want = (ctx->pcg_pipelined < 0 ? geo.peers.world >= 4 : ctx->pcg_pipelined != 0) && tol >= 1e-8;
    int done_its = 0;
    if (want && w.P.pos != nullptr) {
        int fallback = 0;
        int rc = run_pcg_persistent_family(ctx, 1, m, d_row_ptr, d_col, d_val, d_b, d_x, w, geo, tol, max_iter, iters_out, converged,
                                           bb_out, &fallback);
        if (rc != DKMC_OK) return rc;
        if (!fallback && *converged) return rc;
        // table overflow, stagnation or max_iter: the other recurrence carries on from the current x
        if (!fallback) done_its = *iters_out;
    }
    int rc = run_pcg_persistent_family(ctx, 0, m, d_row_ptr, d_col, d_val, d_b, d_x, w, geo, tol, max_iter, iters_out, converged, bb_out,
                                       nullptr);
    *iters_out += done_its;
    return rc;
}

static int run_pcg_persistent_family(dkmc_ctx *ctx, int family, int m, const int *d_row_ptr, const int *d_col, const double *d_val,
                                     const double *d_b, double *d_x, const CgWork &w, const PcgGeometry &geo, double tol,
                                     int max_iter, int *iters_out, int *converged, double *bb_out, int *fallback) {
    PcgArgs a;
    memset(&a, 0, sizeof(a));
    int rc;
    const int n = geo.n_cl;
    DKMC_REQUIRE((size_t)4 + 2 * (size_t)n <= (size_t)kPcgMaxPayload, "too many clustered rows for a reduction slot");
    double *s_vec, *rec, *payload;
    PcgSync *sync;
    if ((rc = ensure<double>(ctx, S_CG_S, (size_t)m, &s_vec))) return rc;
    if ((rc = ensure<double>(ctx, S_CL_REC, (size_t)4 * n + 8, &rec))) return rc;
    if ((rc = ensure<double>(ctx, S_DIST_RED, (size_t)4 + 2 * (size_t)n + 8, &payload))) return rc;
    int *clk;
    if ((rc = ensure<int>(ctx, S_PCG_CLK, (size_t)2 * n + 8, &clk))) return rc;
    // one per sequence owner (the single-GPU window and the distributed one count separately)
    const int sync_slot = geo.peers.world > 1 ? S_PCG_SYNC2 : S_PCG_SYNC;
    const bool fresh = ctx->slot_ptr[sync_slot] == nullptr;
    if ((rc = ensure<PcgSync>(ctx, sync_slot, 1, &sync))) return rc;
    if (fresh) DKMC_CUDA(cudaMemsetAsync(sync, 0, sizeof(PcgSync), ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(&sync->n_global, 0, sizeof(unsigned int), ctx->stream));
    static const bool want_prof = getenv("DKMC_PCG_PROF") != nullptr;
    long long *prof = nullptr;
    if (want_prof) {
        const bool pfresh = ctx->slot_ptr[S_PCG_PROF] == nullptr;
        if ((rc = ensure<long long>(ctx, S_PCG_PROF, kPcgProfLen, &prof))) return rc;
        if (pfresh) DKMC_CUDA(cudaMemsetAsync(prof, 0, kPcgProfLen * sizeof(long long), ctx->stream));
    }
    a.m = m; a.ra = geo.ra; a.rb = geo.rb; a.t0 = geo.t0; a.t1 = geo.t1; a.n_cl = n; a.max_iter = max_iter;
    a.row_ptr = d_row_ptr; a.col = d_col; a.val = d_val; a.dinv = w.dinv; a.b = d_b; a.tile_info = w.tile_row;
    a.x = d_x; a.r = w.r[0]; a.w = w.Ap; a.p = w.p; a.s = s_vec;
    if (family == 1) {
        if ((rc = ensure<double>(ctx, S_CG_PZ, (size_t)m, &a.z))) return rc;
        if ((rc = ensure<double>(ctx, S_CG_PN, (size_t)m, &a.nvec))) return rc;
        a.g2_off = geo.peers.vec2_off;
    }
    a.P = w.P;
    a.cs = rec; a.cr = rec + 2 * (size_t)n;
    a.cl_kind = clk; a.gl_list = clk + n + 4;
    for (int q = 0; q < DKMC_MAX_RANKS; ++q) a.row_end_all[q] = geo.row_end_all[q];
    a.payload = payload; a.partials = w.partials; a.sync = sync; a.sc = w.sc; a.tol = tol;
    a.rseq0 = *geo.pseq_r; a.hseq0 = *geo.pseq_h;
    a.peers = geo.peers; a.halo = geo.halo; a.prof = prof;
    const int rows = geo.rb - geo.ra, nt = geo.t1 - geo.t0;
    int variant = 0;
    int grid = ctx->num_sms * pcg_ctas_per_sm(ctx, family, &variant);
    int need = ceil_div(rows > 0 ? rows : 1, kSpmvThreads);
    if (nt > need) need = nt;
    if (grid > need) grid = need;
    if ((size_t)3 * grid > (size_t)ctx->num_sms * 32) grid = ctx->num_sms * 32 / 3;   // partials capacity (cg_workspace)
    if (grid > 2048) grid = 2048;
    DKMC_CUDA(cudaMemsetAsync(&w.sc->pad, 0, sizeof(int), ctx->stream));
    static const bool trace = getenv("DKMC_SOLVE_TRACE") != nullptr;
    static cudaEvent_t tev0 = nullptr, tev1 = nullptr;
    if (trace && !tev0) { cudaEventCreate(&tev0); cudaEventCreate(&tev1); }
    if (trace) cudaEventRecord(tev0, ctx->stream);
    {
        const PcgKernel kern = (prof ? kPcgKernelsProf : kPcgKernels)[family][variant];
        DKMC_SET_CARVEOUT_FN(kern);
        kern<<<grid, kSpmvThreads, 0, ctx->stream>>>(a);
        ctx->launches++;
        cudaError_t err__ = cudaPeekAtLastError();
        if (err__ != cudaSuccess) {
            set_error("%s:%d: launch of pcg_persistent_kernel failed: %s", __FILE__, __LINE__, cudaGetErrorString(err__));
            return DKMC_ERR_CUDA;
        }
    }
    if (trace) cudaEventRecord(tev1, ctx->stream);
    CgScalars h;
    DKMC_CUDA(cudaMemcpyAsync(&h, w.sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (trace) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tev0, tev1);
        fprintf(stderr, "dkmc pcg launch: rank %d family %d grid %d (variant %d) rows %d tiles %d n_cl %d: %d its, %.3f ms, pad %d\n",
                geo.peers.rank, family, grid, variant, rows, nt, n, h.iters, ms, h.pad);
    }
    if (h.pad == 5 && family == 1 && fallback) {   // table overflow on some rank: every rank left after the first barrier
        *geo.pseq_r = h.rseq_end;
        *geo.pseq_h = h.hseq_end;
        *fallback = 1;
        return DKMC_OK;
    }
    if (h.pad != 0) {
        set_error("persistent PCG: a wait timed out (code %d: 2 local barrier, 3 neighbour halo, 4 reduction) — a CTA could "
                  "not become resident or a peer GPU did not answer", h.pad);
        return DKMC_ERR_CUDA;
    }
    *geo.pseq_r = h.rseq_end;
    *geo.pseq_h = h.hseq_end;
    *iters_out = h.iters;
    *converged = (h.rz <= h.stop) ? 1 : 0;
    if (bb_out) *bb_out = h.bb;
    return DKMC_OK;
}

static int self_window(dkmc_ctx *ctx, int m, SelfWindow **out) {
    SelfWindow *sw = static_cast<SelfWindow *>(ctx->selfwin);
    if (!sw) { sw = new SelfWindow(); ctx->selfwin = sw; }
    if (sw->m_cap < m) {
        if (sw->base) { DKMC_CUDA(cudaStreamSynchronize(ctx->stream)); DKMC_CUDA(cudaFree(sw->base)); sw->base = nullptr; }
        memset(&sw->peers, 0, sizeof(sw->peers));
        const size_t bytes = window_layout(m + m / 8, 1, &sw->peers);
        DKMC_CUDA(cudaMalloc(&sw->base, bytes));
        DKMC_CUDA(cudaMemsetAsync(sw->base, 0, bytes, ctx->stream));
        sw->m_cap = m + m / 8;
        sw->peers.world = 1; sw->peers.rank = 0; sw->peers.base[0] = sw->base;
        // a new window starts with zeroed flags and slots, but the local barriers' generation counters (PcgSync)
        // are monotonic: keep counting both sequences (the LL words of a reduction are matched by equality)
    }
    *out = sw;
    return DKMC_OK;
}

static void free_order(dkmc_ctx *ctx);

void free_solver_state(dkmc_ctx *ctx) {
    free_order(ctx);
    SelfWindow *sw = static_cast<SelfWindow *>(ctx->selfwin);
    if (sw) {
        if (sw->base) cudaFree(sw->base);
        delete sw;
        ctx->selfwin = nullptr;
    }
}

static int run_pcg_persistent_single(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col,
                                     const double *d_val, const double *d_b, double *d_x, const CgWork &w, double tol,
                                     int max_iter, int *iters_out, int *converged, double *bb_out) {
    (void)nnz;
    SelfWindow *sw;
    int rc;
    if ((rc = self_window(ctx, m, &sw))) return rc;
    PcgGeometry geo;
    memset(&geo, 0, sizeof(geo));
    geo.ra = 0; geo.rb = m; geo.t0 = 0; geo.t1 = w.num_tiles; geo.n_cl = w.n_cl;
    geo.row_end_all[0] = m;
    geo.peers = sw->peers;
    geo.pseq_r = &sw->pseq_r; geo.pseq_h = &sw->pseq_h;
    return run_pcg_persistent(ctx, m, d_row_ptr, d_col, d_val, d_b, d_x, w, geo, tol, max_iter, iters_out, converged, bb_out);
}

// dst[0..k) = sum (or max) over ranks of src[0..k)   (src may alias dst)
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(P2pPeers P, unsigned long long seq, int k,
                                                            const double *src, double *dst, int *err, int op_max) {
    const int buf = (int)(seq & 1ull);
    for (int r = 0; r < P.world; ++r) {
        double *slot = reinterpret_cast<double *>(P.base[r] + P.red_off) + ((size_t)buf * P.world + P.rank) * kP2pRedCap;
        for (int j = threadIdx.x; j < k; j += blockDim.x) slot[j] = src[j];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < P.world; r += blockDim.x) st_release_sys(p2p_flag(P, r, buf, P.rank), seq);
    p2p_wait_all(P, buf, seq, err);
    __syncthreads();
    const volatile double *mine = reinterpret_cast<const volatile double *>(P.base[P.rank] + P.red_off) +
                                  (size_t)buf * P.world * kP2pRedCap;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        double acc = mine[j];
        for (int r = 1; r < P.world; ++r) {
            const double v = mine[(size_t)r * kP2pRedCap + j];
            acc = op_max ? fmax(acc, v) : acc + v;
        }
        dst[j] = acc;
    }
}

// window[i] = src[i] over own rows, boundary rows also into the neighbours' windows; the last CTA
// raises the flags and waits for the neighbours' rows (same protocol as dist_direction_p2p_kernel)
__global__ void __launch_bounds__(256) p2p_scatter_rows_kernel(int ra, int rb, const double *__restrict__ src, P2pPeers P,
                                                               P2pHalo H, unsigned long long hseq, unsigned int *counter,
                                                               int *err) {
    __shared__ bool is_last;
    double *mine = reinterpret_cast<double *>(P.base[P.rank]);
    bool pushed = false;
    for (int i = ra + blockIdx.x * blockDim.x + threadIdx.x; i < rb; i += gridDim.x * blockDim.x) {
        const double v = src[i];
        mine[i] = v;
        for (int sgm = 0; sgm < H.n_send; ++sgm)
            if (i >= H.send_begin[sgm] && i < H.send_end[sgm]) {
                reinterpret_cast<double *>(P.base[H.send_peer[sgm]])[i] = v;
                pushed = true;
            }
    }
    if (__syncthreads_or(pushed)) __threadfence_system();
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int buf = 2 + (int)(hseq & 1ull);
    for (int sgm = threadIdx.x; sgm < H.n_send; sgm += blockDim.x) st_release_sys(p2p_flag(P, H.send_peer[sgm], buf, P.rank), hseq);
    for (int sgm = threadIdx.x; sgm < H.n_recv; sgm += blockDim.x) {
        const unsigned long long *f = p2p_flag(P, P.rank, buf, H.recv_peer[sgm]);
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < hseq) {
            if (clock64() - t0 > kP2pTimeoutCycles) { *err = 1; break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0u;
}

// x[i] = window_of_owner(i)[i] for the rows of the other ranks (peer reads over NVLink)
struct P2pRows { int world; int begin[DKMC_MAX_RANKS], end[DKMC_MAX_RANKS]; };
__global__ void __launch_bounds__(256) p2p_pull_rows_kernel(P2pPeers P, P2pRows R, double *__restrict__ x) {
    for (int q = 0; q < R.world; ++q) {
        if (q == P.rank) continue;
        const double *theirs = reinterpret_cast<const double *>(P.base[q]);
        for (int i = R.begin[q] + blockIdx.x * blockDim.x + threadIdx.x; i < R.end[q]; i += gridDim.x * blockDim.x)
            x[i] = __ldcv(theirs + i);
    }
}

// pushes my boundary rows of p (it lives at offset 0 of the window) into the neighbours' p vectors,
// then waits for theirs.  One CTA: the halo of an x-slab is a few hundred KB.
__global__ void __launch_bounds__(1024) p2p_halo_kernel(P2pPeers P, P2pHalo H, unsigned long long seq, int *err) {
    const int buf = 2 + (int)(seq & 1ull);
    const double *mine = reinterpret_cast<const double *>(P.base[P.rank]);
    for (int sgm = 0; sgm < H.n_send; ++sgm) {
        double *theirs = reinterpret_cast<double *>(P.base[H.send_peer[sgm]]);
        for (int i = H.send_begin[sgm] + threadIdx.x; i < H.send_end[sgm]; i += blockDim.x) theirs[i] = mine[i];
    }
    __syncthreads();
    for (int sgm = threadIdx.x; sgm < H.n_send; sgm += blockDim.x) st_release_sys(p2p_flag(P, H.send_peer[sgm], buf, P.rank), seq);
    for (int sgm = threadIdx.x; sgm < H.n_recv; sgm += blockDim.x) {
        const unsigned long long *f = p2p_flag(P, P.rank, buf, H.recv_peer[sgm]);
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > kP2pTimeoutCycles) { *err = 1; break; }
        }
    }
    __syncthreads();
}

#define DKMC_NCCL(call)                                                                       \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            ::dkmc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
            return DKMC_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

// partial cluster sums over the members this rank owns; one thread per cluster (first position)
__global__ void cluster_partial_kernel(int n_cl, int ra, int rb, const int *__restrict__ seg_start,
                                       const int *__restrict__ seg_len, const int *__restrict__ mem_row,
                                       const double *v, double *__restrict__ out) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_cl; s += gridDim.x * blockDim.x) {
        double sum = 0.0;
        if (seg_start[s] == s) {
            const int len = seg_len[s];
            for (int k = 0; k < len; ++k) {
                int row = mem_row[s + k];
                if (row >= ra && row < rb) sum += v[row];
            }
        }
        out[s] = sum;
    }
}

// red[slot] = sum over own rows of v_i^2 * dinv_i
__global__ void __launch_bounds__(kVecThreads) dist_sqnorm_kernel(int ra, int rb, const double *__restrict__ v,
                                                                 const double *__restrict__ dinv, double *partials,
                                                                 unsigned int *counter, double *out) {
    __shared__ double red[32];
    double local = 0.0;
    for (int i = ra + blockIdx.x * blockDim.x + threadIdx.x; i < rb; i += gridDim.x * blockDim.x)
        local += v[i] * v[i] * dinv[i];
    double tot = block_sum(local, red);
    grid_sum_finish(tot, partials, counter, out, red);
}

// coarse completion of a dot: out = base + sum_c w_c * a_c * b_c  (one block)
__device__ __forceinline__ double coarse_dot(int n_cl, const int *seg_start, const double *w, const double *a,
                                             const double *b, double *sh) {
    double local = 0.0;
    for (int s = threadIdx.x; s < n_cl; s += blockDim.x)
        if (seg_start[s] == s) local += w[s] * a[s] * b[s];
    return block_sum(local, sh);
}

// after the init all-reduce: red = [pAp | rr, bb, -, s_c(r)[n], s_c(b)[n]]
__global__ void __launch_bounds__(256) dist_scalars_init_kernel(int n_cl, const int *seg_start, const double *w,
                                                               const double *red, double tol, int max_iter,
                                                               CgScalars *sc) {
    __shared__ double sh[32];
    double c1 = coarse_dot(n_cl, seg_start, w, red + 4, red + 4, sh);
    __syncthreads();
    double c2 = coarse_dot(n_cl, seg_start, w, red + 4 + n_cl, red + 4 + n_cl, sh);
    if (threadIdx.x == 0) {
        double rz = red[1] + c1, bb = red[2] + c2;
        sc->rz = rz; sc->bb = bb; sc->beta = 0.0;
        double ref = bb > 0.0 ? bb : rz;
        sc->stop = tol * tol * ref;
        sc->done = (rz <= sc->stop) ? 1 : 0;
        sc->iters = 0; sc->max_iter = max_iter;
    }
}

// after the per-iteration all-reduce: red = [pAp | rr, -, -, s_c(r)[n]]
__global__ void __launch_bounds__(256) dist_scalars_kernel(int n_cl, const int *seg_start, const double *w,
                                                          const double *red, CgScalars *sc) {
    __shared__ double sh[32];
    if (sc->done) return;
    double c1 = coarse_dot(n_cl, seg_start, w, red + 4, red + 4, sh);
    if (threadIdx.x == 0) {
        double rzn = red[1] + c1;
        sc->beta = rzn / sc->rz;
        sc->rz = rzn;
        sc->iters += 1;
        if (rzn <= sc->stop || sc->iters >= sc->max_iter || !(rzn == rzn)) sc->done = 1;
    }
}

// x += alpha p; r -= alpha Ap over own rows; red[0] = sum r^2 dinv (local part)
__global__ void __launch_bounds__(kVecThreads) dist_update_kernel(int ra, int rb, double *x, double *r,
                                                                 const double *__restrict__ p,
                                                                 const double *__restrict__ Ap,
                                                                 const double *__restrict__ dinv, const double *pAp,
                                                                 double *partials, CgScalars *sc, double *out) {
    __shared__ double red[32];
    if (sc->done) return;
    const double alpha = sc->rz / *pAp;
    double local = 0.0;
    for (int i = ra + blockIdx.x * blockDim.x + threadIdx.x; i < rb; i += gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        local += ri * ri * dinv[i];
    }
    double tot = block_sum(local, red);
    grid_sum_finish(tot, partials, &sc->cnt_b, out, red);
}

// p = D^-1 r + W E^-1 s + beta p over own rows, s = all-reduced cluster sums
__global__ void __launch_bounds__(kVecThreads) dist_direction_kernel(int ra, int rb, const double *__restrict__ r,
                                                                    Precond P, const double *__restrict__ csum,
                                                                    double *__restrict__ p, const CgScalars *sc,
                                                                    int first) {
    if (!first && sc->done) return;
    const double beta = first ? 0.0 : sc->beta;
    for (int i = ra + blockIdx.x * blockDim.x + threadIdx.x; i < rb; i += gridDim.x * blockDim.x) {
        double z = r[i] * P.dinv[i];
        const int s = P.pos ? P.pos[i] : -1;
        if (s >= 0) { const int st = P.seg_start[s]; z += P.w[st] * csum[st]; }
        p[i] = first ? z : z + beta * p[i];
    }
}

// max over own rows of |D^-1 v + W E^-1 s| and |x|
__global__ void __launch_bounds__(kVecThreads) dist_inf_norms_kernel(int ra, int rb, const double *__restrict__ v,
                                                                    Precond P, const double *__restrict__ csum,
                                                                    const double *__restrict__ x,
                                                                    unsigned long long *out) {
    double mz = 0.0, mx = 0.0;
    for (int i = ra + blockIdx.x * blockDim.x + threadIdx.x; i < rb; i += gridDim.x * blockDim.x) {
        double z = v[i] * P.dinv[i];
        const int s = P.pos ? P.pos[i] : -1;
        if (s >= 0) { const int st = P.seg_start[s]; z += P.w[st] * csum[st]; }
        mz = fmax(mz, fabs(z));
        mx = fmax(mx, fabs(x[i]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        mz = fmax(mz, __shfl_xor_sync(0xffffffffu, mz, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, (unsigned long long)__double_as_longlong(mz));
        atomicMax(out + 1, (unsigned long long)__double_as_longlong(mx));
    }
}

struct DistWork {
    CgWork w;
    const dkmc_dist_plan *plan;
    int ra, rb, t0, t1, n_cl;
    double *red;          // [4 + 2 n_cl] all-reduce buffer
    const int *seg_start, *seg_len, *mem_row;
};

static int halo_exchange(dkmc_ctx *ctx, const DistWork &d, double *v) {
    DistState *ds = dist_of(ctx);
    const dkmc_dist_plan *pl = d.plan;
    if (pl->n_send == 0 && pl->n_recv == 0) return DKMC_OK;
    DKMC_NCCL(ncclGroupStart());
    for (int k = 0; k < pl->n_send; ++k)
        DKMC_NCCL(ncclSend(v + pl->send_begin[k], (size_t)(pl->send_end[k] - pl->send_begin[k]), ncclDouble, pl->send_peer[k], ds->comm, ctx->stream));
    for (int k = 0; k < pl->n_recv; ++k)
        DKMC_NCCL(ncclRecv(v + pl->recv_begin[k], (size_t)(pl->recv_end[k] - pl->recv_begin[k]), ncclDouble, pl->recv_peer[k], ds->comm, ctx->stream));
    DKMC_NCCL(ncclGroupEnd());
    return DKMC_OK;
}

// sum of k doubles over the ranks, in place: peer memory when the windows are open, else NCCL
static int dist_allreduce(dkmc_ctx *ctx, double *buf, size_t k, CgScalars *sc, bool op_max = false) {
    DistState *ds = dist_of(ctx);
    if (ds->p2p && k <= (size_t)kP2pRedCap) {
        ++ds->seq;
        DKMC_LAUNCH(ctx, p2p_allreduce_kernel, 1, 256, 0, ds->peers, ds->seq, (int)k, buf, buf, &sc->pad, op_max ? 1 : 0);
        return DKMC_OK;
    }
    DKMC_NCCL(ncclAllReduce(buf, buf, k, ncclDouble, op_max ? ncclMax : ncclSum, ds->comm, ctx->stream));
    return DKMC_OK;
}

// halo of the search direction p: pushed into the neighbours' windows when p lives in the window
static int dist_halo_p(dkmc_ctx *ctx, const DistWork &d, double *p, CgScalars *sc) {
    DistState *ds = dist_of(ctx);
    if (!(ds->p2p && p == reinterpret_cast<double *>(ds->win))) return halo_exchange(ctx, d, p);
    const dkmc_dist_plan *pl = d.plan;
    if (pl->n_send == 0 && pl->n_recv == 0) return DKMC_OK;
    P2pHalo H;
    H.n_send = pl->n_send; H.n_recv = pl->n_recv;
    for (int k = 0; k < pl->n_send; ++k) { H.send_peer[k] = pl->send_peer[k]; H.send_begin[k] = pl->send_begin[k]; H.send_end[k] = pl->send_end[k]; }
    for (int k = 0; k < pl->n_recv; ++k) H.recv_peer[k] = pl->recv_peer[k];
    ++ds->hseq;
    DKMC_LAUNCH(ctx, p2p_halo_kernel, 1, 1024, 0, ds->peers, H, ds->hseq, &sc->pad);
    return DKMC_OK;
}

static void fill_halo_desc(const dkmc_dist_plan *pl, P2pHalo *H) {
    memset(H, 0, sizeof(*H));
    H->n_send = pl->n_send; H->n_recv = pl->n_recv;
    for (int k = 0; k < pl->n_send; ++k) { H->send_peer[k] = pl->send_peer[k]; H->send_begin[k] = pl->send_begin[k]; H->send_end[k] = pl->send_end[k]; }
    for (int k = 0; k < pl->n_recv; ++k) H->recv_peer[k] = pl->recv_peer[k];
}

// Makes v's own rows and halo available to this rank's SpMV.  Peer memory: the rows are copied into
// the window (which the search direction does not occupy at that moment) and the boundary rows
// pushed to the neighbours; *use = the window.  Otherwise NCCL send/recv in place; *use = v.
static int dist_halo_any(dkmc_ctx *ctx, const DistWork &d, double *v, CgScalars *sc, const double **use) {
    DistState *ds = dist_of(ctx);
    const int m_rows = d.plan->row_end[ds->world - 1];
    if (ds->p2p && m_rows <= ds->m_cap && !(g_flags & 128)) {
        P2pHalo H;
        fill_halo_desc(d.plan, &H);
        const int rows = d.rb - d.ra;
        int grid = ceil_div(rows > 0 ? rows : 1, 256);
        if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
        DKMC_LAUNCH(ctx, p2p_scatter_rows_kernel, grid, 256, 0, d.ra, d.rb, v, ds->peers, H, ++ds->hseq, &sc->cnt_d, &sc->pad);
        *use = reinterpret_cast<const double *>(ds->win);
        return DKMC_OK;
    }
    int rc = halo_exchange(ctx, d, v);
    *use = v;
    return rc;
}

// the per-op peer-memory kernels record a wait that timed out in sc->pad: turn it into an error (and clear it, so
// that one slow peer does not fail every later call)
static int dist_check_pad(dkmc_ctx *ctx, CgScalars *sc, const char *what) {
    int pad = 0;
    DKMC_CUDA(cudaMemcpyAsync(&pad, &sc->pad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (pad == 0) return DKMC_OK;
    DKMC_CUDA(cudaMemsetAsync(&sc->pad, 0, sizeof(int), ctx->stream));
    set_error("%s: a peer GPU did not answer within the time limit (peer-memory exchange, code %d): the gathered rows are "
              "incomplete", what, pad);
    return DKMC_ERR_CUDA;
}

static int dist_grid(const dkmc_ctx *ctx, int n) {
    int g = ceil_div(n > 0 ? n : 1, kVecThreads);
    int cap = ctx->num_sms * 8;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

static int dist_pcg(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col, const double *d_val,
                    const double *d_b, double *d_x, const DistWork &d, double tol, int max_iter, int check_every,
                    int *iters_out, int *converged) {
    DistState *ds = dist_of(ctx);
    const CgWork &w = d.w;
    const int nt = d.t1 - d.t0, rows = d.rb - d.ra, vg = dist_grid(ctx, rows), n = d.n_cl;
    // open peer windows: the whole solve is ONE persistent kernel per rank (pcg_persistent.cuh) — halo rows
    // pushed over NVLink from the vector phase, one merged reduction per iteration through peer memory
    if (use_persistent_pcg(ctx) && ds->p2p && m <= ds->m_cap && (size_t)4 + 2 * (size_t)n <= (size_t)kPcgMaxPayload) {
        PcgGeometry geo;
        memset(&geo, 0, sizeof(geo));
        geo.ra = d.ra; geo.rb = d.rb; geo.t0 = d.t0; geo.t1 = d.t1; geo.n_cl = n;
        for (int q = 0; q < ds->world; ++q) geo.row_end_all[q] = d.plan->row_end[q];
        geo.peers = ds->peers;
        fill_halo_desc(d.plan, &geo.halo);
        geo.pseq_r = &ds->pseq_r; geo.pseq_h = &ds->pseq_h;
        return run_pcg_persistent(ctx, m, d_row_ptr, d_col, d_val, d_b, d_x, w, geo, tol, max_iter, iters_out, converged, nullptr);
    }
    // fallback (no peer access, or DKMC_LEGACY_CG=1): one kernel per operation, NCCL (or the per-op peer-memory
    // all-reduce) between them — three exchanges per iteration
    double *r = w.r[0];
    int rc;
    const double *x_use = d_x;
    if ((rc = dist_halo_any(ctx, d, d_x, w.sc, &x_use))) return rc;
    if ((rc = launch_spmv<2>(ctx, nt, m, nnz, d_row_ptr, d_col, d_val, x_use, r, w.tile_row + d.t0, d_b, w.dinv, w.partials,
                             &w.sc->cnt_c, &w.sc->resnorm2, nullptr))) return rc;
    DKMC_LAUNCH(ctx, dist_sqnorm_kernel, vg, kVecThreads, 0, d.ra, d.rb, r, w.dinv, w.partials, &w.sc->cnt_a, d.red + 1);
    DKMC_LAUNCH(ctx, dist_sqnorm_kernel, vg, kVecThreads, 0, d.ra, d.rb, d_b, w.dinv, w.partials, &w.sc->cnt_a, d.red + 2);
    if (n > 0) {
        DKMC_LAUNCH(ctx, cluster_partial_kernel, ceil_div(n, 128), 128, 0, n, d.ra, d.rb, d.seg_start, d.seg_len, d.mem_row, r, d.red + 4);
        DKMC_LAUNCH(ctx, cluster_partial_kernel, ceil_div(n, 128), 128, 0, n, d.ra, d.rb, d.seg_start, d.seg_len, d.mem_row, d_b, d.red + 4 + n);
    }
    if ((rc = dist_allreduce(ctx, d.red + 1, (size_t)3 + 2 * (size_t)n, w.sc))) return rc;
    DKMC_LAUNCH(ctx, dist_scalars_init_kernel, 1, 256, 0, n, d.seg_start, w.P.w, d.red, tol, max_iter, w.sc);
    DKMC_LAUNCH(ctx, dist_direction_kernel, vg, kVecThreads, 0, d.ra, d.rb, r, w.P, d.red + 4, w.p, w.sc, 1);
    CgScalars h;
    memset(&h, 0, sizeof(h));
    int launched = 0;
    if (check_every < 1) check_every = 1;
    while (true) {
        for (int k = 0; k < check_every; ++k) {
            if ((rc = dist_halo_p(ctx, d, w.p, w.sc))) return rc;
            if (nt > 0) {
                if ((rc = launch_spmv<1>(ctx, nt, m, nnz, d_row_ptr, d_col, d_val, w.p, w.Ap, w.tile_row + d.t0, w.p, nullptr,
                                         w.partials, &w.sc->cnt_c, d.red + 0, &w.sc->done))) return rc;
            } else
                DKMC_CUDA(cudaMemsetAsync(d.red, 0, sizeof(double), ctx->stream));
            if ((rc = dist_allreduce(ctx, d.red, 1, w.sc))) return rc;
            DKMC_LAUNCH(ctx, dist_update_kernel, vg, kVecThreads, 0, d.ra, d.rb, d_x, r, w.p, w.Ap, w.dinv, d.red + 0,
                        w.partials, w.sc, d.red + 1);
            if (n > 0)
                DKMC_LAUNCH(ctx, cluster_partial_kernel, ceil_div(n, 128), 128, 0, n, d.ra, d.rb, d.seg_start, d.seg_len, d.mem_row, r, d.red + 4);
            if ((rc = dist_allreduce(ctx, d.red + 1, (size_t)3 + (size_t)n, w.sc))) return rc;
            DKMC_LAUNCH(ctx, dist_scalars_kernel, 1, 256, 0, n, d.seg_start, w.P.w, d.red, w.sc);
            DKMC_LAUNCH(ctx, dist_direction_kernel, vg, kVecThreads, 0, d.ra, d.rb, r, w.P, d.red + 4, w.p, w.sc, 0);
        }
        launched += check_every;
        DKMC_CUDA(cudaMemcpyAsync(&h, w.sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, ctx->stream));
        DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h.pad != 0) {
            set_error("distributed PCG: a peer did not answer within the time limit (peer-memory exchange)");
            return DKMC_ERR_CUDA;
        }
        if (h.done || launched >= max_iter) break;
    }
    *iters_out = h.iters;
    *converged = (h.rz <= h.stop) ? 1 : 0;
    return DKMC_OK;
}

// double-double residual over own rows and the per-entry error estimate (max over ALL ranks)
static int dist_true_residual(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col,
                              const double *d_val, const double *d_rhs, double *d_x, const DistWork &d, double *est) {
    DistState *ds = dist_of(ctx);
    const CgWork &w = d.w;
    const int nt = d.t1 - d.t0, rows = d.rb - d.ra, vg = dist_grid(ctx, rows), n = d.n_cl;
    (void)nnz;
    int rc;
    const double *x_use = d_x;
    if ((rc = dist_halo_any(ctx, d, d_x, w.sc, &x_use))) return rc;
    if (nt > 0)
        DKMC_LAUNCH(ctx, residual_dd_kernel, nt, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, x_use, d_rhs, w.dinv, w.res,
                    w.tile_row + d.t0, w.partials, &w.sc->cnt_d, &w.sc->resnorm2);
    if (n > 0) {
        DKMC_LAUNCH(ctx, cluster_partial_kernel, ceil_div(n, 128), 128, 0, n, d.ra, d.rb, d.seg_start, d.seg_len, d.mem_row, w.res, d.red + 4);
        if ((rc = dist_allreduce(ctx, d.red + 4, (size_t)n, w.sc))) return rc;
    }
    unsigned long long *mx = reinterpret_cast<unsigned long long *>(w.sc + 1);
    DKMC_CUDA(cudaMemsetAsync(mx, 0, 2 * sizeof(unsigned long long), ctx->stream));
    DKMC_LAUNCH(ctx, dist_inf_norms_kernel, vg, kVecThreads, 0, d.ra, d.rb, w.res, w.P, d.red + 4, d_x, mx);
    // non-negative doubles: the max of the values is the max of their bit patterns
    if ((rc = dist_allreduce(ctx, reinterpret_cast<double *>(mx), 2, w.sc, true))) return rc;
    double hm[2] = {0.0, 0.0};
    DKMC_CUDA(cudaMemcpyAsync(hm, mx, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    *est = hm[1] > 0 ? hm[0] / hm[1] : hm[0];
    return DKMC_OK;
}

__global__ void axpy_range_kernel(int ra, int rb, double a, const double *__restrict__ xin, double *__restrict__ y) {
    for (int i = ra + blockIdx.x * blockDim.x + threadIdx.x; i < rb; i += gridDim.x * blockDim.x) y[i] += a * xin[i];
}

static int dist_solve_refined(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col, const double *d_val,
                              const double *d_rhs, double *d_x, DistWork &d, const dkmc_solver_opts &o,
                              dkmc_solve_info *info) {
    int iters = 0, conv = 0, total = 0, rc, rounds = 0;
    double est = 0.0;
    const int rows = d.rb - d.ra, vg = dist_grid(ctx, rows);
    if ((rc = dist_pcg(ctx, m, nnz, d_row_ptr, d_col, d_val, d_rhs, d_x, d, o.rel_tol, o.max_iter, o.check_every, &iters, &conv))) return rc;
    total += iters;
    bool all_conv = conv != 0;
    static const bool trace = getenv("DKMC_SOLVE_TRACE") != nullptr;
    if ((rc = dist_true_residual(ctx, m, nnz, d_row_ptr, d_col, d_val, d_rhs, d_x, d, &est))) return rc;
    if (trace && dist_of(ctx)->rank == 0) fprintf(stderr, "dkmc dist solve: first %d its conv %d -> est %.2e\n", iters, conv, est);
    while (rounds < o.refine_rounds && est > o.est_tol) {
        DKMC_CUDA(cudaMemsetAsync(d.w.e, 0, (size_t)m * sizeof(double), ctx->stream));
        if ((rc = dist_pcg(ctx, m, nnz, d_row_ptr, d_col, d_val, d.w.res, d.w.e, d, restart_tolerance(o, est), o.max_iter, o.check_every, &iters, &conv))) return rc;
        if (trace && dist_of(ctx)->rank == 0) fprintf(stderr, "dkmc dist solve: restart: %d its conv %d\n", iters, conv);
        total += iters;
        if (rows > 0) DKMC_LAUNCH(ctx, axpy_range_kernel, vg, kVecThreads, 0, d.ra, d.rb, 1.0, d.w.e, d_x);
        ++rounds;
        if ((rc = dist_true_residual(ctx, m, nnz, d_row_ptr, d_col, d_val, d_rhs, d_x, d, &est))) return rc;
    }
    if (info) { info->iterations = total; info->refinements = rounds; info->rel_residual = 0.0; info->est_error = est; }
    return (all_conv || est <= o.est_tol) ? DKMC_OK : DKMC_ERR_NOT_CONVERGED;
}

// ================================================================ internal row order of the solver
// The public arrays keep the caller's site order (the reference puts all lattice atoms before all interstitials,
// reorder_boundary.py:113-124: CSR bandwidth ~0.7 N, every SpMV gather a cache miss, and an index range is not a
// slab).  dkmc_solver_set_order registers a permutation of the interior rows (e.g. x-major by grid cell); the
// solver then works on P A P^T: the CSR structure is permuted once, every step the assembled values (computed in
// the caller's order, so the diagonal keeps the reference's summation order bit for bit), the right-hand side,
// 1/diag, the site classes and the starting vector are gathered into the internal order and the solution is
// scattered back.  Nothing outside the solver sees the internal order.
struct SolverOrder {
    const int *key_row_ptr = nullptr;
    int m = 0, nnz = 0;
    int *order = nullptr, *inv = nullptr, *rp2 = nullptr, *col2 = nullptr, *map = nullptr;
};

__global__ void order_invert_kernel(int m, const int *__restrict__ order, int *__restrict__ inv, int *__restrict__ len2,
                                    const int *__restrict__ rp, int *bad) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    const int r = order[p];
    if (r < 0 || r >= m) { atomicAdd(bad, 1); return; }
    if (atomicExch(inv + r, p) != -1) atomicAdd(bad, 1);   // not a permutation
    len2[p] = rp[r + 1] - rp[r];
}

// row p of P A P^T: the entries of row order[p] with renumbered columns, ascending; map = where each came from
__global__ void order_fill_kernel(int m, const int *__restrict__ order, const int *__restrict__ inv, const int *__restrict__ rp,
                                  const int *__restrict__ col, const int *__restrict__ rp2, int *__restrict__ col2,
                                  int *__restrict__ map) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    const int r = order[p], a = rp[r], n = rp[r + 1] - a, o = rp2[p];
    for (int k = 0; k < n; ++k) {   // insertion sort into the output row (rows are a few dozen long)
        const int c = inv[col[a + k]];
        int j = k;
        while (j > 0 && col2[o + j - 1] > c) { col2[o + j] = col2[o + j - 1]; map[o + j] = map[o + j - 1]; --j; }
        col2[o + j] = c; map[o + j] = a + k;
    }
}

__global__ void order_gather_vals_kernel(int nnz, const int *__restrict__ map, const double *__restrict__ val, double *__restrict__ val2) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += gridDim.x * blockDim.x) val2[k] = __ldcs(val + __ldg(map + k));
}

__global__ void order_gather_rows_kernel(int m, int NL, const int *__restrict__ order, const double *__restrict__ rhs,
                                         const double *__restrict__ dinv, const unsigned char *__restrict__ cls,
                                         const double *__restrict__ phi, double *__restrict__ rhs2, double *__restrict__ dinv2,
                                         unsigned char *__restrict__ cls2, double *__restrict__ x2) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    const int r = order[p];
    rhs2[p] = rhs[r]; dinv2[p] = dinv[r]; cls2[p] = cls[NL + r]; x2[p] = phi[NL + r];
}

__global__ void order_scatter_kernel(int m, int NL, const int *__restrict__ order, const double *__restrict__ x2, double *__restrict__ phi) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) phi[NL + order[p]] = x2[p];
}

static SolverOrder *order_of(dkmc_ctx *ctx, const dkmc_sparsity *sp) {
    SolverOrder *so = static_cast<SolverOrder *>(ctx->solver_order);
    return (so && so->key_row_ptr == sp->d_row_ptr && so->m == sp->m && so->nnz == sp->nnz) ? so : nullptr;
}

static void free_order(dkmc_ctx *ctx) {
    SolverOrder *so = static_cast<SolverOrder *>(ctx->solver_order);
    if (!so) return;
    cudaFree(so->order); cudaFree(so->inv); cudaFree(so->rp2); cudaFree(so->col2); cudaFree(so->map);
    delete so;
    ctx->solver_order = nullptr;
}

// what the solver works on: the caller's CSR and arrays, or their internal-order images
struct SolveView {
    int m, nnz, NL;                 // NL: offset of interior row 0 in cls (0 in the internal order)
    const int *rp, *col;
    double *val, *rhs, *x;
    const unsigned char *cls;
    SolverOrder *so;
    double *phi;                    // the caller's potential array (scatter target)
    int phi_NL;
};

// val / rhs hold the matrix assembled in the caller's order (and S_CG_DINV / S_CLASS its 1/diag and classes)
static int view_prepare(dkmc_ctx *ctx, const dkmc_sparsity *sp, int NL, double *val, double *rhs, double *phi, CgWork *w,
                        SolveView *v) {
    v->so = order_of(ctx, sp);
    v->phi = phi; v->phi_NL = NL;
    const unsigned char *cls = static_cast<const unsigned char *>(ctx->slot_ptr[S_CLASS]);
    if (!v->so) {
        v->m = sp->m; v->nnz = sp->nnz; v->NL = NL; v->rp = sp->d_row_ptr; v->col = sp->d_col;
        v->val = val; v->rhs = rhs; v->x = phi + NL; v->cls = cls;
        return DKMC_OK;
    }
    const int m = sp->m, nnz = sp->nnz;
    double *val2, *rhs2, *dinv2, *x2;
    unsigned char *cls2;
    int rc;
    if ((rc = ensure<double>(ctx, S_ORD_VAL, (size_t)nnz, &val2))) return rc;
    if ((rc = ensure<double>(ctx, S_ORD_VEC, (size_t)3 * m, &rhs2))) return rc;
    if ((rc = ensure<unsigned char>(ctx, S_ORD_CLS, (size_t)m, &cls2))) return rc;
    dinv2 = rhs2 + m; x2 = rhs2 + 2 * (size_t)m;
    int g = ceil_div(nnz, 256);
    if (g > ctx->num_sms * 16) g = ctx->num_sms * 16;
    DKMC_LAUNCH(ctx, order_gather_vals_kernel, g, 256, 0, nnz, v->so->map, val, val2);
    DKMC_LAUNCH(ctx, order_gather_rows_kernel, ceil_div(m, 256), 256, 0, m, NL, v->so->order, rhs, w->dinv, cls, phi, rhs2, dinv2,
                cls2, x2);
    w->dinv = dinv2;
    w->P.dinv = dinv2;
    v->m = m; v->nnz = nnz; v->NL = 0; v->rp = v->so->rp2; v->col = v->so->col2;
    v->val = val2; v->rhs = rhs2; v->x = x2; v->cls = cls2;
    return DKMC_OK;
}

static int view_finish(dkmc_ctx *ctx, const SolveView &v) {
    if (v.so) DKMC_LAUNCH(ctx, order_scatter_kernel, ceil_div(v.m, 256), 256, 0, v.m, v.phi_NL, v.so->order, v.x, v.phi);
    return DKMC_OK;
}

}  // namespace dkmc

using namespace dkmc;

extern "C" {

int dkmc_spmv(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col, const double *d_val,
              const double *d_x, double *d_y) {
    DKMC_REQUIRE(ctx && d_row_ptr && d_col && d_val && d_x && d_y, "null pointer");
    DKMC_REQUIRE(m > 0 && nnz >= 0, "m, nnz");
    const int4 *tile_row;
    int num_tiles, rc;
    if ((rc = get_tiling(ctx, m, nnz, d_row_ptr, &tile_row, &num_tiles))) return rc;
    return launch_spmv<0>(ctx, num_tiles, m, nnz, d_row_ptr, d_col, d_val, d_x, d_y, tile_row, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr);
}

static int assemble_impl(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double VL, double VR, int rule,
                         double high_G, double low_G, const int *d_site_element, const int *d_site_charge,
                         const int *d_metals, int num_metals, double *d_val, double *d_rhs);

int dkmc_assemble_K(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double Vd, double high_G,
                    double low_G, const int *d_site_element, const int *d_site_charge, const int *d_metals,
                    int num_metals, double *d_val, double *d_rhs) {
    DKMC_REQUIRE(ctx && sp && d_site_element && d_site_charge && d_val && d_rhs, "null pointer");
    return assemble_impl(ctx, sp, N, NL, NR, -Vd / 2, Vd / 2, 0, high_G, low_G, d_site_element, d_site_charge, d_metals,
                         num_metals, d_val, d_rhs);
}

static int assemble_impl(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double VL, double VR, int rule,
                         double high_G, double low_G, const int *d_site_element, const int *d_site_charge,
                         const int *d_metals, int num_metals, double *d_val, double *d_rhs) {
    DKMC_REQUIRE(sp->m == N - NL - NR, "sparsity does not match N, NL, NR");
    unsigned char *cls;
    double *dinv;
    int rc;
    if ((rc = ensure<unsigned char>(ctx, S_CLASS, N, &cls))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_DINV, sp->m, &dinv))) return rc;
    DKMC_LAUNCH(ctx, site_class_kernel, ceil_div(N, 256), 256, 0, N, d_site_element, d_site_charge, d_metals,
                num_metals, cls);
    DKMC_LAUNCH(ctx, assemble_kernel, ceil_div(sp->m, 128), 128, 0, sp->m, N, NL, NR, VL, VR, rule, high_G, low_G, cls,
                sp->d_row_ptr, sp->d_col, sp->d_left_row_ptr, sp->d_left_col, sp->d_right_row_ptr, sp->d_right_col,
                d_val, d_rhs, dinv);
    return DKMC_OK;
}

int dkmc_solve_cg(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col, const double *d_val,
                  const double *d_rhs, double *d_x, const dkmc_solver_opts *opts, dkmc_solve_info *info) {
    DKMC_REQUIRE(ctx && d_row_ptr && d_col && d_val && d_rhs && d_x, "null pointer");
    dkmc_solver_opts o;
    dkmc_default_solver_opts(&o);
    if (opts) o = *opts;
    CgWork w;
    int rc;
    if ((rc = cg_workspace(ctx, m, nnz, d_row_ptr, &w))) return rc;
    DKMC_LAUNCH(ctx, diag_inverse_kernel, ceil_div(m, 256), 256, 0, m, d_row_ptr, d_col, d_val, w.dinv);
    DKMC_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    rc = solve_refined(ctx, m, nnz, d_row_ptr, d_col, d_val, d_rhs, d_x, w, o, info);
    if (rc != DKMC_OK && rc != DKMC_ERR_NOT_CONVERGED) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_b));
    if (info) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
        info->assemble_ms = 0;
        info->solve_ms = ms;
    }
    if (rc == DKMC_ERR_NOT_CONVERGED) set_error("CG did not converge within max_iter=%d", o.max_iter);
    return rc;
}

int dkmc_background_potential_sparse(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int nn,
                                     const int *d_neigh_idx, int NL, int NR, double Vd, double high_G,
                                     double low_G, const int *d_site_element, const int *d_site_charge,
                                     const int *d_metals, int num_metals, double *d_site_potential_boundary,
                                     const dkmc_solver_opts *opts, dkmc_solve_info *info) {
    (void)nn; (void)d_neigh_idx;
    DKMC_REQUIRE(ctx && sp && d_site_element && d_site_charge && d_site_potential_boundary, "null pointer");
    DKMC_REQUIRE(sp->m == N - NL - NR && sp->m > 0, "sparsity does not match N, NL, NR");
    dkmc_solver_opts o;
    dkmc_default_solver_opts(&o);
    if (opts) o = *opts;
    const int m = sp->m;
    double *val, *rhs;
    CgWork w;
    int rc;
    if ((rc = ensure<double>(ctx, S_CG_VAL, (size_t)sp->nnz, &val))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_RHS, m, &rhs))) return rc;
    if ((rc = cg_workspace(ctx, m, sp->nnz, order_of(ctx, sp) ? order_of(ctx, sp)->rp2 : sp->d_row_ptr, &w))) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    if ((rc = dkmc_assemble_K(ctx, sp, N, NL, NR, Vd, high_G, low_G, d_site_element, d_site_charge, d_metals,
                              num_metals, val, rhs))) return rc;
    SolveView v;
    if ((rc = view_prepare(ctx, sp, NL, val, rhs, d_site_potential_boundary, &w, &v))) return rc;
    if (o.cluster_precond) {
        if ((rc = build_clusters(ctx, m, v.NL, v.cls, v.rp, v.col, v.val, &w))) return rc;
    }
    DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    if (g_flags & 4) {  // keep as much of the matrix as the persisting L2 carve-out holds
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->dev);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
        size_t bytes = (size_t)sp->nnz * sizeof(double);
        if (bytes > (size_t)max_window) bytes = (size_t)max_window;
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = val;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = bytes > 0 ? (float)fmin(1.0, 0.9 * (double)max_persist / (double)bytes) : 0.f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        static bool once = false;
        if (!once) { once = true; fprintf(stderr, "dkmc: L2 persist %d MB, window %zu MB, hitRatio %.2f\n", max_persist >> 20, bytes >> 20, attr.accessPolicyWindow.hitRatio); }
    }
    // warm start: the interior of the previous potential (potential_solver_gpu.cu:754)
    rc = solve_refined(ctx, m, sp->nnz, v.rp, v.col, v.val, v.rhs, v.x, w, o, info);
    if (rc != DKMC_OK && rc != DKMC_ERR_NOT_CONVERGED) return rc;
    { int rc2 = view_finish(ctx, v); if (rc2) return rc2; }
    // Dirichlet contacts (potential_solver.cpp:389-403 / potential_solver_gpu.cu:768-771)
    if (NL > 0) DKMC_LAUNCH(ctx, fill_kernel, ceil_div(NL, 256), 256, 0, NL, -Vd / 2, d_site_potential_boundary);
    if (NR > 0) DKMC_LAUNCH(ctx, fill_kernel, ceil_div(NR, 256), 256, 0, NR, Vd / 2, d_site_potential_boundary + (N - NR));
    DKMC_CUDA(cudaEventRecord(ctx->ev_c, ctx->stream));
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_c));
    if (info) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev_a, ctx->ev_b);
        cudaEventElapsedTime(&b, ctx->ev_b, ctx->ev_c);
        info->assemble_ms = a;
        info->solve_ms = b;
    }
    if (rc == DKMC_ERR_NOT_CONVERGED) set_error("CG did not converge within max_iter=%d", o.max_iter);
    return rc;
}


int dkmc_update_CB_edge_sparse(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double Vd, double q,
                               double high_G, double low_G, const int *d_site_element, const int *d_metals,
                               int num_metals, double *d_site_CB_edge, const dkmc_solver_opts *opts,
                               dkmc_solve_info *info) {
    DKMC_REQUIRE(ctx && sp && d_site_element && d_site_CB_edge, "null pointer");
    DKMC_REQUIRE(sp->m == N - NL - NR && sp->m > 0, "sparsity does not match N, NL, NR");
    dkmc_solver_opts o;
    dkmc_default_solver_opts(&o);
    if (opts) o = *opts;
    const int m = sp->m;
    const double VL = q * Vd / 2, VR = -q * Vd / 2;   // potential_solver.cpp:35,41
    double *val, *rhs;
    CgWork w;
    int rc;
    if ((rc = ensure<double>(ctx, S_CG_VAL, (size_t)sp->nnz, &val))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_RHS, m, &rhs))) return rc;
    if ((rc = cg_workspace(ctx, m, sp->nnz, order_of(ctx, sp) ? order_of(ctx, sp)->rp2 : sp->d_row_ptr, &w))) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    // rule 1: high_G iff either site is a metal; the charges do not enter.  No cluster coarse space:
    // under this rule every strongly coupled set hangs on a Dirichlet contact through the metal layers.
    if ((rc = assemble_impl(ctx, sp, N, NL, NR, VL, VR, 1, high_G, low_G, d_site_element, nullptr, d_metals, num_metals, val,
                            rhs))) return rc;
    SolveView v;
    if ((rc = view_prepare(ctx, sp, NL, val, rhs, d_site_CB_edge, &w, &v))) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    rc = solve_refined(ctx, m, sp->nnz, v.rp, v.col, v.val, v.rhs, v.x, w, o, info);
    if (rc != DKMC_OK && rc != DKMC_ERR_NOT_CONVERGED) return rc;
    { int rc2 = view_finish(ctx, v); if (rc2) return rc2; }
    if (NL > 0) DKMC_LAUNCH(ctx, fill_kernel, ceil_div(NL, 256), 256, 0, NL, VL, d_site_CB_edge);
    if (NR > 0) DKMC_LAUNCH(ctx, fill_kernel, ceil_div(NR, 256), 256, 0, NR, VR, d_site_CB_edge + (N - NR));
    DKMC_CUDA(cudaEventRecord(ctx->ev_c, ctx->stream));
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_c));
    if (info) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev_a, ctx->ev_b);
        cudaEventElapsedTime(&b, ctx->ev_b, ctx->ev_c);
        info->assemble_ms = a;
        info->solve_ms = b;
    }
    if (rc == DKMC_ERR_NOT_CONVERGED) set_error("CG did not converge within max_iter=%d", o.max_iter);
    return rc;
}

int dkmc_solver_set_order(dkmc_ctx *ctx, const dkmc_sparsity *sp, const int *d_order) {
    DKMC_REQUIRE(ctx && sp && sp->d_row_ptr && sp->d_col && sp->m > 0, "sparsity");
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    free_order(ctx);
    if (d_order == nullptr) return DKMC_OK;
    const int m = sp->m, nnz = sp->nnz;
    SolverOrder *so = new SolverOrder();
    ctx->solver_order = so;
    DKMC_CUDA(cudaMalloc(&so->order, (size_t)m * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&so->inv, (size_t)m * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&so->rp2, ((size_t)m + 1) * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&so->col2, (size_t)(nnz > 0 ? nnz : 1) * sizeof(int)));
    DKMC_CUDA(cudaMalloc(&so->map, (size_t)(nnz > 0 ? nnz : 1) * sizeof(int)));
    DKMC_CUDA(cudaMemcpyAsync(so->order, d_order, (size_t)m * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(so->inv, 0xff, (size_t)m * sizeof(int), ctx->stream));
    int *tmp;   // row lengths [m] | bad [1] | scan scratch
    int rc;
    const size_t scan_tmp = (size_t)ceil_div(m, kScanTile) + 1;
    if ((rc = ensure<int>(ctx, S_ORD_TMP, (size_t)m + 1 + scan_tmp, &tmp))) return rc;
    int *len2 = tmp, *bad = tmp + m, *scratch = tmp + m + 1;
    DKMC_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
    DKMC_LAUNCH(ctx, order_invert_kernel, ceil_div(m, 256), 256, 0, m, so->order, so->inv, len2, sp->d_row_ptr, bad);
    int h_bad = 0;
    DKMC_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_bad != 0) {
        free_order(ctx);
        set_error("dkmc_solver_set_order: d_order is not a permutation of 0 .. m-1");
        return DKMC_ERR_ARG;
    }
    DKMC_CUDA(cudaMemsetAsync(so->rp2, 0, sizeof(int), ctx->stream));
    if ((rc = inclusive_scan<int>(ctx, len2, m, so->rp2 + 1, scratch))) return rc;
    DKMC_LAUNCH(ctx, order_fill_kernel, ceil_div(m, 128), 128, 0, m, so->order, so->inv, sp->d_row_ptr, sp->d_col, so->rp2, so->col2,
                so->map);
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    so->key_row_ptr = sp->d_row_ptr; so->m = m; so->nnz = nnz;
    return DKMC_OK;
}

int dkmc_solver_csr(dkmc_ctx *ctx, const dkmc_sparsity *sp, const int **d_row_ptr, const int **d_col, const int **d_order) {
    DKMC_REQUIRE(ctx && sp && d_row_ptr && d_col, "null pointer");
    SolverOrder *so = order_of(ctx, sp);
    *d_row_ptr = so ? so->rp2 : sp->d_row_ptr;
    *d_col = so ? so->col2 : sp->d_col;
    if (d_order) *d_order = so ? so->order : nullptr;
    return DKMC_OK;
}

int dkmc_ctx_invalidate(dkmc_ctx *ctx) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->tiling.d_tile_row) { cudaFree(ctx->tiling.d_tile_row); ctx->tiling.d_tile_row = nullptr; }
    ctx->tiling.row_ptr = nullptr; ctx->tiling.m = 0; ctx->tiling.nnz = 0;
    ctx->pw_grid.d_x = nullptr; ctx->pw_grid.d_sigma = nullptr; ctx->pw_grid.N = 0;
    ctx->pw_inc.valid = false; ctx->pw_inc.since_full = 0;
    return DKMC_OK;
}

int dkmc_spmv_tile_nnz(void) { return kSpmvTile; }

int dkmc_ctx_set_legacy_cg(dkmc_ctx *ctx, int on) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    ctx->legacy_cg = on ? 1 : 0;
    return DKMC_OK;
}

int dkmc_ctx_set_pcg_pipelined(dkmc_ctx *ctx, int on) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    ctx->pcg_pipelined = on ? 1 : 0;
    return DKMC_OK;
}

int dkmc_pcg_profile(dkmc_ctx *ctx, double *out8) {
    DKMC_REQUIRE(ctx && out8, "null pointer");
    for (int q = 0; q < 8; ++q) out8[q] = 0.0;
    long long *prof = static_cast<long long *>(ctx->slot_ptr[S_PCG_PROF]);
    if (!prof) return DKMC_OK;
    std::vector<long long> h(kPcgProfLen);
    DKMC_CUDA(cudaMemcpyAsync(h.data(), prof, kPcgProfLen * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(prof, 0, kPcgProfLen * sizeof(long long), ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < 8; ++q) out8[q] = (double)h[q];
    if (getenv("DKMC_PCG_PROF_CTAS") && h[6] > 0) {   // spread over the CTAs (dev aid)
        std::vector<std::pair<long long, int>> sv;
        long long wmin = 1ll << 62, wmax = 0;
        for (int c = 0; c < 2048; ++c) {
            const long long s_ns = h[8 + 2 * c], w_ns = h[9 + 2 * c] >> 12;
            if (s_ns == 0 && w_ns == 0) continue;
            sv.push_back({s_ns, c});
            if (w_ns < wmin) wmin = w_ns;
            if (w_ns > wmax) wmax = w_ns;
        }
        std::sort(sv.begin(), sv.end());
        const double it = (double)h[6] * 1e3;
        const size_t n = sv.size();
        if (n == 0) return DKMC_OK;   // the pipelined kernel keeps no per-CTA times
        fprintf(stderr, "dkmc pcg prof: %zu CTAs; SpMV phase us/it min %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f; wait at barrier R min %.1f max %.1f; slowest (cta@sm):",
                n, sv[0].first / it, sv[n / 10].first / it, sv[n / 2].first / it, sv[n * 9 / 10].first / it, sv[n * 99 / 100].first / it,
                sv[n - 1].first / it, wmin / it, wmax / it);
        for (size_t q = 0; q < 10 && q < n; ++q) {
            const int c = sv[n - 1 - q].second;
            fprintf(stderr, " %d@%d(%.1f)", c, (int)(h[9 + 2 * c] & 0xfff), sv[n - 1 - q].first / it);
        }
        fprintf(stderr, "\n");
    }
    return DKMC_OK;
}

int dkmc_dist_unique_id(char *id128) {
    DKMC_REQUIRE(id128 != nullptr, "id buffer");
    static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId larger than 128 bytes");
    ncclUniqueId id;
    DKMC_NCCL(ncclGetUniqueId(&id));
    memset(id128, 0, 128);
    memcpy(id128, &id, sizeof(id));
    return DKMC_OK;
}

int dkmc_dist_init(dkmc_ctx *ctx, int rank, int world, const char *id128) {
    DKMC_REQUIRE(ctx && id128 && world >= 1 && rank >= 0 && rank < world, "arguments");
    DKMC_REQUIRE(ctx->dist == nullptr, "already initialised");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    DistState *ds = new DistState();
    ds->rank = rank; ds->world = world;
    DKMC_NCCL(ncclCommInitRank(&ds->comm, world, id, rank));
    ctx->dist = ds;
    return DKMC_OK;
}

int dkmc_dist_p2p_alloc(dkmc_ctx *ctx, int m, char *ipc_handle64) {
    DKMC_REQUIRE(ctx && ipc_handle64 && m > 0, "arguments");
    DistState *ds = dist_of(ctx);
    DKMC_REQUIRE(ds != nullptr, "dkmc_dist_init must be called first");
    DKMC_REQUIRE(ds->win == nullptr, "window already allocated");
    static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "cudaIpcMemHandle_t larger than 64 bytes");
    ds->win_bytes = window_layout(m, ds->world, &ds->peers);
    DKMC_CUDA(cudaMalloc(&ds->win, ds->win_bytes));
    DKMC_CUDA(cudaMemset(ds->win, 0, ds->win_bytes));
    ds->m_cap = m;
    cudaIpcMemHandle_t h;
    DKMC_CUDA(cudaIpcGetMemHandle(&h, ds->win));
    memset(ipc_handle64, 0, 64);
    memcpy(ipc_handle64, &h, sizeof(h));
    return DKMC_OK;
}

int dkmc_dist_p2p_open(dkmc_ctx *ctx, const char *ipc_handles64) {
    DKMC_REQUIRE(ctx && ipc_handles64, "arguments");
    DistState *ds = dist_of(ctx);
    DKMC_REQUIRE(ds != nullptr && ds->win != nullptr, "dkmc_dist_p2p_alloc must be called first");
    ds->peers.world = ds->world;
    ds->peers.rank = ds->rank;
    for (int r = 0; r < ds->world; ++r) {
        if (r == ds->rank) { ds->peers.base[r] = ds->win; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handles64 + (size_t)r * 64, sizeof(h));
        void *ptr = nullptr;
        DKMC_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ds->peers.base[r] = static_cast<unsigned char *>(ptr);
    }
    ds->seq = 0;
    ds->hseq = 0;
    ds->p2p = true;
    return DKMC_OK;
}

int dkmc_dist_allgather_rows(dkmc_ctx *ctx, double *d_buf, int n, const int *row_begin, const int *row_end) {
    DKMC_REQUIRE(ctx && d_buf && row_begin && row_end && n > 0, "arguments");
    DistState *ds = dist_of(ctx);
    DKMC_REQUIRE(ds != nullptr, "dkmc_dist_init must be called first");
    const int me = ds->rank;
    if (ds->p2p && n <= ds->m_cap) {
        CgScalars *sc;
        bool fresh = ctx->slot_ptr[S_SCALARS] == nullptr;
        int rc;
        if ((rc = ensure<CgScalars>(ctx, S_SCALARS, 4, &sc))) return rc;
        if (fresh) DKMC_CUDA(cudaMemsetAsync(sc, 0, 4 * sizeof(CgScalars), ctx->stream));
        double *scratch;
        if ((rc = ensure<double>(ctx, S_DIST_RED, 16, &scratch))) return rc;
        // barrier first: nobody may still be reading the windows (e.g. the tail of the previous solve)
        if ((rc = dist_allreduce(ctx, scratch, 1, sc))) return rc;
        P2pHalo H0;
        memset(&H0, 0, sizeof(H0));
        const int rows = row_end[me] - row_begin[me];
        int grid = ceil_div(rows > 0 ? rows : 1, 256);
        if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
        DKMC_LAUNCH(ctx, p2p_scatter_rows_kernel, grid, 256, 0, row_begin[me], row_end[me], d_buf, ds->peers, H0, ++ds->hseq,
                    &sc->cnt_d, &sc->pad);
        if ((rc = dist_allreduce(ctx, scratch, 1, sc))) return rc;
        P2pRows R;
        memset(&R, 0, sizeof(R));
        R.world = ds->world;
        for (int q = 0; q < ds->world; ++q) { R.begin[q] = row_begin[q]; R.end[q] = row_end[q]; }
        DKMC_LAUNCH(ctx, p2p_pull_rows_kernel, ctx->num_sms * 4, 256, 0, ds->peers, R, d_buf);
        if ((rc = dist_allreduce(ctx, scratch, 1, sc))) return rc;
        return dist_check_pad(ctx, sc, "dkmc_dist_allgather_rows");
    }
    DKMC_NCCL(ncclGroupStart());
    for (int r = 0; r < ds->world; ++r) {
        int cnt = row_end[r] - row_begin[r];
        if (cnt > 0) DKMC_NCCL(ncclBroadcast(d_buf + row_begin[r], d_buf + row_begin[r], (size_t)cnt, ncclDouble, r, ds->comm, ctx->stream));
    }
    DKMC_NCCL(ncclGroupEnd());
    return DKMC_OK;
}

int dkmc_dist_finalize(dkmc_ctx *ctx) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    DistState *ds = dist_of(ctx);
    if (ds) {
        cudaStreamSynchronize(ctx->stream);
        if (ds->p2p)
            for (int r = 0; r < ds->world; ++r)
                if (r != ds->rank && ds->peers.base[r]) cudaIpcCloseMemHandle(ds->peers.base[r]);
        if (ds->win) cudaFree(ds->win);
        // the distributed PCG's barrier generations restart with the next DistState's sequences
        if (ctx->slot_ptr[S_PCG_SYNC2]) cudaMemset(ctx->slot_ptr[S_PCG_SYNC2], 0, sizeof(PcgSync));
        if (ds->comm) ncclCommDestroy(ds->comm);
        delete ds;
        ctx->dist = nullptr;
    }
    return DKMC_OK;
}

int dkmc_dist_background_potential(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double Vd,
                                   double high_G, double low_G, const int *d_site_element, const int *d_site_charge,
                                   const int *d_metals, int num_metals, double *d_site_potential_boundary,
                                   const dkmc_dist_plan *plan, const dkmc_solver_opts *opts, dkmc_solve_info *info) {
    DKMC_REQUIRE(ctx && sp && plan && d_site_element && d_site_charge && d_site_potential_boundary, "null pointer");
    DistState *ds = dist_of(ctx);
    DKMC_REQUIRE(ds != nullptr, "dkmc_dist_init must be called first");
    DKMC_REQUIRE(sp->m == N - NL - NR && sp->m > 0, "sparsity does not match N, NL, NR");
    DKMC_REQUIRE(plan->world == ds->world && plan->n_send <= DKMC_MAX_HALO_SEGMENTS && plan->n_recv <= DKMC_MAX_HALO_SEGMENTS, "plan");
    dkmc_solver_opts o;
    dkmc_default_solver_opts(&o);
    if (opts) o = *opts;
    const int m = sp->m;
    double *val, *rhs;
    DistWork d;
    int rc;
    if ((rc = ensure<double>(ctx, S_CG_VAL, (size_t)sp->nnz, &val))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_RHS, m, &rhs))) return rc;
    if ((rc = cg_workspace(ctx, m, sp->nnz, order_of(ctx, sp) ? order_of(ctx, sp)->rp2 : sp->d_row_ptr, &d.w))) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    // replicated: assembly of all rows and the cluster detection (sub-millisecond)
    if ((rc = dkmc_assemble_K(ctx, sp, N, NL, NR, Vd, high_G, low_G, d_site_element, d_site_charge, d_metals, num_metals, val, rhs))) return rc;
    d.n_cl = 0;
    d.seg_start = d.seg_len = d.mem_row = nullptr;
    SolveView v;
    if ((rc = view_prepare(ctx, sp, NL, val, rhs, d_site_potential_boundary, &d.w, &v))) return rc;
    if (o.cluster_precond) {
        if ((rc = build_clusters(ctx, m, v.NL, v.cls, v.rp, v.col, v.val, &d.w))) return rc;
        d.n_cl = d.w.n_cl;
        d.seg_start = d.w.P.seg_start; d.seg_len = d.w.P.seg_len; d.mem_row = d.w.P.mem_row;
    }
    DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    d.plan = plan;
    d.ra = plan->row_begin[ds->rank]; d.rb = plan->row_end[ds->rank];
    // own rows are whole SpMV tiles: locate them in the tiling
    {
        const int T = kSpmvTile;
        int h[2];
        DKMC_CUDA(cudaMemcpyAsync(&h[0], v.rp + d.ra, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        DKMC_CUDA(cudaMemcpyAsync(&h[1], v.rp + d.rb, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
        // tile t holds the rows whose first non-zero index lies in [t*T, (t+1)*T)
        d.t0 = h[0] / T;
        d.t1 = d.rb >= m ? d.w.num_tiles : h[1] / T;
        if (d.rb <= d.ra) d.t1 = d.t0;
    }
    if ((rc = ensure<double>(ctx, S_DIST_RED, (size_t)4 + 2 * (size_t)d.n_cl + 8, &d.red))) return rc;
    // fallback path with open peer windows: the search direction lives in this rank's window (offset 0), where
    // the neighbours' p2p_halo_kernel pushes their boundary rows (the persistent PCG keeps u there instead)
    if (ds->p2p && m <= ds->m_cap && !use_persistent_pcg(ctx)) d.w.p = reinterpret_cast<double *>(ds->win);
    double *x = v.x;
    rc = dist_solve_refined(ctx, m, sp->nnz, v.rp, v.col, v.val, v.rhs, x, d, o, info);
    if (rc != DKMC_OK && rc != DKMC_ERR_NOT_CONVERGED) return rc;
    // all-gather of the solution.  Peer memory: own rows into the window, a barrier, every rank pulls
    // the other ranks' rows from their windows, a barrier (the windows are written again in the next solve)
    if (ds->p2p && m <= ds->m_cap && !(g_flags & 128)) {
        P2pHalo H0;
        memset(&H0, 0, sizeof(H0));
        const int rows = d.rb - d.ra;
        int grid = ceil_div(rows > 0 ? rows : 1, 256);
        if (grid > ctx->num_sms * 4) grid = ctx->num_sms * 4;
        DKMC_LAUNCH(ctx, p2p_scatter_rows_kernel, grid, 256, 0, d.ra, d.rb, x, ds->peers, H0, ++ds->hseq, &d.w.sc->cnt_d, &d.w.sc->pad);
        if ((rc = dist_allreduce(ctx, d.red, 1, d.w.sc))) return rc;   // barrier
        P2pRows R;
        memset(&R, 0, sizeof(R));
        R.world = ds->world;
        for (int q = 0; q < ds->world; ++q) { R.begin[q] = plan->row_begin[q]; R.end[q] = plan->row_end[q]; }
        DKMC_LAUNCH(ctx, p2p_pull_rows_kernel, ctx->num_sms * 4, 256, 0, ds->peers, R, x);
        if ((rc = dist_allreduce(ctx, d.red, 1, d.w.sc))) return rc;   // barrier
        { int rcp = dist_check_pad(ctx, d.w.sc, "dkmc_dist_background_potential (all-gather of the solution)"); if (rcp) return rcp; }
    } else {
    DKMC_NCCL(ncclGroupStart());
    for (int r = 0; r < ds->world; ++r) {
        int cnt = plan->row_end[r] - plan->row_begin[r];
        if (cnt > 0) DKMC_NCCL(ncclBroadcast(x + plan->row_begin[r], x + plan->row_begin[r], (size_t)cnt, ncclDouble, r, ds->comm, ctx->stream));
    }
    DKMC_NCCL(ncclGroupEnd());
    }
    { int rc2 = view_finish(ctx, v); if (rc2) return rc2; }
    if (NL > 0) DKMC_LAUNCH(ctx, fill_kernel, ceil_div(NL, 256), 256, 0, NL, -Vd / 2, d_site_potential_boundary);
    if (NR > 0) DKMC_LAUNCH(ctx, fill_kernel, ceil_div(NR, 256), 256, 0, NR, Vd / 2, d_site_potential_boundary + (N - NR));
    DKMC_CUDA(cudaEventRecord(ctx->ev_c, ctx->stream));
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_c));
    if (info) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev_a, ctx->ev_b);
        cudaEventElapsedTime(&b, ctx->ev_b, ctx->ev_c);
        info->assemble_ms = a;
        info->solve_ms = b;
    }
    if (rc == DKMC_ERR_NOT_CONVERGED) set_error("distributed CG did not converge within max_iter=%d", o.max_iter);
    return rc;
}

}  // extern "C"
