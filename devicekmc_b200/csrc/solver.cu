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
#include "common.cuh"

namespace dkmc {

constexpr int kSpmvThreads = 256;
constexpr int kSpmvTile = 2048;             // nnz per tile (by row start)
constexpr int kSpmvCap = kSpmvTile + 256;   // shared products per block
constexpr int kVecThreads = 256;
constexpr int kMaxPartials = 1 << 16;

struct CgScalars {
    double rz, rz_new, pAp, alpha, beta, bb, stop, resnorm2, bnorm2;
    int done, iters, max_iter, pad;
    unsigned int cnt_a, cnt_b, cnt_c, cnt_d;
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
    cls[i] = metal ? 1 : ((e == DKMC_VACANCY && charge[i] == 0) ? 2 : 0);
}

// conductance rule potential_solver.cpp:325-346: high_G iff (metal & metal) or (uncharged V & uncharged V)
__device__ __forceinline__ double conductance(unsigned char ci, unsigned char cj, double high_G, double low_G) {
    return (ci != 0 && ci == cj) ? high_G : low_G;
}

__global__ void __launch_bounds__(128) assemble_kernel(
    int m, int N, int NL, int NR, double Vd, double high_G, double low_G,
    const unsigned char *__restrict__ cls, const int *__restrict__ row_ptr, const int *__restrict__ col,
    const int *__restrict__ lrp, const int *__restrict__ lcol, const int *__restrict__ rrp,
    const int *__restrict__ rcol, double *__restrict__ val, double *__restrict__ rhs,
    double *__restrict__ dinv) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const double VL = -Vd / 2, VR = Vd / 2;
    const unsigned char ci = cls[r + NL];
    double diag = 0.0, ksub = 0.0;
    // ascending j: left contact, interior, right contact (potential_solver.cpp:350-372)
    for (int p = lrp[r]; p < lrp[r + 1]; ++p) {
        double G = conductance(ci, cls[lcol[p]], high_G, low_G);
        diag = __dadd_rn(diag, G);
        ksub = __dadd_rn(ksub, __dmul_rn(-G, VL));
    }
    int diag_pos = -1;
    for (int p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
        int c = col[p];
        if (c == r) { diag_pos = p; continue; }
        double G = conductance(ci, cls[c + NL], high_G, low_G);
        val[p] = -G;
        diag = __dadd_rn(diag, G);
    }
    for (int p = rrp[r]; p < rrp[r + 1]; ++p) {
        double G = conductance(ci, cls[rcol[p] + (N - NR)], high_G, low_G);
        diag = __dadd_rn(diag, G);
        ksub = __dadd_rn(ksub, __dmul_rn(-G, VR));
    }
    if (diag_pos >= 0) val[diag_pos] = diag;
    rhs[r] = -ksub;  // D*phi = -Ksub (potential_solver.cpp:379,396)
    if (dinv) dinv[r] = 1.0 / diag;
}

// ---------------------------------------------------------------- SpMV tiling
__global__ void tile_rows_kernel(int m, int num_tiles, const int *__restrict__ row_ptr, int *__restrict__ tile_row) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > num_tiles) return;
    if (t == num_tiles) { tile_row[t] = m; return; }
    int target = t * kSpmvTile;
    int lo = 0, hi = m;  // first r in [0,m) with row_ptr[r] >= target
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (row_ptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    tile_row[t] = lo;
}

// y = A x over one nnz tile per block.  MODE 0: y only.  MODE 1: also dot(w, y) -> *dot_out.
// MODE 2: y = b - A x (residual) and rr = sum (y^2 * dinv) -> *dot_out.
template <int MODE>
__global__ void __launch_bounds__(kSpmvThreads) spmv_tile_kernel(
    int m, const int *__restrict__ row_ptr, const int *__restrict__ col, const double *__restrict__ val,
    const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ tile_row,
    const double *__restrict__ w, const double *__restrict__ dinv, double *partials, unsigned int *counter,
    double *dot_out, const int *done_flag) {
    __shared__ double prod[kSpmvCap];
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    const int r0 = tile_row[blockIdx.x], r1 = tile_row[blockIdx.x + 1];
    double local = 0.0;
    if (r0 < r1) {
        const int k0 = row_ptr[r0], k1 = row_ptr[r1];
        const int cnt = k1 - k0;
        if (cnt <= kSpmvCap) {
            // phase 1: coalesced stream of val/col, gather x through the read-only path
            for (int base = 0; base < cnt; base += kSpmvThreads * 4) {
                double v[4];
                int c[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int k = base + u * kSpmvThreads + threadIdx.x;
                    bool ok = k < cnt;
                    v[u] = ok ? __ldcs(val + k0 + k) : 0.0;
                    c[u] = ok ? __ldcs(col + k0 + k) : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int k = base + u * kSpmvThreads + threadIdx.x;
                    if (k < cnt) prod[k] = v[u] * __ldg(x + c[u]);
                }
            }
            __syncthreads();
            // phase 2: each row adds its products in CSR order
            for (int r = r0 + threadIdx.x; r < r1; r += kSpmvThreads) {
                int a = row_ptr[r] - k0, b = row_ptr[r + 1] - k0;
                double s = 0.0;
                for (int k = a; k < b; ++k) s += prod[k];
                if (MODE == 2) { s = w[r] - s; local += s * s * dinv[r]; }
                y[r] = s;
                if (MODE == 1) local += w[r] * s;
            }
        } else {  // rows longer than the staging buffer: direct path
            for (int r = r0 + threadIdx.x; r < r1; r += kSpmvThreads) {
                double s = 0.0;
                for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) s += val[k] * __ldg(x + col[k]);
                if (MODE == 2) { s = w[r] - s; local += s * s * dinv[r]; }
                y[r] = s;
                if (MODE == 1) local += w[r] * s;
            }
        }
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
    double *__restrict__ res, const int *__restrict__ tile_row, double *partials, unsigned int *counter,
    double *res_out) {
    __shared__ double hi[kSpmvCap];
    __shared__ double lo[kSpmvCap];
    __shared__ double red[32];
    const int r0 = tile_row[blockIdx.x], r1 = tile_row[blockIdx.x + 1];
    double local = 0.0;
    if (r0 < r1) {
        const int k0 = row_ptr[r0], k1 = row_ptr[r1];
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

// ---------------------------------------------------------------- CG vector kernels
// z = r * dinv; p = z; rz = r.z; bb = b.(dinv b); sets the stop threshold and clears flags
__global__ void __launch_bounds__(kVecThreads) cg_init_kernel(int m, const double *__restrict__ r,
                                                             const double *__restrict__ b,
                                                             const double *__restrict__ dinv,
                                                             double *__restrict__ p, double tol,
                                                             int max_iter, double *partials, CgScalars *sc) {
    __shared__ double red[32];
    __shared__ double red2[32];
    double lrz = 0.0, lbb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        double ri = r[i], di = dinv[i], bi = b[i];
        double z = ri * di;
        p[i] = z;
        lrz += ri * z;
        lbb += bi * bi * di;
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

// alpha = rz/pAp; x += alpha p; r -= alpha Ap; rz_new = r.(dinv r); last block: beta, convergence
__global__ void __launch_bounds__(kVecThreads) cg_update_kernel(int m, double *__restrict__ x,
                                                               double *__restrict__ r,
                                                               const double *__restrict__ p,
                                                               const double *__restrict__ Ap,
                                                               const double *__restrict__ dinv,
                                                               double *partials, CgScalars *sc) {
    __shared__ double red[32];
    if (sc->done) return;
    const double alpha = sc->rz / sc->pAp;
    double local = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        double pi = p[i];
        x[i] += alpha * pi;
        double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        local += ri * ri * dinv[i];
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

// p = dinv r + beta p   (skipped once converged so that x,r stay final)
__global__ void __launch_bounds__(kVecThreads) cg_direction_kernel(int m, const double *__restrict__ r,
                                                                  const double *__restrict__ dinv,
                                                                  double *__restrict__ p, const CgScalars *sc) {
    if (sc->done) return;
    const double beta = sc->beta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
        p[i] = r[i] * dinv[i] + beta * p[i];
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
static int get_tiling(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int **tile_row, int *num_tiles) {
    SpmvTiling &t = ctx->tiling;
    if (t.row_ptr != d_row_ptr || t.m != m || t.nnz != nnz) {
        if (t.d_tile_row) {
            DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
            DKMC_CUDA(cudaFree(t.d_tile_row));
            t.d_tile_row = nullptr;
        }
        t.num_tiles = ceil_div(nnz > 0 ? nnz : 1, kSpmvTile);
        DKMC_CUDA(cudaMalloc(&t.d_tile_row, ((size_t)t.num_tiles + 1) * sizeof(int)));
        DKMC_LAUNCH(ctx, tile_rows_kernel, ceil_div(t.num_tiles + 1, 256), 256, 0, m, t.num_tiles, d_row_ptr, t.d_tile_row);
        t.row_ptr = d_row_ptr; t.m = m; t.nnz = nnz;
    }
    *tile_row = t.d_tile_row;
    *num_tiles = t.num_tiles;
    return DKMC_OK;
}

static int vec_grid(const dkmc_ctx *ctx, int m) {
    int g = ceil_div(m, kVecThreads);
    int cap = ctx->num_sms * 8;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

struct CgWork {
    double *r, *p, *Ap, *dinv, *res, *e, *partials;
    CgScalars *sc;
    const int *tile_row;
    int num_tiles;
};

static int cg_workspace(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, CgWork *w) {
    int rc;
    if ((rc = ensure<double>(ctx, S_CG_R, m, &w->r))) return rc;
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
    return DKMC_OK;
}

// Preconditioned CG on A x = b starting from x (in/out).  dinv must be set.  Returns iterations.
static int run_pcg(dkmc_ctx *ctx, int m, const int *d_row_ptr, const int *d_col, const double *d_val,
                   const double *d_b, double *d_x, const CgWork &w, double tol, int max_iter, int check_every,
                   int *iters_out, int *converged, double *bb_out) {
    const int vg = vec_grid(ctx, m);
    // r = b - A x and its weighted norm (unused here), then z/p/rz/bb
    DKMC_LAUNCH(ctx, spmv_tile_kernel<2>, w.num_tiles, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, d_x, w.r,
                w.tile_row, d_b, w.dinv, w.partials, &w.sc->cnt_c, &w.sc->resnorm2, (const int *)nullptr);
    DKMC_LAUNCH(ctx, cg_init_kernel, vg, kVecThreads, 0, m, w.r, d_b, w.dinv, w.p, tol, max_iter, w.partials, w.sc);
    CgScalars h;
    memset(&h, 0, sizeof(h));
    int launched = 0;
    if (check_every < 1) check_every = 1;
    while (true) {
        for (int k = 0; k < check_every; ++k) {
            DKMC_LAUNCH(ctx, spmv_tile_kernel<1>, w.num_tiles, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, w.p,
                        w.Ap, w.tile_row, w.p, (const double *)nullptr, w.partials, &w.sc->cnt_c, &w.sc->pAp,
                        &w.sc->done);
            DKMC_LAUNCH(ctx, cg_update_kernel, vg, kVecThreads, 0, m, d_x, w.r, w.p, w.Ap, w.dinv, w.partials, w.sc);
            DKMC_LAUNCH(ctx, cg_direction_kernel, vg, kVecThreads, 0, m, w.r, w.dinv, w.p, w.sc);
        }
        launched += check_every;
        DKMC_CUDA(cudaMemcpyAsync(&h, w.sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, ctx->stream));
        DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h.done || launched >= max_iter) break;
    }
    *iters_out = h.iters;
    *converged = (h.rz <= h.stop) ? 1 : 0;
    if (bb_out) *bb_out = h.bb;
    return DKMC_OK;
}

static int solve_refined(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col,
                         const double *d_val, const double *d_rhs, double *d_x, CgWork &w,
                         const dkmc_solver_opts &o, dkmc_solve_info *info) {
    int iters = 0, conv = 0, total = 0, rc;
    double bb0 = 0.0;
    bool all_conv = true;
    if ((rc = run_pcg(ctx, m, d_row_ptr, d_col, d_val, d_rhs, d_x, w, o.rel_tol, o.max_iter, o.check_every, &iters, &conv, &bb0))) return rc;
    total += iters;
    all_conv &= (conv != 0);
    const int vg = vec_grid(ctx, m);
    int rounds = 0;
    for (int k = 0; k < o.refine_rounds; ++k) {
        DKMC_LAUNCH(ctx, residual_dd_kernel, w.num_tiles, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, d_x, d_rhs,
                    w.dinv, w.res, w.tile_row, w.partials, &w.sc->cnt_d, &w.sc->resnorm2);
        DKMC_LAUNCH(ctx, fill_kernel, vg, kVecThreads, 0, m, 0.0, w.e);
        double ref_tol = o.rel_tol < 1e-7 ? 1e-7 : o.rel_tol;
        if ((rc = run_pcg(ctx, m, d_row_ptr, d_col, d_val, w.res, w.e, w, ref_tol, o.max_iter, o.check_every, &iters, &conv, nullptr))) return rc;
        total += iters;
        DKMC_LAUNCH(ctx, axpy_kernel, vg, kVecThreads, 0, m, 1.0, w.e, d_x);
        ++rounds;
    }
    // final accurate residual for the report
    DKMC_LAUNCH(ctx, residual_dd_kernel, w.num_tiles, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, d_x, d_rhs,
                w.dinv, w.res, w.tile_row, w.partials, &w.sc->cnt_d, &w.sc->resnorm2);
    double h = 0.0;
    DKMC_CUDA(cudaMemcpyAsync(&h, &w.sc->resnorm2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (info) {
        info->iterations = total;
        info->refinements = rounds;
        info->rel_residual = bb0 > 0 ? sqrt(h / bb0) : sqrt(h);
    }
    return all_conv ? DKMC_OK : DKMC_ERR_NOT_CONVERGED;
}

}  // namespace dkmc

using namespace dkmc;

extern "C" {

int dkmc_spmv(dkmc_ctx *ctx, int m, int nnz, const int *d_row_ptr, const int *d_col, const double *d_val,
              const double *d_x, double *d_y) {
    DKMC_REQUIRE(ctx && d_row_ptr && d_col && d_val && d_x && d_y, "null pointer");
    DKMC_REQUIRE(m > 0 && nnz >= 0, "m, nnz");
    const int *tile_row;
    int num_tiles, rc;
    if ((rc = get_tiling(ctx, m, nnz, d_row_ptr, &tile_row, &num_tiles))) return rc;
    DKMC_LAUNCH(ctx, spmv_tile_kernel<0>, num_tiles, kSpmvThreads, 0, m, d_row_ptr, d_col, d_val, d_x, d_y, tile_row,
                (const double *)nullptr, (const double *)nullptr, (double *)nullptr, (unsigned int *)nullptr,
                (double *)nullptr, (const int *)nullptr);
    return DKMC_OK;
}

int dkmc_assemble_K(dkmc_ctx *ctx, const dkmc_sparsity *sp, int N, int NL, int NR, double Vd, double high_G,
                    double low_G, const int *d_site_element, const int *d_site_charge, const int *d_metals,
                    int num_metals, double *d_val, double *d_rhs) {
    DKMC_REQUIRE(ctx && sp && d_site_element && d_site_charge && d_val && d_rhs, "null pointer");
    DKMC_REQUIRE(sp->m == N - NL - NR, "sparsity does not match N, NL, NR");
    unsigned char *cls;
    double *dinv;
    int rc;
    if ((rc = ensure<unsigned char>(ctx, S_CLASS, N, &cls))) return rc;
    if ((rc = ensure<double>(ctx, S_CG_DINV, sp->m, &dinv))) return rc;
    DKMC_LAUNCH(ctx, site_class_kernel, ceil_div(N, 256), 256, 0, N, d_site_element, d_site_charge, d_metals,
                num_metals, cls);
    DKMC_LAUNCH(ctx, assemble_kernel, ceil_div(sp->m, 128), 128, 0, sp->m, N, NL, NR, Vd, high_G, low_G, cls,
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
    if ((rc = cg_workspace(ctx, m, sp->nnz, sp->d_row_ptr, &w))) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    if ((rc = dkmc_assemble_K(ctx, sp, N, NL, NR, Vd, high_G, low_G, d_site_element, d_site_charge, d_metals,
                              num_metals, val, rhs))) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    // warm start: the interior of the previous potential (potential_solver_gpu.cu:754)
    double *x = d_site_potential_boundary + NL;
    rc = solve_refined(ctx, m, sp->nnz, sp->d_row_ptr, sp->d_col, val, rhs, x, w, o, info);
    if (rc != DKMC_OK && rc != DKMC_ERR_NOT_CONVERGED) return rc;
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

}  // extern "C"
