// devicekmc-b200 — CSR SpMV, persistent CTAs fed by TMA bulk copies (sm_100a).
//
// The first SpMV version (spmv_tile_kernel) loads val/col through registers: every block pays a
// DRAM round trip per batch of loads, then the gather latency, then a barrier — ncu shows 62 %
// long-scoreboard + 25 % barrier stalls and ~3 TB/s in a CG iteration.  Here a CTA owns a strided
// set of nnz tiles and keeps a ring of kStages shared-memory stages; one elected thread issues
// `cp.async.bulk` (1-D TMA) copies of the next tiles' val and col arrays, completion is signalled on
// an mbarrier per stage (expect_tx / complete_tx), so DRAM stays busy while the other threads gather
// x, multiply in place and reduce the rows of the current tile.  No register staging, no per-tile
// launch-side dependency chain.
//
// Same arithmetic as spmv_tile_kernel: products in CSR order, each row summed left to right, so
// y is bit-identical; the fused dot is reduced in a fixed order (deterministic).
#pragma once

#include "common.cuh"

namespace dkmc {

constexpr int kTmaThreads = 256;
constexpr int kTmaStages = 3;
constexpr int kTmaCap = 2048 + 8;   // elements per stage: tile (<= 1984 + 63) + alignment shift (<= 3), rounded to 4
constexpr size_t kTmaStageBytes = (size_t)kTmaCap * (sizeof(double) + sizeof(int));
constexpr size_t kTmaSmemBytes = kTmaStages * kTmaStageBytes + 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// same, but lets the hardware park the thread (up to ~`hint_ns`) instead of re-polling: a spinning
// lane otherwise takes issue slots from the warps that do the arithmetic
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity, uint32_t hint_ns = 20000u) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MODE 0: y = A x.  MODE 1: also dot(w, y).  MODE 2: y = w - A x and sum(y^2 dinv).
template <int MODE>
__global__ void __launch_bounds__(kTmaThreads) spmv_tma_kernel(
    int num_tiles, int nnz, const int *__restrict__ row_ptr, const int *__restrict__ col, const double *__restrict__ val,
    const double *__restrict__ x, double *__restrict__ y, const int4 *__restrict__ tile_info,
    const double *__restrict__ w, const double *__restrict__ dinv, double *partials, unsigned int *counter,
    double *dot_out, const int *done_flag) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    auto val_s = [&](int s) { return reinterpret_cast<double *>(smem_raw + (size_t)s * kTmaStageBytes); };
    auto col_s = [&](int s) {
        return reinterpret_cast<int *>(smem_raw + (size_t)s * kTmaStageBytes + (size_t)kTmaCap * sizeof(double));
    };
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kTmaStages * kTmaStageBytes);
    const int tid = threadIdx.x;
    const int nj = blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&bars[s], 1);
        fence_proxy_async();
    }
    __syncthreads();

    // producer: bulk-copies tile j (of this CTA) into stage j % kStages.  The bulk part is a
    // multiple of four elements starting at ka = k0 & ~3 (16-byte aligned in both arrays) and never
    // crosses the end of the arrays; at most three trailing elements are fetched with plain loads.
    auto issue = [&](int j) {
        const int4 ti = tile_info[blockIdx.x + j * gridDim.x];
        const int s = j % kTmaStages;
        const int ka = ti.z & ~3;
        int n = (ti.w - ka + 3) & ~3;
        const int lim = (nnz - ka) & ~3;
        if (n > lim) n = lim;
        if (n > kTmaCap) n = 0;  // oversize tile: consumers take the direct path
        if (n > 0) {
            mbar_expect_tx(&bars[s], (uint32_t)n * 12u);
            tma_load_1d(val_s(s), val + ka, (uint32_t)n * 8u, &bars[s]);
            tma_load_1d(col_s(s), col + ka, (uint32_t)n * 4u, &bars[s]);
        } else {
            mbar_expect_tx(&bars[s], 0u);
        }
    };
    if (tid == 0)
        for (int j = 0; j < kTmaStages - 1 && j < nj; ++j) issue(j);

    double local = 0.0;
    for (int j = 0; j < nj; ++j) {
        const int s = j % kTmaStages;
        const uint32_t parity = (uint32_t)(j / kTmaStages) & 1u;
        if (tid == 0 && j + kTmaStages - 1 < nj) {
            fence_proxy_async();  // generic-proxy writes to that stage (products) before the async-proxy refill
            issue(j + kTmaStages - 1);
        }
        const int4 ti = tile_info[blockIdx.x + j * gridDim.x];
        const int r0 = ti.x, r1 = ti.y, k0 = ti.z, k1 = ti.w;
        const int ka = k0 & ~3;
        const int cnt = k1 - ka;
        // this thread's first row bounds, requested while the tile is still in flight
        int my_r = r0 + tid, ra = 0, rb = 0;
        if (my_r < r1) { ra = row_ptr[my_r]; rb = row_ptr[my_r + 1]; }
        mbar_wait(&bars[s], parity);
        if (cnt <= kTmaCap) {
            double *vs = val_s(s);
            const int *cs = col_s(s);
            int nb = (cnt + 3) & ~3;
            const int lim = (nnz - ka) & ~3;
            if (nb > lim) nb = lim;
            // trailing elements beyond the bulk copy (last tile of the matrix only)
            for (int e = nb + tid; e < cnt; e += kTmaThreads) vs[e] = val[ka + e] * __ldg(x + col[ka + e]);
            // products in place, eight gathers in flight per thread
            for (int base = 0; base < nb; base += 8 * kTmaThreads) {
                double xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    int e = base + u * kTmaThreads + tid;
                    xv[u] = e < nb ? __ldg(x + cs[e]) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    int e = base + u * kTmaThreads + tid;
                    if (e < nb) vs[e] *= xv[u];
                }
            }
            __syncthreads();
            for (int r = my_r; r < r1; r += kTmaThreads) {
                if (r != my_r) { ra = row_ptr[r]; rb = row_ptr[r + 1]; }
                double sum = 0.0;
#pragma unroll 4
                for (int k = ra - ka; k < rb - ka; ++k) sum += vs[k];
                if (MODE == 2) { sum = w[r] - sum; local += sum * sum * dinv[r]; }
                y[r] = sum;
                if (MODE == 1) local += w[r] * sum;
            }
        } else {  // rows too long for a stage: direct path
            for (int r = my_r; r < r1; r += kTmaThreads) {
                double sum = 0.0;
                for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) sum += val[k] * __ldg(x + col[k]);
                if (MODE == 2) { sum = w[r] - sum; local += sum * sum * dinv[r]; }
                y[r] = sum;
                if (MODE == 1) local += w[r] * sum;
            }
        }
        fence_proxy_async();  // order this thread's shared-memory writes before the stage's async refill
        __syncthreads();      // stage s is free for the producer
    }
    if (MODE != 0) {
        double tot = block_sum(local, red);
        grid_sum_finish(tot, partials, counter, dot_out, red);
    }
}

}  // namespace dkmc
