// devicekmc-b200 — block-level inclusive scan (reduce-then-scan, warp-shuffle based).
// Replaces cub::DeviceScan::InclusiveSum (iterative_solvers_gpu.cu:988,995) and
// thrust::inclusive_scan (kmc_events.cu:214).  Deterministic: the association order depends
// only on n, never on scheduling.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace dkmc {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                       // per thread, warp-striped
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 items per block

template <typename T>
__device__ __forceinline__ T warp_scan_t(T v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// pass 1: total of each 2048-item tile
template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_tile_totals(const T *__restrict__ in, long long n,
                                                                T *__restrict__ tile_total) {
    __shared__ T sh[kScanThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * kScanTile + (long long)w * (32 * kScanItems);
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long i = base + k * 32 + lane;
        T v = i < n ? in[i] : T(0);
        // same association as pass 3: chunk-by-chunk warp scans with a running carry
        T inc = warp_scan_t<T>(v, lane);
        T chunk = __shfl_sync(0xffffffffu, inc, 31);
        acc = (k == 0) ? chunk : acc + chunk;
    }
    if (lane == 0) sh[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = sh[0];
        for (int i = 1; i < kScanThreads / 32; ++i) t += sh[i];
        tile_total[blockIdx.x] = t;
    }
}

// pass 2: exclusive scan of the tile totals by one block (sequential carry across chunks)
template <typename T>
__global__ void __launch_bounds__(1024) scan_tile_offsets(T *tile_total, int ntiles) {
    __shared__ T sh[32];
    __shared__ T carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = T(0);
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + threadIdx.x;
        T v = i < ntiles ? tile_total[i] : T(0);
        T inc = warp_scan_t<T>(v, lane);
        if (lane == 31) sh[w] = inc;
        __syncthreads();
        if (w == 0) {
            T s = sh[lane];
            s = warp_scan_t<T>(s, lane);
            sh[lane] = s;
        }
        __syncthreads();
        T woff = w > 0 ? sh[w - 1] : T(0);
        T carry = carry_s;
        T incl = carry + (woff + inc);
        if (i < ntiles) tile_total[i] = incl - v;  // exclusive (exact for ints; doubles use pass-3 form)
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

// doubles: keep the exclusive offsets free of the "incl - v" cancellation
template <typename T>
__global__ void __launch_bounds__(1024) scan_tile_offsets_seq(T *tile_total, int ntiles) {
    // a single warp walks the totals in order: offsets[i] = sum of totals before i
    const int lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    T carry = T(0);
    for (int base = 0; base < ntiles; base += 32) {
        int i = base + lane;
        T v = i < ntiles ? tile_total[i] : T(0);
        T inc = warp_scan_t<T>(v, lane);
        T prev = __shfl_up_sync(0xffffffffu, inc, 1);
        T excl = lane == 0 ? carry : carry + prev;
        if (i < ntiles) tile_total[i] = excl;
        carry = carry + __shfl_sync(0xffffffffu, inc, 31);
    }
}

// pass 3: scan each tile and add its offset
template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_apply(const T *__restrict__ in, long long n,
                                                          const T *__restrict__ tile_offset,
                                                          T *__restrict__ out) {
    __shared__ T sh[kScanThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * kScanTile + (long long)w * (32 * kScanItems);
    T v[kScanItems];
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long i = base + k * 32 + lane;
        v[k] = i < n ? in[i] : T(0);
    }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = warp_scan_t<T>(v[k], lane);
        T chunk = __shfl_sync(0xffffffffu, v[k], 31);
        T prev = acc;
        acc = (k == 0) ? chunk : acc + chunk;
        if (k > 0) v[k] = prev + v[k];
    }
    if (lane == 0) sh[w] = acc;
    __syncthreads();
    T woff = tile_offset[blockIdx.x];
    for (int i = 0; i < w; ++i) woff += sh[i];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long i = base + k * 32 + lane;
        if (i < n) out[i] = woff + v[k];
    }
}

// host driver; `tmp` must hold ceil(n / kScanTile) items
template <typename T>
inline int inclusive_scan(dkmc_ctx *ctx, const T *d_in, long long n, T *d_out, T *tmp) {
    if (n <= 0) return DKMC_OK;
    int ntiles = ceil_div(n, kScanTile);
    DKMC_LAUNCH(ctx, scan_tile_totals<T>, ntiles, kScanThreads, 0, d_in, n, tmp);
    if (std::is_integral<T>::value) {
        DKMC_LAUNCH(ctx, scan_tile_offsets<T>, 1, 1024, 0, tmp, ntiles);
    } else {
        DKMC_LAUNCH(ctx, scan_tile_offsets_seq<T>, 1, 32, 0, tmp, ntiles);
    }
    DKMC_LAUNCH(ctx, scan_apply<T>, ntiles, kScanThreads, 0, d_in, n, tmp, d_out);
    return DKMC_OK;
}

}  // namespace dkmc
