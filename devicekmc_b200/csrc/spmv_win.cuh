// devicekmc-b200 — window-staged SpMV for the Kirchhoff matrix of the CG (sm_100a).
//
// A CSR SpMV moves 12 bytes per non-zero (FP64 value + int32 column) and gathers x through L1/L2,
// one 32-byte sector per non-zero in the worst case.  K, however, has only two distinct
// off-diagonal values (-high_G, -low_G; potential_solver.cpp:325-346) and its structure is static:
// the neighbours of the ~77 consecutive rows of a 2048-nnz tile fall into a handful of contiguous
// index ranges ("runs"; 9 on a cell-ordered device).  So, once per sparsity pattern, every tile gets
// a BLOB in global memory holding
//   * the runs that cover its columns (64-byte aligned pieces of x, the tile's "window"),
//   * a 16-bit code per non-zero: position inside the window (14 bits), a "diagonal" bit and — set by
//     the assembly every KMC step — a "high_G" bit,
//   * 16-bit row starts, the rows ordered by decreasing length (so the 32 lanes of a warp work on rows
//     of similar length), and — per step — the diagonal entries.
// The SpMV then streams ~2.6 bytes per non-zero.  One persistent CTA per SM, warp-specialised:
//   warp 0        one bulk (TMA) copy per tile, blob -> shared-memory ring, mbarrier `full`
//   warps 1-2     window loaders: read the runs from the staged blob, issue 16-byte cp.async copies of
//                 the x pieces next to it, mbarrier `ready` (cp.async.mbarrier.arrive)
//   warps 3-15    consumers: a warp takes a whole tile, one row per lane, and adds the row's products
//                 in CSR order out of shared memory; only the y stores go to global memory
// Products (v * x, rounded) and sums (left to right, rounded) are the same operations in the same
// order as in the CSR kernels, so y is bit-identical to theirs.
#pragma once

#include "common.cuh"
#include "spmv_tma.cuh"

namespace dkmc {

constexpr int kWinThreads = 1024;
constexpr int kWinLoaders = 4;
constexpr int kWinGroup = 3;                     // consumer warps that share one tile
constexpr int kWinConsumers = kWinThreads / 32 - 1 - kWinLoaders;   // 27 = 9 groups
constexpr int kWinSlots = 32;                    // tiles resident in the ring
constexpr int kWinRingBytes = 176 * 1024;
constexpr int kWinMaxRuns = 32;
constexpr int kWinBlock = 8;                     // run granularity in doubles (64 bytes)
constexpr int kWinGap = 1;                       // blocks closer than this are merged into one run
constexpr int kWinMaxTileNnz = 2048;
constexpr int kWinBatch = 4;                     // tile headers the producer prefetches at a time
constexpr size_t kWinSmemBytes = (size_t)kWinRingBytes + 3 * kWinSlots * sizeof(uint64_t) + 2 * kWinSlots * sizeof(int) + 128;

// byte offsets inside a tile's blob; a pure function of the tile's extent
struct WinBlob {
    int o_runs, o_codes, o_rp, o_ord, o_dg, bytes;
};
__device__ __host__ __forceinline__ int win_up(int v, int a) { return (v + a - 1) / a * a; }
__device__ __host__ __forceinline__ WinBlob win_blob(int rows, int n) {
    WinBlob b;
    b.o_runs = 32;                                    // header: r0, r1, k0, k1, nruns, win_len, -, -
    b.o_codes = b.o_runs + kWinMaxRuns * 8;           // int2 per run: (first x index, offset | length << 16)
    b.o_rp = b.o_codes + win_up(2 * n, 16);           // uint16 row starts relative to k0, rows + 1 of them
    b.o_ord = b.o_rp + win_up(2 * (rows + 1), 16);    // uint16 row order (longest rows first)
    b.o_dg = b.o_ord + win_up(2 * rows, 16);          // double diag[rows] (in row order r0..r1)
    b.bytes = win_up(b.o_dg + 8 * rows, 128);
    return b;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- static preprocessing, one CTA per tile
// tile_plan[t] = (blob offset / 128, blob bytes, window doubles [out], staged bytes = blob + window [out]).
// fail bits: 1 tile too long, 2 too many runs, 4 window too large, 8 staged tile larger than the ring
__global__ void __launch_bounds__(256) win_build_kernel(int m, int num_tiles, const int *__restrict__ row_ptr,
                                                        const int *__restrict__ col, const int4 *__restrict__ tile_info,
                                                        int4 *__restrict__ tile_plan, unsigned char *__restrict__ blobs,
                                                        unsigned short *__restrict__ code_base, int *__restrict__ code_pos,
                                                        int *__restrict__ diag_pos, int *__restrict__ fail,
                                                        int *__restrict__ max_chunk) {
    __shared__ int keys[kWinMaxTileNnz];
    __shared__ int run_first[kWinMaxRuns], run_off[kWinMaxRuns], run_len[kWinMaxRuns];
    __shared__ int s_nruns;
    const int t = blockIdx.x;
    const int4 ti = tile_info[t];
    const int r0 = ti.x, r1 = ti.y, k0 = ti.z, k1 = ti.w;
    const int n = k1 - k0, rows = r1 - r0;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (n > kWinMaxTileNnz || rows > kWinMaxTileNnz) {
        if (tid == 0) {
            atomicOr(fail, 1);
            tile_plan[t].z = 0;
        }
        return;
    }
    const WinBlob B = win_blob(rows, n);
    unsigned char *blob = blobs + (size_t)tile_plan[t].x * 128;
    int P = 32;
    while (P < n) P <<= 1;
    for (int i = tid; i < P; i += nt) keys[i] = i < n ? (col[k0 + i] / kWinBlock) : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += nt) {
                int q = i ^ j;
                if (q > i) {
                    int a = keys[i], b = keys[q];
                    bool asc = (i & k) == 0;
                    if ((a > b) == asc) { keys[i] = b; keys[q] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        // runs of 64-byte blocks; blocks closer than `gap` are merged into one run.  The gap is doubled
        // until the tile needs no more than kWinMaxRuns runs (a few tiles straddle two grid columns).
        int nruns = 0, win = 0, bad = 0;
        for (int gap = kWinGap; gap < 0x40000000; gap <<= 1) {
            int first = 0, last = 0, prev = -0x40000000;
            bool open = false;
            nruns = 0;
            win = 0;
            for (int i = 0; i < n; ++i) {
                const int b = keys[i];
                if (b == prev) continue;
                if (!open || b - prev > gap) {
                    if (open) {
                        const int len = (last - first + 1) * kWinBlock;
                        if (nruns < kWinMaxRuns) { run_first[nruns] = first; run_off[nruns] = win; run_len[nruns] = len; }
                        ++nruns;
                        win += len;
                    }
                    first = b;
                    open = true;
                }
                last = b;
                prev = b;
            }
            if (open) {
                const int len = (last - first + 1) * kWinBlock;
                if (nruns < kWinMaxRuns) { run_first[nruns] = first; run_off[nruns] = win; run_len[nruns] = len; }
                ++nruns;
                win += len;
            }
            if (nruns <= kWinMaxRuns || win > 0x3fff) break;
        }
        if (nruns > kWinMaxRuns) { bad |= 2; nruns = kWinMaxRuns; }
        if (win > 0x3fff) { bad |= 4; win = 0; }
        const int chunk = win_up(B.bytes + 8 * win, 128);
        if (chunk > kWinRingBytes) bad |= 8;
        if (bad) atomicOr(fail, bad);
        atomicMax(max_chunk, chunk);
        s_nruns = nruns;
        tile_plan[t].z = win;
        tile_plan[t].w = chunk;
        int *h = reinterpret_cast<int *>(blob);
        h[0] = r0; h[1] = r1; h[2] = k0; h[3] = k1; h[4] = nruns; h[5] = win; h[6] = 0; h[7] = 0;
        int2 *rd = reinterpret_cast<int2 *>(blob + B.o_runs);
        for (int q = 0; q < kWinMaxRuns; ++q)
            rd[q] = q < nruns ? make_int2(run_first[q] * kWinBlock, run_off[q] | (run_len[q] << 16)) : make_int2(0, 0);
    }
    __syncthreads();
    const int nruns = s_nruns;
    unsigned short *codes = reinterpret_cast<unsigned short *>(blob + B.o_codes);
    unsigned short *rp16 = reinterpret_cast<unsigned short *>(blob + B.o_rp);
    unsigned short *ord = reinterpret_cast<unsigned short *>(blob + B.o_ord);
    const long long code_index0 = ((long long)tile_plan[t].x * 128 + B.o_codes) / 2;   // in halfwords from `blobs`
    const long long diag_index0 = ((long long)tile_plan[t].x * 128 + B.o_dg) / 8;      // in doubles from `blobs`
    for (int r = r0 + tid; r < r1; r += nt) {
        const int ra = row_ptr[r], rb = row_ptr[r + 1];
        for (int k = ra; k < rb; ++k) {
            const int c = col[k];
            const int b = c / kWinBlock;
            int q = 0;
            while (q + 1 < nruns && run_first[q + 1] <= b) ++q;
            int l = run_off[q] + (c - run_first[q] * kWinBlock);
            if (l < 0 || l > 0x3fff) l = 0;  // only when the tile is flagged as failed
            const unsigned short e = (unsigned short)(l | (c == r ? 0x8000 : 0));
            code_base[k] = e;
            codes[k - k0] = e;
        }
        rp16[r - r0] = (unsigned short)(ra - k0);
        if (r + 1 == r1) rp16[rows] = (unsigned short)(rb - k0);
        code_pos[r] = (int)(code_index0 + (ra - k0));
        diag_pos[r] = (int)(diag_index0 + (r - r0));
        // rank among the tile's rows by (length descending, row ascending)
        const int len = rb - ra;
        int rank = 0;
        for (int q = r0; q < r1; ++q) {
            const int lq = row_ptr[q + 1] - row_ptr[q];
            rank += (lq > len || (lq == len && q < r)) ? 1 : 0;
        }
        ord[rank] = (unsigned short)(r - r0);
    }
}

// ---------------------------------------------------------------- the SpMV
struct WinMatrix {
    const unsigned char *blobs;
    const int4 *plan;             // per tile: (blob offset / 128, blob bytes, window doubles, -)
    double m_high, m_low;         // -high_G, -low_G
    int num_tiles;
};

// MODE 0: y = A x.  MODE 1: also dot(x, y) (the CG's p.Ap).  MODE 2: y = w - A x and sum(y^2 dinv).
template <int MODE>
__global__ void __launch_bounds__(1024, 1) spmv_win_kernel(
    WinMatrix A, const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w,
    const double *__restrict__ dinv, double *partials, unsigned int *counter, double *dot_out, const int *done_flag,
    int dbg, int ring_bytes, int n_loaders, long long *prof) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    unsigned char *ring = smem_raw;
    const unsigned R = (unsigned)ring_bytes;
    const int n_consumers = (int)(blockDim.x >> 5) - 1 - n_loaders;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + ring_bytes);   // blob landed (TMA)
    uint64_t *ready = full + kWinSlots;                                        // window landed (cp.async)
    uint64_t *empty = ready + kWinSlots;                                       // tile consumed
    int *slot_base = reinterpret_cast<int *>(empty + kWinSlots);
    unsigned *vstart = reinterpret_cast<unsigned *>(slot_base + kWinSlots);   // producer's bookkeeping
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // contiguous chunk of tiles per CTA
    const int per = (A.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = min((int)blockIdx.x * per, A.num_tiles);
    const int nj = min(t_begin + per, A.num_tiles) - t_begin;

    if (tid == 0) {
        for (int s = 0; s < kWinSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 1); mbar_init(&empty[s], kWinGroup); }
        fence_proxy_async();
    }
    __syncthreads();

    double local = 0.0;
    if (warp == 0) {
        // ------------------------------------------------ producer (one thread): one bulk copy per tile
        if (lane == 0) {
            long long t_wait = 0, t_start = clock64();
            unsigned vhead = 0;     // virtual ring offset of the next allocation
            unsigned pos = 0;       // the same, inside the ring
            int tail_j = 0;         // oldest tile whose ring space has not been reclaimed
            int4 nxt[kWinBatch];
#pragma unroll
            for (int u = 0; u < kWinBatch; ++u) nxt[u] = u < nj ? __ldg(A.plan + t_begin + u) : make_int4(0, 0, 0, 0);
            for (int jb = 0; jb < nj; jb += kWinBatch) {
                int4 cur[kWinBatch];
#pragma unroll
                for (int u = 0; u < kWinBatch; ++u) {
                    cur[u] = nxt[u];
                    nxt[u] = jb + kWinBatch + u < nj ? __ldg(A.plan + t_begin + jb + kWinBatch + u) : make_int4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < kWinBatch; ++u) {
                    const int j = jb + u;
                    if (j >= nj) break;
                    const unsigned bbytes = (unsigned)cur[u].y, total = (unsigned)cur[u].w;
                    const int slot = j % kWinSlots;
                    // ring allocation (a tile is contiguous: skip the end of the ring if it does not fit)
                    if (pos + total > R) { vhead += R - pos; pos = 0; }
                    while (tail_j < j) {
                        const bool need_space = vhead + total - vstart[tail_j % kWinSlots] > R;
                        const bool need_slot = tail_j <= j - kWinSlots;
                        if (!need_space && !need_slot) break;
                        const long long t0 = clock64();
                        mbar_wait_parked(&empty[tail_j % kWinSlots], (uint32_t)(tail_j / kWinSlots) & 1u);
                        t_wait += clock64() - t0;
                        ++tail_j;
                    }
                    vstart[slot] = vhead;
                    slot_base[slot] = (int)pos;
                    if (dbg & 4) {  // timing experiment: no copy at all
                        mbar_expect_tx(&full[slot], 0u);
                    } else {
                        mbar_expect_tx(&full[slot], bbytes);
                        tma_load_1d(ring + pos, A.blobs + (size_t)cur[u].x * 128, bbytes, &full[slot]);
                    }
                    pos += total;
                    vhead += total;
                }
            }
            if (prof) { atomicAdd((unsigned long long *)prof + 0, (unsigned long long)t_wait); atomicAdd((unsigned long long *)prof + 1, (unsigned long long)(clock64() - t_start)); }
        }
    } else if (warp <= n_loaders) {
        // ------------------------------------------------ window loaders: one bulk copy per run of x,
        // issued by one lane each, next to the staged blob
        long long t_wait = 0, t_start = clock64();
        for (int j = warp - 1; j < nj; j += n_loaders) {
            const int slot = j % kWinSlots;
            const long long t0 = clock64();
            if (lane == 0) mbar_wait_parked(&full[slot], (uint32_t)(j / kWinSlots) & 1u);
            __syncwarp();
            t_wait += clock64() - t0;
            __syncwarp();
            unsigned char *chunk = ring + slot_base[slot];
            const int *h = reinterpret_cast<const int *>(chunk);
            const int rows = h[1] - h[0], n = h[3] - h[2], nruns = h[4], win_len = h[5];
            const WinBlob B = win_blob(rows, n);
            const int2 *rd = reinterpret_cast<const int2 *>(chunk + B.o_runs);
            double *xs = reinterpret_cast<double *>(chunk + B.bytes);
            if (lane == 0) mbar_expect_tx(&ready[slot], (dbg & 2) ? 0u : (uint32_t)win_len * 8u);
            __syncwarp();
            if (lane < nruns && !(dbg & 2)) {
                const int2 d = rd[lane];
                const int off = d.y & 0xffff, len = (d.y >> 16) & 0xffff;
                tma_load_1d(xs + off, x + d.x, (uint32_t)len * 8u, &ready[slot]);
            }
        }
        if (prof && lane == 0) { atomicAdd((unsigned long long *)prof + 2, (unsigned long long)t_wait); atomicAdd((unsigned long long *)prof + 3, (unsigned long long)(clock64() - t_start)); }
    } else {
        // ------------------------------------------------ consumers: whole tiles, one row per lane
        // a group of kWinGroup warps shares a tile (warp `sub` takes the 32-row passes sub, sub + G, ...):
        // the tile's stay in the ring — what bounds the number of tiles in flight — is G times shorter
        const int c = warp - 1 - n_loaders;
        const int gid = c / kWinGroup, sub = c % kWinGroup, n_groups = n_consumers / kWinGroup;
        long long t_wait = 0, t_start = clock64();
        for (int j = gid; j < nj && gid < n_groups; j += n_groups) {
            const int slot = j % kWinSlots;
            const long long t0 = clock64();
            if (lane == 0) mbar_wait_parked(&ready[slot], (uint32_t)(j / kWinSlots) & 1u);
            __syncwarp();
            t_wait += clock64() - t0;
            __syncwarp();
            const unsigned char *chunk = ring + slot_base[slot];
            const int *h = reinterpret_cast<const int *>(chunk);
            const int r0 = h[0], rows = h[1] - h[0], n = h[3] - h[2];
            const WinBlob B = win_blob(rows, n);
            const unsigned short *code_s = reinterpret_cast<const unsigned short *>(chunk + B.o_codes);
            const unsigned short *rp_s = reinterpret_cast<const unsigned short *>(chunk + B.o_rp);
            const unsigned short *ord_s = reinterpret_cast<const unsigned short *>(chunk + B.o_ord);
            const double *dg_s = reinterpret_cast<const double *>(chunk + B.o_dg);
            const double *xs = reinterpret_cast<const double *>(chunk + B.bytes);
            for (int i = sub * 32 + lane; i < rows && !(dbg & 1); i += 32 * kWinGroup) {
                const int lr = (dbg & 8) ? i : ord_s[i];
                const int ra = rp_s[lr], rb = rp_s[lr + 1];
                const double dg = dg_s[lr];
                double sum = 0.0, xd = 0.0;
#pragma unroll 4
                for (int k = ra; k < rb; ++k) {
                    const unsigned e = code_s[k];
                    const double xv = xs[e & 0x3fffu];
                    double v = (e & 0x4000u) ? A.m_high : A.m_low;
                    if (e & 0x8000u) { v = dg; xd = xv; }
                    sum = __dadd_rn(sum, __dmul_rn(v, xv));
                }
                const int r = r0 + lr;
                if (MODE == 2) { sum = w[r] - sum; local += sum * sum * dinv[r]; }
                y[r] = sum;
                if (MODE == 1) local += xd * sum;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        if (prof && lane == 0) { atomicAdd((unsigned long long *)prof + 4, (unsigned long long)t_wait); atomicAdd((unsigned long long *)prof + 5, (unsigned long long)(clock64() - t_start)); }
    }
    if (MODE != 0) {
        double tot = block_sum(local, red);
        grid_sum_finish(tot, partials, counter, dot_out, red);
    }
}

}  // namespace dkmc
