// devicekmc-b200 — window-staged SpMV for the Kirchhoff matrix of the CG (sm_100a).
//
// A CSR SpMV moves 12 bytes per non-zero (FP64 value + int32 column) and gathers x through L1/L2,
// one 32-byte sector per non-zero in the worst case.  K, however, has only two distinct
// off-diagonal values (-high_G, -low_G; potential_solver.cpp:325-346) and its structure is static:
// the neighbours of the ~77 consecutive rows of a 2048-nnz tile fall into a handful of contiguous
// index ranges ("runs"; 9 on a cell-ordered device).  So, once per sparsity pattern, every tile gets
//   * a list of runs that cover its columns (64-byte aligned pieces of x), and
//   * a 16-bit code per non-zero: position inside the tile's window (14 bits), a "diagonal" bit,
//     and — rewritten by the assembly every KMC step — a "high_G" bit.
// The SpMV then streams 2 bytes per non-zero.  One persistent CTA per SM: a producer warp issues
// bulk (TMA) copies of the tile's codes, row pointers, diagonal entries and the x runs into a
// shared-memory ring; eleven consumer warps each take whole tiles, one row per lane, and add the
// row's products in CSR order from shared memory.  Nothing but the y stores touches global memory
// on the consumer side.  Products (v * x, rounded) and sums (left to right, rounded) are the same
// operations in the same order as the CSR kernels, so y is bit-identical to them.
//
// Algorithmic traffic per SpMV: 2 nnz (codes) + 20 m (row_ptr, diag, y) + 8 m (x, once) + tile
// headers, against the CSR contract's 12 nnz + 20 m.
#pragma once

#include "common.cuh"
#include "spmv_tma.cuh"

namespace dkmc {

constexpr int kWinThreads = 384;                 // 1 producer warp + 11 consumer warps
constexpr int kWinConsumers = kWinThreads / 32 - 1;
constexpr int kWinSlots = 16;                    // tiles resident in the ring (in flight + being summed)
constexpr int kWinRingBytes = 168 * 1024;
constexpr int kWinMaxRuns = 32;                  // one bulk copy per producer lane
constexpr int kWinBlock = 8;                     // run granularity in doubles (64 bytes)
constexpr int kWinGap = 1;                       // blocks are merged into one run when b - prev <= kWinGap
constexpr int kWinMaxTileNnz = 2048;
constexpr int kWinBatch = 8;                     // tile headers the producer prefetches at a time
constexpr size_t kWinSmemBytes = (size_t)kWinRingBytes + 2 * kWinSlots * sizeof(uint64_t) + kWinSlots * 32 + 128;

struct __align__(16) WinTileHdr {
    int r0, r1, k0, k1;          // rows [r0,r1), non-zeros [k0,k1)
    int nruns, win_len;          // x window: nruns runs, win_len doubles in total
    int chunk_bytes, pad;        // ring bytes of the staged tile
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __host__ __forceinline__ int win_round16(int bytes) { return (bytes + 15) & ~15; }

// layout of a staged tile inside the ring (all pieces 16-byte aligned, the x window 128-byte)
struct WinLayout {
    int idx_off, idx_bytes, rp_off, rp_bytes, dg_off, dg_bytes, xs_off, xs_bytes, total;
};
__device__ __host__ __forceinline__ WinLayout win_layout(int r0, int r1, int k0, int k1, int win_len) {
    WinLayout L;
    const int k0a = k0 & ~7, r0a = r0 & ~3, r0d = r0 & ~1;
    L.idx_off = 0;
    L.idx_bytes = win_round16((k1 - k0a) * 2);
    L.rp_off = L.idx_off + L.idx_bytes;
    L.rp_bytes = win_round16((r1 - r0a + 1) * 4);
    L.dg_off = L.rp_off + L.rp_bytes;
    L.dg_bytes = win_round16((r1 - r0d) * 8);
    L.xs_off = (L.dg_off + L.dg_bytes + 127) & ~127;
    L.xs_bytes = win_len * 8;
    L.total = (L.xs_off + L.xs_bytes + 127) & ~127;
    return L;
}

// ---------------------------------------------------------------- static preprocessing, one CTA per tile
// fail bits: 1 tile too long, 2 too many runs, 4 window too large, 8 chunk too large
__global__ void __launch_bounds__(256) win_build_kernel(int m, int num_tiles, const int *__restrict__ row_ptr,
                                                        const int *__restrict__ col, const int4 *__restrict__ tile_info,
                                                        unsigned short *__restrict__ code_base, WinTileHdr *__restrict__ hdr,
                                                        int2 *__restrict__ runs, int *__restrict__ fail,
                                                        int *__restrict__ max_chunk) {
    __shared__ int keys[kWinMaxTileNnz];
    __shared__ int run_first[kWinMaxRuns], run_off[kWinMaxRuns], run_len[kWinMaxRuns];
    __shared__ int s_nruns, s_win;
    const int t = blockIdx.x;
    const int4 ti = tile_info[t];
    const int r0 = ti.x, r1 = ti.y, k0 = ti.z, k1 = ti.w;
    const int n = k1 - k0;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (n > kWinMaxTileNnz) {
        if (tid == 0) {
            atomicOr(fail, 1);
            hdr[t] = WinTileHdr{r0, r1, k0, k1, 0, 0, 0, 0};
        }
        return;
    }
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
        if (win > 0x3fff) bad |= 4;
        const WinLayout L = win_layout(r0, r1, k0, k1, win);
        if (L.total > kWinRingBytes / 4) bad |= 8;
        if (bad) atomicOr(fail, bad);
        atomicMax(max_chunk, L.total);
        s_nruns = nruns;
        s_win = win;
        hdr[t] = WinTileHdr{r0, r1, k0, k1, nruns, win, L.total, 0};
        for (int q = 0; q < kWinMaxRuns; ++q)
            runs[(size_t)t * kWinMaxRuns + q] =
                q < nruns ? make_int2(run_first[q] * kWinBlock, run_off[q] | (run_len[q] << 16)) : make_int2(0, 0);
    }
    __syncthreads();
    const int nruns = s_nruns;
    for (int r = r0 + tid; r < r1; r += nt) {
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            const int c = col[k];
            const int b = c / kWinBlock;
            int q = 0;
            while (q + 1 < nruns && run_first[q + 1] <= b) ++q;
            int l = run_off[q] + (c - run_first[q] * kWinBlock);
            if (l < 0 || l > 0x3fff) l = 0;  // only when the tile is flagged as failed
            code_base[k] = (unsigned short)(l | (c == r ? 0x8000 : 0));
        }
    }
}

// padded copy of row_ptr (bulk copies read whole 16-byte groups)
__global__ void win_copy_rowptr_kernel(int m, int padded, const int *__restrict__ row_ptr, int *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += gridDim.x * blockDim.x)
        out[i] = row_ptr[i <= m ? i : m];
}

// ---------------------------------------------------------------- the SpMV
struct WinMatrix {
    const unsigned short *code;   // [nnz + pad] window position | 0x4000 high_G | 0x8000 diagonal
    const int *rp;                // padded row_ptr
    const double *diag;           // [m + pad]
    const WinTileHdr *hdr;
    const int2 *runs;             // [num_tiles][kWinMaxRuns]: (first x index, offset | length << 16)
    double m_high, m_low;         // -high_G, -low_G
    int num_tiles;
};

struct __align__(16) WinSlotMeta { int r0, r1, k0, k1, base, rp_off, dg_off, xs_off; };

// MODE 0: y = A x.  MODE 1: also dot(x, y) (the CG's p.Ap).  MODE 2: y = w - A x and sum(y^2 dinv).
template <int MODE>
__global__ void __launch_bounds__(kWinThreads, 1) spmv_win_kernel(
    WinMatrix A, const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w,
    const double *__restrict__ dinv, double *partials, unsigned int *counter, double *dot_out, const int *done_flag,
    int dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    unsigned char *ring = smem_raw;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kWinRingBytes);
    uint64_t *empty = full + kWinSlots;
    WinSlotMeta *meta = reinterpret_cast<WinSlotMeta *>(empty + kWinSlots);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // contiguous chunk of tiles per CTA
    const int per = (A.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = min((int)blockIdx.x * per, A.num_tiles);
    const int nj = min(t_begin + per, A.num_tiles) - t_begin;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kWinSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_proxy_async();
    }
    __syncthreads();

    double local = 0.0;
    if (warp == 0) {
        // ------------------------------------------------ producer warp
        unsigned head = 0;     // virtual ring offset of the next allocation
        int tail_j = 0;        // oldest tile whose ring space has not been reclaimed
        unsigned vstart_mine = 0;  // lane s keeps the virtual start of the tile in slot s
        for (int jb = 0; jb < nj; jb += kWinBatch) {
            // headers and run descriptors of the next kWinBatch tiles: all loads in flight together
            const int nb = min(kWinBatch, nj - jb);
            int4 ha = make_int4(0, 0, 0, 0), hb = make_int4(0, 0, 0, 0);
            if (lane < nb) {
                const int4 *hp = reinterpret_cast<const int4 *>(A.hdr + t_begin + jb + lane);
                ha = __ldg(hp);
                hb = __ldg(hp + 1);
            }
            int2 rd[kWinBatch];
#pragma unroll
            for (int u = 0; u < kWinBatch; ++u)
                rd[u] = u < nb ? __ldg(A.runs + (size_t)(t_begin + jb + u) * kWinMaxRuns + lane) : make_int2(0, 0);
#pragma unroll
            for (int u = 0; u < kWinBatch; ++u) {
                if (u >= nb) break;
                const int j = jb + u;
                const int r0 = __shfl_sync(0xffffffffu, ha.x, u), r1 = __shfl_sync(0xffffffffu, ha.y, u);
                const int k0 = __shfl_sync(0xffffffffu, ha.z, u), k1 = __shfl_sync(0xffffffffu, ha.w, u);
                const int nruns = __shfl_sync(0xffffffffu, hb.x, u), win_len = __shfl_sync(0xffffffffu, hb.y, u);
                const WinLayout L = win_layout(r0, r1, k0, k1, win_len);
                const int slot = j % kWinSlots;
                // ring allocation (a tile is contiguous: skip the end of the ring if it does not fit)
                unsigned pos = head % (unsigned)kWinRingBytes;
                if (pos + (unsigned)L.total > (unsigned)kWinRingBytes) head += (unsigned)kWinRingBytes - pos;
                while (tail_j < j) {
                    const unsigned vs = __shfl_sync(0xffffffffu, vstart_mine, tail_j % kWinSlots);
                    const bool need_space = head + (unsigned)L.total - vs > (unsigned)kWinRingBytes;
                    const bool need_slot = tail_j <= j - kWinSlots;
                    if (!need_space && !need_slot) break;
                    mbar_wait(&empty[tail_j % kWinSlots], (uint32_t)(tail_j / kWinSlots) & 1u);
                    ++tail_j;
                }
                if (lane == slot) vstart_mine = head;
                const int base = (int)(head % (unsigned)kWinRingBytes);
                head += (unsigned)L.total;
                unsigned char *chunk = ring + base;
                const int k0a = k0 & ~7, r0a = r0 & ~3, r0d = r0 & ~1;
                if (lane == 0) {
                    meta[slot] = WinSlotMeta{r0, r1, k0, k1, base, L.rp_off, L.dg_off, L.xs_off};
                    mbar_expect_tx(&full[slot], (uint32_t)(L.idx_bytes + L.rp_bytes + L.dg_bytes + ((dbg & 2) ? 0 : L.xs_bytes)));
                    tma_load_1d(chunk + L.idx_off, A.code + k0a, (uint32_t)L.idx_bytes, &full[slot]);
                    tma_load_1d(chunk + L.rp_off, A.rp + r0a, (uint32_t)L.rp_bytes, &full[slot]);
                    tma_load_1d(chunk + L.dg_off, A.diag + r0d, (uint32_t)L.dg_bytes, &full[slot]);
                }
                __syncwarp();
                if (lane < nruns && !(dbg & 2)) {
                    const int start = rd[u].x, off = rd[u].y & 0xffff, len = (rd[u].y >> 16) & 0xffff;
                    tma_load_1d(chunk + L.xs_off + off * 8, x + start, (uint32_t)len * 8u, &full[slot]);
                }
            }
        }
    } else {
        // ------------------------------------------------ consumer warps: whole tiles, one row per lane
        const int c = warp - 1;
        for (int j = c; j < nj; j += kWinConsumers) {
            const int slot = j % kWinSlots;
            mbar_wait(&full[slot], (uint32_t)(j / kWinSlots) & 1u);
            const WinSlotMeta mt = meta[slot];
            const unsigned char *chunk = ring + mt.base;
            const unsigned short *code_s = reinterpret_cast<const unsigned short *>(chunk) - (mt.k0 & ~7);
            const int *rp_s = reinterpret_cast<const int *>(chunk + mt.rp_off) - (mt.r0 & ~3);
            const double *dg_s = reinterpret_cast<const double *>(chunk + mt.dg_off) - (mt.r0 & ~1);
            const double *xs = reinterpret_cast<const double *>(chunk + mt.xs_off);
            for (int r = mt.r0 + lane; r < mt.r1 && !(dbg & 1); r += 32) {
                const int ra = rp_s[r], rb = rp_s[r + 1];
                const double dg = dg_s[r];
                double sum = 0.0, xd = 0.0;
#pragma unroll 4
                for (int k = ra; k < rb; ++k) {
                    const unsigned e = code_s[k];
                    const double xv = xs[e & 0x3fffu];
                    double v = (e & 0x4000u) ? A.m_high : A.m_low;
                    if (e & 0x8000u) { v = dg; xd = xv; }
                    sum = __dadd_rn(sum, __dmul_rn(v, xv));
                }
                if (MODE == 2) { sum = w[r] - sum; local += sum * sum * dinv[r]; }
                y[r] = sum;
                if (MODE == 1) local += xd * sum;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
    if (MODE != 0) {
        double tot = block_sum(local, red);
        grid_sum_finish(tot, partials, counter, dot_out, red);
    }
}

}  // namespace dkmc
