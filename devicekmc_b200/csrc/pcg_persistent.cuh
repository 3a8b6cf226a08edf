// devicekmc-b200 — the preconditioned CG of a5 as ONE persistent kernel per solve, for one GPU and for
// x-slab partitioned ranks alike (replaces the reference's launch sequence, iterative_solvers_gpu.cu:424-455:
// cusparseSpMV + 8 cuBLAS-1 calls + 4 host-synchronous scalar reads per iteration).
//
// Recurrence: Chronopoulos-Gear CG — algebraically the CG of iterative_solvers_gpu.cu:424-455 on the
// preconditioned system, rearranged so that both inner products of an iteration are taken at the same
// point (one reduction per iteration instead of two):
//     p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; w = A u
//     gamma = (r,u), delta = (w,u)            <- the ONE reduction
//     beta = gamma / gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old)
// M^-1 = D^-1 + W E^-1 W^T (Jacobi + one coarse unknown per uncharged-vacancy cluster, solver.cu).  The cluster
// sums W^T r follow the same recurrences (W^T s = W^T w + beta W^T s, W^T r -= alpha W^T s) from W^T w, which is
// computed beside the SpMV.  A cluster whose members all lie in one rank's rows (almost all of them) is that
// rank's private business: its sum is read back from local memory after the barrier.  Only the clusters that
// straddle a slab face travel in the reduction, beside the two scalars — so the preconditioner costs no extra
// synchronisation and a few words of NVLink traffic.
//
// An iteration is two phases separated by grid barriers:
//   V  vector phase over the own rows (p, s, x, r, u; gamma partial); boundary rows of u are stored straight
//      into the neighbours' windows over NVLink.  Barrier H: local arrival counter; the last CTA raises the
//      halo flags at the neighbours; every CTA waits for the local release and the neighbours' flags.
//   S  SpMV w = A u over the own nnz tiles (delta partial) + cluster rows (W^T w partial).  Barrier R: the
//      last CTA to arrive adds the CTA partials in index order and stores [gamma, delta, W^T w] into every
//      rank's slot, then raises its flag everywhere; every CTA waits for all ranks' flags in its own window
//      and adds the slots in rank order — identical bits on all ranks, and no second local barrier.
// One GPU is the same code with world = 1 (the "window" is then ordinary device memory).
// Scalars live in registers of every CTA; the host launches once per solve and reads the result.
//
// Memory-model notes: vectors that change inside the kernel (u, r, p, s, w, x) are read with plain loads
// (never __ldg / ld.global.nc: the kernel outlives the usual per-launch L1 invalidation); the acquire loads
// that end every barrier invalidate L1, so plain loads afterwards observe what other SMs / GPUs wrote.
#pragma once
// (included by solver.cu inside namespace dkmc, after the peer-window plumbing)

struct PcgSync {
    unsigned int arrive_h, arrive_r;   // arrival counters of the two barrier kinds (reset by the last arriver)
    unsigned long long gen_h;          // local release of barrier H: the halo sequence number reached
    unsigned int n_global, pad;        // clusters that straddle a slab face (length of gl_list; zeroed per launch)
    unsigned long long gen_r;          // local release of barrier R: the reduction sequence number reached
    double sums[2][4];                 // the reduced scalars of reduction rseq, double-buffered by its parity
    double loc[2][4];                  // pipelined PCG: this rank's own sums of reduction rseq (for whoever completes it)
    unsigned int arrive_s, pad2;       // pipelined PCG: arrivals at the end of the SpMV phase (the first completes the reduction)
};

struct PcgArgs {
    int m, ra, rb, t0, t1, n_cl, max_iter;
    const int *row_ptr, *col;
    const double *val, *dinv, *b;
    const int4 *tile_info;
    double *x, *r, *w, *p, *s;           // full-length vectors; only the own rows [ra, rb) are touched
    Precond P;                           // cluster tables (pos may be null: plain Jacobi)
    double *cs, *cr;                     // [2][n_cl] cluster-sum recurrences W^T s, W^T r (replicated)
    double *payload;                     // [4 + 2 n_cl] this rank's cluster sums (entry 4 + s; b-sums at 4 + n_cl + s)
    int *cl_kind;                        // [n_cl] at a cluster's first position: 0 not mine, 1 all members mine, 2 straddles ranks
    int *gl_list;                        // [n_cl] first positions of the straddling clusters (any order)
    int row_end_all[DKMC_MAX_RANKS];     // last row + 1 of every rank (rank of a row)
    double *z, *nvec, *g2;               // pipelined PCG only: z, n = A m, and the second gather buffer (own window + g2_off)
    size_t g2_off;                       // byte offset of the second gather buffer inside every rank's window
    double *partials;                    // [3][gridDim] CTA partial sums
    PcgSync *sync;
    CgScalars *sc;
    double tol;
    unsigned long long rseq0, hseq0;     // sequence numbers already used by earlier solves
    P2pPeers peers;
    P2pHalo halo;
    long long *prof;                     // optional [8]: time spent by CTA 0 per phase (ns)
};

constexpr int kPcgFlagR = 4, kPcgFlagH = 6;   // flag banks of the persistent kernel (0-3: the per-op kernels)

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// polling loads: relaxed (no fence per poll); the waiter fences once after the flag has arrived
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long pcg_now() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Reductions travel in the second bank of slots as self-validating words (the "LL" idea of NCCL): a double goes
// as two 8-byte words {low half | seq << 32}, {high half | seq << 32}, each one store, so a reader that finds the
// sequence number in both words has the value — no flag after the data, hence no system-scope fence at the
// sender and no acquire at the reader (a release.sys store alone cost ~8 us per reduction).
__device__ __forceinline__ unsigned long long *pcg_slot(const P2pPeers &P, int at_rank, int buf, int of_rank) {
    return reinterpret_cast<unsigned long long *>(P.base[at_rank] + P.red2_off) + ((size_t)buf * P.world + of_rank) * kP2pRedCap;
}
constexpr int kPcgMaxPayload = kP2pRedCap / 2;   // doubles per reduction (two words each)

__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void pcg_ll_store(unsigned long long *slot, int j, double v, unsigned int seq) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v), tag = (unsigned long long)seq << 32;
    st_relaxed_sys(slot + 2 * j, tag | (bits & 0xffffffffull));
    st_relaxed_sys(slot + 2 * j + 1, tag | (bits >> 32));
}
// entry j of rank q's contribution to reduction `seq`, as it arrived here (spins until it has)
__device__ __forceinline__ double pcg_ll_load(const P2pPeers &P, int buf, int q, int j, unsigned int seq, int *err) {
    const unsigned long long *w = pcg_slot(P, P.rank, buf, q) + 2 * j;
    unsigned long long a = ld_relaxed_sys(w), b = ld_relaxed_sys(w + 1);
    if ((unsigned int)(a >> 32) != seq || (unsigned int)(b >> 32) != seq) {
        const long long t0 = clock64();
        do {
            a = ld_relaxed_sys(w); b = ld_relaxed_sys(w + 1);
            if (clock64() - t0 > kP2pTimeoutCycles) { *err = 4; break; }
        } while ((unsigned int)(a >> 32) != seq || (unsigned int)(b >> 32) != seq);
    }
    return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
}
// sum over the ranks (in rank order) of entry j of reduction `seq` — one system-scope round trip per rank: only
// for the few clusters that straddle ranks (the scalars of a reduction are polled by one thread per rank)
__device__ __forceinline__ double pcg_reduced(const P2pPeers &P, unsigned long long rseq, int j, int *err) {
    const int buf = (int)(rseq & 1ull);
    double acc = 0.0;
    for (int q = 0; q < P.world; ++q) acc += pcg_ll_load(P, buf, q, j, (unsigned int)rseq, err);
    return acc;
}
// the sum over cluster s0 (first position) of the latest reduction: private clusters from local memory
__device__ __forceinline__ double pcg_cluster_sum(const PcgArgs &a, unsigned long long rseq, int s0, int off, int *err) {
    if (__ldcg(a.cl_kind + s0) == 1) return __ldcg(a.payload + 4 + off + s0);
    return pcg_reduced(a.peers, rseq, 4 + off + s0, err);
}

__device__ __forceinline__ bool pcg_push(const P2pHalo &H, const P2pPeers &P, int i, double v) {
    bool pushed = false;
    for (int sgm = 0; sgm < H.n_send; ++sgm)
        if (i >= H.send_begin[sgm] && i < H.send_end[sgm]) {
            reinterpret_cast<double *>(P.base[H.send_peer[sgm]])[i] = v;
            pushed = true;
        }
    return pushed;
}

// Barrier H: the own rows of u (and the boundary rows pushed to the neighbours) are complete on return,
// and so are the neighbours' boundary rows in this rank's window.
__device__ __forceinline__ void pcg_barrier_halo(const PcgArgs &a, unsigned long long hseq, bool pushed) {
    const P2pPeers &P = a.peers;
    __shared__ bool s_last_h;
    (void)pushed;
    const int buf = kPcgFlagH + (int)(hseq & 1ull);
    __syncthreads();
    if (threadIdx.x == 0) {
        // gpu scope is enough also for the rows this CTA stored into the neighbours' windows: the chain
        // store -> fence.gpu -> arrival (atomic) -> last CTA's fence -> its st.release.sys flag -> the neighbour's
        // ld.acquire.sys is a causality chain in the PTX memory model, and system-wide fences in every CTA are
        // what made these exchanges slow (15-20 us per barrier, measured)
        __threadfence();
        s_last_h = atomicAdd(&a.sync->arrive_h, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last_h) {
        // the local release and the neighbours' flags leave from different warps: a system-scope release costs
        // microseconds, and nothing orders the three stores among themselves
        const int wid = (int)threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0 && wid <= a.halo.n_send && wid < (int)blockDim.x / 32) {
            __threadfence();
            if (wid == 0) {
                a.sync->arrive_h = 0u;
                st_release_gpu(&a.sync->gen_h, hseq);
                for (int sgm = (int)blockDim.x / 32 - 1; sgm < a.halo.n_send; ++sgm)   // more neighbours than warps (never on x-slabs)
                    st_release_sys(p2p_flag(P, a.halo.send_peer[sgm], buf, P.rank), hseq);
            } else {
                st_release_sys(p2p_flag(P, a.halo.send_peer[wid - 1], buf, P.rank), hseq);
            }
        }
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_relaxed_gpu(&a.sync->gen_h) < hseq)
            if (clock64() - t0 > kP2pTimeoutCycles) { a.sc->pad = 2; break; }
        for (int sgm = 0; sgm < a.halo.n_recv; ++sgm) {
            const unsigned long long *f = p2p_flag(P, P.rank, buf, a.halo.recv_peer[sgm]);
            while (ld_relaxed_sys(f) < hseq)
                if (clock64() - t0 > kP2pTimeoutCycles) { a.sc->pad = 3; break; }
        }
        // the acquire that ends the wait (and invalidates L1: plain loads now see what other SMs / GPUs wrote)
        if (a.halo.n_recv > 0) (void)ld_acquire_sys(p2p_flag(P, P.rank, buf, a.halo.recv_peer[0]));
        else (void)ld_acquire_gpu(&a.sync->gen_h);
    }
    __syncthreads();
}

// Barrier R + all-reduce.  On entry every CTA has written its `nparts` partial sums (partials[q * grid + cta])
// and the cluster warps their entries of a.payload[4 ..).  The last CTA to arrive adds the partials in index
// order, sends [sums | entries of the straddling clusters] to every other rank (LL words), collects the other
// ranks' sums (one thread per (entry, rank), all polls in flight together), adds them in rank order — identical
// bits on all ranks — and releases the local barrier with the result.  Everybody else polls ONE local word with
// ONE thread: with every CTA polling the LL words themselves (the first version) thousands of system-scope
// loads per round queued on a single L2 line, and the barrier cost 9 us on one GPU and 16-20 us on four.
// out[0 .. 3] = global sums.  The straddling clusters' entries are read where they are needed
// (pcg_cluster_sum).  with_b: they also send their second array (set-up).  Returns true after a timeout.
__device__ __forceinline__ bool pcg_barrier_reduce(const PcgArgs &a, unsigned long long rseq, int nparts, bool with_b,
                                                   double *out, double *sh) {
    const P2pPeers &P = a.peers;
    __shared__ bool s_last;
    __shared__ double s_loc[4];
    __shared__ int s_err;
    const int buf = (int)(rseq & 1ull);
    const unsigned int seq = (unsigned int)rseq;
    const int n = a.n_cl;
    __syncthreads();
    if (threadIdx.x == 0) {
        s_err = 0;
        __threadfence();
        s_last = (atomicAdd(&a.sync->arrive_r, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        const int G = (int)gridDim.x;
        // the CTA partials in index order; the loads of all arrays in flight together
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < G; i += blockDim.x) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q < nparts) acc[q] += __ldcg(a.partials + (size_t)q * G + i);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (q < nparts) acc[q] = block_sum(acc[q], sh);
            if (threadIdx.x == 0) s_loc[q] = q < nparts ? acc[q] : 0.0;
        }
        if (threadIdx.x == 0) { s_loc[3] = 0.0; a.sync->arrive_r = 0u; }
        __syncthreads();
        if (P.world > 1) {
            // the scalars to the other ranks; the straddling clusters' entries to every rank, this one included
            const int ng = (int)__ldcg(&a.sync->n_global);
            const int per = with_b ? 2 : 1, K = 4 + per * ng;
            for (int q = 0; q < P.world; ++q) {
                unsigned long long *slot = pcg_slot(P, q, buf, P.rank);
                for (int e = threadIdx.x; e < K; e += blockDim.x) {
                    if (e < 4) {
                        if (q != P.rank) pcg_ll_store(slot, e, s_loc[e], seq);
                    } else {
                        const int gi = (e - 4) / per, second = (e - 4) - gi * per;
                        const int j = 4 + second * n + __ldcg(a.gl_list + gi);
                        pcg_ll_store(slot, j, __ldcg(a.payload + j), seq);
                    }
                }
            }
            // the other ranks' scalars: thread (e, q) waits for entry e of rank q; sums in rank order
            double *s_w = sh;   // [8][4] staged values of eight ranks at a time
            for (int q0 = 0; q0 < P.world; q0 += 8) {
                const int e = (int)threadIdx.x & 3, q = q0 + ((int)threadIdx.x >> 2);
                __syncthreads();
                if (threadIdx.x < 32 && q < P.world) {
                    int err = 0;
                    s_w[threadIdx.x] = q == P.rank ? s_loc[e] : pcg_ll_load(P, buf, q, e, seq, &err);
                    if (err) { s_err = err; a.sc->pad = err; }
                }
                __syncthreads();
                if (threadIdx.x < 4) {
                    double acc2 = q0 == 0 ? 0.0 : out[threadIdx.x];
                    for (int u = 0; u < 8 && q0 + u < P.world; ++u) acc2 += s_w[4 * u + threadIdx.x];
                    out[threadIdx.x] = acc2;
                }
            }
            __syncthreads();
        } else {
            if (threadIdx.x < 4) out[threadIdx.x] = s_loc[threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x < 4) a.sync->sums[buf][threadIdx.x] = out[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            st_release_gpu(&a.sync->gen_r, rseq);
        }
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (ld_relaxed_gpu(&a.sync->gen_r) < rseq)
                if (clock64() - t0 > kP2pTimeoutCycles) { s_err = 2; a.sc->pad = 2; break; }
            (void)ld_acquire_gpu(&a.sync->gen_r);
        }
        __syncthreads();
        if (threadIdx.x < 4) out[threadIdx.x] = __ldcg(&a.sync->sums[buf][threadIdx.x]);
    }
    __syncthreads();
    return s_err != 0;
}

// sum over a row of A times g, lanes striding the row (fixed order: deterministic); result in every lane
__device__ __forceinline__ double pcg_warp_row_dot(const PcgArgs &a, const double *g, int row, int lane) {
    double s = 0.0;
    for (int k = __ldg(a.row_ptr + row) + lane, e = __ldg(a.row_ptr + row + 1); k < e; k += 32)
        s += __ldg(a.val + k) * g[__ldg(a.col + k)];
    return warp_sum(s);
}

// experiment switch (never the default): gathers through the non-coherent path
#ifdef DKMC_PCG_GATHER_NC
#define PCG_GATHER(ptr) __ldg(ptr)
#else
#define PCG_GATHER(ptr) (*(ptr))
#endif

// SpMV over the own tiles: y = A g (MODE 1, returns the CTA's share of y.g) or y = b - A g (MODE 2).
// The tile algorithm of spmv_tile_kernel: the tile's val/col streamed coalesced in two batches of four per
// thread, products parked in shared memory, every row added in CSR order.  What bounds it is latency times
// the tiles in flight per SM (measured: a software-pipelined version with all sixteen loads of the next tile
// in flight needed 64 registers, ran 4 CTAs per SM instead of 6 and was slower), so the phase is written to
// live in the 40 registers that keep six CTAs resident.
template <int MODE>
__device__ __forceinline__ double pcg_spmv_tiles(const PcgArgs &a, const double *g, double *y, double *prod, int t_first,
                                                 int t_end, int t_step) {
    double local = 0.0;
    for (int t = t_first; t < t_end; t += t_step) {
        const int4 ti = __ldg(a.tile_info + t);
        const int r0 = ti.x, r1 = ti.y;
        if (r0 < r1) {
            const int k0 = ti.z, k1 = ti.w;
            const int ka = k0 & ~1;
            const int my_r = r0 + (int)threadIdx.x;
            int ra = 0, rb = 0;
            if (my_r < r1) { ra = __ldg(a.row_ptr + my_r); rb = __ldg(a.row_ptr + my_r + 1); }
            if (k1 - ka <= kSpmvCap) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double v[4];
                    int c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + (h * 4 + u) * kSpmvThreads + threadIdx.x;
                        const bool ok = k < k1;
                        v[u] = ok ? __ldcs(a.val + k) : 0.0;
                        c[u] = ok ? __ldcs(a.col + k) : 0;
                    }
                    // ptxas otherwise issues the first gather right after the first column load and sinks the
                    // value loads next to their products — the warp then stalls with two loads in flight.  Make
                    // every gather address depend on ALL eight loads of the batch (columns are non-negative, so
                    // `zero` is 0 — which the assembler cannot know).
                    int zero;
                    {
                        const int c_or = c[0] | c[1] | c[2] | c[3];
                        const int v_or = __double2hiint(v[0]) | __double2hiint(v[1]) | __double2hiint(v[2]) | __double2hiint(v[3]);
                        asm("{ .reg .u32 t; shr.u32 t, %1, 31; and.b32 %0, t, %2; }" : "=r"(zero) : "r"(c_or), "r"(v_or));
                    }
                    double xg[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xg[u] = PCG_GATHER(g + c[u] + zero);   // masked entries read g[0]: harmless
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + (h * 4 + u) * kSpmvThreads + threadIdx.x;
                        if (k < k1) prod[k - ka] = v[u] * xg[u];
                    }
                }
                __syncthreads();
                for (int r = my_r; r < r1; r += kSpmvThreads) {
                    if (r != my_r) { ra = __ldg(a.row_ptr + r); rb = __ldg(a.row_ptr + r + 1); }
                    double s = 0.0;
#pragma unroll 4
                    for (int k = ra - ka; k < rb - ka; ++k) s += prod[k];
                    if (MODE == 2) s = __ldg(a.b + r) - s;
                    y[r] = s;
                    if (MODE == 1) local += g[r] * s;
                }
            } else {  // rows too long for the staging buffer
                for (int r = my_r; r < r1; r += kSpmvThreads) {
                    double s = 0.0;
                    for (int k = __ldg(a.row_ptr + r); k < __ldg(a.row_ptr + r + 1); ++k) s += __ldg(a.val + k) * g[__ldg(a.col + k)];
                    if (MODE == 2) s = __ldg(a.b + r) - s;
                    y[r] = s;
                    if (MODE == 1) local += g[r] * s;
                }
            }
        }
        __syncthreads();   // prod is reused by the next tile
    }
    return local;
}

// cluster rows: payload[4 + off + s] = sum over the members of cluster s this rank owns of f(row), with
// f = (A g)_row (MODE 1) or b_row - (A g)_row (MODE 2, and b_row into the second array).  One warp per cluster.
template <int MODE>
__device__ __forceinline__ void pcg_cluster_rows(const PcgArgs &a, const double *g) {
    const int n = a.n_cl;
    if (n <= 0) return;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    // dealt out one per CTA from the END of the grid (the first CTAs are the ones that hold one more SpMV tile),
    // the next round to the next warp: a CTA's warps work in parallel, so the phase grows by one cluster at most
    const int G = (int)gridDim.x;
    for (int s = (G - 1 - (int)blockIdx.x) + (int)(threadIdx.x >> 5) * G; s < n; s += G * wpb) {
        double acc = 0.0, accb = 0.0;
        if (__ldg(a.P.seg_start + s) != s || __ldcg(a.cl_kind + s) == 0) continue;   // not a first position / no member here
        {
            const int len = __ldg(a.P.seg_len + s);
            for (int k = 0; k < len; ++k) {
                const int row = __ldg(a.P.mem_row + s + k);
                if (row < a.ra || row >= a.rb) continue;
                const double d = pcg_warp_row_dot(a, g, row, lane);
                if (MODE == 2) { const double bv = __ldg(a.b + row); acc += bv - d; accb += bv; }
                else acc += d;
            }
        }
        if (lane == 0) {
            a.payload[4 + s] = acc;
            if (MODE == 2) a.payload[4 + n + s] = accb;
        }
    }
}

// MINB = CTAs per SM the register budget allows: 6 (40 registers) as the stand-alone SpMV kernel, 5 (48), 4 (64).
// The scalars of the recurrence live in SHARED memory between the phases (every CTA keeps its own identical
// copy), not in registers: a phase then needs no more registers than the same code as a kernel of its own.
struct PcgState {
    double alpha, beta, gamma, stop, bb, gmin;
    int done, err, it_min;
};

template <int MINB, bool PROF>
__global__ void __launch_bounds__(kSpmvThreads, MINB) pcg_persistent_kernel(const PcgArgs a) {
    __shared__ __align__(16) double prod[kSpmvCap];
    __shared__ double red[32];
    __shared__ double s_out[4];
    __shared__ PcgState st;
    const P2pPeers &P = a.peers;
    const int G = (int)gridDim.x, B = (int)blockDim.x, cta = (int)blockIdx.x, tid = (int)threadIdx.x;
    const int n = a.n_cl;
    double *g = reinterpret_cast<double *>(P.base[P.rank]);   // u: own rows + the neighbours' boundary rows
    const bool clustered = a.P.pos != nullptr && n > 0;
    // vector phases: a contiguous chunk of the own rows per CTA, so that only the CTAs at the slab faces store
    // into the neighbours' windows
    const int chunk = (a.rb - a.ra + G - 1) / G;
    const int v0 = a.ra + cta * chunk, v1 = min(a.rb, v0 + chunk);
    long long tp = 0, prof[6] = {0, 0, 0, 0, 0, 0}, cp_s = 0, cp_w = 0, cp_t = 0;
    const bool do_prof = PROF && cta == 0 && tid == 0;
    const bool cta_prof = PROF && tid == 0;   // every CTA: its SpMV-phase time and its wait at barrier R
    if (do_prof) tp = pcg_now();
#define PCG_PROF(slot) do { if (PROF && do_prof) { const long long now__ = pcg_now(); prof[slot] += now__ - tp; tp = now__; } } while (0)

    // sequence numbers: set-up uses halo exchanges hseq0+1, +2 and reductions rseq0+1, +2; iteration `it` (from 0)
    // uses hseq0+3+it and rseq0+3+it
    // ---- init 0: u := x over the own rows (the SpMV gathers from the window), boundary rows to the neighbours
    {
        bool pushed = false;
        for (int i = v0 + tid; i < v1; i += B) {
            const double v = a.x[i];
            g[i] = v;
            pushed |= pcg_push(a.halo, P, i, v);
        }
        // clusters: whose business are they?  Members are sorted by row inside a cluster, so the ranks of the
        // first and the last member decide: both mine -> private (1); same other rank -> not mine (0); else the
        // cluster straddles a slab face (2) and its sums travel in the reductions (same verdict on every rank)
        for (int s = cta * B + tid; s < n; s += G * B)
            if (__ldg(a.P.seg_start + s) == s) {
                const int first = __ldg(a.P.mem_row + s), last = __ldg(a.P.mem_row + s + __ldg(a.P.seg_len + s) - 1);
                int qf = 0, ql = 0;
                while (qf < P.world - 1 && first >= a.row_end_all[qf]) ++qf;
                while (ql < P.world - 1 && last >= a.row_end_all[ql]) ++ql;
                const int kind = qf != ql ? 2 : (qf == P.rank ? 1 : 0);
                a.cl_kind[s] = kind;
                if (kind == 2) a.gl_list[atomicAdd(&a.sync->n_global, 1u)] = s;
            }
        pcg_barrier_halo(a, a.hseq0 + 1, pushed);
    }
    // ---- init 1: r = b - A x; cluster sums of r and of b
    pcg_spmv_tiles<2>(a, g, a.r, prod, a.t0 + cta, a.t1, G);
    if (clustered) pcg_cluster_rows<2>(a, g);
    pcg_barrier_reduce(a, a.rseq0 + 1, 0, true, s_out, red);
    // ---- init 2: u = M^-1 r (into the window), gamma and b.M^-1 b partials, recurrence state
    {
        const unsigned long long rs_ = a.rseq0 + 1;
        int lerr = 0;
        double lg = 0.0, lbb = 0.0;
        bool pushed = false;
        for (int i = v0 + tid; i < v1; i += B) {
            const double ri = a.r[i], bi = __ldg(a.b + i), di = __ldg(a.dinv + i);
            double un = ri * di, zb = bi * di;
            const int sp = clustered ? __ldg(a.P.pos + i) : -1;
            if (sp >= 0) {
                const int s0 = __ldg(a.P.seg_start + sp);
                const double we = __ldg(a.P.w + s0);
                un += we * pcg_cluster_sum(a, rs_, s0, 0, &lerr);
                zb += we * pcg_cluster_sum(a, rs_, s0, n, &lerr);
            }
            g[i] = un;
            pushed |= pcg_push(a.halo, P, i, un);
            lg += ri * un;
            lbb += bi * zb;
        }
        for (int s = cta * B + tid; s < n; s += G * B)
            if (__ldg(a.P.seg_start + s) == s && __ldcg(a.cl_kind + s) != 0) { a.cr[s] = pcg_cluster_sum(a, rs_, s, 0, &lerr); a.cs[s] = 0.0; }
        if (lerr) a.sc->pad = lerr;
        lg = block_sum(lg, red);
        __syncthreads();
        lbb = block_sum(lbb, red);
        if (tid == 0) { a.partials[cta] = lg; a.partials[2 * (size_t)G + cta] = lbb; }
        pcg_barrier_halo(a, a.hseq0 + 2, pushed);
    }
    // ---- init 3: w = A u; delta partial; cluster sums of w
    {
        double ld = pcg_spmv_tiles<1>(a, g, a.w, prod, a.t0 + cta, a.t1, G);
        if (clustered) pcg_cluster_rows<1>(a, g);
        ld = block_sum(ld, red);
        if (tid == 0) a.partials[(size_t)G + cta] = ld;
        const bool err = pcg_barrier_reduce(a, a.rseq0 + 2, 3, false, s_out, red);
        if (tid == 0) {
            const double gamma = s_out[0], delta = s_out[1], bb = s_out[2];
            st.gamma = gamma; st.bb = bb;
            st.stop = a.tol * a.tol * (bb > 0.0 ? bb : gamma);
            st.alpha = gamma / delta; st.beta = 0.0;
            st.done = (gamma <= st.stop || !(gamma == gamma)) ? 1 : 0;
            st.err = err ? 1 : 0;
        }
        __syncthreads();
    }
    PCG_PROF(0);

    int it = 0;
    for (; it < a.max_iter && !st.done && !st.err; ++it) {
        // ---- V: p, s, x, r, u over the own rows; gamma partial
        {
            const double alpha = st.alpha, beta = st.beta;
            const unsigned long long rs_ = a.rseq0 + 2 + it;   // the latest reduction
            int lerr = 0;
            const int par = it & 1;
            const double *cs_old = a.cs + (size_t)par * n, *cr_old = a.cr + (size_t)par * n;
            double *cs_new = a.cs + (size_t)(par ^ 1) * n, *cr_new = a.cr + (size_t)(par ^ 1) * n;
            const bool first = it == 0;
            double lg = 0.0;
            bool pushed = false;
            for (int i = v0 + tid; i < v1; i += B) {
                // every load before the first store: the vectors may alias as far as the compiler knows, and a
                // load behind a store would wait for it (three dependent round trips per pass instead of one)
                const double ui = g[i], wi = a.w[i], xi = a.x[i], ro = a.r[i], di = __ldg(a.dinv + i);
                const int sp = clustered ? __ldg(a.P.pos + i) : -1;
                double pi = ui, si = wi;
                if (!first) { pi += beta * a.p[i]; si += beta * a.s[i]; }
                const double ri = ro - alpha * si;
                a.p[i] = pi;
                a.s[i] = si;
                a.x[i] = xi + alpha * pi;
                a.r[i] = ri;
                double un = ri * di;
                if (sp >= 0) {
                    const int s0 = __ldg(a.P.seg_start + sp);
                    const double csn = pcg_cluster_sum(a, rs_, s0, 0, &lerr) + beta * cs_old[s0];
                    un += __ldg(a.P.w + s0) * (cr_old[s0] - alpha * csn);
                }
                g[i] = un;
                pushed |= pcg_push(a.halo, P, i, un);
                lg += ri * un;
            }
            for (int s = cta * B + tid; s < n; s += G * B)
                if (__ldg(a.P.seg_start + s) == s && __ldcg(a.cl_kind + s) != 0) {
                    const double csn = pcg_cluster_sum(a, rs_, s, 0, &lerr) + beta * cs_old[s];
                    cs_new[s] = csn;
                    cr_new[s] = cr_old[s] - alpha * csn;
                }
            if (lerr) a.sc->pad = lerr;
            lg = block_sum(lg, red);
            if (tid == 0) a.partials[cta] = lg;
            PCG_PROF(1);
            pcg_barrier_halo(a, a.hseq0 + 3 + it, pushed);
            PCG_PROF(2);
        }
        // ---- S: w = A u; delta partial; cluster sums of w; the one reduction of the iteration
        {
            if (PROF && cta_prof) cp_t = pcg_now();
            double ld = pcg_spmv_tiles<1>(a, g, a.w, prod, a.t0 + cta, a.t1, G);
            PCG_PROF(3);
            if (clustered) pcg_cluster_rows<1>(a, g);
            ld = block_sum(ld, red);
            if (tid == 0) a.partials[(size_t)G + cta] = ld;
            PCG_PROF(4);
            if (PROF && cta_prof) { const long long now = pcg_now(); cp_s += now - cp_t; cp_t = now; }
            const bool err = pcg_barrier_reduce(a, a.rseq0 + 3 + it, 2, false, s_out, red);
            if (PROF && cta_prof) cp_w += pcg_now() - cp_t;
            PCG_PROF(5);
            if (tid == 0) {
                const double gamma_new = s_out[0], delta = s_out[1];
                const double beta = gamma_new / st.gamma;
                st.alpha = gamma_new / (delta - beta * gamma_new / st.alpha);
                st.beta = beta;
                st.gamma = gamma_new;
                st.done = (gamma_new <= st.stop || !(gamma_new == gamma_new)) ? 1 : 0;
                st.err = err ? 1 : 0;
            }
            __syncthreads();
        }
    }
    if (PROF && cta_prof) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.prof[8 + 2 * cta] += cp_s;
        a.prof[9 + 2 * cta] = (a.prof[9 + 2 * cta] & ~0xfffll) + (cp_w << 12) + (long long)(smid & 0xfffu);   // wait | SM id
    }
    if (cta == 0 && tid == 0) {
        CgScalars *sc = a.sc;
        sc->rz = st.gamma; sc->bb = st.bb; sc->stop = st.stop; sc->iters = it; sc->max_iter = a.max_iter;
        sc->done = st.done;
        sc->alpha = st.alpha; sc->beta = st.beta;
        sc->rseq_end = a.rseq0 + 2 + it; sc->hseq_end = a.hseq0 + 2 + it;
        if (PROF) {
            for (int q = 0; q < 6; ++q) a.prof[q] += prof[q];
            a.prof[6] += it;
            a.prof[7] += 1;
        }
    }
#undef PCG_PROF
}
