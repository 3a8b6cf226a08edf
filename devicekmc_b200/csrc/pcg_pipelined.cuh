// devicekmc-b200 — the preconditioned CG of a5 as ONE persistent kernel with ONE synchronisation point per
// iteration: pipelined CG (Ghysels & Vanroose 2014) — algebraically the CG of iterative_solvers_gpu.cu:424-455,
// with w = A u carried by a recurrence so that both inner products of an iteration are known BEFORE its matrix
// product and their reduction travels while the product runs:
//     gamma = (r, u), delta = (w, u)                     <- reduction started ...
//     m = M^-1 w ; n = A m                               <- ... and hidden behind the halo exchange and the SpMV
//     beta = gamma / gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old)
//     z = n + beta z ; s = w + beta s ; p = u + beta p ; x += alpha p ; r -= alpha s ; w -= alpha z
// u = M^-1 r and q = M^-1 s are not stored: M^-1 = D^-1 + W E^-1 W^T is cheap enough to apply where it is needed
// (one multiply and, for the few clustered rows, one cluster term), which also removes two drifting recurrences.
//
// The Chronopoulos-Gear kernel (pcg_persistent.cuh) needs two grid-wide synchronisations per iteration — the halo
// of u before the product, the reduction after it — and on several GPUs each of them costs 10-20 us of NVLink
// round trips.  Here an iteration is
//   V  vector phase over the own rows; partial sums; m into the OTHER gather buffer, boundary rows pushed to the
//      neighbours.  ONE arrival: the last CTA adds the partial sums, sends them to the other ranks (LL words),
//      raises the halo flags and releases the local barrier; everybody waits for the release and the halo only.
//   S  n = A m over the own tiles; the first CTA to finish collects the other ranks' sums (long arrived) and
//      publishes the result, the others read it when they get there.
// "Owner computes": a CTA owns a contiguous range of nnz tiles and, in the vector phase, exactly the rows of those
// tiles, so n needs no barrier between S and V; the gather vector is double-buffered, so a CTA may run ahead into
// the next vector phase while others still gather.
// Cluster sums W^T n: every CTA keeps, in shared memory, the recurrences of the clusters that have a member among
// its rows and computes W^T n for them itself from the complete vector m (redundantly, identical bits).  Clusters
// that straddle a slab face exchange per-rank partial sums point to point (LL words, sent before the tiles, read
// after them).
// Attainable accuracy: the true residual of a pipelined solve stagnates two orders above the classic recurrences'
// (measured on the 100 k-site matrix: 7e-11 instead of 9e-13 relative) — irrelevant here, because the driver
// restarts on the double-double residual anyway (solve_refined) and the total iteration count is the same.
#pragma once
// (included by solver.cu inside namespace dkmc, after pcg_persistent.cuh)

constexpr int kPipeCap = 24;   // clusters per CTA kept in shared memory (more: the host falls back to the other kernel)

struct PipeTab {
    int n;
    int s0[kPipeCap];                    // first position of the cluster in the sorted member list
    int kind[kPipeCap];                  // 1 all members on this rank, 2 straddles ranks qf .. ql
    int qf[kPipeCap], ql[kPipeCap];
    int sender[kPipeCap];                // this CTA holds this rank's first member: it sends the rank's partial sums
    double w[kPipeCap];                  // 1 / (1_c^T A 1_c)
    double cn[kPipeCap], cb[kPipeCap];   // latest exchanged sums (W^T n; set-up: W^T r and W^T b)
    double cz[kPipeCap], cs[kPipeCap], cw[kPipeCap], cr[kPipeCap];   // recurrences W^T z, W^T s, W^T w, W^T r
};

__device__ __forceinline__ int pipe_rank_of(const PcgArgs &a, int row) {
    int q = 0;
    while (q < a.peers.world - 1 && row >= a.row_end_all[q]) ++q;
    return q;
}

__device__ __forceinline__ int pipe_find(const PipeTab &t, int s0) {
    for (int k = 0; k < t.n; ++k)
        if (t.s0[k] == s0) return k;
    return -1;
}

// Cluster sums of the table, part A (before the tiles): every entry's warp adds f(row) over the members —
// all of them for a private cluster, this rank's for a straddling one — with f = (A g)_row (MODE 1) or
// b_row - (A g)_row and b_row (MODE 2).  Private clusters are done; the sender of a straddling cluster ships the
// partial sums to every participant rank.
template <int MODE>
__device__ __forceinline__ void pipe_cluster_partials(const PcgArgs &a, const double *g, PipeTab &t, unsigned long long cseq) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const P2pPeers &P = a.peers;
    for (int k = wid; k < t.n; k += nw) {
        const int s0 = t.s0[k], len = __ldg(a.P.seg_len + s0), kind = t.kind[k];
        if (kind == 2 && !t.sender[k]) continue;
        double acc = 0.0, accb = 0.0;
        for (int j = 0; j < len; ++j) {
            const int row = __ldg(a.P.mem_row + s0 + j);
            if (kind == 2 && (row < a.ra || row >= a.rb)) continue;
            const double d = pcg_warp_row_dot(a, g, row, lane);
            if (MODE == 2) { const double bv = __ldg(a.b + row); acc += bv - d; accb += bv; }
            else acc += d;
        }
        if (lane == 0) {
            if (kind == 1) { t.cn[k] = acc; t.cb[k] = accb; }
            else {
                const int buf = (int)(cseq & 1ull);
                for (int q = t.qf[k]; q <= t.ql[k]; ++q) {
                    unsigned long long *slot = pcg_slot(P, q, buf, P.rank);
                    pcg_ll_store(slot, 4 + s0, acc, (unsigned int)cseq);
                    if (MODE == 2) pcg_ll_store(slot, 4 + a.n_cl + s0, accb, (unsigned int)cseq);
                }
            }
        }
    }
}
// part B (after the tiles): the holders of a straddling cluster add the participants' partial sums in rank order
template <int MODE>
__device__ __forceinline__ void pipe_cluster_collect(const PcgArgs &a, PipeTab &t, unsigned long long cseq, int *err) {
    const P2pPeers &P = a.peers;
    const int buf = (int)(cseq & 1ull);
    for (int k = threadIdx.x; k < t.n; k += blockDim.x)
        if (t.kind[k] == 2) {
            double acc = 0.0, accb = 0.0;
            for (int q = t.qf[k]; q <= t.ql[k]; ++q) {
                acc += pcg_ll_load(P, buf, q, 4 + t.s0[k], (unsigned int)cseq, err);
                if (MODE == 2) accb += pcg_ll_load(P, buf, q, 4 + a.n_cl + t.s0[k], (unsigned int)cseq, err);
            }
            if (*err) a.sc->pad = 4;
            t.cn[k] = acc; t.cb[k] = accb;
        }
}

// The one synchronisation of an iteration.  Every CTA has written its `nparts` partial sums.  The last CTA to
// arrive adds them in index order, sends them to the other ranks (reduction `rs`), keeps a copy for whoever
// completes the reduction on this rank, raises the halo flags at the neighbours and releases the local barrier.
// On return the gather buffer written in this phase is complete on this GPU, and so are the neighbours'
// boundary rows in it.  nparts = 0: no reduction rides on this barrier.
__device__ __forceinline__ void pipe_barrier(const PcgArgs &a, unsigned long long hseq, unsigned long long rs, int nparts,
                                             double *sh) {
    const P2pPeers &P = a.peers;
    __shared__ bool s_last_p;
    __shared__ double s_locp[4];
    const int hbuf = kPcgFlagH + (int)(hseq & 1ull);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last_p = atomicAdd(&a.sync->arrive_h, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last_p) {
        if (nparts > 0) {
            __threadfence();
            const int G = (int)gridDim.x, rbuf = (int)(rs & 1ull);
            double acc[3] = {0.0, 0.0, 0.0};
            for (int i = threadIdx.x; i < G; i += blockDim.x) {
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (q < nparts) acc[q] += __ldcg(a.partials + (size_t)q * G + i);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (q < nparts) acc[q] = block_sum(acc[q], sh);
                if (threadIdx.x == 0) s_locp[q] = q < nparts ? acc[q] : 0.0;
            }
            if (threadIdx.x == 0) s_locp[3] = 0.0;
            __syncthreads();
            // thread (e, q): entry e to rank q
            for (int i = threadIdx.x; i < 4 * P.world; i += blockDim.x) {
                const int e = i & 3, q = i >> 2;
                if (q != P.rank) pcg_ll_store(pcg_slot(P, q, rbuf, P.rank), e, s_locp[e], (unsigned int)rs);
            }
            if (threadIdx.x < 4) {
                a.sync->loc[rbuf][threadIdx.x] = s_locp[threadIdx.x];
                if (P.world == 1) a.sync->sums[rbuf][threadIdx.x] = s_locp[threadIdx.x];
            }
            __syncthreads();
        }
        const int wid = (int)threadIdx.x >> 5, nw = (int)blockDim.x >> 5;
        if ((threadIdx.x & 31) == 0 && wid <= a.halo.n_send) {
            __threadfence();
            if (wid == 0) {
                a.sync->arrive_h = 0u;
                if (nparts > 0 && P.world == 1) st_release_gpu(&a.sync->gen_r, rs);   // one GPU: the reduction is complete
                st_release_gpu(&a.sync->gen_h, hseq);
                for (int sgm = nw - 1; sgm < a.halo.n_send; ++sgm)
                    st_release_sys(p2p_flag(P, a.halo.send_peer[sgm], hbuf, P.rank), hseq);
            } else {
                st_release_sys(p2p_flag(P, a.halo.send_peer[wid - 1], hbuf, P.rank), hseq);
            }
        }
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_relaxed_gpu(&a.sync->gen_h) < hseq)
            if (clock64() - t0 > kP2pTimeoutCycles) { a.sc->pad = 2; break; }
        for (int sgm = 0; sgm < a.halo.n_recv; ++sgm) {
            const unsigned long long *f = p2p_flag(P, P.rank, hbuf, a.halo.recv_peer[sgm]);
            while (ld_relaxed_sys(f) < hseq)
                if (clock64() - t0 > kP2pTimeoutCycles) { a.sc->pad = 3; break; }
        }
        if (a.halo.n_recv > 0) (void)ld_acquire_sys(p2p_flag(P, P.rank, hbuf, a.halo.recv_peer[0]));
        else (void)ld_acquire_gpu(&a.sync->gen_h);
    }
    __syncthreads();
}

// The result of reduction `rs` (started at the previous pipe_barrier): out[0 .. 3].  On several GPUs the first CTA
// to get here adds the ranks' sums in rank order and publishes them; everybody else polls one local word.
__device__ __forceinline__ bool pipe_reduced(const PcgArgs &a, unsigned long long rs, double *out, double *sh) {
    const P2pPeers &P = a.peers;
    __shared__ bool s_first;
    __shared__ int s_errp;
    const int rbuf = (int)(rs & 1ull);
    __syncthreads();
    if (threadIdx.x == 0) {
        s_errp = 0;
        s_first = false;
        if (P.world > 1) {
            const unsigned int t = atomicAdd(&a.sync->arrive_s, 1u);
            s_first = t == 0u;
            if (t == gridDim.x - 1) a.sync->arrive_s = 0u;   // nobody arrives again before the next pipe_barrier has passed
        }
    }
    __syncthreads();
    if (s_first) {
        double *s_w = sh;
        for (int q0 = 0; q0 < P.world; q0 += 8) {
            const int e = (int)threadIdx.x & 3, q = q0 + ((int)threadIdx.x >> 2);
            __syncthreads();
            if (threadIdx.x < 32 && q < P.world) {
                int err = 0;
                s_w[threadIdx.x] = q == P.rank ? __ldcg(&a.sync->loc[rbuf][e]) : pcg_ll_load(P, rbuf, q, e, (unsigned int)rs, &err);
                if (err) { s_errp = err; a.sc->pad = err; }
            }
            __syncthreads();
            if (threadIdx.x < 4) {
                double acc2 = q0 == 0 ? 0.0 : out[threadIdx.x];
                for (int u = 0; u < 8 && q0 + u < P.world; ++u) acc2 += s_w[4 * u + threadIdx.x];
                out[threadIdx.x] = acc2;
            }
        }
        __syncthreads();
        if (threadIdx.x < 4) a.sync->sums[rbuf][threadIdx.x] = out[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            st_release_gpu(&a.sync->gen_r, rs);
        }
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (ld_relaxed_gpu(&a.sync->gen_r) < rs)
                if (clock64() - t0 > kP2pTimeoutCycles) { s_errp = 2; a.sc->pad = 2; break; }
            (void)ld_acquire_gpu(&a.sync->gen_r);
        }
        __syncthreads();
        if (threadIdx.x < 4) out[threadIdx.x] = __ldcg(&a.sync->sums[rbuf][threadIdx.x]);
    }
    __syncthreads();
    return s_errp != 0;
}

template <int MINB, bool PROF>
__global__ void __launch_bounds__(kSpmvThreads, MINB) pcg_pipelined_kernel(const PcgArgs a) {
    __shared__ __align__(16) double prod[kSpmvCap];
    __shared__ double red[32];
    __shared__ double s_out[4];
    __shared__ PcgState st;
    __shared__ PipeTab tab;
    const P2pPeers &P = a.peers;
    const int G = (int)gridDim.x, B = (int)blockDim.x, cta = (int)blockIdx.x, tid = (int)threadIdx.x;
    const int n = a.n_cl;
    const bool clustered = a.P.pos != nullptr && n > 0;
    // the two gather buffers: this rank's window at offset 0 and at g2_off
    double *gbuf[2] = {reinterpret_cast<double *>(P.base[P.rank]), reinterpret_cast<double *>(P.base[P.rank] + a.g2_off)};
    // owner computes: a contiguous range of tiles and exactly their rows
    const int nt = a.t1 - a.t0;
    const int tb = a.t0 + (int)(((long long)cta * nt) / G), te = a.t0 + (int)(((long long)(cta + 1) * nt) / G);
    int v0 = a.rb, v1 = a.rb;
    if (tb < te) { v0 = __ldg(a.tile_info + tb).x; v1 = __ldg(a.tile_info + te - 1).y; }
    long long tp = 0, prof[6] = {0, 0, 0, 0, 0, 0};
    const bool do_prof = PROF && cta == 0 && tid == 0;
    if (do_prof) tp = pcg_now();
#define PIPE_PROF(slot) do { if (PROF && do_prof) { const long long now__ = pcg_now(); prof[slot] += now__ - tp; tp = now__; } } while (0)
    auto push = [&](int which, int i, double v) {
        for (int sgm = 0; sgm < a.halo.n_send; ++sgm)
            if (i >= a.halo.send_begin[sgm] && i < a.halo.send_end[sgm])
                reinterpret_cast<double *>(P.base[a.halo.send_peer[sgm]] + (which ? a.g2_off : 0))[i] = v;
    };
    // sequence numbers: halo exchanges hseq0+1 .. +3 in the set-up, hseq0+4+it in iteration it; reductions
    // rseq0+1 (table overflow anywhere?), rseq0+2 (set-up), rseq0+3+it (started in iteration it); cluster exchanges
    // rseq0+1, +2, +3 in the set-up, rseq0+4+it in iteration it
    if (tid == 0) { tab.n = 0; st.err = 0; st.done = 0; }
    __syncthreads();
    // ---- set-up 0: gather buffer 0 := x; the table of the clusters that have a member among the own rows
    for (int i = v0 + tid; i < v1; i += B) {
        const double v = a.x[i];
        gbuf[0][i] = v;
        push(0, i, v);
        const int sp = clustered ? __ldg(a.P.pos + i) : -1;
        if (sp >= 0) {
            const int s0 = __ldg(a.P.seg_start + sp);
            // the member of this cluster with the smallest row among the own rows registers it
            if (sp == s0 || __ldg(a.P.mem_row + sp - 1) < v0) {
                const int k = atomicAdd(&tab.n, 1);
                if (k < kPipeCap) {
                    const int first = __ldg(a.P.mem_row + s0), last = __ldg(a.P.mem_row + s0 + __ldg(a.P.seg_len + s0) - 1);
                    const int qf = pipe_rank_of(a, first), ql = pipe_rank_of(a, last);
                    tab.s0[k] = s0; tab.qf[k] = qf; tab.ql[k] = ql;
                    tab.kind[k] = qf == ql ? 1 : 2;
                    tab.sender[k] = (sp == s0 || __ldg(a.P.mem_row + sp - 1) < a.ra) ? 1 : 0;
                    tab.w[k] = __ldg(a.P.w + s0);
                    tab.cz[k] = 0.0; tab.cs[k] = 0.0;
                }
            }
        }
    }
    __syncthreads();
    // too many clusters among one CTA's rows, on any rank: everybody leaves and the host takes the other kernel
    if (tid == 0) { a.partials[cta] = tab.n > kPipeCap ? 1.0 : 0.0; if (tab.n > kPipeCap) tab.n = 0; }
    pipe_barrier(a, a.hseq0 + 1, a.rseq0 + 1, 1, red);
    pipe_reduced(a, a.rseq0 + 1, s_out, red);
    if (s_out[0] != 0.0) {
        if (cta == 0 && tid == 0) { a.sc->pad = 5; a.sc->rseq_end = a.rseq0 + 1; a.sc->hseq_end = a.hseq0 + 1; }
        return;
    }
    __syncthreads();
    // ---- set-up 1: r = b - A x over the own tiles; W^T r and W^T b for the table
    int lerr = 0;
    if (clustered) pipe_cluster_partials<2>(a, gbuf[0], tab, a.rseq0 + 1);
    pcg_spmv_tiles<2>(a, gbuf[0], a.r, prod, tb, te, 1);
    if (clustered) pipe_cluster_collect<2>(a, tab, a.rseq0 + 1, &lerr);
    __syncthreads();
    if (tid < tab.n) tab.cr[tid] = tab.cn[tid];
    __syncthreads();
    // ---- set-up 2: u = M^-1 r into gather buffer 1; gamma = r.u and b.M^-1 b partial sums
    {
        double lg = 0.0, lbb = 0.0;
        for (int i = v0 + tid; i < v1; i += B) {
            const double ri = a.r[i], bi = __ldg(a.b + i), di = __ldg(a.dinv + i);
            double un = ri * di, zb = bi * di;
            const int sp = clustered ? __ldg(a.P.pos + i) : -1;
            if (sp >= 0) {
                const int k = pipe_find(tab, __ldg(a.P.seg_start + sp));
                un += tab.w[k] * tab.cr[k];
                zb += tab.w[k] * tab.cb[k];
            }
            gbuf[1][i] = un;
            push(1, i, un);
            lg += ri * un;
            lbb += bi * zb;
        }
        lg = block_sum(lg, red);
        __syncthreads();
        lbb = block_sum(lbb, red);
        if (tid == 0) { a.partials[cta] = lg; a.partials[2 * (size_t)G + cta] = lbb; }
        pipe_barrier(a, a.hseq0 + 2, 0, 0, red);
    }
    // ---- set-up 3: w = A u over the own tiles; W^T w; delta = w.u; m = M^-1 w into gather buffer 0
    {
        if (clustered) pipe_cluster_partials<1>(a, gbuf[1], tab, a.rseq0 + 2);
        pcg_spmv_tiles<1>(a, gbuf[1], a.w, prod, tb, te, 1);
        if (clustered) pipe_cluster_collect<1>(a, tab, a.rseq0 + 2, &lerr);
        __syncthreads();
        if (tid < tab.n) tab.cw[tid] = tab.cn[tid];
        __syncthreads();
        double ld = 0.0;
        for (int i = v0 + tid; i < v1; i += B) {
            const double wi = a.w[i], di = __ldg(a.dinv + i);
            double mi = wi * di;
            const int sp = clustered ? __ldg(a.P.pos + i) : -1;
            if (sp >= 0) { const int k = pipe_find(tab, __ldg(a.P.seg_start + sp)); mi += tab.w[k] * tab.cw[k]; }
            ld += wi * gbuf[1][i];
            gbuf[0][i] = mi;
            push(0, i, mi);
        }
        ld = block_sum(ld, red);
        if (tid == 0) a.partials[(size_t)G + cta] = ld;
        pipe_barrier(a, a.hseq0 + 3, a.rseq0 + 2, 3, red);   // reduction 2: gamma, delta, b.M^-1 b
    }
    // ---- set-up 4 (= the S phase before the first iteration): n = A m; W^T n
    if (clustered) pipe_cluster_partials<1>(a, gbuf[0], tab, a.rseq0 + 3);
    pcg_spmv_tiles<1>(a, gbuf[0], a.nvec, prod, tb, te, 1);
    if (clustered) pipe_cluster_collect<1>(a, tab, a.rseq0 + 3, &lerr);
    PIPE_PROF(0);

    int it = 0;
    for (;; ++it) {
        // ---- the scalars of this iteration: reduction rseq0 + 2 + it was started before the product just finished
        {
            const bool err = pipe_reduced(a, a.rseq0 + 2 + it, s_out, red);
            PIPE_PROF(5);
            if (tid == 0) {
                const double gamma = s_out[0], delta = s_out[1];
                if (it == 0) {
                    st.bb = s_out[2];
                    st.stop = a.tol * a.tol * (st.bb > 0.0 ? st.bb : gamma);
                    st.beta = 0.0;
                    st.alpha = gamma / delta;
                } else {
                    const double beta = gamma / st.gamma;
                    st.alpha = gamma / (delta - beta * gamma / st.alpha);
                    st.beta = beta;
                }
                st.gamma = gamma;
                // the recurrence residual of a pipelined CG stops falling where rounding in the extra recurrences
                // takes over (two orders above the classic CG's floor): leave when it has not set a new minimum
                // for 64 iterations — the host then lets the other kernel carry on from the current x
                if (it == 0 || gamma < st.gmin) { st.gmin = gamma; st.it_min = it; }
                st.done = (gamma <= st.stop || !(gamma == gamma) || it >= a.max_iter || it - st.it_min >= 64) ? 1 : 0;
                st.err = (err || lerr) ? 1 : 0;
            }
            __syncthreads();
            if (st.done || st.err) break;
        }
        const double alpha = st.alpha, beta = st.beta;
        // cluster recurrences: W^T z, W^T s (from the old W^T w), W^T r, W^T w
        if (tid < tab.n) {
            const double czn = tab.cn[tid] + beta * tab.cz[tid], csn = tab.cw[tid] + beta * tab.cs[tid];
            tab.cz[tid] = czn; tab.cs[tid] = csn;
            tab.cb[tid] = tab.cr[tid];                 // W^T r before the update (for p = u + beta p)
            tab.cr[tid] -= alpha * csn;
            tab.cw[tid] -= alpha * czn;
        }
        __syncthreads();
        // ---- V: z, s, p, x, r, w over the own rows; gamma and delta partial sums; m into the other gather buffer
        {
            const int wb = (it + 1) & 1;          // S of this iteration reads gbuf[wb]; the product above read gbuf[it & 1]
            double *gw = wb ? gbuf[1] : gbuf[0];
            const bool first = it == 0;
            double lg = 0.0, ld = 0.0;
            for (int i = v0 + tid; i < v1; i += B) {
                const double ni = a.nvec[i], wi = a.w[i], ri = a.r[i], xi = a.x[i], di = __ldg(a.dinv + i);
                const int sp = clustered ? __ldg(a.P.pos + i) : -1;
                double zi = ni, si = wi, ui = ri * di, pi;
                double c_r_old = 0.0, c_r = 0.0, c_w = 0.0;
                if (sp >= 0) {
                    const int k = pipe_find(tab, __ldg(a.P.seg_start + sp));
                    const double wk = tab.w[k];
                    c_r_old = wk * tab.cb[k]; c_r = wk * tab.cr[k]; c_w = wk * tab.cw[k];
                }
                ui += c_r_old;
                pi = ui;
                if (!first) { zi += beta * a.z[i]; si += beta * a.s[i]; pi += beta * a.p[i]; }
                const double rn = ri - alpha * si, wn = wi - alpha * zi;
                a.z[i] = zi; a.s[i] = si; a.p[i] = pi;
                a.x[i] = xi + alpha * pi;
                a.r[i] = rn; a.w[i] = wn;
                const double un = rn * di + c_r;
                const double mn = wn * di + c_w;
                gw[i] = mn;
                push(wb, i, mn);
                lg += rn * un;
                ld += wn * un;
            }
            lg = block_sum(lg, red);
            __syncthreads();
            ld = block_sum(ld, red);
            if (tid == 0) { a.partials[cta] = lg; a.partials[(size_t)G + cta] = ld; }
            PIPE_PROF(1);
            pipe_barrier(a, a.hseq0 + 4 + it, a.rseq0 + 3 + it, 2, red);
            PIPE_PROF(2);
            // ---- S: n = A m over the own tiles; W^T n for the table
            if (clustered) pipe_cluster_partials<1>(a, gw, tab, a.rseq0 + 4 + it);
            PIPE_PROF(4);
            pcg_spmv_tiles<1>(a, gw, a.nvec, prod, tb, te, 1);
            PIPE_PROF(3);
            if (clustered) pipe_cluster_collect<1>(a, tab, a.rseq0 + 4 + it, &lerr);
            PIPE_PROF(4);
        }
    }
    if (cta == 0 && tid == 0) {
        CgScalars *sc = a.sc;
        sc->rz = st.gamma; sc->bb = st.bb; sc->stop = st.stop; sc->iters = it; sc->max_iter = a.max_iter;
        sc->done = st.done;
        sc->alpha = st.alpha; sc->beta = st.beta;
        sc->rseq_end = a.rseq0 + 5 + it; sc->hseq_end = a.hseq0 + 4 + it;
        if (st.err && sc->pad == 0) sc->pad = 4;
        if (PROF) {
            for (int q = 0; q < 6; ++q) a.prof[q] += prof[q];
            a.prof[6] += it;
            a.prof[7] += 1;
        }
    }
#undef PIPE_PROF
}
