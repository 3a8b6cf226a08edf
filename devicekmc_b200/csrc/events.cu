// devicekmc-b200 — per-event rate table and residence-time event selection.
//   a7  KMCProcess::update_events_and_rates (CPU semantics)   KMCProcess.cpp:67-164
//       build_event_list                                       kmc_events.cu:34-126
//   a8  KMCProcess::executeKMCStep loop (CPU semantics)        KMCProcess.cpp:297-358
//       execute_kmc_step_gpu / zero_out_events                 kmc_events.cu:146-365,129-143
// The reference re-runs thrust::inclusive_scan over all N*nn rates, a thrust::upper_bound and
// ~13 one-element cudaMemcpys for EVERY executed event.  Here the prefix sums are kept as a
// hierarchy of warp-wide (32-ary) segmented scans: row sums, then sums of 32 rows, ... up to
// one top group.  Selecting an event is a root-to-leaf walk (one warp scan + ballot per
// level = the binary search), executing it zeroes O(nn^2) entries and re-scans only the
// touched segments.  The whole residence-time loop of a KMC step is ONE persistent CTA; the
// host only supplies the pre-drawn uniforms of the reference's mt19937 stream.
//
// Bit-exactness of the selected index against the reference's strict left-to-right sum: both
// summation orders are within delta of the exact prefix sums, where delta is bounded from an
// exponent histogram of the rates (see select_walk).  If the target lies farther than delta
// from both bracketing prefix sums the sequential sum provably picks the same index; otherwise
// the selection is replayed with the exact sequential association.
//
// Compiled with -fmad=false so the rate arithmetic rounds like the reference's x86-64 build.
#include "common.cuh"
#include "scan.cuh"

namespace dkmc {

constexpr int kRateWarps = 8;             // warps (= rows in flight) per rate-table block
constexpr int kLoopThreads = 1024;        // the persistent event-loop CTA
constexpr int kMaxNN = 256;               // max neighbours per site supported by the loop
constexpr int kHistBuckets = 2048;        // one per FP64 exponent
constexpr int kTopCap = 1152;             // doubles of the prefix hierarchy's top levels kept in the loop's shared memory

struct Levels {
    int n_levels;                          // level 0 = row sums (N entries)
    int size[kMaxLevels];
    long long off[kMaxLevels];             // offsets into the level buffer (doubles)
};

struct EvState {
    double event_time, psum_last, delta_last;
    int n_events, n_used, done, n_fallback, n_none, status, pad0, pad1;
};

// ---------------------------------------------------------------- rate of one (i, j) pair
struct SiteI {
    int e, q, layer;
    double phi, x, y, z;
};

__device__ __forceinline__ void rate_entry(const SiteI &si, int j, const int *__restrict__ element,
                                           const int *__restrict__ charge, const int *__restrict__ layer,
                                           const double *__restrict__ pb, const double *__restrict__ pc,
                                           const double *__restrict__ x, const double *__restrict__ y,
                                           const double *__restrict__ z, const double *__restrict__ layerE,
                                           double ly, double lz, int pbc, double T_bg, double freq,
                                           double sigma, double k, int &type, double &P) {
    type = DKMC_NULL_EVENT;
    P = 0.0;
    const int ej = element[j];
    double E = 0.0, zf = 0.0;
    if (si.e == DKMC_DEFECT && ej == DKMC_O_EL) {
        double phij = pb[j] + pc[j];
        E = 2 * (si.phi - phij);
        zf = layerE[0 * kMaxLayers + layer[j]];
        type = DKMC_VACANCY_GENERATION;
    } else if (si.e == DKMC_OXYGEN_DEFECT && ej == DKMC_VACANCY) {
        double r = 1e-10 * site_dist_exact(si.x, si.y, si.z, x[j], y[j], z[j], ly, lz, pbc);
        double self_int_V = v_solve_ref(r, 2, sigma, k);
        int cs = si.q - charge[j];
        double phij = pb[j] + pc[j];
        E = cs * (si.phi - phij + (cs / 2) * self_int_V);
        zf = layerE[1 * kMaxLayers + layer[j]];
        type = DKMC_VACANCY_RECOMBINATION;
    } else if (si.e == DKMC_VACANCY && ej == DKMC_O_EL) {
        double self_int_V = 0.0;
        if (si.q != 0) {
            double r = 1e-10 * site_dist_exact(si.x, si.y, si.z, x[j], y[j], z[j], ly, lz, pbc);
            self_int_V = v_solve_ref(r, si.q, sigma, k);
        }
        double phij = pb[j] + pc[j];
        E = (si.q - charge[j]) * (si.phi - phij + self_int_V);
        zf = layerE[2 * kMaxLayers + si.layer];  // CPU path: layer of i (KMCProcess.cpp:134)
        type = DKMC_VACANCY_DIFFUSION;
    } else if (si.e == DKMC_OXYGEN_DEFECT && ej == DKMC_DEFECT) {
        double self_int_V = 0.0;
        if (si.q != 0) {
            double r = 1e-10 * site_dist_exact(si.x, si.y, si.z, x[j], y[j], z[j], ly, lz, pbc);
            self_int_V = v_solve_ref(r, 2, sigma, k);
        }
        double phij = pb[j] + pc[j];
        E = (si.q - charge[j]) * (si.phi - phij - self_int_V);
        zf = layerE[3 * kMaxLayers + layer[j]];
        type = DKMC_ION_DIFFUSION;
    }
    if (type != DKMC_NULL_EVENT) {
        double EA = zf - E - 0.0;
        P = exp(-1 * EA / (kBoltzmann * T_bg)) * freq;
    }
}

// lane-local sequential sum of the lane's `c` contiguous row entries, then the warp scan.
// The SAME routine produces the stored row sum and the cumulative used by the search.
__device__ __forceinline__ double row_scan(const double *vals, int c, int lane, double &lane_sum) {
    double s = 0.0;
    for (int t = 0; t < c; ++t) s = (t == 0) ? vals[0] : s + vals[t];
    lane_sum = s;
    return warp_inclusive_scan(s, lane);
}

// One warp per row: rates of the row's nn slots (lane owns slots [lane*c, lane*c + c)),
// the row sum, and the exponent histogram of the non-zero rates.
__global__ void __launch_bounds__(kRateWarps * 32) rate_rows_kernel(
    int N, int nn, int c, const int *__restrict__ neigh, const int *__restrict__ layer,
    const double *__restrict__ lattice, int pbc, const double *__restrict__ T_bg_p,
    const double *__restrict__ freq_p, const double *__restrict__ sigma_p, const double *__restrict__ k_p,
    const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
    const double *__restrict__ pb, const double *__restrict__ pc, const int *__restrict__ element,
    const int *__restrict__ charge, const double *__restrict__ layerE, int *__restrict__ ev_type,
    double *__restrict__ ev_prob, double *__restrict__ rowsum, int *hist) {
    __shared__ int sh_hist[kHistBuckets];
    const bool do_hist = hist != nullptr;
    if (do_hist) {
        for (int b = threadIdx.x; b < kHistBuckets; b += blockDim.x) sh_hist[b] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const double T_bg = *T_bg_p, freq = *freq_p, sigma = *sigma_p, kc = *k_p;
    const double ly = lattice[1], lz = lattice[2];
    for (int i = blockIdx.x * kRateWarps + w; i < N; i += gridDim.x * kRateWarps) {
        SiteI si;
        si.e = element[i];
        const bool active = (si.e == DKMC_DEFECT || si.e == DKMC_OXYGEN_DEFECT || si.e == DKMC_VACANCY);
        if (active) {
            si.q = charge[i]; si.layer = layer[i];
            si.phi = pb[i] + pc[i];
            si.x = x[i]; si.y = y[i]; si.z = z[i];
        }
        double vals[8];
        const size_t base = (size_t)i * nn;
        for (int t = 0; t < c; ++t) {
            int s = lane * c + t;
            int type = DKMC_NULL_EVENT;
            double P = 0.0;
            if (s < nn) {
                if (active) {
                    int j = neigh[base + s];
                    if (j >= 0 && j < N)
                        rate_entry(si, j, element, charge, layer, pb, pc, x, y, z, layerE, ly, lz, pbc, T_bg, freq,
                                   sigma, kc, type, P);
                }
                ev_type[base + s] = type;
                ev_prob[base + s] = P;
                if (do_hist && P != 0.0) atomicAdd(&sh_hist[(int)((__double_as_longlong(P) >> 52) & 0x7ff)], 1);
            }
            vals[t] = P;
        }
        double lane_sum;
        double inc = row_scan(vals, c, lane, lane_sum);
        if (lane == 31 && rowsum) rowsum[i] = inc;
    }
    if (do_hist) {
        __syncthreads();
        for (int b = threadIdx.x; b < kHistBuckets; b += blockDim.x)
            if (sh_hist[b]) atomicAdd(hist + b, sh_hist[b]);
    }
}

// level k+1 from level k: one warp per group of 32 children
__global__ void __launch_bounds__(256) level_build_kernel(int child_size, const double *__restrict__ child,
                                                          int parent_size, double *__restrict__ parent) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= parent_size) return;
    int idx = g * 32 + lane;
    double v = idx < child_size ? child[idx] : 0.0;
    double inc = warp_inclusive_scan(v, lane);
    if (lane == 31) parent[g] = inc;
}

// ---------------------------------------------------------------- the persistent event loop
struct LoopArgs {
    int N, nn, c;
    const int *neigh;
    int *ev_type;
    double *ev_prob;
    double *levels;
    Levels lv;
    int *element, *charge;
    const double *freq_p;
    const double *uniforms;
    int n_uniforms;
    int *events_out;       // 4 ints per event
    int max_events, events_base;
    const int *hist;
    EvState *state;
    int exact_mode;        // 1 = always replay exactly (test hook)
    int *nz_rows;          // scratch of select_exact: N ints
    double *batch_sum;     // scratch of select_exact: one running sum per batch of 32 row segments
};

__device__ __forceinline__ double ldv(const double *p) { return __ldcg(p); }
__device__ __forceinline__ int ldi(const int *p) { return __ldcg(p); }

// root-to-leaf walk by one warp.  out: idx (-1 = none), psum, flag (0 ok, 1 none, 2 needs exact)
// The top levels of the prefix hierarchy (1 + 32 + 1024 sums at 1 M sites) live in the loop's shared memory:
// level `lev` >= top_lev is read and written there (s_top, offsets relative to level top_lev's), the lower ones in
// global memory through L2 — three of the six dependent L2 round trips of a selection, and most of the re-scan.
struct TopLevels {
    double *s_top;
    int top_lev;
    long long top_off;
};
__device__ __forceinline__ double ld_level(const LoopArgs &a, const TopLevels &T, int lev, int idx) {
    return lev >= T.top_lev ? T.s_top[a.lv.off[lev] - T.top_off + idx] : ldv(a.levels + a.lv.off[lev] + idx);
}
__device__ __forceinline__ void st_level(const LoopArgs &a, const TopLevels &T, int lev, int idx, double v) {
    if (lev >= T.top_lev) T.s_top[a.lv.off[lev] - T.top_off + idx] = v;
    else __stcg(a.levels + a.lv.off[lev] + idx, v);
}

__device__ void select_walk(const LoopArgs &a, const TopLevels &T, double u, const double *errA, const int *errB, int lane,
                            int &idx_out, double &psum_out, int &flag_out, double &delta_out) {
    const Levels &lv = a.lv;
    double prefix = 0.0, number = 0.0, psum = 0.0, delta = 0.0;
    int g = 0, flag = 0;
    for (int lev = lv.n_levels - 1; lev >= 0; --lev) {
        int idx = g * 32 + lane;
        double v = idx < lv.size[lev] ? ld_level(a, T, lev, idx) : 0.0;
        double inc = warp_inclusive_scan(v, lane);
        if (lev == lv.n_levels - 1) {
            psum = __shfl_sync(0xffffffffu, inc, 31);
            number = u * psum;
            // Rounding-error bound of the strictly sequential sum of the current rates, from the
            // exponent histogram.  One rounded add errs by at most min(|a|, |b|) and by at most
            // u * |result| (u = 2^-53), so every sequential prefix sum is within
            //     E = sum_i min(x_i, u * Psum)  <=  E' = A[b_tau] + tau * B[b_tau],  tau = 2^-52 Psum
            // of the exact one (E' takes tau = 2u Psum and the upper edge of every exponent bucket: E' >= E,
            // up to 4 E).  A prefix sum of this walk goes through at most ~45 rounded adds of partial sums
            // <= Psum (per level five in the warp scan and one for the running prefix, a few more inside the
            // row): within 45 tau of the exact one.  The target u1 * Psum inherits the difference of the two
            // totals (<= E + 36 tau) plus one rounding (<= tau).  The sequential search therefore brackets the
            // same entry if the target is farther than 2 E + 82 tau from both bracketing prefix sums of this
            // walk; delta = 8 E' + 128 tau keeps a factor >= 4 on the dominant term.  (An earlier
            // 64 (E' + tau) sent 8x more selections than necessary to the exact replay, which at 4 M sites
            // costs ~250 ms.  tests/test_cpu_selection_margin.py replays this logic on the CPU.)
            double tau = psum * 2.220446049250313e-16 * 1.000001;
            int bt = (int)((__double_as_longlong(tau) >> 52) & 0x7ff);
            delta = 8.0 * (errA[bt] + tau * (double)errB[bt]) + 128.0 * tau;
        }
        double glob = prefix + inc;
        unsigned mask = __ballot_sync(0xffffffffu, glob > number);
        if (mask == 0u) { flag = (lev == lv.n_levels - 1) ? 1 : 2; break; }
        int l = __ffs(mask) - 1;
        double below = __shfl_sync(0xffffffffu, glob, l > 0 ? l - 1 : 0);
        if (l > 0) prefix = below;
        g = idx - lane + l;  // entry index at this level == group index one level down
    }
    int idx_sel = -1;
    if (flag == 0) {
        // g is the row; entries of the row, c per lane
        const size_t base = (size_t)g * a.nn;
        double vals[8];
        for (int t = 0; t < a.c; ++t) {
            int s = lane * a.c + t;
            vals[t] = s < a.nn ? ldv(a.ev_prob + base + s) : 0.0;
        }
        double lane_sum;
        double inc = row_scan(vals, a.c, lane, lane_sum);
        double excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0.0;
        int my_t = -1;
        double cum_before = 0.0, cum_at = 0.0, part = 0.0;
        for (int t = 0; t < a.c; ++t) {
            double prev = part;
            part = (t == 0) ? vals[0] : part + vals[t];
            double cum = prefix + (excl + part);
            if (my_t < 0 && cum > number) {
                my_t = t;
                cum_at = cum;
                cum_before = prefix + (excl + (t == 0 ? 0.0 : prev));
            }
        }
        unsigned mask = __ballot_sync(0xffffffffu, my_t >= 0);
        if (mask == 0u) flag = 2;
        else {
            int l = __ffs(mask) - 1;
            int t_sel = __shfl_sync(0xffffffffu, my_t, l);
            double cb = __shfl_sync(0xffffffffu, cum_before, l);
            double ca = __shfl_sync(0xffffffffu, cum_at, l);
            idx_sel = (int)base + l * a.c + t_sel;
            if (!(number - cb > delta) || !(ca - number > delta)) flag = 2;
        }
    }
    idx_out = idx_sel; psum_out = psum; flag_out = flag; delta_out = delta;
}

// Exact replay of utils.h:91-99 + std::upper_bound (rare path): the strictly left-to-right sum of
// the current rates.  Adding a zero changes nothing (x + 0.0 == x, and 0.0 + p == p), so only the
// non-zero entries have to be added, in table order.  The whole CTA cooperates, one thread adds:
//   A. ordered compaction of the rows whose row sum is not 0 (rates are >= 0: the sum is 0 iff every
//      entry is) into nz_rows — one coalesced pass over the N row sums;
//   B. items = (non-zero row, segment of 32 slots), one per warp and batch: the warp compacts the
//      non-zero entries of its segment into shared memory in slot order; thread 0 adds the batch
//      sequentially and records the running sum after each batch;
//   C. Psum = the last running sum; the first batch whose running sum exceeds u * Psum is replayed
//      from its starting value to find the entry.
// Cost: O(N / threads) + O(non-zero rows) instead of one thread walking all N rows twice (9 s per
// call at 4 M sites).  Must be called by ALL threads of the CTA; every thread returns the same result.
__device__ void select_exact(const LoopArgs &a, double u, int &idx_out, double &psum_out) {
    constexpr int kWarps = kLoopThreads / 32;
    static_assert(kWarps == 32, "the warp-count scan below maps one warp of the CTA to one lane");
    __shared__ int s_wcnt[2][kWarps];
    __shared__ double s_val[kWarps][32];
    __shared__ unsigned char s_slot[kWarps][32];
    __shared__ int s_cnt[kWarps], s_row[kWarps];
    __shared__ double s_acc;
    __shared__ int s_found, s_first_batch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *rowsum = a.levels + a.lv.off[0];
    // ---- A
    int total = 0, buf = 0;
    for (int base = 0; base < a.N; base += kLoopThreads, buf ^= 1) {
        const int r = base + tid;
        const bool nz = r < a.N && ldv(rowsum + r) != 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) s_wcnt[buf][warp] = __popc(m);
        __syncthreads();
        const int c = s_wcnt[buf][lane];
        const int inc = warp_inclusive_scan_int(c, lane);
        const int woff = __shfl_sync(0xffffffffu, inc - c, warp);
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (nz) a.nz_rows[total + woff + __popc(m & ((1u << lane) - 1u))] = r;
        total += tot;
    }
    __syncthreads();   // nz_rows is read below by other threads than its writers
    const int nseg = (a.nn + 31) / 32;
    const long long n_items = (long long)total * nseg;
    const int n_batches = (int)((n_items + kWarps - 1) / kWarps);
    // one batch: every warp stages its item; returns after the staging barrier
    auto stage = [&](int batch) {
        const long long item = (long long)batch * kWarps + warp;
        int cnt = 0, row = -1;
        if (item < n_items) {
            row = a.nz_rows[item / nseg];
            const int slot = (int)(item % nseg) * 32 + lane;
            const double v = slot < a.nn ? ldv(a.ev_prob + (size_t)row * a.nn + slot) : 0.0;
            const unsigned m = __ballot_sync(0xffffffffu, v != 0.0);
            if (v != 0.0) {
                const int pos = __popc(m & ((1u << lane) - 1u));
                s_val[warp][pos] = v;
                s_slot[warp][pos] = (unsigned char)slot;
            }
            cnt = __popc(m);
        }
        if (lane == 0) { s_cnt[warp] = cnt; s_row[warp] = row; }
        __syncthreads();
    };
    // ---- B
    if (tid == 0) { s_acc = 0.0; s_found = -1; s_first_batch = n_batches; }
    for (int batch = 0; batch < n_batches; ++batch) {
        stage(batch);
        if (tid == 0) {
            double acc = s_acc;
            for (int w = 0; w < kWarps; ++w)
                for (int k = 0; k < s_cnt[w]; ++k) acc = acc + s_val[w][k];
            s_acc = acc;
            a.batch_sum[batch] = acc;
        }
        __syncthreads();
    }
    __syncthreads();
    const double psum = s_acc;
    const double number = u * psum;
    // ---- C: first batch whose running sum exceeds the target (running sums never decrease)
    for (int b0 = tid; b0 < n_batches; b0 += kLoopThreads)
        if (a.batch_sum[b0] > number) atomicMin(&s_first_batch, b0);
    __syncthreads();
    const int fb = s_first_batch;
    if (fb < n_batches) {
        stage(fb);
        if (tid == 0) {
            double acc = fb > 0 ? a.batch_sum[fb - 1] : 0.0;
            for (int w = 0; w < kWarps && s_found < 0; ++w)
                for (int k = 0; k < s_cnt[w]; ++k) {
                    acc = acc + s_val[w][k];
                    if (acc > number) { s_found = s_row[w] * a.nn + (int)s_slot[w][k]; break; }
                }
        }
        __syncthreads();
    }
    idx_out = s_found;
    psum_out = psum;
    __syncthreads();   // the shared scratch may be reused by the next call
}

__global__ void __launch_bounds__(kLoopThreads, 1) event_loop_kernel(LoopArgs a) {
    __shared__ double errA[kHistBuckets];   // sum_{b' < b} count * 2^(b'+1)
    __shared__ int errB[kHistBuckets];      // sum_{b' >= b} count
    __shared__ int s_rows[2 + 2 * kMaxNN];
    __shared__ int s_idx, s_flag, s_done, s_used, s_nev, s_nfb, s_nnone;
    __shared__ double s_psum, s_delta, s_time;
    __shared__ double s_top[kTopCap];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = kLoopThreads / 32;
    // the levels from top_lev up fit the shared copy (every level is padded to a multiple of 32 entries)
    TopLevels T;
    T.s_top = s_top;
    T.top_lev = a.lv.n_levels;
    const long long lv_end = a.lv.off[a.lv.n_levels - 1] + (a.lv.size[a.lv.n_levels - 1] + 31) / 32 * 32;
    while (T.top_lev > 1 && lv_end - a.lv.off[T.top_lev - 1] <= kTopCap) --T.top_lev;
    T.top_off = T.top_lev < a.lv.n_levels ? a.lv.off[T.top_lev] : lv_end;
    for (long long k = tid; k < lv_end - T.top_off; k += kLoopThreads) s_top[k] = ldv(a.levels + T.top_off + k);

    // ---- error-bound tables from the exponent histogram
    for (int b = tid; b < kHistBuckets; b += kLoopThreads) errB[b] = a.hist[b];
    __syncthreads();
    if (warp == 0) {
        // serial-in-chunks prefix/suffix over 2048 buckets: 64 per lane
        double accA = 0.0;
        double laneA[1];
        int cntB = 0;
        for (int k = 0; k < 64; ++k) {
            int b = lane * 64 + k;
            int ee = b + 1 < 2047 ? b + 1 : 2046;
            double ub = __longlong_as_double((long long)ee << 52);
            accA += (double)errB[b] * ub;
            cntB += errB[b];
        }
        laneA[0] = accA;
        double incA = warp_inclusive_scan(accA, lane);
        int incB = warp_inclusive_scan_int(cntB, lane);
        double exA = incA - laneA[0];
        int totalB = __shfl_sync(0xffffffffu, incB, 31);
        int sufB = totalB - (incB - cntB);  // count in buckets >= lane*64
        double runA = exA;
        for (int k = 0; k < 64; ++k) {
            int b = lane * 64 + k;
            int cnt = errB[b];
            int ee = b + 1 < 2047 ? b + 1 : 2046;
            double ub = __longlong_as_double((long long)ee << 52);
            errA[b] = runA;
            errB[b] = sufB;
            runA += (double)cnt * ub;
            sufB -= cnt;
        }
    }
    if (tid == 0) {
        s_done = a.state->done; s_used = 0; s_nev = a.state->n_events; s_nfb = a.state->n_fallback;
        s_nnone = a.state->n_none; s_time = a.state->event_time; s_psum = a.state->psum_last; s_delta = 0.0;
    }
    __syncthreads();
    const double inv_freq = 1.0 / *a.freq_p;
    const int n_touch = 2 + 2 * a.nn;

    while (true) {
        if (s_done || s_used + 2 > a.n_uniforms) break;
        const double u1 = a.uniforms[s_used];
        // ---- select
        if (warp == 0) {
            int idx, flag;
            double psum, delta;
            select_walk(a, T, u1, errA, errB, lane, idx, psum, flag, delta);
            if (a.exact_mode == 1 && flag == 0) flag = 2;
            if (lane == 0) { s_idx = idx; s_flag = flag; s_psum = psum; s_delta = delta; }
        }
        __syncthreads();
        if (s_flag == 2) {   // uniform: s_flag is shared
            int idx;
            double psum;
            select_exact(a, u1, idx, psum);
            if (tid == 0) { s_idx = idx; s_psum = psum; s_flag = idx >= 0 ? 0 : 1; s_nfb += 1; }
            __syncthreads();
        }
        const int idx = s_flag == 0 ? s_idx : -1;
        if (idx >= 0) {
            const int i = idx / a.nn;
            const int j = ldi(a.neigh + idx);
            // touched rows: i, j, neighbours of i, neighbours of j
            for (int t = tid; t < n_touch; t += kLoopThreads) {
                int r;
                if (t == 0) r = i;
                else if (t == 1) r = j;
                else if (t < 2 + a.nn) r = ldi(a.neigh + (size_t)i * a.nn + (t - 2));
                else r = ldi(a.neigh + (size_t)j * a.nn + (t - 2 - a.nn));
                s_rows[t] = r;
            }
            if (tid == 0) {
                const int type = ldi(a.ev_type + idx);
                int ei = ldi(a.element + i), ej = ldi(a.element + j), qi = ldi(a.charge + i), qj = ldi(a.charge + j);
                switch (type) {  // KMCProcess.cpp:187-256
                case DKMC_VACANCY_GENERATION: ei = DKMC_OXYGEN_DEFECT; qi = -2; ej = DKMC_VACANCY; qj = 2; break;
                case DKMC_VACANCY_RECOMBINATION: ei = DKMC_DEFECT; qi = 0; ej = DKMC_O_EL; qj = 0; break;
                case DKMC_VACANCY_DIFFUSION:
                case DKMC_ION_DIFFUSION: { int te = ei; ei = ej; ej = te; int tq = qi; qi = qj; qj = tq; break; }
                default: break;
                }
                __stcg(a.element + i, ei); __stcg(a.element + j, ej);
                __stcg(a.charge + i, qi); __stcg(a.charge + j, qj);
                int slot = a.events_base + s_nev;
                if (a.events_out && slot < a.max_events) {
                    a.events_out[4 * slot + 0] = idx; a.events_out[4 * slot + 1] = i;
                    a.events_out[4 * slot + 2] = j; a.events_out[4 * slot + 3] = type;
                }
            }
            __syncthreads();
            // zero conflicting events (KMCProcess.cpp:330-352) and re-scan the touched rows
            double *rowsum = a.levels + a.lv.off[0];
            if (a.c <= 2) {
                // two rows of a warp at a time, all their loads requested before the first is used (the rounds
                // are independent: without this a warp pays one L2 round trip per row)
                for (int t0 = warp; t0 < n_touch; t0 += 2 * nwarps) {
                    double pv[2][2];
                    int jv[2][2], rr[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int t = t0 + h * nwarps;
                        rr[h] = t < n_touch ? s_rows[t] : -1;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int sl = lane * a.c + k;
                            const bool ok = rr[h] >= 0 && k < a.c && sl < a.nn;
                            pv[h][k] = ok ? ldv(a.ev_prob + (size_t)rr[h] * a.nn + sl) : 0.0;
                            jv[h][k] = ok ? ldi(a.neigh + (size_t)rr[h] * a.nn + sl) : -1;
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = rr[h];
                        if (r < 0) continue;
                        const size_t base = (size_t)r * a.nn;
                        double vals[2] = {0.0, 0.0};
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int sl = lane * a.c + k;
                            if (k < a.c && sl < a.nn) {
                                double p = pv[h][k];
                                // rows i and j go entirely (also when reached again as a neighbour row, so that
                                // duplicate visits of a row write identical results)
                                const bool kill = (r == i || r == j) || jv[h][k] == i || jv[h][k] == j;
                                if (kill) {
                                    if (p != 0.0) __stcg(a.ev_prob + base + sl, 0.0);
                                    __stcg(a.ev_type + base + sl, (int)DKMC_NULL_EVENT);
                                    p = 0.0;
                                }
                                vals[k] = p;
                            }
                        }
                        double lane_sum;
                        double inc = row_scan(vals, a.c, lane, lane_sum);
                        if (lane == 31) __stcg(rowsum + r, inc);
                    }
                }
            } else {
                for (int t = warp; t < n_touch; t += nwarps) {
                    const int r = s_rows[t];
                    if (r < 0) continue;
                    const size_t base = (size_t)r * a.nn;
                    double vals[8];
                    for (int k = 0; k < a.c; ++k) {
                        int s = lane * a.c + k;
                        double p = 0.0;
                        if (s < a.nn) {
                            p = ldv(a.ev_prob + base + s);
                            bool kill = (r == i || r == j);
                            if (!kill) { int jj = ldi(a.neigh + base + s); kill = (jj == i || jj == j); }
                            if (kill) {
                                if (p != 0.0) __stcg(a.ev_prob + base + s, 0.0);
                                __stcg(a.ev_type + base + s, (int)DKMC_NULL_EVENT);
                                p = 0.0;
                            }
                        }
                        vals[k] = p;
                    }
                    double lane_sum;
                    double inc = row_scan(vals, a.c, lane, lane_sum);
                    if (lane == 31) __stcg(rowsum + r, inc);
                }
            }
            __syncthreads();
            // propagate upwards: re-scan every touched group, level by level (a warp's groups all requested first)
            for (int lev = 1; lev < a.lv.n_levels; ++lev) {
                const int csize = a.lv.size[lev - 1];
                for (int t0 = warp; t0 < n_touch; t0 += 4 * nwarps) {
                    double v[4];
                    int g[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const int t = t0 + h * nwarps;
                        const int r = t < n_touch ? s_rows[t] : -1;
                        g[h] = r >= 0 ? (r >> (5 * lev)) : -1;
                        const int ci = g[h] * 32 + lane;
                        v[h] = (g[h] >= 0 && ci < csize) ? ld_level(a, T, lev - 1, ci) : 0.0;
                    }
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        if (g[h] < 0) continue;
                        const double inc = warp_inclusive_scan(v[h], lane);
                        if (lane == 31) st_level(a, T, lev, g[h], inc);
                    }
                }
                __syncthreads();
            }
        }
        if (tid == 0) {
            const double u2 = a.uniforms[s_used + 1];
            s_time = -log(u2) / s_psum;  // KMCProcess.cpp:354
            s_used += 2;
            if (idx >= 0) s_nev += 1; else s_nnone += 1;
            if (!(s_time < inv_freq)) s_done = 1;
            // a rate overflowed (exp of a huge energy): the sum is inf or NaN, no entry can be selected and the drawn
            // time is 0 for ever — stop with an error status instead of burning every batch of uniforms
            if (!(s_psum == s_psum) || s_psum > 1.7e308) { s_done = 1; a.state->status = 1; }
        }
        __syncthreads();
    }
    // the shared top levels back to global memory (continue calls, the tests' dkmc_last_event_tables)
    for (long long k = tid; k < lv_end - T.top_off; k += kLoopThreads) __stcg(a.levels + T.top_off + k, s_top[k]);
    if (tid == 0) {
        a.state->event_time = s_time; a.state->psum_last = s_psum; a.state->delta_last = s_delta;
        a.state->n_events = s_nev; a.state->n_used = s_used; a.state->done = s_done;
        a.state->n_fallback = s_nfb; a.state->n_none = s_nnone;
    }
}

// ---------------------------------------------------------------- exported scan + search primitive
__global__ void upper_bound_kernel(long long n, const double *__restrict__ cum, double u, long long *idx,
                                   double *psum) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double total = cum[n - 1];
    double number = u * total;
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = lo + (hi - lo) / 2;
        if (cum[mid] > number) hi = mid; else lo = mid + 1;
    }
    *idx = lo;
    *psum = total;
}

static int build_levels_desc(int N, Levels *lv, long long *total) {
    int n = N, k = 0;
    long long off = 0;
    while (true) {
        if (k >= kMaxLevels) return DKMC_ERR_ARG;
        lv->size[k] = n;
        lv->off[k] = off;
        off += (n + 31) / 32 * 32;
        ++k;
        if (n <= 32) break;
        n = (n + 31) / 32;
    }
    lv->n_levels = k;
    *total = off;
    return DKMC_OK;
}

static int launch_rate_table(dkmc_ctx *ctx, int N, int nn, const int *d_neigh, const int *d_layer,
                             const double *d_lattice, int pbc, const double *d_T_bg, const double *d_freq,
                             const double *d_sigma, const double *d_k, const double *d_x, const double *d_y,
                             const double *d_z, const double *d_pb, const double *d_pc, const int *d_element,
                             const int *d_charge, int *d_type, double *d_prob, double *d_rowsum, int *d_hist) {
    int c = (nn + 31) / 32;
    DKMC_REQUIRE(c <= 8, "nn must be <= 256");
    DKMC_REQUIRE(ctx->n_layers > 0, "dkmc_set_layer_energies must be called first");
    int grid = ceil_div(N, kRateWarps);
    int cap = ctx->num_sms * 16;
    if (grid > cap) grid = cap;
    DKMC_LAUNCH(ctx, rate_rows_kernel, grid, kRateWarps * 32, 0, N, nn, c, d_neigh, d_layer, d_lattice, pbc, d_T_bg,
                d_freq, d_sigma, d_k, d_x, d_y, d_z, d_pb, d_pc, d_element, d_charge, ctx->d_layerE, d_type, d_prob,
                d_rowsum, d_hist);
    return DKMC_OK;
}

static int run_loop(dkmc_ctx *ctx, const double *uniforms, int n_uniforms, int *events_out, int max_events,
                    dkmc_step_info *info) {
    auto &ev = ctx->ev;
    DKMC_REQUIRE(ev.active, "no KMC step in progress");
    DKMC_REQUIRE(uniforms && n_uniforms >= 2, "need at least two uniforms");
    double *d_u;
    int *d_events;
    EvState *d_state;
    int rc;
    if ((rc = ensure<double>(ctx, S_EV_UNIFORMS, (size_t)n_uniforms, &d_u))) return rc;
    int cap_events = max_events > 0 ? max_events : 1;
    // the event record lives in its own slot and must survive across continue calls
    if (ctx->slot_cap[S_EV_EVENTS] < (size_t)cap_events * 16) {
        DKMC_REQUIRE(ev.events_done == 0, "max_events must not grow between dkmc_kmc_step_continue calls");
    }
    if ((rc = ensure<int>(ctx, S_EV_EVENTS, (size_t)cap_events * 4, &d_events))) return rc;
    d_state = static_cast<EvState *>(ctx->slot_ptr[S_EV_STATE]);
    DKMC_CUDA(cudaMemcpyAsync(d_u, uniforms, (size_t)n_uniforms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LoopArgs a;
    a.N = ev.N; a.nn = ev.nn; a.c = (ev.nn + 31) / 32;
    a.neigh = ev.d_neigh;
    a.ev_type = static_cast<int *>(ctx->slot_ptr[S_EV_TYPE]);
    a.ev_prob = static_cast<double *>(ctx->slot_ptr[S_EV_PROB]);
    a.levels = static_cast<double *>(ctx->slot_ptr[S_EV_LEVELS]);
    a.lv.n_levels = ev.n_levels;
    for (int k = 0; k < kMaxLevels; ++k) { a.lv.size[k] = ev.level_size[k]; a.lv.off[k] = ev.level_off[k]; }
    a.element = ev.d_element; a.charge = ev.d_charge;
    a.freq_p = ev.d_freq;
    a.uniforms = d_u; a.n_uniforms = n_uniforms;
    a.events_out = events_out ? d_events : nullptr;
    a.max_events = max_events; a.events_base = 0;
    a.hist = static_cast<int *>(ctx->slot_ptr[S_EV_SCRATCH]);
    {
        const size_t items = (size_t)ev.N * (size_t)((ev.nn + 31) / 32);
        if ((rc = ensure<int>(ctx, S_EV_NZROWS, (size_t)ev.N, &a.nz_rows))) return rc;
        if ((rc = ensure<double>(ctx, S_EV_BATCHSUM, items / (kLoopThreads / 32) + 2, &a.batch_sum))) return rc;
    }
    a.state = d_state;
    a.exact_mode = ctx->exact_select;
    DKMC_LAUNCH(ctx, event_loop_kernel, 1, kLoopThreads, 0, a);
    EvState h;
    DKMC_CUDA(cudaMemcpyAsync(&h, d_state, sizeof(EvState), cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (events_out && h.n_events > 0) {
        int ncopy = h.n_events < max_events ? h.n_events : max_events;
        DKMC_CUDA(cudaMemcpyAsync(events_out, d_events, (size_t)ncopy * 16, cudaMemcpyDeviceToHost, ctx->stream));
        DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ev.events_done = h.n_events;
    if (info) {
        info->n_events = h.n_events;
        info->n_used = h.n_used;
        info->n_exact_fallbacks = h.n_fallback;
        info->event_time = h.event_time;
    }
    if (h.status != 0) {
        ev.active = false;
        set_error("event loop: the sum of the rates is not finite (a rate overflowed) after %d events", h.n_events);
        return DKMC_ERR_ARG;
    }
    if (!h.done) {
        set_error("event loop consumed all %d uniforms after %d events; call dkmc_kmc_step_continue", n_uniforms, h.n_events);
        return DKMC_ERR_RNG_EXHAUSTED;
    }
    ev.active = false;
    return DKMC_OK;
}

}  // namespace dkmc

using namespace dkmc;

extern "C" {

int dkmc_build_event_list(dkmc_ctx *ctx, int N, int nn, const int *d_neigh_idx, const int *d_site_layer,
                          const double *d_lattice, int pbc, const double *d_T_bg, const double *d_freq,
                          const double *d_sigma, const double *d_k, const double *d_x, const double *d_y,
                          const double *d_z, const double *d_potential_boundary, const double *d_potential_charge,
                          const int *d_site_element, const int *d_site_charge, int *d_event_type,
                          double *d_event_prob) {
    DKMC_REQUIRE(ctx && d_neigh_idx && d_site_layer && d_lattice && d_T_bg && d_freq && d_sigma && d_k && d_x && d_y &&
                     d_z && d_potential_boundary && d_potential_charge && d_site_element && d_site_charge &&
                     d_event_type && d_event_prob, "null pointer");
    int rc = launch_rate_table(ctx, N, nn, d_neigh_idx, d_site_layer, d_lattice, pbc, d_T_bg, d_freq, d_sigma, d_k,
                               d_x, d_y, d_z, d_potential_boundary, d_potential_charge, d_site_element, d_site_charge,
                               d_event_type, d_event_prob, nullptr, nullptr);
    if (rc) return rc;
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return DKMC_OK;
}

int dkmc_inclusive_scan(dkmc_ctx *ctx, long long n, const double *d_in, double *d_out) {
    DKMC_REQUIRE(ctx && d_in && d_out && n > 0, "arguments");
    double *tmp;
    int rc;
    if ((rc = ensure<double>(ctx, S_SCAN_TMP, (size_t)ceil_div(n, kScanTile) + 1, &tmp))) return rc;
    return inclusive_scan<double>(ctx, d_in, n, d_out, tmp);
}

int dkmc_select_event(dkmc_ctx *ctx, long long n, const double *d_cum, double u, long long *idx, double *Psum) {
    DKMC_REQUIRE(ctx && d_cum && idx && Psum && n > 0, "arguments");
    void *out;
    int rc;
    if ((rc = ensure_slot(ctx, S_SEL_OUT, 16, &out))) return rc;
    DKMC_LAUNCH(ctx, upper_bound_kernel, 1, 32, 0, n, d_cum, u, static_cast<long long *>(out),
                reinterpret_cast<double *>(static_cast<char *>(out) + 8));
    char h[16];
    DKMC_CUDA(cudaMemcpyAsync(h, out, 16, cudaMemcpyDeviceToHost, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(idx, h, 8);
    memcpy(Psum, h + 8, 8);
    return DKMC_OK;
}

int dkmc_execute_kmc_step(dkmc_ctx *ctx, int N, int nn, const int *d_neigh_idx, const int *d_site_layer,
                          const double *d_lattice, int pbc, const double *d_T_bg, const double *d_freq,
                          const double *d_sigma, const double *d_k, const double *d_x, const double *d_y,
                          const double *d_z, const double *d_potential_boundary, const double *d_potential_charge,
                          int *d_site_element, int *d_site_charge, const double *uniforms, int n_uniforms,
                          int *events_out, int max_events, dkmc_step_info *info) {
    DKMC_REQUIRE(ctx && d_neigh_idx && d_site_layer && d_lattice && d_T_bg && d_freq && d_sigma && d_k && d_x && d_y &&
                     d_z && d_potential_boundary && d_potential_charge && d_site_element && d_site_charge,
                 "null pointer");
    DKMC_REQUIRE(N > 0 && nn > 0 && nn <= kMaxNN, "N > 0 and 0 < nn <= 256");
    DKMC_REQUIRE((long long)N * nn < 2147483647ll, "N * nn must fit 32 bits: event indices are ints (about 41 M sites at nn = 52)");
    Levels lv;
    long long lv_total;
    int rc;
    if ((rc = build_levels_desc(N, &lv, &lv_total))) { set_error("too many sites for the level hierarchy"); return rc; }
    int *d_type, *d_hist;
    double *d_prob, *d_levels;
    EvState *d_state;
    const size_t total = (size_t)N * nn;
    if ((rc = ensure<int>(ctx, S_EV_TYPE, total, &d_type))) return rc;
    if ((rc = ensure<double>(ctx, S_EV_PROB, total, &d_prob))) return rc;
    if ((rc = ensure<double>(ctx, S_EV_LEVELS, (size_t)lv_total, &d_levels))) return rc;
    if ((rc = ensure<int>(ctx, S_EV_SCRATCH, kHistBuckets, &d_hist))) return rc;
    if ((rc = ensure<EvState>(ctx, S_EV_STATE, 1, &d_state))) return rc;
    DKMC_CUDA(cudaMemsetAsync(d_hist, 0, kHistBuckets * sizeof(int), ctx->stream));
    DKMC_CUDA(cudaMemsetAsync(d_state, 0, sizeof(EvState), ctx->stream));
    DKMC_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    if ((rc = launch_rate_table(ctx, N, nn, d_neigh_idx, d_site_layer, d_lattice, pbc, d_T_bg, d_freq, d_sigma, d_k,
                                d_x, d_y, d_z, d_potential_boundary, d_potential_charge, d_site_element,
                                d_site_charge, d_type, d_prob, d_levels + lv.off[0], d_hist))) return rc;
    for (int k = 1; k < lv.n_levels; ++k)
        DKMC_LAUNCH(ctx, level_build_kernel, ceil_div(lv.size[k], 8), 256, 0, lv.size[k - 1], d_levels + lv.off[k - 1],
                    lv.size[k], d_levels + lv.off[k]);
    DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    auto &ev = ctx->ev;
    ev.N = N; ev.nn = nn; ev.d_neigh = d_neigh_idx; ev.d_element = d_site_element; ev.d_charge = d_site_charge;
    ev.d_freq = d_freq; ev.n_levels = lv.n_levels;
    for (int k = 0; k < kMaxLevels; ++k) { ev.level_size[k] = lv.size[k]; ev.level_off[k] = lv.off[k]; }
    ev.events_done = 0;
    ev.active = true;
    rc = run_loop(ctx, uniforms, n_uniforms, events_out, max_events, info);
    if (rc != DKMC_OK && rc != DKMC_ERR_RNG_EXHAUSTED) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_c, ctx->stream));
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_c));
    if (info) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev_a, ctx->ev_b);
        cudaEventElapsedTime(&b, ctx->ev_b, ctx->ev_c);
        info->rate_ms = a;
        info->loop_ms = b;
    }
    return rc;
}

int dkmc_kmc_step_continue(dkmc_ctx *ctx, const double *uniforms, int n_uniforms, int *events_out, int max_events,
                           dkmc_step_info *info) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    int rc = run_loop(ctx, uniforms, n_uniforms, events_out, max_events, info);
    if (rc != DKMC_OK && rc != DKMC_ERR_RNG_EXHAUSTED) return rc;
    DKMC_CUDA(cudaEventRecord(ctx->ev_c, ctx->stream));
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_c));
    if (info) {   // this call's share of the loop; the caller adds it to the first call's loop_ms
        float b = 0;
        cudaEventElapsedTime(&b, ctx->ev_b, ctx->ev_c);
        info->rate_ms = 0.0;
        info->loop_ms = b;
    }
    return rc;
}

int dkmc_last_event_tables(dkmc_ctx *ctx, const int **d_event_type, const double **d_event_prob) {
    DKMC_REQUIRE(ctx && d_event_type && d_event_prob, "null pointer");
    *d_event_type = static_cast<const int *>(ctx->slot_ptr[S_EV_TYPE]);
    *d_event_prob = static_cast<const double *>(ctx->slot_ptr[S_EV_PROB]);
    return DKMC_OK;
}

}  // extern "C"
