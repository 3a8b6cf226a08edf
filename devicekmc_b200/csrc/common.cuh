// devicekmc-b200 — shared host/device plumbing for the sm_100a kernels.
// Context (workspace arena + stream), error handling, launch accounting and the warp/block
// reduction primitives every stage uses.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "dkmc.h"

namespace dkmc {

constexpr int kWarp = 32;
constexpr int kMaxLayers = 16;
constexpr int kNumSlots = 96;
constexpr int kMaxLevels = 8;

void set_error(const char *fmt, ...);
int window_carveout();  // shared-memory carve-out (percent) common to the kernels of the overlap window

#define DKMC_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            ::dkmc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                   \
                              cudaGetErrorString(err__));                                     \
            return DKMC_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define DKMC_REQUIRE(cond, msg)                                                               \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::dkmc::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);        \
            return DKMC_ERR_ARG;                                                              \
        }                                                                                     \
    } while (0)

// Scratch slots: a persistent arena of named device buffers that only ever grow.
enum Slot : int {
    S_SCALARS = 0,   // CG scalars + reduction partials + counters
    S_PARTIALS,
    S_CG_R, S_CG_P, S_CG_AP, S_CG_Z, S_CG_DINV, S_CG_RES, S_CG_E, S_CG_VAL, S_CG_RHS,
    S_SPMV_TILES,
    S_CLASS,         // per-site conductance class byte
    S_NB_CELL_OF, S_NB_CELL_START, S_NB_CELL_FILL, S_NB_CELL_SITES, S_NB_DEG, S_NB_BOUNDS,
    S_SP_CNT, S_SCAN_BLOCK,
    S_PW_FLAGS, S_PW_SRC, S_PW_COUNT,
    S_EV_TYPE, S_EV_PROB, S_EV_LEVELS, S_EV_STATE, S_EV_UNIFORMS, S_EV_EVENTS, S_EV_SCRATCH,
    S_SCAN_TMP, S_SEL_OUT, S_CL_INT, S_CL_KEYS, S_CL_W, S_DIST_RED, S_PW_TILECTR, S_CG_S, S_CL_REC, S_PCG_SYNC, S_PCG_SYNC2, S_PCG_PROF, S_PCG_CLK, S_CG_PZ, S_CG_PN, S_ORD_VAL, S_ORD_VEC, S_ORD_CLS, S_ORD_TMP, S_PW_BOX, S_PW_CELLS, S_PW_SRC2, S_PW_IDX2,
    S_NB_HOSTPOS, S_NB_HOSTTAB, S_PW_PREVQ, S_PW_DQ, S_SNAP_STAGE, S_EV_NZROWS, S_EV_BATCHSUM,
    S_LAST
};
static_assert(S_LAST <= kNumSlots, "increase kNumSlots");

struct SpmvTiling {
    const int *row_ptr = nullptr;  // key
    int m = 0, nnz = 0;
    int num_tiles = 0;
    int *d_tile_row = nullptr;     // [num_tiles+1] first row of each nnz tile
};

}  // namespace dkmc

struct dkmc_ctx {
    cudaStream_t stream = nullptr;
    int dev = 0;
    int num_sms = 148;
    long long launches = 0;
    void *slot_ptr[dkmc::kNumSlots] = {};
    size_t slot_cap[dkmc::kNumSlots] = {};
    // rate-table layer energies: [E_gen | E_rec | E_Vdiff | E_Odiff] x kMaxLayers
    double *d_layerE = nullptr;
    int n_layers = 0;
    // neighbour-grid cache between count and fill
    struct {
        int N = 0, ncx = 0, ncy = 0, ncz = 0, pbc = 0;
        double minx = 0, miny = 0, minz = 0, wx = 0, wy = 0, wz = 0;
        bool valid = false;
    } grid;
    dkmc::SpmvTiling tiling;
    // event loop state for dkmc_kmc_step_continue
    struct {
        int N = 0, nn = 0;
        const int *d_neigh = nullptr;
        int *d_element = nullptr, *d_charge = nullptr;
        const double *d_freq = nullptr;
        int n_levels = 0;
        long long level_off[dkmc::kMaxLevels + 1] = {};
        int level_size[dkmc::kMaxLevels + 1] = {};
        int events_done = 0;
        bool active = false;
    } ev;
    int exact_select = 0;
    void *dist = nullptr;  // DistState (NCCL communicator) when running slab-partitioned
    int legacy_cg = 0;     // 1: one kernel per CG operation (round-1 path) instead of the persistent PCG
    void *solver_order = nullptr;   // SolverOrder (solver.cu): internal row order registered by dkmc_solver_set_order
    void *selfwin = nullptr;   // SelfWindow (solver.cu): the persistent PCG's window on one GPU
    int pw_regs_per_thread = 80;   // registers of the overlapped pairwise kernel (pairwise.cu refreshes it)
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    // side stream: the pairwise sum (FP64-bound) runs there while the CG (HBM-bound) runs on `stream`
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_pw0 = nullptr, ev_pw1 = nullptr;
    struct {
        bool active = false;
        int pbc = 0, N = 0, row_begin = 0, row_end = 0;
        const int *d_charge = nullptr;
        double *d_out = nullptr;
    } pw_pending;
    int pcg_pipelined = -1;          // persistent PCG: pipelined recurrences (one synchronisation per iteration); -1 = from four GPUs on
    int pw_far_field = 1;            // cell-list kernel: far-field formula for the runs beyond t = kErfcFarT0
    int pw_use_cells = 1;            // skip sources beyond the distance where erfc is exactly 0 (non-periodic devices)
    struct { const double *d_x = nullptr, *d_sigma = nullptr; const void *box = nullptr; int N = 0; double cutoff_sigmas = 0.0; } pw_grid;
    double pw_cutoff_sigmas = 0.0;   // 0: exact; > 0: truncate the pairwise sum at this many sigma (opt-in)
    // opt-in incremental update of phi_c: only the sites whose charge changed since the last
    // call are summed, phi_c += delta; a full sum every `refresh_every` calls bounds the rounding drift
    struct {
        int refresh_every = 0;           // 0: off
        int since_full = 0;
        bool valid = false;              // prev_charge mirrors the charges d_out was computed from
        const int *d_charge = nullptr; const double *d_out = nullptr; int N = 0, row_begin = 0, row_end = 0, pbc = 0;
        long long full_sums = 0, delta_sums = 0;
    } pw_inc;
    // snapshot copies: staged on `stream`, moved to the host on `io_stream`
    cudaStream_t io_stream = nullptr;
    cudaEvent_t ev_snap_staged = nullptr, ev_snap_done = nullptr;
    bool snap_pending = false;
    int pw_side_threads = 128;
    int pw_side_blocks_per_sm = 3;   // residency of the pairwise kernel while it shares the SMs with the CG
};

namespace dkmc {

void free_solver_state(dkmc_ctx *ctx);   // solver.cu: the single-GPU window of the persistent PCG

// returns a device buffer of at least `bytes` for `slot` (contents NOT preserved on growth)
int ensure_slot(dkmc_ctx *ctx, int slot, size_t bytes, void **out);

template <typename T>
inline int ensure(dkmc_ctx *ctx, int slot, size_t count, T **out) {
    void *p = nullptr;
    int rc = ensure_slot(ctx, slot, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return rc;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

#define DKMC_LAUNCH(ctx, kernel, grid, block, smem, ...) \
    DKMC_LAUNCH_ON(ctx, (ctx)->stream, kernel, grid, block, smem, __VA_ARGS__)

// Files whose kernels run while the pairwise sum shares the SMs with the CG define
// DKMC_CARVEOUT_MAXSHARED: the L1/shared split of an SM cannot change while CTAs are resident, so
// every kernel of that window asks for the SAME carve-out (window_carveout(): 60 % shared memory,
// the rest L1 for the SpMV's x gathers) — otherwise a CTA that needs another split waits until the
// SM has drained, which serialises the two streams.
#ifdef DKMC_CARVEOUT_MAXSHARED
#define DKMC_SET_CARVEOUT(kernel)                                                             \
    do {                                                                                      \
        static bool cfg__ = false;                                                            \
        if (!cfg__) {                                                                         \
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,      \
                                 ::dkmc::window_carveout());                                  \
            cfg__ = true;                                                                     \
        }                                                                                     \
    } while (0)
// the same for a kernel chosen at run time (function pointer): set on every launch, it is a cheap driver call
#define DKMC_SET_CARVEOUT_FN(fn) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, ::dkmc::window_carveout())
#else
#define DKMC_SET_CARVEOUT(kernel) do { } while (0)
#define DKMC_SET_CARVEOUT_FN(fn) do { } while (0)
#endif

#define DKMC_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                             \
    do {                                                                                      \
        DKMC_SET_CARVEOUT(kernel);                                                            \
        kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                             \
        (ctx)->launches++;                                                                    \
        cudaError_t err__ = cudaPeekAtLastError();                                            \
        if (err__ != cudaSuccess) {                                                           \
            ::dkmc::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel,  \
                              cudaGetErrorString(err__));                                     \
            return DKMC_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

// ---------------------------------------------------------------- device primitives
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// inclusive scan across the warp, fixed (deterministic) Hillis-Steele order
__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ int warp_inclusive_scan_int(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// block-wide sum; result valid in thread 0.  `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (w == 0) {
        r = lane < nw ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// Deterministic grid-wide sum: every block deposits its partial, the last block to arrive
// adds the partials in index order and publishes the total.  `counter` must be 0 on entry and
// is reset to 0 by the last block.  Returns true in thread 0 of the last block only.
__device__ __forceinline__ bool grid_sum_finish(double block_total, double *partials,
                                                unsigned int *counter, double *result,
                                                double *sh) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_total;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // fixed order (thread t adds partials t, t+B, t+2B, ...), loads issued eight at a time so the
    // tail of the launch costs one or two L2 round trips instead of gridDim/blockDim of them
    double acc = 0.0;
    const int n = (int)gridDim.x, B = (int)blockDim.x;
    int i = threadIdx.x;
    for (; i + 7 * B < n; i += 8 * B) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(partials + i + u * B);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; i < n; i += B) acc += __ldcg(partials + i);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        *result = acc;
        *counter = 0u;
        __threadfence();
        return true;
    }
    return false;
}

// utils.cpp:100-137 with every operation individually rounded (no FMA contraction), so the
// `dist < nn_dist` predicate matches the reference's x86-64 build bit for bit.
__device__ __forceinline__ double site_dist_exact(double x1, double y1, double z1, double x2,
                                                  double y2, double z2, double ly, double lz,
                                                  int pbc) {
    if (pbc == 1) {
        double dx = __dsub_rn(x1, x2);
        double fy = __ddiv_rn(__dsub_rn(y1, y2), ly);
        fy = __dsub_rn(fy, round(fy));
        double fz = __ddiv_rn(__dsub_rn(z1, z2), lz);
        fz = __dsub_rn(fz, round(fz));
        double dy = __dmul_rn(fy, ly), dz = __dmul_rn(fz, lz);
        return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    }
    double ax = __dsub_rn(x2, x1), ay = __dsub_rn(y2, y1), az = __dsub_rn(z2, z1);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
}

constexpr double kElementaryCharge = 1.60217663e-19;  // Device.h:113
constexpr double kBoltzmann = 8.617333262e-5;          // KMCProcess.h:35

// utils.h:102 — potential of a gaussian charge distribution, reference operation order
__device__ __forceinline__ double v_solve_ref(double r_dist, int charge, double sigma, double k) {
    return (double)charge * erfc(r_dist / (sigma * sqrt(2.0))) * k * kElementaryCharge / r_dist;
}

}  // namespace dkmc
