// devicekmc-b200 — context, error text, device selection (C-ABI plumbing).
// Replaces get_gpu_info / set_gpu (kmc_events.cu:15-32) and the per-call cudaMalloc/cudaFree
// churn of the reference (potential_solver_gpu.cu:397-493,735-779) with a persistent arena.
#include <cstdlib>

#include "common.cuh"

namespace dkmc {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int window_carveout() {
    static const int pct = [] { const char *e = getenv("DKMC_CARVEOUT"); int v = e ? atoi(e) : 60; return v < 0 ? 0 : (v > 100 ? 100 : v); }();
    return pct;
}

int ensure_slot(dkmc_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (ctx->slot_cap[slot] < bytes) {
        if (ctx->slot_ptr[slot]) {
            DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
            DKMC_CUDA(cudaFree(ctx->slot_ptr[slot]));
            ctx->slot_ptr[slot] = nullptr;
            ctx->slot_cap[slot] = 0;
        }
        size_t cap = bytes + bytes / 8 + 256;
        DKMC_CUDA(cudaMalloc(&ctx->slot_ptr[slot], cap));
        ctx->slot_cap[slot] = cap;
    }
    *out = ctx->slot_ptr[slot];
    return DKMC_OK;
}

// fused staging pass of dkmc_snapshot_begin: one coalesced read of the five site arrays
__global__ void __launch_bounds__(256) snapshot_stage_kernel(
    int N, const int *__restrict__ element, const int *__restrict__ charge, const double *__restrict__ pb,
    const double *__restrict__ pc, const double *__restrict__ power, int *__restrict__ s_el, int *__restrict__ s_q,
    double *__restrict__ s_pot, double *__restrict__ s_pow) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    s_el[i] = element[i];
    s_q[i] = charge[i];
    s_pot[i] = __dadd_rn(pb[i], pc[i]);   // Device.cpp:250: site_potential_boundary[i] + site_potential_charge[i]
    if (power != nullptr) s_pow[i] = power[i];
}

}  // namespace dkmc

using namespace dkmc;

extern "C" {

int dkmc_version(void) { return DKMC_VERSION; }

const char *dkmc_last_error(void) { return g_err; }

int dkmc_device_count(int *count) {
    DKMC_REQUIRE(count != nullptr, "count");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return DKMC_ERR_NO_DEVICE;
    }
    return DKMC_OK;
}

int dkmc_get_gpu_info(char *name, int name_cap, int dev) {
    DKMC_REQUIRE(name != nullptr && name_cap > 0, "name buffer");
    cudaDeviceProp prop;
    DKMC_CUDA(cudaSetDevice(dev));
    DKMC_CUDA(cudaGetDeviceProperties(&prop, dev));
    strncpy(name, prop.name, (size_t)name_cap - 1);
    name[name_cap - 1] = '\0';
    return DKMC_OK;
}

int dkmc_set_gpu(int dev) {
    DKMC_CUDA(cudaSetDevice(dev));
    return DKMC_OK;
}

int dkmc_ctx_create(dkmc_ctx **out) {
    DKMC_REQUIRE(out != nullptr, "ctx out");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        set_error("no CUDA device: the devicekmc-b200 hot path has no CPU fallback");
        return DKMC_ERR_NO_DEVICE;
    }
    dkmc_ctx *ctx = new dkmc_ctx();
    DKMC_CUDA(cudaGetDevice(&ctx->dev));
    DKMC_CUDA(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->dev));
    DKMC_CUDA(cudaMalloc(&ctx->d_layerE, 4 * kMaxLayers * sizeof(double)));
    DKMC_CUDA(cudaMemset(ctx->d_layerE, 0, 4 * kMaxLayers * sizeof(double)));
    DKMC_CUDA(cudaEventCreate(&ctx->ev_a));
    DKMC_CUDA(cudaEventCreate(&ctx->ev_b));
    DKMC_CUDA(cudaEventCreate(&ctx->ev_c));
    DKMC_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    DKMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    DKMC_CUDA(cudaEventCreate(&ctx->ev_pw0));
    DKMC_CUDA(cudaEventCreate(&ctx->ev_pw1));
    DKMC_CUDA(cudaStreamCreateWithFlags(&ctx->io_stream, cudaStreamNonBlocking));
    DKMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_snap_staged, cudaEventDisableTiming));
    DKMC_CUDA(cudaEventCreateWithFlags(&ctx->ev_snap_done, cudaEventDisableTiming));
    if (const char *e = getenv("DKMC_LEGACY_CG")) ctx->legacy_cg = atoi(e) ? 1 : 0;
    if (const char *e = getenv("DKMC_PCG_PIPELINED")) ctx->pcg_pipelined = atoi(e) ? 1 : 0;
    if (const char *e = getenv("DKMC_PW_SIDE_BPS")) { int v = atoi(e); if (v > 0) ctx->pw_side_blocks_per_sm = v; }
    if (const char *e = getenv("DKMC_PW_SHARE")) {   // experiments: "blocks_per_sm,threads" of the overlapped pairwise sum
        int b = 0, t = 0;
        if (sscanf(e, "%d,%d", &b, &t) == 2 && b >= 1 && b <= 16 && t >= 32 && t <= 256 && t % 32 == 0) {
            ctx->pw_side_blocks_per_sm = b;
            ctx->pw_side_threads = t;
        }
    }
    *out = ctx;
    return DKMC_OK;
}

int dkmc_ctx_destroy(dkmc_ctx *ctx) {
    if (!ctx) return DKMC_OK;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->io_stream) { cudaStreamSynchronize(ctx->io_stream); cudaStreamDestroy(ctx->io_stream); }
    if (ctx->ev_snap_staged) cudaEventDestroy(ctx->ev_snap_staged);
    if (ctx->ev_snap_done) cudaEventDestroy(ctx->ev_snap_done);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_pw0) cudaEventDestroy(ctx->ev_pw0);
    if (ctx->ev_pw1) cudaEventDestroy(ctx->ev_pw1);
    for (int s = 0; s < kNumSlots; ++s)
        if (ctx->slot_ptr[s]) cudaFree(ctx->slot_ptr[s]);
    if (ctx->tiling.d_tile_row) cudaFree(ctx->tiling.d_tile_row);
    free_solver_state(ctx);
    if (ctx->d_layerE) cudaFree(ctx->d_layerE);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_c) cudaEventDestroy(ctx->ev_c);
    delete ctx;
    return DKMC_OK;
}

// ---------------------------------------------------------------- snapshot
// What Device::writeSnapshot prints per site (Device.cpp:236-252) besides the static positions:
// element, potential_boundary + potential_charge (one rounded add, as on the host), power; the charge
// goes along for restarts.  One fused pass stages them (40 B read, 24 B written per site); the event
// loop may then mutate the live arrays while the staged copy drains to the host on its own stream.
int dkmc_snapshot_begin(dkmc_ctx *ctx, int N, const int *d_site_element, const int *d_site_charge,
                        const double *d_site_potential_boundary, const double *d_site_potential_charge,
                        const double *d_site_power, int *h_element, int *h_charge, double *h_potential,
                        double *h_power) {
    DKMC_REQUIRE(ctx && d_site_element && d_site_charge && d_site_potential_boundary && d_site_potential_charge &&
                 h_element && h_charge && h_potential, "null pointer");
    DKMC_REQUIRE(N > 0, "N must be positive");
    DKMC_REQUIRE((d_site_power == nullptr) == (h_power == nullptr), "d_site_power and h_power go together");
    DKMC_REQUIRE(!ctx->snap_pending, "a snapshot is already in flight: call dkmc_snapshot_wait");
    double *stage;
    int rc;
    if ((rc = ensure<double>(ctx, S_SNAP_STAGE, (size_t)3 * N + 2, &stage))) return rc;
    double *s_pot = stage, *s_pow = stage + N;
    int *s_el = reinterpret_cast<int *>(stage + 2 * (size_t)N), *s_q = s_el + N;
    DKMC_LAUNCH(ctx, snapshot_stage_kernel, ceil_div(N, 256), 256, 0, N, d_site_element, d_site_charge,
                d_site_potential_boundary, d_site_potential_charge, d_site_power, s_el, s_q, s_pot, s_pow);
    DKMC_CUDA(cudaEventRecord(ctx->ev_snap_staged, ctx->stream));
    DKMC_CUDA(cudaStreamWaitEvent(ctx->io_stream, ctx->ev_snap_staged, 0));
    DKMC_CUDA(cudaMemcpyAsync(h_element, s_el, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, ctx->io_stream));
    DKMC_CUDA(cudaMemcpyAsync(h_charge, s_q, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, ctx->io_stream));
    DKMC_CUDA(cudaMemcpyAsync(h_potential, s_pot, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->io_stream));
    if (h_power) DKMC_CUDA(cudaMemcpyAsync(h_power, s_pow, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->io_stream));
    DKMC_CUDA(cudaEventRecord(ctx->ev_snap_done, ctx->io_stream));
    // `stream` does NOT wait for the drain; the staging buffer is protected by snap_pending (a second
    // begin is refused until dkmc_snapshot_wait has seen ev_snap_done)
    ctx->snap_pending = true;
    return DKMC_OK;
}

int dkmc_snapshot_ready(dkmc_ctx *ctx, int *ready) {
    DKMC_REQUIRE(ctx && ready, "null pointer");
    *ready = 1;
    if (!ctx->snap_pending) return DKMC_OK;
    cudaError_t e = cudaEventQuery(ctx->ev_snap_done);
    if (e == cudaErrorNotReady) { *ready = 0; return DKMC_OK; }
    DKMC_CUDA(e);
    return DKMC_OK;
}

int dkmc_snapshot_wait(dkmc_ctx *ctx) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    if (!ctx->snap_pending) return DKMC_OK;
    DKMC_CUDA(cudaEventSynchronize(ctx->ev_snap_done));
    ctx->snap_pending = false;
    return DKMC_OK;
}

int dkmc_ctx_set_stream(dkmc_ctx *ctx, void *cuda_stream) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return DKMC_OK;
}

int dkmc_ctx_synchronize(dkmc_ctx *ctx) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return DKMC_OK;
}

int dkmc_ctx_launch_count(dkmc_ctx *ctx, long long *count) {
    DKMC_REQUIRE(ctx != nullptr && count != nullptr, "ctx/count");
    *count = ctx->launches;
    return DKMC_OK;
}

int dkmc_ctx_set_exact_select(dkmc_ctx *ctx, int mode) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    ctx->exact_select = mode;
    return DKMC_OK;
}

int dkmc_set_layer_energies(dkmc_ctx *ctx, int n_layers, const double *E_gen, const double *E_rec,
                            const double *E_Vdiff, const double *E_Odiff) {
    DKMC_REQUIRE(ctx != nullptr, "ctx");
    DKMC_REQUIRE(n_layers > 0 && n_layers <= kMaxLayers, "n_layers must be in 1..16");
    double host[4 * kMaxLayers] = {};
    for (int l = 0; l < n_layers; ++l) {
        host[0 * kMaxLayers + l] = E_gen[l];
        host[1 * kMaxLayers + l] = E_rec[l];
        host[2 * kMaxLayers + l] = E_Vdiff[l];
        host[3 * kMaxLayers + l] = E_Odiff[l];
    }
    DKMC_CUDA(cudaMemcpyAsync(ctx->d_layerE, host, sizeof(host), cudaMemcpyHostToDevice, ctx->stream));
    DKMC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->n_layers = n_layers;
    return DKMC_OK;
}

void dkmc_default_solver_opts(dkmc_solver_opts *o) {
    if (!o) return;
    o->rel_tol = 1e-12;
    o->max_iter = 20000;
    o->refine_rounds = 4;
    o->check_every = 32;
    o->cluster_precond = 1;
    o->refine_tol = 1e-6;
    o->est_tol = 1e-13;
    // experiments only
    if (const char *e = getenv("DKMC_REL_TOL")) { double v = atof(e); if (v > 0.0) o->rel_tol = v; }
    if (const char *e = getenv("DKMC_REFINE_TOL")) { double v = atof(e); if (v > 0.0) o->refine_tol = v; }
}

}  // extern "C"
