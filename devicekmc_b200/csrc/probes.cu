// devicekmc-b200 — FP64-pipe peak probe.  MEASURED_PEAKS.json carries HBM and bf16 tensor peaks
// only; the pairwise Coulomb kernel is bound by the FP64 FMA pipe, so its roofline denominator
// is measured here: 8 independent DFMA chains per thread, all SMs, timed with CUDA events.
#include "common.cuh"

namespace dkmc {

__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double seed, double *out) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.9999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s;  // keeps the chains alive
}

}  // namespace dkmc

using namespace dkmc;

extern "C" int dkmc_probe_fp64_tflops(dkmc_ctx *ctx, double *tflops) {
    DKMC_REQUIRE(ctx && tflops, "null pointer");
    void *buf;
    int rc;
    if ((rc = ensure_slot(ctx, S_SEL_OUT, 16, &buf))) return rc;
    const int iters = 1 << 15, threads = 256, blocks = ctx->num_sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        DKMC_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
        DKMC_LAUNCH(ctx, dfma_probe_kernel, blocks, threads, 0, iters, 1.0 + rep, static_cast<double *>(buf));
        DKMC_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
        DKMC_CUDA(cudaEventSynchronize(ctx->ev_b));
        float ms = 0;
        DKMC_CUDA(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
        double tf = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    return DKMC_OK;
}
