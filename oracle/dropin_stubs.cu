// TEST INFRASTRUCTURE ONLY — link-time stubs for the reference GPU entry points that are OUT OF
// SCOPE of devicekmc-b200 (current solver, heat solver, dense-LU potential;
// SURVEY.md §2 rows 14/15, §8f).  They let the unmodified reference host (kmc_main.cpp ...) link
// against our shim for the drop-in check; reaching one of them aborts with a clear message.
#include "gpu_solvers.h"

#include <cstdio>
#include <cstdlib>

static void out_of_scope(const char *name) {
    fprintf(stderr, "devicekmc-b200 drop-in: %s is outside the field-and-rate hot path (run with solve_current = 0, "
                    "solve_heating_* = 0)\n", name);
    abort();
}

extern "C" {
void background_potential_gpu(cusolverDnHandle_t, GPUBuffers &, const int, const int, const int, const double, const int,
                              const double, const double, const double, const int, int) {
    out_of_scope("background_potential_gpu (dense LU variant)");
}
void update_power_gpu(cublasHandle_t, cusolverDnHandle_t, GPUBuffers &, const int, const int, const int, const double,
                      const int, const double, const double, const double, const double, const double, const double,
                      const double, const double, int, double *, const bool, const bool, const double) {
    out_of_scope("update_power_gpu");
}
void update_power_gpu_sparse(cublasHandle_t, cusolverDnHandle_t, GPUBuffers &, const int, const int, const int,
                             const double, const int, const double, const double, const double, const double,
                             const double, const double, const double, const double, int, double *, const bool,
                             const bool, const double) {
    out_of_scope("update_power_gpu_sparse");
}
void update_power_gpu_split(cublasHandle_t, cusolverDnHandle_t, GPUBuffers &, const int, const int, const int,
                            const double, const int, const double, const double, const double, const double,
                            const double, const double, const double, const double, int, double *, const bool,
                            const bool, const double) {
    out_of_scope("update_power_gpu_split");
}
void update_temperatureglobal_gpu(const double *, double *, const int, const double, const double, const double,
                                  const double, const double) {
    out_of_scope("update_temperatureglobal_gpu");
}
}
