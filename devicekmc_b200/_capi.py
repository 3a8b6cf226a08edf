"""ctypes binding of include/dkmc.h (libdkmc_b200.so).

There is no CPU fallback: importing the package works without a GPU (so host-side logic can be
tested), but every compute entry point raises if the library is missing or no device exists.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libdkmc_b200.so")

DKMC_OK = 0
DKMC_ERR_CUDA = 1
DKMC_ERR_ARG = 2
DKMC_ERR_NOT_CONVERGED = 3
DKMC_ERR_RNG_EXHAUSTED = 4
DKMC_ERR_NO_DEVICE = 5


class DkmcError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"dkmc status {status}: {msg}")
        self.status = status


class Sparsity(C.Structure):
    _fields_ = [("m", C.c_int), ("nnz", C.c_int), ("left_nnz", C.c_int), ("right_nnz", C.c_int),
                ("d_row_ptr", C.c_void_p), ("d_col", C.c_void_p),
                ("d_left_row_ptr", C.c_void_p), ("d_left_col", C.c_void_p),
                ("d_right_row_ptr", C.c_void_p), ("d_right_col", C.c_void_p)]


class SolverOpts(C.Structure):
    _fields_ = [("rel_tol", C.c_double), ("max_iter", C.c_int), ("refine_rounds", C.c_int),
                ("check_every", C.c_int), ("cluster_precond", C.c_int), ("refine_tol", C.c_double),
                ("est_tol", C.c_double)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("refinements", C.c_int), ("rel_residual", C.c_double),
                ("assemble_ms", C.c_double), ("solve_ms", C.c_double), ("est_error", C.c_double)]


MAX_RANKS, MAX_HALO = 64, 64


class DistPlan(C.Structure):
    _fields_ = [("world", C.c_int), ("row_begin", C.c_int * MAX_RANKS), ("row_end", C.c_int * MAX_RANKS),
                ("n_send", C.c_int), ("n_recv", C.c_int),
                ("send_peer", C.c_int * MAX_HALO), ("send_begin", C.c_int * MAX_HALO), ("send_end", C.c_int * MAX_HALO),
                ("recv_peer", C.c_int * MAX_HALO), ("recv_begin", C.c_int * MAX_HALO), ("recv_end", C.c_int * MAX_HALO)]


class StepInfo(C.Structure):
    _fields_ = [("n_events", C.c_int), ("n_used", C.c_int), ("n_exact_fallbacks", C.c_int),
                ("event_time", C.c_double), ("rate_ms", C.c_double), ("loop_ms", C.c_double)]


# every symbol include/dkmc.h declares (tests check the .so exports each one)
EXPORTS = [
    "dkmc_version", "dkmc_last_error", "dkmc_get_gpu_info", "dkmc_set_gpu", "dkmc_device_count",
    "dkmc_ctx_create", "dkmc_ctx_destroy", "dkmc_ctx_set_stream", "dkmc_ctx_synchronize",
    "dkmc_ctx_launch_count", "dkmc_set_layer_energies", "dkmc_neighbor_count", "dkmc_neighbor_fill", "dkmc_neighbor_table_host", "dkmc_snapshot_begin", "dkmc_snapshot_ready", "dkmc_snapshot_wait",
    "dkmc_initialize_sparsity", "dkmc_free_sparsity", "dkmc_update_charge", "dkmc_default_solver_opts",
    "dkmc_background_potential_sparse", "dkmc_update_CB_edge_sparse", "dkmc_assemble_K", "dkmc_spmv", "dkmc_solve_cg", "dkmc_pcg_profile", "dkmc_ctx_set_legacy_cg", "dkmc_ctx_set_pcg_pipelined", "dkmc_solver_set_order", "dkmc_solver_csr", "dkmc_ctx_invalidate",
    "dkmc_poisson_gridless", "dkmc_poisson_gridless_rows", "dkmc_poisson_gridless_begin",
    "dkmc_poisson_gridless_join", "dkmc_ctx_set_pairwise_share", "dkmc_ctx_set_pairwise_cells", "dkmc_ctx_set_pairwise_cutoff", "dkmc_ctx_set_pairwise_incremental", "dkmc_pairwise_incremental_counts",
    "dkmc_pairwise_pairs_evaluated", "dkmc_ctx_set_pairwise_far_field", "dkmc_pairwise_pairs_far", "dkmc_build_event_list",
    "dkmc_inclusive_scan", "dkmc_select_event", "dkmc_execute_kmc_step", "dkmc_kmc_step_continue",
    "dkmc_ctx_set_exact_select", "dkmc_last_event_tables", "dkmc_probe_fp64_tflops", "dkmc_spmv_tile_nnz", "dkmc_dist_unique_id", "dkmc_dist_init",
    "dkmc_dist_finalize", "dkmc_dist_background_potential", "dkmc_dist_p2p_alloc", "dkmc_dist_p2p_open", "dkmc_dist_allgather_rows",
]

_lib = None


def load() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m devicekmc_b200.build` "
                "(devicekmc-b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.dkmc_last_error.restype = C.c_char_p
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        lib.dkmc_ctx_create.argtypes = [C.POINTER(vp)]
        lib.dkmc_ctx_destroy.argtypes = [vp]
        lib.dkmc_ctx_set_stream.argtypes = [vp, vp]
        lib.dkmc_ctx_synchronize.argtypes = [vp]
        lib.dkmc_ctx_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
        lib.dkmc_ctx_set_exact_select.argtypes = [vp, ci]
        lib.dkmc_set_layer_energies.argtypes = [vp, ci, vp, vp, vp, vp]
        lib.dkmc_neighbor_count.argtypes = [vp, ci, vp, vp, vp, vp, ci, cd, C.POINTER(ci)]
        lib.dkmc_neighbor_fill.argtypes = [vp, ci, vp, vp, vp, vp, ci, cd, ci, vp]
        lib.dkmc_snapshot_begin.argtypes = [vp, ci] + [vp] * 9
        lib.dkmc_snapshot_ready.argtypes = [vp, C.POINTER(ci)]
        lib.dkmc_snapshot_wait.argtypes = [vp]
        lib.dkmc_neighbor_table_host.argtypes = [vp, ci, vp, vp, vp, vp, ci, cd, C.POINTER(ci), vp]
        lib.dkmc_initialize_sparsity.argtypes = [vp, ci, ci, vp, ci, ci, C.POINTER(Sparsity)]
        lib.dkmc_free_sparsity.argtypes = [vp, C.POINTER(Sparsity)]
        lib.dkmc_update_charge.argtypes = [vp, vp, vp, vp, ci, ci, vp, ci]
        lib.dkmc_default_solver_opts.argtypes = [C.POINTER(SolverOpts)]
        lib.dkmc_default_solver_opts.restype = None
        lib.dkmc_background_potential_sparse.argtypes = [vp, C.POINTER(Sparsity), ci, ci, vp, ci, ci, cd, cd, cd,
                                                         vp, vp, vp, ci, vp, C.POINTER(SolverOpts),
                                                         C.POINTER(SolveInfo)]
        lib.dkmc_update_CB_edge_sparse.argtypes = [vp, C.POINTER(Sparsity), ci, ci, ci, cd, cd, cd, cd, vp, vp, ci, vp,
                                                   C.POINTER(SolverOpts), C.POINTER(SolveInfo)]
        lib.dkmc_assemble_K.argtypes = [vp, C.POINTER(Sparsity), ci, ci, ci, cd, cd, cd, vp, vp, vp, ci, vp, vp]
        lib.dkmc_spmv.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
        lib.dkmc_pcg_profile.argtypes = [vp, vp]
        lib.dkmc_ctx_set_legacy_cg.argtypes = [vp, ci]
        lib.dkmc_ctx_set_pcg_pipelined.argtypes = [vp, ci]
        lib.dkmc_solver_set_order.argtypes = [vp, C.POINTER(Sparsity), vp]
        lib.dkmc_solver_csr.argtypes = [vp, C.POINTER(Sparsity), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        lib.dkmc_solve_cg.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, C.POINTER(SolverOpts), C.POINTER(SolveInfo)]
        lib.dkmc_poisson_gridless.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.dkmc_poisson_gridless_rows.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, ci, ci, vp]
        lib.dkmc_poisson_gridless_begin.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, ci, ci, vp]
        lib.dkmc_poisson_gridless_join.argtypes = [vp, C.POINTER(cd)]
        lib.dkmc_ctx_set_pairwise_share.argtypes = [vp, ci, ci]
        lib.dkmc_ctx_set_pairwise_cells.argtypes = [vp, ci]
        lib.dkmc_ctx_set_pairwise_cutoff.argtypes = [vp, cd]
        lib.dkmc_ctx_set_pairwise_incremental.argtypes = [vp, ci]
        lib.dkmc_pairwise_incremental_counts.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        lib.dkmc_pairwise_pairs_evaluated.argtypes = [vp, C.POINTER(C.c_longlong)]
        lib.dkmc_pairwise_pairs_far.argtypes = [vp, C.POINTER(C.c_longlong)]
        lib.dkmc_ctx_set_pairwise_far_field.argtypes = [vp, C.c_int]
        lib.dkmc_build_event_list.argtypes = [vp, ci, ci, vp, vp, vp, ci] + [vp] * 13
        lib.dkmc_inclusive_scan.argtypes = [vp, C.c_longlong, vp, vp]
        lib.dkmc_select_event.argtypes = [vp, C.c_longlong, vp, cd, C.POINTER(C.c_longlong), C.POINTER(cd)]
        lib.dkmc_execute_kmc_step.argtypes = [vp, ci, ci, vp, vp, vp, ci] + [vp] * 11 + [vp, ci, vp, ci,
                                                                                       C.POINTER(StepInfo)]
        lib.dkmc_kmc_step_continue.argtypes = [vp, vp, ci, vp, ci, C.POINTER(StepInfo)]
        lib.dkmc_last_event_tables.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        lib.dkmc_dist_unique_id.argtypes = [C.c_char_p]
        lib.dkmc_dist_init.argtypes = [vp, ci, ci, C.c_char_p]
        lib.dkmc_dist_finalize.argtypes = [vp]
        lib.dkmc_dist_p2p_alloc.argtypes = [vp, ci, C.c_char_p]
        lib.dkmc_dist_p2p_open.argtypes = [vp, C.c_char_p]
        lib.dkmc_dist_allgather_rows.argtypes = [vp, vp, ci, vp, vp]
        lib.dkmc_dist_background_potential.argtypes = [vp, C.POINTER(Sparsity), ci, ci, ci, cd, cd, cd, vp, vp, vp, ci, vp,
                                                       C.POINTER(DistPlan), C.POINTER(SolverOpts), C.POINTER(SolveInfo)]
        lib.dkmc_probe_fp64_tflops.argtypes = [vp, C.POINTER(cd)]
        lib.dkmc_get_gpu_info.argtypes = [C.c_char_p, ci, ci]
        lib.dkmc_set_gpu.argtypes = [ci]
        lib.dkmc_device_count.argtypes = [C.POINTER(ci)]
        _lib = lib
    return _lib


def check(status: int, allow=()):
    if status != DKMC_OK and status not in allow:
        raise DkmcError(status, load().dkmc_last_error().decode(errors="replace"))
    return status
