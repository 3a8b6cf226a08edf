"""Distributed (slab-partitioned) PCG: builds the row/halo plan and drives
dkmc_dist_background_potential.  NCCL communicator bootstrap goes through torch.distributed."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi, slab
from ._capi import DistPlan, SolveInfo, check


def make_plan(row_ptr: np.ndarray, col: np.ndarray, world: int, rank: int, tile: int) -> DistPlan:
    rows = slab.split_rows_by_nnz(row_ptr, world, tile)
    halos = slab.halo_ranges(row_ptr, col, rows)
    recv, send = slab.halo_plan(rows, halos)
    if world > _capi.MAX_RANKS or len(recv[rank]) > _capi.MAX_HALO or len(send[rank]) > _capi.MAX_HALO:
        raise ValueError("halo plan exceeds the C-ABI limits")
    pl = DistPlan()
    pl.world = world
    for r, (a, b) in enumerate(rows):
        pl.row_begin[r], pl.row_end[r] = a, b
    pl.n_send, pl.n_recv = len(send[rank]), len(recv[rank])
    for k, (peer, b, e) in enumerate(send[rank]):
        pl.send_peer[k], pl.send_begin[k], pl.send_end[k] = peer, b, e
    for k, (peer, b, e) in enumerate(recv[rank]):
        pl.recv_peer[k], pl.recv_begin[k], pl.recv_end[k] = peer, b, e
    return pl


def _device_ints(ptr: int, n: int) -> np.ndarray:
    """device int32 array -> numpy (one-off, at set-up)"""
    import torch
    out = np.empty(n, np.int32)
    torch.cuda.synchronize()
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    err = rt.cudaMemcpy(out.ctypes.data, ptr, out.nbytes, 2)
    if err != 0:
        raise RuntimeError(f"cudaMemcpy failed: {err}")
    return out


class DistributedSolver:
    def __init__(self, sim: "slab.SlabSim", p2p: bool = True):
        import torch.distributed as dist
        self.sim = sim
        lib = sim.dev.ctx.lib
        idbuf = C.create_string_buffer(128)
        if sim.rank == 0:
            check(lib.dkmc_dist_unique_id(idbuf))
        box = [bytes(idbuf.raw)]
        dist.broadcast_object_list(box, src=0)
        check(lib.dkmc_dist_init(sim.dev.ctx.h, sim.rank, sim.world, box[0]))
        sp = sim.sp
        # the CSR structure the solver works on: the caller's, or its image in the registered internal row order
        rp_ptr, col_ptr = C.c_void_p(), C.c_void_p()
        check(lib.dkmc_solver_csr(sim.dev.ctx.h, C.byref(sp), C.byref(rp_ptr), C.byref(col_ptr), None))
        row_ptr = _device_ints(rp_ptr.value, sp.m + 1)
        col = _device_ints(col_ptr.value, sp.nnz)
        self.plan = make_plan(row_ptr, col, sim.world, sim.rank, lib.dkmc_spmv_tile_nnz())
        self.p2p = False
        if p2p and os.environ.get("DKMC_P2P", "1") != "0":
            # peer-memory windows: CUDA IPC handles all-gathered through torch.distributed
            hbuf = C.create_string_buffer(64)
            ok = lib.dkmc_dist_p2p_alloc(sim.dev.ctx.h, sim.dev.N, hbuf) == 0   # room for a per-site array
            handles = [None] * sim.world
            dist.all_gather_object(handles, bytes(hbuf.raw) if ok else b"")
            if all(len(h) == 64 for h in handles):
                ok = lib.dkmc_dist_p2p_open(sim.dev.ctx.h, b"".join(handles)) == 0
            else:
                ok = False
            flags = [None] * sim.world
            dist.all_gather_object(flags, bool(ok))
            self.p2p = all(flags)   # every rank must have mapped every window

    def solve(self, Vd: float, info: SolveInfo, opts=None):
        s = self.sim
        dev, buf, p = s.dev, s.buf, s.p
        st = dev.ctx.lib.dkmc_dist_background_potential(
            dev.ctx.h, C.byref(s.sp), dev.N, s.nc, s.nc, float(Vd), float(p.high_G), float(p.low_G),
            buf.site_element.data_ptr(), buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), buf.num_metal_types_,
            buf.site_potential_boundary.data_ptr(), C.byref(self.plan), C.byref(opts) if opts is not None else None,
            C.byref(info))
        check(st, allow=(_capi.DKMC_ERR_NOT_CONVERGED,))
        return st

    def close(self):
        self.sim.dev.ctx.lib.dkmc_dist_finalize(self.sim.dev.ctx.h)
