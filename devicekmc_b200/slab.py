"""Slab-partitioned multi-GPU execution of the hot path: one process per GPU (torchrun),
torch.distributed/NCCL for the plumbing.

Sites are ordered x-major (structures.tile_device), so a contiguous index range is a spatial slab.

  stage                     partition                          exchange (NCCL over NVLink)
  ------------------------  ---------------------------------  --------------------------------------
  charge state machine      replicated (integer, 30 us)        none
  K assembly + PCG          interior rows split by nnz tiles   halo of p (contiguous index ranges to the
                                                               two x-neighbours), 2 small all-reduces
                                                               per iteration, all-gather of phi_b
  pairwise Coulomb sum      target rows [i0, i1) per rank,     all-gather of phi_c (8 B per site)
                            ALL charged sources (replicated)
  rate table + event loop   replicated: every rank runs the    none — events are serial by construction;
                            identical loop on the same         a per-event all-gather of rate sums costs
                            mt19937 uniforms                   more than the whole per-event work

The host-side planning (row/tile ranges, halo widths, owner of a global rate target) lives here
and is covered by world_size-2 gloo tests on the CPU (tests/test_slab_gloo.py).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import List, Tuple

import numpy as np


# ---------------------------------------------------------------------------- planning (pure host logic)
def split_even(n: int, world: int) -> List[Tuple[int, int]]:
    """contiguous [begin, end) ranges of n items over `world` ranks, sizes differing by <= 1"""
    base, rem = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def chunk_padded(n: int, world: int) -> int:
    """equal chunk length c with world * c >= n (for all_gather_into_tensor)"""
    return (n + world - 1) // world


def split_rows_by_nnz(row_ptr: np.ndarray, world: int, tile: int) -> List[Tuple[int, int]]:
    """interior rows split so that every rank gets the same number of `tile`-nnz SpMV tiles
    (tiles are defined by row START, exactly as the CUDA kernel's tiling does)"""
    m = len(row_ptr) - 1
    nnz = int(row_ptr[m])
    ntiles = max(1, (nnz + tile - 1) // tile)
    tile_row = np.searchsorted(row_ptr[:m], np.arange(ntiles) * tile, side="left")
    tile_row = np.append(tile_row, m)
    out = []
    for t0, t1 in split_even(ntiles, world):
        out.append((int(tile_row[t0]), int(tile_row[t1])))
    return out


def halo_ranges(row_ptr: np.ndarray, col: np.ndarray, rows: List[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """per rank: [lo, hi) = the smallest contiguous column range its rows touch"""
    out = []
    for ra, rb in rows:
        if rb <= ra:
            out.append((ra, ra))
            continue
        c = col[row_ptr[ra]:row_ptr[rb]]
        out.append((int(min(c.min(), ra)), int(max(c.max() + 1, rb))))
    return out


def halo_plan(rows: List[Tuple[int, int]], halos: List[Tuple[int, int]]):
    """For every rank the list of (peer, begin, end) segments to RECEIVE (columns it needs that
    another rank owns).  With x-major ordering these are one segment from each x-neighbour; for an
    arbitrary ordering the plan degenerates to (almost) an all-gather."""
    world = len(rows)
    recv = [[] for _ in range(world)]
    for r in range(world):
        lo, hi = halos[r]
        for q in range(world):
            if q == r:
                continue
            b, e = max(lo, rows[q][0]), min(hi, rows[q][1])
            if b < e:
                recv[r].append((q, b, e))
    send = [[] for _ in range(world)]
    for r in range(world):
        for q, b, e in recv[r]:
            send[q].append((r, b, e))
    return recv, send


def owner_of_target(rank_sums: np.ndarray, u: float):
    """the global rate-sum scan that picks the owning rank: exclusive offsets of the per-rank rate
    sums, first rank whose inclusive sum exceeds u * total.  Returns (rank, local target, total)."""
    incl = np.cumsum(rank_sums)
    total = float(incl[-1])
    number = u * total
    r = int(np.searchsorted(incl, number, side="right"))
    if r >= len(rank_sums):
        return -1, 0.0, total
    return r, number - (float(incl[r - 1]) if r > 0 else 0.0), total


# ---------------------------------------------------------------------------- GPU execution
class SlabSim:
    """One rank's view of the slab-partitioned simulation."""

    def __init__(self, arrays, p, rank: int, world: int, distributed_cg: bool = True):
        import torch
        import torch.distributed as dist
        import devicekmc_b200 as D
        self.torch, self.dist, self.rank, self.world, self.p = torch, dist, rank, world, p
        el, x, y, z = arrays
        self.dev = D.Device([], p, arrays=(el, x, y, z))
        self.sim = D.KMCProcess(self.dev, p.freq)
        self.buf = D.GPUBuffers(self.sim.layers, self.sim.site_layer, self.sim.freq, self.dev, p.metals)
        N = self.dev.N
        self.nc = p.num_atoms_contact
        # phi_c lives in a padded buffer so that the all-gather has equal chunks
        self.chunk = chunk_padded(N, world)
        self._pc_full = torch.zeros(self.chunk * world, dtype=torch.float64, device="cuda")
        self.buf.site_potential_charge = self._pc_full[:N]
        # the other per-step arrays likewise (slab-wise host <-> device syncs + one all-gather, sync_slab_*)
        self._full = {"site_potential_charge": self._pc_full}
        if world > 1:
            for name in self.buf._SYNCED:
                if name in self._full:
                    continue
                old = getattr(self.buf, name)
                full = torch.zeros(self.chunk * world, dtype=old.dtype, device="cuda")
                full[:N].copy_(old)
                self._full[name] = full
                setattr(self.buf, name, full[:N])
        self.i0 = min(N, rank * self.chunk)
        self.i1 = min(N, (rank + 1) * self.chunk)
        self._rows_b = np.array([min(N, r * self.chunk) for r in range(world)], np.int32)
        self._rows_e = np.array([min(N, (r + 1) * self.chunk) for r in range(world)], np.int32)
        self.buf.sync_HostToGPU(self.dev)
        self.sp = self.buf.sparsity(self.nc, self.nc)
        # the pairwise kernel's share of each SM while it runs beside the CG: the fewer targets a rank
        # owns, the smaller the share it needs to finish before the CG does
        # (measured at 1 M sites, round 2: with two CTAs of 128 threads per SM the sum ends before the CG at every
        # N >= 2 and the CG is no slower than beside one CTA — its iterations are latency-bound there)
        share = (3, 128) if world == 1 else (2, 128)
        if os.environ.get("DKMC_PW_SHARE"):     # experiments: "blocks_per_sm,threads"
            share = tuple(int(v) for v in os.environ["DKMC_PW_SHARE"].split(","))
        from ._capi import check
        check(self.dev.ctx.lib.dkmc_ctx_set_pairwise_share(self.dev.ctx.h, share[0], share[1]))
        self._share = list(share)
        # the share follows the measured times (strong scaling shrinks the sum with N, weak scaling does not): one more
        # CTA per SM when the sum ends after the CG, one fewer when it ends in less than half the CG's time
        self._share_auto = not os.environ.get("DKMC_PW_SHARE") and not os.environ.get("DKMC_PW_SERIAL")
        self.dcg = None
        if distributed_cg and world > 1:
            from . import _dist
            self.dcg = _dist.DistributedSolver(self)
            if not self.dcg.p2p and os.environ.get("DKMC_P2P", "1") != "0":
                # no peer access between these GPUs: three NCCL calls per CG iteration cost more than the
                # partition saves at this size, so every rank solves the whole system instead
                self.dcg.close()
                self.dcg = None

    def sync_slab_HostToGPU(self):
        """gpu_buffers.cpp:10-37 for a host that is partitioned like the device: every rank uploads only the rows of
        its slab (N / world sites instead of all N through the one host's PCIe lanes); one all-gather per array over
        NVLink completes the replicated device arrays the event loop needs."""
        if self.world == 1:
            return self.buf.sync_HostToGPU(self.dev)
        a, b = self.i0, self.i1
        for name in self.buf._SYNCED:
            full = self._full[name]
            if b > a:
                full[a:b].copy_(self.buf._pinned(self.dev, name)[a:b], non_blocking=True)
            mine = full[self.rank * self.chunk:(self.rank + 1) * self.chunk].clone()
            self.dist.all_gather_into_tensor(full, mine)
        self.buf.T_bg.fill_(self.dev.T_bg)

    def sync_slab_GPUToHost(self):
        """gpu_buffers.cpp:39-55, slab-wise: the host arrays of this rank are current in its slab's rows"""
        if self.world == 1:
            return self.buf.sync_GPUToHost(self.dev)
        a, b = self.i0, self.i1
        if b > a:
            for name in self.buf._SYNCED:
                self.buf._pinned(self.dev, name)[a:b].copy_(self._full[name][a:b], non_blocking=True)
        self.torch.cuda.current_stream().synchronize()

    def slab_bytes(self) -> int:
        return (self.i1 - self.i0) * (4 + 4 + 8 + 8 + 8)

    def step(self, Vd: float, record_events: int = 0):
        import devicekmc_b200 as D
        from . import _capi
        from ._capi import SolveInfo, check
        dev, buf, p, lib = self.dev, self.buf, self.p, self.dev.ctx.lib
        dev.updateCharge(buf, p.metals)
        info = SolveInfo()
        t0 = time.perf_counter()
        # pairwise: my target rows against all charged sources, on the side stream while the CG runs
        pw_args = (dev.ctx.h, dev.pbc, dev.N, buf.lattice.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(),
                   buf.site_x.data_ptr(), buf.site_y.data_ptr(), buf.site_z.data_ptr(), buf.site_charge.data_ptr(),
                   self.i0, self.i1, self._pc_full.data_ptr())
        if self.i1 > self.i0:
            check(lib.dkmc_poisson_gridless_begin(*pw_args))
        pw_ms = C.c_double(0.0)
        if os.environ.get("DKMC_PW_SERIAL"):    # experiment: the pairwise sum BEFORE the CG instead of beside it
            check(lib.dkmc_poisson_gridless_join(dev.ctx.h, C.byref(pw_ms)))
        if self.dcg is not None:
            st = self.dcg.solve(Vd, info)
        else:
            st = lib.dkmc_background_potential_sparse(
                dev.ctx.h, C.byref(self.sp), dev.N, buf.nn_, buf.neigh_idx.data_ptr(), self.nc, self.nc, float(Vd),
                float(p.high_G), float(p.low_G), buf.site_element.data_ptr(), buf.site_charge.data_ptr(),
                buf.metal_types.data_ptr(), buf.num_metal_types_, buf.site_potential_boundary.data_ptr(), None,
                C.byref(info))
            check(st, allow=(3,))
        if not os.environ.get("DKMC_PW_SERIAL"):
            check(lib.dkmc_poisson_gridless_join(dev.ctx.h, C.byref(pw_ms)))
        if self.world > 1:
            if self.dcg is not None and self.dcg.p2p:
                # every rank pulls the other ranks' target rows from their peer windows (no NCCL call)
                check(lib.dkmc_dist_allgather_rows(dev.ctx.h, self._pc_full.data_ptr(), dev.N,
                                                   self._rows_b.ctypes.data, self._rows_e.ctypes.data))
            else:
                mine = self._pc_full[self.rank * self.chunk:(self.rank + 1) * self.chunk].clone()
                self.dist.all_gather_into_tensor(self._pc_full, mine)
        t = self.sim.executeKMCStep(buf, dev, record_events=record_events)
        if self._share_auto and info.iterations > 0:
            want = self._share[0]
            if pw_ms.value > 1.05 * info.solve_ms:
                want = min(3, want + 1)
            elif pw_ms.value < 0.5 * info.solve_ms:
                want = max(1 if self.world > 1 else 2, want - 1)
            # (every rank decides for itself: the ranks' grids need not agree, and a KMC step stays free of NCCL calls)
            if want != self._share[0]:
                self._share[0] = want
                check(lib.dkmc_ctx_set_pairwise_share(dev.ctx.h, want, self._share[1]))
        return {"cg_iterations": info.iterations, "pairwise_share": self._share[0], "cg_converged": st == _capi.DKMC_OK, "cg_est_error": info.est_error,
                "solve_ms": info.solve_ms, "assemble_ms": info.assemble_ms,
                "pairwise_ms": pw_ms.value,
                "events": self.sim.last_info.n_events, "fallbacks": self.sim.last_info.n_exact_fallbacks,
                "loop_ms": self.sim.last_info.loop_ms, "rate_ms": self.sim.last_info.rate_ms, "step_time": t}


def iv_ramp(points: int = 200, v_max: float = 4.0) -> np.ndarray:
    """the SET/RESET sweep of BASELINE config 5: 0 -> v_max -> 0 V in `points` bias points"""
    up = points // 2
    return np.concatenate([np.linspace(0.0, v_max, up, endpoint=False), np.linspace(v_max, 0.0, points - up)])


def bias_loop(s: "SlabSim", V_switch, t_switch, max_steps_per_point=None, max_steps=None, scale_warm_start=True, on_step=None,
              record_events: int = 0):
    """The reference's bias-point loop (kmc_main.cpp:136-279) over a SlabSim (1 or N ranks): for every bias point
    host -> device sync, then KMC steps (charge, potential at Vd, events) until kmc_time reaches t_switch, then
    device -> host sync (where the reference writes its snapshot).  max_steps_per_point / max_steps bound the work
    (the reference itself only stops when the physical time has elapsed: at a bias where the rates are high that
    is ~1e11 steps per point).  scale_warm_start: the boundary potential is linear in Vd while the structure does
    not change, so the previous solution is rescaled by Vd / Vd_prev before it starts the CG (the reference starts
    from the unscaled previous potential, potential_solver_gpu.cu:754).  Returns the per-step stats."""
    stats, total = [], 0
    t_switch = np.broadcast_to(np.asarray(t_switch, float), (len(V_switch),))
    vd_prev = None
    for k, Vd in enumerate(V_switch):
        s.buf.sync_HostToGPU(s.dev)                                     # kmc_main.cpp:172
        kmc_time, count = 0.0, 0
        while kmc_time < t_switch[k]:
            if scale_warm_start and vd_prev not in (None, 0.0) and vd_prev != Vd:
                s.buf.site_potential_boundary.mul_(float(Vd) / float(vd_prev))
            vd_prev = float(Vd)
            st = s.step(float(Vd), record_events=record_events)
            st["Vd"], st["bias_point"] = float(Vd), k
            stats.append(st)
            if on_step:
                on_step(st)
            kmc_time += st["step_time"]
            count += 1; total += 1
            if (max_steps_per_point and count >= max_steps_per_point) or (max_steps and total >= max_steps):
                break
        s.buf.sync_GPUToHost(s.dev)                                     # kmc_main.cpp:282-286
        if max_steps and total >= max_steps:
            break
    return stats


def multi_vs_single_step(multi: "SlabSim", arrays, p, Vd: float) -> dict:
    """One step from multi's current state through the distributed path and through a single-GPU SlabSim built
    on this rank (same state, same KMC random stream); returns the differences."""
    import torch
    single = SlabSim(arrays, p, 0, 1, distributed_cg=False)
    for name in ("site_element", "site_charge", "site_potential_boundary", "site_potential_charge"):
        getattr(single.buf, name).copy_(getattr(multi.buf, name))
    single.sim.random_generator._bg.state = multi.sim.random_generator._bg.state
    a = multi.step(Vd, record_events=1 << 16)
    ev_m = multi.sim.last_events.copy()
    b = single.step(Vd, record_events=1 << 16)
    ev_s = single.sim.last_events.copy()
    rel = lambda u, v: float((u - v).abs().max() / v.abs().max())
    return {"phi_b_rel": rel(multi.buf.site_potential_boundary, single.buf.site_potential_boundary),
            "phi_c_rel": rel(multi.buf.site_potential_charge, single.buf.site_potential_charge),
            "elements_equal": bool(torch.equal(multi.buf.site_element, single.buf.site_element)),
            "charges_equal": bool(torch.equal(multi.buf.site_charge, single.buf.site_charge)),
            "events_equal": bool(a["events"] == b["events"] and np.array_equal(ev_m, ev_s)), "events": int(a["events"]),
            "cg_converged": bool(a["cg_converged"] and b["cg_converged"])}


def bench_multi_gpu(args, metric: str, unit: str):
    """bench.py --gpus N under torchrun: strong scaling of one KMC step on the same workload"""
    import torch
    import torch.distributed as dist
    import bench
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    el, x, y, z, lat, nc, p = bench.workload(args.workload)
    el = bench.substoichiometric(el, p)
    s = SlabSim((el, x, y, z), p, rank, world, distributed_cg=not getattr(args, 'replicated_cg', False))
    stats = []
    sampler = bench.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        s.step(args.vd)
    launches0 = s.dev.ctx.launch_count()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ckpt = bench.StateCheckpoint(s.buf, s.sim)     # the e2e run below times the same steps of the trajectory
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stats.append(s.step(args.vd))
    e1.record(); e1.synchronize()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # device time, max over ranks
    launches = s.dev.ctx.launch_count() - launches0
    prof = None
    if os.environ.get("DKMC_PCG_PROF"):    # in-kernel phase profile of the persistent PCG (rank 0's CTA 0), dev aid
        buf8 = (C.c_double * 8)()
        s.dev.ctx.lib.dkmc_pcg_profile(s.dev.ctx.h, buf8)
        it = max(buf8[6], 1.0)
        prof = {"us_per_iteration": {"vector": buf8[1] / it / 1e3, "barrier_halo": buf8[2] / it / 1e3, "spmv_tiles": buf8[3] / it / 1e3,
                                     "cluster_rows": buf8[4] / it / 1e3, "barrier_reduce": buf8[5] / it / 1e3},
                "setup_us_per_solve": buf8[0] / max(buf8[7], 1.0) / 1e3, "iterations": buf8[6], "solves": buf8[7],
                "note": "profiling kernels carry extra registers: the timings of such a run are not bench values"}
        allp = [None] * world
        dist.all_gather_object(allp, prof["us_per_iteration"])
        prof["per_rank"] = {k: [round(q[k], 1) for q in allp] for k in allp[0]}
    # e2e: host buffers in and out every step, on every rank
    ckpt.restore(s.buf, s.sim, s.dev)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        s.sync_slab_HostToGPU()
        s.step(args.vd)
        s.sync_slab_GPUToHost()
    e1.record(); e1.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    # all ranks must hold the same state, bit for bit (replicated event loop on identical inputs)
    import hashlib
    digest = hashlib.sha256(s.buf.site_element.cpu().numpy().tobytes() + s.buf.site_charge.cpu().numpy().tobytes()).hexdigest()
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    consistent = len(set(digests)) == 1
    # parity, driver-visible: one more step through the distributed path AND through the single-GPU path
    # (replicated on every rank) from the same state; integers identical, potentials within 1e-10
    parity = multi_vs_single_step(s, (el, x, y, z), p, args.vd)
    worst = [None] * world
    dist.all_gather_object(worst, parity)
    parity = {"phi_b_rel": max(w["phi_b_rel"] for w in worst), "phi_c_rel": max(w["phi_c_rel"] for w in worst),
              "elements_equal": all(w["elements_equal"] for w in worst), "charges_equal": all(w["charges_equal"] for w in worst),
              "events_equal": all(w["events_equal"] for w in worst), "events": worst[0]["events"],
              "cg_converged": all(w["cg_converged"] for w in worst), "tolerance": 1e-10,
              "what": "one extra step after the timed region: distributed path vs single-GPU path on every rank, same state and random stream"}
    parity["ok"] = bool(parity["phi_b_rel"] <= 1e-10 and parity["phi_c_rel"] <= 1e-10 and parity["elements_equal"] and
                        parity["charges_equal"] and parity["events_equal"] and parity["cg_converged"])
    if rank == 0:
        clocks = sampler.stop()
        value = args.steps / (ms.item() * 1e-3)
        med = lambda k: float(np.median([t[k] for t in stats]))
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms.item() / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": bench.workload_config(args.workload, s.dev.N, s.buf.nn_, args.vd),
                "arm": {"solver_order": "internal x-major cell order (input in the reference's site order)" if s.buf.solver_order_applied else "caller's (already x-major)",
                        "partition": f"x-slabs over {world} ranks: rows by nnz tiles (CG), targets by site (pairwise)",
                        "cg": ("slab-partitioned persistent PCG, exchange over NVLink peer memory (CUDA IPC)" if s.dcg is not None and s.dcg.p2p
                               else "slab-partitioned, NCCL" if s.dcg is not None else "replicated"),
                        "pairwise_share_ctas_per_sm": stats[-1].get("pairwise_share")},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": args.steps / (ms2.item() * 1e-3), "unit": unit, "h2d_bytes_per_step": s.slab_bytes() * world,
                        "d2h_bytes_per_step": s.slab_bytes() * world,
                        "note": "host partitioned like the device: every rank moves its slab's rows (h2d/d2h bytes are the "
                                "job's totals) and one all-gather per array over NVLink completes the replicated device arrays"},
                "stage_ms": {"cg_solve": med("solve_ms"), "pairwise_concurrent": med("pairwise_ms"),
                             "assemble": med("assemble_ms"), "rate_table": med("rate_ms"),
                             "event_loop": med("loop_ms"),
                             "cg_us_per_iteration": 1e3 * float(np.sum([t["solve_ms"] for t in stats])) /
                                                    max(1, int(np.sum([t["cg_iterations"] for t in stats])))},
                "per_step": {"events": [t["events"] for t in stats], "cg_iterations": [t["cg_iterations"] for t in stats]},
                "pcg_profile": prof, "ranks_consistent": consistent, "ranks_state_sha256": digests[0][:16], "parity": parity,
                "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


# weak scaling of BASELINE config 5: ~0.5 M sites per GPU (tiles of the 2.5 nm cell: ny x nz)
RAMP_TILES = {1: (7, 8), 2: (10, 11), 4: (14, 15), 8: (21, 20)}


def bench_ramp(args, metric: str, unit: str):
    """bench.py --ramp [--gpus N]: the I-V sweep of BASELINE config 5 through the reference's bias-point loop
    (bias_loop), weak scaling: the device grows with N (RAMP_TILES, ~0.5 M sites per GPU; N = 8 is the 4 M-site
    device).  One KMC step per bias point (--ramp-steps-per-point), W untimed points, then K timed ones starting
    at --ramp-start; value = KMC steps per second of the whole job."""
    import torch
    import torch.distributed as dist
    import bench
    from . import structures as S
    from .host import KMCParameters
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29571")
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ny, nz = RAMP_TILES[world]
    el, x, y, z, lat, nc = S.tile_device(ny, nz)
    p = KMCParameters(lattice=tuple(lat), num_atoms_contact=nc, num_atoms_first_layer=nc)
    el = bench.substoichiometric(el, p)
    s = SlabSim((el, x, y, z), p, rank, world)
    ramp = iv_ramp(args.ramp_points, args.ramp_vmax)
    a, w, k = args.ramp_start, args.warmup, args.steps
    sampler = bench.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    bias_loop(s, ramp[a:a + w], 1e-3, max_steps_per_point=args.ramp_steps_per_point)
    launches0 = s.dev.ctx.launch_count()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = bias_loop(s, ramp[a + w:a + w + k], 1e-3, max_steps_per_point=args.ramp_steps_per_point)
    e1.record(); e1.synchronize()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = s.dev.ctx.launch_count() - launches0
    import hashlib
    digest = hashlib.sha256(s.buf.site_element.cpu().numpy().tobytes() + s.buf.site_charge.cpu().numpy().tobytes()).hexdigest()
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    if rank == 0:
        clocks = sampler.stop()
        n = len(stats)
        med = lambda key: float(np.median([t[key] for t in stats]))
        line = {"metric": metric, "value": n / (ms.item() * 1e-3), "unit": unit, "n_gpus": world, "steps": n, "warmup": w,
                "ms_per_step": ms.item() / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"iv_ramp_{ny}x{nz}", "sites": s.dev.N, "sites_per_gpu": s.dev.N // world, "nn": s.buf.nn_,
                           "ramp": f"0 -> {args.ramp_vmax} -> 0 V in {args.ramp_points} bias points, t_switch 1e-3 s, through the "
                                   f"reference's bias-point loop (kmc_main.cpp:136-279); bias points {a + w} .. {a + w + k - 1}",
                           "Vd_first_last": [float(stats[0]["Vd"]), float(stats[-1]["Vd"])],
                           "steps_per_bias_point": args.ramp_steps_per_point, "l2": "inputs larger than L2"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": n / (ms.item() * 1e-3), "unit": unit,
                        "h2d_bytes_per_step": s.buf.h2d_bytes() // args.ramp_steps_per_point,
                        "d2h_bytes_per_step": s.buf.d2h_bytes() // args.ramp_steps_per_point,
                        "note": "the bias-point loop itself moves the site arrays host -> device and back at every bias point "
                                "(kmc_main.cpp:172,282): the timed region includes those copies"},
                "stage_ms": {"cg_solve": med("solve_ms"), "pairwise_concurrent": med("pairwise_ms"), "assemble": med("assemble_ms"),
                             "rate_table": med("rate_ms"), "event_loop": med("loop_ms")},
                "per_step": {"Vd": [round(t["Vd"], 4) for t in stats], "events": [t["events"] for t in stats],
                             "cg_iterations": [t["cg_iterations"] for t in stats]},
                "ranks_consistent": len(set(digests)) == 1, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
