#!/usr/bin/env python
"""Benchmark of the DeviceKMC field-and-rate hot path on B200 (BASELINE.json metric:
"KMC steps/sec at 1/2/4/8 B200 (1M sites); CG SpMV HBM GB/s vs peak").

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # CPU arm on the box's host cores

One "step" = one pass of kmc_main.cpp:175-279 with solve_potential = perturb_structure = 1:
updateCharge -> K assembly + CG -> pairwise Coulomb sum -> rate table -> residence-time loop.
Prints ONE JSON line (rank 0).  The oracle (oracle/) is used only by the cpu_baseline /
--impl reference legs, as the thing timed on the CPU — never inside the GPU path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries ONE JSON line: keep NCCL's "NCCL version ..." banner (printed to stdout at the VERSION level) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "kmc_steps_per_sec"
UNIT = "steps/s"


def workload(name):
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import KMCParameters
    if name in S.REAL_WORKLOADS:      # a shipped structure, sites in the reference's order
        data, ny, nz = S.REAL_WORKLOADS[name]
        el, x, y, z, lat, nc, _vd = S.tile_structure(data, ny, nz)
    else:
        ny, nz = S.WORKLOADS[name]
        el, x, y, z, lat, nc = S.tile_device(ny, nz)
    p = KMCParameters(lattice=tuple(lat), num_atoms_contact=nc, num_atoms_first_layer=nc)
    return el, x, y, z, lat, nc, p


def workload_config(name, sites, nn, Vd):
    """the `config` object both arms print (same keys, same workload); arm-specific details go to `arm`"""
    return {"workload": name, "sites": int(sites), "nn": int(nn), "Vd": float(Vd),
            "l2": "inputs larger than L2 (matrix 12*nnz bytes, rate table 16*N*nn bytes)"}


def workload_vd(name, vd):
    """the bias of a workload: --vd, else the V_switch of the structure's parameters.txt (10 V for the synthetic tiles)"""
    from devicekmc_b200 import structures as S
    if vd is not None:
        return float(vd)
    if name in S.REAL_WORKLOADS:
        return S.load_structure(S.REAL_WORKLOADS[name][0])[6]
    return 10.0


def substoichiometric(el, p):
    """Device::makeSubstoichiometric (Device.cpp:202-233) on host arrays, reference random stream"""
    from devicekmc_b200.host import RandomNumberGenerator, DEFECT, OXYGEN_DEFECT, O_EL, VACANCY
    rng = RandomNumberGenerator(p.rnd_seed)
    atom_ind = np.nonzero((el != DEFECT) & (el != OXYGEN_DEFECT))[0]
    n_add = int(p.initial_vacancy_concentration * np.count_nonzero(el == O_EL))
    el = el.copy()
    while n_add > 0:
        u = rng.getRandomNumbers(4096)
        for loc in (u * len(atom_ind)).astype(np.int64):
            if n_add > 0 and el[atom_ind[loc]] == O_EL:
                el[atom_ind[loc]] = VACANCY
                n_add -= 1
    return el


class StateCheckpoint:
    """The simulation state after the warm-up steps (device arrays + the KMC random stream), so that
    the resident-input run, the e2e run and every variant time the SAME steps of the trajectory (the
    workload is not stationary: the events per step fall as the device relaxes)."""
    NAMES = ("site_element", "site_charge", "site_potential_boundary", "site_potential_charge")

    def __init__(self, buf, sim):
        self.t = {n: getattr(buf, n).clone() for n in self.NAMES}
        self.rng = sim.random_generator._bg.state

    def restore(self, buf, sim, dev):
        for n, t in self.t.items():
            getattr(buf, n).copy_(t)
        sim.random_generator._bg.state = self.rng
        buf.sync_GPUToHost(dev)      # the host arrays mirror the device: an e2e step's H2D carries this state


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------- CPU arm
def cpu_step_sample(name, threads=None, pair_rows=None, Vd=10.0, events_per_step=350, event_sample=300, steps=1):
    """KMC steps of the oracle (CPU restatement of the reference's algorithm, sparse K) on the same
    workload, each a BOUNDED sample (~25 s): the pairwise O(N*N_charged) sum is timed on a sample
    of target rows and scaled to N; the residence-time loop is timed on its first `event_sample`
    events and scaled to `events_per_step` (what the GPU arm executes per step on this workload);
    every other stage runs in full.  Like the GPU arm's timed steps, the timed CPU steps are WARM:
    one untimed step comes first, and the CG of a timed step starts from the previous potential.
    The reference's own CPU build (oracle/_ref) cannot run this workload: it allocates a dense
    N x N K (potential_solver.cpp:301) = 8.5 TB at 1 M sites.
    Returns (steps/s, threads, sample description, stage seconds of the last timed step)."""
    from oracle import oracle as O
    threads = threads or os.cpu_count()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    el, x, y, z, lat, nc, p = workload(name)
    el = substoichiometric(el, p)
    N = len(x)
    from devicekmc_b200.host import DEFAULT_LAYERS
    t0 = time.perf_counter(); nb, nn = O.neighbor_list(x, y, z, lat, p.pbc, p.nn_dist, method=1); t_init = time.perf_counter() - t0
    layer = O.site_layers(x, [l.start_x for l in DEFAULT_LAYERS], [l.end_x for l in DEFAULT_LAYERS])
    E = np.array([[l.E_gen_0, l.E_rec_1, l.E_diff_2, l.E_diff_3] for l in DEFAULT_LAYERS])
    rng = O.Rng(1)
    st = {"el": el, "q": np.zeros(N, np.int32), "pb": None}
    rows_used, ncharged_last, its_last = 0, 0, 0

    def one_step():
        nonlocal rows_used, ncharged_last, its_last
        t = {}
        t0 = time.perf_counter(); q = O.update_charge(nb, st["el"], p.metals, st["q"]); t["charge"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        pb, info = O.background_potential(nb, nc, nc, st["el"], q, p.metals, p.high_G, p.low_G, Vd, phi0=st["pb"], tol=1e-12,
                                          refine=1)
        t["potential_boundary"] = time.perf_counter() - t0
        ncharged = int(np.count_nonzero(q))
        rows = pair_rows or max(1024, min(N, int(2.0e8 * threads / 8 / max(ncharged, 1))))
        rows = min(rows, N)
        r0 = (N - rows) // 2
        t0 = time.perf_counter()
        pc_rows = O.poisson_gridless(x, y, z, lat, p.pbc, q, p.sigma, p.k, rows=(r0, r0 + rows))
        t["pairwise_sample"] = time.perf_counter() - t0
        t["pairwise_scaled"] = t["pairwise_sample"] * N / rows
        pc = np.zeros(N); pc[r0:r0 + rows] = pc_rows
        t0 = time.perf_counter()
        et, ep = O.rate_table(nb, layer, lat, p.pbc, p.background_temp, p.freq, p.sigma, p.k, x, y, z, pb, pc, st["el"], q, E)
        t["rate_table"] = time.perf_counter() - t0
        O.set_event_limit(event_sample)
        try:
            t0 = time.perf_counter(); _, ev, el2, q2 = O.kmc_events(nb, et, ep, st["el"], q, p.freq, rng)
            t["event_loop_sample"] = time.perf_counter() - t0
        finally:
            O.set_event_limit(0)
        n_ev = max(len(ev), 1)
        t["event_loop_scaled"] = t["event_loop_sample"] * events_per_step / n_ev
        t["events_sampled"] = n_ev
        st.update(el=el2, q=q2, pb=pb)
        rows_used, ncharged_last, its_last = rows, ncharged, int(info[0])
        return t

    warm = one_step()                     # untimed: cold CG, first events
    secs, t = [], {}
    for _ in range(max(1, steps)):
        t = one_step()
        secs.append(t["charge"] + t["potential_boundary"] + t["pairwise_scaled"] + t["rate_table"] + t["event_loop_scaled"])
    t["init_neighbors"] = t_init
    t["warmup_potential_boundary_cold"] = warm["potential_boundary"]
    sample = (f"oracle port (the reference's algorithm: Jacobi-preconditioned CG on the sparse K, not the GPU arm's cluster "
              f"coarse space), {len(secs)} warm step(s) of {name} after one untimed step (N={N}, N_charged={ncharged_last}, "
              f"{its_last} CG its from the previous potential); pairwise timed on {rows_used} of {N} target rows and scaled; "
              f"event loop timed on its first {int(t['events_sampled'])} events and scaled to {events_per_step} events per "
              f"step; all other stages in full")
    t["_N"], t["_nn"] = N, nn
    return len(secs) / float(np.sum(secs)), threads, sample, t


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a bounded run (~2 min on 16 cores): one untimed (cold) step, then at most 2 timed warm steps whatever --steps says
    value, cores, smp, tim = cpu_step_sample(args.workload, Vd=args.vd, events_per_step=args.events_per_step,
                                             steps=max(1, min(args.steps, 2)))
    N_, nn_ = tim.pop("_N"), tim.pop("_nn")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, N_, nn_, args.vd),
            "arm": {"events_per_step": args.events_per_step,
                    "events_per_step_source": "median of what the GPU arm executes per step in the driver's window "
                                              "(5 warm-up + 20 timed steps of tiled_1M at 10 V: 49)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": smp,
                             "stage_seconds": {k: round(v, 4) for k, v in tim.items()},
                             "why_port": "the reference CPU build needs a dense N x N K (8.5 TB at 1M sites)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import torch
    import devicekmc_b200 as D
    from devicekmc_b200._capi import check
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from devicekmc_b200 import slab
        return slab.bench_multi_gpu(args, METRIC, UNIT)
    torch.cuda.set_device(local)
    el, x, y, z, lat, nc, p = workload(args.workload)
    el = substoichiometric(el, p)
    t0 = time.perf_counter()
    dev = D.Device([], p, arrays=(el, x, y, z))
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    sp = buf.sparsity(nc, nc)
    torch.cuda.synchronize()
    init_s = time.perf_counter() - t0
    Vd = args.vd
    stats = []

    def step(e2e=False):
        if e2e:
            buf.sync_HostToGPU(dev)
        dev.updateCharge(buf, p.metals)
        ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ep0.record()
        o = dev.updatePotential(buf, p, Vd, n_contact=nc)
        ep1.record()
        if not o["cg_converged"]:
            raise RuntimeError(f"the CG of the boundary potential did not converge ({o['cg_iterations']} iterations, "
                               f"error estimate {o['cg_est_error']:.1e}): not a benchmark step")
        sim.executeKMCStep(buf, dev)
        if e2e:
            buf.sync_GPUToHost(dev)
        o["potential_ms"] = ep0.elapsed_time(ep1)
        i = sim.last_info
        o.update(events=i.n_events, fallbacks=i.n_exact_fallbacks, rate_ms=i.rate_ms, loop_ms=i.loop_ms)
        return o

    # clocks are sampled from the warm-up on (same load as the timed steps; the timed region alone
    # lasts only a few hundred ms, too short for nvidia-smi's sampling period)
    sampler = ClockSampler(local); sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ckpt = StateCheckpoint(buf, sim)
    launches0 = dev.ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stats.append(step())
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    launches = dev.ctx.launch_count() - launches0
    clocks = sampler.stop()
    value = args.steps / (ms * 1e-3)

    # e2e: the SAME steps through the host-facing API with HOST buffers (pinned H2D in, D2H out every step)
    ckpt.restore(buf, sim, dev)
    torch.cuda.synchronize()
    e0.record()
    e2e_stats = [step(e2e=True) for _ in range(args.steps)]
    e1.record(); e1.synchronize()
    e2e_value = args.steps / (e0.elapsed_time(e1) * 1e-3)
    e2e_same = [s_["events"] for s_ in e2e_stats] == [s_["events"] for s_ in stats]   # same steps of the trajectory

    # ---- variant (NOT the headline): pairwise sum truncated at 10 sigma (opt-in approximation)
    ckpt.restore(buf, sim, dev)
    check(dev.ctx.lib.dkmc_ctx_set_pairwise_cutoff(dev.ctx.h, 10.0))
    e0.record()
    vstats = [step() for _ in range(args.steps)]
    e1.record(); e1.synchronize()
    variant = {"pairwise_cutoff_10_sigma": {
        "value": args.steps / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
        "potential_ms": float(np.median([s["potential_ms"] for s in vstats])),
        "events": [s["events"] for s in vstats],
        "note": "opt-in: |delta phi_c| < 1e-21 of max|phi_c|; changes the work of the pairwise stage, so it is not the headline"}}
    check(dev.ctx.lib.dkmc_ctx_set_pairwise_cutoff(dev.ctx.h, 0.0))

    # ---- variant (NOT the headline): phi_c updated by the charge differences since the previous step
    ckpt.restore(buf, sim, dev)
    dev.ctx.set_pairwise_incremental(32)
    # the mode's first call is a full sum: prime it on the checkpointed charges, outside the timed region
    check(dev.ctx.lib.dkmc_poisson_gridless(dev.ctx.h, dev.pbc, dev.N, buf.lattice.data_ptr(), buf.sigma.data_ptr(),
                                            buf.k.data_ptr(), buf.site_x.data_ptr(), buf.site_y.data_ptr(),
                                            buf.site_z.data_ptr(), buf.site_charge.data_ptr(),
                                            buf.site_potential_charge.data_ptr()))
    torch.cuda.synchronize()
    e0.record()
    vstats = [step() for _ in range(args.steps)]
    e1.record(); e1.synchronize()
    variant["pairwise_incremental"] = {
        "value": args.steps / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
        "potential_ms": float(np.median([s["potential_ms"] for s in vstats])),
        "pairwise_ms": float(np.median([s["pairwise_ms"] for s in vstats])),
        "events": [s["events"] for s in vstats],
        "note": "opt-in (SURVEY 8f-2): O(N * n_changed) sum of the charge differences, full sum every 32nd step; "
                "norm-wise 1e-13, changes the work of the pairwise stage, so it is not the headline"}
    dev.ctx.set_pairwise_incremental(0)

    # ---- rooflines, measured live with CUDA events on the launching stream
    lib = dev.ctx.lib
    hbm_peak, peak_src = measured_peaks()
    m, nnz = sp.m, sp.nnz
    val = torch.empty(nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(m, dtype=torch.float64, device="cuda")
    check(lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, Vd, p.high_G, p.low_G, buf.site_element.data_ptr(),
                              buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(), rhs.data_ptr()))
    xv = torch.rand(m, dtype=torch.float64, device="cuda"); yv = torch.empty_like(xv)
    reps = 50
    for _ in range(5):
        check(lib.dkmc_spmv(dev.ctx.h, m, nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
    e0.record()
    for _ in range(reps):
        check(lib.dkmc_spmv(dev.ctx.h, m, nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
    e1.record(); e1.synchronize()
    spmv_ms = e0.elapsed_time(e1) / reps
    spmv_bytes = 12.0 * nnz + 20.0 * m
    spmv_gbs = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    # the same product on the structure the SOLVER works on: for inputs in the reference's site order that is the
    # matrix in the registered internal row order (dkmc_solver_set_order), otherwise the same CSR again
    spmv_int_ms = spmv_ms
    if buf.solver_order_applied:
        rp2, col2 = C.c_void_p(), C.c_void_p()
        check(lib.dkmc_solver_csr(dev.ctx.h, C.byref(sp), C.byref(rp2), C.byref(col2), None))
        for _ in range(5):
            check(lib.dkmc_spmv(dev.ctx.h, m, nnz, rp2, col2, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
        e0.record()
        for _ in range(reps):
            check(lib.dkmc_spmv(dev.ctx.h, m, nnz, rp2, col2, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
        e1.record(); e1.synchronize()
        spmv_int_ms = e0.elapsed_time(e1) / reps
    fp64 = C.c_double(0)
    check(lib.dkmc_probe_fp64_tflops(dev.ctx.h, C.byref(fp64)))
    ncharged = int((buf.site_charge != 0).sum().item())
    # the pairwise kernel alone (inside a step it shares the SMs with the CG)
    for _ in range(2):
        check(lib.dkmc_poisson_gridless(dev.ctx.h, dev.pbc, dev.N, buf.lattice.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(),
                                        buf.site_x.data_ptr(), buf.site_y.data_ptr(), buf.site_z.data_ptr(),
                                        buf.site_charge.data_ptr(), buf.site_potential_charge.data_ptr()))
    e0.record()
    for _ in range(3):
        check(lib.dkmc_poisson_gridless(dev.ctx.h, dev.pbc, dev.N, buf.lattice.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(),
                                        buf.site_x.data_ptr(), buf.site_y.data_ptr(), buf.site_z.data_ptr(),
                                        buf.site_charge.data_ptr(), buf.site_potential_charge.data_ptr()))
    e1.record(); e1.synchronize()
    pair_ms = e0.elapsed_time(e1) / 3
    pairs_total = float(dev.N) * ncharged - ncharged
    pe = C.c_longlong(-1)
    check(lib.dkmc_pairwise_pairs_evaluated(dev.ctx.h, C.byref(pe)))
    # the cell-list kernel skips the pairs whose erfc factor is exactly 0 (r > 131 A): count the rest
    pairs = float(pe.value) if pe.value > 0 else pairs_total
    pair_tflops = 200.0 * pairs / (pair_ms * 1e-3) / 1e12
    # scan primitive over the N*nn rate table
    n_tab = dev.N * buf.nn_
    tab = torch.rand(n_tab, dtype=torch.float64, device="cuda"); cum = torch.empty_like(tab)
    for _ in range(3):
        check(lib.dkmc_inclusive_scan(dev.ctx.h, n_tab, tab.data_ptr(), cum.data_ptr()))
    e0.record()
    for _ in range(10):
        check(lib.dkmc_inclusive_scan(dev.ctx.h, n_tab, tab.data_ptr(), cum.data_ptr()))
    e1.record(); e1.synchronize()
    scan_ms = e0.elapsed_time(e1) / 10
    rate_ms = float(np.median([s["rate_ms"] for s in stats]))
    med = lambda k: float(np.median([s[k] for s in stats]))
    # cg_solve and pairwise run CONCURRENTLY (main / side stream): "potential" is the wall time of both
    shares = {"potential": med("potential_ms"), "assemble": med("assemble_ms"), "cg_solve": med("solve_ms"),
              "pairwise": med("pairwise_ms"), "rate_table": rate_ms, "event_loop": med("loop_ms")}
    traffic, pipe_pct = {}, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get(args.workload, {})
        pipe_pct = tj.get("pairwise_fp64_pipe_pct")
    rooflines = {
        "spmv": {"bound": "hbm", "achieved": spmv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": spmv_gbs / hbm_peak,
                 "traffic": None, "bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms, "peak_source": peak_src,
                 "order": "caller's site order" + (" (the reference's: lattice atoms before interstitials)" if buf.solver_order_applied else " (x-major grid cells)")},
        "spmv_solver_order": {"bound": "hbm", "achieved": spmv_bytes / (spmv_int_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": spmv_bytes / (spmv_int_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                              "bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_int_ms,
                              "order": "internal row order of the solver (x-major grid cells)" if buf.solver_order_applied
                                       else "same structure as `spmv`"},
        "pairwise": {"bound": "fp64", "achieved": pair_tflops, "peak": fp64.value, "unit": "TFLOP/s",
                     "frac": pair_tflops / fp64.value if fp64.value else None, "traffic": None,
                     "flops_per_pair": 200, "pairs": pairs, "pairs_all": pairs_total,
                     "pairs_skipped_exact_zero": 1.0 - pairs / pairs_total, "ms_per_launch": pair_ms,
                     "peak_source": "measured here: DFMA-chain probe (dkmc_probe_fp64_tflops)"},
        "rate_table": {"bound": "hbm", "achieved": 16.0 * n_tab / (rate_ms * 1e-3) / 1e9, "peak": hbm_peak,
                       "unit": "GB/s", "frac": 16.0 * n_tab / (rate_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None},
        "scan": {"bound": "hbm", "achieved": 16.0 * n_tab / (scan_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                 "frac": 16.0 * n_tab / (scan_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None},
    }
    # one CG iteration inside a step (beside the pairwise sum): SURVEY 8d's 12 nnz + 92 m bytes over the measured
    # time per iteration of the timed steps
    its_total = max(1, int(np.sum([s["cg_iterations"] for s in stats])))
    cg_us = 1e3 * float(np.sum([s["solve_ms"] for s in stats])) / its_total
    cg_bytes = 12.0 * nnz + 92.0 * m
    rooflines["cg_iteration"] = {"bound": "hbm", "achieved": cg_bytes / (cg_us * 1e-6) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": cg_bytes / (cg_us * 1e-6) / 1e9 / hbm_peak, "traffic": None, "bytes_per_iteration": cg_bytes,
                                 "us_per_iteration": cg_us,
                                 "note": "in-step: the persistent PCG shares the SMs with the pairwise sum (FP64 pipe)"}
    for k_, r_ in rooflines.items():
        r_["traffic"] = traffic.get(k_)
    # the contract counts 200 flop per pair; the kernel executes ~62 FP64 instructions per
    # pair, so `frac` > 1 is work saved, not a faster pipe: the pipe utilisation is the ncu figure
    rooflines["pairwise"]["fp64_pipe_active_pct_ncu"] = pipe_pct
    pf = C.c_longlong(-1)
    check(lib.dkmc_pairwise_pairs_far(dev.ctx.h, C.byref(pf)))
    rooflines["pairwise"]["pairs_far_field_formula"] = float(pf.value) if pf.value >= 0 else None
    rooflines["pairwise"]["note"] = ("achieved = 200 contract flop/pair / time; the kernel needs ~40 FP64 instructions per "
                                     "far-field pair (no square root, one division) and ~62 per near pair, so frac > 1 is work "
                                     "saved; pipe utilisation = fp64_pipe_active_pct_ncu")
    dominant = max((k_ for k_ in shares if k_ != "potential"), key=shares.get)
    roof_key = {"pairwise": "pairwise", "cg_solve": "spmv_solver_order", "assemble": "spmv_solver_order", "rate_table": "rate_table",
                "event_loop": "rate_table"}[dominant]
    roofline = dict(rooflines[roof_key]); roofline["kernel"] = roof_key; roofline["dominant_stage"] = dominant

    cpu = None
    if not args.no_cpu_baseline:
        v, cores, smp, tim = cpu_step_sample(args.workload, Vd=Vd, events_per_step=int(np.median([s["events"] for s in stats])))
        tim.pop("_N"); tim.pop("_nn")
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": smp,
               "stage_seconds": {k: round(t, 4) for k, t in tim.items()}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, dev.N, buf.nn_, Vd),
            "arm": {"interior_rows": m, "nnz": nnz, "n_charged": ncharged, "init_seconds": round(init_s, 3),
                    "overlap": "pairwise sum on a side stream, concurrent with the CG",
                    "solver_order": ("internal x-major grid-cell order (input in the reference's site order)"
                                     if buf.solver_order_applied else "caller's (already x-major by grid cell)")},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": buf.h2d_bytes(), "d2h_bytes_per_step": buf.d2h_bytes(),
                    "same_steps_as_value": bool(e2e_same)},
            "roofline": roofline, "rooflines": rooflines, "stage_ms": shares, "variants": variant,
            "per_step": {"events": [s["events"] for s in stats], "exact_fallbacks": [s["fallbacks"] for s in stats],
                         "cg_iterations": [s["cg_iterations"] for s in stats]},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tiled_1M")
    ap.add_argument("--vd", type=float, default=None, help="bias [V]; default: the workload's own (10 V for the synthetic tiles)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--events-per-step", type=int, default=49,
                    help="--impl reference: events per KMC step the bounded event-loop sample is scaled to (the median of "
                         "what the GPU arm executes per step in the driver's 5 + 20 window on tiled_1M)")
    ap.add_argument("--replicated-cg", action="store_true",
                    help="N>1: every rank runs the whole CG beside its share of the pairwise sum, instead of the default "
                         "slab-partitioned CG whose per-iteration exchange goes through NVLink peer memory")
    ap.add_argument("--distributed-cg", action="store_true", help=argparse.SUPPRESS)  # the default since round 1
    ap.add_argument("--ramp", action="store_true",
                    help="BASELINE config 5: the I-V sweep through the reference's bias-point loop, weak scaling (the device grows "
                         "with --gpus: ~0.5 M sites per GPU, 4 M sites at 8)")
    ap.add_argument("--ramp-points", type=int, default=200)
    ap.add_argument("--ramp-vmax", type=float, default=4.0)
    ap.add_argument("--ramp-start", type=int, default=60, help="first bias point of the warm-up (point 60 of 200 = 2.4 V)")
    ap.add_argument("--ramp-steps-per-point", type=int, default=1)
    args = ap.parse_args()
    args.vd = workload_vd(args.workload, args.vd)
    if args.impl == "reference":
        reference_arm(args)
    elif args.ramp:
        from devicekmc_b200 import slab
        slab.bench_ramp(args, METRIC, UNIT)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
