"""dev aid: the CG at 1M sites — persistent-kernel PCG vs the one-kernel-per-operation path, alone and
overlapped with the pairwise sum, for several grid sizes (DKMC_PCG_CPS, read once per process, hence one
subprocess per configuration).   python tools/pcg_experiment.py [workload]"""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, os.getcwd())

CHILD = r'''
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, devicekmc_b200 as D
from devicekmc_b200._capi import check
name = sys.argv[1]
el, x, y, z, lat, nc, p = bench.workload(name); el = bench.substoichiometric(el, p)
dev = D.Device([], p, arrays=(el, x, y, z)); sim = D.KMCProcess(dev, p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
lib = dev.ctx.lib
for s in range(3):
    dev.updateCharge(buf, p.metals); dev.updatePotential(buf, p, 10.0, n_contact=nc); sim.executeKMCStep(buf, dev)
dev.updateCharge(buf, p.metals)
w0 = buf.site_potential_boundary.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for legacy in (1, 0):
    check(lib.dkmc_ctx_set_legacy_cg(dev.ctx.h, legacy))
    for overlap in (False, True):
        ts, outs = [], []
        prof = (C.c_double * 8)()
        check(lib.dkmc_pcg_profile(dev.ctx.h, prof))
        for rep in range(3):
            buf.site_potential_boundary.copy_(w0)
            torch.cuda.synchronize(); e0.record()
            out = dev.updatePotential(buf, p, 10.0, n_contact=nc, overlap=overlap)
            e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1)); outs.append(out)
        check(lib.dkmc_pcg_profile(dev.ctx.h, prof))
        pb = buf.site_potential_boundary.clone()
        if ref is None: ref = pb
        err = float((pb - ref).abs().max() / ref.abs().max())
        o = outs[-1]
        line = "%s overlap=%d total %s ms | cg %.2f ms %d its = %.1f us/it | pw %.2f ms | est %.1e conv %s | vs first %.1e" % (
            "per-op    " if legacy else "persistent", overlap, "/".join("%.1f" % t for t in ts), o["solve_ms"], o["cg_iterations"],
            1e3 * o["solve_ms"] / max(o["cg_iterations"], 1), o["pairwise_ms"], o["cg_est_error"], o["cg_converged"], err)
        if prof[6] > 0:
            it = prof[6]
            line += " | prof us/it: V %.1f barH %.1f S %.1f cl %.1f barR %.1f; setup %.0f us/solve (%d solves)" % (
                prof[1] / it / 1e3, prof[2] / it / 1e3, prof[3] / it / 1e3, prof[4] / it / 1e3, prof[5] / it / 1e3,
                prof[0] / max(prof[7], 1) / 1e3, prof[7])
        print(line, flush=True)
'''

name = sys.argv[1] if len(sys.argv) > 1 else "tiled_1M"
configs = sys.argv[2:] or ["6,0,3,0", "6,0,3,0p", "5,1,2,1", "4,2,2,2"]
for cfg in configs:
    env = dict(os.environ)
    if cfg.endswith("p"):      # "...p": with the in-kernel profile (its kernels carry extra registers: timings are distorted)
        cfg = cfg[:-1]
        env.update(DKMC_PCG_PROF="1", DKMC_PCG_PROF_CTAS="1")
    if cfg:
        env["DKMC_PCG_CPS"] = cfg
    print("=== DKMC_PCG_CPS=%r (alone_cps,alone_variant,overlap_cps,overlap_variant; variant 0/1/2 = 40/48/64 registers)" % cfg, flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD, name], env=env, capture_output=True, text=True)
    print(r.stdout, end="")
    print("\n".join(l for l in r.stderr.splitlines() if "pcg prof" in l))
    if r.returncode != 0:
        print("FAILED:", r.stderr[-3000:])
