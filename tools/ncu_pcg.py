"""dev aid for ncu: one short persistent-PCG solve at 1M sites (30 iterations, no refinement) after a warm step.
ncu --set full --import-source on --clock-control none -k regex:pcg_persistent -c 1 -o gpurun_out/x python tools/ncu_pcg.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench, devicekmc_b200 as D
from devicekmc_b200._capi import SolverOpts
name = sys.argv[1] if len(sys.argv) > 1 else "tiled_1M"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
el, x, y, z, lat, nc, p = bench.workload(name); el = bench.substoichiometric(el, p)
dev = D.Device([], p, arrays=(el, x, y, z)); sim = D.KMCProcess(dev, p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
dev.updateCharge(buf, p.metals)
o = SolverOpts(); dev.ctx.lib.dkmc_default_solver_opts(o)
o.max_iter = iters; o.refine_rounds = 0
out = dev.updatePotential(buf, p, 10.0, n_contact=nc, overlap=False, opts=o)
torch.cuda.synchronize()
print("iterations", out["cg_iterations"], "solve ms", out["solve_ms"])
