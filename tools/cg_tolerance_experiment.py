"""dev aid: iterations / time / accuracy of the refined PCG for several tolerance settings at 1M
sites, from a realistic warm start (the previous step's potential)"""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, devicekmc_b200 as D
name = sys.argv[1] if len(sys.argv) > 1 else "tiled_1M"
el,x,y,z,lat,nc,p = bench.workload(name); el = bench.substoichiometric(el,p)
dev = D.Device([],p,arrays=(el,x,y,z)); sim = D.KMCProcess(dev,p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
for s in range(2):
    dev.updateCharge(buf,p.metals); dev.updatePotential(buf,p,10.0,n_contact=nc); sim.executeKMCStep(buf,dev)
dev.updateCharge(buf,p.metals)
w0 = buf.site_potential_boundary.clone()
def run(**kw):
    o = D.SolverOpts(); dev.ctx.lib.dkmc_default_solver_opts(C.byref(o))
    for k,v in kw.items(): setattr(o,k,v)
    buf.site_potential_boundary.copy_(w0)
    out = dev.updatePotential(buf,p,10.0,n_contact=nc,opts=o)
    return out, buf.site_potential_boundary.clone()
_, ref = run(rel_tol=1e-13, refine_rounds=8, est_tol=1e-17, refine_tol=1e-8)
print("ref done")
for kw in [dict(), dict(refine_tol=1e-4), dict(refine_tol=1e-3), dict(rel_tol=1e-9), dict(rel_tol=1e-9, refine_tol=1e-4), dict(rel_tol=1e-8, refine_tol=1e-3, refine_rounds=8),
           dict(rel_tol=1e-10, refine_tol=1e-5), dict(est_tol=1e-13), dict(rel_tol=1e-9, refine_tol=1e-4, est_tol=1e-13), dict(cluster_precond=0, refine_rounds=8)]:
    out, sol = run(**kw)
    err = float((sol-ref).abs().max()/ref.abs().max())
    print(kw, "iters", out["cg_iterations"], "rounds", out["cg_refinements"], "ms %.1f" % out["solve_ms"], "est %.1e" % out["cg_est_error"], "err vs ref %.1e" % err, flush=True)
