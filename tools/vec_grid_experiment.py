"""dev aid: CG microseconds per iteration at 1M sites (CG alone, no pairwise overlap) for several residencies
of the grid-stride vector kernels (DKMC_VEC_CPS = CTAs per SM; 16 = the default = one pass per thread)"""
import os, subprocess, sys
code = r'''
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, devicekmc_b200 as D
el,x,y,z,lat,nc,p = bench.workload("tiled_1M"); el = bench.substoichiometric(el,p)
dev = D.Device([],p,arrays=(el,x,y,z)); sim = D.KMCProcess(dev,p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
res=[]
for s in range(6):
    dev.updateCharge(buf,p.metals); o = dev.updatePotential(buf,p,10.0,n_contact=nc,overlap=(s%2==1)); sim.executeKMCStep(buf,dev)
    res.append((o["solve_ms"], o["cg_iterations"], o["overlap"]))
print("VEC_CPS", os.environ.get("DKMC_VEC_CPS","16"), "us/iter alone", [round(1e3*a/b,1) for a,b,c in res[2:] if not c],
      "overlapped", [round(1e3*a/b,1) for a,b,c in res[2:] if c], "iters", [b for a,b,c in res[2:]])
'''
for f in sys.argv[1:] or ["16", "5"]:
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DKMC_VEC_CPS=f), capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
