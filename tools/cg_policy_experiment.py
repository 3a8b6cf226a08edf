"""dev aid: CG microseconds per iteration at 1M sites for the DKMC_FLAGS cache-policy variants"""
import os, subprocess, sys
code = r'''
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, devicekmc_b200 as D
el,x,y,z,lat,nc,p = bench.workload("tiled_1M"); el = bench.substoichiometric(el,p)
dev = D.Device([],p,arrays=(el,x,y,z)); sim = D.KMCProcess(dev,p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
res=[]
for s in range(5):
    dev.updateCharge(buf,p.metals); o = dev.updatePotential(buf,p,10.0,n_contact=nc); sim.executeKMCStep(buf,dev)
    res.append((o["solve_ms"], o["cg_iterations"]))
print("FLAGS", os.environ.get("DKMC_FLAGS","0"), "us/iter", [round(1e3*a/b,1) for a,b in res[2:]], "iters", [b for a,b in res[2:]])
'''
for f in sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "7"]:
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DKMC_FLAGS=f), capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
    for l in r.stderr.splitlines():
        if l.startswith("dkmc:"): print("   ", l)
