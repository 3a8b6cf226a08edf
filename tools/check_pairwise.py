"""prints the element-wise relative error of the CUDA pairwise sum against the oracle (dev aid)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import devicekmc_b200 as D
from oracle import oracle as O
for name, pbc in (("2.5nm_9k", 0), ("2.5nm_9k", 1), ("tiled_100k", 0)):
    el, x, y, z, lat, nc, p = bench.workload(name)
    import dataclasses
    p = dataclasses.replace(p, pbc=pbc)
    el = bench.substoichiometric(el, p)
    dev = D.Device([], p, arrays=(el, x, y, z)); sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
    dev.updateCharge(buf, p.metals)
    out = dev.updatePotential(buf, p, 5.0, n_contact=nc)
    q = buf.site_charge.cpu().numpy()
    rows = (0, dev.N) if dev.N < 20000 else (dev.N // 2, dev.N // 2 + 20000)
    ref = O.poisson_gridless(x, y, z, lat, pbc, q, p.sigma, p.k, rows=rows)
    got = buf.site_potential_charge.cpu().numpy()[rows[0]:rows[1]]
    nz = ref != 0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    print(name, "pbc", pbc, "max rel err", rel.max(), "at |ref|", np.abs(ref[nz])[rel.argmax()], "min|ref|", np.abs(ref[nz]).min(),
          "cg", out["cg_iterations"], out["cg_est_error"], "pair ms", out["pairwise_ms"])
