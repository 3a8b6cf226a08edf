"""One KMC step of the benchmark workload between cudaProfilerStart/Stop, for ncu
(`--profile-from-start off`).  Numbers printed under a profiler are never bench values."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
import devicekmc_b200 as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="tiled_1M")
ap.add_argument("--vd", type=float, default=10.0)
ap.add_argument("--warm", type=int, default=2)
args = ap.parse_args()

el, x, y, z, lat, nc, p = bench.workload(args.workload)
el = bench.substoichiometric(el, p)
dev = D.Device([], p, arrays=(el, x, y, z))
sim = D.KMCProcess(dev, p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
buf.sync_HostToGPU(dev)


def step():
    dev.updateCharge(buf, p.metals)
    out = dev.updatePotential(buf, p, args.vd, n_contact=nc)
    sim.executeKMCStep(buf, dev)
    return out


for _ in range(args.warm):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled step:", out, "events", sim.last_info.n_events)
