"""Where does the e2e step (host buffers in / out every step) spend its time beyond the resident step?
python tools/e2e_probe.py [workload]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import devicekmc_b200 as D

name = sys.argv[1] if len(sys.argv) > 1 else "tiled_1M"
el, x, y, z, lat, nc, p = bench.workload(name)
el = bench.substoichiometric(el, p)
dev = D.Device([], p, arrays=(el, x, y, z))
sim = D.KMCProcess(dev, p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
buf.sync_HostToGPU(dev)
torch.cuda.synchronize()


def timed(fn, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    return round(float(np.median(ts)), 3)


def step(e2e):
    if e2e: buf.sync_HostToGPU(dev)
    dev.updateCharge(buf, p.metals)
    dev.updatePotential(buf, p, 10.0, n_contact=nc)
    sim.executeKMCStep(buf, dev)
    if e2e: buf.sync_GPUToHost(dev)


for _ in range(4): step(False)
print("h2d sync ms", timed(lambda: buf.sync_HostToGPU(dev)), "bytes", buf.h2d_bytes())
print("d2h sync ms", timed(lambda: buf.sync_GPUToHost(dev)))
a = torch.empty(buf.N_, dtype=torch.float64).pin_memory(); b = torch.empty(buf.N_, dtype=torch.float64, device="cuda")
print("8 MB pinned h2d ms", timed(lambda: b.copy_(a, non_blocking=True)), "d2h ms", timed(lambda: a.copy_(b, non_blocking=True)))
print("pinned flags", {k: v[0].is_pinned() for k, v in buf._pin.items() if hasattr(v[0], "is_pinned")})
print("resident step ms", timed(lambda: step(False), 6))
print("e2e step ms", timed(lambda: step(True), 6))
def parts():
    out = {}
    for nm, fn in [("h2d", lambda: buf.sync_HostToGPU(dev)), ("charge", lambda: dev.updateCharge(buf, p.metals)),
                   ("potential", lambda: dev.updatePotential(buf, p, 10.0, n_contact=nc)),
                   ("events", lambda: sim.executeKMCStep(buf, dev)), ("d2h", lambda: buf.sync_GPUToHost(dev))]:
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); out[nm] = round((time.perf_counter() - t) * 1e3, 2)
    return out
for _ in range(4): print(parts())
