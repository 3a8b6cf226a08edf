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
# one public SpMV launch on the step's matrix (spmv_tile_kernel<0>: the SpMV roofline's kernel; inside a step the
# same tile code runs as a phase of the persistent PCG kernel)
import ctypes as C  # noqa: E402
from devicekmc_b200._capi import check  # noqa: E402
sp = buf.sparsity(nc, nc)
val = torch.empty(sp.nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(sp.m, dtype=torch.float64, device="cuda")
check(dev.ctx.lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, args.vd, p.high_G, p.low_G, buf.site_element.data_ptr(),
                                  buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(), rhs.data_ptr()))
xv = torch.rand(sp.m, dtype=torch.float64, device="cuda"); yv = torch.empty_like(xv)
for _ in range(3):
    check(dev.ctx.lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = step()
check(dev.ctx.lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled step:", out, "events", sim.last_info.n_events)
