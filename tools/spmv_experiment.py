"""dev aid: standalone timing of the CSR and window-staged SpMV kernels at 1M sites"""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, devicekmc_b200 as D
from devicekmc_b200._capi import check
name = sys.argv[1] if len(sys.argv) > 1 else "tiled_1M"
el, x, y, z, lat, nc, p = bench.workload(name); el = bench.substoichiometric(el, p)
dev = D.Device([], p, arrays=(el, x, y, z)); sim = D.KMCProcess(dev, p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
dev.updateCharge(buf, p.metals)
sp = buf.sparsity(nc, nc); lib = dev.ctx.lib
val = torch.empty(sp.nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(sp.m, dtype=torch.float64, device="cuda")
check(lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, 10.0, p.high_G, p.low_G, buf.site_element.data_ptr(),
                          buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(), rhs.data_ptr()))
pad = (sp.m + 7) // 8 * 8
xw = torch.zeros(pad, dtype=torch.float64, device="cuda"); xw[:sp.m] = torch.rand(sp.m, dtype=torch.float64, device="cuda")
y1 = torch.empty(sp.m, dtype=torch.float64, device="cuda"); y2 = torch.empty_like(y1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(f, reps=50):
    for _ in range(5): f()
    e0.record()
    for _ in range(reps): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
vcopy = val.clone(); y0 = torch.empty_like(y1)
csr = lambda: check(lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, vcopy.data_ptr(), xw.data_ptr(), y0.data_ptr()))
pk = lambda: check(lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xw.data_ptr(), y1.data_ptr()))
win = lambda: check(lib.dkmc_spmv_window(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xw.data_ptr(), pad, y2.data_ptr()))
B = 12.0 * sp.nnz + 20.0 * sp.m
t = timeit(csr); print("csr    %.1f us  contract %.0f GB/s" % (t * 1e3, B / t / 1e6))
t = timeit(pk); print("packed %.1f us  contract %.0f GB/s  actual-bytes %.0f GB/s  identical %s" % (t * 1e3, B / t / 1e6, (4.0 * sp.nnz + 28.0 * sp.m) / t / 1e6, bool(torch.equal(y0, y1))))
if os.environ.get("DKMC_WINDOW_SPMV") != "1": sys.exit(0)
t = timeit(win); print("window %.1f us  contract %.0f GB/s  identical %s  dbg=%s" % (t * 1e3, B / t / 1e6, bool(torch.equal(y1, y2)), os.environ.get("DKMC_WIN_DBG", "0")))
