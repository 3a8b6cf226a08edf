"""dev aid: KMC step time at 1M sites with the pairwise sum serial vs overlapped with the CG on the
side stream, for several residencies (CTAs per SM) of the overlapped pairwise kernel"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, devicekmc_b200 as D
from devicekmc_b200._capi import check
name = sys.argv[1] if len(sys.argv) > 1 else "tiled_1M"
el, x, y, z, lat, nc, p = bench.workload(name); el = bench.substoichiometric(el, p)
dev = D.Device([], p, arrays=(el, x, y, z)); sim = D.KMCProcess(dev, p.freq)
buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals); buf.sync_HostToGPU(dev)
for s in range(3):
    dev.updateCharge(buf, p.metals); dev.updatePotential(buf, p, 10.0, n_contact=nc); sim.executeKMCStep(buf, dev)
dev.updateCharge(buf, p.metals)
w0 = buf.site_potential_boundary.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(overlap, share=None):
    if share: check(dev.ctx.lib.dkmc_ctx_set_pairwise_share(dev.ctx.h, share[0], share[1]))
    ts = []
    for rep in range(3):
        buf.site_potential_boundary.copy_(w0)
        torch.cuda.synchronize(); e0.record()
        out = dev.updatePotential(buf, p, 10.0, n_contact=nc, overlap=overlap)
        e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return ts, out
t, out = run(False)
ref_c = buf.site_potential_charge.clone(); ref_b = buf.site_potential_boundary.clone()
fmt = lambda ts: "/".join("%.1f" % v for v in ts)
pe = C.c_longlong(0); dev.ctx.lib.dkmc_pairwise_pairs_evaluated(dev.ctx.h, C.byref(pe))
nchg = int((buf.site_charge != 0).sum().item())
print("pairs evaluated %.3e of %.3e (%.1f%%)" % (pe.value, dev.N * nchg, 100.0 * pe.value / (dev.N * nchg)))
print("serial      total %s ms  cg %.2f pw %.2f iters %d" % (fmt(t), out["solve_ms"], out["pairwise_ms"], out["cg_iterations"]), flush=True)
for share in ((2, 128), (2, 160), (2, 192), (3, 96), (4, 96), (5, 64)):
    t, out = run(True, share)
    same_c = bool(torch.equal(ref_c, buf.site_potential_charge))
    err_b = float((ref_b - buf.site_potential_boundary).abs().max() / ref_b.abs().max())
    print("overlap share=%s total %s ms  cg %.2f pw %.2f iters %d phi_c identical %s phi_b diff %.1e" % (share, fmt(t), out["solve_ms"], out["pairwise_ms"], out["cg_iterations"], same_c, err_b), flush=True)
# standalone SpMV (CTAs per SM from DKMC_SPMV_CPS)
sp = buf.sparsity(nc, nc)
val = torch.empty(sp.nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(sp.m, dtype=torch.float64, device="cuda")
lib = dev.ctx.lib
check(lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, 10.0, p.high_G, p.low_G, buf.site_element.data_ptr(),
                          buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(), rhs.data_ptr()))
xv = torch.rand(sp.m, dtype=torch.float64, device="cuda"); yv = torch.empty_like(xv)
for _ in range(5): check(lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
e0.record()
for _ in range(50): check(lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xv.data_ptr(), yv.data_ptr()))
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 50
print("spmv standalone CPS=%s: %.1f us, %.0f GB/s" % (os.environ.get("DKMC_SPMV_CPS", "3"), ms * 1e3, (12.0 * sp.nnz + 20.0 * sp.m) / ms / 1e6))

pad = (sp.m + 7) // 8 * 8
xw = torch.zeros(pad, dtype=torch.float64, device="cuda"); xw[:sp.m] = xv
yw = torch.empty_like(xv)
st = lib.dkmc_spmv_window(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xw.data_ptr(), pad, yw.data_ptr())
if st == 0:
    for _ in range(5): lib.dkmc_spmv_window(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xw.data_ptr(), pad, yw.data_ptr())
    e0.record()
    for _ in range(50): lib.dkmc_spmv_window(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), xw.data_ptr(), pad, yw.data_ptr())
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print("spmv window-staged: %.1f us, contract %.0f GB/s, identical to CSR %s" % (ms * 1e3, (12.0 * sp.nnz + 20.0 * sp.m) / ms / 1e6, bool(torch.equal(yv, yw))))
else:
    print("window-staged format not applicable:", lib.dkmc_last_error().decode())
