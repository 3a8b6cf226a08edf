"""BASELINE configs 2 and 3 on the reference's SHIPPED structures, sites in the reference's own order
(structures/single_devices/7.5nm_potential, structures/crossbars/10nm_5pitch — data/*.npz, tools/make_structure_data.py):
the path through the C-ABI against the oracle, with the solver's internal row order (dkmc_solver_set_order)
registered automatically because the input is not spatially sorted."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel_inf(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _build(name, solver_order="auto"):
    import bench
    import devicekmc_b200 as D
    el, x, y, z, lat, nc, p = bench.workload(name)
    el = bench.substoichiometric(el, p)
    dev = D.Device([], p, arrays=(el, x, y, z))
    dev.solver_order = solver_order
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    dev.updateCharge(buf, p.metals)
    return dict(p=p, dev=dev, sim=sim, buf=buf, nc=nc, Vd=bench.workload_vd(name, None))


@pytest.fixture(scope="module")
def dev75():
    return _build("device_7.5nm")


def test_7p5nm_is_the_shipped_structure_in_reference_order(dev75):
    d = dev75["dev"]
    assert d.N == 85071 and dev75["nc"] == 12960 and dev75["Vd"] == 10.0
    # the reference's order: lattice atoms first, interstitials (`d`) last — not spatially sorted
    from devicekmc_b200 import structures
    sl = slice(dev75["nc"], d.N - dev75["nc"])
    order = structures.cell_order(d.site_x[sl], d.site_y[sl], d.site_z[sl], x0=float(d.site_x.min()))
    assert not np.array_equal(order, np.arange(len(order)))


def test_7p5nm_graph_charge_csr_assembly_bit_exact(dev75, O, torch):
    from devicekmc_b200._capi import check
    p, dev, buf, nc = dev75["p"], dev75["dev"], dev75["buf"], dev75["nc"]
    nb, nn = O.neighbor_list(dev.site_x, dev.site_y, dev.site_z, dev.lattice, p.pbc, p.nn_dist, method=1)
    assert nn == dev.max_num_neighbors
    assert np.array_equal(dev.neigh_idx.reshape(dev.N, -1), nb)
    q = buf.site_charge.cpu().numpy()
    assert np.array_equal(q, O.update_charge(nb, dev.site_element, p.metals, np.zeros(dev.N, np.int32)))
    sp = buf.sparsity(nc, nc)
    cs = O.csr_structure(nb, nc, nc)
    rt = C.CDLL("libcudart.so.12")

    def ints(ptr, n):
        out = np.empty(n, np.int32)
        torch.cuda.synchronize()
        assert rt.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(ptr), C.c_size_t(out.nbytes), 2) == 0
        return out
    assert np.array_equal(ints(sp.d_row_ptr, sp.m + 1), cs["row_ptr"])
    assert np.array_equal(ints(sp.d_col, sp.nnz), cs["col"])
    # assembled values and right-hand side in the CALLER's order (the internal order is the solver's business)
    val = torch.empty(sp.nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(sp.m, dtype=torch.float64, device="cuda")
    check(dev.ctx.lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, 10.0, p.high_G, p.low_G, buf.site_element.data_ptr(),
                                      buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(),
                                      rhs.data_ptr()))
    v_ref, b_ref = O.assemble_K(nb, nc, nc, dev.site_element, q, p.metals, p.high_G, p.low_G, 10.0, cs["row_ptr"], cs["col"])
    assert np.array_equal(val.cpu().numpy(), v_ref) and np.array_equal(rhs.cpu().numpy(), b_ref)
    dev75["nb"], dev75["q"] = nb, q


def test_7p5nm_potentials_match_oracle_with_internal_order(dev75, O, torch):
    """config 2: constant-voltage forming step at 10 V on the shipped 85 k-site device"""
    from devicekmc_b200._capi import check
    p, dev, buf, nc, Vd = dev75["p"], dev75["dev"], dev75["buf"], dev75["nc"], dev75["Vd"]
    buf.sparsity(nc, nc)
    assert buf.solver_order_applied                      # reference-ordered input: x-major cell order inside the solver
    out = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert out["cg_converged"] and out["cg_est_error"] <= 1e-13
    nb, q = dev75["nb"], dev75["q"]
    ref, _ = O.background_potential(nb, nc, nc, dev.site_element, q, p.metals, p.high_G, p.low_G, Vd, refine=3)
    pb = buf.site_potential_boundary.cpu().numpy()
    assert rel_inf(pb, ref) <= TOL
    rows = (dev.N // 2, dev.N // 2 + 3000)
    refc = O.poisson_gridless(dev.site_x, dev.site_y, dev.site_z, dev.lattice, p.pbc, q, p.sigma, p.k, rows=rows)
    pc = buf.site_potential_charge.cpu().numpy()
    assert np.abs(pc[rows[0]:rows[1]] - refc).max() <= TOL * np.abs(refc).max()
    assert np.all(np.abs(pc[rows[0]:rows[1]] - refc) <= TOL * np.abs(refc) + 1e-300)
    # the same solve in the caller's order: same solution to rounding, more iterations' worth of time
    check(dev.ctx.lib.dkmc_solver_set_order(dev.ctx.h, C.byref(buf.sparsity(nc, nc)), None))
    buf.site_potential_boundary.zero_()
    out2 = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert out2["cg_converged"]
    assert rel_inf(buf.site_potential_boundary.cpu().numpy(), pb) <= 1e-12
    check(dev.ctx.lib.dkmc_solver_set_order(dev.ctx.h, C.byref(buf.sparsity(nc, nc)), buf._order_tensor.data_ptr()))
    dev75["pb"], dev75["pc"] = pb, pc


def test_7p5nm_step_matches_oracle_events(dev75, O, torch):
    """rate table and the residence-time loop on the device's own potentials: event for event"""
    import devicekmc_b200 as D
    p, dev, sim, buf, nc, Vd = dev75["p"], dev75["dev"], dev75["sim"], dev75["buf"], dev75["nc"], dev75["Vd"]
    dev.updatePotential(buf, p, Vd, n_contact=nc)
    gpb, gpc = buf.site_potential_boundary.cpu().numpy(), buf.site_potential_charge.cpu().numpy()
    nb, q = dev75["nb"], dev75["q"]
    et, ep = O.rate_table(nb, sim.site_layer, dev.lattice, p.pbc, dev.T_bg, sim.freq, dev.sigma, dev.k, dev.site_x, dev.site_y,
                          dev.site_z, gpb, gpc, dev.site_element, q, buf.E_host.T)
    rng = O.Rng(D.host.RND_SEED_KMC)
    t_ref, ev_ref, el_ref, q_ref = O.kmc_events(nb, et, ep, dev.site_element, q, sim.freq, rng)
    t = sim.executeKMCStep(buf, dev, record_events=65536)
    assert len(ev_ref) > 0
    assert np.array_equal(sim.last_events, ev_ref)
    assert np.array_equal(buf.site_element.cpu().numpy(), el_ref)
    assert np.array_equal(buf.site_charge.cpu().numpy(), q_ref)
    assert abs(t - t_ref) <= 1e-12 * abs(t_ref)


def test_crossbar_2x2_potential_residual_and_order(torch):
    """config 3's device (the 10 nm crossbar tiled 2 x 2, 443 k sites, 15 V): the solve with the internal order
    meets its own error estimate, and its solution satisfies the maximum principle of the Laplace problem"""
    s = _build("crossbar_2x2")
    p, dev, buf, nc, Vd = s["p"], s["dev"], s["buf"], s["nc"], s["Vd"]
    assert dev.N == 4 * 110813 and Vd == 15.0
    out = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert buf.solver_order_applied
    assert out["cg_converged"] and out["cg_est_error"] <= 1e-13
    pb = buf.site_potential_boundary
    assert float(pb.max()) <= Vd / 2 * (1 + 1e-12) and float(pb.min()) >= -Vd / 2 * (1 + 1e-12)
    assert torch.all(pb[:nc] == -Vd / 2) and torch.all(pb[-nc:] == Vd / 2)


def test_28k_tile_matches_the_reference_build(O, torch):
    """the GPU path against the reference's own CPU build at 28 k sites (s_tile28k.npz: dense K + dgesv), input in
    the reference's site order copy after copy, state with uncharged-vacancy clusters"""
    import os
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "s_tile28k.npz")
    if not os.path.exists(path):
        pytest.skip("s_tile28k.npz not generated")
    import devicekmc_b200 as D
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import KMCParameters
    g = np.load(path)
    el0, x, y, z, lat, nc = S.tile_device(1, 3, order="tile")
    p = KMCParameters(lattice=tuple(lat), num_atoms_contact=nc, num_atoms_first_layer=nc)
    dev = D.Device([], p, arrays=(g["element"].astype(np.int32), x, y, z))
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    dev.updateCharge(buf, p.metals)
    assert np.array_equal(buf.site_charge.cpu().numpy(), g["charge"].astype(np.int32))
    out = dev.updatePotential(buf, p, float(g["Vd"]), n_contact=nc)
    assert out["cg_converged"] and buf.solver_order_applied
    ref_b, ref_c = g["potential_boundary"], g["potential_charge"]
    # vs the reference's dgesv (1e-9 off its own matrix) and, tighter, vs the oracle
    assert rel_inf(buf.site_potential_boundary.cpu().numpy(), ref_b) <= 5e-9
    nb = dev.neigh_idx.reshape(dev.N, -1)
    q = buf.site_charge.cpu().numpy()
    pb_o, _ = O.background_potential(nb, nc, nc, dev.site_element, q, p.metals, p.high_G, p.low_G, float(g["Vd"]), refine=3)
    assert rel_inf(buf.site_potential_boundary.cpu().numpy(), pb_o) <= TOL
    pc = buf.site_potential_charge.cpu().numpy()
    assert np.all(np.abs(pc - ref_c) <= TOL * np.abs(ref_c) + 1e-300)


def test_solver_order_rejects_what_is_not_a_permutation_and_invalidate_rebuilds(dev75, torch):
    """dkmc_solver_set_order checks its argument; dkmc_ctx_invalidate drops the address-keyed caches and the next
    solve rebuilds them (same solution)"""
    from devicekmc_b200 import _capi
    from devicekmc_b200._capi import check
    p, dev, buf, nc, Vd = dev75["p"], dev75["dev"], dev75["buf"], dev75["nc"], dev75["Vd"]
    sp = buf.sparsity(nc, nc)
    lib = dev.ctx.lib
    bad = buf._order_tensor.clone()
    bad[1] = bad[0]                                           # a repeated row
    assert lib.dkmc_solver_set_order(dev.ctx.h, C.byref(sp), bad.data_ptr()) == _capi.DKMC_ERR_ARG
    bad[1] = sp.m                                             # out of range
    assert lib.dkmc_solver_set_order(dev.ctx.h, C.byref(sp), bad.data_ptr()) == _capi.DKMC_ERR_ARG
    # a rejected order leaves the solver in the caller's order: still the same solution
    out = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert out["cg_converged"]
    ref = buf.site_potential_boundary.clone()
    check(lib.dkmc_solver_set_order(dev.ctx.h, C.byref(sp), buf._order_tensor.data_ptr()))
    check(lib.dkmc_ctx_invalidate(dev.ctx.h))
    buf.site_potential_boundary.zero_()
    out = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert out["cg_converged"]
    assert float((buf.site_potential_boundary - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
