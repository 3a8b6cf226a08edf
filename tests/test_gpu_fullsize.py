"""Parity at BASELINE.json's full size (tiled_1M: 1 033 890 sites, the bench workload), through the
C-ABI: integer stages bit-exact against the oracle on the whole device; floating-point stages against
the oracle on what it can finish in seconds (the sparse restatement's assembly, a band of target rows
of the pairwise sum, the rate table on the device's own potentials, the first 25 events) plus
size-independent properties (residual of the solve, linearity in Vd, superposition of charges)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def big():
    import bench
    import devicekmc_b200 as D
    el, x, y, z, lat, nc, p = bench.workload("tiled_1M")
    el = bench.substoichiometric(el, p)
    dev = D.Device([], p, arrays=(el, x, y, z))
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    return dict(p=p, dev=dev, sim=sim, buf=buf, nc=nc, nb=dev.neigh_idx.reshape(dev.N, -1))


def _charges(big):
    """site charges of the initial structure (device run), computed once"""
    if "q" not in big:
        big["dev"].updateCharge(big["buf"], big["p"].metals)
        big["q"] = big["buf"].site_charge.cpu().numpy()
    return big["q"]


def _potentials(big):
    """the 10 V state: phi_b and phi_c of the initial structure (device run), computed once"""
    if "pb" not in big:
        _charges(big)
        assert big["dev"].updatePotential(big["buf"], big["p"], 10.0, n_contact=big["nc"])["cg_converged"]
        big["pb"] = big["buf"].site_potential_boundary.cpu().numpy()
        big["pc"] = big["buf"].site_potential_charge.cpu().numpy()
    return big["pb"], big["pc"]


def test_fullsize_graph_charge_and_csr_bit_exact(big, O):
    p, dev, buf, nc = big["p"], big["dev"], big["buf"], big["nc"]
    assert dev.N == 1033890
    nb, nn = O.neighbor_list(dev.site_x, dev.site_y, dev.site_z, dev.lattice, p.pbc, p.nn_dist, method=1)
    assert nn == dev.max_num_neighbors == 51
    assert np.array_equal(big["nb"], nb)
    # undirected graph: every edge has its mirror
    i = np.repeat(np.arange(dev.N, dtype=np.int64), nn)[(nb >= 0).ravel()]; j = nb[nb >= 0].astype(np.int64)
    assert np.array_equal(np.sort(i * dev.N + j), np.sort(j * dev.N + i))
    q = _charges(big)
    assert np.array_equal(q, O.update_charge(nb, dev.site_element, p.metals, np.zeros(dev.N, np.int32)))
    sp = buf.sparsity(nc, nc)
    ref = O.csr_structure(nb, nc, nc)
    from test_gpu_parity import _from_ptr
    import torch
    assert (sp.m, sp.nnz) == (dev.N - 2 * nc, len(ref["col"]))
    assert np.array_equal(_from_ptr(torch, sp.d_row_ptr, sp.m + 1), ref["row_ptr"])
    assert np.array_equal(_from_ptr(torch, sp.d_col, sp.nnz), ref["col"])


def test_fullsize_assembly_bit_exact_and_solve_properties(big, O):
    import scipy.sparse as ss
    import torch
    from devicekmc_b200._capi import check
    p, dev, buf, nc, nb = big["p"], big["dev"], big["buf"], big["nc"], big["nb"]
    q, csr = _charges(big), O.csr_structure(nb, nc, nc)
    sp = buf.sparsity(nc, nc)
    Vd = 10.0
    val = torch.zeros(sp.nnz, dtype=torch.float64, device="cuda"); rhs = torch.zeros(sp.m, dtype=torch.float64, device="cuda")
    check(dev.ctx.lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, Vd, p.high_G, p.low_G, buf.site_element.data_ptr(),
                                      buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(),
                                      rhs.data_ptr()))
    oval, orhs = O.assemble_K(nb, nc, nc, dev.site_element, q, p.metals, p.high_G, p.low_G, Vd, csr["row_ptr"], csr["col"])
    assert np.array_equal(val.cpu().numpy(), oval) and np.array_equal(rhs.cpu().numpy(), orhs)
    out = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert out["cg_converged"]
    phi = buf.site_potential_boundary.cpu().numpy()
    assert np.array_equal(phi[:nc], np.full(nc, -Vd / 2)) and np.array_equal(phi[-nc:], np.full(nc, Vd / 2))
    # property 1: the residual of the interior system, scaled as an error (D^-1 r), is at rounding level
    A = ss.csr_matrix((oval, csr["col"], csr["row_ptr"]), shape=(sp.m, sp.m))
    r = orhs - A @ phi[nc:-nc]
    assert np.abs(r / A.diagonal()).max() <= 1e-11 * np.abs(phi).max()
    # property 2: discrete maximum principle of the Kirchhoff system: all potentials between the contacts
    assert phi.min() >= -Vd / 2 * (1 + 1e-12) and phi.max() <= Vd / 2 * (1 + 1e-12)
    # property 3: linear in Vd (same charges, so the same matrix)
    buf.site_potential_boundary.zero_()
    assert dev.updatePotential(buf, p, -2.5 * Vd, n_contact=nc)["cg_converged"]
    phi2 = buf.site_potential_boundary.cpu().numpy()
    assert np.abs(phi2 + 2.5 * phi).max() <= TOL * np.abs(phi2).max()
    # back to 10 V from the -25 V solution: the same potential again (this is the state the next tests use)
    big.pop("pb", None)
    pb, _ = _potentials(big)
    assert np.abs(pb - phi).max() <= TOL * np.abs(phi).max()


def test_fullsize_pairwise_rows_and_superposition(big, O):
    from devicekmc_b200._capi import check
    p, dev, buf = big["p"], big["dev"], big["buf"]
    q, (_, pc) = _charges(big), _potentials(big)
    # oracle on bands of target rows (contacts, the oxide, the far end)
    for a, b in [(0, 400), (dev.N // 2, dev.N // 2 + 400), (dev.N - 400, dev.N)]:
        ref = O.poisson_gridless(dev.site_x, dev.site_y, dev.site_z, dev.lattice, p.pbc, q, p.sigma, p.k, rows=(a, b))
        nz = ref != 0
        assert (np.abs(pc[a:b][nz] - ref[nz]) <= TOL * np.abs(ref[nz])).all()
        assert np.abs(pc[a:b] - ref).max() <= 1e-13 * np.abs(pc).max()
    # superposition: phi_c(q) = phi_c(q+) + phi_c(q-)
    import torch
    lib, h = dev.ctx.lib, dev.ctx.h
    parts = []
    for sign in (1, -1):
        qs = torch.from_numpy(np.where(np.sign(q) == sign, q, 0).astype(np.int32)).cuda()
        out = torch.empty(dev.N, dtype=torch.float64, device="cuda")
        check(lib.dkmc_poisson_gridless(h, dev.pbc, dev.N, buf.lattice.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(),
                                        buf.site_x.data_ptr(), buf.site_y.data_ptr(), buf.site_z.data_ptr(), qs.data_ptr(),
                                        out.data_ptr()))
        parts.append(out.cpu().numpy())
    assert np.abs(parts[0] + parts[1] - pc).max() <= 1e-13 * np.abs(pc).max()


def test_fullsize_rate_table_and_first_events(big, O):
    import devicekmc_b200 as D
    from devicekmc_b200._capi import StepInfo, check, DKMC_ERR_RNG_EXHAUSTED
    p, dev, sim, buf, nb = big["p"], big["dev"], big["sim"], big["buf"], big["nb"]
    q, (pb, pc) = _charges(big), _potentials(big)
    o_type, o_prob = O.rate_table(nb, sim.site_layer, dev.lattice, dev.pbc, dev.T_bg, sim.freq, dev.sigma, dev.k, dev.site_x,
                                  dev.site_y, dev.site_z, pb, pc, dev.site_element, q, buf.E_host.T)
    import torch
    n = dev.N * buf.nn_
    et = torch.empty(n, dtype=torch.int32, device="cuda"); ep = torch.empty(n, dtype=torch.float64, device="cuda")
    lib, h = dev.ctx.lib, dev.ctx.h
    args = (h, dev.N, buf.nn_, buf.neigh_idx.data_ptr(), buf.site_layer.data_ptr(), buf.lattice.data_ptr(), dev.pbc,
            buf.T_bg.data_ptr(), buf.freq.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(), buf.site_x.data_ptr(),
            buf.site_y.data_ptr(), buf.site_z.data_ptr(), buf.site_potential_boundary.data_ptr(),
            buf.site_potential_charge.data_ptr(), buf.site_element.data_ptr(), buf.site_charge.data_ptr())
    check(lib.dkmc_build_event_list(*args, et.data_ptr(), ep.data_ptr()))
    et, ep = et.cpu().numpy(), ep.cpu().numpy()
    assert np.array_equal(et, o_type)
    nz = o_prob != 0
    assert np.array_equal(ep[~nz], o_prob[~nz]) and (np.abs(ep[nz] - o_prob[nz]) <= TOL * o_prob[nz]).all()
    # the first 25 events of the step (the step itself runs thousands): the device loop is given
    # exactly 50 uniforms of the reference stream and must stop, exhausted, after 25 events
    n_ev = 25
    rng = O.Rng(D.host.RND_SEED_KMC)
    O.set_event_limit(n_ev)
    try:
        _, ev_ref, el_ref, q_ref = O.kmc_events(nb, et, ep, dev.site_element, q, sim.freq, rng, max_events=n_ev)
    finally:
        O.set_event_limit(0)
    assert len(ev_ref) == n_ev
    u = D.RandomNumberGenerator(D.host.RND_SEED_KMC).getRandomNumbers(2 * n_ev)
    el0, q0 = buf.site_element.clone(), buf.site_charge.clone()
    import time
    # exact_mode 1: every selection goes through the exact sequential replay (the rare path of the
    # default mode) — at this size it walks ~10^4 non-zero rows of 10^6 per selection
    for exact_mode in (0, 1):
        buf.site_element.copy_(el0); buf.site_charge.copy_(q0)
        lib.dkmc_ctx_set_exact_select(h, exact_mode)
        info = StepInfo()
        ev = np.zeros((n_ev, 4), np.int32)
        t0 = time.perf_counter()
        try:
            st = lib.dkmc_execute_kmc_step(*args, u.ctypes.data_as(C.c_void_p), 2 * n_ev, ev.ctypes.data_as(C.c_void_p),
                                           n_ev, C.byref(info))
        finally:
            lib.dkmc_ctx_set_exact_select(h, 0)
        assert st == DKMC_ERR_RNG_EXHAUSTED and info.n_events == n_ev and info.n_used == 2 * n_ev
        assert np.array_equal(ev, ev_ref), exact_mode                    # (idx, i, j, type): bit-exact
        assert np.array_equal(buf.site_element.cpu().numpy(), el_ref)
        assert np.array_equal(buf.site_charge.cpu().numpy(), q_ref)
        if exact_mode == 1:
            assert info.n_exact_fallbacks == n_ev
            assert time.perf_counter() - t0 < 5.0, "the exact replay must stay a millisecond-scale path"
