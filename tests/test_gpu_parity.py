"""Parity of the CUDA path (through the C-ABI, include/dkmc.h) against the oracle and the golden
fixtures produced by the reference's own CPU build.  Integer outputs bit-exact; potentials and
rates within 1e-10 relative (BASELINE.json north_star); trajectories event for event."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10  # north_star: potentials and rates within 1e-10 relative


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def make_sim(case, pbc=None):
    import devicekmc_b200 as D
    p = case["p"]
    if pbc is not None:
        import dataclasses
        p = dataclasses.replace(p, pbc=pbc)
    dev = D.Device([], p, arrays=(case["element"], case["x"], case["y"], case["z"]))
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    return p, dev, sim, buf


@pytest.fixture(scope="module")
def sim0(base_case):
    return make_sim(base_case)


def rel_inf(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def elementwise_rel(a, b):
    nz = b != 0
    assert np.array_equal(a[~nz], b[~nz])
    return (np.abs(a[nz] - b[nz]) / np.abs(b[nz])).max() if nz.any() else 0.0


# ------------------------------------------------------------------ a1 neighbour graph
@pytest.mark.parametrize("pbc", [0, 1])
def test_neighbor_list_bit_exact(base_case, golden_step0, O, pbc):
    import hashlib
    p, dev, sim, buf = make_sim(base_case, pbc=pbc)
    nb, nn = O.neighbor_list(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], pbc, p.nn_dist, method=0)
    assert dev.max_num_neighbors == nn
    got = dev.neigh_idx.reshape(dev.N, nn)
    assert np.array_equal(got, nb)
    if pbc == 0:
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(golden_step0["neigh_sha"])


def test_neighbor_list_ragged_and_tiny(O):
    """degenerate inputs: coincident sites (dist 0, i != j must stay neighbours), an isolated
    site (empty row), N smaller than a warp"""
    import devicekmc_b200 as D
    rng = np.random.default_rng(0)
    x = np.array([0.0, 0.0, 1.0, 50.0, 2.0, 2.5, 3.0]); y = np.array([0.0, 0.0, 1.0, 50.0, 0.0, 0.5, 3.0])
    z = np.array([0.0, 0.0, 0.5, 50.0, 1.0, 0.0, 0.1])
    el = np.full(len(x), D.host.Hf_EL, np.int32)
    p = D.KMCParameters(lattice=(60.0, 60.0, 60.0))
    dev = D.Device([], p, arrays=(el, x, y, z))
    nb, nn = O.neighbor_list(x, y, z, p.lattice, 0, p.nn_dist, method=0)
    assert dev.max_num_neighbors == nn
    assert np.array_equal(dev.neigh_idx.reshape(len(x), nn), nb)
    assert (nb[3] == -1).all() and 1 in nb[0]


# ------------------------------------------------------------------ a2 CSR structure
def test_sparsity_bit_exact(sim0, O, torch):
    p, dev, sim, buf = sim0
    nc = p.num_atoms_contact
    sp = buf.sparsity(nc, nc)
    ref = O.csr_structure(dev.neigh_idx.reshape(dev.N, -1), nc, nc)
    m = dev.N - 2 * nc
    assert (sp.m, sp.nnz, sp.left_nnz, sp.right_nnz) == (m, len(ref["col"]), len(ref["left_col"]), len(ref["right_col"]))

    for name, ptr, n in [("row_ptr", sp.d_row_ptr, m + 1), ("col", sp.d_col, sp.nnz),
                         ("left_row_ptr", sp.d_left_row_ptr, m + 1), ("left_col", sp.d_left_col, sp.left_nnz),
                         ("right_row_ptr", sp.d_right_row_ptr, m + 1), ("right_col", sp.d_right_col, sp.right_nnz)]:
        got = _from_ptr(torch, ptr, n)
        assert np.array_equal(got, ref[name]), name


def _from_ptr(torch, ptr, n, dtype=np.int32):
    """raw device pointer -> numpy, through the CUDA runtime torch already loaded"""
    out = np.empty(n, dtype)
    if n == 0:
        return out
    torch.cuda.synchronize()
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    err = rt.cudaMemcpy(out.ctypes.data, ptr, out.nbytes, 2)  # cudaMemcpyDeviceToHost
    assert err == 0, err
    return out


# ------------------------------------------------------------------ a3 charge
def test_update_charge_bit_exact(sim0, O, golden_step0, torch):
    p, dev, sim, buf = sim0
    buf.sync_HostToGPU(dev)
    dev.updateCharge(buf, p.metals)
    torch.cuda.synchronize()
    q = buf.site_charge.cpu().numpy()
    ref = O.update_charge(dev.neigh_idx.reshape(dev.N, -1), dev.site_element, p.metals, np.zeros(dev.N, np.int32))
    assert np.array_equal(q, ref)
    assert np.array_equal(q, golden_step0["charge"].astype(np.int32))


# ------------------------------------------------------------------ a4/a5 assembly, SpMV, solve
def _assemble(sim0, O, torch, Vd):
    p, dev, sim, buf = sim0
    nc = p.num_atoms_contact
    lib = dev.ctx.lib
    dev.updateCharge(buf, p.metals)
    sp = buf.sparsity(nc, nc)
    val = torch.zeros(sp.nnz, dtype=torch.float64, device="cuda")
    rhs = torch.zeros(sp.m, dtype=torch.float64, device="cuda")
    from devicekmc_b200._capi import check
    check(lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, Vd, p.high_G, p.low_G, buf.site_element.data_ptr(),
                              buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(),
                              rhs.data_ptr()))
    torch.cuda.synchronize()
    q = buf.site_charge.cpu().numpy()
    csr = O.csr_structure(dev.neigh_idx.reshape(dev.N, -1), nc, nc)
    oval, orhs = O.assemble_K(dev.neigh_idx.reshape(dev.N, -1), nc, nc, dev.site_element, q, p.metals, p.high_G, p.low_G,
                              Vd, csr["row_ptr"], csr["col"])
    return sp, val, rhs, csr, oval, orhs


def test_assemble_K_bit_exact(sim0, O, torch):
    sp, val, rhs, csr, oval, orhs = _assemble(sim0, O, torch, 1.5)
    assert np.array_equal(val.cpu().numpy(), oval)      # same ascending-j sequential sums
    assert np.array_equal(rhs.cpu().numpy(), orhs)


def test_spmv_matches_csr(sim0, O, torch):
    p, dev, sim, buf = sim0
    sp, val, rhs, csr, oval, orhs = _assemble(sim0, O, torch, 1.5)
    import scipy.sparse as ss
    A = ss.csr_matrix((oval, csr["col"], csr["row_ptr"]), shape=(sp.m, sp.m))
    rng = np.random.default_rng(1)
    xh = rng.standard_normal(sp.m)
    x = torch.from_numpy(xh).cuda(); y = torch.empty_like(x)
    from devicekmc_b200._capi import check
    check(dev.ctx.lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), x.data_ptr(), y.data_ptr()))
    torch.cuda.synchronize()
    ref = A @ xh
    scale = np.abs(A) @ np.abs(xh)
    assert (np.abs(y.cpu().numpy() - ref) <= 4e-15 * scale).all()


@pytest.mark.parametrize("Vd", [1.5, 0.0, -3.0])
def test_background_potential(sim0, O, golden_step0, torch, Vd):
    p, dev, sim, buf = sim0
    nc = p.num_atoms_contact
    buf.sync_HostToGPU(dev)
    dev.updateCharge(buf, p.metals)
    buf.site_potential_boundary.zero_()
    out = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert out["cg_converged"]
    phi = buf.site_potential_boundary.cpu().numpy()
    q = buf.site_charge.cpu().numpy()
    ref, info = O.background_potential(dev.neigh_idx.reshape(dev.N, -1), nc, nc, dev.site_element, q, p.metals,
                                       p.high_G, p.low_G, Vd, refine=3)
    assert np.array_equal(phi[:nc], np.full(nc, -Vd / 2)) and np.array_equal(phi[-nc:], np.full(nc, Vd / 2))
    if Vd == 0.0:
        assert np.abs(phi).max() == 0.0
        return
    assert rel_inf(phi, ref) <= TOL
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    assert (np.abs(phi[big] - ref[big]) / np.abs(ref[big])).max() <= TOL
    if Vd == 1.5:
        # the reference's dense dgesv solution is itself ~2e-9 from the exact one
        assert rel_inf(phi, golden_step0["potential_boundary"]) <= 5e-9
    # warm start from the converged solution converges immediately
    out2 = dev.updatePotential(buf, p, Vd, n_contact=nc)
    assert out2["cg_iterations"] <= out["cg_iterations"]
    assert rel_inf(buf.site_potential_boundary.cpu().numpy(), ref) <= TOL


def test_background_potential_cluster_preconditioner(sim0, O, torch):
    """plain Jacobi (the reference's preconditioner) and Jacobi + cluster coarse space reach the
    same solution; the coarse space needs far fewer iterations"""
    import devicekmc_b200 as D
    p, dev, sim, buf = sim0
    nc = p.num_atoms_contact
    # plant uncharged-vacancy clusters: turn a few oxygen neighbours of vacancies into vacancies
    el = dev.site_element.copy()
    nb = dev.neigh_idx.reshape(dev.N, -1)
    vac = np.nonzero(el == D.host.VACANCY)[0][:40]
    for v in vac:
        for j in nb[v]:
            if j >= 0 and el[j] == D.host.O_EL:
                el[j] = D.host.VACANCY
                break
    saved = dev.site_element.copy()
    dev.site_element[...] = el
    try:
        buf.sync_HostToGPU(dev)
        dev.updateCharge(buf, p.metals)
        q = buf.site_charge.cpu().numpy()
        assert np.count_nonzero((el == D.host.VACANCY) & (q == 0)) >= 20
        ref, _ = O.background_potential(nb, nc, nc, el, q, p.metals, p.high_G, p.low_G, 8.0, refine=3)
        its = {}
        for flag in (1, 0):
            o = D.SolverOpts()
            dev.ctx.lib.dkmc_default_solver_opts(C.byref(o))
            o.cluster_precond = flag
            o.refine_rounds = 4
            buf.site_potential_boundary.zero_()
            out = dev.updatePotential(buf, p, 8.0, n_contact=nc, opts=o)
            assert out["cg_converged"]
            assert rel_inf(buf.site_potential_boundary.cpu().numpy(), ref) <= TOL, flag
            its[flag] = out["cg_iterations"]
        assert its[1] < its[0]
    finally:
        dev.site_element[...] = saved
        buf.sync_HostToGPU(dev)


# ------------------------------------------------------------------ 8f-3 CB-edge Laplace solve
def test_cb_edge_laplace(sim0, O, torch):
    """Device::setLaplacePotential / update_CB_edge_gpu_sparse: `metal || metal` rule, contacts at
    +-q Vd / 2, against the oracle and the fixture written by the reference's CPU build"""
    from conftest import GOLDEN
    p, dev, sim, buf = sim0
    g = np.load(os.path.join(GOLDEN, "s_cb_edge.npz"))
    Vd, nc = float(g["Vd"]), int(g["n_contact"])
    assert nc == p.num_atoms_first_layer
    dev.site_CB_edge[...] = 0.0
    out = dev.setLaplacePotential(buf, p, Vd)
    assert out["cg_converged"]
    cb = np.array(dev.site_CB_edge)                                    # synced back to the host, as the reference does
    assert np.array_equal(cb, buf.site_CB_edge.cpu().numpy())
    assert np.array_equal(cb[:nc], np.full(nc, p.q * Vd / 2)) and np.array_equal(cb[-nc:], np.full(nc, -p.q * Vd / 2))
    ref, _ = O.laplace_cb_edge(dev.neigh_idx.reshape(dev.N, -1), nc, nc, dev.site_element, p.metals, p.high_G, p.low_G, Vd)
    assert rel_inf(cb, ref) <= TOL
    assert rel_inf(cb, g["cb_edge"]) <= TOL
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    assert (np.abs(cb[big] - ref[big]) / np.abs(ref[big])).max() <= TOL
    # next bias point: warm start from the previous one (the reference's v_soln, potential_solver_gpu.cu:648)
    out2 = dev.setLaplacePotential(buf, p, -2.0 * Vd)
    assert out2["cg_converged"]
    assert rel_inf(np.array(dev.site_CB_edge), -2.0 * ref) <= TOL
    # the other per-step arrays are untouched by the bias-point solve
    buf.sync_HostToGPU(dev)


# ------------------------------------------------------------------ a6 pairwise
@pytest.mark.parametrize("pbc", [0, 1])
def test_poisson_gridless(base_case, O, golden_step0, torch, pbc):
    p, dev, sim, buf = make_sim(base_case, pbc=pbc)
    dev.updateCharge(buf, p.metals)
    dev.updatePotential(buf, p, 0.0, n_contact=p.num_atoms_contact)
    got = buf.site_potential_charge.cpu().numpy()
    q = buf.site_charge.cpu().numpy()
    ref = O.poisson_gridless(dev.site_x, dev.site_y, dev.site_z, dev.lattice, pbc, q, p.sigma, p.k)
    assert elementwise_rel(got, ref) <= TOL
    if pbc == 0:
        assert elementwise_rel(got, golden_step0["potential_charge"]) <= TOL


def test_poisson_gridless_no_charges(sim0, torch):
    p, dev, sim, buf = sim0
    buf.site_charge.zero_()
    buf.site_potential_charge.fill_(7.0)
    from devicekmc_b200._capi import check
    check(dev.ctx.lib.dkmc_poisson_gridless(dev.ctx.h, dev.pbc, dev.N, buf.lattice.data_ptr(), buf.sigma.data_ptr(),
                                            buf.k.data_ptr(), buf.site_x.data_ptr(), buf.site_y.data_ptr(),
                                            buf.site_z.data_ptr(), buf.site_charge.data_ptr(),
                                            buf.site_potential_charge.data_ptr()))
    assert float(buf.site_potential_charge.abs().max()) == 0.0


@pytest.mark.parametrize("pbc", [0, 1])
def test_poisson_gridless_incremental(base_case, O, torch, pbc):
    """opt-in 8f-2: phi_c updated by the charge DIFFERENCES since the previous step (cell-list kernel
    for pbc = 0, all-pairs kernel for pbc = 1), full sum every 4th call; against the oracle's full
    sum after every step of a short trajectory"""
    p, dev, sim, buf = make_sim(base_case, pbc=pbc)
    nc = p.num_atoms_contact
    dev.ctx.set_pairwise_incremental(4)
    scale = None
    for step in range(6):
        dev.updateCharge(buf, p.metals)
        dev.updatePotential(buf, p, 6.0, n_contact=nc)
        got = buf.site_potential_charge.cpu().numpy()
        q = buf.site_charge.cpu().numpy()
        ref = O.poisson_gridless(dev.site_x, dev.site_y, dev.site_z, dev.lattice, pbc, q, p.sigma, p.k)
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= 1e-13 * scale, step
        if step in (0, 4):
            assert elementwise_rel(got, ref) <= TOL          # a full sum: element-wise as in the default mode
        sim.executeKMCStep(buf, dev)
    assert dev.ctx.pairwise_incremental_counts() == (2, 4)
    # a different output array, or switching the mode off and on, falls back to a full sum
    dev.ctx.set_pairwise_incremental(0)
    dev.updateCharge(buf, p.metals)
    dev.updatePotential(buf, p, 6.0, n_contact=nc)
    assert dev.ctx.pairwise_incremental_counts() == (2, 4)


def test_trajectory_with_incremental_pairwise(base_case, torch):
    """the 6 V golden trajectory of the reference run, event for event, with phi_c carried
    incrementally from step to step (full sum every 3rd step)"""
    import hashlib
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "s_traj_6V.npz"))
    p, dev, sim, buf = make_sim(base_case)
    dev.ctx.set_pairwise_incremental(3)
    for s, Vd in enumerate(g["Vd"]):
        dev.updateCharge(buf, p.metals)
        assert dev.updatePotential(buf, p, float(Vd), n_contact=p.num_atoms_contact)["cg_converged"]
        t = sim.executeKMCStep(buf, dev, record_events=65536)
        assert np.array_equal(sim.last_events[:, 1:3], g["ev_ij"][g["ev_ptr"][s]:g["ev_ptr"][s + 1]]), f"step {s}"
        assert abs(t - g["step_time"][s]) <= 1e-7 * abs(g["step_time"][s]), f"step {s}"
        assert hashlib.sha256(buf.site_element.cpu().numpy().tobytes()).hexdigest() == str(g["el_sha"][s])
    assert dev.ctx.pairwise_incremental_counts() == (2, 4)


# ------------------------------------------------------------------ a7 rate table
def _load_golden_state(sim0, golden_step0, torch):
    p, dev, sim, buf = sim0
    dev.site_element[...] = golden_step0["element"].astype(np.int32)
    dev.site_charge[...] = golden_step0["charge"].astype(np.int32)
    dev.site_potential_boundary[...] = golden_step0["potential_boundary"]
    dev.site_potential_charge[...] = golden_step0["potential_charge"]
    buf.sync_HostToGPU(dev)
    torch.cuda.synchronize()


def test_rate_table(sim0, golden_step0, O, torch):
    p, dev, sim, buf = sim0
    _load_golden_state(sim0, golden_step0, torch)
    n = dev.N * buf.nn_
    et = torch.empty(n, dtype=torch.int32, device="cuda"); ep = torch.empty(n, dtype=torch.float64, device="cuda")
    from devicekmc_b200._capi import check
    check(dev.ctx.lib.dkmc_build_event_list(
        dev.ctx.h, dev.N, buf.nn_, buf.neigh_idx.data_ptr(), buf.site_layer.data_ptr(), buf.lattice.data_ptr(), dev.pbc,
        buf.T_bg.data_ptr(), buf.freq.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(), buf.site_x.data_ptr(),
        buf.site_y.data_ptr(), buf.site_z.data_ptr(), buf.site_potential_boundary.data_ptr(),
        buf.site_potential_charge.data_ptr(), buf.site_element.data_ptr(), buf.site_charge.data_ptr(), et.data_ptr(),
        ep.data_ptr()))
    et, ep = et.cpu().numpy(), ep.cpu().numpy()
    g_type = np.full(n, 4, np.int32); g_prob = np.zeros(n)
    g_type[golden_step0["ev_idx"]] = golden_step0["ev_type"]; g_prob[golden_step0["ev_idx"]] = golden_step0["ev_prob"]
    assert np.array_equal(et, g_type)                       # integer: bit-exact
    assert elementwise_rel(ep, g_prob) <= TOL
    o_type, o_prob = O.rate_table(dev.neigh_idx.reshape(dev.N, -1), sim.site_layer, dev.lattice, dev.pbc, dev.T_bg,
                                  sim.freq, dev.sigma, dev.k, dev.site_x, dev.site_y, dev.site_z,
                                  dev.site_potential_boundary, dev.site_potential_charge, dev.site_element,
                                  dev.site_charge, buf.E_host.T)
    assert np.array_equal(et, o_type) and elementwise_rel(ep, o_prob) <= TOL


# ------------------------------------------------------------------ scan + search primitive
@pytest.mark.parametrize("n", [1, 31, 2048, 2049, 479349, 3_000_001])
def test_inclusive_scan_and_select(sim0, O, torch, n):
    p, dev, sim, buf = sim0
    from devicekmc_b200._capi import check
    rng = np.random.default_rng(n)
    # integer-valued rates: every summation order is exact, so indices must agree bit for bit
    v = np.where(rng.random(n) < 0.03, rng.integers(1, 1000, n), 0).astype(np.float64)
    if v.sum() == 0:
        v[n // 2] = 5.0
    d_in = torch.from_numpy(v).cuda(); d_out = torch.empty_like(d_in)
    check(dev.ctx.lib.dkmc_inclusive_scan(dev.ctx.h, n, d_in.data_ptr(), d_out.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), np.cumsum(v))
    for u in [0.0, 0.25, 0.999999, rng.random()]:
        idx = C.c_longlong(-1); ps = C.c_double(0)
        check(dev.ctx.lib.dkmc_select_event(dev.ctx.h, n, d_out.data_ptr(), u, C.byref(idx), C.byref(ps)))
        oi, ops = O.select_event(v, u)
        assert (idx.value, ps.value) == (oi, ops)
    # general doubles: scan within a few ulp of the exact prefix sums
    w = rng.random(n) * np.exp(rng.normal(0, 5, n))
    d_in = torch.from_numpy(w).cuda()
    check(dev.ctx.lib.dkmc_inclusive_scan(dev.ctx.h, n, d_in.data_ptr(), d_out.data_ptr()))
    exact = np.cumsum(w.astype(np.longdouble)).astype(np.float64)
    assert rel_inf(d_out.cpu().numpy(), exact) <= 1e-13


# ------------------------------------------------------------------ a8 event loop
@pytest.mark.parametrize("exact_mode", [0, 1])
def test_kmc_step_events_match_oracle(sim0, golden_step0, O, torch, exact_mode):
    import devicekmc_b200 as D
    p, dev, sim, buf = sim0
    _load_golden_state(sim0, golden_step0, torch)
    # boost the field so that several events run in one step
    buf.site_potential_boundary.mul_(4.0)
    pb = buf.site_potential_boundary.cpu().numpy()
    o_type, o_prob = O.rate_table(dev.neigh_idx.reshape(dev.N, -1), sim.site_layer, dev.lattice, dev.pbc, dev.T_bg,
                                  sim.freq, dev.sigma, dev.k, dev.site_x, dev.site_y, dev.site_z, pb,
                                  dev.site_potential_charge, dev.site_element, dev.site_charge, buf.E_host.T)
    rng = O.Rng(D.host.RND_SEED_KMC)
    t_ref, ev_ref, el_ref, q_ref = O.kmc_events(dev.neigh_idx.reshape(dev.N, -1), o_type, o_prob, dev.site_element,
                                                dev.site_charge, sim.freq, rng)
    sim.random_generator.setSeed(D.host.RND_SEED_KMC)
    dev.ctx.lib.dkmc_ctx_set_exact_select(dev.ctx.h, exact_mode)
    try:
        t = sim.executeKMCStep(buf, dev, record_events=4096)
    finally:
        dev.ctx.lib.dkmc_ctx_set_exact_select(dev.ctx.h, 0)
    assert len(ev_ref) >= 2
    assert np.array_equal(sim.last_events, ev_ref)                     # (idx, i, j, type): bit-exact
    assert np.array_equal(buf.site_element.cpu().numpy(), el_ref)
    assert np.array_equal(buf.site_charge.cpu().numpy(), q_ref)
    assert abs(t - t_ref) <= 1e-12 * abs(t_ref)
    assert sim.last_info.n_used == 2 * (len(ev_ref))
    if exact_mode == 1:
        assert sim.last_info.n_exact_fallbacks == len(ev_ref)
    # the host generator was advanced exactly as the reference's would be
    assert sim.random_generator.getRandomNumber() == rng.uniform()


def test_kmc_step_no_events(sim0, torch):
    """all rates zero: upper_bound finds nothing, event_time = -ln(u)/0 = inf, nothing executed
    (the reference reads out of bounds here, SURVEY.md notes; we define it)"""
    p, dev, sim, buf = sim0
    dev.site_element[...] = np.where(np.isin(dev.site_element, [0, 1, 2]), 4, dev.site_element)  # no d/Od/V
    buf.sync_HostToGPU(dev)
    before = buf.site_element.clone()
    t = sim.executeKMCStep(buf, dev, record_events=16)
    assert np.isinf(t) and sim.last_info.n_events == 0 and sim.last_info.n_used == 2
    assert torch.equal(before, buf.site_element)


def test_kmc_step_continue_on_exhausted_uniforms(sim0, golden_step0, O, torch):
    import devicekmc_b200 as D
    p, dev, sim, buf = sim0
    _load_golden_state(sim0, golden_step0, torch)
    buf.site_potential_boundary.mul_(4.0)
    sim.random_generator.setSeed(D.host.RND_SEED_KMC)
    t_full = sim.executeKMCStep(buf, dev, record_events=4096)
    ev_full = sim.last_events.copy()
    _load_golden_state(sim0, golden_step0, torch)
    buf.site_potential_boundary.mul_(4.0)
    sim.random_generator.setSeed(D.host.RND_SEED_KMC)
    sim.batch_uniforms = 2                                   # one event per call
    try:
        t = sim.executeKMCStep(buf, dev, record_events=4096)
    finally:
        sim.batch_uniforms = 4096
    assert t == t_full and np.array_equal(sim.last_events, ev_full)


# ------------------------------------------------------------------ trajectories vs the reference run
@pytest.mark.parametrize("name", ["s_traj_ramp.npz", "s_traj_6V.npz", "s_traj_6V_pbc.npz"])
def test_trajectory_matches_reference(base_case, torch, name):
    import hashlib
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, name))
    pbc = int(g["pbc"]) if "pbc" in g.files else 0                  # s_traj_6V_pbc: periodic in y and z
    p, dev, sim, buf = make_sim(base_case, pbc=pbc)
    if pbc:
        assert np.array_equal((dev.neigh_idx.reshape(dev.N, -1) >= 0).sum(1), g["degree"].astype(np.int64))
    nc = p.num_atoms_contact
    for s, Vd in enumerate(g["Vd"]):
        dev.updateCharge(buf, p.metals)
        out = dev.updatePotential(buf, p, float(Vd), n_contact=nc)
        assert out["cg_converged"]
        if s == 0:
            assert rel_inf(buf.site_potential_boundary.cpu().numpy(), g["pb0"]) <= 5e-9 or float(Vd) == 0.0
            assert elementwise_rel(buf.site_potential_charge.cpu().numpy(), g["pc0"]) <= TOL
        t = sim.executeKMCStep(buf, dev, record_events=65536)
        ev_ref = g["ev_ij"][g["ev_ptr"][s]:g["ev_ptr"][s + 1]]
        assert np.array_equal(sim.last_events[:, 1:3], ev_ref), f"step {s}: executed events differ"
        assert abs(t - g["step_time"][s]) <= 1e-7 * abs(g["step_time"][s]), f"step {s}"
        el = buf.site_element.cpu().numpy()
        assert hashlib.sha256(el.tobytes()).hexdigest() == str(g["el_sha"][s])
    assert np.array_equal(buf.site_element.cpu().numpy(), g["element_last"].astype(np.int32))


def test_bias_point_loop_reproduces_the_reference_run(torch):
    """BASELINE config 1 through the bias-point loop (kmc_main.cpp:136-279): the shipped 500-point ramp
    (0 -> 12 V, t_switch 1e-3 s), 100 KMC steps — the same bias point for every step, the same events and the same
    KMC times as the reference's CPU run (s_traj_ramp100.npz), with the warm start rescaled by Vd / Vd_prev"""
    from conftest import GOLDEN
    import bench
    from devicekmc_b200 import slab, structures as S
    from devicekmc_b200.host import KMCParameters
    g = np.load(os.path.join(GOLDEN, "s_traj_ramp100.npz"))
    el, x, y, z, lat, nc = S.load_base_cell()
    p = KMCParameters(lattice=tuple(lat), num_atoms_contact=nc, num_atoms_first_layer=nc)
    el = bench.substoichiometric(el, p)
    sim = slab.SlabSim((el, x, y, z), p, 0, 1)
    V_switch = np.linspace(0.0, 12.0, 500)          # test_2.5nm/parameters.txt:41
    seen = []

    def on_step(st):
        seen.append((st["Vd"], sim.sim.last_events[:, 1:3].copy(), st["step_time"], st["cg_converged"]))
    n = len(g["Vd"])
    slab.bias_loop(sim, V_switch, 1e-3, max_steps=n, on_step=on_step, record_events=65536)
    assert len(seen) == n
    for s_, (Vd, ev, t, conv) in enumerate(seen):
        assert conv
        assert abs(Vd - float(g["Vd"][s_])) <= 1e-12, f"step {s_}: bias point"
        ev_ref = g["ev_ij"][g["ev_ptr"][s_]:g["ev_ptr"][s_ + 1]]
        assert np.array_equal(ev, ev_ref), f"step {s_}: executed events differ"
        assert abs(t - g["step_time"][s_]) <= 1e-7 * abs(g["step_time"][s_]), f"step {s_}"
    assert np.array_equal(sim.buf.site_element.cpu().numpy(), g["element_last"].astype(np.int32))


# ------------------------------------------------------------------ 8f-4 snapshots that do not stall the step
def test_snapshot_is_a_consistent_copy_while_the_step_goes_on(base_case, golden_step0, torch, tmp_path):
    import hashlib
    from conftest import GOLDEN
    sim1 = make_sim(base_case)
    p, dev, sim, buf = sim1
    _load_golden_state(sim1, golden_step0, torch)
    snap = buf.snapshot_begin(dev)
    # the state is mutated right away: events on the live arrays, a new potential
    sim.executeKMCStep(buf, dev)
    buf.site_potential_boundary.add_(1.0)
    buf.site_charge.zero_()
    with pytest.raises(Exception):
        buf.snapshot_begin(dev)                       # one snapshot in flight per context
    el, q, pot = snap.wait()
    assert np.array_equal(el, golden_step0["element"].astype(np.int32))
    assert np.array_equal(q, golden_step0["charge"].astype(np.int32))
    assert np.array_equal(pot, golden_step0["potential_boundary"] + golden_step0["potential_charge"])
    assert not np.array_equal(buf.site_element.cpu().numpy(), el)      # the live state did move on
    # the file is the one the reference's Device::writeSnapshot wrote for this state
    snap.write("snapshot_0.xyz", str(tmp_path))
    g = np.load(os.path.join(GOLDEN, "s_snapshot.npz"))
    assert hashlib.sha256((tmp_path / "snapshot_0.xyz").read_bytes()).hexdigest() == str(g["sha256"])
    # and a second snapshot works once the first has been waited for
    snap2 = buf.snapshot_begin(dev)
    el2, q2, _ = snap2.wait()
    assert np.array_equal(el2, buf.site_element.cpu().numpy()) and not q2.any()


# ------------------------------------------------------------------ window-staged SpMV + stream overlap
@pytest.fixture(scope="module")
def cell_sim():
    """3 x 4 tiles of the base cell in x-major cell order (the benchmark devices' order): the order in
    which the window-staged SpMV format applies.  Returns (p, dev, sim, buf, nc)."""
    import bench
    import devicekmc_b200 as D
    el, x, y, z, lat, nc, p = bench.workload("tiled_100k")
    el = bench.substoichiometric(el, p)
    dev = D.Device([], p, arrays=(el, x, y, z))
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    dev.updateCharge(buf, p.metals)
    return p, dev, sim, buf, nc


def test_spmv_reads_the_values_it_is_given(cell_sim, torch):
    """dkmc_spmv / dkmc_solve_cg take ANY CSR values: no matrix format is cached behind the caller's back
    (round 1 kept packed / window-staged copies keyed by the pointer; changing val in place went unnoticed)"""
    from devicekmc_b200._capi import check
    p, dev, sim, buf, nc = cell_sim
    lib = dev.ctx.lib
    sp = buf.sparsity(nc, nc)
    val = torch.zeros(sp.nnz, dtype=torch.float64, device="cuda")
    rhs = torch.zeros(sp.m, dtype=torch.float64, device="cuda")
    check(lib.dkmc_assemble_K(dev.ctx.h, C.byref(sp), dev.N, nc, nc, 10.0, p.high_G, p.low_G, buf.site_element.data_ptr(),
                              buf.site_charge.data_ptr(), buf.metal_types.data_ptr(), len(p.metals), val.data_ptr(),
                              rhs.data_ptr()))
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    x = torch.randn(sp.m, dtype=torch.float64, device="cuda", generator=g)
    y1, y2 = torch.empty_like(x), torch.empty_like(x)
    check(lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), x.data_ptr(), y1.data_ptr()))
    val.mul_(2.0)                                   # same address, other matrix
    check(lib.dkmc_spmv(dev.ctx.h, sp.m, sp.nnz, sp.d_row_ptr, sp.d_col, val.data_ptr(), x.data_ptr(), y2.data_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(y2, 2.0 * y1)


def test_persistent_pcg_matches_per_op_pcg(cell_sim, torch):
    """the persistent-kernel PCG (Chronopoulos-Gear recurrence, cluster sums by recurrence) and the round-1
    one-kernel-per-operation PCG solve the same system to the same refined accuracy"""
    from devicekmc_b200._capi import check
    p, dev, sim, buf, nc = cell_sim
    lib = dev.ctx.lib
    out = {}
    for legacy in (1, 0):
        check(lib.dkmc_ctx_set_legacy_cg(dev.ctx.h, legacy))
        buf.site_potential_boundary.zero_()
        o = dev.updatePotential(buf, p, 10.0, n_contact=nc, overlap=False)
        assert o["cg_converged"] and o["cg_est_error"] <= 1e-13
        out[legacy] = (buf.site_potential_boundary.clone(), o["cg_iterations"])
    a, b = out[1][0], out[0][0]
    assert float((a - b).abs().max()) <= 1e-12 * float(a.abs().max())
    assert out[0][1] <= 1.3 * out[1][1] + 20      # the rearranged recurrence costs no extra iterations to speak of


def test_potential_overlap_matches_serial(cell_sim, O, torch):
    """pairwise sum on the side stream concurrently with the CG == the two run one after the other;
    and both within 1e-10 of the oracle on the cell-ordered device"""
    p, dev, sim, buf, nc = cell_sim
    Vd = 10.0
    buf.site_potential_boundary.zero_()
    o1 = dev.updatePotential(buf, p, Vd, n_contact=nc, overlap=False)
    b1, c1 = buf.site_potential_boundary.clone(), buf.site_potential_charge.clone()
    buf.site_potential_boundary.zero_(); buf.site_potential_charge.zero_()
    o2 = dev.updatePotential(buf, p, Vd, n_contact=nc, overlap=True)
    assert o1["cg_converged"] and o2["cg_converged"]
    assert torch.equal(c1, buf.site_potential_charge)
    # the persistent PCG sizes its grid by what is resident beside the pairwise CTAs, so the partial sums of
    # its reductions are grouped differently in the two runs: same solution to rounding, not the same bits
    b2 = buf.site_potential_boundary
    assert float((b1 - b2).abs().max()) <= 1e-12 * float(b1.abs().max())
    nb = dev.neigh_idx.reshape(dev.N, -1)
    q = buf.site_charge.cpu().numpy()
    ref, _ = O.background_potential(nb, nc, nc, dev.site_element, q, p.metals, p.high_G, p.low_G, Vd, refine=3)
    assert rel_inf(b1.cpu().numpy(), ref) <= TOL
    rows = (dev.N // 2, dev.N // 2 + 2000)
    refc = O.poisson_gridless(dev.site_x, dev.site_y, dev.site_z, dev.lattice, p.pbc, q, p.sigma, p.k, rows=rows)
    assert rel_inf(c1.cpu().numpy()[rows[0]:rows[1]], refc) <= TOL


def test_pairwise_cell_list_and_cutoff(cell_sim, O, torch):
    """cell-list kernel (skips exact zeros) == all-pairs kernel to rounding; the opt-in 10-sigma
    truncation stays far inside the 1e-10 parity bound (norm-wise) and evaluates far fewer pairs"""
    from devicekmc_b200._capi import check
    p, dev, sim, buf, nc = cell_sim
    lib = dev.ctx.lib
    args = (dev.ctx.h, dev.pbc, dev.N, buf.lattice.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(),
            buf.site_x.data_ptr(), buf.site_y.data_ptr(), buf.site_z.data_ptr(), buf.site_charge.data_ptr())
    out = {}
    pairs = {}
    for name, cells, cut in (("all", 0, 0.0), ("cells", 1, 0.0), ("cut10", 1, 10.0)):
        check(lib.dkmc_ctx_set_pairwise_cells(dev.ctx.h, cells))
        check(lib.dkmc_ctx_set_pairwise_cutoff(dev.ctx.h, cut))
        buf.site_potential_charge.fill_(7.0)
        check(lib.dkmc_poisson_gridless(*args, buf.site_potential_charge.data_ptr()))
        out[name] = buf.site_potential_charge.clone()
        pe = C.c_longlong(0)
        check(lib.dkmc_pairwise_pairs_evaluated(dev.ctx.h, C.byref(pe)))
        pairs[name] = pe.value
    check(lib.dkmc_ctx_set_pairwise_cells(dev.ctx.h, 1))
    check(lib.dkmc_ctx_set_pairwise_cutoff(dev.ctx.h, 0.0))
    scale = float(out["all"].abs().max())
    assert scale > 0
    assert float((out["cells"] - out["all"]).abs().max()) <= 1e-13 * scale
    big = out["all"].abs() > 1e-6 * scale
    assert float(((out["cells"] - out["all"]).abs()[big] / out["all"].abs()[big]).max()) <= 1e-12
    assert float((out["cut10"] - out["all"]).abs().max()) <= 1e-15 * scale
    assert pairs["cut10"] < pairs["cells"] <= dev.N * int((buf.site_charge != 0).sum().item())
