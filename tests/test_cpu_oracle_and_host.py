"""CPU-side tests (run with -m "not gpu"): the oracle against the golden vectors produced by the
reference's own CPU build, the host-side logic, and the C-ABI library's exported symbols.
No compute call into libdkmc_b200.so happens here (there is no GPU in this container)."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def graph(base_case, O):
    p = base_case["p"]
    nb, nn = O.neighbor_list(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], p.pbc, p.nn_dist, method=1)
    return nb, nn


# ------------------------------------------------------------------ oracle vs golden (reference run)
def test_oracle_neighbor_list_matches_reference(base_case, golden_step0, graph, O):
    nb, nn = graph
    assert nn == int(golden_step0["nn"]) == 51
    assert hashlib.sha256(np.ascontiguousarray(nb).tobytes()).hexdigest() == str(golden_step0["neigh_sha"])
    assert np.array_equal((nb >= 0).sum(1), golden_step0["degree"].astype(np.int64))
    # brute force (the reference's own loop order) == cell list, also with periodic y/z
    p = base_case["p"]
    for pbc in (0, 1):
        a, na = O.neighbor_list(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], pbc, p.nn_dist, method=0)
        b, nb_ = O.neighbor_list(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], pbc, p.nn_dist, method=1)
        assert na == nb_ and np.array_equal(a, b)


def test_oracle_charge_and_layers_match_reference(base_case, golden_step0, graph, O):
    from devicekmc_b200.host import DEFAULT_LAYERS
    nb, nn = graph
    p = base_case["p"]
    q = O.update_charge(nb, base_case["element"], p.metals, np.zeros(len(nb), np.int32))
    assert np.array_equal(q, golden_step0["charge"].astype(np.int32))
    layer = O.site_layers(base_case["x"], [l.start_x for l in DEFAULT_LAYERS], [l.end_x for l in DEFAULT_LAYERS])
    assert np.array_equal(layer, golden_step0["site_layer"].astype(np.int32))
    E = np.array([[l.E_gen_0, l.E_rec_1, l.E_diff_2, l.E_diff_3] for l in DEFAULT_LAYERS])
    assert np.array_equal(E, golden_step0["layers"])


def test_oracle_pairwise_bit_exact_vs_reference(base_case, golden_step0, O):
    p = base_case["p"]
    q = golden_step0["charge"].astype(np.int32)
    pc = O.poisson_gridless(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], p.pbc, q, p.sigma, p.k)
    assert np.array_equal(pc, golden_step0["potential_charge"])
    rows = O.poisson_gridless(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], p.pbc, q, p.sigma, p.k,
                              rows=(100, 300))
    assert np.array_equal(rows, golden_step0["potential_charge"][100:300])


def test_oracle_background_potential_vs_reference_dgesv(base_case, golden_step0, graph, O):
    """The reference's dense LU is ~2e-9 (norm-wise) from the exact solution of its own matrix:
    cond(K) ~ 1e7 from uncharged-vacancy clusters.  The oracle's binary128-residual
    refinement converges to 1e-16, so the two agree only to the reference's own accuracy."""
    nb, nn = graph
    p = base_case["p"]
    nc = int(golden_step0["n_contact"])
    q = golden_step0["charge"].astype(np.int32)
    Vd = float(golden_step0["Vd"])
    phi, info = O.background_potential(nb, nc, nc, base_case["element"], q, p.metals, p.high_G, p.low_G, Vd, refine=3)
    ref = golden_step0["potential_boundary"]
    assert np.abs(phi - ref).max() / np.abs(ref).max() <= 5e-9
    assert info[2] <= 1e-13                      # binary128 residual of the oracle's solution
    # the refinement has converged: one more round changes nothing at 1e-14
    phi2, _ = O.background_potential(nb, nc, nc, base_case["element"], q, p.metals, p.high_G, p.low_G, Vd, refine=5)
    assert np.abs(phi - phi2).max() / np.abs(phi2).max() <= 1e-14
    # and the oracle's solution has the SMALLER binary128 residual of the two
    csr = O.csr_structure(nb, nc, nc)
    val, rhs = O.assemble_K(nb, nc, nc, base_case["element"], q, p.metals, p.high_G, p.low_G, Vd, csr["row_ptr"], csr["col"])
    _, i_ref = O.solve(csr["row_ptr"], csr["col"], val, rhs, x0=ref[nc:-nc], max_iter=0, refine=0)
    _, i_orc = O.solve(csr["row_ptr"], csr["col"], val, rhs, x0=phi[nc:-nc], max_iter=0, refine=0)
    assert i_orc[2] <= i_ref[2]


def test_oracle_cb_edge_vs_reference_dgesv(base_case, graph, O):
    """Device::setLaplacePotential (potential_solver.cpp:4-139), the `metal || metal`
    rule and the +-q Vd / 2 contacts, against the fixture written by the reference's CPU build.
    No vacancy clusters under this rule, so the reference's dgesv is accurate here (cond ~ 1e3)."""
    g = np.load(os.path.join(GOLDEN, "s_cb_edge.npz"))
    nb, nn = graph
    p = base_case["p"]
    nc = int(g["n_contact"])
    cb, info = O.laplace_cb_edge(nb, nc, nc, base_case["element"], p.metals, p.high_G, p.low_G, float(g["Vd"]))
    ref = g["cb_edge"]
    assert np.array_equal(cb[:nc], ref[:nc]) and np.array_equal(cb[-nc:], ref[-nc:])   # contacts: +-q Vd / 2
    assert cb[0] == 1.60217663e-19 * float(g["Vd"]) / 2 and cb[-1] == -cb[0]
    assert np.abs(cb - ref).max() / np.abs(ref).max() <= 1e-12
    # the charges do not enter: vacancies (uncharged or not) change nothing
    el2 = base_case["element"].copy()
    el2[el2 == 3] = 2
    cb2, _ = O.laplace_cb_edge(nb, nc, nc, el2, p.metals, p.high_G, p.low_G, float(g["Vd"]))
    assert np.array_equal(cb, cb2)


def test_oracle_rate_table_bit_exact_vs_reference(base_case, golden_step0, graph, O):
    nb, nn = graph
    p = base_case["p"]
    et, ep = O.rate_table(nb, golden_step0["site_layer"].astype(np.int32), base_case["lattice"], p.pbc, p.background_temp,
                          p.freq, p.sigma, p.k, base_case["x"], base_case["y"], base_case["z"],
                          golden_step0["potential_boundary"], golden_step0["potential_charge"], base_case["element"],
                          golden_step0["charge"].astype(np.int32), golden_step0["layers"])
    g_type = np.full(len(et), 4, np.int32); g_prob = np.zeros(len(ep))
    g_type[golden_step0["ev_idx"]] = golden_step0["ev_type"]; g_prob[golden_step0["ev_idx"]] = golden_step0["ev_prob"]
    assert np.array_equal(et, g_type) and np.array_equal(ep, g_prob)


def test_oracle_rates_with_oxygen_ions_bit_exact_vs_reference(base_case, graph, O):
    """crafted state (tests/golden/make_golden.py rates_ions): 80 oxygen ions in the oxide, so that the
    table holds recombination (type 1) and ion-diffusion (type 3) entries as well — all four rate
    formulas of KMCProcess.cpp:67-164 against the reference's own table, bit for bit"""
    from devicekmc_b200.host import DEFAULT_LAYERS
    g = np.load(os.path.join(GOLDEN, "s_rates_ions.npz"))
    nb, nn = graph
    p = base_case["p"]
    el = g["element"].astype(np.int32)
    assert np.count_nonzero(el == 1) == 80
    q = O.update_charge(nb, el, p.metals, np.zeros(len(el), np.int32))
    assert np.array_equal(q, g["charge"].astype(np.int32))
    assert (q[el == 1] == -2).any()
    layer = O.site_layers(base_case["x"], [l.start_x for l in DEFAULT_LAYERS], [l.end_x for l in DEFAULT_LAYERS])
    E = np.array([[l.E_gen_0, l.E_rec_1, l.E_diff_2, l.E_diff_3] for l in DEFAULT_LAYERS])
    et, ep = O.rate_table(nb, layer, base_case["lattice"], p.pbc, p.background_temp, p.freq, p.sigma, p.k, base_case["x"],
                          base_case["y"], base_case["z"], g["potential_boundary"], g["potential_charge"], el, q, E)
    g_type = np.full(len(et), 4, np.int32); g_prob = np.zeros(len(ep))
    g_type[g["ev_idx"]] = g["ev_type"]; g_prob[g["ev_idx"]] = g["ev_prob"]
    assert set(np.unique(g["ev_type"]).tolist()) == {0, 1, 2, 3}
    assert np.array_equal(et, g_type) and np.array_equal(ep, g_prob)
    # the pairwise potential of this state (negative and positive charges) is bit-exact as well
    pc = O.poisson_gridless(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], p.pbc, q, p.sigma, p.k)
    assert np.array_equal(pc, g["potential_charge"])


def test_rng_stream_matches_reference(golden_step0, O):
    from devicekmc_b200.host import RandomNumberGenerator, RND_SEED_KMC
    want = golden_step0["kmc_rng_first16"]
    o = O.Rng(RND_SEED_KMC)
    assert np.array_equal(o.uniforms(16), want)
    r = RandomNumberGenerator(RND_SEED_KMC)
    assert np.array_equal(r.peek(16), want)          # peek does not consume
    assert np.array_equal(r.getRandomNumbers(16), want)
    r2 = RandomNumberGenerator(RND_SEED_KMC); r2.advance(5)
    assert r2.getRandomNumber() == want[5]


@pytest.mark.parametrize("name", ["s_traj_ramp.npz", "s_traj_6V.npz", "s_traj_6V_pbc.npz", "s_traj_10V.npz"])
def test_oracle_trajectory_matches_reference(base_case, graph, O, name):
    """the kmc_main.cpp:175-279 loop with the oracle's stages, event for event against the
    reference run (the potentials differ by the reference's own dgesv error, see above);
    s_traj_6V_pbc: the same device periodic in y and z (pbc = 1)"""
    from devicekmc_b200.host import DEFAULT_LAYERS, RND_SEED_KMC
    g = np.load(os.path.join(GOLDEN, name))
    p = base_case["p"]
    pbc = int(g["pbc"]) if "pbc" in g.files else 0
    if pbc:
        nb, nn = O.neighbor_list(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], pbc, p.nn_dist, method=1)
        assert np.array_equal((nb >= 0).sum(1), g["degree"].astype(np.int64))
        assert (nb >= 0).sum() > (graph[0] >= 0).sum()               # the periodic images add neighbours
    else:
        nb, nn = graph
    nc = p.num_atoms_contact
    el = base_case["element"].copy()
    q = np.zeros(len(el), np.int32)
    layer = O.site_layers(base_case["x"], [l.start_x for l in DEFAULT_LAYERS], [l.end_x for l in DEFAULT_LAYERS])
    E = np.array([[l.E_gen_0, l.E_rec_1, l.E_diff_2, l.E_diff_3] for l in DEFAULT_LAYERS])
    rng = O.Rng(RND_SEED_KMC)
    phi = np.zeros(len(el))
    executed = set()
    for s, Vd in enumerate(g["Vd"][:8]):
        q = O.update_charge(nb, el, p.metals, q)
        phi, _ = O.background_potential(nb, nc, nc, el, q, p.metals, p.high_G, p.low_G, float(Vd), phi0=phi, refine=1)
        pc = O.poisson_gridless(base_case["x"], base_case["y"], base_case["z"], base_case["lattice"], pbc, q, p.sigma, p.k)
        if s == 0 and pbc:
            assert np.array_equal(pc, g["pc0"])                       # periodic pairwise sum: bit-exact
        et, ep = O.rate_table(nb, layer, base_case["lattice"], pbc, p.background_temp, p.freq, p.sigma, p.k,
                              base_case["x"], base_case["y"], base_case["z"], phi, pc, el, q, E)
        t, ev, el, q = O.kmc_events(nb, et, ep, el, q, p.freq, rng)
        assert np.array_equal(ev[:, 1:3], g["ev_ij"][g["ev_ptr"][s]:g["ev_ptr"][s + 1]]), f"step {s}"
        assert abs(t - g["step_time"][s]) <= 1e-7 * abs(g["step_time"][s])
        assert hashlib.sha256(el.tobytes()).hexdigest() == str(g["el_sha"][s])
        executed.update(ev[:, 3].tolist())
    if name == "s_traj_ramp.npz":
        assert {0, 1, 2} <= executed                 # generation, recombination, vacancy diffusion
    if name == "s_traj_10V.npz":
        assert g["ev_ptr"][1] >= 20                  # many events inside one step (stale table, conflicts zeroed)


def test_oracle_event_selection_edge_cases(O):
    # all-zero table: upper_bound finds nothing
    idx, ps = O.select_event(np.zeros(10), 0.3)
    assert idx == 10 and ps == 0.0
    # zeros between non-zeros never win; u = 0 picks the first non-zero entry
    v = np.array([0.0, 2.0, 0.0, 0.0, 3.0, 0.0])
    assert O.select_event(v, 0.0)[0] == 1
    assert O.select_event(v, 0.39)[0] == 1 and O.select_event(v, 0.41)[0] == 4
    # an entry absorbed by rounding (1e-30 next to 1.0) is never selected by the sequential sum
    w = np.array([1.0, 1e-30, 1.0])
    assert O.select_event(w, 0.5)[0] == 2


# ------------------------------------------------------------------ live oracle vs the reference build
def test_oracle_vs_reference_build_live(O):
    """where /root/reference and oracle/_ref exist (this container): the restatement against the
    UNMODIFIED reference objects on a state the fixtures do not cover (pbc = 1 distances)"""
    from oracle import refsim
    if not refsim.available() or not os.path.isdir("/root/reference"):
        pytest.skip("reference build not present (GPU box)")
    rng = np.random.default_rng(5)
    lat = np.array([100.0, 25.575, 25.575])
    for _ in range(200):
        a, b = rng.uniform(-30, 60, 3), rng.uniform(-30, 60, 3)
        for pbc in (0, 1):
            assert O.site_dist(a, b, lat, pbc) == refsim.site_dist(a, b, lat, pbc)


# ------------------------------------------------------------------ host logic
def test_parameters_parser_and_layers():
    from devicekmc_b200.host import KMCParameters, Ti_EL, N_EL
    txt = os.path.join(ROOT, "tests", "golden", "parameters_2.5nm.txt")
    p = KMCParameters.from_file(txt)
    assert p.freq == 1e14 and p.metals == (Ti_EL, N_EL) and p.nn_dist == 3.5 and p.pbc == 0
    assert p.num_atoms_contact == 144 and p.rnd_seed == 4 and abs(p.k - 8.987552e9 / 23.0) == 0
    assert len(p.V_switch) == len(p.t_switch) == 3 and p.V_switch[1] == 0.0240480961923848


def test_snapshot_file_matches_reference_writer(base_case, golden_step0, tmp_path):
    """the snapshot writer against the file Device::writeSnapshot (Device.cpp:236-252)
    of the reference build wrote for the s_step0 state — byte for byte"""
    from devicekmc_b200.host import write_snapshot
    g = np.load(os.path.join(GOLDEN, "s_snapshot.npz"))
    path = tmp_path / "snapshot_0.xyz"
    write_snapshot(str(path), base_case["element"], base_case["x"], base_case["y"], base_case["z"],
                   golden_step0["potential_boundary"] + golden_step0["potential_charge"])
    raw = path.read_bytes()
    lines = raw.decode().split("\n")
    assert lines[:6] == g["head"].tolist() and lines[-4:] == g["tail"].tolist()
    assert lines[3002:3008] == g["sample"].tolist()
    assert len(raw) == int(g["size"])
    assert hashlib.sha256(raw).hexdigest() == str(g["sha256"])


def test_substoichiometric_draw_matches_reference_for_other_seeds():
    """bench.substoichiometric (the host-side restatement of Device::makeSubstoichiometric that prepares every
    benchmark device) against the reference's own draw for seeds / concentrations other than the shipped ones"""
    import dataclasses
    import bench
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import KMCParameters, VACANCY
    g = np.load(os.path.join(GOLDEN, "s_substoich.npz"))
    el0 = S.load_base_cell()[0]
    for seed, conc, want, nv in zip(g["seed"], g["conc"], g["sha"], g["n_vacancies"]):
        p = dataclasses.replace(KMCParameters(), rnd_seed=int(seed), initial_vacancy_concentration=float(conc))
        el = bench.substoichiometric(el0, p)
        assert int((el == VACANCY).sum()) == int(nv)
        assert hashlib.sha256(np.ascontiguousarray(el, np.int32).tobytes()).hexdigest() == str(want), (seed, conc)


def test_tiled_device_structure(O):
    from devicekmc_b200 import structures as S
    el, x, y, z, lat, nc = S.tile_device(2, 2)
    assert len(x) == 4 * 9399 and nc == 4 * 144
    assert np.unique(x[:nc]).size == 1 and np.unique(x[-nc:]).size == 1       # contacts first / last
    nb, nn = O.neighbor_list(x, y, z, lat, 0, 3.5)
    assert nn == 51
    deg = (nb >= 0).sum(1)
    assert deg.min() >= 1
    i = np.repeat(np.arange(len(x)), nn)[(nb >= 0).ravel()]; j = nb[nb >= 0]
    assert set(zip(i.tolist()[:5000], j.tolist()[:5000])) <= set(zip(j.tolist(), i.tolist()))  # symmetric graph
    # every ordering is a permutation of the same sites
    e2, x2, y2, z2, _, _ = S.tile_device(2, 2, order="tile")
    assert sorted(zip(x.round(6), y.round(6), z.round(6))) == sorted(zip(x2.round(6), y2.round(6), z2.round(6)))


# ------------------------------------------------------------------ C-ABI surface
def test_capi_library_exports_every_declared_symbol():
    from devicekmc_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "dkmc.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char \*)\s*(dkmc_[A-Za-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = C.CDLL(_capi.LIB_PATH)          # loads without a GPU; no compute call is made
    for name in declared:
        assert hasattr(lib, name), name
    lib.dkmc_version.restype = C.c_int
    assert lib.dkmc_version() == 100


def test_capi_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/dkmc.h must compile as C11 (no C++ or torch types in any
    signature) and link against the library from a C program"""
    import subprocess
    from devicekmc_b200 import _capi
    src = tmp_path / "cabi.c"
    src.write_text('#include "dkmc.h"\n'
                   'int main(void) { dkmc_ctx *c = 0; (void)c; return dkmc_version() == DKMC_VERSION ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    exe = tmp_path / "cabi"
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ldkmc_b200",
                           f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"])
    assert subprocess.run([str(exe)]).returncode == 0          # dkmc_version() needs no GPU


def test_cpp_example_builds_and_fails_loudly_without_gpu(base_case, tmp_path):
    """examples/kmc_loop.cpp: every call of the loop type-checks against include/dkmc.h, the program links,
    and without a GPU it stops at dkmc_ctx_create with the library's error text (no CPU fallback)"""
    import subprocess
    import torch
    from conftest import build_cpp_example
    from devicekmc_b200.host import write_xyz
    exe = build_cpp_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: tests/test_gpu_zz_ions.py runs the example")
    xyz = tmp_path / "device.xyz"
    write_xyz(str(xyz), base_case["element"], base_case["x"], base_case["y"], base_case["z"])
    r = subprocess.run([exe, str(xyz), "108.97557", "25.575", "25.575", "144", "6.0", "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "dkmc_ctx_create" in r.stderr and "no CUDA device" in r.stderr, (r.stdout, r.stderr)


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import devicekmc_b200 as D
    with pytest.raises(Exception):
        D.Context()
    src = "".join(open(os.path.join(ROOT, "devicekmc_b200", f)).read() for f in ("host.py", "_capi.py", "structures.py", "slab.py")
                  if os.path.exists(os.path.join(ROOT, "devicekmc_b200", f)))
    assert "oracle" not in src.replace("# oracle", ""), "the product path must never import the oracle"


def test_oracle_matches_reference_build_on_the_28k_tile():
    """the sparse oracle pinned against the reference's own CPU build beyond the 9 k-site cell — the base
    cell tiled 1 x 3 (28 197 sites, the largest the reference's dense K + dgesv handles comfortably here), state
    with uncharged-vacancy clusters, 6 V.  Integers bit-exact; phi_c bit-exact; phi_b to what the reference's LU
    itself delivers (its dgesv is ~1e-9 off its own matrix, see test_background_potential_matches_reference)."""
    import os
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "s_tile28k.npz")
    if not os.path.exists(path):
        pytest.skip("s_tile28k.npz not generated (tests/golden/make_golden.py tile28k)")
    import hashlib
    from oracle import oracle as O
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import KMCParameters, VACANCY
    g = np.load(path)
    el0, x, y, z, lat, nc = S.tile_device(1, 3, order="tile")
    assert len(x) == int(g["N"]) == 28197 and nc == int(g["n_contact"])
    p = KMCParameters(lattice=tuple(lat), num_atoms_contact=nc, num_atoms_first_layer=nc)
    nb, nn = O.neighbor_list(x, y, z, lat, p.pbc, p.nn_dist, method=1)
    assert nn == int(g["nn"])
    assert hashlib.sha256(np.ascontiguousarray(nb).tobytes()).hexdigest() == str(g["neigh_sha"])
    el = g["element"].astype(np.int32)                      # after the reference's own 5 % draw
    import bench
    assert np.array_equal(bench.substoichiometric(el0, p), el)
    q = O.update_charge(nb, el, p.metals, np.zeros(len(x), np.int32))
    assert np.array_equal(q, g["charge"].astype(np.int32))
    unch = (el == VACANCY) & (q == 0)
    clustered = unch & np.where(nb >= 0, unch[np.clip(nb, 0, None)], False).any(axis=1)
    assert clustered.sum() >= 4                             # the state the cluster coarse space exists for
    Vd = float(g["Vd"])
    pb, _ = O.background_potential(nb, nc, nc, el, q, p.metals, p.high_G, p.low_G, Vd, refine=3)
    assert np.abs(pb - g["potential_boundary"]).max() <= 5e-9 * np.abs(g["potential_boundary"]).max()
    pc = O.poisson_gridless(x, y, z, lat, p.pbc, q, p.sigma, p.k)
    assert np.array_equal(pc, g["potential_charge"])
