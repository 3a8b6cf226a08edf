"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference CPU build
(oracle/_ref/libdkmc_ref.so, see oracle/Makefile).  Runs only where /root/reference exists.

    python tests/golden/make_golden.py

Fixtures (config 0 of BASELINE.json: structures/single_devices/test_2.5nm, rnd_seed = 4,
rnd_seed_kmc = 1, CPU build, 1 process):

  s_step0.npz      state after Device ctor + makeSubstoichiometric + updateCharge + potentials at
                   Vd = 1.5 V + the rate table (non-zero entries) — every stage of one step
  s_traj_ramp.npz  the kmc_main.cpp:136-279 loop on the shipped V_switch ramp, first 12 KMC steps:
                   per step Vd, executed (i, j) pairs, step time, sha256 of site_element/charge
  s_traj_ramp100.npz  the same, the full config-1 window: 100 KMC steps (77 bias points)
  s_traj_6V.npz    the same loop at constant Vd = 6 V, 6 KMC steps (many events per step)
  s_substoich.npz  Device::makeSubstoichiometric (Device.cpp:202-233) for other seeds / concentrations than the
                   shipped (4, 0.05): sha256 of site_element after the reference's draw
  s_rates_ions.npz crafted state: 80 interstitial sites of the oxide turned into oxygen ions (Od) next to the
                   5 % vacancies, then updateCharge + potentials at 3 V + the rate table: pins the recombination
                   and ion-diffusion rates (event types 1 and 3), which short trajectories hardly reach
  s_traj_10V.npz   the same at constant 10 V, 8 KMC steps: generation, then diffusion of the generated
                   oxygen ions (the one event type the other trajectories never execute)
  s_traj_6V_pbc.npz  the same at constant 6 V with pbc = 1 (periodic in y and z), 4 KMC steps
  s_snapshot.npz   Device::writeSnapshot (Device.cpp:236-252) of the s_step0 state: sha256 and size of the
                   file, its first and last lines
  s_tile28k.npz    SURVEY 8c's "~25 k-site tile": the base cell tiled 1 x 3 (28 197 sites, the reference's site order
                   copy after copy), where the reference's dense K (6.4 GB) and dgesv still run: element after the
                   reference's own 5 % draw, charges, phi_b at 6 V (dense LU), phi_c — pins the sparse oracle and the
                   GPU path against the reference beyond the 9 k-site cell, on a state with uncharged-vacancy clusters
  s_cb_edge.npz    Device::setLaplacePotential (CPU branch, potential_solver.cpp:4-139) at Vd = 1.5 V:
                   site_CB_edge of all sites
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refsim as R  # noqa: E402

REF = "/root/reference/structures/single_devices/test_2.5nm/"
OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def new_sim(pbc=0):
    params = REF + "parameters.txt"
    if pbc:
        import re
        import tempfile
        txt, n = re.subn(r"pbc = 0", "pbc = 1", open(params).read())
        assert n == 1
        f = tempfile.NamedTemporaryFile("w", suffix="_parameters.txt", delete=False)
        f.write(txt); f.close()
        params = f.name
    return R.RefSim(params, REF + "reordered_device_2.5.xyz")


def tile28k():
    import re
    import tempfile
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import write_xyz
    el, x, y, z, lat, nc = S.tile_device(1, 3, order="tile")
    d = tempfile.mkdtemp()
    xyz = os.path.join(d, "tile_1x3.xyz")
    write_xyz(xyz, el, x, y, z)
    txt = open(REF + "parameters.txt").read()
    txt, n1 = re.subn(r"lattice = [^\n/]*", "lattice = %r %r %r " % (float(lat[0]), float(lat[1]), float(lat[2])), txt, count=1)
    txt, n2 = re.subn(r"num_atoms_first_layer = \d+", "num_atoms_first_layer = %d" % nc, txt, count=1)
    txt, n3 = re.subn(r"num_atoms_contact = \d+", "num_atoms_contact = %d" % nc, txt, count=1)
    assert n1 == n2 == n3 == 1
    par = os.path.join(d, "parameters.txt")
    open(par, "w").write(txt)
    s = R.RefSim(par, xyz)
    assert s.N == len(x) and s.num_atoms_contact == nc
    nb = s.neigh_idx()
    el1 = s.element()
    s.update_charge()
    q = s.charge()
    Vd = 6.0
    s.background_potential(Vd)
    s.poisson_gridless()
    pb, pc = s.potential_boundary(), s.potential_charge()
    np.savez_compressed(os.path.join(OUT, "s_tile28k.npz"), N=s.N, nn=s.nn, Vd=Vd, n_contact=nc, lattice=s.lattice,
                        neigh_sha=sha(nb), element=el1.astype(np.int8), charge=q.astype(np.int8), potential_boundary=pb,
                        potential_charge=pc)
    print("s_tile28k: N", s.N, "nn", s.nn, "charged", int((q != 0).sum()))
    s.close()


def step0():
    s = new_sim()
    x, y, z = s.positions()
    nb = s.neigh_idx()
    el = s.element()
    s.update_charge()
    q = s.charge()
    Vd = 1.5
    s.background_potential(Vd)
    s.poisson_gridless()
    pb, pc = s.potential_boundary(), s.potential_charge()
    et, ep = s.rate_table()
    nz = np.nonzero(et != 4)[0]
    np.savez_compressed(
        os.path.join(OUT, "s_step0.npz"),
        N=s.N, nn=s.nn, Vd=Vd, n_contact=s.num_atoms_contact, pbc=s.pbc,
        params=np.array([s.sigma, s.k, s.T_bg, s.freq, s.nn_dist, s.high_G, s.low_G]), lattice=s.lattice,
        metals=s.metals, layers=s.layers, site_layer=s.site_layer().astype(np.int8),
        neigh_sha=sha(nb), degree=(nb >= 0).sum(1).astype(np.int8),
        element=el.astype(np.int8), charge=q.astype(np.int8),
        potential_boundary=pb, potential_charge=pc,
        ev_idx=nz.astype(np.int32), ev_type=et[nz].astype(np.int8), ev_prob=ep[nz],
        kmc_rng_first16=s.peek_kmc_rng(16))
    print("step0 written", len(nz), "events")


def snapshot():
    import tempfile
    s = new_sim()
    s.update_charge()
    s.background_potential(1.5)
    s.poisson_gridless()
    g = np.load(os.path.join(OUT, "s_step0.npz"))
    assert np.array_equal(s.potential_boundary(), g["potential_boundary"]) and np.array_equal(s.potential_charge(), g["potential_charge"])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            os.mkdir("Results")
            s.write_snapshot("snapshot_0.xyz", "Results")
            raw = open(os.path.join(d, "Results", "snapshot_0.xyz"), "rb").read()
        finally:
            os.chdir(cwd)
    lines = raw.decode().split("\n")
    np.savez_compressed(os.path.join(OUT, "s_snapshot.npz"), sha256=hashlib.sha256(raw).hexdigest(), size=len(raw),
                        head=np.array(lines[:6]), tail=np.array(lines[-4:]), sample=np.array(lines[2 + 3000:2 + 3006]))
    print("snapshot written", len(raw), "bytes;", lines[2], "|", lines[3002])


def substoich():
    import re
    import tempfile
    base = open(REF + "parameters.txt").read()
    seeds, concs, shas, nvac = [], [], [], []
    for seed, conc in [(1, 0.02), (7, 0.1), (123, 0.05), (4, 0.2)]:
        txt, n1 = re.subn(r"rnd_seed = \d+", f"rnd_seed = {seed}", base)
        txt, n2 = re.subn(r"initial_vacancy_concentration = [0-9.]+", f"initial_vacancy_concentration = {conc}", txt)
        assert n1 == 1 and n2 == 1
        f = tempfile.NamedTemporaryFile("w", suffix="_parameters.txt", delete=False)
        f.write(txt); f.close()
        s = R.RefSim(f.name, REF + "reordered_device_2.5.xyz")
        assert s.rnd_seed == seed
        el = s.element()
        seeds.append(seed); concs.append(conc); shas.append(sha(el)); nvac.append(int((el == 2).sum()))
        print("substoich", seed, conc, nvac[-1], shas[-1][:12])
    np.savez_compressed(os.path.join(OUT, "s_substoich.npz"), seed=np.array(seeds), conc=np.array(concs),
                        sha=np.array(shas), n_vacancies=np.array(nvac))


def rates_ions():
    s = new_sim()
    el = s.element()
    x, _, _ = s.positions()
    cand = np.nonzero((el == 0) & (x > 3.0) & (x < 48.0))[0]          # DEFECT (empty interstitial) sites of the oxide
    pick = np.random.default_rng(11).choice(cand, 80, replace=False)
    el[pick] = 1                                                        # OXYGEN_DEFECT
    s.set_element(el)
    s.update_charge()
    Vd = 3.0
    s.background_potential(Vd)
    s.poisson_gridless()
    et, ep = s.rate_table()
    nz = np.nonzero(et != 4)[0]
    np.savez_compressed(os.path.join(OUT, "s_rates_ions.npz"), Vd=Vd, element=el.astype(np.int8),
                        charge=s.charge().astype(np.int8), potential_boundary=s.potential_boundary(),
                        potential_charge=s.potential_charge(), ev_idx=nz.astype(np.int32), ev_type=et[nz].astype(np.int8),
                        ev_prob=ep[nz])
    print("rates_ions written; types:", {int(t): int((et[nz] == t).sum()) for t in np.unique(et[nz])})


def cb_edge():
    s = new_sim()
    Vd = 1.5
    cb = s.laplace_cb_edge(Vd)
    np.savez_compressed(os.path.join(OUT, "s_cb_edge.npz"), Vd=Vd, n_contact=s.num_atoms_first_layer, cb_edge=cb)
    print("cb_edge written", cb[:2], cb[-2:])


def trajectory(name, schedule, nsteps, pbc=0):
    """schedule: list of (Vd, t) bias points as in kmc_main.cpp:136-279"""
    s = new_sim(pbc)
    assert s.pbc == pbc
    rec = dict(Vd=[], step_time=[], ev_ptr=[0], ev_ij=[], el_sha=[], q_sha=[], n_charged=[])
    count = 0
    for Vd, t in schedule:
        kmc_time = 0.0
        while kmc_time < t and count < nsteps:
            s.update_charge()
            s.background_potential(Vd)
            s.poisson_gridless()
            if count == 0:
                rec["pb0"], rec["pc0"] = s.potential_boundary(), s.potential_charge()
            rec["n_charged"].append(int(np.count_nonzero(s.charge())))
            dt, ev = s.kmc_step()
            kmc_time += dt
            rec["Vd"].append(Vd); rec["step_time"].append(dt)
            rec["ev_ij"].append(ev); rec["ev_ptr"].append(rec["ev_ptr"][-1] + len(ev))
            rec["el_sha"].append(sha(s.element())); rec["q_sha"].append(sha(s.charge()))
            count += 1
            print(name, "step", count, "Vd", Vd, "events", len(ev), "dt", dt, flush=True)
        if count >= nsteps:
            break
    rec["pbc"] = pbc
    rec["degree"] = (s.neigh_idx() >= 0).sum(1).astype(np.int8)
    rec["pb_last"], rec["pc_last"] = s.potential_boundary(), s.potential_charge()
    rec["element_last"] = s.element().astype(np.int8)
    rec["ev_ij"] = np.concatenate(rec["ev_ij"]).astype(np.int32) if rec["ev_ptr"][-1] else np.zeros((0, 2), np.int32)
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v) for k, v in rec.items()})


if __name__ == "__main__":
    import devicekmc_b200.host as H
    p = H.KMCParameters.from_file(REF + "parameters.txt")
    which = sys.argv[1:] or ["step0", "ramp", "6V", "10V", "6V_pbc", "cb_edge", "snapshot", "rates_ions", "substoich"]
    if "tile28k" in which:
        # ~25 minutes and ~7 GB: the reference's O(N^2) set-up, dense 27 333^2 K and dgesv at 28 k sites
        tile28k()
    if "step0" in which:
        step0()
    if "substoich" in which:
        substoich()
    if "rates_ions" in which:
        rates_ions()
    if "snapshot" in which:
        snapshot()
    if "cb_edge" in which:
        cb_edge()
    if "ramp" in which:
        trajectory("s_traj_ramp.npz", list(zip(p.V_switch, p.t_switch)), 12)
    if "ramp100" in which:
        # BASELINE.json config 1 in full: the first 100 KMC steps of the shipped ramp (~15 min of the reference CPU build)
        trajectory("s_traj_ramp100.npz", list(zip(p.V_switch, p.t_switch)), 100)
    if "6V" in which:
        trajectory("s_traj_6V.npz", [(6.0, 1.0)], 6)
    if "10V" in which:
        trajectory("s_traj_10V.npz", [(10.0, 1.0)], 8)
    if "6V_pbc" in which:
        trajectory("s_traj_6V_pbc.npz", [(6.0, 1.0)], 4, pbc=1)
