"""Parity on a crafted state with oxygen ions (tests/golden/s_rates_ions.npz, written by the
reference's CPU build): recombination and ion-diffusion entries in the rate table, negative and
positive charges in the pairwise sum, and the event loop executing on such a table.
(Sorted last on purpose: added at the very end of round 1, when the GPU budget of the round was spent;
the same inputs pass bit for bit in the oracle, tests/test_cpu_oracle_and_host.py.)"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_rate_table_charges_and_events_with_oxygen_ions(base_case):
    import torch
    import devicekmc_b200 as D
    from devicekmc_b200._capi import check
    from oracle import oracle as O
    g = np.load(os.path.join(GOLDEN, "s_rates_ions.npz"))
    p = base_case["p"]
    el = g["element"].astype(np.int32)
    dev = D.Device([], p, arrays=(el, base_case["x"], base_case["y"], base_case["z"]))
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    # a3: charges with ions present
    dev.updateCharge(buf, p.metals)
    q = buf.site_charge.cpu().numpy()
    assert np.array_equal(q, g["charge"].astype(np.int32))
    # a6: pairwise sum of positive and negative charges (device potential; the boundary potential comes from
    # the fixture so that the rate table sees exactly the reference's inputs)
    dev.updatePotential(buf, p, float(g["Vd"]), n_contact=p.num_atoms_contact)
    pc = buf.site_potential_charge.cpu().numpy()
    ref = g["potential_charge"]
    nz = ref != 0
    assert (np.abs(pc[nz] - ref[nz]) <= TOL * np.abs(ref[nz])).all() and np.array_equal(pc[~nz], ref[~nz])
    pb = buf.site_potential_boundary.cpu().numpy()
    assert np.abs(pb - g["potential_boundary"]).max() <= 5e-9 * np.abs(pb).max()     # reference dgesv accuracy
    buf.site_potential_boundary.copy_(torch.from_numpy(g["potential_boundary"]))
    buf.site_potential_charge.copy_(torch.from_numpy(g["potential_charge"]))
    # a7: all four event types
    n = dev.N * buf.nn_
    et = torch.empty(n, dtype=torch.int32, device="cuda"); ep = torch.empty(n, dtype=torch.float64, device="cuda")
    check(dev.ctx.lib.dkmc_build_event_list(
        dev.ctx.h, dev.N, buf.nn_, buf.neigh_idx.data_ptr(), buf.site_layer.data_ptr(), buf.lattice.data_ptr(), dev.pbc,
        buf.T_bg.data_ptr(), buf.freq.data_ptr(), buf.sigma.data_ptr(), buf.k.data_ptr(), buf.site_x.data_ptr(),
        buf.site_y.data_ptr(), buf.site_z.data_ptr(), buf.site_potential_boundary.data_ptr(),
        buf.site_potential_charge.data_ptr(), buf.site_element.data_ptr(), buf.site_charge.data_ptr(), et.data_ptr(),
        ep.data_ptr()))
    et, ep = et.cpu().numpy(), ep.cpu().numpy()
    g_type = np.full(n, 4, np.int32); g_prob = np.zeros(n)
    g_type[g["ev_idx"]] = g["ev_type"]; g_prob[g["ev_idx"]] = g["ev_prob"]
    assert set(np.unique(g["ev_type"]).tolist()) == {0, 1, 2, 3}
    assert np.array_equal(et, g_type)                                       # integer: bit-exact
    nzp = g_prob != 0
    assert np.array_equal(ep[~nzp], g_prob[~nzp])
    assert (np.abs(ep[nzp] - g_prob[nzp]) <= TOL * g_prob[nzp]).all()
    # a8: the event loop on this table against the oracle on the same (device-built) table
    nb = dev.neigh_idx.reshape(dev.N, -1)
    rng = O.Rng(D.host.RND_SEED_KMC)
    t_ref, ev_ref, el_ref, q_ref = O.kmc_events(nb, et, ep, el, q, sim.freq, rng)
    t = sim.executeKMCStep(buf, dev, record_events=4096)
    assert np.array_equal(sim.last_events, ev_ref)
    assert np.array_equal(buf.site_element.cpu().numpy(), el_ref) and np.array_equal(buf.site_charge.cpu().numpy(), q_ref)
    assert abs(t - t_ref) <= 1e-12 * abs(t_ref)


def test_cpp_host_example_reproduces_the_reference_trajectory(base_case, tmp_path):
    """examples/kmc_loop.cpp — a C++ host over the C-ABI with std::mt19937 as the reference uses it — on the
    shipped device at constant 6 V: the KMC times of the reference's CPU run (s_traj_6V.npz), step for step"""
    import re
    import subprocess
    from conftest import build_cpp_example
    from devicekmc_b200.host import write_xyz
    g = np.load(os.path.join(GOLDEN, "s_traj_6V.npz"))
    exe = build_cpp_example(tmp_path)
    xyz = tmp_path / "device.xyz"
    write_xyz(str(xyz), base_case["element"], base_case["x"], base_case["y"], base_case["z"])
    r = subprocess.run([exe, str(xyz), "108.97557", "25.575", "25.575", "144", "6.0", str(len(g["Vd"]))],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    times = [float(v) for v in re.findall(r"KMC time is: (\S+)", r.stdout)]
    events = [int(v) for v in re.findall(r"(\d+) events", r.stdout)]
    assert events == np.diff(g["ev_ptr"]).tolist()
    assert np.allclose(times, np.cumsum(g["step_time"]), rtol=2e-5)          # printed with 6 significant digits
