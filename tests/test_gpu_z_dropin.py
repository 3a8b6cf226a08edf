"""Drop-in check: the UNMODIFIED reference host program (kmc_main.cpp, Device, KMCProcess,
potential_solver, gpu_buffers — compiled with -DUSE_CUDA from /root/reference by oracle/Makefile
`dropin`) linked against OUR shim + libdkmc_b200.so instead of the reference's .cu files.
The binary (oracle/_ref/runKMC_b200) is built where the reference sources exist and travels to the
GPU box; here it runs config 0 and must reproduce the golden CPU-reference trajectory."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

BINARY = os.path.join(ROOT, "oracle", "_ref", "runKMC_b200")
# the same objects with Device::constructSiteNeighborList replaced at link time by the cell-list
# builder of devicekmc_b200/shim/device_setup_shim.cpp (make -C oracle dropin_fast)
BINARY_FAST = os.path.join(ROOT, "oracle", "_ref", "runKMC_b200_fastinit")


def _kmc_times(out_txt):
    return [float(v) for v in re.findall(r"KMC time is: (\S+)", out_txt)]


@pytest.mark.parametrize("BINARY", [BINARY, BINARY_FAST], ids=["reference_setup", "cell_list_setup"])
def test_reference_host_runs_on_our_library(tmp_path, BINARY):
    if not os.path.exists(BINARY):
        pytest.skip(f"{BINARY} not built (make -C oracle dropin dropin_fast, needs /root/reference)")
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import write_xyz
    el, x, y, z, lat, nc = S.load_base_cell()
    xyz = tmp_path / "device.xyz"
    write_xyz(str(xyz), el, x, y, z)
    params = open(os.path.join(GOLDEN, "parameters_2.5nm.txt")).read()
    params = re.sub(r"restart_xyz_file = \S+", f"restart_xyz_file = {xyz}", params)
    params = re.sub(r"solve_current = 1", "solve_current = 0", params)
    assert "solve_current = 0" in params
    (tmp_path / "parameters.txt").write_text(params)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([BINARY, "parameters.txt"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "devicekmc-b200:" not in r.stderr, r.stderr[-2000:]
    out = (tmp_path / "output.txt").read_text()
    times = _kmc_times(out)
    g = np.load(os.path.join(GOLDEN, "s_traj_ramp.npz"))
    # golden step times -> kmc_time as kmc_main accumulates it (reset at every bias point)
    want, acc, last_v = [], 0.0, None
    for v, dt in zip(g["Vd"], g["step_time"]):
        if v != last_v:
            acc, last_v = 0.0, v
        acc += dt
        want.append(acc)
    assert len(times) >= len(want)
    # output.txt carries 6 significant digits
    assert np.allclose(times[:len(want)], want, rtol=2e-5), (times[:len(want)], want)


def test_reference_host_on_100k_sites_with_cell_list_setup(tmp_path):
    """BASELINE.json config 1 scale (3 x 4 tiles, 112 788 sites, constant 6 V): the unmodified
    reference host program — its O(N^2) neighbour loop replaced at link time — on our
    library, against our own host mirror (devicekmc_b200.host) driving the same C-ABI: same
    structure, same seeds, so the same events and the same KMC times step for step."""
    if not os.path.exists(BINARY_FAST):
        pytest.skip("oracle/_ref/runKMC_b200_fastinit not built (make -C oracle dropin_fast)")
    import bench
    import devicekmc_b200 as D
    from devicekmc_b200.host import write_xyz
    el, x, y, z, lat, nc, p = bench.workload("tiled_100k")
    Vd, t_switch = 6.0, 3e-12
    xyz = tmp_path / "device.xyz"
    write_xyz(str(xyz), el, x, y, z)
    params = open(os.path.join(GOLDEN, "parameters_2.5nm.txt")).read()
    for pat, new in [(r"restart_xyz_file = \S+", f"restart_xyz_file = {xyz}"),
                     (r"solve_current = 1", "solve_current = 0"),
                     (r"lattice = [^/\n]+", "lattice = " + " ".join(repr(float(v)) for v in lat) + " "),
                     (r"V_switch = [^\n]+", f"V_switch = {Vd}"),
                     (r"t_switch = [^/\n]+", f"t_switch = {t_switch} "),
                     (r"num_atoms_first_layer = \d+", f"num_atoms_first_layer = {nc}"),
                     (r"num_atoms_contact = \d+", f"num_atoms_contact = {nc}")]:
        params, n = re.subn(pat, new, params)
        assert n == 1, pat
    (tmp_path / "parameters.txt").write_text(params)
    env = dict(os.environ, OMP_NUM_THREADS="8")
    r = subprocess.run([BINARY_FAST, "parameters.txt"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "devicekmc-b200:" not in r.stderr, r.stderr[-2000:]
    assert "Maximum number of neighbors in device is: 51" in r.stdout
    times = _kmc_times((tmp_path / "output.txt").read_text())
    assert 1 <= len(times) <= 200

    # the same loop (kmc_main.cpp:136-279) through the host mirror
    dev = D.Device([], p, arrays=(el, x, y, z))
    dev.makeSubstoichiometric(p.initial_vacancy_concentration)
    sim = D.KMCProcess(dev, p.freq)
    buf = D.GPUBuffers(sim.layers, sim.site_layer, sim.freq, dev, p.metals)
    buf.sync_HostToGPU(dev)
    want, kmc_time = [], 0.0
    while kmc_time < t_switch and len(want) < 200:
        dev.updateCharge(buf, p.metals)
        assert dev.updatePotential(buf, p, Vd)["cg_converged"]
        kmc_time += sim.executeKMCStep(buf, dev)
        want.append(kmc_time)
    assert len(times) == len(want), (times, want)
    assert np.allclose(times, want, rtol=2e-5), (times, want)       # output.txt carries 6 significant digits
