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


def test_reference_host_runs_on_our_library(tmp_path):
    if not os.path.exists(BINARY):
        pytest.skip("oracle/_ref/runKMC_b200 not built (make -C oracle dropin, needs /root/reference)")
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
    times = [float(v) for v in re.findall(r"KMC time is: (\S+)", out)]
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
