"""Times the C++ drop-in: the UNMODIFIED reference host program (kmc_main.cpp, Device, KMCProcess, GPUBuffers — with
the O(N^2) neighbour loop replaced at link time) linked against this library
(oracle/_ref/runKMC_b200_fastinit, built where /root/reference exists).  The reference host logs its own wall time
per KMC superstep into output.txt ("Z - calculation time - KMC superstep [s]"); this tool runs it on a synthetic
tile device at constant voltage and prints the median beside the step count.
    python tools/time_dropin.py tiled_1M [Vd] [t_switch]
Numbers of this tool are host wall clock of the reference's own loop (its logging included), not bench values."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from devicekmc_b200.host import write_xyz  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tiled_100k"
Vd = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
t_switch = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5e-13
BINARY = os.path.join(ROOT, "oracle", "_ref", "runKMC_b200_fastinit")
GOLDEN = os.path.join(ROOT, "tests", "golden")
el, x, y, z, lat, nc, p = bench.workload(name)
d = tempfile.mkdtemp()
xyz = os.path.join(d, "device.xyz")
t0 = time.time(); write_xyz(xyz, el, x, y, z); t_write = time.time() - t0
params = open(os.path.join(GOLDEN, "parameters_2.5nm.txt")).read()
for pat, new in [(r"restart_xyz_file = \S+", f"restart_xyz_file = {xyz}"),
                 (r"solve_current = 1", "solve_current = 0"),
                 (r"lattice = [^/\n]+", "lattice = " + " ".join(repr(float(v)) for v in lat) + " "),
                 (r"V_switch = [^\n]+", f"V_switch = {Vd}"),
                 (r"t_switch = [^/\n]+", f"t_switch = {t_switch} "),
                 (r"log_freq = \d+", "log_freq = 100000"),
                 (r"num_atoms_first_layer = \d+", f"num_atoms_first_layer = {nc}"),
                 (r"num_atoms_contact = \d+", f"num_atoms_contact = {nc}")]:
    params, n = re.subn(pat, new, params)
    assert n == 1, pat
open(os.path.join(d, "parameters.txt"), "w").write(params)
t0 = time.time()
r = subprocess.run([BINARY, "parameters.txt"], cwd=d, env=dict(os.environ, OMP_NUM_THREADS="8"), capture_output=True, text=True,
                   timeout=900)
wall = time.time() - t0
assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
out = open(os.path.join(d, "output.txt")).read()
steps = [float(v) for v in re.findall(r"KMC superstep \[s\]: (\S+)", out)]
pot = [float(v) for v in re.findall(r"potential from boundaries \[s\]: (\S+)", out)]
ev = [float(v) for v in re.findall(r"kmc events \[s\]: (\S+)", out)]
warm = steps[3:] if len(steps) > 6 else steps
print(json.dumps({"tool": "time_dropin", "workload": name, "sites": int(len(x)), "Vd": Vd, "kmc_steps": len(steps),
                  "superstep_ms_all": [round(1e3 * s, 2) for s in steps[:24]],
                  "superstep_ms_median_after_3": round(1e3 * float(np.median(warm)), 2) if warm else None,
                  "steps_per_s_median_after_3": round(1.0 / float(np.median(warm)), 2) if warm else None,
                  "potential_boundary_ms_median": round(1e3 * float(np.median(pot)), 2) if pot else None,
                  "kmc_events_ms_median": round(1e3 * float(np.median(ev)), 2) if ev else None,
                  "process_wall_s": round(wall, 1), "xyz_write_s": round(t_write, 1),
                  "host": "unmodified reference C++ host (kmc_main.cpp loop, its per-step logging included) on libdkmc_b200.so"}))
