import os, re, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd())
from devicekmc_b200 import structures as S
from devicekmc_b200.host import write_xyz
d = tempfile.mkdtemp()
el, x, y, z, lat, nc = S.load_base_cell()
write_xyz(d + "/device.xyz", el, x, y, z)
params = open("tests/golden/parameters_2.5nm.txt").read()
params = re.sub(r"restart_xyz_file = \S+", f"restart_xyz_file = {d}/device.xyz", params)
params = re.sub(r"solve_current = 1", "solve_current = 0", params)
open(d + "/parameters.txt", "w").write(params)
try:
    r = subprocess.run(["/root/repo/oracle/_ref/runKMC_b200", "parameters.txt"], cwd=d, capture_output=True, text=True, timeout=60, env=dict(os.environ, OMP_NUM_THREADS="4"))
    print("rc", r.returncode); print(r.stdout[-3000:]); print("ERR", r.stderr[-2000:])
except subprocess.TimeoutExpired as e:
    print("TIMEOUT"); print((e.stdout or b"")[-3000:]); print("ERR", (e.stderr or b"")[-2000:])
print(os.listdir(d))
if os.path.exists(d + "/output.txt"): print(open(d + "/output.txt").read()[-2500:])
