"""torchrun --nproc-per-node N tools/check_multi_gpu.py [--workload tiled_100k] [--oracle]
Parity of the slab-partitioned path.  Every rank runs, on the same inputs,
  * the distributed path (rows / targets / rate-table slabs over the ranks), and
  * the single-GPU path (replicated);
integers (elements, charges, executed events) must be identical, potentials within 1e-10.  With --oracle rank 0
also checks the distributed potentials of step 0 against the CPU oracle: phi_b on the whole device, phi_c on
bands of target rows (test infrastructure: the oracle is never part of the path).  Used by tests/test_gpu_multi.py."""
import argparse
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from devicekmc_b200 import slab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="tiled_100k")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--vd", type=float, default=None)
ap.add_argument("--oracle", action="store_true")
args = ap.parse_args()
args.vd = bench.workload_vd(args.workload, args.vd)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
el, x, y, z, lat, nc, p = bench.workload(args.workload)
el = bench.substoichiometric(el, p)
multi = slab.SlabSim((el, x, y, z), p, rank, world, distributed_cg=True)
single = slab.SlabSim((el, x, y, z), p, 0, 1, distributed_cg=False)
TOL = 1e-10
ok = True
for s in range(args.steps):
    a = multi.step(args.vd, record_events=1 << 16)
    ev_m = multi.sim.last_events.copy()
    pb_m = multi.buf.site_potential_boundary.cpu().numpy(); pc_m = multi.buf.site_potential_charge.cpu().numpy()
    b = single.step(args.vd, record_events=1 << 16)
    ev_s = single.sim.last_events.copy()
    pb_s = single.buf.site_potential_boundary.cpu().numpy(); pc_s = single.buf.site_potential_charge.cpu().numpy()
    e_b = np.abs(pb_m - pb_s).max() / np.abs(pb_s).max()
    e_c = np.abs(pc_m - pc_s).max() / np.abs(pc_s).max()
    same_el = bool(torch.equal(multi.buf.site_element, single.buf.site_element))
    same_q = bool(torch.equal(multi.buf.site_charge, single.buf.site_charge))
    same_ev = a["events"] == b["events"] and np.array_equal(ev_m, ev_s)
    good = e_b <= TOL and e_c <= TOL and same_el and same_q and same_ev and a["cg_converged"]
    if args.oracle and s == 0 and rank == 0:
        from oracle import oracle as O   # the checker
        dev = multi.dev
        nb = dev.neigh_idx.reshape(dev.N, -1)
        q0 = O.update_charge(nb, el, p.metals, np.zeros(dev.N, np.int32))
        pb_o, _ = O.background_potential(nb, nc, nc, el, q0, p.metals, p.high_G, p.low_G, args.vd)
        o_b = np.abs(pb_m - pb_o).max() / np.abs(pb_o).max()
        o_c = 0.0
        for r0 in (0, dev.N // 2 - 256, dev.N - 512):
            pc_o = O.poisson_gridless(x, y, z, lat, p.pbc, q0, p.sigma, p.k, rows=(r0, r0 + 512))
            o_c = max(o_c, float(np.abs(pc_m[r0:r0 + 512] - pc_o).max() / np.abs(pc_o).max()))
        good = good and o_b <= TOL and o_c <= TOL
        print(f"rank 0 step 0 vs ORACLE: phi_b rel {o_b:.2e} phi_c (3 bands) rel {o_c:.2e} {'OK' if o_b <= TOL and o_c <= TOL else 'MISMATCH'}", flush=True)
    if args.oracle and s == 0:
        dist.barrier()      # rank 0 spent seconds in the CPU oracle: the others wait here, not inside a kernel
    ok &= good
    print(f"rank {rank} step {s}: phi_b rel {e_b:.2e} phi_c rel {e_c:.2e} element== {same_el} charge== {same_q} "
          f"events {a['events']}/{b['events']} identical {same_ev} cg its {a['cg_iterations']}/{b['cg_iterations']} "
          f"{'OK' if good else 'MISMATCH'}", flush=True)
# every rank must hold the same state, bit for bit
h = hashlib.sha256(multi.buf.site_element.cpu().numpy().tobytes() + multi.buf.site_charge.cpu().numpy().tobytes()).hexdigest()
hs = [None] * world
dist.all_gather_object(hs, h)
ok &= len(set(hs)) == 1
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ranks hold identical element/charge arrays:", len(set(hs)) == 1, flush=True)
    print("MULTI-GPU PARITY", "OK" if t.item() == 1.0 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
