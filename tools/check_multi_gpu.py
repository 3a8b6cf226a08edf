"""torchrun --nproc-per-node N tools/check_multi_gpu.py [--workload tiled_100k]
Parity of the slab-partitioned path against the single-GPU path on the same inputs: every rank runs
both (the single-GPU path is replicated), integers must be identical, potentials within 1e-10."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import devicekmc_b200 as D  # noqa: E402
from devicekmc_b200 import slab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="tiled_100k")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--vd", type=float, default=10.0)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
el, x, y, z, lat, nc, p = bench.workload(args.workload)
el = bench.substoichiometric(el, p)
multi = slab.SlabSim((el, x, y, z), p, rank, world, distributed_cg=True)
single = slab.SlabSim((el, x, y, z), p, 0, 1, distributed_cg=False)
ok = True
for s in range(args.steps):
    a = multi.step(args.vd)
    b = single.step(args.vd)
    pb_m, pb_s = multi.buf.site_potential_boundary.cpu().numpy(), single.buf.site_potential_boundary.cpu().numpy()
    pc_m, pc_s = multi.buf.site_potential_charge.cpu().numpy(), single.buf.site_potential_charge.cpu().numpy()
    e_b = np.abs(pb_m - pb_s).max() / np.abs(pb_s).max()
    e_c = np.abs(pc_m - pc_s).max() / np.abs(pc_s).max()
    same_el = bool(torch.equal(multi.buf.site_element, single.buf.site_element))
    same_q = bool(torch.equal(multi.buf.site_charge, single.buf.site_charge))
    good = e_b <= 1e-10 and e_c <= 1e-10 and same_el and same_q and a["events"] == b["events"]
    ok &= good
    print(f"rank {rank} step {s}: phi_b rel {e_b:.2e} phi_c rel {e_c:.2e} element== {same_el} charge== {same_q} "
          f"events {a['events']}/{b['events']} cg its {a['cg_iterations']}/{b['cg_iterations']} "
          f"solve ms {a['solve_ms']:.2f}/{b['solve_ms']:.2f} {'OK' if good else 'MISMATCH'}", flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU PARITY", "OK" if t.item() == 1.0 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
