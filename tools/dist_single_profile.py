"""dev aid: the slab-partitioned PCG code path with world = 1 (all rows on one rank), so that ncu can
list its kernels:  ncu --metrics gpu__time_duration.sum --profile-from-start off --csv python tools/dist_single_profile.py"""
import os, sys
sys.path.insert(0, os.getcwd())
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29555")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
import torch, torch.distributed as dist
import bench
from devicekmc_b200 import slab, _dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
el, x, y, z, lat, nc, p = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "tiled_1M"); el = bench.substoichiometric(el, p)
s = slab.SlabSim((el, x, y, z), p, 0, 1, distributed_cg=False)
s.dcg = _dist.DistributedSolver(s)
for _ in range(3): out = s.step(10.0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = s.step(10.0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(out)
dist.destroy_process_group()
