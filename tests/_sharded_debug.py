"""torchrun debug worker (scratch): per-step difference between a sharded run and the single-rank run."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravhopper_b200 import ic_raw
from gravhopper_b200.sharded import ShardedSimulation

n, steps = int(sys.argv[1]), int(sys.argv[2])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=21)
sim = ShardedSimulation(x, v, m, 0.005, 5e-5, algorithm="tree", precision="fp32", rank=rank, world=world, device=local)
single = None
if rank == 0:
    single = ShardedSimulation(x, v, m, 0.005, 5e-5, algorithm="tree", precision="fp32")
for s in range(1, steps + 1):
    sim.step()
    pos, vel = sim.gather_state()
    if rank == 0:
        single.step()
        p1, v1 = single.gather_state()
        dp = np.abs(pos - p1).max(axis=1) / np.abs(p1).max()
        dv = np.abs(vel - v1).max(axis=1) / np.abs(v1).max()
        print("step %d pos median %.2e p99 %.2e max %.2e | vel median %.2e p99 %.2e max %.2e" %
              (s, np.median(dp), np.percentile(dp, 99), dp.max(), np.median(dv), np.percentile(dv, 99), dv.max()), flush=True)
dist.barrier()
dist.destroy_process_group()
