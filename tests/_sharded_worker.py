"""torchrun worker for tests/test_gpu_multi.py: run a sharded leapfrog and dump rank 0's gathered state."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravhopper_b200 import ic_raw  # noqa: E402
from gravhopper_b200.sharded import ShardedSimulation  # noqa: E402

out, n, alg, prec, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=21)
sim = ShardedSimulation(x, v, m, 0.005, 5e-5, algorithm=alg, precision=prec, rank=rank, world=world, device=local)
sim.run(steps)
pos, vel = sim.gather_state()
if rank == 0:
    np.savez(out, pos=pos, vel=vel)
dist.barrier()
dist.destroy_process_group()
