"""Short driver for ncu captures: one direct fp32, one direct fp64 and one tree evaluation with
device-resident inputs, or a few steps of the device-resident engine (scratch tool; numbers printed
under a profiler are never bench values).
usage: profile_kernels.py direct32|direct64|tree32|tree64|engine_tree32|engine_direct32 [N] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gravhopper_b200 import _jbgrav as J, ic_raw

which = sys.argv[1] if len(sys.argv) > 1 else "direct32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if "tree" in which:
    x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42); eps, dt = 0.05, 1.0
else:
    x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=42); eps, dt = 5e-5, 0.005
if which.startswith("engine"):
    # the path bench.py times: state resident, fused epilogue, splitter sort from the second step on
    from gravhopper_b200.sharded import ShardedSimulation
    sim = ShardedSimulation(x, v, m, dt, eps, algorithm="tree" if "tree" in which else "direct", theta=0.7,
                            precision="fp32" if which.endswith("32") else "fp64")
    sim.run(steps)
    sim.shard.synchronize()
    print("done", which, n, steps)
    sys.exit(0)
tx = torch.from_numpy(np.ascontiguousarray(x)).cuda(); tm = torch.from_numpy(m).cuda()
for _ in range(steps):
    if which == "direct32":
        J.direct_summation(tx, tm, eps, precision="fp32")
    elif which == "direct64":
        J.direct_summation(tx, tm, eps, precision="fp64")
    elif which == "tree32":
        J.tree_force(tx, tm, eps, 0.7, precision="fp32")
    else:
        J.tree_force(tx, tm, eps, 0.7, precision="fp64")
torch.cuda.synchronize()
print("done", which, n)
