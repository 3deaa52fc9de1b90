"""Short driver for ncu captures: one direct fp32, one direct fp64 and one tree evaluation with
device-resident inputs (scratch tool; numbers printed under a profiler are never bench values)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gravhopper_b200 import _jbgrav as J, ic_raw

which = sys.argv[1] if len(sys.argv) > 1 else "direct32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
if which.startswith("tree"):
    x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42); eps = 0.05
else:
    x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=42); eps = 5e-5
tx = torch.from_numpy(np.ascontiguousarray(x)).cuda(); tm = torch.from_numpy(m).cuda()
for _ in range(3):
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
