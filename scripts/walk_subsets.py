"""Walk cost vs target subset (scratch): all targets, contiguous Morton half, alternate Morton blocks, random half."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gravhopper_b200 import _jbgrav as J, ic_raw
from gravhopper_b200.sharded import morton_order
n = 1 << 22
x, v, m = ic_raw.Hernquist(n, 1.0, 1e10, seed=42)
x = np.ascontiguousarray(x)
order = morton_order(x)
xs = np.ascontiguousarray(x[order]); ms = m[order]
tx, tm = torch.from_numpy(xs).cuda(), torch.from_numpy(ms).cuda()
blocks = np.arange(n) // 2048
subsets = {"all_self": None, "contig_half": np.arange(n // 2), "alt_blocks": np.nonzero(blocks % 2 == 0)[0],
           "random_half": np.sort(np.random.default_rng(0).choice(n, n // 2, replace=False)),
           "alt_blocks_8": np.nonzero(blocks % 8 == 0)[0]}
for name, sel in subsets.items():
    for rep in range(2):
        if sel is None:
            J.tree_force(tx, tm, 0.05, 0.7, precision="fp32")
        else:
            tt = torch.from_numpy(np.ascontiguousarray(xs[sel])).cuda()
            J.tree_force_position(tx, tm, tt, 0.05, 0.7, precision="fp32")
    torch.cuda.synchronize()
    print("done", name, flush=True)
