"""Walk cost vs target subset and block size (scratch)."""
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
subsets = {"all_self": None}
for bs in (2048, 16384, 65536, 262144):
    blk = np.arange(n) // bs
    subsets["alt8_block%d" % bs] = np.nonzero(blk % 8 == 3)[0]
J.tree_stats(True)
for name, sel in subsets.items():
    if sel is None:
        J.tree_force(tx, tm, 0.05, 0.7, precision="fp32")
    else:
        tt = torch.from_numpy(np.ascontiguousarray(xs[sel])).cuda()
        J.tree_force_position(tx, tm, tt, 0.05, 0.7, precision="fp32")
    torch.cuda.synchronize()
    st = J.tree_stats()
    print("STATS", name, "warps", st["warps"], "mean warp entries %.0f" % (st["warp_entries"] / st["warps"]), "max", st["warp_entries_max"],
          "visited/target %.0f accepted/target %.0f" % (st["visited"] / (st["warps"] * 32), st["accepted"] / (st["warps"] * 32)), flush=True)
J.tree_stats(False)
for wb in (0,):
    for name, sel in subsets.items():
        for rep in range(2):
            if sel is None:
                J.tree_force(tx, tm, 0.05, 0.7, precision="fp32")
            else:
                tt = torch.from_numpy(np.ascontiguousarray(xs[sel])).cuda()
                J.tree_force_position(tx, tm, tt, 0.05, 0.7, precision="fp32")
        torch.cuda.synchronize()
        print("done", wb, name, flush=True)
